/*
 * Closed forms of the block contents, shared by device kernels and host code.
 *
 * Pattern: LocalWorker::preWriteIntegrityCheckFillBuf (reference source/workers/LocalWorker.cpp:
 * 2091-2128): file byte x = byte (x % 8) of the little-endian u64 ((x & ~7) + salt), wrap-around
 * mod 2^64.
 *
 * Random fill: counter-based replacement of preWriteBufRandRefill/-Cuda (:2209-2277). The
 * reference draws from a self-seeded serial PRNG (unpinned content); here every u64 word is a
 * pure function of (seed, blockCounter, word index) so any thread can produce any word.
 *   blockKey    = splitmix64_mix(seed + blockCounter * ELB_CTR_MULT)
 *   word k      = splitmix64_mix(blockKey + (k+1) * GOLDEN)   (k = byte position / 8 in block)
 *   remainder v = splitmix64_mix(blockKey)                    (the one repeated u64, :2228)
 * i.e. word k is output k of a standard SplitMix64 stream seeded with blockKey.
 */
#ifndef ELB_PATTERNS_CUH_
#define ELB_PATTERNS_CUH_

#include <stdint.h>

#if defined(__CUDACC__)
#define ELB_HD __host__ __device__ __forceinline__
#else
#define ELB_HD static inline
#endif

#define ELB_GOLDEN 0x9E3779B97F4A7C15ULL
#define ELB_CTR_MULT 0xD1342543DE82EF95ULL

ELB_HD uint64_t elb_splitmix64_mix(uint64_t z)
{
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

ELB_HD uint64_t elb_rand_block_key(uint64_t seed, uint64_t blockCounter)
{
	return elb_splitmix64_mix(seed + blockCounter * ELB_CTR_MULT);
}

ELB_HD uint64_t elb_rand_word(uint64_t blockKey, uint64_t wordIdx)
{
	return elb_splitmix64_mix(blockKey + (wordIdx + 1) * ELB_GOLDEN);
}

ELB_HD uint64_t elb_rand_remainder_val(uint64_t blockKey)
{
	return elb_splitmix64_mix(blockKey);
}

/* varFillLen of the GPU path: (len*pct)/100 rounded down to a multiple of 4
 * (LocalWorker.cpp:2251-2256) */
ELB_HD uint64_t elb_rand_var_fill_len(uint64_t len, unsigned pct)
{
	uint64_t varFillLen = (len * pct) / 100;
	return varFillLen - (varFillLen % 4);
}

/* 8 pattern bytes starting at (arbitrary) file position filePos, as a little-endian u64. */
ELB_HD uint64_t elb_pattern_bytes8(uint64_t filePos, uint64_t salt)
{
	const unsigned shiftBits = (unsigned)(filePos & 7) * 8;
	const uint64_t w0 = (filePos & ~7ULL) + salt;

	if(!shiftBits)
		return w0;

	const uint64_t w1 = w0 + 8;

	return (w0 >> shiftBits) | (w1 << (64 - shiftBits) );
}

/* single pattern byte at file position filePos */
ELB_HD uint8_t elb_pattern_byte(uint64_t filePos, uint64_t salt)
{
	const uint64_t w0 = (filePos & ~7ULL) + salt;
	return (uint8_t)(w0 >> ( (filePos & 7) * 8) );
}

/* single random-fill byte at block-relative position pos */
ELB_HD uint8_t elb_rand_byte(uint64_t pos, uint64_t blockKey, uint64_t varFillLen,
	uint64_t remainderVal)
{
	if(pos < varFillLen)
		return (uint8_t)(elb_rand_word(blockKey, pos >> 3) >> ( (pos & 7) * 8) );

	return (uint8_t)(remainderVal >> ( ( (pos - varFillLen) & 7) * 8) );
}

/* 8 random-fill bytes starting at block-relative position pos, as a little-endian u64 */
ELB_HD uint64_t elb_rand_bytes8(uint64_t pos, uint64_t blockKey, uint64_t varFillLen,
	uint64_t remainderVal)
{
	if( !(pos & 7) && ( (pos + 8) <= varFillLen) )
		return elb_rand_word(blockKey, pos >> 3); // fast path: aligned word inside var part

	if(pos >= varFillLen)
	{ // inside the constant remainder: rotate the repeated u64
		const unsigned rotBits = (unsigned)( (pos - varFillLen) & 7) * 8;
		return rotBits ?
			( (remainderVal >> rotBits) | (remainderVal << (64 - rotBits) ) ) : remainderVal;
	}

	// straddles the boundary or unaligned: compose byte-wise (rare path: keep its code small)
	uint64_t val = 0;

#if defined(__CUDA_ARCH__)
	#pragma unroll 1
#endif
	for(unsigned i = 0; i < 8; i++)
		val |= (uint64_t)elb_rand_byte(pos + i, blockKey, varFillLen, remainderVal) << (i * 8);

	return val;
}

#endif /* ELB_PATTERNS_CUH_ */
