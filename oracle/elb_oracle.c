/*
 * TEST INFRASTRUCTURE ONLY — CPU restatement of elbencho's LocalWorker hot path in plain C.
 * (See elb_oracle.h for who may use this and how its parity is pinned.)
 *
 * Every function cites the reference lines it follows; paths are relative to the reference's
 * source/ directory. This is a restatement of behaviour, not a copy: the reference is C++ with
 * member-function pointers, std::chrono and exceptions, this is C with explicit state.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <linux/aio_abi.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/syscall.h>
#include <sys/types.h>
#include <time.h>
#include <unistd.h>

#include "elb_oracle.h"

#define ORC_MIN(a, b) ( ( (a) < (b) ) ? (a) : (b) )
#define ORC_MKFILE_MODE (S_IRUSR | S_IWUSR | S_IRGRP | S_IWGRP | S_IROTH | S_IWOTH)
#define ORC_MKDIR_MODE (S_IRWXU | S_IRWXG | S_IRWXO)
#define ORC_PATH_BUF_LEN 64 /* workers/LocalWorker.cpp:62 */
#define ORC_AIO_MAX_WAIT_SEC 5 /* workers/LocalWorker.cpp:60 */
#define ORC_AIO_MAX_EVENTS 4   /* workers/LocalWorker.cpp:61 */

/* ============================================================================================
 * Block modifiers / checkers
 * ========================================================================================== */

/* workers/LocalWorker.cpp:2091-2128 (preWriteIntegrityCheckFillBuf): walk the buffer in pieces
 * that never cross an 8-byte file-offset boundary; each piece is the matching slice of the
 * little-endian u64 (alignedOffset + salt). */
void orc_fill_pattern(char* buf, size_t bufLen, uint64_t fileOffset, uint64_t salt)
{
	const size_t checkSumLen = sizeof(uint64_t);
	size_t numBytesDone = 0;
	size_t numBytesLeft = bufLen;
	uint64_t currentOffset = fileOffset;

	while(numBytesLeft)
	{
		uint64_t checkSumStartOffset = currentOffset - (currentOffset % checkSumLen);
		uint64_t checkSum = checkSumStartOffset + salt;
		const char* checkSumArray = (const char*)&checkSum;
		uint64_t checkSumArrayStartIdx = currentOffset - checkSumStartOffset;
		size_t checkSumCopyLen = ORC_MIN( (uint64_t)numBytesLeft,
			(uint64_t)checkSumLen - checkSumArrayStartIdx);

		memcpy(&buf[numBytesDone], &checkSumArray[checkSumArrayStartIdx], checkSumCopyLen);

		numBytesDone += checkSumCopyLen;
		numBytesLeft -= checkSumCopyLen;
		currentOffset += checkSumCopyLen;
	}
}

/* workers/LocalWorker.cpp:2137-2179 (postReadIntegrityCheckVerifyBuf): malloc a scratch block,
 * regenerate the expected bytes, memcmp; on difference scan for the first bad byte and produce
 * the reference's exception text. Additionally counts all differing bytes (the reference stops
 * at the first; the count is what the GPU kernel's warp-reduced counter is compared against). */
int orc_verify_pattern(const char* buf, size_t bufLen, uint64_t fileOffset, uint64_t salt,
	uint64_t* outNumMismatchBytes, uint64_t* outFirstMismatchIdx, unsigned* outExpected,
	unsigned* outActual, char* errBuf, size_t errBufLen)
{
	if(outNumMismatchBytes)
		*outNumMismatchBytes = 0;
	if(outFirstMismatchIdx)
		*outFirstMismatchIdx = ~0ULL;
	if(errBuf && errBufLen)
		errBuf[0] = 0;

	if(!bufLen)
		return 0; // :2140-2141

	char* verifyBuf = (char*)malloc(bufLen);
	if(!verifyBuf)
	{
		if(errBuf)
			snprintf(errBuf, errBufLen, "Buffer alloc for verification buffer failed. Size: %zu",
				bufLen);
		return -1;
	}

	orc_fill_pattern(verifyBuf, bufLen, fileOffset, salt);

	int compareRes = memcmp(buf, verifyBuf, bufLen);

	if(!compareRes)
	{
		free(verifyBuf);
		return 0;
	}

	uint64_t numMismatch = 0;
	uint64_t firstIdx = ~0ULL;

	for(size_t i = 0; i < bufLen; i++)
	{
		if(verifyBuf[i] == buf[i])
			continue;

		if(firstIdx == ~0ULL)
		{
			firstIdx = i;

			unsigned expectedVal = (unsigned char)verifyBuf[i];
			unsigned actualVal = (unsigned char)buf[i];

			if(outExpected)
				*outExpected = expectedVal;
			if(outActual)
				*outActual = actualVal;

			if(errBuf) // text of :2174-2177
				snprintf(errBuf, errBufLen, "Data verification failed. Offset: %llu; "
					"Expected value: %u; Actual value: %u",
					(unsigned long long)(fileOffset + i), expectedVal, actualVal);
		}

		numMismatch++;
	}

	if(outNumMismatchBytes)
		*outNumMismatchBytes = numMismatch;
	if(outFirstMismatchIdx)
		*outFirstMismatchIdx = firstIdx;

	free(verifyBuf);

	return 1;
}

/* workers/LocalWorker.cpp:2185-2203 (bufFill) */
void orc_buf_fill(char* buf, uint64_t fillValue, size_t bufLen)
{
	size_t numBytesDone = 0;

	for(uint64_t i = 0; i < (bufLen / sizeof(uint64_t) ); i++)
	{
		memcpy(buf, &fillValue, sizeof(uint64_t) );
		buf += sizeof(uint64_t);
		numBytesDone += sizeof(uint64_t);
	}

	if(numBytesDone == bufLen)
		return;

	memcpy(buf, &fillValue, bufLen - numBytesDone);
}

/* ============================================================================================
 * PRNGs
 * ========================================================================================== */

static inline uint64_t orc_rol64(uint64_t x, int k)
{
	return (x << k) | (x >> (64 - k) );
}

/* toolkits/random/RandAlgoXoshiro256ss.h:76-91 */
uint64_t orc_xoshiro256ss_next(orc_xoshiro256ss* st)
{
	uint64_t* s = st->s;
	uint64_t const result = orc_rol64(s[1] * 5, 7) * 9;
	uint64_t const t = s[1] << 17;

	s[2] ^= s[0];
	s[3] ^= s[1];
	s[1] ^= s[2];
	s[0] ^= s[3];

	s[2] ^= t;
	s[3] = orc_rol64(s[3], 45);

	return result;
}

/* toolkits/random/RandAlgoXoshiro256ss.h:46-65 */
void orc_xoshiro256ss_fill_buf(orc_xoshiro256ss* st, char* buf, uint64_t bufLen)
{
	uint64_t numBytesDone = 0;

	for(uint64_t i = 0; i < (bufLen / sizeof(uint64_t) ); i++)
	{
		uint64_t val = orc_xoshiro256ss_next(st);
		memcpy(buf, &val, sizeof(val) );
		buf += sizeof(uint64_t);
		numBytesDone += sizeof(uint64_t);
	}

	if(numBytesDone == bufLen)
		return;

	uint64_t randUint64 = orc_xoshiro256ss_next(st);
	memcpy(buf, &randUint64, bufLen - numBytesDone);
}

/* toolkits/random/RandAlgoGoldenPrime.h:12-17 */
#define ORC_GOLDEN_RESEED_SIZE (256 * 1024)
static const uint64_t orcGoldenPrimes[] =
	{0x9e37fffffffc0001ULL, 0x9e3779b97f4a7c15ULL, 0xbf58476d1ce4e5b9ULL, 0x94d049bb133111ebULL};
#define ORC_GOLDEN_PRIMES_LEN (sizeof(orcGoldenPrimes) / sizeof(orcGoldenPrimes[0] ) )

/* RandAlgoGoldenPrime.h:37-42 (seed constructor); the seeder's state is injected because the
 * reference fills it from std::random_device (RandAlgoXoshiro256ss.h:22-29) */
void orc_goldenprime_init(orc_goldenprime* st, uint64_t seed, const uint64_t seederState[4])
{
	memcpy(st->stateSeeder.s, seederState, sizeof(st->stateSeeder.s) );
	st->state = seed;
	st->currentGoldenPrimeIdx = seed % ORC_GOLDEN_PRIMES_LEN;
}

/* RandAlgoGoldenPrime.h:126-132 */
static inline uint64_t orc_goldenprime_next_internal(orc_goldenprime* st)
{
	st->state *= orcGoldenPrimes[st->currentGoldenPrimeIdx];
	st->state >>= 3;
	return st->state;
}

uint64_t orc_goldenprime_next(orc_goldenprime* st)
{
	return orc_goldenprime_next_internal(st);
}

/* RandAlgoGoldenPrime.h:114-118 */
static void orc_goldenprime_reseed(orc_goldenprime* st)
{
	st->state = orc_xoshiro256ss_next(&st->stateSeeder);
	st->currentGoldenPrimeIdx = (st->currentGoldenPrimeIdx + 1) % ORC_GOLDEN_PRIMES_LEN;
}

/* RandAlgoGoldenPrime.h:89-109 */
static void orc_goldenprime_fill_no_reseed(orc_goldenprime* st, char* buf, uint64_t bufLen)
{
	uint64_t numBytesDone = 0;

	for(uint64_t i = 0; i < (bufLen / sizeof(uint64_t) ); i++)
	{
		uint64_t val = orc_goldenprime_next_internal(st);
		memcpy(buf, &val, sizeof(val) );
		buf += sizeof(uint64_t);
		numBytesDone += sizeof(uint64_t);
	}

	if(numBytesDone == bufLen)
		return;

	uint64_t randUint64 = orc_goldenprime_next_internal(st);
	memcpy(buf, &randUint64, bufLen - numBytesDone);
}

/* RandAlgoGoldenPrime.h:58-83 */
void orc_goldenprime_fill_buf(orc_goldenprime* st, char* buf, uint64_t bufLen)
{
	uint64_t numBytesDone = 0;

	for(uint64_t chunkLen = bufLen - numBytesDone;
		chunkLen >= ORC_GOLDEN_RESEED_SIZE;
		chunkLen = bufLen - numBytesDone)
	{
		chunkLen = ORC_GOLDEN_RESEED_SIZE;

		orc_goldenprime_reseed(st);
		orc_goldenprime_fill_no_reseed(st, buf, chunkLen);

		buf += chunkLen;
		numBytesDone += chunkLen;
	}

	orc_goldenprime_reseed(st);
	orc_goldenprime_fill_no_reseed(st, buf, bufLen - numBytesDone);
}

/* workers/LocalWorker.cpp:2209-2230 (preWriteBufRandRefill), rwmix skip rule left to the caller */
void orc_rand_refill_goldenprime(orc_goldenprime* st, char* buf, size_t bufLen, unsigned pct)
{
	const uint64_t varFillLen = ( (uint64_t)bufLen * pct) / 100;
	const size_t constFillRemainderLen = bufLen - varFillLen;

	orc_goldenprime_fill_buf(st, buf, varFillLen);

	if(!constFillRemainderLen)
		return;

	orc_buf_fill(&buf[varFillLen], orc_goldenprime_next(st), constFillRemainderLen);
}

/* ---- CPU twin of the project's counter-based random fill -----------------------------------
 * Layout = the reference's GPU path (workers/LocalWorker.cpp:2236-2277): varFillLen =
 * (bufLen*pct)/100 rounded down to a multiple of 4 (:2251-2256), random u64 words from the start
 * of the block (tail < 8 bytes takes the low bytes of one more word, like fillBuf), remainder =
 * one repeated u64 (bufFill semantics, phase starts at varFillLen).
 * Content = SplitMix64 stream seeded per block (the reference's content is self-seeded, i.e.
 * unpinned): blockKey = mix(seed + ctr*0xD1342543DE82EF95); word k = mix(blockKey + (k+1)*G);
 * remainder value = mix(blockKey). */
static inline uint64_t orc_splitmix64_mix(uint64_t z)
{
	z = (z ^ (z >> 30) ) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27) ) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

void orc_fill_random_ctr(char* buf, uint64_t bufLen, unsigned pct, uint64_t seed,
	uint64_t blockCounter)
{
	const uint64_t golden = 0x9E3779B97F4A7C15ULL;
	const uint64_t blockKey = orc_splitmix64_mix(seed + blockCounter * 0xD1342543DE82EF95ULL);

	uint64_t varFillLen = (bufLen * pct) / 100;
	if(varFillLen % sizeof(int) )
		varFillLen -= (varFillLen % sizeof(int) );

	const uint64_t constFillRemainderLen = bufLen - varFillLen;

	uint64_t state = blockKey;
	uint64_t numBytesDone = 0;

	while(numBytesDone < varFillLen)
	{
		state += golden;
		uint64_t val = orc_splitmix64_mix(state);
		uint64_t copyLen = ORC_MIN( (uint64_t)sizeof(val), varFillLen - numBytesDone);
		memcpy(&buf[numBytesDone], &val, copyLen);
		numBytesDone += copyLen;
	}

	if(!constFillRemainderLen)
		return;

	orc_buf_fill(&buf[varFillLen], orc_splitmix64_mix(blockKey), constFillRemainderLen);
}

/* ============================================================================================
 * Offset generators (toolkits/offsetgen/OffsetGenerator.h, OffsetGenRandomAlignedFullCoverageV2.h)
 * ========================================================================================== */

/* ---- xoshiro256++, one lane (RandAlgoXoshiro256ppSIMD.h:100-135 with NwayInternal 1) ---- */
uint64_t orc_xoshiro256pp_next(orc_xoshiro256pp* st)
{
	const uint64_t x = st->s[0] + st->s[3];
	const uint64_t result = ( (x << 23) | (x >> 41) ) + st->s[0];
	const uint64_t t = st->s[1] << 17;

	st->s[2] ^= st->s[0];
	st->s[3] ^= st->s[1];
	st->s[1] ^= st->s[2];
	st->s[0] ^= st->s[3];
	st->s[2] ^= t;
	st->s[3] = (st->s[3] << 45) | (st->s[3] >> 19);

	return result;
}

/* ---- std::mt19937_64 as the C++ standard defines it ([rand.predef]: w 64, n 312, m 156, r 31,
 * a 0xb5026f5aa96619e9, u 29, d 0x5555555555555555, s 17, b 0x71d67fffeda60000, t 37,
 * c 0xfff7eee000000000, l 43, f 6364136223846793005), what RandAlgoMT19937.h:22 instantiates ---- */
void orc_mt19937_64_seed(orc_mt19937_64* st, uint64_t seed)
{
	st->mt[0] = seed;

	for(unsigned i = 1; i < 312; i++)
		st->mt[i] = 6364136223846793005ULL * (st->mt[i - 1] ^ (st->mt[i - 1] >> 62) ) + i;

	st->idx = 312;
}

uint64_t orc_mt19937_64_next(orc_mt19937_64* st)
{
	if(st->idx >= 312)
	{ // regenerate the whole state block
		const uint64_t upperMask = 0xFFFFFFFF80000000ULL; // upper w-r bits
		const uint64_t lowerMask = 0x000000007FFFFFFFULL; // lower r bits

		for(unsigned i = 0; i < 312; i++)
		{
			const uint64_t y = (st->mt[i] & upperMask) | (st->mt[ (i + 1) % 312] & lowerMask);
			uint64_t next = st->mt[ (i + 156) % 312] ^ (y >> 1);

			if(y & 1)
				next ^= 0xB5026F5AA96619E9ULL;

			st->mt[i] = next;
		}

		st->idx = 0;
	}

	uint64_t z = st->mt[st->idx++];

	z ^= (z >> 29) & 0x5555555555555555ULL;
	z ^= (z << 17) & 0x71D67FFFEDA60000ULL;
	z ^= (z << 37) & 0xFFF7EEE000000000ULL;
	z ^= (z >> 43);

	return z;
}

/* ---- RandAlgoSelectorTk::stringToAlgo (RandAlgoSelectorTk.cpp:37-53) with injected state ---- */
int orc_randalgo_init(orc_randalgo* st, int algo, const uint64_t state[4])
{
	static const uint64_t zeroState[4] = {0, 0, 0, 0};

	if(!state)
		state = zeroState;

	st->algo = algo;

	switch(algo)
	{
		case ELB_OFFSETALGO_XOSHIRO256SS:
			memcpy(st->u.xoshiroSS.s, state, sizeof(st->u.xoshiroSS.s) );
			return 0;
		case ELB_OFFSETALGO_GOLDENPRIME:
			orc_goldenprime_init(&st->u.goldenPrime, state[0], state);
			return 0;
		case ELB_OFFSETALGO_XOSHIRO256PP:
			memcpy(st->u.xoshiroPP.s, state, sizeof(st->u.xoshiroPP.s) );
			return 0;
		case ELB_OFFSETALGO_MT19937:
			orc_mt19937_64_seed(&st->u.mt, state[0] );
			return 0;
		default:
			return -1;
	}
}

uint64_t orc_randalgo_next(orc_randalgo* st)
{
	switch(st->algo)
	{
		case ELB_OFFSETALGO_GOLDENPRIME: return orc_goldenprime_next(&st->u.goldenPrime);
		case ELB_OFFSETALGO_XOSHIRO256PP: return orc_xoshiro256pp_next(&st->u.xoshiroPP);
		case ELB_OFFSETALGO_MT19937: return orc_mt19937_64_next(&st->u.mt);
		default: return orc_xoshiro256ss_next(&st->u.xoshiroSS);
	}
}

orc_randalgo* orc_randalgo_create(int algo, const uint64_t state[4])
{
	orc_randalgo* st = (orc_randalgo*)calloc(1, sizeof(*st) );

	if(st && orc_randalgo_init(st, algo, state) )
	{
		free(st);
		return NULL;
	}

	return st;
}

void orc_randalgo_destroy(orc_randalgo* st)
{
	free(st);
}

struct orc_offsetgen
{
	int kind;
	uint64_t numBytesTotal;
	uint64_t numBytesLeft;
	uint64_t startOffset;   // sequential/reverse/strided; "offset" of random aligned
	uint64_t currentOffset;
	uint64_t blockSize;
	uint64_t numDataSetThreads; // strided
	// RandAlgoRange (toolkits/random/RandAlgoRange.h:14-56)
	orc_randalgo rand;
	uint64_t rangeStart;
	uint64_t rangeLengthPlusOne;
	// CoveringRandomGenerator (OffsetGenRandomAlignedFullCoverageV2.h:9-203)
	uint64_t rangeLen;
	uint64_t covMin, covMax, covRangeSize, covCount, covM, covState, covNextSeed;
};

static uint64_t orc_range_next(orc_offsetgen* g) // RandAlgoRange.h:50-54
{
	return (orc_randalgo_next(&g->rand) % g->rangeLengthPlusOne) + g->rangeStart;
}

static void orc_range_reset(orc_offsetgen* g, uint64_t min, uint64_t max) // RandAlgoRange.h:40-48
{
	g->rangeStart = min;
	g->rangeLengthPlusOne = max - min + 1;
}

static uint64_t orc_next_power_of_two(uint64_t n) // FullCoverageV2.h:190-201
{
	if(n == 0)
		return 1;
	n--;
	n |= n >> 1;
	n |= n >> 2;
	n |= n >> 4;
	n |= n >> 8;
	n |= n >> 16;
	n |= n >> 32;
	n++;
	return n;
}

/* FullCoverageV2.h:25-105 (constructor); the seed replaces std::random_device. Each later
 * re-seed (reset(), :155-164) takes the next value of a simple counter sequence derived from the
 * injected seed so that runs stay reproducible. */
static void orc_cov_init(orc_offsetgen* g, uint64_t minVal, uint64_t maxVal)
{
	g->covMin = minVal;
	g->covMax = maxVal;
	g->covCount = 0;
	g->covRangeSize = (maxVal - minVal) + 1;
	g->covM = orc_next_power_of_two(g->covRangeSize);
	if( (g->covM == 0) && (g->covRangeSize > 0) )
		g->covM = g->covRangeSize;
	g->covState = (uint32_t)g->covNextSeed; // random_device yields 32 bits
	g->covNextSeed = g->covNextSeed * 6364136223846793005ULL + 1442695040888963407ULL;
	g->covState %= g->covM;
}

static uint64_t orc_cov_next(orc_offsetgen* g) // FullCoverageV2.h:115-139
{
	if( !(g->covCount < g->covRangeSize) )
	{ // reset(): new permutation (:155-164)
		g->covCount = 0;
		g->covState = (uint32_t)g->covNextSeed;
		g->covNextSeed = g->covNextSeed * 6364136223846793005ULL + 1442695040888963407ULL;
		g->covState %= g->covM;
	}

	uint64_t val;
	do
	{
		g->covState = (6364136223846793005ULL * g->covState + 1442695040888963407ULL) % g->covM;
		val = g->covState;
	} while(val >= g->covRangeSize);

	g->covCount++;
	return g->covMin + val;
}

static uint64_t orc_cov_start_block(uint64_t offset, uint64_t blockSize) // :293-296
{
	return blockSize ? (offset / blockSize) : 0;
}

static uint64_t orc_cov_end_block(uint64_t offset, uint64_t rangeLen, uint64_t blockSize) // :301-305
{
	return orc_cov_start_block(offset, blockSize) +
		( (blockSize && (rangeLen / blockSize) ) ? (rangeLen / blockSize) - 1 : 0);
}

void orc_offsetgen_reset(orc_offsetgen* g)
{
	g->numBytesLeft = g->numBytesTotal;

	switch(g->kind)
	{
		case ORC_OFFGEN_SEQUENTIAL: // OffsetGenerator.h:68-72
		case ORC_OFFGEN_STRIDED:    // :343-347
			g->currentOffset = g->startOffset;
			break;

		case ORC_OFFGEN_REVERSE_SEQ: // :127-146
		{
			if(!g->numBytesTotal)
			{
				g->currentOffset = 0;
				break;
			}

			uint64_t lastBlockRemainder = g->numBytesTotal % g->blockSize;

			if(lastBlockRemainder)
				g->currentOffset = g->startOffset + g->numBytesTotal - lastBlockRemainder;
			else
				g->currentOffset = g->startOffset + g->numBytesTotal - g->blockSize;
		} break;

		case ORC_OFFGEN_RANDOM_ALIGNED_FULLCOV: // FullCoverageV2.h:245-250
		{
			g->covCount = 0;
			g->covState = (uint32_t)g->covNextSeed;
			g->covNextSeed = g->covNextSeed * 6364136223846793005ULL + 1442695040888963407ULL;
			g->covState %= g->covM;
		} break;

		default: // random / random aligned (:214-215, :285-286)
			break;
	}
}

void orc_offsetgen_reset_range(orc_offsetgen* g, uint64_t len, uint64_t offset)
{
	g->numBytesTotal = len;
	g->numBytesLeft = len;

	switch(g->kind)
	{
		case ORC_OFFGEN_SEQUENTIAL: // :74-80
		case ORC_OFFGEN_STRIDED:    // :349-355
			g->startOffset = offset;
			g->currentOffset = offset;
			break;

		case ORC_OFFGEN_REVERSE_SEQ: // :148-156
			g->startOffset = offset;
			g->currentOffset = offset;
			orc_offsetgen_reset(g);
			break;

		case ORC_OFFGEN_RANDOM: // :217-227
		{
			uint64_t minLenAndBlockSize = ORC_MIN(g->blockSize, len);
			orc_range_reset(g, offset, offset + len - minLenAndBlockSize);
		} break;

		case ORC_OFFGEN_RANDOM_ALIGNED: // :288-302
		{
			uint64_t minLenAndBlockSize = ORC_MIN(g->blockSize, len);
			g->startOffset = offset;

			if(minLenAndBlockSize)
				orc_range_reset(g, 0, (len - minLenAndBlockSize) / minLenAndBlockSize);
			else
				orc_range_reset(g, 0, 0);
		} break;

		case ORC_OFFGEN_RANDOM_ALIGNED_FULLCOV: // FullCoverageV2.h:257-266
			g->rangeLen = len;
			orc_cov_init(g, orc_cov_start_block(offset, g->blockSize),
				orc_cov_end_block(offset, len, g->blockSize) );
			break;
	}
}

orc_offsetgen* orc_offsetgen_create(int kind, uint64_t numBytesTotal, uint64_t len,
	uint64_t offset, uint64_t blockSize, uint64_t numDataSetThreads,
	const uint64_t randState[4], uint64_t lcgSeed)
{
	return orc_offsetgen_create_algo(kind, numBytesTotal, len, offset, blockSize,
		numDataSetThreads, ELB_OFFSETALGO_XOSHIRO256SS, randState, lcgSeed);
}

orc_offsetgen* orc_offsetgen_create_algo(int kind, uint64_t numBytesTotal, uint64_t len,
	uint64_t offset, uint64_t blockSize, uint64_t numDataSetThreads, int randAlgo,
	const uint64_t randState[4], uint64_t lcgSeed)
{
	orc_offsetgen* g = (orc_offsetgen*)calloc(1, sizeof(*g) );
	if(!g)
		return NULL;

	g->kind = kind;
	g->blockSize = blockSize;
	g->numDataSetThreads = numDataSetThreads;
	g->covNextSeed = lcgSeed;

	if(orc_randalgo_init(&g->rand, randAlgo, randState) )
	{
		free(g);
		return NULL;
	}

	switch(kind)
	{
		case ORC_OFFGEN_SEQUENTIAL: // OffsetGenerator.h:51-54
		case ORC_OFFGEN_STRIDED:    // :326-330
			g->numBytesTotal = len;
			g->numBytesLeft = len;
			g->startOffset = offset;
			g->currentOffset = offset;
			break;

		case ORC_OFFGEN_REVERSE_SEQ: // :109-114
			g->numBytesTotal = len;
			g->numBytesLeft = len;
			g->startOffset = offset;
			g->currentOffset = offset;
			orc_offsetgen_reset(g);
			break;

		case ORC_OFFGEN_RANDOM: // :191-199
			g->numBytesTotal = numBytesTotal;
			g->numBytesLeft = numBytesTotal;
			orc_range_reset(g, offset, offset + len - ORC_MIN(blockSize, len) );
			break;

		case ORC_OFFGEN_RANDOM_ALIGNED: // :255-268
		{
			uint64_t minLen = ORC_MIN(blockSize, len);
			g->numBytesTotal = numBytesTotal;
			g->numBytesLeft = numBytesTotal;
			g->startOffset = offset;
			orc_range_reset(g, 0, !minLen ? 0 : (len - minLen) / minLen);
		} break;

		case ORC_OFFGEN_RANDOM_ALIGNED_FULLCOV: // FullCoverageV2.h:219-228
			g->numBytesTotal = numBytesTotal;
			g->numBytesLeft = numBytesTotal;
			g->rangeLen = len;
			orc_cov_init(g, orc_cov_start_block(offset, blockSize),
				orc_cov_end_block(offset, len, blockSize) );
			break;

		default:
			free(g);
			return NULL;
	}

	return g;
}

void orc_offsetgen_destroy(orc_offsetgen* g)
{
	free(g);
}

uint64_t orc_offsetgen_next_offset(orc_offsetgen* g)
{
	switch(g->kind)
	{
		case ORC_OFFGEN_RANDOM: // :229-230
			return orc_range_next(g);
		case ORC_OFFGEN_RANDOM_ALIGNED: // :304-305
			return g->startOffset + (orc_range_next(g) * g->blockSize);
		case ORC_OFFGEN_RANDOM_ALIGNED_FULLCOV: // FullCoverageV2.h:268-269
			return orc_cov_next(g) * g->blockSize;
		default: // :82-83, :158-159, :357-358
			return g->currentOffset;
	}
}

uint64_t orc_offsetgen_next_block_size(const orc_offsetgen* g)
{
	if(g->kind == ORC_OFFGEN_REVERSE_SEQ) // :164-165
		return ORC_MIN(g->startOffset + g->numBytesTotal - g->currentOffset, g->blockSize);

	return ORC_MIN(g->numBytesLeft, g->blockSize); // :88-89 etc.
}

uint64_t orc_offsetgen_bytes_total(const orc_offsetgen* g)
{
	return g->numBytesTotal;
}

uint64_t orc_offsetgen_bytes_left(const orc_offsetgen* g)
{
	return g->numBytesLeft;
}

void orc_offsetgen_add_bytes_submitted(orc_offsetgen* g, uint64_t numBytes)
{
	g->numBytesLeft -= numBytes;

	switch(g->kind)
	{
		case ORC_OFFGEN_SEQUENTIAL: // :97-101
			g->currentOffset += numBytes;
			break;
		case ORC_OFFGEN_REVERSE_SEQ: // :173-177
			g->currentOffset -= g->blockSize;
			break;
		case ORC_OFFGEN_STRIDED: // :372-376
			g->currentOffset += (g->blockSize * g->numDataSetThreads);
			break;
		default:
			break;
	}
}

/* ============================================================================================
 * LatencyHistogram (LatencyHistogram.h) / UnitTk
 * ========================================================================================== */

void orc_histogram_reset(elb_histogram* h) // :113-123
{
	memset(h->buckets, 0, sizeof(h->buckets) );
	h->numStoredValues = 0;
	h->numMicroSecTotal = 0;
	h->minMicroSecLat = ~0ULL;
	h->maxMicroSecLat = 0;
}

void orc_histogram_add_latency(elb_histogram* h, uint64_t latencyMicroSec) // :50-77
{
	h->numStoredValues++;
	h->numMicroSecTotal += latencyMicroSec;

	if(latencyMicroSec < h->minMicroSecLat)
		h->minMicroSecLat = latencyMicroSec;

	if(latencyMicroSec > h->maxMicroSecLat)
		h->maxMicroSecLat = latencyMicroSec;

	size_t bucketIndex;

	if(!latencyMicroSec)
		bucketIndex = 0;
	else
		bucketIndex = (size_t)(log2( (double)latencyMicroSec) * 4);

	if(bucketIndex >= ELB_LATHISTO_NUMBUCKETS)
		bucketIndex = ELB_LATHISTO_NUMBUCKETS - 1;

	h->buckets[bucketIndex]++;
}

void orc_histogram_merge(elb_histogram* dst, const elb_histogram* src) // :187-202
{
	for(size_t i = 0; i < ELB_LATHISTO_NUMBUCKETS; i++)
		dst->buckets[i] += src->buckets[i];

	dst->numStoredValues += src->numStoredValues;
	dst->numMicroSecTotal += src->numMicroSecTotal;

	if(src->minMicroSecLat < dst->minMicroSecLat)
		dst->minMicroSecLat = src->minMicroSecLat;

	if(src->maxMicroSecLat > dst->maxMicroSecLat)
		dst->maxMicroSecLat = src->maxMicroSecLat;
}

double orc_histogram_percentile(const elb_histogram* h, double percentage) // :140-159
{
	size_t numValuesSoFar = 0;
	double log2BucketSize = 1.0 / 4;

	for(size_t bucketIndex = 0; bucketIndex < ELB_LATHISTO_NUMBUCKETS; bucketIndex++)
	{
		numValuesSoFar += h->buckets[bucketIndex];

		double percentileSoFar = (double)numValuesSoFar / h->numStoredValues;

		if(percentileSoFar >= (percentage / 100) )
			return pow(2, (bucketIndex + 1) * log2BucketSize);
	}

	return 0;
}

uint64_t orc_per_sec_from_usec(uint64_t totalValue, uint64_t elapsedUSec) // toolkits/UnitTk.h:48-56
{
	const double numUSecsPerSec = 1000000;
	return (uint64_t)(totalValue * (numUSecsPerSec / elapsedUSec) );
}

/* ============================================================================================
 * The CPU LocalWorker
 * ========================================================================================== */

typedef struct orc_shared
{
	const elb_cfg* cfg;
	int benchPhase;
	uint64_t blockSize;   // after ProgArgs-style normalisation
	uint64_t fileSize;
	uint64_t randomAmount;
	uint32_t numDataSetThreads;
	int* pathFDs;
	struct timespec phaseStartT;
	pthread_mutex_t mutex;
	int numWorkersDone;
	int stoneWallTriggered;
	struct orc_worker* workers;
} orc_shared;

typedef struct orc_worker
{
	orc_shared* shared;
	uint64_t rank;
	char* ioBuf;
	char** ioBufVec;       // iodepth buffers of the aio loop (ioBufVec[0] == ioBuf)
	struct iocb* iocbVec;  // libaioContext.iocbVec
	struct timespec* ioStartTimeVec; // libaioContext.ioStartTimeVec (tv_sec < 0: invalidated)
	aio_context_t aioContext;
	int aioInitialized;
	orc_ratelimiter rateLimiter;
	orc_offsetgen* offsetGen;
	uint64_t numIOPSSubmitted;
	int currentFD; // fdVec[0] of sequential/dir mode
	orc_goldenprime blockVarAlgo;
	int isRWMixReader; // --rwmixthr reader in a write phase (workers/LocalWorker.cpp:1028-1041)
	orc_worker_result* res;
	elb_liveops stoneWallOps;
	elb_liveops stoneWallOpsReadMix; // Worker.h:203-209 snapshots both counter sets
	char errTmp[512];
} orc_worker;

static uint64_t orc_elapsed_usec(const struct timespec* start)
{
	struct timespec now;
	clock_gettime(CLOCK_MONOTONIC, &now);
	int64_t nsec = (int64_t)(now.tv_sec - start->tv_sec) * 1000000000LL +
		(now.tv_nsec - start->tv_nsec);
	return (uint64_t)(nsec / 1000);
}

/* ProgArgs.cpp:1531-1561, 1599-1637: block size / file size / random amount normalisation */
static void orc_normalize(orc_shared* sh)
{
	const elb_cfg* cfg = sh->cfg;

	sh->blockSize = cfg->blockSize;
	sh->fileSize = cfg->fileSize;
	sh->randomAmount = cfg->randomAmount;
	sh->numDataSetThreads = cfg->numDataSetThreads ? cfg->numDataSetThreads : cfg->numThreads;

	if(sh->blockSize > sh->fileSize)
		sh->blockSize = sh->fileSize; // :1531-1540

	if( (cfg->useDirectIO || cfg->useRandomOffsets || cfg->useStridedAccess) && sh->fileSize &&
		sh->blockSize && (sh->fileSize % sh->blockSize) )
		sh->fileSize -= (sh->fileSize % sh->blockSize); // :1543-1555

	if(!sh->randomAmount && (cfg->pathType != ELB_PATH_DIR) && cfg->useRandomOffsets)
		sh->randomAmount = sh->fileSize * cfg->numPaths; // :1558-1561

	if(cfg->useRandomOffsets && !cfg->useRandomUnaligned && sh->blockSize &&
		(sh->randomAmount % sh->blockSize) && (cfg->pathType != ELB_PATH_DIR) )
		sh->randomAmount -= (sh->randomAmount % sh->blockSize); // :1585-1596

	if(cfg->pathType != ELB_PATH_DIR)
	{
		const uint64_t blockSetSize = sh->blockSize * sh->numDataSetThreads;

		if(cfg->useRandomOffsets && !cfg->useRandomUnaligned && blockSetSize &&
			(sh->randomAmount % blockSetSize) )
			sh->randomAmount -= (sh->randomAmount % blockSetSize); // :1625-1637
	}
}

static void orc_worker_fail(orc_worker* w, const char* msg)
{
	w->res->hadError = 1;
	snprintf(w->res->errorMsg, sizeof(w->res->errorMsg), "%s", msg);
}

/* workers/LocalWorker.cpp:2051-2074 */
static void orc_calc_file_idx_and_offset(uint64_t rwOffsetGenNext, uint64_t fileSize,
	int isSingleFile, size_t* outFileIdx, uint64_t* outFileOffset)
{
	if(isSingleFile)
	{
		*outFileIdx = 0;
		*outFileOffset = rwOffsetGenNext;
	}
	else
	{
		*outFileIdx = rwOffsetGenNext / fileSize;
		*outFileOffset = rwOffsetGenNext % fileSize;
	}
}

/* toolkits/RateLimiter.h:13-66: a budget per second; whoever would exceed it sleeps until the
 * second is over. initStart at phase start (workers/LocalWorker.cpp:1293-1299, 1331-1337). */
void orc_ratelimiter_init_start(orc_ratelimiter* rl, uint64_t limitPerSec)
{
	rl->limitPerSec = limitPerSec;
	rl->numDoneThisSec = 0;
	clock_gettime(CLOCK_MONOTONIC, &rl->startT);
}

/* @return 1 if the caller had to wait (RateLimiter::wait), 0 = go on immediately.
 * limitPerSec 0 = noOpRateLimiter (workers/LocalWorker.cpp:2364-2367). */
int orc_ratelimiter_wait(orc_ratelimiter* rl, uint64_t nextSize)
{
	if(!rl->limitPerSec)
		return 0;

	const uint64_t elapsedMicroSec = orc_elapsed_usec(&rl->startT);

	if(elapsedMicroSec >= 1000000)
	{ // 1s elapsed without exceeding the rate limit => reset for next second
		rl->numDoneThisSec = nextSize;
		clock_gettime(CLOCK_MONOTONIC, &rl->startT);
		return 0;
	}

	if( (rl->numDoneThisSec + nextSize) > rl->limitPerSec)
	{ // next r/w op would exceed rate limit => wait until end of second
		struct timespec wakeT = rl->startT;
		wakeT.tv_sec += 1;

		while(clock_nanosleep(CLOCK_MONOTONIC, TIMER_ABSTIME, &wakeT, NULL) == EINTR)
			;

		rl->numDoneThisSec = nextSize;
		clock_gettime(CLOCK_MONOTONIC, &rl->startT);
		return 1;
	}

	rl->numDoneThisSec += nextSize;
	return 0;
}

/* fill (write phase) as funcPreWriteBlockModifier does it (:1256-1265, rwmix exception :2213) */
static void orc_pre_write_modify(orc_worker* w, char* buf, size_t blockSize, uint64_t offset)
{
	const elb_cfg* cfg = w->shared->cfg;

	if(cfg->integrityCheckSalt)
		orc_fill_pattern(buf, blockSize, offset, cfg->integrityCheckSalt);
	else
	if(cfg->blockVariancePercent &&
		!( ( (w->rank + w->numIOPSSubmitted) % 100) < cfg->rwMixReadPercent) )
		orc_rand_refill_goldenprime(&w->blockVarAlgo, buf, blockSize, cfg->blockVariancePercent);
}

/* aioReadPrepper / aioWritePrepper / aioRWMixPrepper (workers/LocalWorker.cpp:2400-2440): which
 * opcode the next request gets */
static void orc_aio_prep(orc_worker* w, struct iocb* cb, int fd, char* buf, size_t count,
	uint64_t offset, int isRead)
{
	const elb_cfg* cfg = w->shared->cfg;
	int doRead = isRead;

	if(!isRead && cfg->rwMixReadPercent &&
		( ( (w->rank + w->numIOPSSubmitted) % 100) < cfg->rwMixReadPercent) )
		doRead = 1; // aioRWMixPrepper

	memset(cb, 0, sizeof(*cb) );
	cb->aio_lio_opcode = doRead ? IOCB_CMD_PREAD : IOCB_CMD_PWRITE;
	cb->aio_fildes = fd;
	cb->aio_buf = (uint64_t)(uintptr_t)buf;
	cb->aio_nbytes = count;
	cb->aio_offset = offset;
}

/* prepare request slot ioVecIdx with the generator's next block and submit it: the body that
 * phase 1 (:1819-1868) and the resubmission of phase 2 (:1977-2027) share. Order of the steps as
 * in the reference: prep, START STAMP, rate limiter (a wait invalidates the stamps of ALL pending
 * requests, :1843-1845), block modifier, io_submit of this one iocb. @return 0 ok, -1 error. */
static int orc_aio_submit_next(orc_worker* w, const int* fdVec, int isSingleFile, int isRead,
	size_t ioVecIdx, size_t numSlots)
{
	orc_shared* sh = w->shared;
	orc_offsetgen* gen = w->offsetGen;
	const uint64_t rwOffsetGenNext = orc_offsetgen_next_offset(gen);
	const size_t blockSize = orc_offsetgen_next_block_size(gen);
	uint64_t currentOffset;
	size_t fileHandlesIdx;
	struct iocb* cb = &w->iocbVec[ioVecIdx];
	struct iocb* cbPtr = cb;

	orc_calc_file_idx_and_offset(rwOffsetGenNext, sh->fileSize, isSingleFile, &fileHandlesIdx,
		&currentOffset);

	orc_aio_prep(w, cb, fdVec[fileHandlesIdx], w->ioBufVec[ioVecIdx], blockSize, currentOffset,
		isRead);
	cb->aio_data = ioVecIdx; // the vec index of this request

	clock_gettime(CLOCK_MONOTONIC, &w->ioStartTimeVec[ioVecIdx] );

	if(orc_ratelimiter_wait(&w->rateLimiter, blockSize) )
		for(size_t i = 0; i < numSlots; i++)
			w->ioStartTimeVec[i].tv_sec = -1; // time_point::min(): not counted (:1935, 1946)

	if(!isRead)
		orc_pre_write_modify(w, w->ioBufVec[ioVecIdx], blockSize, currentOffset);

	if(syscall(SYS_io_submit, w->aioContext, 1L, &cbPtr) != 1)
	{
		snprintf(w->errTmp, sizeof(w->errTmp),
			"Async IO submission (io_submit) failed. SysErr: %s", strerror(errno) );
		orc_worker_fail(w, w->errTmp);
		return -1;
	}

	w->numIOPSSubmitted++;
	orc_offsetgen_add_bytes_submitted(gen, blockSize);

	return 0;
}

/* workers/LocalWorker.cpp:1795-2037 (aioBlockSized) on the raw kernel AIO ABI (the reference uses
 * libaio's wrappers of the same syscalls): phase 1 seeds up to iodepth requests one io_submit
 * each, phase 2 reaps 1..ORC_AIO_MAX_EVENTS completions at a time, checks, verifies (read phase),
 * accounts (latency only if the start stamp was not invalidated) and reuses the slot for the next
 * block. Returns like the reference: total bytes, -1 (errno set), the partial byte count, or -2
 * after a verify failure / submission error (message set). */
static int64_t orc_aio_block_sized(orc_worker* w, const int* fdVec, size_t numFDs, int isRead)
{
	orc_shared* sh = w->shared;
	const elb_cfg* cfg = sh->cfg;
	orc_offsetgen* gen = w->offsetGen;
	const size_t maxIODepth = cfg->ioDepth;
	const int isSingleFile = (numFDs == 1);
	size_t numPending = 0;
	uint64_t numBytesDone = 0;
	struct io_event ioEvents[ORC_AIO_MAX_EVENTS];

	// P H A S E 1: initial seed of io submissions up to full ioDepth
	while(orc_offsetgen_bytes_left(gen) && (numPending < maxIODepth) )
	{
		if(orc_aio_submit_next(w, fdVec, isSingleFile, isRead, numPending, maxIODepth) )
			return -2;

		numPending++;
	}

	// P H A S E 2: wait for submissions to complete and submit new requests if bytes left
	while(numPending)
	{
		struct timespec ioTimeout = {ORC_AIO_MAX_WAIT_SEC, 0};

		long eventsRes = syscall(SYS_io_getevents, w->aioContext, 1L, (long)ORC_AIO_MAX_EVENTS,
			ioEvents, &ioTimeout);

		if(!eventsRes || ( (eventsRes < 0) && (errno == EINTR) ) )
			continue; // timeout expired (only there to check interruptions)

		if(eventsRes < 0)
		{
			snprintf(w->errTmp, sizeof(w->errTmp),
				"Getting async IO events (io_getevents) failed. NumPending: %zu; SysErr: %s",
				numPending, strerror(errno) );
			orc_worker_fail(w, w->errTmp);
			return -2;
		}

		for(long eventIdx = 0; eventIdx < eventsRes; eventIdx++)
		{
			const struct io_event* event = &ioEvents[eventIdx];
			struct iocb* cb = (struct iocb*)(uintptr_t)event->obj;
			const size_t ioVecIdx = (size_t)event->data;

			if(event->res2 || (event->res != (int64_t)cb->aio_nbytes) )
			{ // unexpected result (:1900-1927)
				if(event->res2)
				{
					snprintf(w->errTmp, sizeof(w->errTmp), "Async IO framework error. "
						"res: %lld; res2: %lld", (long long)event->res, (long long)event->res2);
					orc_worker_fail(w, w->errTmp);
					return -2;
				}

				if(event->res < 0)
				{
					errno = -(int)event->res;
					return -1;
				}

				return (int64_t)(numBytesDone + event->res); // partial read/write
			}

			const int wasRead = (cb->aio_lio_opcode == IOCB_CMD_PREAD);

			if(isRead && cfg->integrityCheckSalt)
			{ // funcPostReadBlockChecker (:1938-1940): only set in a read phase (:1318-1319)
				int verifyRes = orc_verify_pattern( (char*)(uintptr_t)cb->aio_buf, cb->aio_nbytes,
					cb->aio_offset, cfg->integrityCheckSalt, NULL, NULL, NULL, NULL, w->errTmp,
					sizeof(w->errTmp) );
				if(verifyRes)
				{
					orc_worker_fail(w, w->errTmp);
					return -2;
				}
			}

			const int latencyValid = (w->ioStartTimeVec[ioVecIdx].tv_sec >= 0);
			const uint64_t ioElapsedMicroSec =
				latencyValid ? orc_elapsed_usec(&w->ioStartTimeVec[ioVecIdx] ) : 0;

			numBytesDone += event->res;

			if( (wasRead && (sh->benchPhase == ELB_PHASE_CREATEFILES) ) )
			{ // read in a write phase => rwmix read stats (:1951-1963; no histogram kept here)
				__atomic_fetch_add(&w->res->liveOpsReadMix.numBytesDone, (uint64_t)event->res,
					__ATOMIC_RELAXED);
				__atomic_fetch_add(&w->res->liveOpsReadMix.numIOPSDone, 1, __ATOMIC_RELAXED);
			}
			else
			{
				if(latencyValid) // (:1966-1969)
					orc_histogram_add_latency(&w->res->iopsLatHisto, ioElapsedMicroSec);

				__atomic_fetch_add(&w->res->liveOps.numBytesDone, (uint64_t)event->res,
					__ATOMIC_RELAXED);
				__atomic_fetch_add(&w->res->liveOps.numIOPSDone, 1, __ATOMIC_RELAXED);
			}

			if(!orc_offsetgen_bytes_left(gen) )
			{
				numPending--;
				continue;
			}

			// request complete, so reuse iocb for the next request
			if(orc_aio_submit_next(w, fdVec, isSingleFile, isRead, ioVecIdx, maxIODepth) )
				return -2;
		}
	}

	return orc_offsetgen_bytes_total(gen);
}

/* funcRWBlockSized (workers/LocalWorker.cpp:1243-1244, 1305-1306): the sync loop for iodepth 1,
 * else the aio loop */
static int64_t orc_rw_block_sized(orc_worker* w, const int* fdVec, size_t numFDs, int isRead);

static int64_t orc_rw_any_block_sized(orc_worker* w, const int* fdVec, size_t numFDs, int isRead)
{
	if( (w->shared->cfg->ioDepth > 1) && w->aioInitialized)
		return orc_aio_block_sized(w, fdVec, numFDs, isRead);

	return orc_rw_block_sized(w, fdVec, numFDs, isRead);
}

/* workers/LocalWorker.cpp:1669-1781 (rwBlockSized) with the CPU policy of :1188-1355:
 * write phase: modifier = pattern fill if salt != 0, else random refill if blockvarpct, else noop;
 * read phase: checker = pattern verify if salt != 0. Returns like the reference: total bytes,
 * -1 (errno set) or the partial byte count. A verify failure sets w->res->hadError. */
static int64_t orc_rw_block_sized(orc_worker* w, const int* fdVec, size_t numFDs, int isRead)
{
	orc_shared* sh = w->shared;
	const elb_cfg* cfg = sh->cfg;
	orc_offsetgen* gen = w->offsetGen;
	const int isSingleFile = (numFDs == 1);
	const unsigned rwMixReadPercent = cfg->rwMixReadPercent;

	while(orc_offsetgen_bytes_left(gen) )
	{
		const uint64_t rwOffsetGenNext = orc_offsetgen_next_offset(gen);
		const size_t currentBlockSize = orc_offsetgen_next_block_size(gen);
		uint64_t currentOffset;
		size_t fileHandleIdx;
		int isRWMixRead = 0;
		ssize_t rwRes;

		orc_calc_file_idx_and_offset(rwOffsetGenNext, sh->fileSize, isSingleFile,
			&fileHandleIdx, &currentOffset);

		orc_ratelimiter_wait(&w->rateLimiter, currentBlockSize); // funcRWRateLimiter (:1689)

		struct timespec ioStartT;
		clock_gettime(CLOCK_MONOTONIC, &ioStartT);

		if(!isRead) // funcPreWriteBlockModifier (:1693-1694)
			orc_pre_write_modify(w, w->ioBuf, currentBlockSize, currentOffset);

		if(isRead)
		{ // this is a read, but could be a rwmix read thread (:1697-1706)
			isRWMixRead = w->isRWMixReader;
			rwRes = pread(fdVec[fileHandleIdx], w->ioBuf, currentBlockSize, currentOffset);
		}
		else
		if(rwMixReadPercent && ( ( (w->rank + w->numIOPSSubmitted) % 100) < rwMixReadPercent) )
		{ // :1708-1718
			isRWMixRead = 1;
			rwRes = pread(fdVec[fileHandleIdx], w->ioBuf, currentBlockSize, currentOffset);
		}
		else
		{
			rwRes = pwrite(fdVec[fileHandleIdx], w->ioBuf, currentBlockSize, currentOffset);

			// pwriteAndReadWrapper for --verifydirect/--readinline (:2533-2554)
			if( (rwRes > 0) && (cfg->doDirectVerify || cfg->doReadInline) )
				rwRes = pread(fdVec[fileHandleIdx], w->ioBuf, rwRes, currentOffset);
		}

		if(rwRes <= 0)
			return (rwRes < 0) ? rwRes :
				(int64_t)(orc_offsetgen_bytes_total(gen) - orc_offsetgen_bytes_left(gen) );

		if( (isRead || (!isRWMixRead && cfg->doDirectVerify) ) && cfg->integrityCheckSalt)
		{ // funcPostReadBlockChecker (:1318-1319; --verifydirect :1279-1280)
			int verifyRes = orc_verify_pattern(w->ioBuf, currentBlockSize, currentOffset,
				cfg->integrityCheckSalt, NULL, NULL, NULL, NULL, w->errTmp, sizeof(w->errTmp) );
			if(verifyRes)
			{
				orc_worker_fail(w, w->errTmp);
				return -2;
			}
		}

		uint64_t ioElapsedMicroSec = orc_elapsed_usec(&ioStartT);

		if(isRWMixRead)
		{
			__atomic_fetch_add(&w->res->liveOpsReadMix.numBytesDone, (uint64_t)rwRes,
				__ATOMIC_RELAXED);
			__atomic_fetch_add(&w->res->liveOpsReadMix.numIOPSDone, 1, __ATOMIC_RELAXED);
		}
		else
		{
			orc_histogram_add_latency(&w->res->iopsLatHisto, ioElapsedMicroSec);
			__atomic_fetch_add(&w->res->liveOps.numBytesDone, (uint64_t)rwRes, __ATOMIC_RELAXED);
			__atomic_fetch_add(&w->res->liveOps.numIOPSDone, 1, __ATOMIC_RELAXED);
		}

		w->numIOPSSubmitted++;
		orc_offsetgen_add_bytes_submitted(gen, rwRes);
	}

	return orc_offsetgen_bytes_total(gen);
}

static void orc_set_io_error(orc_worker* w, int isRead, int64_t rwRes, uint64_t expected,
	const char* path)
{
	char msg[512];

	if(rwRes == -2)
		return; // verify failure, message already set

	if(rwRes == -1) // workers/LocalWorker.cpp:3657-3663
		snprintf(msg, sizeof(msg), "File %s failed. Path: %s; SysErr: %s",
			isRead ? "read" : "write", path, strerror(errno) );
	else // :3665-3670
		snprintf(msg, sizeof(msg), "Unexpected short file %s. Path: %s; Bytes %s: %lld; "
			"Expected %s: %llu; Hint: Consider initial sequential write or adding "
			"\"--trunctosize\" to ensure full file size.",
			isRead ? "read" : "write", path, isRead ? "read" : "written", (long long)rwRes,
			isRead ? "read" : "written", (unsigned long long)expected);

	orc_worker_fail(w, msg);
}

/* workers/LocalWorker.cpp:3564-3729 (fileModeIterateFilesSeq) */
static void orc_file_mode_iterate_seq(orc_worker* w, int isRead)
{
	orc_shared* sh = w->shared;
	const elb_cfg* cfg = sh->cfg;
	const size_t numFiles = cfg->numPaths;
	const uint64_t fileSize = sh->fileSize;
	const uint64_t blockSize = sh->blockSize;
	const size_t numThreads = sh->numDataSetThreads;

	const uint64_t numBlocksPerFile = (fileSize / blockSize) + ( (fileSize % blockSize) ? 1 : 0);
	const uint64_t numBlocksTotal = numBlocksPerFile * numFiles;
	const uint64_t standardWorkerNumBlocks = numBlocksTotal / numThreads;

	uint64_t thisWorkerNumBlocks = standardWorkerNumBlocks;
	if( (w->rank == (numThreads - 1) ) && (numBlocksTotal % numThreads) )
		thisWorkerNumBlocks = numBlocksTotal - (standardWorkerNumBlocks * (numThreads - 1) );

	uint64_t startBlock = w->rank * standardWorkerNumBlocks;
	uint64_t endBlock = startBlock + thisWorkerNumBlocks;

	if(startBlock >= endBlock)
	{
		w->res->gotPhaseWork = 0;
		return;
	}

	uint64_t currentBlockIdx = startBlock;

	while(currentBlockIdx < endBlock)
	{
		const uint64_t currentFileIndex = currentBlockIdx / numBlocksPerFile;
		w->currentFD = sh->pathFDs[currentFileIndex];

		const uint64_t currentBlockInFile = currentBlockIdx % numBlocksPerFile;
		const uint64_t currentIOStart = currentBlockInFile * blockSize;
		const uint64_t remainingWorkerLen = (endBlock - currentBlockIdx) * blockSize;
		const uint64_t remainingFileLen = fileSize - (currentBlockInFile * blockSize);
		const uint64_t currentIOLen = ORC_MIN(remainingWorkerLen, remainingFileLen);

		orc_offsetgen_reset_range(w->offsetGen, currentIOLen, currentIOStart);

		int64_t rwRes = orc_rw_any_block_sized(w, &w->currentFD, 1, isRead);

		if( (rwRes < 0) || ( (uint64_t)rwRes != currentIOLen) )
		{
			orc_set_io_error(w, isRead, rwRes, currentIOLen, cfg->paths[currentFileIndex] );
			return;
		}

		const uint64_t numBlocksDone = (currentIOLen / blockSize) +
			( (currentIOLen % blockSize) ? 1 : 0);

		currentBlockIdx += numBlocksDone;
	}
}

/* workers/LocalWorker.cpp:3478-3556 (fileModeIterateFilesRand) */
static void orc_file_mode_iterate_rand(orc_worker* w, int isRead)
{
	orc_shared* sh = w->shared;

	int64_t rwRes = orc_rw_any_block_sized(w, sh->pathFDs, sh->cfg->numPaths, isRead);

	if( (rwRes < 0) || ( (uint64_t)rwRes != orc_offsetgen_bytes_total(w->offsetGen) ) )
		orc_set_io_error(w, isRead, rwRes, orc_offsetgen_bytes_total(w->offsetGen),
			sh->cfg->paths[0] );
}

/* workers/LocalWorker.cpp:1119-1164 (initPhaseRWOffsetGen) + the per-iterator overrides of
 * :3494-3513. Seeds: randOffsetSeed expands to a xoshiro256** state through splitmix64 and to the
 * full-coverage LCG seed (the reference self-seeds both). */
static void orc_init_offset_gen(orc_worker* w, int isWritePhase)
{
	orc_shared* sh = w->shared;
	const elb_cfg* cfg = sh->cfg;
	const uint64_t blockSize = sh->blockSize;
	const uint64_t fileSize = sh->fileSize;
	const uint32_t numDataSetThreads = sh->numDataSetThreads;
	const int isDir = (cfg->pathType == ELB_PATH_DIR);

	uint64_t seedBase = cfg->randOffsetSeed ? cfg->randOffsetSeed : (uint64_t)time(NULL);
	uint64_t randState[4];
	uint64_t sm = seedBase + w->rank * 0x9E3779B97F4A7C15ULL;
	for(int i = 0; i < 4; i++)
	{
		sm += 0x9E3779B97F4A7C15ULL;
		randState[i] = orc_splitmix64_mix(sm);
	}
	uint64_t lcgSeed = randState[0] ^ randState[1];

	if(w->offsetGen)
		orc_offsetgen_destroy(w->offsetGen);

	if(!isDir && (cfg->useRandomOffsets || cfg->useStridedAccess) )
	{ // fileModeIterateFilesRand (:3486-3513)
		const uint64_t numBlocksPerFile = fileSize / blockSize;
		const uint64_t numBlocksTotal = numBlocksPerFile * cfg->numPaths;
		const uint64_t randomAmount = sh->randomAmount / numDataSetThreads;
		const uint64_t rangeLen = blockSize * (numBlocksTotal / numDataSetThreads);
		const uint64_t rangeOffset = w->rank * blockSize * (numBlocksTotal / numDataSetThreads);

		if(cfg->useStridedAccess)
			w->offsetGen = orc_offsetgen_create(ORC_OFFGEN_STRIDED, rangeLen, rangeLen,
				blockSize * w->rank, blockSize, numDataSetThreads, randState, lcgSeed);
		else
		if(cfg->useRandomUnaligned)
			w->offsetGen = orc_offsetgen_create_algo(ORC_OFFGEN_RANDOM, randomAmount, rangeLen,
				rangeOffset, blockSize, numDataSetThreads, cfg->randOffsetAlgo, randState, lcgSeed);
		else
		if(cfg->useExplicitRandOffsetAlgo || !isWritePhase)
			w->offsetGen = orc_offsetgen_create_algo(ORC_OFFGEN_RANDOM_ALIGNED, randomAmount,
				rangeLen, rangeOffset, blockSize, numDataSetThreads, cfg->randOffsetAlgo, randState,
				lcgSeed);
		else
			w->offsetGen = orc_offsetgen_create(ORC_OFFGEN_RANDOM_ALIGNED_FULLCOV, randomAmount,
				rangeLen, rangeOffset, blockSize, numDataSetThreads, randState, lcgSeed);

		return;
	}

	// :1129-1163 (dir mode: randomAmount = fileSize)
	const uint64_t randomAmount = isDir ? fileSize : (sh->randomAmount / numDataSetThreads);

	if(cfg->doReverseSeqOffsets)
		w->offsetGen = orc_offsetgen_create(ORC_OFFGEN_REVERSE_SEQ, fileSize, fileSize, 0,
			blockSize, numDataSetThreads, randState, lcgSeed);
	else
	if(!cfg->useRandomOffsets && !cfg->useStridedAccess)
		w->offsetGen = orc_offsetgen_create(ORC_OFFGEN_SEQUENTIAL, fileSize, fileSize, 0,
			blockSize, numDataSetThreads, randState, lcgSeed);
	else
	if(cfg->useRandomUnaligned)
		w->offsetGen = orc_offsetgen_create(ORC_OFFGEN_RANDOM, randomAmount, fileSize, 0,
			blockSize, numDataSetThreads, randState, lcgSeed);
	else
	if(cfg->useExplicitRandOffsetAlgo || !isWritePhase)
		w->offsetGen = orc_offsetgen_create(ORC_OFFGEN_RANDOM_ALIGNED, randomAmount, fileSize, 0,
			blockSize, numDataSetThreads, randState, lcgSeed);
	else
		w->offsetGen = orc_offsetgen_create(ORC_OFFGEN_RANDOM_ALIGNED_FULLCOV, randomAmount,
			fileSize, 0, blockSize, numDataSetThreads, randState, lcgSeed);
}

/* workers/LocalWorker.cpp:7062-7082 (getDirModeOpenFlags) */
static int orc_dir_mode_open_flags(const elb_cfg* cfg, int benchPhase)
{
	int openFlags;

	if(benchPhase == ELB_PHASE_CREATEFILES)
	{
		openFlags = O_CREAT | O_RDWR;
		if(cfg->doTruncate)
			openFlags |= O_TRUNC;
	}
	else
		openFlags = O_RDONLY;

	if(cfg->useDirectIO)
		openFlags |= O_DIRECT;

	return openFlags;
}

/* workers/LocalWorker.cpp:3022-3248 (dirModeIterateFiles) incl. :7097-7161 (open + prep) */
static void orc_dir_mode_iterate_files(orc_worker* w, int benchPhase)
{
	orc_shared* sh = w->shared;
	const elb_cfg* cfg = sh->cfg;
	const int haveSubdirs = (cfg->numDirs > 0);
	const size_t numDirs = haveSubdirs ? cfg->numDirs : 1;
	const size_t numFiles = cfg->numFiles;
	const uint64_t fileSize = sh->fileSize;
	const int openFlags = orc_dir_mode_open_flags(cfg, benchPhase);
	const size_t workerDirRank = cfg->doDirSharing ? 0 : w->rank;
	char currentPath[ORC_PATH_BUF_LEN];
	char msg[512];

	for(size_t dirIndex = 0; dirIndex < numDirs; dirIndex++)
	{
		for(size_t fileIndex = 0; fileIndex < numFiles; fileIndex++)
		{
			int printRes;

			if(haveSubdirs)
				printRes = snprintf(currentPath, ORC_PATH_BUF_LEN, "r%zu/d%zu/r%zu-f%zu",
					workerDirRank, dirIndex, (size_t)w->rank, fileIndex);
			else
				printRes = snprintf(currentPath, ORC_PATH_BUF_LEN, "r%zu-f%zu",
					(size_t)w->rank, fileIndex);

			if(printRes >= ORC_PATH_BUF_LEN)
			{
				orc_worker_fail(w, "file path too long for static buffer.");
				return;
			}

			unsigned pathFDsIndex = (w->rank + dirIndex) % cfg->numPaths;

			orc_offsetgen_reset(w->offsetGen);

			struct timespec ioStartT;
			clock_gettime(CLOCK_MONOTONIC, &ioStartT);

			if( (benchPhase == ELB_PHASE_CREATEFILES) || (benchPhase == ELB_PHASE_READFILES) )
			{
				const int isRead = (benchPhase == ELB_PHASE_READFILES);

				int fd = openat(sh->pathFDs[pathFDsIndex], currentPath, openFlags,
					ORC_MKFILE_MODE);
				if(fd == -1)
				{
					snprintf(msg, sizeof(msg), "File open failed. Path: %s/%s; SysErr: %s",
						cfg->paths[pathFDsIndex], currentPath, strerror(errno) );
					orc_worker_fail(w, msg);
					return;
				}

				if(!isRead && cfg->doTruncToSize && (ftruncate(fd, fileSize) == -1) )
				{
					orc_worker_fail(w, "Unable to set file size through ftruncate.");
					close(fd);
					return;
				}

				if(!isRead && cfg->doPreallocFile && posix_fallocate(fd, 0, fileSize) )
				{
					orc_worker_fail(w, "Unable to preallocate file size through posix_fallocate.");
					close(fd);
					return;
				}

				w->currentFD = fd;

				int64_t rwRes = orc_rw_any_block_sized(w, &w->currentFD, 1, isRead);

				if( (rwRes < 0) || ( (uint64_t)rwRes != fileSize) )
				{
					snprintf(msg, sizeof(msg), "%s/%s", cfg->paths[pathFDsIndex], currentPath);
					orc_set_io_error(w, isRead, rwRes, fileSize, msg);
					close(fd);
					return;
				}

				if(close(fd) == -1)
				{
					orc_worker_fail(w, "File close failed.");
					return;
				}
			}

			if(benchPhase == ELB_PHASE_STATFILES)
			{
				struct stat statBuf;

				if(fstatat(sh->pathFDs[pathFDsIndex], currentPath, &statBuf, 0) == -1)
				{
					orc_worker_fail(w, "File stat failed.");
					return;
				}
			}

			if(benchPhase == ELB_PHASE_DELETEFILES)
			{
				int unlinkRes = unlinkat(sh->pathFDs[pathFDsIndex], currentPath, 0);

				if( (unlinkRes == -1) && (!cfg->ignoreDelErrors || (errno != ENOENT) ) )
				{
					snprintf(msg, sizeof(msg), "File delete failed. Path: %s/%s; SysErr: %s",
						cfg->paths[pathFDsIndex], currentPath, strerror(errno) );
					orc_worker_fail(w, msg);
					return;
				}
			}

			if(w->isRWMixReader)
			{ // (:3233-3237)
				__atomic_fetch_add(&w->res->liveOpsReadMix.numEntriesDone, 1, __ATOMIC_RELAXED);
				continue;
			}

			orc_histogram_add_latency(&w->res->entriesLatHisto, orc_elapsed_usec(&ioStartT) );
			__atomic_fetch_add(&w->res->liveOps.numEntriesDone, 1, __ATOMIC_RELAXED);
		}
	}
}

/* workers/LocalWorker.cpp:2778-2912 (dirModeIterateDirs) */
static void orc_dir_mode_iterate_dirs(orc_worker* w, int benchPhase)
{
	orc_shared* sh = w->shared;
	const elb_cfg* cfg = sh->cfg;
	const size_t numDirs = cfg->numDirs;
	const int ignoreDelErrors = cfg->doDirSharing ? 1 : cfg->ignoreDelErrors;
	const size_t workerDirRank = cfg->doDirSharing ? 0 : w->rank;
	char currentPath[ORC_PATH_BUF_LEN];
	char msg[512];

	if(!numDirs)
		return;

	if(benchPhase == ELB_PHASE_CREATEDIRS)
	{
		for(unsigned pathFDsIndex = 0; pathFDsIndex < cfg->numPaths; pathFDsIndex++)
		{
			snprintf(currentPath, ORC_PATH_BUF_LEN, "r%zu", workerDirRank);

			int mkdirRes = mkdirat(sh->pathFDs[pathFDsIndex], currentPath, ORC_MKDIR_MODE);

			if( (mkdirRes == -1) && (errno != EEXIST) )
			{
				snprintf(msg, sizeof(msg), "Rank directory creation failed. Path: %s/%s; "
					"SysErr: %s", cfg->paths[pathFDsIndex], currentPath, strerror(errno) );
				orc_worker_fail(w, msg);
				return;
			}
		}
	}

	for(size_t dirIndex = 0; dirIndex < numDirs; dirIndex++)
	{
		snprintf(currentPath, ORC_PATH_BUF_LEN, "r%zu/d%zu", workerDirRank, dirIndex);

		unsigned pathFDsIndex = (w->rank + dirIndex) % cfg->numPaths;

		struct timespec ioStartT;
		clock_gettime(CLOCK_MONOTONIC, &ioStartT);

		if(benchPhase == ELB_PHASE_CREATEDIRS)
		{
			int mkdirRes = mkdirat(sh->pathFDs[pathFDsIndex], currentPath, ORC_MKDIR_MODE);

			if( (mkdirRes == -1) && (errno != EEXIST) )
			{
				snprintf(msg, sizeof(msg), "Directory creation failed. Path: %s/%s; SysErr: %s",
					cfg->paths[pathFDsIndex], currentPath, strerror(errno) );
				orc_worker_fail(w, msg);
				return;
			}
		}

		if(benchPhase == ELB_PHASE_DELETEDIRS)
		{
			int rmdirRes = unlinkat(sh->pathFDs[pathFDsIndex], currentPath, AT_REMOVEDIR);

			if( (rmdirRes == -1) && ( (errno != ENOENT) || !ignoreDelErrors) )
			{
				snprintf(msg, sizeof(msg), "Directory deletion failed. Path: %s/%s; SysErr: %s",
					cfg->paths[pathFDsIndex], currentPath, strerror(errno) );
				orc_worker_fail(w, msg);
				return;
			}
		}

		orc_histogram_add_latency(&w->res->entriesLatHisto, orc_elapsed_usec(&ioStartT) );
		__atomic_fetch_add(&w->res->liveOps.numEntriesDone, 1, __ATOMIC_RELAXED);
	}

	if(benchPhase == ELB_PHASE_DELETEDIRS)
	{
		for(unsigned pathFDsIndex = 0; pathFDsIndex < cfg->numPaths; pathFDsIndex++)
		{
			snprintf(currentPath, ORC_PATH_BUF_LEN, "r%zu", workerDirRank);

			int rmdirRes = unlinkat(sh->pathFDs[pathFDsIndex], currentPath, AT_REMOVEDIR);

			if( (rmdirRes == -1) && ( (errno != ENOENT) || !ignoreDelErrors) )
			{
				snprintf(msg, sizeof(msg), "Directory deletion failed. Path: %s/%s; SysErr: %s",
					cfg->paths[pathFDsIndex], currentPath, strerror(errno) );
				orc_worker_fail(w, msg);
				return;
			}
		}
	}
}

/* workers/LocalWorker.cpp:3736-3767 (fileModeDeleteFiles): every worker tries to delete every
 * file, starting at a rank-dependent index; ENOENT is ignored. */
static void orc_file_mode_delete_files(orc_worker* w)
{
	const elb_cfg* cfg = w->shared->cfg;
	const size_t numFiles = cfg->numPaths;
	char msg[512];

	for(size_t fileIndex = 0; fileIndex < numFiles; fileIndex++)
	{
		const char* path = cfg->paths[ (w->rank + fileIndex) % numFiles];

		int unlinkRes = unlink(path);

		if( (unlinkRes == -1) && (errno != ENOENT) )
		{
			snprintf(msg, sizeof(msg), "File delete failed. Path: %s; SysErr: %s", path,
				strerror(errno) );
			orc_worker_fail(w, msg);
			return;
		}

		__atomic_fetch_add(&w->res->liveOps.numEntriesDone, 1, __ATOMIC_RELAXED);
	}
}

/* Worker.cpp:33-55 (incNumWorkersDone): the first finisher that had work snapshots every worker's
 * live ops as the stonewall ("first done") result. */
static void orc_inc_num_workers_done(orc_worker* w)
{
	orc_shared* sh = w->shared;
	const uint32_t numWorkersTotal = sh->cfg->numThreads;

	pthread_mutex_lock(&sh->mutex);

	int lastFinisherTrigger = sh->cfg->runAsService ?
		0 : ( (sh->numWorkersDone + 1) == (int)numWorkersTotal);
	int triggerStoneWall = (!sh->stoneWallTriggered &&
		(w->res->gotPhaseWork || lastFinisherTrigger) );

	sh->numWorkersDone++;

	if(triggerStoneWall)
	{
		sh->stoneWallTriggered = 1;

		for(uint32_t i = 0; i < numWorkersTotal; i++)
		{
			orc_worker* other = &sh->workers[i];
			other->stoneWallOps.numEntriesDone =
				__atomic_load_n(&other->res->liveOps.numEntriesDone, __ATOMIC_RELAXED);
			other->stoneWallOps.numBytesDone =
				__atomic_load_n(&other->res->liveOps.numBytesDone, __ATOMIC_RELAXED);
			other->stoneWallOps.numIOPSDone =
				__atomic_load_n(&other->res->liveOps.numIOPSDone, __ATOMIC_RELAXED);
			other->stoneWallOpsReadMix.numEntriesDone =
				__atomic_load_n(&other->res->liveOpsReadMix.numEntriesDone, __ATOMIC_RELAXED);
			other->stoneWallOpsReadMix.numBytesDone =
				__atomic_load_n(&other->res->liveOpsReadMix.numBytesDone, __ATOMIC_RELAXED);
			other->stoneWallOpsReadMix.numIOPSDone =
				__atomic_load_n(&other->res->liveOpsReadMix.numIOPSDone, __ATOMIC_RELAXED);
		}
	}

	pthread_mutex_unlock(&sh->mutex);
}

/* workers/LocalWorker.cpp:177-396 (run), one phase */
static void* orc_worker_thread(void* arg)
{
	orc_worker* w = (orc_worker*)arg;
	orc_shared* sh = w->shared;
	const elb_cfg* cfg = sh->cfg;
	const int benchPhase = sh->benchPhase;
	const int isDir = (cfg->pathType == ELB_PATH_DIR);

	w->res->gotPhaseWork = 1;

	switch(benchPhase)
	{
		case ELB_PHASE_CREATEDIRS:
		case ELB_PHASE_DELETEDIRS:
			orc_dir_mode_iterate_dirs(w, benchPhase);
			break;

		case ELB_PHASE_CREATEFILES:
		case ELB_PHASE_READFILES:
		{
			// initThreadPhaseVars (:1028-1041): rwmix reader threads read in the write phase
			w->isRWMixReader = (benchPhase == ELB_PHASE_CREATEFILES) &&
				( (w->rank - cfg->rankOffset) < cfg->numRWMixReadThreads);

			const int isRead = (benchPhase == ELB_PHASE_READFILES) || w->isRWMixReader;

			orc_init_offset_gen(w, benchPhase == ELB_PHASE_CREATEFILES);

			// per-thread rate limit (:1293-1299 write side, :1331-1337 read side incl. rwmix readers)
			orc_ratelimiter_init_start(&w->rateLimiter,
				isRead ? cfg->limitReadBps : cfg->limitWriteBps);

			if(isDir)
				orc_dir_mode_iterate_files(w, isRead ? ELB_PHASE_READFILES : benchPhase);
			else
			if(cfg->useRandomOffsets || cfg->useStridedAccess)
				orc_file_mode_iterate_rand(w, isRead);
			else
				orc_file_mode_iterate_seq(w, isRead);
		} break;

		case ELB_PHASE_STATFILES:
			if(isDir)
			{
				orc_init_offset_gen(w, 0);
				orc_dir_mode_iterate_files(w, benchPhase);
			}
			break;

		case ELB_PHASE_DELETEFILES:
			if(isDir)
			{
				orc_init_offset_gen(w, 0);
				orc_dir_mode_iterate_files(w, benchPhase);
			}
			else
				orc_file_mode_delete_files(w);
			break;

		default:
			break;
	}

	// finishPhase (:433-453)
	if(!w->res->hadError && w->res->gotPhaseWork)
		w->res->elapsedUSec = orc_elapsed_usec(&sh->phaseStartT);
	else
		w->res->elapsedUSec = 0;

	orc_inc_num_workers_done(w);

	return NULL;
}

void orc_expected_per_worker(const elb_cfg* cfg, int benchPhase, uint64_t* outEntries,
	uint64_t* outBytes)
{
	orc_shared sh;
	memset(&sh, 0, sizeof(sh) );
	sh.cfg = cfg;
	orc_normalize(&sh);

	*outEntries = 0;
	*outBytes = 0;

	if(cfg->pathType == ELB_PATH_DIR)
	{ // workers/WorkerManager.cpp:354-400
		const uint64_t numDirs = cfg->numDirs ? cfg->numDirs : 1;

		switch(benchPhase)
		{
			case ELB_PHASE_CREATEDIRS:
			case ELB_PHASE_DELETEDIRS:
				*outEntries = cfg->numDirs;
				break;
			case ELB_PHASE_CREATEFILES:
			case ELB_PHASE_READFILES:
				*outEntries = numDirs * cfg->numFiles;
				*outBytes = *outEntries * sh.fileSize;
				break;
			case ELB_PHASE_DELETEFILES:
			case ELB_PHASE_STATFILES:
				*outEntries = numDirs * cfg->numFiles;
				break;
			default:
				break;
		}
	}
	else
	{ // :455-478
		*outEntries = cfg->numPaths;

		if( (benchPhase == ELB_PHASE_CREATEFILES) || (benchPhase == ELB_PHASE_READFILES) )
			*outBytes = cfg->useRandomOffsets ?
				(sh.randomAmount / sh.numDataSetThreads) :
				( (*outEntries * sh.fileSize) / sh.numDataSetThreads);
	}
}

int orc_run_phase(const elb_cfg* cfg, int benchPhase, orc_worker_result* results,
	elb_phase_results* outPhaseResults)
{
	orc_shared sh;
	memset(&sh, 0, sizeof(sh) );
	sh.cfg = cfg;
	sh.benchPhase = benchPhase;
	orc_normalize(&sh);
	pthread_mutex_init(&sh.mutex, NULL);

	const uint32_t numThreads = cfg->numThreads;
	const int isDir = (cfg->pathType == ELB_PATH_DIR);
	int retVal = 0;

	// ProgArgs::prepareBenchPathFDsVec (ProgArgs.cpp:1859-1935) + prepareFileSize
	sh.pathFDs = (int*)malloc(sizeof(int) * cfg->numPaths);
	for(uint32_t i = 0; i < cfg->numPaths; i++)
		sh.pathFDs[i] = -1;

	for(uint32_t i = 0; i < cfg->numPaths; i++)
	{
		int openFlags = 0;

		if(isDir)
			openFlags |= (O_DIRECTORY | O_RDONLY);
		else
		{
			openFlags |= (benchPhase == ELB_PHASE_READFILES) ? O_RDONLY : O_RDWR;

			if(cfg->useDirectIO)
				openFlags |= O_DIRECT;

			if(benchPhase == ELB_PHASE_CREATEFILES)
				openFlags |= O_CREAT;
		}

		if(!isDir && (benchPhase == ELB_PHASE_DELETEFILES) )
			continue; // nothing to open for unlink by path

		sh.pathFDs[i] = open(cfg->paths[i], openFlags, ORC_MKFILE_MODE);

		if(sh.pathFDs[i] == -1)
		{
			for(uint32_t r = 0; r < numThreads; r++)
			{
				results[r].hadError = 1;
				snprintf(results[r].errorMsg, sizeof(results[r].errorMsg),
					"Unable to open benchmark path: %s; SysErr: %s", cfg->paths[i],
					strerror(errno) );
			}
			retVal = -1;
			goto cleanup;
		}

		if(!isDir && (benchPhase == ELB_PHASE_CREATEFILES) )
		{
			if(cfg->doTruncate && (ftruncate(sh.pathFDs[i], 0) == -1) )
				retVal = -1;
			if(cfg->doTruncToSize && (ftruncate(sh.pathFDs[i], sh.fileSize) == -1) )
				retVal = -1;
			if(cfg->doPreallocFile && posix_fallocate(sh.pathFDs[i], 0, sh.fileSize) )
				retVal = -1;
		}
	}

	orc_worker* workers = (orc_worker*)calloc(numThreads, sizeof(orc_worker) );
	pthread_t* threads = (pthread_t*)calloc(numThreads, sizeof(pthread_t) );
	sh.workers = workers;

	for(uint32_t i = 0; i < numThreads; i++)
	{
		orc_worker* w = &workers[i];
		w->shared = &sh;
		w->rank = cfg->rankOffset + i;
		w->res = &results[i];
		memset(w->res, 0, sizeof(*w->res) );
		orc_histogram_reset(&w->res->iopsLatHisto);
		orc_histogram_reset(&w->res->entriesLatHisto);

		// allocIOBuffer (workers/LocalWorker.cpp:1362-1396): page aligned, random prefill
		if(posix_memalign( (void**)&w->ioBuf, sysconf(_SC_PAGESIZE),
			sh.blockSize ? sh.blockSize : 1) )
		{
			retVal = -1;
			w->ioBuf = NULL;
		}

		/* allocIOBuffer for iodepth > 1 (:1362-1396: one buffer per depth entry) + initLibAio
		   (:455-480) */
		if( (cfg->ioDepth > 1) && w->ioBuf && sh.blockSize)
		{
			w->ioBufVec = (char**)calloc(cfg->ioDepth, sizeof(char*) );
			w->iocbVec = (struct iocb*)calloc(cfg->ioDepth, sizeof(struct iocb) );
			w->ioStartTimeVec = (struct timespec*)calloc(cfg->ioDepth, sizeof(struct timespec) );
			w->ioBufVec[0] = w->ioBuf;

			for(uint32_t d = 1; d < cfg->ioDepth; d++)
				if(posix_memalign( (void**)&w->ioBufVec[d], sysconf(_SC_PAGESIZE), sh.blockSize) )
					retVal = -1;

			if(!retVal && (syscall(SYS_io_setup, (unsigned)cfg->ioDepth, &w->aioContext) == 0) )
				w->aioInitialized = 1;
			else
				retVal = -1;
		}

		uint64_t seederState[4];
		uint64_t sm = (cfg->blockVarianceSeed ? cfg->blockVarianceSeed : 0x1234567ULL) + w->rank;
		for(int k = 0; k < 4; k++)
		{
			sm += 0x9E3779B97F4A7C15ULL;
			seederState[k] = orc_splitmix64_mix(sm);
		}
		orc_goldenprime_init(&w->blockVarAlgo, seederState[0] | 1, seederState);

		if(w->ioBuf)
			orc_xoshiro256ss_fill_buf(&w->blockVarAlgo.stateSeeder, w->ioBuf, sh.blockSize);
	}

	if(retVal)
		goto cleanup_workers;

	clock_gettime(CLOCK_MONOTONIC, &sh.phaseStartT); // WorkerManager.cpp:311

	for(uint32_t i = 0; i < numThreads; i++)
		pthread_create(&threads[i], NULL, orc_worker_thread, &workers[i] );

	for(uint32_t i = 0; i < numThreads; i++)
		pthread_join(threads[i], NULL);

	for(uint32_t i = 0; i < numThreads; i++)
		if(results[i].hadError)
			retVal = -1;

	// Statistics::generatePhaseResults (Statistics.cpp:1641-1764)
	if(outPhaseResults)
	{
		elb_phase_results* pr = outPhaseResults;
		memset(pr, 0, sizeof(*pr) );
		orc_histogram_reset(&pr->iopsLatHisto);
		orc_histogram_reset(&pr->entriesLatHisto);
		pr->firstFinishUSec = ~0ULL;

		for(uint32_t i = 0; i < numThreads; i++)
		{
			orc_worker_result* r = &results[i];

			if(r->elapsedUSec)
			{
				if(r->elapsedUSec < pr->firstFinishUSec)
					pr->firstFinishUSec = r->elapsedUSec;
				if(r->elapsedUSec > pr->lastFinishUSec)
					pr->lastFinishUSec = r->elapsedUSec;
			}

			pr->opsTotal.numEntriesDone += r->liveOps.numEntriesDone;
			pr->opsTotal.numBytesDone += r->liveOps.numBytesDone;
			pr->opsTotal.numIOPSDone += r->liveOps.numIOPSDone;
			pr->opsReadMixTotal.numEntriesDone += r->liveOpsReadMix.numEntriesDone;
			pr->opsReadMixTotal.numBytesDone += r->liveOpsReadMix.numBytesDone;
			pr->opsReadMixTotal.numIOPSDone += r->liveOpsReadMix.numIOPSDone;
			pr->opsStoneWallTotal.numEntriesDone += workers[i].stoneWallOps.numEntriesDone;
			pr->opsStoneWallTotal.numBytesDone += workers[i].stoneWallOps.numBytesDone;
			pr->opsStoneWallTotal.numIOPSDone += workers[i].stoneWallOps.numIOPSDone;
			// (Statistics.cpp:1697: opsStoneWallTotalReadMix)
			pr->opsStoneWallReadMixTotal.numEntriesDone +=
				workers[i].stoneWallOpsReadMix.numEntriesDone;
			pr->opsStoneWallReadMixTotal.numBytesDone += workers[i].stoneWallOpsReadMix.numBytesDone;
			pr->opsStoneWallReadMixTotal.numIOPSDone += workers[i].stoneWallOpsReadMix.numIOPSDone;
			orc_histogram_merge(&pr->iopsLatHisto, &r->iopsLatHisto);
			orc_histogram_merge(&pr->entriesLatHisto, &r->entriesLatHisto);

			pr->numWorkersDone++;
			if(r->hadError)
				pr->numWorkersDoneWithError++;
		}

		if(pr->firstFinishUSec == ~0ULL)
			pr->firstFinishUSec = 0;

		if(pr->lastFinishUSec)
		{
			pr->opsPerSec.numEntriesDone =
				orc_per_sec_from_usec(pr->opsTotal.numEntriesDone, pr->lastFinishUSec);
			pr->opsPerSec.numBytesDone =
				orc_per_sec_from_usec(pr->opsTotal.numBytesDone, pr->lastFinishUSec);
			pr->opsPerSec.numIOPSDone =
				orc_per_sec_from_usec(pr->opsTotal.numIOPSDone, pr->lastFinishUSec);
		}

		if(pr->firstFinishUSec)
		{
			pr->opsStoneWallPerSec.numEntriesDone = orc_per_sec_from_usec(
				pr->opsStoneWallTotal.numEntriesDone, pr->firstFinishUSec);
			pr->opsStoneWallPerSec.numBytesDone = orc_per_sec_from_usec(
				pr->opsStoneWallTotal.numBytesDone, pr->firstFinishUSec);
			pr->opsStoneWallPerSec.numIOPSDone = orc_per_sec_from_usec(
				pr->opsStoneWallTotal.numIOPSDone, pr->firstFinishUSec);
			// (Statistics.cpp:1721-1722)
			pr->opsStoneWallReadMixPerSec.numEntriesDone = orc_per_sec_from_usec(
				pr->opsStoneWallReadMixTotal.numEntriesDone, pr->firstFinishUSec);
			pr->opsStoneWallReadMixPerSec.numBytesDone = orc_per_sec_from_usec(
				pr->opsStoneWallReadMixTotal.numBytesDone, pr->firstFinishUSec);
			pr->opsStoneWallReadMixPerSec.numIOPSDone = orc_per_sec_from_usec(
				pr->opsStoneWallReadMixTotal.numIOPSDone, pr->firstFinishUSec);
		}
	}

cleanup_workers:
	for(uint32_t i = 0; i < numThreads; i++)
	{
		if(workers[i].aioInitialized)
			syscall(SYS_io_destroy, workers[i].aioContext);
		for(uint32_t d = 1; workers[i].ioBufVec && (d < cfg->ioDepth); d++)
			free(workers[i].ioBufVec[d] );
		free(workers[i].ioBufVec);
		free(workers[i].iocbVec);
		free(workers[i].ioStartTimeVec);
		free(workers[i].ioBuf);
		if(workers[i].offsetGen)
			orc_offsetgen_destroy(workers[i].offsetGen);
	}

	free(workers);
	free(threads);

cleanup:
	for(uint32_t i = 0; i < cfg->numPaths; i++)
		if(sh.pathFDs[i] != -1)
			close(sh.pathFDs[i] );

	free(sh.pathFDs);
	pthread_mutex_destroy(&sh.mutex);

	return retVal;
}

/* ============================================================================================
 * Micro-benchmarks of the CPU block modifiers (one thread)
 * ========================================================================================== */

static double orc_now_sec(void)
{
	struct timespec now;
	clock_gettime(CLOCK_MONOTONIC, &now);
	return now.tv_sec + now.tv_nsec * 1e-9;
}

double orc_bench_fill_pattern(size_t blockSize, size_t numBlocks)
{
	char* buf = NULL;
	if(posix_memalign( (void**)&buf, 4096, blockSize) )
		return 0;

	double startT = orc_now_sec();

	for(size_t i = 0; i < numBlocks; i++)
	{
		orc_fill_pattern(buf, blockSize, (uint64_t)i * blockSize, 1);
		__asm__ volatile("" : : "r"(buf) : "memory");
	}

	double elapsed = orc_now_sec() - startT;
	free(buf);

	return ( (double)blockSize * numBlocks) / elapsed;
}

double orc_bench_verify_pattern(size_t blockSize, size_t numBlocks)
{
	char* buf = NULL;
	if(posix_memalign( (void**)&buf, 4096, blockSize) )
		return 0;

	orc_fill_pattern(buf, blockSize, 0, 1);

	double startT = orc_now_sec();
	int numBad = 0;

	for(size_t i = 0; i < numBlocks; i++)
	{
		numBad += orc_verify_pattern(buf, blockSize, 0, 1, NULL, NULL, NULL, NULL, NULL, 0);
		__asm__ volatile("" : : "r"(buf) : "memory");
	}

	double elapsed = orc_now_sec() - startT;
	free(buf);

	return numBad ? 0 : ( ( (double)blockSize * numBlocks) / elapsed);
}
