/*
 * TEST INFRASTRUCTURE ONLY — C ABI around the REFERENCE's own PRNG and offset generator headers,
 * compiled from /root/reference/source where they lie (nothing is copied into this repo) into
 * oracle/_ref/libelb_ref.so by oracle/Makefile. Used to pin the oracle restatement
 * (tests/test_oracle_vs_ref.py) and to generate tests/golden/*.json (tests/golden/make_golden.py).
 *
 * The reference self-seeds its generators from std::random_device and keeps the state private;
 * the `private`/`protected` -> `public` redefinition below is how this harness injects a known
 * state. It changes access control only, not behaviour.
 */
#include <cstdint>
#include <cstring>
#include <memory>
#include <random>
#include <sstream>
#include <string>

#define private public
#define protected public
#include "workers/WorkerException.h"
#include "toolkits/random/RandAlgoSelectorTk.h"
#include "toolkits/random/RandAlgoRange.h"
#include "toolkits/random/RandAlgoXoshiro256ppSIMD.h"
#include "toolkits/offsetgen/OffsetGenerator.h"
#include "toolkits/offsetgen/OffsetGenRandomAlignedFullCoverageV2.h"
#undef private
#undef protected

extern "C" {

/* ---- PRNGs ---- */

void* ref_xoshiro256ss_create(const uint64_t state[4])
{
	RandAlgoXoshiro256ss* algo = new RandAlgoXoshiro256ss();
	memcpy(algo->state.s, state, sizeof(algo->state.s) );
	return algo;
}

void* ref_goldenprime_create(uint64_t seed, const uint64_t seederState[4])
{
	RandAlgoGoldenPrime* algo = new RandAlgoGoldenPrime(seed);
	memcpy(algo->stateSeeder.state.s, seederState, sizeof(algo->stateSeeder.state.s) );
	return algo;
}

/* algo values = enum elb_offset_rand_algo; the object is what RandAlgoSelectorTk::stringToAlgo
 * returns for "balanced_single" / "fast" / "balanced" / "strong", with an injected state */
void* ref_randalgo_create(int algoType, const uint64_t state[4])
{
	switch(algoType)
	{
		case 0:
			return ref_xoshiro256ss_create(state);

		case 1:
			return ref_goldenprime_create(state[0], state);

		case 2:
		{ // (the selector instantiates the default template argument, RandAlgoSelectorTk.cpp:49)
			RandAlgoXoshiro256ppSIMD<>* algo = new RandAlgoXoshiro256ppSIMD<>();
			for(int i = 0; i < 4; i++)
				algo->state.s[i][0] = state[i]; // next() advances lane 0 only
			return algo;
		}

		case 3:
		{
			RandAlgoMT19937* algo = new RandAlgoMT19937();
			algo->randGen.seed(state[0] );
			return algo;
		}

		default:
			return NULL;
	}
}

uint64_t ref_randalgo_next(void* algo)
{
	return ( (RandAlgoInterface*)algo)->next();
}

void ref_randalgo_fill_buf(void* algo, char* buf, uint64_t bufLen)
{
	( (RandAlgoInterface*)algo)->fillBuf(buf, bufLen);
}

void ref_randalgo_destroy(void* algo)
{
	delete (RandAlgoInterface*)algo;
}

/* ---- offset generators ---- */

struct RefOffsetGen
{
	std::unique_ptr<RandAlgoInterface> randAlgo;
	std::unique_ptr<OffsetGenerator> gen;
};

void* ref_offsetgen_create_algo(int kind, uint64_t numBytesTotal, uint64_t len, uint64_t offset,
	uint64_t blockSize, uint64_t numDataSetThreads, int algoType, const uint64_t randState[4],
	uint64_t lcgState);

/* kind values = enum orc_offsetgen_kind of oracle/elb_oracle.h */
void* ref_offsetgen_create(int kind, uint64_t numBytesTotal, uint64_t len, uint64_t offset,
	uint64_t blockSize, uint64_t numDataSetThreads, const uint64_t randState[4],
	uint64_t lcgState)
{
	return ref_offsetgen_create_algo(kind, numBytesTotal, len, offset, blockSize,
		numDataSetThreads, 0, randState, lcgState);
}

void* ref_offsetgen_create_algo(int kind, uint64_t numBytesTotal, uint64_t len, uint64_t offset,
	uint64_t blockSize, uint64_t numDataSetThreads, int algoType, const uint64_t randState[4],
	uint64_t lcgState)
{
	static const uint64_t zeroState[4] = {0, 0, 0, 0};

	RefOffsetGen* ref = new RefOffsetGen();
	ref->randAlgo.reset( (RandAlgoInterface*)ref_randalgo_create(algoType,
		randState ? randState : zeroState) );

	if(!ref->randAlgo)
	{
		delete ref;
		return NULL;
	}

	switch(kind)
	{
		case 0: ref->gen.reset(new OffsetGenSequential(len, offset, blockSize) ); break;
		case 1: ref->gen.reset(new OffsetGenReverseSeq(len, offset, blockSize) ); break;
		case 2: ref->gen.reset(new OffsetGenRandom(numBytesTotal, *ref->randAlgo, len, offset,
			blockSize) ); break;
		case 3: ref->gen.reset(new OffsetGenRandomAligned(numBytesTotal, *ref->randAlgo, len,
			offset, blockSize) ); break;
		case 4: ref->gen.reset(new OffsetGenStrided(len, offset, blockSize,
			numDataSetThreads) ); break;
		case 5:
		{
			OffsetGenRandomAlignedFullCoverageV2* cov = new OffsetGenRandomAlignedFullCoverageV2(
				numBytesTotal, len, offset, blockSize);
			// inject the LCG start state (reference: random_device() % m, FullCoverageV2.h:93-99)
			cov->randomGen.current_lcg_state_ = lcgState % cov->randomGen.m_lcg_;
			ref->gen.reset(cov);
		} break;
		default:
			delete ref;
			return NULL;
	}

	return ref;
}

void ref_offsetgen_destroy(void* g) { delete (RefOffsetGen*)g; }
void ref_offsetgen_reset(void* g) { ( (RefOffsetGen*)g)->gen->reset(); }
void ref_offsetgen_reset_range(void* g, uint64_t len, uint64_t offset)
	{ ( (RefOffsetGen*)g)->gen->reset(len, offset); }
uint64_t ref_offsetgen_next_offset(void* g) { return ( (RefOffsetGen*)g)->gen->getNextOffset(); }
uint64_t ref_offsetgen_next_block_size(void* g)
	{ return ( (RefOffsetGen*)g)->gen->getNextBlockSizeToSubmit(); }
uint64_t ref_offsetgen_bytes_total(void* g) { return ( (RefOffsetGen*)g)->gen->getNumBytesTotal(); }
uint64_t ref_offsetgen_bytes_left(void* g)
	{ return ( (RefOffsetGen*)g)->gen->getNumBytesLeftToSubmit(); }
void ref_offsetgen_add_bytes_submitted(void* g, uint64_t numBytes)
	{ ( (RefOffsetGen*)g)->gen->addBytesSubmitted(numBytes); }

/* set the LCG state of a full-coverage generator (after reset(len, offset) re-created it) */
void ref_offsetgen_fullcov_set_state(void* g, uint64_t lcgState)
{
	OffsetGenRandomAlignedFullCoverageV2* cov =
		dynamic_cast<OffsetGenRandomAlignedFullCoverageV2*>( ( (RefOffsetGen*)g)->gen.get() );
	if(cov)
		cov->randomGen.current_lcg_state_ = lcgState % cov->randomGen.m_lcg_;
}

uint64_t ref_offsetgen_fullcov_modulus(void* g)
{
	OffsetGenRandomAlignedFullCoverageV2* cov =
		dynamic_cast<OffsetGenRandomAlignedFullCoverageV2*>( ( (RefOffsetGen*)g)->gen.get() );
	return cov ? cov->randomGen.m_lcg_ : 0;
}

} // extern "C"
