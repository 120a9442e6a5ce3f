/*
 * TEST INFRASTRUCTURE ONLY — CPU oracle for elbencho's LocalWorker hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this. The product library (elbencho_b200/libelbencho_b200.so) never links, loads or calls
 * anything in oracle/.
 *
 * Parity pinning (see oracle/README.md):
 *  - PRNGs and offset generators are checked bit-for-bit against the reference's own headers
 *    compiled into oracle/_ref/libelb_ref.so (tests/test_oracle_vs_ref.py, tests/golden/).
 *  - The --verify pattern is a closed form (byte x of the file = byte x%8 of LE u64 (x&~7)+salt),
 *    checked against hand-derived vectors and against the line-by-line restatement.
 *  - Random-fill CONTENT: parity unpinned in the reference itself (self-seeded PRNG / cuRAND);
 *    orc_fill_random_ctr is the CPU twin of the project's counter-based generator.
 */
#ifndef ELB_ORACLE_H_
#define ELB_ORACLE_H_

#include <stddef.h>
#include <stdint.h>
#include <time.h>

#include "elbencho_b200.h" /* shared config/stats structs only */

#ifdef __cplusplus
extern "C" {
#endif

/* ---- block modifiers / checkers (LocalWorker.cpp:2091-2230) ---- */
void orc_fill_pattern(char* buf, size_t bufLen, uint64_t fileOffset, uint64_t salt);
/* returns 0 if equal, 1 on mismatch (then errBuf holds the reference's exception text) */
int orc_verify_pattern(const char* buf, size_t bufLen, uint64_t fileOffset, uint64_t salt,
	uint64_t* outNumMismatchBytes, uint64_t* outFirstMismatchIdx, unsigned* outExpected,
	unsigned* outActual, char* errBuf, size_t errBufLen);
void orc_buf_fill(char* buf, uint64_t fillValue, size_t bufLen);

/* ---- PRNGs (toolkits/random) with injectable state ---- */
typedef struct orc_xoshiro256ss { uint64_t s[4]; } orc_xoshiro256ss;
uint64_t orc_xoshiro256ss_next(orc_xoshiro256ss* st);
void orc_xoshiro256ss_fill_buf(orc_xoshiro256ss* st, char* buf, uint64_t bufLen);

typedef struct orc_goldenprime
{
	orc_xoshiro256ss stateSeeder;
	uint64_t state;
	unsigned currentGoldenPrimeIdx;
} orc_goldenprime;
void orc_goldenprime_init(orc_goldenprime* st, uint64_t seed, const uint64_t seederState[4]);
uint64_t orc_goldenprime_next(orc_goldenprime* st);
void orc_goldenprime_fill_buf(orc_goldenprime* st, char* buf, uint64_t bufLen);

/* preWriteBufRandRefill host layout (LocalWorker.cpp:2209-2230) on top of golden prime */
void orc_rand_refill_goldenprime(orc_goldenprime* st, char* buf, size_t bufLen, unsigned pct);

/* CPU twin of the project's counter-based random fill (GPU layout of LocalWorker.cpp:2236-2277) */
void orc_fill_random_ctr(char* buf, uint64_t bufLen, unsigned pct, uint64_t seed,
	uint64_t blockCounter);

/* ---- offset generators (toolkits/offsetgen) ---- */
enum orc_offsetgen_kind
{
	ORC_OFFGEN_SEQUENTIAL = 0,
	ORC_OFFGEN_REVERSE_SEQ = 1,
	ORC_OFFGEN_RANDOM = 2,
	ORC_OFFGEN_RANDOM_ALIGNED = 3,
	ORC_OFFGEN_STRIDED = 4,
	ORC_OFFGEN_RANDOM_ALIGNED_FULLCOV = 5,
};

/* ---- the four --randalgo generators behind one next() (RandAlgoSelectorTk.cpp:37-53);
 * algo values = enum elb_offset_rand_algo. state: 4 words for the xoshiro variants, word 0 = seed
 * for golden prime (its seeder is never consulted by next() ) and for mt19937_64. ---- */
typedef struct orc_xoshiro256pp { uint64_t s[4]; } orc_xoshiro256pp; /* lane 0 of the SIMD class */
uint64_t orc_xoshiro256pp_next(orc_xoshiro256pp* st);

typedef struct orc_mt19937_64 { uint64_t mt[312]; unsigned idx; } orc_mt19937_64;
void orc_mt19937_64_seed(orc_mt19937_64* st, uint64_t seed);
uint64_t orc_mt19937_64_next(orc_mt19937_64* st);

typedef struct orc_randalgo
{
	int algo;
	union
	{
		orc_xoshiro256ss xoshiroSS;
		orc_goldenprime goldenPrime;
		orc_xoshiro256pp xoshiroPP;
		orc_mt19937_64 mt;
	} u;
} orc_randalgo;

int orc_randalgo_init(orc_randalgo* st, int algo, const uint64_t state[4]); /* 0 ok, -1 bad algo */
uint64_t orc_randalgo_next(orc_randalgo* st);
orc_randalgo* orc_randalgo_create(int algo, const uint64_t state[4]); /* heap, for ctypes */
void orc_randalgo_destroy(orc_randalgo* st);

typedef struct orc_offsetgen orc_offsetgen;

/* randState: xoshiro256** state for RANDOM/RANDOM_ALIGNED; lcgSeed: initial LCG state for FULLCOV
 * (the reference takes both from std::random_device) */
orc_offsetgen* orc_offsetgen_create(int kind, uint64_t numBytesTotal, uint64_t len,
	uint64_t offset, uint64_t blockSize, uint64_t numDataSetThreads,
	const uint64_t randState[4], uint64_t lcgSeed);
orc_offsetgen* orc_offsetgen_create_algo(int kind, uint64_t numBytesTotal, uint64_t len,
	uint64_t offset, uint64_t blockSize, uint64_t numDataSetThreads, int randAlgo,
	const uint64_t randState[4], uint64_t lcgSeed);
void orc_offsetgen_destroy(orc_offsetgen* g);
void orc_offsetgen_reset(orc_offsetgen* g);
void orc_offsetgen_reset_range(orc_offsetgen* g, uint64_t len, uint64_t offset);
uint64_t orc_offsetgen_next_offset(orc_offsetgen* g);
uint64_t orc_offsetgen_next_block_size(const orc_offsetgen* g);
uint64_t orc_offsetgen_bytes_total(const orc_offsetgen* g);
uint64_t orc_offsetgen_bytes_left(const orc_offsetgen* g);
void orc_offsetgen_add_bytes_submitted(orc_offsetgen* g, uint64_t numBytes);

/* ---- LatencyHistogram / UnitTk ---- */
void orc_histogram_reset(elb_histogram* h);
void orc_histogram_add_latency(elb_histogram* h, uint64_t latencyMicroSec);
void orc_histogram_merge(elb_histogram* dst, const elb_histogram* src);
double orc_histogram_percentile(const elb_histogram* h, double percentage);
uint64_t orc_per_sec_from_usec(uint64_t totalValue, uint64_t elapsedUSec);

/* ---- RateLimiter (toolkits/RateLimiter.h:13-66) ---- */
typedef struct orc_ratelimiter
{
	uint64_t limitPerSec;    // 0 = no limit
	uint64_t numDoneThisSec;
	struct timespec startT;  // when the current second started
} orc_ratelimiter;

void orc_ratelimiter_init_start(orc_ratelimiter* rl, uint64_t limitPerSec);
int orc_ratelimiter_wait(orc_ratelimiter* rl, uint64_t nextSize); // 1 = had to wait

/* ---- the CPU LocalWorker: run one phase with cfg->numThreads threads (rwBlockSized for
 * iodepth 1, aioBlockSized on kernel AIO for iodepth > 1) ---- */
typedef struct orc_worker_result
{
	elb_liveops liveOps;
	elb_liveops liveOpsReadMix;
	elb_histogram iopsLatHisto;
	elb_histogram entriesLatHisto;
	uint64_t elapsedUSec;
	int32_t gotPhaseWork;
	int32_t hadError;
	char errorMsg[512];
} orc_worker_result;

/* Runs benchPhase on the host CPU the way LocalWorker::run does for that phase (file open/prepare
 * like ProgArgs::prepareBenchPathFDsVec). results must hold cfg->numThreads entries. Returns 0 if
 * all workers succeeded. outPhaseResults may be NULL. */
int orc_run_phase(const elb_cfg* cfg, int benchPhase, orc_worker_result* results,
	elb_phase_results* outPhaseResults);

/* expected totals per worker (WorkerManager::getPhaseNumEntriesAndBytes, :333-487) */
void orc_expected_per_worker(const elb_cfg* cfg, int benchPhase, uint64_t* outEntries,
	uint64_t* outBytes);

/* micro-benchmarks of the CPU block modifiers: returns bytes/s of one thread */
double orc_bench_fill_pattern(size_t blockSize, size_t numBlocks);
double orc_bench_verify_pattern(size_t blockSize, size_t numBlocks);

#ifdef __cplusplus
}
#endif

#endif /* ELB_ORACLE_H_ */
