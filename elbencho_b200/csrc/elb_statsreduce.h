/*
 * Live statistics across the GPUs of one box.
 *
 * The reference sums the LiveOps / LiveLatency counters of all LocalWorkers on the host
 * (source/Statistics.cpp:414-470 live loop, :2728-2804 service status tree). Here the workers of
 * one process are spread over several GPUs ("one CUDA context per --gpuids entry"), and part of
 * their live state never leaves the GPU: the verify mismatch / verified / filled byte counters
 * are written by the kernels into per-worker counter blocks in HBM. The reducer keeps one small
 * slab per GPU, lets a gather kernel add the counter blocks of that GPU's workers to the
 * host-side counters of the same workers, and sums the slabs of all GPUs into the slab of the
 * first GPU with ONE grouped ncclReduce over NVLink. This is the only collective of the data
 * path, and it only carries statistics.
 *
 * NCCL is bound at run time (dlopen "libnccl.so.2", override with ELB_NCCL_LIB). With a single
 * GPU, or when NCCL cannot be loaded or initialised, the same snapshot is produced by summing on
 * the host (noted once on stderr for the multi-GPU case).
 */
#ifndef ELB_STATSREDUCE_H_
#define ELB_STATSREDUCE_H_

#include <cuda_runtime.h>

#include <mutex>
#include <string>
#include <vector>

#include "elbencho_b200.h"

namespace elb
{

class Manager;
class Worker;

/* slab layout: u64 slots, all reduced with a sum */
enum LiveSlot
{
	LiveSlot_OPS = 0,                       // entries, bytes, iops
	LiveSlot_OPS_READMIX = 3,               // entries, bytes, iops
	LiveSlot_LAT = 6,                       // the 8 values of elb_livelat
	LiveSlot_WORKERS_DONE = 14,
	LiveSlot_WORKERS_TOTAL = 15,
	LiveSlot_DEVCTR = 16,                   // ELB_DEVCTR_NUM values, gathered on the device
	LiveSlot_NUM = LiveSlot_DEVCTR + ELB_DEVCTR_NUM,
};

/* phase-end slab: the 4 latency histograms of a phase (IOPS, IOPS read-mix, entries, entries
 * read-mix; LatencyHistogram.h:28-45) as three regions, reduced with sum, min and max */
enum HistoSlot
{
	HistoSlot_NUMHISTOS = 4,
	HistoSlot_WORDS_PER_HISTO = ELB_LATHISTO_NUMBUCKETS + 2, // buckets, numStored, numMicroSecTotal
	HistoSlot_SUM = 0,
	HistoSlot_DEVCTR = HistoSlot_SUM + HistoSlot_NUMHISTOS * HistoSlot_WORDS_PER_HISTO,
	HistoSlot_MIN = HistoSlot_DEVCTR + ELB_DEVCTR_NUM, // end of the sum region
	HistoSlot_MAX = HistoSlot_MIN + HistoSlot_NUMHISTOS,
	HistoSlot_NUM = HistoSlot_MAX + HistoSlot_NUMHISTOS,
};

class LiveStatsReducer
{
	public:
		explicit LiveStatsReducer(Manager& manager);
		~LiveStatsReducer();

		/* sum over all workers of the manager. consumes the live latency counters
		   (add-and-reset, like LiveLatency::getAndResetAll of the reference) */
		void snapshot(elb_live_snapshot& out);

		/* phase end (all workers done): merged histograms [iops, iopsReadMix, entries,
		   entriesReadMix] and the summed device counter blocks, through ncclReduce sum / min / max
		   (SURVEY.md §8e). @return false if NCCL is not in use (caller merges on the host) */
		bool reducePhaseEnd(elb_histogram outHistos[4], uint64_t outDevCounters[ELB_DEVCTR_NUM] );

		bool usesNccl() const { return ncclReady && !ncclBroken; }
		const std::string& getNcclNote() const { return ncclNote; }

	private:
		struct PerGPU
		{
			int gpuID{-1};
			std::vector<Worker*> workers;
			cudaStream_t stream{NULL};
			uint64_t* devSend{NULL};       // LiveSlot_NUM
			uint64_t* devRecv{NULL};       // LiveSlot_NUM (meaningful on the root)
			uint64_t** devCtrPtrs{NULL};   // device array of the workers' counter block addresses
			uint64_t* hostSend{NULL};      // pinned, LiveSlot_NUM
			uint64_t* hostRecv{NULL};      // pinned, LiveSlot_NUM
			uint64_t* devHistoSend{NULL};  // HistoSlot_NUM
			uint64_t* devHistoRecv{NULL};  // HistoSlot_NUM
			uint64_t* hostHisto{NULL};     // pinned, HistoSlot_NUM (send staging; recv on the root)
			uint64_t** hostCtrPtrs{NULL};  // pinned
			void* comm{NULL};              // ncclComm_t
		};

		Manager& manager;
		std::vector<PerGPU> gpus;
		bool deviceReady{false};
		bool ncclReady{false};
		bool ncclBroken{false};
		std::string ncclNote;
		std::mutex mutex;

		void initNccl();
		void collectHostPart(PerGPU& gpu, uint64_t* slots);
		bool snapshotDevice(uint64_t* outSlots, bool& outUsedNccl);
		void snapshotHost(uint64_t* outSlots);
		void releaseDeviceState();
};

} // namespace elb

#endif /* ELB_STATSREDUCE_H_ */
