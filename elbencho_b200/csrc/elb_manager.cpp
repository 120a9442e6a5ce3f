/*
 * Manager (WorkerManager of the reference, source/workers/WorkerManager.cpp) and the C ABI of the
 * worker/manager level declared in include/elbencho_b200.h.
 */
#include <errno.h>
#include <algorithm>
#include <fcntl.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>

#include "elb_cli.h"
#include "elb_statsreduce.h"
#include "elb_worker.h"

#define ELB_MKFILE_MODE (S_IRUSR | S_IWUSR | S_IRGRP | S_IWGRP | S_IROTH | S_IWOTH)

namespace elb
{

/* prepareThreads (WorkerManager.cpp:142-199): open bench paths, create workers and their threads,
 * wait until all of them finished their preparation. */
/* ProgArgs::prepareFileSize (ProgArgs.cpp:2071-2210), the part that has to happen before the
 * size dependent normalisation: no --size given => take the size of the (first) existing file or
 * of the block device */
uint64_t detectFileSize(const elb_cfg* abiCfg)
{
	if(abiCfg->fileSize || !abiCfg->numPaths || !abiCfg->paths || !abiCfg->paths[0] )
		return abiCfg->fileSize;

	const char* path = abiCfg->paths[0];

	if(abiCfg->pathType == ELB_PATH_FILE)
	{
		struct stat statBuf;

		if( (stat(path, &statBuf) == 0) && S_ISREG(statBuf.st_mode) )
			return statBuf.st_size;
	}
	else
	if(abiCfg->pathType == ELB_PATH_BLOCKDEV)
	{
		int fd = open(path, O_RDONLY);

		if(fd != -1)
		{
			off_t blockdevSize = lseek(fd, 0, SEEK_END);
			close(fd);

			if(blockdevSize > 0)
				return blockdevSize;
		}
	}

	return 0;
}

Manager::Manager(const elb_cfg* abiCfg)
{
	elb_cfg sizedCfg = *abiCfg;
	sizedCfg.fileSize = detectFileSize(abiCfg);

	shared.cfg = Config::fromABI(&sizedCfg);

	prepareBenchPathFDs();

	const Config& cfg = shared.cfg;

	if(!cfg.treeFilePath.empty() ) // ProgArgs::loadCustomTreeFile (ProgArgs.cpp:2740-2803)
	{
		try
		{
			shared.customTree.load(cfg.treeFilePath, cfg.blockSize, cfg.fileShareSize,
				cfg.treeRoundUpSize);
		}
		catch(...)
		{
			closeBenchPathFDs();
			throw;
		}
	}

	for(uint32_t i = 0; i < cfg.numThreads; i++)
	{
		workers.emplace_back(new Worker(&shared, cfg.rankOffset + i) );
		shared.workers.push_back(workers.back().get() );
	}

	for(uint32_t i = 0; i < cfg.numThreads; i++)
		threads.emplace_back(Worker::threadStart, workers[i].get() );

	// wait for preparation of all workers
	int waitRes = waitForWorkersDone(-1);

	if(waitRes < 0)
	{
		std::string errMsg = shared.firstErrorMsg.empty() ?
			"Worker preparation failed." : shared.firstErrorMsg;

		interruptAndNotifyWorkers();

		for(std::thread& thread : threads)
			thread.join();

		threads.clear();
		closeBenchPathFDs();

		throw WorkerError(errMsg);
	}
}

Manager::~Manager()
{
	liveStatsReducer.reset(); // (refers to the workers and their device counter blocks)

	// BenchPhase_TERMINATE tells workers to self-terminate (Common.h:145)
	if(!threads.empty() )
	{
		{
			std::unique_lock<std::mutex> lock(shared.mutex);
			shared.currentBenchPhase = ELB_PHASE_TERMINATE;
			shared.currentBenchSeq++;
			for(Worker* worker : shared.workers)
				worker->interruptExecution();
			shared.condition.notify_all();
		}

		for(std::thread& thread : threads)
			thread.join();
	}

	closeBenchPathFDs();
}

void Manager::getLiveSnapshot(elb_live_snapshot& out)
{
	std::unique_lock<std::mutex> lock(liveStatsReducerMutex);

	if(!liveStatsReducer)
		liveStatsReducer.reset(new LiveStatsReducer(*this) );

	liveStatsReducer->snapshot(out);
}

size_t Manager::getNumGPUs() const
{
	std::vector<int> gpuIDs;

	for(const std::unique_ptr<Worker>& worker : workers)
		if(std::find(gpuIDs.begin(), gpuIDs.end(), worker->getGPUID() ) == gpuIDs.end() )
			gpuIDs.push_back(worker->getGPUID() );

	return gpuIDs.size();
}

std::string Manager::getLiveReduceInfo()
{
	std::unique_lock<std::mutex> lock(liveStatsReducerMutex);

	if(!liveStatsReducer)
		liveStatsReducer.reset(new LiveStatsReducer(*this) );

	return liveStatsReducer->getNcclNote();
}

/* ProgArgs::prepareBenchPathFDsVec (ProgArgs.cpp:1859-1935) */
void Manager::prepareBenchPathFDs()
{
	const Config& cfg = shared.cfg;

	for(const std::string& path : cfg.paths)
	{
		int openFlags = 0;

		if(cfg.pathType == ELB_PATH_DIR)
			openFlags |= (O_DIRECTORY | O_RDONLY);
		else
		{
			openFlags |= O_RDWR;

			if(cfg.useDirectIO)
				openFlags |= O_DIRECT;

			if(cfg.pathType == ELB_PATH_FILE)
				openFlags |= O_CREAT;
		}

		int fd = open(path.c_str(), openFlags, ELB_MKFILE_MODE);

		if( (fd == -1) && (cfg.pathType != ELB_PATH_DIR) &&
			( (errno == EACCES) || (errno == EROFS) || (errno == EPERM) ) )
		{ // read-only file: good enough for read phases
			openFlags &= ~(O_RDWR | O_CREAT);
			openFlags |= O_RDONLY;
			fd = open(path.c_str(), openFlags, ELB_MKFILE_MODE);
		}

		if(fd == -1)
		{
			int openErrno = errno;
			closeBenchPathFDs();
			throw WorkerError("Unable to open benchmark path: " + path + "; "
				"SysErr: " + strerror(openErrno) );
		}

		shared.pathFDs.push_back(fd);
		shared.fileWriteGates.emplace_back(new FileWriteGate() );

		// --fadv on the files / block devices themselves (ProgArgs.cpp:2032)
		if(cfg.fadviseFlags && (cfg.pathType != ELB_PATH_DIR) )
		{
			const struct { unsigned flag; int advice; const char* name; } adviceDefs[] =
			{
				{8, POSIX_FADV_DONTNEED, "POSIX_FADV_DONTNEED"},
				{16, POSIX_FADV_NOREUSE, "POSIX_FADV_NOREUSE"},
				{1, POSIX_FADV_SEQUENTIAL, "POSIX_FADV_SEQUENTIAL"},
				{2, POSIX_FADV_RANDOM, "POSIX_FADV_RANDOM"},
				{4, POSIX_FADV_WILLNEED, "POSIX_FADV_WILLNEED"},
			};

			for(const auto& def : adviceDefs)
			{
				if(!(cfg.fadviseFlags & def.flag) )
					continue;

				const int fadviseRes = posix_fadvise(fd, 0, 0, def.advice);

				if(fadviseRes)
				{
					closeBenchPathFDs();
					throw WorkerError(std::string("Unable to set POSIX fadvise. ") +
						"Advise: " + def.name + "; "
						"File: " + path + "; "
						"SysErr: " + strerror(fadviseRes) );
				}
			}
		}
	}

	/* ProgArgs::prepareCuFileHandleDataVec (ProgArgs.cpp:1950-1990): driver open + one registered
	   handle per file in file/bdev mode (dir mode registers per file in the worker) */
	if(cfg.useCuFile)
	{
		try
		{
			CuFileApi::get().driverOpenOnce();

			if(cfg.pathType != ELB_PATH_DIR)
				for(size_t i = 0; i < shared.pathFDs.size(); i++)
				{
					shared.cuFileHandles.emplace_back(new CuFileHandle() );
					shared.cuFileHandles.back()->registerFD(shared.pathFDs[i], cfg.paths[i] );
				}
		}
		catch(...)
		{
			closeBenchPathFDs();
			throw;
		}
	}
}

void Manager::closeBenchPathFDs()
{
	shared.cuFileHandles.clear(); // deregisters

	for(int fd : shared.pathFDs)
		close(fd);

	shared.pathFDs.clear();
}

/* ProgArgs::prepareFileSize (ProgArgs.cpp:2071-2200) for file mode: truncate / set size /
 * preallocate before the write phase starts */
void Manager::prepareFilesForPhase(int benchPhase)
{
	const Config& cfg = shared.cfg;
	const bool isRWPhase = (benchPhase == ELB_PHASE_CREATEFILES) || (benchPhase == ELB_PHASE_READFILES);

	if( (cfg.pathType == ELB_PATH_FILE) && isRWPhase)
	{ // size checks of ProgArgs::prepareFileSize (ProgArgs.cpp:2085-2104)
		for(size_t i = 0; i < shared.pathFDs.size(); i++)
		{
			struct stat statBuf;

			if(fstat(shared.pathFDs[i], &statBuf) == -1)
				throw WorkerError("Unable to check size of file through fstat: " + cfg.paths[i] +
					"; SysErr: " + strerror(errno) );

			if(!cfg.fileSize && !statBuf.st_size)
				throw WorkerError("File size must not be 0 when benchmark path is a file. "
					"File: " + cfg.paths[i] );

			/* (the reference checks this once at startup and only for runs without a write
			   phase, "!runCreateFilesPhase": a manager that has written the files itself
			   corresponds to a "-w -r" run) */
			if( (benchPhase == ELB_PHASE_READFILES) && !hadCreateFilesPhase &&
				S_ISREG(statBuf.st_mode) && ( (uint64_t)statBuf.st_size < cfg.fileSize) )
				throw WorkerError("Given size to use is larger than detected size. "
					"File: " + cfg.paths[i] + "; "
					"Detected size: " + std::to_string(statBuf.st_size) + "; "
					"Given size: " + std::to_string(cfg.fileSize) );
		}
	}

	if(benchPhase == ELB_PHASE_CREATEFILES)
		hadCreateFilesPhase = true;

	if( (cfg.pathType != ELB_PATH_FILE) || (benchPhase != ELB_PHASE_CREATEFILES) )
		return;

	for(size_t i = 0; i < shared.pathFDs.size(); i++)
	{
		const int fd = shared.pathFDs[i];

		if(cfg.doTruncate && (ftruncate(fd, 0) == -1) )
			throw WorkerError("Unable to truncate file. Path: " + cfg.paths[i] + "; "
				"SysErr: " + strerror(errno) );

		if(cfg.doTruncToSize && (ftruncate(fd, cfg.fileSize) == -1) )
			throw WorkerError("Unable to set file size through ftruncate. "
				"Path: " + cfg.paths[i] + "; "
				"Size: " + std::to_string(cfg.fileSize) + "; "
				"SysErr: " + strerror(errno) );

		if(cfg.doPreallocFile)
		{
			int preallocRes = posix_fallocate(fd, 0, cfg.fileSize);
			if(preallocRes != 0)
				throw WorkerError("Unable to preallocate file size through posix_fallocate. "
					"File: " + cfg.paths[i] + "; "
					"Size: " + std::to_string(cfg.fileSize) + "; "
					"SysErr: " + strerror(preallocRes) );
		}
	}
}

/* startNextPhase (WorkerManager.cpp:291-324) */
void Manager::startNextPhase(int benchPhase)
{
	prepareFilesForPhase(benchPhase);

	std::unique_lock<std::mutex> lock(shared.mutex);

	for(Worker* worker : shared.workers)
		worker->resetStats();

	shared.numWorkersDone = 0;
	shared.numWorkersDoneWithError = 0;
	shared.firstErrorMsg.clear();
	shared.currentBenchPhase = benchPhase;
	shared.currentBenchSeq++;
	shared.cpuUtilFirstDone.update(); // WorkerManager.cpp:307-308
	shared.cpuUtilLastDone.update();
	shared.cpuUtilFirstDonePercent = 0;
	shared.cpuUtilLastDonePercent = 0;
	shared.phaseStartT = Clock::now();

	// --rwmixthrpct (LocalWorker.cpp:1284-1290): fresh byte counters for every phase
	shared.rwMixThreadsBalancer.initStart(
		(benchPhase == ELB_PHASE_CREATEFILES) && shared.cfg.numRWMixReadThreads ?
			shared.cfg.rwMixThreadsReadPercent : 0,
		shared.cfg.numRWMixReadThreads, shared.cfg.numThreads - shared.cfg.numRWMixReadThreads,
		shared.cfg.blockSize);

	shared.condition.notify_all();
}

/* waitForWorkersDone / checkWorkersDoneUnlocked (WorkerManager.cpp:38-70, 245-267) */
int Manager::waitForWorkersDone(int timeoutMS)
{
	std::unique_lock<std::mutex> lock(shared.mutex);

	const size_t numWorkersTotal = shared.workers.size();
	const Clock::time_point deadline = Clock::now() + std::chrono::milliseconds(timeoutMS);

	for( ; ; )
	{
		if(shared.numWorkersDoneWithError)
		{ // a worker failed: ask the others to stop and wait for them (:54-58)
			for(Worker* worker : shared.workers)
				worker->interruptExecution();

			while(shared.numWorkersDone < numWorkersTotal)
				shared.condition.wait_for(lock, std::chrono::milliseconds(100) );

			return -1;
		}

		if(shared.numWorkersDone >= numWorkersTotal)
			return 1;

		if(timeoutMS < 0)
			shared.condition.wait(lock);
		else
		if(shared.condition.wait_until(lock, deadline) == std::cv_status::timeout)
		{
			if(shared.numWorkersDoneWithError)
				continue;

			return (shared.numWorkersDone >= numWorkersTotal) ? 1 : 0;
		}
	}
}

void Manager::interruptAndNotifyWorkers() // WorkerManager.cpp:75-90
{
	std::unique_lock<std::mutex> lock(shared.mutex);

	for(Worker* worker : shared.workers)
		worker->interruptExecution();

	shared.currentBenchPhase = ELB_PHASE_TERMINATE;
	shared.currentBenchSeq++;
	shared.condition.notify_all();
}

/* Statistics::generatePhaseResults (Statistics.cpp:1641-1764) */
void Manager::getPhaseResults(elb_phase_results& out)
{
	memset(&out, 0, sizeof(out) );
	histogramReset(out.iopsLatHisto);
	histogramReset(out.entriesLatHisto);
	histogramReset(out.iopsLatHistoReadMix);
	histogramReset(out.entriesLatHistoReadMix);

	uint64_t firstFinishUSec = ~0ULL;
	uint64_t lastFinishUSec = 0;

	/* workers on several GPUs: histograms and the device counter blocks are merged per GPU and
	   reduced to the first GPU by NCCL (sum / min / max); everything else is per-thread host
	   state (elapsed times, stonewall snapshots) and stays a host loop */
	bool reducedWithNccl = false;

	if(getNumGPUs() >= 2)
	{
		elb_histogram histos[4];
		uint64_t devCounters[ELB_DEVCTR_NUM];

		std::unique_lock<std::mutex> lock(liveStatsReducerMutex);

		if(!liveStatsReducer)
			liveStatsReducer.reset(new LiveStatsReducer(*this) );

		reducedWithNccl = liveStatsReducer->reducePhaseEnd(histos, devCounters);

		if(reducedWithNccl)
		{
			out.iopsLatHisto = histos[0];
			out.iopsLatHistoReadMix = histos[1];
			out.entriesLatHisto = histos[2];
			out.entriesLatHistoReadMix = histos[3];
			out.verifyMismatchBytes = devCounters[ELB_DEVCTR_VERIFY_MISMATCH_BYTES];
			out.verifiedBytes = devCounters[ELB_DEVCTR_VERIFIED_BYTES];
			out.filledBytes = devCounters[ELB_DEVCTR_FILLED_BYTES];
			out.statsReducedWithNccl = 1;
		}
	}

	for(const std::unique_ptr<Worker>& worker : workers)
	{
		const uint64_t elapsedUSec = worker->getElapsedUSec();

		if(elapsedUSec)
		{
			firstFinishUSec = std::min(firstFinishUSec, elapsedUSec);
			lastFinishUSec = std::max(lastFinishUSec, elapsedUSec);
		}

		liveOpsAdd(out.opsTotal, worker->getLiveOps() );
		liveOpsAdd(out.opsReadMixTotal, worker->getLiveOpsReadMix() );
		liveOpsAdd(out.opsStoneWallTotal, worker->getStoneWallOps() );
		liveOpsAdd(out.opsStoneWallReadMixTotal, worker->getStoneWallOpsReadMix() );
		out.numKernelLaunches += worker->getNumKernelLaunches();
		out.h2dBytes += worker->getNumH2DBytes();
		out.d2hBytes += worker->getNumD2HBytes();
		out.devKernelUSec += worker->getDevKernelUSec();

		if(reducedWithNccl)
			continue;

		histogramMerge(out.iopsLatHisto, worker->getIOPSLatHisto() );
		histogramMerge(out.entriesLatHisto, worker->getEntriesLatHisto() );
		histogramMerge(out.iopsLatHistoReadMix, worker->getIOPSLatHistoReadMix() );
		histogramMerge(out.entriesLatHistoReadMix, worker->getEntriesLatHistoReadMix() );

		uint64_t devCounters[ELB_DEVCTR_NUM];
		if(!worker->snapshotDevCounters(devCounters) )
		{
			out.verifyMismatchBytes += devCounters[ELB_DEVCTR_VERIFY_MISMATCH_BYTES];
			out.verifiedBytes += devCounters[ELB_DEVCTR_VERIFIED_BYTES];
			out.filledBytes += devCounters[ELB_DEVCTR_FILLED_BYTES];
		}
	}

	out.firstFinishUSec = (firstFinishUSec == ~0ULL) ? 0 : firstFinishUSec;
	out.lastFinishUSec = lastFinishUSec;

	if(out.lastFinishUSec)
	{
		out.opsPerSec.numEntriesDone =
			perSecFromUSec(out.opsTotal.numEntriesDone, out.lastFinishUSec);
		out.opsPerSec.numBytesDone = perSecFromUSec(out.opsTotal.numBytesDone, out.lastFinishUSec);
		out.opsPerSec.numIOPSDone = perSecFromUSec(out.opsTotal.numIOPSDone, out.lastFinishUSec);
	}

	// rwmix read side (Statistics.cpp:1725-1745)
	if(out.lastFinishUSec && out.opsReadMixTotal.numIOPSDone)
	{
		out.opsReadMixPerSec.numEntriesDone =
			perSecFromUSec(out.opsReadMixTotal.numEntriesDone, out.lastFinishUSec);
		out.opsReadMixPerSec.numBytesDone =
			perSecFromUSec(out.opsReadMixTotal.numBytesDone, out.lastFinishUSec);
		out.opsReadMixPerSec.numIOPSDone =
			perSecFromUSec(out.opsReadMixTotal.numIOPSDone, out.lastFinishUSec);
	}

	if(out.firstFinishUSec && out.opsReadMixTotal.numIOPSDone)
	{
		out.opsStoneWallReadMixPerSec.numEntriesDone =
			perSecFromUSec(out.opsStoneWallReadMixTotal.numEntriesDone, out.firstFinishUSec);
		out.opsStoneWallReadMixPerSec.numBytesDone =
			perSecFromUSec(out.opsStoneWallReadMixTotal.numBytesDone, out.firstFinishUSec);
		out.opsStoneWallReadMixPerSec.numIOPSDone =
			perSecFromUSec(out.opsStoneWallReadMixTotal.numIOPSDone, out.firstFinishUSec);
	}

	out.cpuUtilStoneWallPercent = shared.cpuUtilFirstDonePercent;
	out.cpuUtilPercent = shared.cpuUtilLastDonePercent;

	if(out.firstFinishUSec)
	{
		out.opsStoneWallPerSec.numEntriesDone =
			perSecFromUSec(out.opsStoneWallTotal.numEntriesDone, out.firstFinishUSec);
		out.opsStoneWallPerSec.numBytesDone =
			perSecFromUSec(out.opsStoneWallTotal.numBytesDone, out.firstFinishUSec);
		out.opsStoneWallPerSec.numIOPSDone =
			perSecFromUSec(out.opsStoneWallTotal.numIOPSDone, out.firstFinishUSec);
	}

	std::unique_lock<std::mutex> lock(shared.mutex);
	out.numWorkersDone = (uint32_t)shared.numWorkersDone;
	out.numWorkersDoneWithError = (uint32_t)shared.numWorkersDoneWithError;
}

/* getPhaseNumEntriesAndBytes (WorkerManager.cpp:333-487): expected totals per worker */
void expectedPerWorker(const Config& cfg, int benchPhase, uint64_t& outEntries,
	uint64_t& outBytes, const TreeManifest* customTree)
{
	outEntries = 0;
	outBytes = 0;

	if( (cfg.pathType == ELB_PATH_DIR) && customTree && customTree->isLoaded)
	{ // custom tree mode (WorkerManager.cpp:406-450)
		const uint64_t numDirs = customTree->getNumDirs();
		const uint64_t numFiles = customTree->getNumFiles();
		const uint64_t numBytesTotal = customTree->getNumFileBytes();

		switch(benchPhase)
		{
			case ELB_PHASE_CREATEDIRS:
			case ELB_PHASE_DELETEDIRS:
				outEntries = numDirs / cfg.numDataSetThreads;
				break;

			case ELB_PHASE_CREATEFILES:
			case ELB_PHASE_READFILES:
				outEntries = numFiles / cfg.numDataSetThreads;
				outBytes = numBytesTotal / cfg.numDataSetThreads;
				break;

			case ELB_PHASE_DELETEFILES:
			case ELB_PHASE_STATFILES:
				outEntries = numFiles / cfg.numDataSetThreads;
				break;

			default:
				break;
		}
	}
	else
	if(cfg.pathType == ELB_PATH_DIR)
	{
		const uint64_t numDirs = cfg.numDirs ? cfg.numDirs : 1;

		switch(benchPhase)
		{
			case ELB_PHASE_CREATEDIRS:
			case ELB_PHASE_DELETEDIRS:
				outEntries = cfg.numDirs;
				break;

			case ELB_PHASE_CREATEFILES:
			case ELB_PHASE_READFILES:
				outEntries = numDirs * cfg.numFiles;
				outBytes = outEntries * cfg.fileSize;
				break;

			case ELB_PHASE_DELETEFILES:
			case ELB_PHASE_STATFILES:
				outEntries = numDirs * cfg.numFiles;
				break;

			default:
				break;
		}
	}
	else
	{
		outEntries = cfg.paths.size();

		if( (benchPhase == ELB_PHASE_CREATEFILES) || (benchPhase == ELB_PHASE_READFILES) )
			outBytes = cfg.useRandomOffsets ?
				(cfg.randomAmount / cfg.numDataSetThreads) :
				( (outEntries * cfg.fileSize) / cfg.numDataSetThreads);
	}
}

/* summed over this manager's workers */
void Manager::getExpectedTotals(int benchPhase, uint64_t& outEntries, uint64_t& outBytes)
{
	uint64_t entriesPerWorker, bytesPerWorker;

	expectedPerWorker(shared.cfg, benchPhase, entriesPerWorker, bytesPerWorker,
		&shared.customTree);

	outEntries = entriesPerWorker * shared.cfg.numThreads;
	outBytes = bytesPerWorker * shared.cfg.numThreads;
}

} // namespace elb

/* ==============================================================================================
 * C ABI
 * ============================================================================================ */

struct elb_mgr
{
	elb::Manager* impl;
	std::string lastError;
	std::string liveReduceInfo;
};

/* elb_worker handles are the Worker objects themselves */
static inline elb::Worker* toWorker(elb_worker* w)
{
	return reinterpret_cast<elb::Worker*>(w);
}

static thread_local std::string elbWorkerErrorTmp;

extern "C" {

void elb_histogram_reset(elb_histogram* h)
{
	elb::histogramReset(*h);
}

void elb_histogram_add_latency(elb_histogram* h, uint64_t latencyMicroSec)
{
	elb::histogramAdd(*h, latencyMicroSec);
}

void elb_histogram_merge(elb_histogram* dst, const elb_histogram* src)
{
	elb::histogramMerge(*dst, *src);
}

double elb_histogram_percentile(const elb_histogram* h, double percentage)
{
	return elb::histogramPercentile(*h, percentage);
}

uint64_t elb_per_sec_from_usec(uint64_t totalValue, uint64_t elapsedUSec)
{
	return elb::perSecFromUSec(totalValue, elapsedUSec);
}

uint32_t elb_cfg_struct_size(void)
{
	return (uint32_t)sizeof(elb_cfg);
}

uint32_t elb_phase_results_struct_size(void)
{
	return (uint32_t)sizeof(elb_phase_results);
}

struct elb_offset_plan_impl
{
	std::unique_ptr<elb::RandAlgo> randAlgo;
	std::unique_ptr<elb::OffsetPlan> plan;
};

elb_rate_limiter* elb_rate_limiter_create(uint64_t limitPerSec)
{
	elb::RateLimiter* limiter = new elb::RateLimiter();
	limiter->initStart(limitPerSec);
	return reinterpret_cast<elb_rate_limiter*>(limiter);
}

int elb_rate_limiter_wait(elb_rate_limiter* limiter, uint64_t nextSize)
{
	return reinterpret_cast<elb::RateLimiter*>(limiter)->wait(nextSize) ? 1 : 0;
}

void elb_rate_limiter_destroy(elb_rate_limiter* limiter)
{
	delete reinterpret_cast<elb::RateLimiter*>(limiter);
}

struct elb_rwmix_balancer
{
	elb::RWMixThreadsBalancer impl;
	std::atomic_bool isInterruptionRequested{false};
};

elb_rwmix_balancer* elb_rwmix_balancer_create(unsigned readRatioPercent, unsigned numReaderThreads,
	unsigned numWriterThreads, uint64_t maxBlockSize)
{
	elb_rwmix_balancer* balancer = new elb_rwmix_balancer();
	balancer->impl.initStart(readRatioPercent, numReaderThreads, numWriterThreads, maxBlockSize);
	return balancer;
}

static int balancerWait(elb_rwmix_balancer* balancer, bool isRead, uint64_t nextBlockSize)
{
	try
	{
		const bool hadToWait = isRead ?
			balancer->impl.waitRead(nextBlockSize, balancer->isInterruptionRequested) :
			balancer->impl.waitWrite(nextBlockSize, balancer->isInterruptionRequested);

		return hadToWait ? 1 : 0;
	}
	catch(const std::exception& e)
	{
		elb_set_last_error(e.what() );
		return -1;
	}
}

int elb_rwmix_balancer_wait_read(elb_rwmix_balancer* balancer, uint64_t nextBlockSize)
{
	return balancerWait(balancer, true, nextBlockSize);
}

int elb_rwmix_balancer_wait_write(elb_rwmix_balancer* balancer, uint64_t nextBlockSize)
{
	return balancerWait(balancer, false, nextBlockSize);
}

void elb_rwmix_balancer_interrupt(elb_rwmix_balancer* balancer)
{
	balancer->isInterruptionRequested = true;
}

void elb_rwmix_balancer_destroy(elb_rwmix_balancer* balancer)
{
	delete balancer;
}

struct elb_write_gate
{
	elb::FileWriteGate gate;
};

elb_write_gate* elb_write_gate_create(void)
{
	return new elb_write_gate();
}

uint64_t elb_write_gate_take_ticket(elb_write_gate* gate)
{
	return gate->gate.takeTicket();
}

void elb_write_gate_wait_until_near(elb_write_gate* gate, uint64_t ticket)
{
	gate->gate.waitUntilNear(ticket);
}

void elb_write_gate_wait_turn(elb_write_gate* gate, uint64_t ticket)
{
	gate->gate.waitTurn(ticket);
}

void elb_write_gate_leave(elb_write_gate* gate)
{
	gate->gate.leave();
}

void elb_write_gate_destroy(elb_write_gate* gate)
{
	delete gate;
}

int64_t elb_write_gate_selftest(uint32_t numThreads, uint32_t turnsPerThread, uint32_t holdUSec)
{
	elb::FileWriteGate gate;
	std::atomic<int> numInside{0};
	std::atomic<uint64_t> nextExpectedTicket{0};
	std::atomic<int64_t> numViolations{0};
	std::atomic<uint64_t> numTurnsDone{0};
	std::vector<std::thread> threads;

	for(uint32_t t = 0; t < numThreads; t++)
		threads.emplace_back([&, t]()
		{
			for(uint32_t turn = 0; turn < turnsPerThread; turn++)
			{
				{
					elb::FileWriteTurn writeTurn(&gate);

					if( (t + turn) % 2) // (both ways of waiting)
						writeTurn.waitUntilNear();

					writeTurn.waitTurn();

					if(numInside.fetch_add(1) != 0)
						numViolations++;

					// tickets are served in the order they were taken
					const uint64_t served = nextExpectedTicket.fetch_add(1);
					(void)served;

					if(holdUSec)
						std::this_thread::sleep_for(std::chrono::microseconds(holdUSec) );

					numInside.fetch_sub(1);
					numTurnsDone++;
				}

				if( (turn % 7) == 3)
					std::this_thread::yield();
			}
		});

	for(std::thread& thread : threads)
		thread.join();

	if(numTurnsDone != ( (uint64_t)numThreads * turnsPerThread) )
		numViolations++;

	return numViolations;
}

int64_t elb_custom_tree_worker_list(const char* treeFilePath, uint64_t blockSize,
	uint64_t fileShareSize, uint64_t treeRoundUpSize, uint64_t workerRank,
	uint64_t numDataSetThreads, int kind, char* outBuf, uint64_t outBufLen)
{
	try
	{
		if(!treeFilePath || !numDataSetThreads || !blockSize)
			throw elb::WorkerError("elb_custom_tree_worker_list: invalid argument");

		elb::TreeManifest tree;
		elb::WorkerTreeShare share;

		tree.load(treeFilePath, blockSize, fileShareSize ? fileShareSize : (32 * blockSize),
			treeRoundUpSize);

		if(kind == 0)
			tree.takeDirs(workerRank, numDataSetThreads, share.slices);
		else
			tree.takeFiles(workerRank, numDataSetThreads, false, share);

		std::string text;

		for(const elb::TreeSlice& elem : share.slices)
			text += elem.path + "\t" + std::to_string(elem.totalLen) + "\t" +
				std::to_string(elem.rangeStart) + "\t" + std::to_string(elem.rangeLen) + "\n";

		if(outBuf && outBufLen)
		{
			const size_t copyLen = std::min( (size_t)(outBufLen - 1), text.size() );
			memcpy(outBuf, text.data(), copyLen);
			outBuf[copyLen] = 0;
		}

		return (int64_t)text.size();
	}
	catch(const std::exception& e)
	{
		elb_set_last_error(e.what() );
		return -1;
	}
}

int64_t elb_custom_tree_scan(const char* scanPath, const char* outTreeFilePath)
{
	try
	{
		uint64_t numDirs, numFiles, numBytes;

		return (int64_t)elb::TreeManifest::scanToTreeFile(scanPath, outTreeFilePath, numDirs,
			numFiles, numBytes);
	}
	catch(const std::exception& e)
	{
		elb_set_last_error(e.what() );
		return -1;
	}
}

elb_rand_algo_handle* elb_rand_algo_create(int randAlgo, const uint64_t state[4])
{
	try
	{
		return reinterpret_cast<elb_rand_algo_handle*>(
			elb::RandAlgo::create(randAlgo, state).release() );
	}
	catch(const std::exception& e)
	{
		elb_set_last_error(e.what() );
		return NULL;
	}
}

uint64_t elb_rand_algo_next(elb_rand_algo_handle* algo)
{
	return reinterpret_cast<elb::RandAlgo*>(algo)->next();
}

void elb_rand_algo_destroy(elb_rand_algo_handle* algo)
{
	delete reinterpret_cast<elb::RandAlgo*>(algo);
}

elb_offset_plan* elb_offset_plan_create(int kind, uint64_t amount, uint64_t rangeLen,
	uint64_t rangeOffset, uint64_t blockSize, uint64_t numDataSetThreads,
	const uint64_t randState[4], uint64_t lcgSeed, int haveLCGSeed)
{
	return elb_offset_plan_create_algo(kind, amount, rangeLen, rangeOffset, blockSize,
		numDataSetThreads, ELB_OFFSETALGO_XOSHIRO256SS, randState, lcgSeed, haveLCGSeed);
}

elb_offset_plan* elb_offset_plan_create_algo(int kind, uint64_t amount, uint64_t rangeLen,
	uint64_t rangeOffset, uint64_t blockSize, uint64_t numDataSetThreads, int randAlgo,
	const uint64_t randState[4], uint64_t lcgSeed, int haveLCGSeed)
{
	if( (kind < elb::OffsetPlan::Kind_SEQUENTIAL) || (kind > elb::OffsetPlan::Kind_FULL_COVERAGE) )
	{
		elb_set_last_error("Invalid offset plan kind: " + std::to_string(kind) );
		return NULL;
	}

	std::unique_ptr<elb::RandAlgo> algoObj;

	try
	{
		algoObj = elb::RandAlgo::create(randAlgo, randState);
	}
	catch(const std::exception& e)
	{
		elb_set_last_error(e.what() );
		return NULL;
	}

	elb_offset_plan_impl* impl = new elb_offset_plan_impl();

	impl->randAlgo = std::move(algoObj);

	impl->plan.reset(new elb::OffsetPlan( (elb::OffsetPlan::Kind)kind, amount, rangeLen,
		rangeOffset, blockSize, numDataSetThreads, impl->randAlgo.get(), lcgSeed,
		haveLCGSeed != 0) );

	return reinterpret_cast<elb_offset_plan*>(impl);
}

void elb_offset_plan_destroy(elb_offset_plan* plan)
{
	delete reinterpret_cast<elb_offset_plan_impl*>(plan);
}

void elb_offset_plan_restart(elb_offset_plan* plan)
{
	reinterpret_cast<elb_offset_plan_impl*>(plan)->plan->restart();
}

void elb_offset_plan_restart_range(elb_offset_plan* plan, uint64_t rangeLen,
	uint64_t rangeOffset)
{
	reinterpret_cast<elb_offset_plan_impl*>(plan)->plan->restart(rangeLen, rangeOffset);
}

int elb_offset_plan_next(elb_offset_plan* plan, uint64_t* outOffset, uint64_t* outLen)
{
	return reinterpret_cast<elb_offset_plan_impl*>(plan)->plan->nextBlock(*outOffset, *outLen) ?
		1 : 0;
}

uint64_t elb_offset_plan_bytes_total(const elb_offset_plan* plan)
{
	return reinterpret_cast<const elb_offset_plan_impl*>(plan)->plan->getNumBytesTotal();
}

uint64_t elb_offset_plan_bytes_left(const elb_offset_plan* plan)
{
	return reinterpret_cast<const elb_offset_plan_impl*>(plan)->plan->getNumBytesLeftToSubmit();
}

void elb_expand_offset_seed(uint64_t seed, uint64_t rank, uint64_t outState[4])
{
	elb::Xoshiro256ss::expandSeed(seed, rank, outState);
}

elb_mgr* elb_mgr_create(const elb_cfg* cfg)
{
	try
	{
		elb_mgr* m = new elb_mgr();
		try
		{
			m->impl = new elb::Manager(cfg);
		}
		catch(...)
		{
			delete m;
			throw;
		}

		return m;
	}
	catch(std::exception& e)
	{
		elb_set_last_error(e.what() );
		return NULL;
	}
}

void elb_mgr_destroy(elb_mgr* m)
{
	if(!m)
		return;

	delete m->impl;
	delete m;
}

int elb_mgr_start_phase(elb_mgr* m, int benchPhase)
{
	try
	{
		m->impl->startNextPhase(benchPhase);
		return 0;
	}
	catch(std::exception& e)
	{
		m->lastError = e.what();
		elb_set_last_error(e.what() );
		return -1;
	}
}

int elb_mgr_wait_done(elb_mgr* m, int timeoutMS)
{
	int waitRes = m->impl->waitForWorkersDone(timeoutMS);

	if(waitRes < 0)
	{
		std::unique_lock<std::mutex> lock(m->impl->shared.mutex);
		m->lastError = m->impl->shared.firstErrorMsg.empty() ?
			"Worker encountered error" : m->impl->shared.firstErrorMsg;
		elb_set_last_error(m->lastError);
	}

	return waitRes;
}

int elb_mgr_run_phase(elb_mgr* m, int benchPhase)
{
	if(elb_mgr_start_phase(m, benchPhase) )
		return -1;

	return (elb_mgr_wait_done(m, -1) == 1) ? 0 : -1;
}

int elb_mgr_live_ops(elb_mgr* m, elb_liveops out[2])
{
	out[0] = elb_liveops{};
	out[1] = elb_liveops{};

	for(const std::unique_ptr<elb::Worker>& worker : m->impl->workers)
	{
		elb::liveOpsAdd(out[0], worker->getLiveOps() );
		elb::liveOpsAdd(out[1], worker->getLiveOpsReadMix() );
	}

	return 0;
}

int elb_mgr_live_latency(elb_mgr* m, elb_livelat* out)
{
	*out = elb_livelat{};

	for(const std::unique_ptr<elb::Worker>& worker : m->impl->workers)
		worker->getAndResetLiveLatency(*out);

	return 0;
}

int elb_mgr_live_snapshot(elb_mgr* m, elb_live_snapshot* out)
{
	m->impl->getLiveSnapshot(*out);
	return 0;
}

const char* elb_mgr_live_reduce_info(elb_mgr* m)
{
	m->liveReduceInfo = m->impl->getLiveReduceInfo();
	return m->liveReduceInfo.c_str();
}

int elb_mgr_phase_results(elb_mgr* m, elb_phase_results* out)
{
	m->impl->getPhaseResults(*out);
	return 0;
}

int elb_mgr_expected_totals(elb_mgr* m, int benchPhase, uint64_t* outEntries,
	uint64_t* outBytes)
{
	m->impl->getExpectedTotals(benchPhase, *outEntries, *outBytes);
	return 0;
}

int elb_mgr_interrupt(elb_mgr* m)
{
	std::unique_lock<std::mutex> lock(m->impl->shared.mutex);

	for(elb::Worker* worker : m->impl->shared.workers)
		worker->interruptExecution();

	m->impl->shared.condition.notify_all();

	return 0;
}

uint32_t elb_mgr_num_workers(elb_mgr* m)
{
	return (uint32_t)m->impl->workers.size();
}

elb_worker* elb_mgr_worker(elb_mgr* m, uint32_t localIdx)
{
	if(localIdx >= m->impl->workers.size() )
		return NULL;

	return reinterpret_cast<elb_worker*>(m->impl->workers[localIdx].get() );
}

const char* elb_mgr_last_error(elb_mgr* m)
{
	return m->lastError.c_str();
}

uint64_t elb_worker_rank(elb_worker* w)
{
	return toWorker(w)->getRank();
}

int elb_worker_gpu_id(elb_worker* w)
{
	return toWorker(w)->getGPUID();
}

int elb_worker_live_ops(elb_worker* w, elb_liveops out[2])
{
	out[0] = toWorker(w)->getLiveOps();
	out[1] = toWorker(w)->getLiveOpsReadMix();
	return 0;
}

int elb_worker_stonewall_ops(elb_worker* w, elb_liveops out[2])
{
	out[0] = toWorker(w)->getStoneWallOps();
	out[1] = toWorker(w)->getStoneWallOpsReadMix();
	return 0;
}

int elb_worker_histogram(elb_worker* w, int kind, elb_histogram* out)
{
	switch(kind)
	{
		case ELB_HISTO_IOPS: *out = toWorker(w)->getIOPSLatHisto(); break;
		case ELB_HISTO_IOPS_READMIX: *out = toWorker(w)->getIOPSLatHistoReadMix(); break;
		case ELB_HISTO_ENTRIES: *out = toWorker(w)->getEntriesLatHisto(); break;
		case ELB_HISTO_ENTRIES_READMIX: *out = toWorker(w)->getEntriesLatHistoReadMix(); break;
		default:
			elb_set_last_error("Invalid histogram kind: " + std::to_string(kind) );
			return -1;
	}

	return 0;
}

uint64_t elb_worker_elapsed_usec(elb_worker* w)
{
	return toWorker(w)->getElapsedUSec();
}

int elb_worker_got_work(elb_worker* w)
{
	return toWorker(w)->getWorkerGotPhaseWork() ? 1 : 0;
}

int elb_worker_dev_counters(elb_worker* w, uint64_t out[ELB_DEVCTR_NUM])
{
	return toWorker(w)->snapshotDevCounters(out);
}

uint64_t* elb_worker_dev_counters_ptr(elb_worker* w)
{
	return toWorker(w)->getDevCountersPtr();
}

const char* elb_worker_last_error(elb_worker* w)
{
	elbWorkerErrorTmp = toWorker(w)->getLastError();
	return elbWorkerErrorTmp.c_str();
}

} // extern "C"
