/*
 * Coordinator: what the reference's main() + Coordinator do for a local run
 * (source/Main.cpp:13-68, source/Coordinator.cpp:31-142 main, :248-272 runBenchmarkPhase,
 * :298-374 runBenchmarks, :380-413 runSyncAndDropCaches, :418-440 signal handling) on top of the
 * native manager, plus the live statistics loop (source/Statistics.cpp:180-300, 1232-1345).
 *
 * Exported as elb_cli_main(argc, argv) so that the tiny elbencho-b200 executable, tests and any
 * embedding program share one implementation.
 */
#include <signal.h>
#include <sys/ioctl.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>

#include "elb_cli.h"
#include "elb_service.h"
#include "elb_worker.h"

namespace elb
{

static volatile sig_atomic_t gotUserInterruptSignal = 0;

static void interruptSignalHandler(int signal) // Coordinator.cpp:418-440
{
	(void)signal;
	gotUserInterruptSignal = 1;
}

static std::string isoDateNow(bool withMillis)
{
	struct timespec now;
	clock_gettime(CLOCK_REALTIME, &now);

	struct tm localTimeInfo;
	localtime_r(&now.tv_sec, &localTimeInfo);

	char dateBuf[64];
	char zoneBuf[16];
	strftime(dateBuf, sizeof(dateBuf), "%FT%T", &localTimeInfo);
	strftime(zoneBuf, sizeof(zoneBuf), "%z", &localTimeInfo);

	std::ostringstream out;
	out << dateBuf;

	if(withMillis)
		out << "." << std::setfill('0') << std::setw(3) << (now.tv_nsec / 1000000);

	out << zoneBuf;

	return out.str();
}

std::string stats::elapsedSecToHumanStr(uint64_t elapsedSec) // UnitTk::elapsedSecToHumanStr
{
	std::ostringstream out;
	const uint64_t numHours = elapsedSec / 3600;
	const uint64_t numMin = (elapsedSec % 3600) / 60;
	const uint64_t numSec = elapsedSec % 60;

	if(numHours)
		out << numHours << "h";
	if(numHours || numMin)
		out << numMin << "m";
	out << numSec << "s";

	return out.str();
}

class Coordinator
{
	public:
		explicit Coordinator(ProgArgs& progArgs) : progArgs(progArgs) {}

		int main();

	private:
		ProgArgs& progArgs;
		std::unique_ptr<Manager> manager;
		CPUUtil liveCpuUtil;
		uint64_t phaseCounter{0};

		void runBenchmarks();
		void runBenchmarkPhase(int benchPhase);
		void runSyncAndDropCaches();
		bool liveCSVHeaderPrinted{false};
		bool isPhaseTimeExpired{false}; // WorkersSharedData::isPhaseTimeExpired
		void printLiveStatsCSV(int benchPhase, const elb_liveops liveOps[2],
			const elb_liveops oldLiveOps[2], const elb_livelat& liveLat, uint64_t intervalUSec,
			size_t numWorkersDone, uint64_t expectedEntries, uint64_t expectedBytes,
			uint64_t elapsedMS);
		void printLiveStatsLine(int benchPhase, const elb_liveops liveOps[2],
			const elb_liveops oldLiveOps[2], uint64_t intervalUSec, size_t numWorkersDone,
			uint64_t elapsedSec);
		void printPhaseResultsEverywhere(int benchPhase, const std::string& isoStartDate);
		void printDryRunInfo();
};

/* Statistics::printDryRunInfo (:2809-2846) */
void Coordinator::printDryRunInfo()
{
	ProgArgs::ABIConfig abiConfig;
	progArgs.toABIConfig(abiConfig);

	if(abiConfig.gpuIDs.empty() )
	{ // "all" or master mode: the GPU list does not matter for the expected totals
		abiConfig.gpuIDs.push_back(0);
		abiConfig.cfg.gpuIDs = abiConfig.gpuIDs.data();
		abiConfig.cfg.numGPUIDs = 1;
	}

	abiConfig.cfg.fileSize = detectFileSize(&abiConfig.cfg);

	Config cfg = Config::fromABI(&abiConfig.cfg);

	TreeManifest customTree;

	if(!cfg.treeFilePath.empty() )
		customTree.load(cfg.treeFilePath, cfg.blockSize, cfg.fileShareSize, cfg.treeRoundUpSize);

	const int phases[] = {ELB_PHASE_CREATEDIRS, ELB_PHASE_DELETEDIRS, ELB_PHASE_CREATEFILES,
		ELB_PHASE_READFILES, ELB_PHASE_DELETEFILES, ELB_PHASE_STATFILES};
	const bool enabled[] = {progArgs.runCreateDirsPhase, progArgs.runDeleteDirsPhase,
		progArgs.runCreateFilesPhase, progArgs.runReadPhase, progArgs.runDeleteFilesPhase,
		progArgs.runStatFilesPhase};

	for(size_t i = 0; i < (sizeof(phases) / sizeof(phases[0] ) ); i++)
	{
		if(!enabled[i] )
			continue;

		uint64_t entriesPerThread, bytesPerThread;
		expectedPerWorker(cfg, phases[i], entriesPerThread, bytesPerThread, &customTree);

		stats::printDryRunPhaseInfo(progArgs, phases[i], entriesPerThread, bytesPerThread,
			std::cout);
	}
}

/* Statistics::printSingleLineLiveStatsLine (:180-285), brief form on one console line */
void Coordinator::printLiveStatsLine(int benchPhase, const elb_liveops liveOps[2],
	const elb_liveops oldLiveOps[2], uint64_t intervalUSec, size_t numWorkersDone,
	uint64_t elapsedSec)
{
	const uint64_t mib = 1024 * 1024;
	const bool isRWMixPhase = (liveOps[1].numBytesDone || liveOps[1].numEntriesDone);
	const bool isDirMode = (progArgs.benchPathType == ELB_PATH_DIR);
	const std::string entryType = stats::phaseEntryType(benchPhase, false);

	auto perSec = [&](uint64_t newVal, uint64_t oldVal)
		{ return intervalUSec ? perSecFromUSec(newVal - oldVal, intervalUSec) : 0; };

	std::ostringstream line;
	line << "\x1b[2K\r" << stats::phaseName(benchPhase, progArgs) << ": ";

	if(isDirMode)
	{
		if(!isRWMixPhase)
			line << perSec(liveOps[0].numEntriesDone, oldLiveOps[0].numEntriesDone) << " " <<
				entryType << "/s; " <<
				perSec(liveOps[0].numBytesDone, oldLiveOps[0].numBytesDone) / mib << " MiB/s; ";
		else
			line << "wr=[" <<
				perSec(liveOps[0].numEntriesDone, oldLiveOps[0].numEntriesDone) << " " <<
				entryType << "/s; " <<
				perSec(liveOps[0].numBytesDone, oldLiveOps[0].numBytesDone) / mib << " MiB/s] "
				"rd=[" <<
				perSec(liveOps[1].numBytesDone, oldLiveOps[1].numBytesDone) / mib << " MiB/s]; ";

		line << (liveOps[0].numEntriesDone + liveOps[1].numEntriesDone) << " " << entryType <<
			"; " << (liveOps[0].numBytesDone + liveOps[1].numBytesDone) / mib << " MiB; ";
	}
	else
	{
		if(!isRWMixPhase)
			line << perSec(liveOps[0].numIOPSDone, oldLiveOps[0].numIOPSDone) << " IOPS; " <<
				perSec(liveOps[0].numBytesDone, oldLiveOps[0].numBytesDone) / mib << " MiB/s; " <<
				liveOps[0].numBytesDone / mib << " MiB; ";
		else
			line << "wr=[" <<
				perSec(liveOps[0].numIOPSDone, oldLiveOps[0].numIOPSDone) << " IOPS; " <<
				perSec(liveOps[0].numBytesDone, oldLiveOps[0].numBytesDone) / mib << " MiB/s; " <<
				liveOps[0].numBytesDone / mib << " MiB] rd=[" <<
				perSec(liveOps[1].numIOPSDone, oldLiveOps[1].numIOPSDone) << " IOPS; " <<
				perSec(liveOps[1].numBytesDone, oldLiveOps[1].numBytesDone) / mib << " MiB/s; " <<
				liveOps[1].numBytesDone / mib << " MiB]; ";
	}

	liveCpuUtil.update();

	line << (manager->workers.size() - numWorkersDone) << " threads; " <<
		liveCpuUtil.getCPUUtilPercent() << "% CPU; " << stats::elapsedSecToHumanStr(elapsedSec);

	std::string lineStr = line.str();

	struct winsize consoleSize;
	if( (ioctl(STDOUT_FILENO, TIOCGWINSZ, &consoleSize) == 0) && consoleSize.ws_col)
	{
		const size_t usableLineLen = consoleSize.ws_col + 5 - 2; // 5 hidden control chars
		if(lineStr.length() > usableLineLen)
			lineStr.resize(usableLineLen);
	}

	std::cout << lineStr << std::flush;
}

/* Statistics::prepLiveCSVFile + printLiveStatsCSV (:2944-3113): one "Total" line per interval
 * (plus a "Read" line in rwmix phases) with the aggregate of all local workers */
void Coordinator::printLiveStatsCSV(int benchPhase, const elb_liveops liveOps[2],
	const elb_liveops oldLiveOps[2], const elb_livelat& liveLat, uint64_t intervalUSec,
	size_t numWorkersDone, uint64_t expectedEntries, uint64_t expectedBytes, uint64_t elapsedMS)
{
	const bool toStdout = (progArgs.liveCSVFilePath == "stdout");
	std::ofstream fileStream;
	bool printHeaders = toStdout ? !liveCSVHeaderPrinted : false;

	if(!toStdout)
	{
		struct stat statBuf;
		printHeaders = (stat(progArgs.liveCSVFilePath.c_str(), &statBuf) != 0) ||
			!statBuf.st_size;

		fileStream.open(progArgs.liveCSVFilePath, std::ofstream::app);

		if(!fileStream)
			throw ProgError("Unable to open live stats csv file: " + progArgs.liveCSVFilePath);
	}

	std::ostream& out = toStdout ? std::cout : fileStream;

	if(printHeaders)
		out << "ISO Date,Label,Phase,RuntimeMS,Rank,MixType,Done%,DoneBytes,MiB/s,IOPS,Entries,"
			"Entries/s,Lat Ent us,Lat IO us,Active,CPU,Service," << std::endl;

	liveCSVHeaderPrinted = true;

	const bool isRWMixPhase = (liveOps[1].numBytesDone || liveOps[1].numEntriesDone);
	const bool isDirMode = (progArgs.benchPathType == ELB_PATH_DIR);
	const std::string isoDate = isoDateNow(true);
	const std::string phaseName = stats::phaseName(benchPhase, progArgs);
	std::string label = progArgs.benchLabel;
	std::replace(label.begin(), label.end(), ',', ' ');

	liveCpuUtil.update();

	auto perSec = [&](uint64_t newVal, uint64_t oldVal)
		{ return intervalUSec ? perSecFromUSec(newVal - oldVal, intervalUSec) : 0; };
	auto percentDone = [&](const elb_liveops& ops)
	{
		uint64_t percent = 0;

		if(expectedBytes)
			percent = (100 * ops.numBytesDone) / expectedBytes;
		else
		if(expectedEntries)
			percent = (100 * ops.numEntriesDone) / expectedEntries;

		return std::min(percent, (uint64_t)100);
	};
	auto avg = [](uint64_t sum, uint64_t num) { return num ? (sum / num) : 0; };

	for(int mixIdx = 0; mixIdx < (isRWMixPhase ? 2 : 1); mixIdx++)
	{
		const elb_liveops& ops = liveOps[mixIdx];
		const elb_liveops& oldOps = oldLiveOps[mixIdx];

		out << isoDate << "," << label << "," << phaseName << "," << elapsedMS << "," <<
			"Total" << "," << (isRWMixPhase ? (mixIdx ? "Read" : "Write") : "") << "," <<
			percentDone(ops) << "," << ops.numBytesDone << "," <<
			(perSec(ops.numBytesDone, oldOps.numBytesDone) / (1024 * 1024) ) << "," <<
			perSec(ops.numIOPSDone, oldOps.numIOPSDone) << "," <<
			(isDirMode ? ops.numEntriesDone : 0) << "," <<
			(isDirMode ? perSec(ops.numEntriesDone, oldOps.numEntriesDone) : 0) << "," <<
			(mixIdx ? avg(liveLat.avgEntriesLatReadMixMicrosSecsSum,
				liveLat.numAvgEntriesLatReadMixValues) :
				avg(liveLat.avgEntriesLatMicroSecsSum, liveLat.numAvgEntriesLatValues) ) << "," <<
			(mixIdx ? avg(liveLat.avgIOLatReadMixMicroSecsSum, liveLat.numAvgIOLatReadMixValues) :
				avg(liveLat.avgIOLatMicroSecsSum, liveLat.numAvgIOLatValues) ) << "," <<
			(manager->workers.size() - numWorkersDone) << "," <<
			liveCpuUtil.getCPUUtilPercent() << "," << "" << "," << std::endl;
	}

	if(!progArgs.useExtendedLiveCSV)
		return;

	/* --livecsvex: one more line per worker with its totals so far; no per-second and latency
	   values there (Statistics.cpp:3116-3225) */
	const size_t numWorkers = manager->workers.size();
	const uint64_t bytesPerWorker = numWorkers ? (expectedBytes / numWorkers) : 0;
	const uint64_t entriesPerWorker = numWorkers ? (expectedEntries / numWorkers) : 0;

	for(size_t workerIdx = 0; workerIdx < numWorkers; workerIdx++)
	{
		const elb_liveops workerOps[2] = { manager->workers[workerIdx]->getLiveOps(),
			manager->workers[workerIdx]->getLiveOpsReadMix() };

		for(int mixIdx = 0; mixIdx < (isRWMixPhase ? 2 : 1); mixIdx++)
		{
			const elb_liveops& ops = workerOps[mixIdx];
			uint64_t workerPercentDone = 0;

			if(bytesPerWorker)
				workerPercentDone = (100 * ops.numBytesDone) / bytesPerWorker;
			else
			if(entriesPerWorker)
				workerPercentDone = (100 * ops.numEntriesDone) / entriesPerWorker;

			out << isoDate << "," << label << "," << phaseName << "," << elapsedMS << "," <<
				workerIdx << "," << (isRWMixPhase ? (mixIdx ? "Read" : "Write") : "") << "," <<
				std::min(workerPercentDone, (uint64_t)100) << "," << ops.numBytesDone << "," <<
				"" << "," << "" << "," << (isDirMode ? ops.numEntriesDone : 0) << "," <<
				"" << "," << "" << "," << "" << "," << ",,," << std::endl;
		}
	}
}

/* Statistics::printPhaseResults (:1568-1632): console + optional txt/csv/json files */
void Coordinator::printPhaseResultsEverywhere(int benchPhase, const std::string& isoStartDate)
{
	elb_phase_results res;
	manager->getPhaseResults(res);

	std::vector<uint64_t> elapsedUSecVec;
	for(const std::unique_ptr<Worker>& worker : manager->workers)
		if(worker->getElapsedUSec() )
			elapsedUSecVec.push_back(worker->getElapsedUSec() );

	const bool haveResults = !elapsedUSecVec.empty();

	if(!haveResults)
		std::cout << "Skipping stats print due to unavailable worker results." << std::endl;
	else
		stats::printPhaseResults(progArgs, benchPhase, res, elapsedUSecVec, std::cout);

	if(!progArgs.resFilePath.empty() )
	{
		std::ofstream fileStream(progArgs.resFilePath, std::ofstream::app);

		if(!fileStream)
			std::cerr << "ERROR: Opening results file failed: " << progArgs.resFilePath <<
				std::endl;
		else
		{
			if(!haveResults)
				fileStream << "Skipping stats print due to unavailable worker results." <<
					std::endl;
			else
				stats::printPhaseResults(progArgs, benchPhase, res, elapsedUSecVec, fileStream);

			fileStream << std::endl;
		}
	}

	if(haveResults && !progArgs.csvFilePath.empty() )
	{
		std::vector<std::string> labels, values;
		stats::csvLabelsAndValues(progArgs, benchPhase, res, isoDateNow(false), labels, values);

		// labels line only for a new/empty file (Coordinator.cpp checkCSVFileCompatibility)
		bool needLabels = !progArgs.noCSVLabels;
		{
			std::ifstream existing(progArgs.csvFilePath);
			if(existing && (existing.peek() != std::ifstream::traits_type::eof() ) )
				needLabels = false;
		}

		std::ofstream fileStream(progArgs.csvFilePath, std::ofstream::app);

		if(!fileStream)
			std::cerr << "ERROR: Opening results CSV file failed: " << progArgs.csvFilePath <<
				std::endl;
		else
		{
			auto joinCSV = [](const std::vector<std::string>& vec)
			{
				std::string line;
				for(size_t i = 0; i < vec.size(); i++)
					line += (i ? "," : "") + vec[i];
				return line;
			};

			if(needLabels)
				fileStream << joinCSV(labels) << std::endl;

			fileStream << joinCSV(values) << std::endl;
		}
	}

	if(haveResults && !progArgs.jsonFilePath.empty() )
	{
		std::ofstream fileStream(progArgs.jsonFilePath, std::ofstream::app);

		if(!fileStream)
			std::cerr << "ERROR: Opening results JSON file failed: " << progArgs.jsonFilePath <<
				std::endl;
		else
			fileStream << stats::phaseResultsJSON(progArgs, benchPhase, res, phaseCounter,
				isoStartDate) << std::endl;
	}
}

/* Coordinator::runBenchmarkPhase (:248-272) + Statistics live loop (:1284-1345) */
void Coordinator::runBenchmarkPhase(int benchPhase)
{
	// don't start the next phase if the time limit of the previous one expired (:234-241)
	if(isPhaseTimeExpired)
		throw ProgTimeLimit();

	const std::string isoStartDate = isoDateNow(true);
	const Clock::time_point phaseStartT = Clock::now();

	phaseCounter++;
	manager->startNextPhase(benchPhase);

	elb_liveops oldLiveOps[2] = {};
	Clock::time_point lastLiveT = phaseStartT;
	bool printedLiveLine = false;
	int waitRes;

	liveCpuUtil.update();

	const bool showLiveLine = !progArgs.disableLiveStats &&
		(isatty(STDOUT_FILENO) || getenv("ELB_FORCE_LIVESTATS") );
	const bool writeLiveCSV = !progArgs.liveCSVFilePath.empty();
	const bool showLive = showLiveLine || writeLiveCSV;
	const bool useLiveReduce = (manager->getNumGPUs() >= 2);

	uint64_t expectedEntries = 0, expectedBytes = 0; // of all local workers, for "Done%"
	manager->getExpectedTotals(benchPhase, expectedEntries, expectedBytes);

	for( ; ; )
	{
		waitRes = manager->waitForWorkersDone( (int)progArgs.liveStatsSleepMS);

		if(waitRes)
			break;

		if(gotUserInterruptSignal)
		{ // (Coordinator.cpp:54-60: interrupt workers, results of this phase are not printed)
			manager->interruptAndNotifyWorkers();
			throw ProgError("Received interrupt signal. Stopping workers...");
		}

		const uint64_t elapsedSec = std::chrono::duration_cast<std::chrono::seconds>(
			Clock::now() - phaseStartT).count();

		if(progArgs.timeLimitSecs && (elapsedSec >= progArgs.timeLimitSecs) )
		{ // WorkerManager::checkPhaseTimeLimit (:109-128): friendly interruption, results stay
			isPhaseTimeExpired = true;
			std::unique_lock<std::mutex> lock(manager->shared.mutex);
			for(Worker* worker : manager->shared.workers)
				worker->interruptExecution();
		}

		if(!showLive)
			continue;

		elb_liveops liveOps[2] = {};
		elb_livelat liveLat = {};
		size_t numWorkersDone;

		if(useLiveReduce)
		{ // workers on several GPUs: per-GPU partial sums, reduced over NVLink by NCCL
			elb_live_snapshot snapshot;
			manager->getLiveSnapshot(snapshot);
			liveOps[0] = snapshot.ops;
			liveOps[1] = snapshot.opsReadMix;
			liveLat = snapshot.lat;
		}
		else
			for(const std::unique_ptr<Worker>& worker : manager->workers)
			{
				liveOpsAdd(liveOps[0], worker->getLiveOps() );
				liveOpsAdd(liveOps[1], worker->getLiveOpsReadMix() );

				if(writeLiveCSV)
					worker->getAndResetLiveLatency(liveLat);
			}

		{
			std::unique_lock<std::mutex> lock(manager->shared.mutex);
			numWorkersDone = manager->shared.numWorkersDone;
		}

		const Clock::time_point nowT = Clock::now();
		const uint64_t intervalUSec =
			std::chrono::duration_cast<std::chrono::microseconds>(nowT - lastLiveT).count();

		if(writeLiveCSV)
			printLiveStatsCSV(benchPhase, liveOps, oldLiveOps, liveLat, intervalUSec,
				numWorkersDone, expectedEntries, expectedBytes,
				std::chrono::duration_cast<std::chrono::milliseconds>(nowT - phaseStartT).count() );

		if(showLiveLine)
		{
			printLiveStatsLine(benchPhase, liveOps, oldLiveOps, intervalUSec, numWorkersDone,
				elapsedSec);

			if(progArgs.useBriefLiveStatsNewLine)
				std::cout << std::endl; // --live1n
			else
				printedLiveLine = true;
		}

		oldLiveOps[0] = liveOps[0];
		oldLiveOps[1] = liveOps[1];
		lastLiveT = nowT;
	}

	if(printedLiveLine)
		std::cout << "\x1b[2K\r" << std::flush; // delete the live stats line

	if(waitRes < 0)
	{
		std::string errMsg;
		{
			std::unique_lock<std::mutex> lock(manager->shared.mutex);
			errMsg = manager->shared.firstErrorMsg;
		}

		if(!errMsg.empty() )
			std::cerr << "ERROR: " << errMsg << std::endl;

		throw ProgError("Worker encountered error"); // WorkerManager.cpp:54-58
	}

	if( (benchPhase != ELB_PHASE_SYNC) && (benchPhase != ELB_PHASE_DROPCACHES) )
		printPhaseResultsEverywhere(benchPhase, isoStartDate);
	else
	{ // (sync and dropcache phases print only their elapsed time)
		elb_phase_results res;
		manager->getPhaseResults(res);
		std::vector<uint64_t> elapsedUSecVec(1, res.lastFinishUSec);
		stats::printPhaseResults(progArgs, benchPhase, res, elapsedUSecVec, std::cout);
	}
}

void waitForUserDefinedStartTime(const ProgArgs& progArgs) // Coordinator.cpp:149-158
{
	if(!progArgs.startTime)
		return;

	if(time(NULL) > (time_t)progArgs.startTime)
		throw ProgError("Defined start time has already passed. Aborting.");

	if(!progArgs.disableLiveStats)
		std::cout << "Waiting for start time: " << (progArgs.startTime - time(NULL) ) << "s" <<
			std::endl;

	while(time(NULL) < (time_t)progArgs.startTime)
		usleep(1000);
}

void Coordinator::runSyncAndDropCaches() // Coordinator.cpp:380-413
{
	if(progArgs.runSyncPhase)
		runBenchmarkPhase(ELB_PHASE_SYNC);

	if(progArgs.runDropCachesPhase)
		runBenchmarkPhase(ELB_PHASE_DROPCACHES);
}

void Coordinator::runBenchmarks() // Coordinator.cpp:298-374
{
	struct BenchPhaseConfig { int benchPhase; bool runPhase; };

	const BenchPhaseConfig allBenchPhases[] =
	{
		{ELB_PHASE_CREATEDIRS, progArgs.runCreateDirsPhase},
		{ELB_PHASE_CREATEFILES, progArgs.runCreateFilesPhase},
		{ELB_PHASE_STATFILES, progArgs.runStatFilesPhase},
		{ELB_PHASE_READFILES, progArgs.runReadPhase},
		{ELB_PHASE_DELETEFILES, progArgs.runDeleteFilesPhase},
		{ELB_PHASE_DELETEDIRS, progArgs.runDeleteDirsPhase},
	};

	std::vector<int> enabledPhases;
	for(const BenchPhaseConfig& phaseConfig : allBenchPhases)
		if(phaseConfig.runPhase)
			enabledPhases.push_back(phaseConfig.benchPhase);

	for(uint64_t iterationIndex = 0; iterationIndex < progArgs.iterations; iterationIndex++)
	{
		if(progArgs.iterations > 1)
			std::cout << "[Starting iteration " << (iterationIndex + 1) << " of " <<
				progArgs.iterations << "...]" << std::endl;

		stats::printPhaseResultsTableHeader(std::cout);

		runSyncAndDropCaches();

		for(size_t phaseIdx = 0; phaseIdx < enabledPhases.size(); phaseIdx++)
		{
			runBenchmarkPhase(enabledPhases[phaseIdx] );

			runSyncAndDropCaches();

			if( (phaseIdx < (enabledPhases.size() - 1) ) && progArgs.nextPhaseDelaySecs)
				sleep( (unsigned)progArgs.nextPhaseDelaySecs);
		}
	}
}

int Coordinator::main() // Coordinator.cpp:31-142
{
	try
	{
		if(progArgs.runAsService)
			return serviceMain(progArgs);

		if(progArgs.interruptServices || progArgs.quitServices)
			return masterInterruptOrQuitServices(progArgs);

		if(!progArgs.treeScanPath.empty() )
		{ // ProgArgs::scanCustomTree (ProgArgs.cpp:2674-2735): write the tree file first
			uint64_t numDirs, numFiles, numBytes;

			TreeManifest::scanToTreeFile(progArgs.treeScanPath, progArgs.treeFilePath, numDirs,
				numFiles, numBytes);

			std::cout << "Directory scan done. Dirs: " << numDirs << "; Files: " << numFiles <<
				"; Bytes: " << numBytes << "; Treefile: " << progArgs.treeFilePath << std::endl;

			if(progArgs.benchPaths.empty() )
				return EXIT_SUCCESS;
		}

		if(progArgs.doDryRun)
		{
			printDryRunInfo();
			return EXIT_SUCCESS;
		}

		if(!progArgs.hosts.empty() )
			return masterMain(progArgs);

		if(progArgs.gpuIDsStr == "all")
		{ // ProgArgs.cpp:2540-2553
			int numGPUs = 0;
			cudaError_t countRes = cudaGetDeviceCount(&numGPUs);

			if( (countRes != cudaSuccess) || !numGPUs)
				throw ProgError(std::string("No GPUs found for \"--gpuids all\". CUDA Error: ") +
					cudaGetErrorString(countRes) );

			for(int gpuID = 0; gpuID < numGPUs; gpuID++)
				progArgs.gpuIDs.push_back(gpuID);
		}

		ProgArgs::ABIConfig abiConfig;
		progArgs.toABIConfig(abiConfig);

		manager.reset(new Manager(&abiConfig.cfg) );

		/* workers on several GPUs: create the NCCL communicators of the statistics reduce now,
		   while the workers are idle (and before any result output) */
		if(manager->getNumGPUs() >= 2)
			manager->getLiveReduceInfo();

		// the normalised values are what results files report (e.g. block size reduced to file size)
		progArgs.blockSize = manager->shared.cfg.blockSize;
		progArgs.fileSize = manager->shared.cfg.fileSize;

		struct sigaction sigAction;
		memset(&sigAction, 0, sizeof(sigAction) );
		sigAction.sa_handler = interruptSignalHandler;
		sigaction(SIGINT, &sigAction, NULL);
		sigaction(SIGTERM, &sigAction, NULL);

		waitForUserDefinedStartTime(progArgs);

		runBenchmarks();

		manager.reset();
	}
	catch(ProgTimeLimit& e)
	{ // a user-defined time limit, not an error (Coordinator.cpp:111-116)
		std::cout << e.what() << std::endl;
		manager.reset();
	}
	catch(std::exception& e)
	{
		std::cerr << "ERROR: " << e.what() << std::endl;
		manager.reset();
		return EXIT_FAILURE;
	}

	return EXIT_SUCCESS;
}

} // namespace elb

extern "C" int64_t elb_format_phase_results(int argc, char** argv, int benchPhase,
	const elb_phase_results* results, int format, char* outBuf, uint64_t outBufLen)
{
	try
	{
		elb::ProgArgs progArgs(argc, argv);
		std::ostringstream out;

		if(format == 0)
		{
			std::vector<uint64_t> elapsedUSecVec;
			elapsedUSecVec.push_back(results->firstFinishUSec);
			if(results->lastFinishUSec != results->firstFinishUSec)
				elapsedUSecVec.push_back(results->lastFinishUSec);

			elb::stats::printPhaseResultsTableHeader(out);
			elb::stats::printPhaseResults(progArgs, benchPhase, *results, elapsedUSecVec, out);
		}
		else
		if(format == 1)
		{
			std::vector<std::string> labels, values;
			elb::stats::csvLabelsAndValues(progArgs, benchPhase, *results,
				"2026-01-01T00:00:00+0000", labels, values);

			for(size_t i = 0; i < labels.size(); i++)
				out << (i ? "," : "") << labels[i];
			out << std::endl;
			for(size_t i = 0; i < values.size(); i++)
				out << (i ? "," : "") << values[i];
			out << std::endl;
		}
		else
		if(format == 2)
			out << elb::stats::phaseResultsJSON(progArgs, benchPhase, *results, 1,
				"2026-01-01T00:00:00.000+0000") << std::endl;
		else
			throw elb::ProgError("Invalid format: " + std::to_string(format) );

		const std::string text = out.str();

		if(outBuf && outBufLen)
		{
			const size_t copyLen = std::min( (size_t)(outBufLen - 1), text.size() );
			memcpy(outBuf, text.data(), copyLen);
			outBuf[copyLen] = 0;
		}

		return (int64_t)text.size();
	}
	catch(std::exception& e)
	{
		elb_set_last_error(e.what() );
		return -1;
	}
}

/* kind 0: UnitTk::latencyUsToHumanStr, 1: elapsedMSToHumanStr, 2: elapsedSecToHumanStr,
 * 3: LatencyHistogram::getHistogramStr, 4: getPercentileStr(percentage) */
extern "C" int64_t elb_format_value(int kind, uint64_t value, double percentage,
	const elb_histogram* histo, char* outBuf, uint64_t outBufLen)
{
	std::string text;

	switch(kind)
	{
		case 0: text = elb::stats::latencyUsToHumanStr(value); break;
		case 1: text = elb::stats::elapsedMSToHumanStr(value); break;
		case 2: text = elb::stats::elapsedSecToHumanStr(value); break;
		case 3:
		case 4:
		{
			if(!histo)
			{
				elb_set_last_error("elb_format_value: histogram missing");
				return -1;
			}

			text = (kind == 3) ?
				elb::stats::histogramStr(*histo) : elb::stats::percentileStr(*histo, percentage);
		} break;

		default:
			elb_set_last_error("elb_format_value: invalid kind " + std::to_string(kind) );
			return -1;
	}

	if(outBuf && outBufLen)
	{
		const size_t copyLen = std::min( (size_t)(outBufLen - 1), text.size() );
		memcpy(outBuf, text.data(), copyLen);
		outBuf[copyLen] = 0;
	}

	return (int64_t)text.size();
}

/* UnitTk::numHumanToBytesBinary (toolkits/UnitTk.cpp:18-76); error text via elb_last_error() */
extern "C" int elb_num_human_to_bytes(const char* numHuman, uint64_t* outBytes)
{
	try
	{
		*outBytes = elb::ProgArgs::numHumanToBytesBinary(numHuman ? numHuman : "");
		return 0;
	}
	catch(std::exception& e)
	{
		elb_set_last_error(e.what() );
		return -1;
	}
}

extern "C" void elb_simple128_hash(const char* input, char out[33])
{
	const std::string hash = elb::ProgArgs::simple128Hash(input ? input : "");
	memcpy(out, hash.c_str(), std::min(hash.size() + 1, (size_t)33) );
	out[32] = 0;
}

extern "C" int elb_cli_main(int argc, char** argv)
{
	try
	{
		elb::ProgArgs progArgs(argc, argv);

		if(progArgs.printHelp)
		{
			std::cout << elb::ProgArgs::helpText();
			return EXIT_SUCCESS;
		}

		if(progArgs.printVersion)
		{
			std::cout << "elbencho-b200 (GPU worker for NVIDIA Blackwell, sm_100a)" << std::endl;
			std::cout << "Compatible with: elbencho 3.1-4 (service protocol 3.1.1)" << std::endl;
			std::cout << "Included optional build features: cuda cufile(dlopen) aio(raw syscalls)" <<
				std::endl;
			return EXIT_SUCCESS;
		}

		elb::Coordinator coordinator(progArgs);

		return coordinator.main();
	}
	catch(elb::ProgError& e)
	{
		std::cerr << "ERROR: " << e.what() << std::endl;
		return EXIT_FAILURE;
	}
	catch(std::exception& e)
	{
		std::cerr << "ERROR: " << e.what() << std::endl;
		return EXIT_FAILURE;
	}
}
