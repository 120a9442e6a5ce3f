/*
 * TEST INFRASTRUCTURE ONLY - C ABI around the REFERENCE's own PathStore (source/PathStore.{h,cpp}
 * with toolkits/Base64Encoder.cpp and Logger.cpp), compiled from /root/reference/source where they
 * lie into oracle/_ref/libelb_ref.so. Pins the product's custom tree partition
 * (elb_custom_tree_worker_list) to the real implementation.
 */
#include <cstring>
#include <string>

#include "PathStore.h"
#include "ProgException.h"

extern "C" {

/* same text format as elb_custom_tree_worker_list: "<path>\t<totalLen>\t<rangeStart>\t<rangeLen>\n".
 * The tree is loaded the way ProgArgs::loadCustomTreeFile does (ProgArgs.cpp:2740-2803) and split
 * the way LocalWorker::prepareCustomTreePathStores does (LocalWorker.cpp:1520-1560). */
int64_t ref_custom_tree_worker_list(const char* treeFilePath, uint64_t blockSize,
	uint64_t fileShareSize, uint64_t treeRoundUpSize, uint64_t workerRank,
	uint64_t numDataSetThreads, int kind, char* outBuf, uint64_t outBufLen)
{
	try
	{
		if(!fileShareSize)
			fileShareSize = 32 * blockSize; // FILESHAREBLOCKFACTOR (ProgArgs.cpp:52, 1291-1292)

		PathStore sublist;
		sublist.setBlockSize(blockSize);

		if(kind == 0)
		{
			PathStore dirs;
			dirs.loadDirsFromFile(treeFilePath);
			dirs.sortByPathLen();
			dirs.getWorkerSublistNonShared(workerRank, numDataSetThreads, false, sublist);
		}
		else
		{
			PathStore filesNonShared, filesShared;

			filesNonShared.setBlockSize(blockSize);
			filesNonShared.loadFilesFromFile(treeFilePath, 0, fileShareSize - 1, treeRoundUpSize);
			filesNonShared.sortByFileSize();

			filesShared.setBlockSize(blockSize);
			filesShared.loadFilesFromFile(treeFilePath, fileShareSize, ~0ULL, treeRoundUpSize);

			filesNonShared.getWorkerSublistNonShared(workerRank, numDataSetThreads, false, sublist);
			filesShared.getWorkerSublistShared(workerRank, numDataSetThreads, false, sublist);
		}

		std::string text;

		for(const PathStoreElem& elem : sublist.getPaths() )
			text += elem.path + "\t" + std::to_string(elem.totalLen) + "\t" +
				std::to_string(elem.rangeStart) + "\t" + std::to_string(elem.rangeLen) + "\n";

		if(outBuf && outBufLen)
		{
			const size_t copyLen = (text.size() < (outBufLen - 1) ) ? text.size() : (outBufLen - 1);
			memcpy(outBuf, text.data(), copyLen);
			outBuf[copyLen] = 0;
		}

		return (int64_t)text.size();
	}
	catch(ProgException& e)
	{
		if(outBuf && outBufLen)
		{
			strncpy(outBuf, e.what(), outBufLen - 1);
			outBuf[outBufLen - 1] = 0;
		}

		return -1;
	}
}

} // extern "C"
