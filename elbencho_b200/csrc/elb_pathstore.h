/*
 * Custom tree mode (--treefile): the list of directories and files to work on, and its split into
 * per-worker sublists. Reference: source/PathStore.{h,cpp} (line format :24-31 and :75-90, sort
 * orders :192-212, non-shared sublist :258-300, shared sublist :322-437),
 * ProgArgs::loadCustomTreeFile (ProgArgs.cpp:2740-2803), FileTk::scanCustomTree
 * (toolkits/FileTk.cpp:387-470), LocalWorker::prepareCustomTreePathStores
 * (workers/LocalWorker.cpp:1520-1560).
 *
 * Tree file lines:  "d <relative_path>"  and  "f <size_in_bytes> <relative_path>"; anything else
 * is ignored; a "# encoding=base64" header line says that the paths are base64 encoded.
 */
#ifndef ELB_PATHSTORE_H_
#define ELB_PATHSTORE_H_

#include <dirent.h>
#include <stdint.h>
#include <sys/stat.h>

#include <algorithm>
#include <fstream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "elb_host.h"

namespace elb
{

struct PathStoreElem
{
	std::string path;       // relative to the benchmark directory
	uint64_t totalLen{0};   // file size
	uint64_t rangeStart{0}; // this worker's part of the file
	uint64_t rangeLen{0};
};

namespace base64
{
	inline std::string encode(const std::string& raw)
	{
		static const char table[] =
			"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
		std::string out;
		size_t i = 0;

		for( ; (i + 2) < raw.size(); i += 3)
		{
			const unsigned triple = ( (unsigned char)raw[i] << 16) |
				( (unsigned char)raw[i + 1] << 8) | (unsigned char)raw[i + 2];
			out += table[(triple >> 18) & 63];
			out += table[(triple >> 12) & 63];
			out += table[(triple >> 6) & 63];
			out += table[triple & 63];
		}

		if(i < raw.size() )
		{
			const bool haveTwo = ( (i + 1) < raw.size() );
			const unsigned triple = ( (unsigned char)raw[i] << 16) |
				(haveTwo ? ( (unsigned char)raw[i + 1] << 8) : 0);
			out += table[(triple >> 18) & 63];
			out += table[(triple >> 12) & 63];
			out += haveTwo ? table[(triple >> 6) & 63] : '=';
			out += '=';
		}

		return out;
	}

	inline std::string decode(const std::string& encoded)
	{
		std::string out;
		unsigned accumulator = 0;
		int numBits = 0;

		for(const char c : encoded)
		{
			int value;

			if( (c >= 'A') && (c <= 'Z') ) value = c - 'A';
			else if( (c >= 'a') && (c <= 'z') ) value = c - 'a' + 26;
			else if( (c >= '0') && (c <= '9') ) value = c - '0' + 52;
			else if(c == '+') value = 62;
			else if(c == '/') value = 63;
			else continue; // padding or whitespace

			accumulator = (accumulator << 6) | (unsigned)value;
			numBits += 6;

			if(numBits >= 8)
			{
				numBits -= 8;
				out += (char)( (accumulator >> numBits) & 0xFF);
			}
		}

		return out;
	}
}

class PathStore
{
	public:
		void setBlockSize(uint64_t newBlockSize) { blockSize = newBlockSize; }

		uint64_t getNumBlocksTotal() const { return numBlocksTotal; }
		uint64_t getNumBytesTotal() const { return numBytesTotal; }
		size_t getNumPaths() const { return paths.size(); }
		const std::vector<PathStoreElem>& getPaths() const { return paths; }

		void clear()
		{
			paths.clear();
			numBlocksTotal = 0;
			numBytesTotal = 0;
		}

		/* PathStore.cpp:32-73 */
		void loadDirsFromFile(const std::string& treeFilePath)
		{
			const bool isBase64Encoding = checkBase64Encoding(treeFilePath);
			std::ifstream fileStream(treeFilePath.c_str() );

			if(!fileStream)
				throw WorkerError("Opening input file failed: " + treeFilePath);

			std::string lineStr;

			for(unsigned lineNum = 0; std::getline(fileStream, lineStr); lineNum++)
			{
				std::istringstream lineStream(lineStr);
				std::string linePrefixStr;

				lineStream >> linePrefixStr;

				if(linePrefixStr != "d")
					continue;

				PathStoreElem newElem;

				std::getline(lineStream, newElem.path);
				trim(newElem.path);

				if(isBase64Encoding)
					newElem.path = base64::decode(newElem.path);

				if(newElem.path.empty() )
					throw WorkerError("Encountered invalid directory line without path in input "
						"file. File: " + treeFilePath + "; "
						"Line number: " + std::to_string(lineNum) );

				paths.push_back(newElem);
			}
		}

		/* PathStore.cpp:89-155: files with minFileSize <= (rounded up) size <= maxFileSize */
		void loadFilesFromFile(const std::string& treeFilePath, uint64_t minFileSize,
			uint64_t maxFileSize, uint64_t roundUpSize)
		{
			const bool isBase64Encoding = checkBase64Encoding(treeFilePath);
			std::ifstream fileStream(treeFilePath.c_str() );

			if(!fileStream)
				throw WorkerError("Opening input file failed: " + treeFilePath);

			std::string lineStr;

			for(unsigned lineNum = 0; std::getline(fileStream, lineStr); lineNum++)
			{
				std::istringstream lineStream(lineStr);
				std::string linePrefixStr;

				lineStream >> linePrefixStr;

				if(linePrefixStr != "f")
					continue;

				PathStoreElem newElem;

				if(!(lineStream >> newElem.totalLen) )
					throw WorkerError("Encountered invalid file line without size in input file. "
						"File: " + treeFilePath + "; "
						"Line number: " + std::to_string(lineNum) );

				if(roundUpSize && (newElem.totalLen % roundUpSize) )
					newElem.totalLen = newElem.totalLen - (newElem.totalLen % roundUpSize) +
						roundUpSize;

				newElem.rangeLen = newElem.totalLen;

				const uint64_t fileSize = newElem.totalLen;

				if( (fileSize < minFileSize) || (fileSize > maxFileSize) )
					continue;

				std::getline(lineStream, newElem.path);
				trim(newElem.path);

				if(isBase64Encoding)
					newElem.path = base64::decode(newElem.path);

				if(newElem.path.empty() )
					throw WorkerError("Encountered invalid file line without path in input file. "
						"File: " + treeFilePath + "; "
						"Line number: " + std::to_string(lineNum) );

				paths.push_back(newElem);
				numBlocksTotal += numBlocksOf(fileSize);
				numBytesTotal += fileSize;
			}
		}

		/* parents before their subdirs, same order on all hosts (PathStore.cpp:192-197) */
		void sortByPathLen()
		{
			std::stable_sort(paths.begin(), paths.end(),
				[](const PathStoreElem& a, const PathStoreElem& b)
				{
					return (a.path.size() < b.path.size() ) ||
						( (a.path.size() == b.path.size() ) && (a.path < b.path) );
				} );
		}

		/* balance for "every n-th element per worker" (PathStore.cpp:206-211) */
		void sortByFileSize()
		{
			std::stable_sort(paths.begin(), paths.end(),
				[](const PathStoreElem& a, const PathStoreElem& b)
				{
					return (a.totalLen < b.totalLen) ||
						( (a.totalLen == b.totalLen) && (a.path < b.path) );
				} );
		}

		void randomShuffle(uint64_t seed) // PathStore.cpp:217-241 (seed 0 = random_device)
		{
			std::mt19937 generator(seed ? (unsigned)seed : std::random_device()() );
			std::shuffle(paths.begin(), paths.end(), generator);
		}

		/* whole files: elements workerRank, workerRank + n, ... (PathStore.cpp:258-300) */
		void getWorkerSublistNonShared(uint64_t workerRank, uint64_t numDataSetThreads,
			bool throwOnFileSmallerBlock, PathStore& outPathStore) const
		{
			for(size_t idx = workerRank; idx < paths.size(); idx += numDataSetThreads)
			{
				const PathStoreElem& elem = paths[idx];

				if(throwOnFileSmallerBlock && (elem.totalLen < blockSize) )
					throw WorkerError("Found file that is smaller than block size. Consider using "
						"\"--treeroundup\". (\"--nodiocheck\" disables this check.) "
						"File: " + elem.path + "; "
						"FileSize: " + std::to_string(elem.totalLen) + "; "
						"BlockSize: " + std::to_string(blockSize) );

				outPathStore.paths.push_back(elem);
				outPathStore.numBlocksTotal += numBlocksOf(elem.totalLen);
				outPathStore.numBytesTotal += elem.totalLen;
			}
		}

		/* all blocks of all files form one sequence; worker r gets the contiguous share
		   [r * standard, ...) of it, the last worker absorbs the remainder; a file that straddles a
		   boundary is handed out as ranges (PathStore.cpp:322-437) */
		void getWorkerSublistShared(uint64_t workerRank, uint64_t numDataSetThreads,
			bool throwOnSliceSmallerBlock, PathStore& outPathStore) const
		{
			if(paths.empty() )
				return;

			const uint64_t standardWorkerNumBlocks = numBlocksTotal / numDataSetThreads;
			uint64_t thisWorkerNumBlocks = standardWorkerNumBlocks;

			if( (workerRank == (numDataSetThreads - 1) ) && (numBlocksTotal % numDataSetThreads) )
				thisWorkerNumBlocks =
					numBlocksTotal - (standardWorkerNumBlocks * (numDataSetThreads - 1) );

			const uint64_t startBlock = workerRank * standardWorkerNumBlocks;
			const uint64_t endBlock = startBlock + thisWorkerNumBlocks;

			uint64_t currentBlockIdx = 0; // first block of the current file in the global sequence
			uint64_t numBlocksLeft = thisWorkerNumBlocks;

			for(size_t idx = 0; (idx < paths.size() ) && numBlocksLeft; idx++)
			{
				const PathStoreElem& elem = paths[idx];
				const uint64_t fileSize = elem.totalLen;
				const uint64_t numFileBlocks = numBlocksOf(fileSize);

				if( (currentBlockIdx + numFileBlocks) <= startBlock)
				{ // file ends before our share starts
					currentBlockIdx += numFileBlocks;
					continue;
				}

				if(currentBlockIdx >= endBlock)
					break;

				uint64_t rangeStart = 0;
				uint64_t remainingFileBlocks = numFileBlocks;

				if(currentBlockIdx < startBlock)
				{ // our share starts inside this file
					const uint64_t innerFileBlockOffset = startBlock - currentBlockIdx;
					rangeStart = innerFileBlockOffset * blockSize;
					remainingFileBlocks = numFileBlocks - innerFileBlockOffset;
				}

				uint64_t rangeLen;

				if(numBlocksLeft < remainingFileBlocks)
				{ // our share ends inside this file (so not with its possibly partial last block)
					rangeLen = numBlocksLeft * blockSize;
					numBlocksLeft = 0;
				}
				else
				{
					rangeLen = fileSize - rangeStart;
					numBlocksLeft -= remainingFileBlocks;
				}

				PathStoreElem slice = elem;
				slice.rangeStart = rangeStart;
				slice.rangeLen = rangeLen;

				if(throwOnSliceSmallerBlock && (rangeLen < blockSize) )
					throw WorkerError("Found file slice that is smaller than block size. Consider "
						"using \"--treeroundup\". (\"--nodiocheck\" disables this check.) "
						"File: " + elem.path + "; "
						"RangeStart: " + std::to_string(rangeStart) + "; "
						"RangeLength: " + std::to_string(rangeLen) + "; "
						"BlockSize: " + std::to_string(blockSize) );

				outPathStore.paths.push_back(slice);
				outPathStore.numBytesTotal += rangeLen;

				currentBlockIdx += numFileBlocks;
			}

			outPathStore.numBlocksTotal += thisWorkerNumBlocks;
		}

		/* FileTk::scanCustomTree (toolkits/FileTk.cpp:387-470): walk scanPath recursively and
		   write a tree file with base64 encoded relative paths. @return number of entries */
		static uint64_t scanToTreeFile(const std::string& scanPath,
			const std::string& outTreeFilePath, uint64_t& outNumDirs, uint64_t& outNumFiles,
			uint64_t& outNumBytes)
		{
			std::ofstream fileStream(outTreeFilePath, std::ofstream::out | std::ofstream::trunc);

			if(!fileStream)
				throw WorkerError("Opening tree scan results file failed: " + outTreeFilePath);

			fileStream << "# encoding=base64" << std::endl;

			outNumDirs = outNumFiles = outNumBytes = 0;

			std::vector<std::string> pendingDirs(1, ""); // relative paths

			while(!pendingDirs.empty() )
			{
				const std::string relativeDir = pendingDirs.back();
				pendingDirs.pop_back();

				const std::string absoluteDir =
					relativeDir.empty() ? scanPath : (scanPath + "/" + relativeDir);
				DIR* dirHandle = opendir(absoluteDir.c_str() );

				if(!dirHandle)
					throw WorkerError("Unable to scan directory: " + absoluteDir + "; "
						"SysErr: " + strerror(errno) );

				std::vector<std::string> entryNames;

				for(struct dirent* entry = readdir(dirHandle); entry; entry = readdir(dirHandle) )
				{
					const std::string name = entry->d_name;

					if( (name != ".") && (name != "..") )
						entryNames.push_back(name);
				}

				closedir(dirHandle);
				std::sort(entryNames.begin(), entryNames.end() );

				for(const std::string& name : entryNames)
				{
					const std::string relativePath =
						relativeDir.empty() ? name : (relativeDir + "/" + name);
					struct stat statBuf;

					if(lstat( (scanPath + "/" + relativePath).c_str(), &statBuf) == -1)
						continue;

					if(S_ISREG(statBuf.st_mode) )
					{
						outNumFiles++;
						outNumBytes += statBuf.st_size;
						fileStream << "f " << statBuf.st_size << " " <<
							base64::encode(relativePath) << std::endl;
					}
					else
					if(S_ISDIR(statBuf.st_mode) )
					{
						outNumDirs++;
						fileStream << "d " << base64::encode(relativePath) << std::endl;
						pendingDirs.push_back(relativePath);
					}
				}
			}

			return outNumDirs + outNumFiles;
		}

	private:
		uint64_t blockSize{0};
		uint64_t numBlocksTotal{0};
		uint64_t numBytesTotal{0};
		std::vector<PathStoreElem> paths;

		uint64_t numBlocksOf(uint64_t fileSize) const
		{
			if(!blockSize) // (block size can be zero for the dir store)
				return 0;

			return (fileSize / blockSize) + ( (fileSize % blockSize) ? 1 : 0);
		}

		static void trim(std::string& text)
		{
			const char* blanks = " \t\r\n";
			const size_t first = text.find_first_not_of(blanks);
			const size_t last = text.find_last_not_of(blanks);

			text = (first == std::string::npos) ? "" : text.substr(first, last - first + 1);
		}

		/* PathStore.cpp:167-185 */
		static bool checkBase64Encoding(const std::string& treeFilePath)
		{
			std::ifstream fileStream(treeFilePath.c_str() );

			if(!fileStream)
				throw WorkerError("Opening input file failed: " + treeFilePath);

			std::string lineStr;

			while(std::getline(fileStream, lineStr) )
			{
				if(lineStr == "# encoding=base64")
					return true;

				if(!lineStr.empty() && (lineStr[0] != '#') )
					break;
			}

			return false;
		}
};

/* dirs and files of a custom tree (PathStore.h:106-111) */
struct CustomTree
{
	PathStore dirs;
	PathStore filesNonShared; // file size < fileShareSize: whole files round robin over workers
	PathStore filesShared;    // file size >= fileShareSize: block ranges shared between workers

	bool isLoaded{false};

	/* ProgArgs::loadCustomTreeFile (ProgArgs.cpp:2740-2803) */
	void load(const std::string& treeFilePath, uint64_t blockSize, uint64_t fileShareSize,
		uint64_t treeRoundUpSize)
	{
		dirs.clear();
		filesNonShared.clear();
		filesShared.clear();

		dirs.loadDirsFromFile(treeFilePath);
		dirs.sortByPathLen();

		filesNonShared.setBlockSize(blockSize);
		filesNonShared.loadFilesFromFile(treeFilePath, 0, fileShareSize - 1, treeRoundUpSize);
		filesNonShared.sortByFileSize();

		filesShared.setBlockSize(blockSize);
		filesShared.loadFilesFromFile(treeFilePath, fileShareSize, ~0ULL, treeRoundUpSize);

		isLoaded = true;
	}
};

} // namespace elb

#endif /* ELB_PATHSTORE_H_ */
