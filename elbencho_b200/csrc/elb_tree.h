/*
 * Custom tree mode (--treefile): the manifest of directories and files to work on, and each
 * worker's share of it.
 *
 * What has to equal the reference (checked against the reference's own PathStore.cpp, compiled in
 * place into oracle/_ref, by tests/test_custom_tree.py): which worker gets which file or which
 * byte range of a file, in which order (source/PathStore.cpp:192-212 sort orders, :258-300 whole
 * files, :322-437 shared files; ProgArgs::loadCustomTreeFile, ProgArgs.cpp:2740-2803), the tree
 * file format (PathStore.cpp:24-31, 75-90) and the error texts.
 *
 * How it is built here: the tree file is parsed ONCE into a TreeManifest (the reference reads it
 * three times into three stores). Files below --sharesize are kept sorted by size and handed out
 * with a stride; files from --sharesize on keep their file order together with the prefix sums of
 * their block counts, so a worker finds its first and last shared file with a binary search over
 * the prefix sums and cuts its block range [rank * standard, ...) directly, instead of walking the
 * whole list and counting blocks down.
 *
 * Tree file lines:  "d <relative_path>"  and  "f <size_in_bytes> <relative_path>"; anything else
 * is ignored; a "# encoding=base64" line ahead of the first entry says that the paths are base64
 * encoded.
 */
#ifndef ELB_TREE_H_
#define ELB_TREE_H_

#include <dirent.h>
#include <stdint.h>
#include <sys/stat.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <random>
#include <string>
#include <tuple>
#include <vector>

#include "elb_host.h"

namespace elb
{

/* a file (or directory) of the tree, or the byte range of a file that one worker works on */
struct TreeSlice
{
	std::string path;       // relative to the benchmark directory
	uint64_t totalLen{0};   // file size
	uint64_t rangeStart{0}; // this worker's part of the file
	uint64_t rangeLen{0};

	bool coversWholeFile() const { return rangeLen == totalLen; }
};

namespace base64
{
	inline std::string encode(const std::string& raw)
	{
		static const char alphabet[] =
			"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
		std::string out;
		out.reserve( (raw.size() + 2) / 3 * 4);

		for(size_t pos = 0; pos < raw.size(); pos += 3)
		{
			const size_t numIn = std::min<size_t>(3, raw.size() - pos);
			unsigned group = 0;

			for(size_t k = 0; k < 3; k++)
				group = (group << 8) | ( (k < numIn) ? (unsigned char)raw[pos + k] : 0);

			for(size_t k = 0; k < 4; k++)
				out += (k <= numIn) ? alphabet[(group >> (18 - 6 * k) ) & 63] : '=';
		}

		return out;
	}

	inline std::string decode(const std::string& encoded)
	{
		std::string out;
		unsigned bits = 0;
		int numBits = 0;

		for(const unsigned char c : encoded)
		{
			unsigned value;

			if( (c >= 'A') && (c <= 'Z') ) value = c - 'A';
			else if( (c >= 'a') && (c <= 'z') ) value = 26 + (c - 'a');
			else if( (c >= '0') && (c <= '9') ) value = 52 + (c - '0');
			else if(c == '+') value = 62;
			else if(c == '/') value = 63;
			else continue; // padding, whitespace

			bits = (bits << 6) | value;
			numBits += 6;

			if(numBits >= 8)
			{
				numBits -= 8;
				out += (char)( (bits >> numBits) & 0xFF);
			}
		}

		return out;
	}
}

/* what one worker got out of the manifest: its files and slices in processing order */
class WorkerTreeShare
{
	public:
		std::vector<TreeSlice> slices;
		uint64_t numBlocks{0};
		uint64_t numBytes{0};

		void clear()
		{
			slices.clear();
			numBlocks = numBytes = 0;
		}

		bool empty() const { return slices.empty(); }
		size_t size() const { return slices.size(); }

		/* --treerand (PathStore.cpp:217-241): seed 0 = self-seeded like the reference */
		void shuffle(uint64_t seed)
		{
			std::mt19937 generator(seed ? (unsigned)seed : std::random_device()() );
			std::shuffle(slices.begin(), slices.end(), generator);
		}
};

class TreeManifest
{
	public:
		bool isLoaded{false};

		const std::vector<TreeSlice>& getDirs() const { return dirs; }
		size_t getNumDirs() const { return dirs.size(); }
		size_t getNumFiles() const { return smallFiles.size() + bigFiles.size(); }
		uint64_t getNumFileBytes() const { return smallBytes + bigBytes; }

		/**
		 * Parse the tree file (one pass). Files whose (rounded up) size is below fileShareSize are
		 * "small" (whole files per worker), the others "big" (block ranges shared between workers).
		 */
		void load(const std::string& treeFilePath, uint64_t newBlockSize, uint64_t fileShareSize,
			uint64_t roundUpSize)
		{
			*this = TreeManifest();
			blockSize = newBlockSize;

			std::ifstream stream(treeFilePath.c_str() );

			if(!stream)
				throw WorkerError("Opening input file failed: " + treeFilePath);

			bool pathsAreBase64 = false;
			bool sawEntryOrText = false; // the encoding header only counts ahead of the first entry
			std::string line;

			for(unsigned lineNum = 0; std::getline(stream, line); lineNum++)
			{
				if(!sawEntryOrText)
				{
					if(line == "# encoding=base64")
						pathsAreBase64 = true;

					if(!line.empty() && (line[0] != '#') )
						sawEntryOrText = true;
				}

				const char* cursor = skipBlanks(line.c_str() );
				const char* tokenEnd = skipToken(cursor);
				const size_t tokenLen = tokenEnd - cursor;

				if( (tokenLen != 1) || ( (*cursor != 'd') && (*cursor != 'f') ) )
					continue;

				const bool isDirLine = (*cursor == 'd');
				TreeSlice entry;
				cursor = tokenEnd;

				if(!isDirLine)
				{ // size field
					cursor = skipBlanks(cursor);

					char* numberEnd = NULL;
					const bool startsWithDigit = (*cursor >= '0') && (*cursor <= '9');
					const unsigned long long size =
						startsWithDigit ? strtoull(cursor, &numberEnd, 10) : 0;

					if(!startsWithDigit)
						throw WorkerError("Encountered invalid file line without size in input "
							"file. File: " + treeFilePath + "; "
							"Line number: " + std::to_string(lineNum) );

					entry.totalLen = size;
					cursor = numberEnd;

					if(roundUpSize && (entry.totalLen % roundUpSize) )
						entry.totalLen += roundUpSize - (entry.totalLen % roundUpSize);

					entry.rangeLen = entry.totalLen;
				}

				entry.path = trimmed(cursor);

				if(pathsAreBase64)
					entry.path = base64::decode(entry.path);

				if(entry.path.empty() )
					throw WorkerError(std::string("Encountered invalid ") +
						(isDirLine ? "directory" : "file") + " line without path in input "
						"file. File: " + treeFilePath + "; "
						"Line number: " + std::to_string(lineNum) );

				if(isDirLine)
					dirs.push_back(std::move(entry) );
				else
				if(entry.totalLen < fileShareSize)
				{
					smallBytes += entry.totalLen;
					smallFiles.push_back(std::move(entry) );
				}
				else
				{
					bigBytes += entry.totalLen;
					bigFiles.push_back(std::move(entry) );
				}
			}

			/* parents ahead of their subdirs, same order on every host (PathStore.cpp:192-197) */
			std::sort(dirs.begin(), dirs.end(), [](const TreeSlice& a, const TreeSlice& b)
				{ return std::make_tuple(a.path.size(), std::cref(a.path) ) <
					std::make_tuple(b.path.size(), std::cref(b.path) ); } );

			/* small files by size, so that "every n-th file" is balanced (PathStore.cpp:206-211) */
			std::sort(smallFiles.begin(), smallFiles.end(), [](const TreeSlice& a, const TreeSlice& b)
				{ return std::make_tuple(a.totalLen, std::cref(a.path) ) <
					std::make_tuple(b.totalLen, std::cref(b.path) ); } );

			/* big files stay in file order; blocksAhead[i] = blocks of all big files before file i */
			blocksAhead.resize(bigFiles.size() + 1);
			blocksAhead[0] = 0;

			for(size_t i = 0; i < bigFiles.size(); i++)
				blocksAhead[i + 1] = blocksAhead[i] + blocksOf(bigFiles[i].totalLen);

			isLoaded = true;
		}

		/* dirs rank, rank + n, ... of the sorted list */
		void takeDirs(uint64_t workerRank, uint64_t numDataSetThreads,
			std::vector<TreeSlice>& outDirs) const
		{
			for(size_t i = workerRank; i < dirs.size(); i += numDataSetThreads)
				outDirs.push_back(dirs[i] );
		}

		/**
		 * The worker's files: its stride of the small files, then its block range of the big ones.
		 *
		 * @strict throw if a file or slice is smaller than one block (direct random I/O).
		 */
		void takeFiles(uint64_t workerRank, uint64_t numDataSetThreads, bool strict,
			WorkerTreeShare& outShare) const
		{
			takeSmallFiles(workerRank, numDataSetThreads, strict, outShare);
			takeBigFileRanges(workerRank, numDataSetThreads, strict, outShare);
		}

		void takeSmallFiles(uint64_t workerRank, uint64_t numDataSetThreads, bool strict,
			WorkerTreeShare& outShare) const
		{
			for(size_t i = workerRank; i < smallFiles.size(); i += numDataSetThreads)
			{
				const TreeSlice& file = smallFiles[i];

				if(strict && (file.totalLen < blockSize) )
					throw WorkerError("Found file that is smaller than block size. Consider using "
						"\"--treeroundup\". (\"--nodiocheck\" disables this check.) "
						"File: " + file.path + "; "
						"FileSize: " + std::to_string(file.totalLen) + "; "
						"BlockSize: " + std::to_string(blockSize) );

				outShare.slices.push_back(file);
				outShare.numBlocks += blocksOf(file.totalLen);
				outShare.numBytes += file.totalLen;
			}
		}

		/**
		 * All blocks of all big files form one sequence. Worker r owns blocks
		 * [r * standard, r * standard + standard) of it, the last worker also the remainder
		 * (PathStore.cpp:334-352). A file that crosses a boundary is cut at the block boundary; a
		 * range that reaches the end of a file takes the (possibly partial) last block with it.
		 * Empty files take part only when they lie strictly inside a worker's range (the
		 * reference's walk behaves like that, :365-380).
		 */
		void takeBigFileRanges(uint64_t workerRank, uint64_t numDataSetThreads, bool strict,
			WorkerTreeShare& outShare) const
		{
			if(bigFiles.empty() )
				return;

			const uint64_t numBlocksAll = blocksAhead.back();
			const uint64_t standardShare = numBlocksAll / numDataSetThreads;
			const bool isLastWorker = (workerRank + 1 == numDataSetThreads);
			const uint64_t firstBlock = workerRank * standardShare;
			const uint64_t endBlock = isLastWorker ? numBlocksAll : (firstBlock + standardShare);

			outShare.numBlocks += (endBlock - firstBlock);

			if(endBlock == firstBlock)
				return;

			// first file that ends behind firstBlock: blocksAhead[i + 1] > firstBlock
			size_t fileIdx = std::upper_bound(blocksAhead.begin() + 1, blocksAhead.end(),
				firstBlock) - (blocksAhead.begin() + 1);

			for( ; (fileIdx < bigFiles.size() ) && (blocksAhead[fileIdx] < endBlock); fileIdx++)
			{
				const TreeSlice& file = bigFiles[fileIdx];
				const uint64_t fileFirstBlock = blocksAhead[fileIdx];
				const uint64_t fileEndBlock = blocksAhead[fileIdx + 1];

				const uint64_t cutFirst = std::max(firstBlock, fileFirstBlock) - fileFirstBlock;
				const bool reachesFileEnd = (fileEndBlock <= endBlock);

				TreeSlice slice = file;
				slice.rangeStart = cutFirst * blockSize;
				slice.rangeLen = reachesFileEnd ? (file.totalLen - slice.rangeStart) :
					( (endBlock - fileFirstBlock - cutFirst) * blockSize);

				if(strict && (slice.rangeLen < blockSize) )
					throw WorkerError("Found file slice that is smaller than block size. Consider "
						"using \"--treeroundup\". (\"--nodiocheck\" disables this check.) "
						"File: " + file.path + "; "
						"RangeStart: " + std::to_string(slice.rangeStart) + "; "
						"RangeLength: " + std::to_string(slice.rangeLen) + "; "
						"BlockSize: " + std::to_string(blockSize) );

				outShare.numBytes += slice.rangeLen;
				outShare.slices.push_back(std::move(slice) );
			}
		}

		/**
		 * --treescan (FileTk::scanCustomTree, toolkits/FileTk.cpp:387-470): walk scanPath and write
		 * a tree file with base64 encoded relative paths, directory by directory with sorted
		 * entries. @return number of entries
		 */
		static uint64_t scanToTreeFile(const std::string& scanPath,
			const std::string& outTreeFilePath, uint64_t& outNumDirs, uint64_t& outNumFiles,
			uint64_t& outNumBytes)
		{
			std::ofstream out(outTreeFilePath, std::ofstream::out | std::ofstream::trunc);

			if(!out)
				throw WorkerError("Opening tree scan results file failed: " + outTreeFilePath);

			out << "# encoding=base64" << std::endl;
			outNumDirs = outNumFiles = outNumBytes = 0;

			std::vector<std::string> todo(1); // relative dir paths still to be listed ("" = root)

			while(!todo.empty() )
			{
				const std::string relDir = std::move(todo.back() );
				todo.pop_back();

				const std::string absDir = relDir.empty() ? scanPath : (scanPath + "/" + relDir);
				std::vector<std::string> names;

				if(DIR* dir = opendir(absDir.c_str() ) )
				{
					while(struct dirent* entry = readdir(dir) )
						if(strcmp(entry->d_name, ".") && strcmp(entry->d_name, "..") )
							names.emplace_back(entry->d_name);

					closedir(dir);
				}
				else
					throw WorkerError("Unable to scan directory: " + absDir + "; "
						"SysErr: " + strerror(errno) );

				std::sort(names.begin(), names.end() );

				for(const std::string& name : names)
				{
					const std::string relPath = relDir.empty() ? name : (relDir + "/" + name);
					struct stat info;

					if(lstat( (scanPath + "/" + relPath).c_str(), &info) == -1)
						continue;

					if(S_ISDIR(info.st_mode) )
					{
						out << "d " << base64::encode(relPath) << std::endl;
						outNumDirs++;
						todo.push_back(relPath);
					}
					else
					if(S_ISREG(info.st_mode) )
					{
						out << "f " << info.st_size << " " << base64::encode(relPath) << std::endl;
						outNumFiles++;
						outNumBytes += info.st_size;
					}
				}
			}

			return outNumDirs + outNumFiles;
		}

	private:
		uint64_t blockSize{0};
		std::vector<TreeSlice> dirs;       // sorted: short paths first
		std::vector<TreeSlice> smallFiles; // sorted by size
		std::vector<TreeSlice> bigFiles;   // file order
		std::vector<uint64_t> blocksAhead; // bigFiles.size() + 1 prefix sums of block counts
		uint64_t smallBytes{0};
		uint64_t bigBytes{0};

		uint64_t blocksOf(uint64_t fileSize) const
		{
			return blockSize ? ( (fileSize / blockSize) + ( (fileSize % blockSize) ? 1 : 0) ) : 0;
		}

		static bool isBlank(char c) { return (c == ' ') || (c == '\t') || (c == '\r') || (c == '\n'); }

		static const char* skipBlanks(const char* text)
		{
			while(*text && isBlank(*text) )
				text++;
			return text;
		}

		static const char* skipToken(const char* text)
		{
			while(*text && !isBlank(*text) )
				text++;
			return text;
		}

		static std::string trimmed(const char* text)
		{
			text = skipBlanks(text);
			size_t len = strlen(text);

			while(len && isBlank(text[len - 1] ) )
				len--;

			return std::string(text, len);
		}
};

} // namespace elb

#endif /* ELB_TREE_H_ */
