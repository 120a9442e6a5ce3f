/*
 * Exploration of the host side of the staged path (round 2): where does the time of
 * pread -> pinned ring -> H2D  and  D2H -> pinned ring -> pwrite  go on the GPU box?
 *
 *   pcie   : pinned H2D / D2H bandwidth from a buffer on NUMA node 0 / 1, by chunk size and streams
 *   tmpfs  : raw pread scaling of one tmpfs file by reader count, reader node, page node, buffer size
 *   wfiles : raw pwrite scaling across files (one writer per file)
 *   chase  : pread -> H2D per block with a small ring (cache resident) vs a big ring
 *   wpipe  : single-file write pipeline variants (gate kinds, dedicated writer, ring size)
 *
 * Not product code: a measurement tool, results go to profiles/.
 * Build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/explore_hostpath.bin \
 *        scripts/explore_hostpath.cu -lpthread
 */
#include <cuda_runtime.h>
#include <dirent.h>
#include <fcntl.h>
#include <pthread.h>
#include <sched.h>
#include <sys/stat.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

typedef std::chrono::steady_clock Clock;
static const uint64_t MiB = 1ULL << 20;
static const uint64_t GiB = 1ULL << 30;

#define CK(x) do { cudaError_t e_ = (x); if(e_ != cudaSuccess) { \
	fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while(0)

static double secsSince(Clock::time_point t)
{ return std::chrono::duration<double>(Clock::now() - t).count(); }

static std::vector<int> nodeCPUs(int node)
{
	std::vector<int> cpus;
	std::ifstream f("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist");
	std::string s;
	if(!f || !std::getline(f, s) )
		return cpus;
	std::stringstream ss(s);
	std::string el;
	while(std::getline(ss, el, ',') )
	{
		size_t dash = el.find('-');
		int a = atoi(el.c_str() );
		int b = (dash == std::string::npos) ? a : atoi(el.substr(dash + 1).c_str() );
		for(int c = a; c <= b; c++)
			cpus.push_back(c);
	}
	return cpus;
}

/* node >= 0: run on that node's cpus, memory from that node; -1: anywhere */
static void bindNode(int node)
{
	cpu_set_t set;
	CPU_ZERO(&set);
	if(node < 0)
	{
		for(int c = 0; c < CPU_SETSIZE; c++)
			CPU_SET(c, &set);
		sched_setaffinity(0, sizeof(set), &set);
		syscall(SYS_set_mempolicy, 0, NULL, 0);
		return;
	}
	for(int c : nodeCPUs(node) )
		CPU_SET(c, &set);
	sched_setaffinity(0, sizeof(set), &set);
	unsigned long mask[16] = {};
	mask[0] = 1UL << node;
	syscall(SYS_set_mempolicy, 2 /*MPOL_BIND*/, mask, sizeof(mask) * 8);
}

static std::string g_dir = "/dev/shm";

/* ---------------------------------------------------------------------------------------------- */

static void runPcie()
{
	for(int node = 0; node <= 1; node++)
	{
		if(nodeCPUs(node).empty() )
			continue;
		bindNode(node);
		const uint64_t bufLen = 512 * MiB;
		char* host;
		char* dev;
		CK(cudaHostAlloc( (void**)&host, bufLen, cudaHostAllocDefault) );
		memset(host, 1, bufLen);
		CK(cudaMalloc( (void**)&dev, bufLen) );
		for(int dir = 0; dir < 2; dir++)
			for(uint64_t chunk : {1 * MiB, 16 * MiB})
				for(int nstreams : {1, 4})
				{
					std::vector<cudaStream_t> streams(nstreams);
					for(auto& s : streams)
						CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) );
					const uint64_t total = 8 * GiB;
					CK(cudaDeviceSynchronize() );
					Clock::time_point t0 = Clock::now();
					uint64_t off = 0;
					for(uint64_t done = 0, i = 0; done < total; done += chunk, i++)
					{
						if(dir == 0)
							CK(cudaMemcpyAsync(dev + off, host + off, chunk, cudaMemcpyHostToDevice,
								streams[i % nstreams]) );
						else
							CK(cudaMemcpyAsync(host + off, dev + off, chunk, cudaMemcpyDeviceToHost,
								streams[i % nstreams]) );
						off = (off + chunk) % bufLen;
					}
					CK(cudaDeviceSynchronize() );
					double secs = secsSince(t0);
					printf("{\"test\":\"pcie\",\"host_node\":%d,\"dir\":\"%s\",\"chunk_mib\":%llu,"
						"\"streams\":%d,\"gib_s\":%.2f}\n", node, dir ? "d2h" : "h2d",
						(unsigned long long)(chunk / MiB), nstreams, total / (double)GiB / secs);
					fflush(stdout);
					for(auto& s : streams)
						cudaStreamDestroy(s);
				}
		cudaFreeHost(host);
		cudaFree(dev);
	}
	bindNode(-1);
}

/* ---------------------------------------------------------------------------------------------- */

struct Barrier
{
	std::mutex m;
	std::condition_variable cv;
	int count, waiting = 0, gen = 0;
	explicit Barrier(int n) : count(n) {}
	void wait()
	{
		std::unique_lock<std::mutex> l(m);
		int g = gen;
		if(++waiting == count)
		{
			gen++;
			waiting = 0;
			cv.notify_all();
		}
		else
			cv.wait(l, [&] { return g != gen; });
	}
};

/* every thread runs fn(threadIdx) between two barriers; returns seconds of the slowest */
template <typename F>
static double runThreads(int n, int node, F fn)
{
	Barrier startBar(n + 1);
	std::vector<std::thread> threads;
	std::atomic<int> done{0};
	for(int i = 0; i < n; i++)
		threads.emplace_back([&, i]()
		{
			bindNode(node);
			startBar.wait();
			fn(i);
			done++;
		});
	startBar.wait();
	Clock::time_point t0 = Clock::now();
	for(auto& t : threads)
		t.join();
	return secsSince(t0);
}

static double writeFiles(const std::vector<std::string>& paths, uint64_t fileSize, int writersPerFile,
	int node)
{
	std::vector<int> fds;
	for(const std::string& p : paths)
	{
		unlink(p.c_str() );
		fds.push_back(open(p.c_str(), O_CREAT | O_RDWR, 0600) );
	}
	const int n = (int)paths.size() * writersPerFile;
	double secs = runThreads(n, node, [&](int idx)
	{
		const int fd = fds[idx / writersPerFile];
		const int w = idx % writersPerFile;
		const uint64_t share = fileSize / writersPerFile;
		char* buf = (char*)aligned_alloc(4096, MiB);
		memset(buf, idx + 1, MiB);
		for(uint64_t off = w * share; off < (w + 1) * share; off += MiB)
			if(pwrite(fd, buf, MiB, off) != (ssize_t)MiB)
			{
				perror("pwrite");
				exit(1);
			}
		free(buf);
	});
	for(int fd : fds)
		close(fd);
	return secs;
}

static double readFile(const std::string& path, uint64_t fileSize, int readers, int node,
	uint64_t bufLen)
{
	const int fd = open(path.c_str(), O_RDONLY);
	double secs = runThreads(readers, node, [&](int idx)
	{
		const uint64_t share = fileSize / readers;
		char* buf = (char*)aligned_alloc(4096, bufLen);
		memset(buf, 0, bufLen);
		uint64_t bufOff = 0;
		for(uint64_t off = idx * share; off < (idx + 1) * share; off += MiB)
		{
			if(pread(fd, buf + bufOff, MiB, off) != (ssize_t)MiB)
			{
				perror("pread");
				exit(1);
			}
			bufOff = (bufOff + MiB) % bufLen;
		}
		free(buf);
	});
	close(fd);
	return secs;
}

/* kernel stack sampling of all threads of this process (best effort; needs root) */
static void sampleStacks(std::atomic<bool>& stop, std::map<std::string, int>& histo)
{
	while(!stop)
	{
		DIR* d = opendir("/proc/self/task");
		if(!d)
			return;
		struct dirent* e;
		while( (e = readdir(d) ) )
		{
			if(e->d_name[0] == '.')
				continue;
			std::ifstream f(std::string("/proc/self/task/") + e->d_name + "/stack");
			std::string line, key;
			int n = 0;
			while(std::getline(f, line) && (n < 5) )
			{
				size_t p = line.find("] ");
				std::string fn = (p == std::string::npos) ? line : line.substr(p + 2);
				size_t plus = fn.find('+');
				if(plus != std::string::npos)
					fn = fn.substr(0, plus);
				key += fn + "<";
				n++;
			}
			if(!key.empty() )
				histo[key]++;
		}
		closedir(d);
		usleep(2000);
	}
}

static void runTmpfs(uint64_t fileSize, bool withStacks)
{
	const std::string path = g_dir + "/xp_tmpfs.bin";
	for(int pageNode : {0, 1})
	{
		if(nodeCPUs(pageNode).empty() )
			continue;
		double wsecs = writeFiles({path}, fileSize, 1, pageNode);
		printf("{\"test\":\"tmpfs_write\",\"page_node\":%d,\"writers\":1,\"gib_s\":%.2f}\n", pageNode,
			fileSize / (double)GiB / wsecs);
		fflush(stdout);
		for(int readers : {8, 16, 32, 64})
			for(int node : {0, 1, -1})
				for(uint64_t bufLen : {1 * MiB, 32 * MiB})
				{
					if( (pageNode == 1) && ( (readers != 16) || (bufLen != MiB) ) )
						continue;
					double secs = readFile(path, fileSize, readers, node, bufLen);
					printf("{\"test\":\"tmpfs_read\",\"page_node\":%d,\"readers\":%d,\"reader_node\":%d,"
						"\"buf_mib\":%llu,\"gib_s\":%.2f}\n", pageNode, readers, node,
						(unsigned long long)(bufLen / MiB), fileSize / (double)GiB / secs);
					fflush(stdout);
				}
		if(withStacks && (pageNode == 0) )
		{
			std::atomic<bool> stop{false};
			std::map<std::string, int> histo;
			std::thread sampler(sampleStacks, std::ref(stop), std::ref(histo) );
			for(int rep = 0; rep < 3; rep++)
				readFile(path, fileSize, 64, -1, MiB);
			stop = true;
			sampler.join();
			std::vector<std::pair<int, std::string> > sorted;
			for(auto& kv : histo)
				sorted.push_back({kv.second, kv.first});
			std::sort(sorted.rbegin(), sorted.rend() );
			for(size_t i = 0; i < std::min(sorted.size(), (size_t)12); i++)
				printf("{\"test\":\"stacks_read64\",\"n\":%d,\"stack\":\"%s\"}\n", sorted[i].first,
					sorted[i].second.c_str() );
			fflush(stdout);
		}
	}
	// several writers on one file: inode lock behaviour
	for(int writers : {1, 2, 4, 16})
	{
		double wsecs = writeFiles({path}, fileSize / 2, writers, 0);
		printf("{\"test\":\"tmpfs_write\",\"page_node\":0,\"writers\":%d,\"gib_s\":%.2f}\n", writers,
			(fileSize / 2) / (double)GiB / wsecs);
		fflush(stdout);
	}
	// rewrite of existing pages (no allocation)
	{
		const int fd = open(path.c_str(), O_RDWR);
		char* buf = (char*)aligned_alloc(4096, MiB);
		memset(buf, 7, MiB);
		Clock::time_point t0 = Clock::now();
		for(uint64_t off = 0; off < fileSize / 2; off += MiB)
			if(pwrite(fd, buf, MiB, off) != (ssize_t)MiB)
				exit(1);
		double secs = secsSince(t0);
		printf("{\"test\":\"tmpfs_rewrite\",\"writers\":1,\"gib_s\":%.2f}\n",
			(fileSize / 2) / (double)GiB / secs);
		close(fd);
		free(buf);
	}
	unlink(path.c_str() );
}

static void runWFiles(uint64_t fileSize)
{
	for(int nfiles : {1, 2, 4, 8, 16})
		for(int node : {-1, 0})
		{
			std::vector<std::string> paths;
			for(int i = 0; i < nfiles; i++)
				paths.push_back(g_dir + "/xp_wfiles_" + std::to_string(i) + ".bin");
			double secs = writeFiles(paths, fileSize, 1, node);
			printf("{\"test\":\"wfiles\",\"files\":%d,\"node\":%d,\"gib_s\":%.2f}\n", nfiles, node,
				nfiles * fileSize / (double)GiB / secs);
			fflush(stdout);
			for(auto& p : paths)
				unlink(p.c_str() );
		}
}

/* ---------------------------------------------------------------------------------------------- */

/* pread -> H2D per block through a ring of numSlots x 1 MiB per reader */
static void runChase(uint64_t fileSize)
{
	const std::string path = g_dir + "/xp_chase.bin";
	writeFiles({path}, fileSize, 4, 0);
	for(int readers : {8, 16, 32})
		for(int node : {0, 1, -1})
			for(int numSlots : {2, 4, 32})
			{
				const int fd = open(path.c_str(), O_RDONLY);
				double secs = runThreads(readers, node, [&](int idx)
				{
					CK(cudaSetDevice(0) );
					char* host;
					char* dev;
					cudaStream_t stream;
					CK(cudaHostAlloc( (void**)&host, numSlots * MiB, cudaHostAllocDefault) );
					memset(host, 0, numSlots * MiB);
					CK(cudaMalloc( (void**)&dev, numSlots * MiB) );
					CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) );
					std::vector<cudaEvent_t> events(numSlots);
					for(auto& e : events)
						CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming) );
					const uint64_t share = fileSize / readers;
					uint64_t i = 0;
					for(uint64_t off = idx * share; off < (idx + 1) * share; off += MiB, i++)
					{
						const int slot = i % numSlots;
						if(i >= (uint64_t)numSlots)
							CK(cudaEventSynchronize(events[slot]) );
						if(pread(fd, host + slot * MiB, MiB, off) != (ssize_t)MiB)
							exit(1);
						CK(cudaMemcpyAsync(dev + slot * MiB, host + slot * MiB, MiB,
							cudaMemcpyHostToDevice, stream) );
						CK(cudaEventRecord(events[slot], stream) );
					}
					CK(cudaStreamSynchronize(stream) );
					for(auto& e : events)
						cudaEventDestroy(e);
					cudaStreamDestroy(stream);
					cudaFreeHost(host);
					cudaFree(dev);
				});
				close(fd);
				printf("{\"test\":\"chase_read\",\"readers\":%d,\"node\":%d,\"slots\":%d,\"gib_s\":%.2f,"
					"\"note\":\"includes per-thread alloc/free\"}\n", readers, node, numSlots,
					fileSize / (double)GiB / secs);
				fflush(stdout);
			}
	unlink(path.c_str() );
}

/* ---------------------------------------------------------------------------------------------- */

struct TicketGate
{
	std::atomic<uint64_t> next{0};
	std::atomic<uint64_t> serving{0};
	uint64_t take() { return next.fetch_add(1); }
	void wait(uint64_t ticket)
	{
		int spins = 0;
		while(serving.load(std::memory_order_acquire) != ticket)
		{
			if(++spins > 2000)
			{
				sched_yield();
				spins = 0;
			}
			else
				__builtin_ia32_pause();
		}
	}
	void release() { serving.fetch_add(1, std::memory_order_release); }
};

/* single-file write pipeline: workers produce blocks by D2H into their ring, then write them.
 * gate: 0 none, 1 mutex, 2 ticket spin (FIFO), 3 dedicated writer thread */
static void runWPipe(uint64_t fileSize)
{
	const std::string path = g_dir + "/xp_wpipe.bin";
	for(int workers : {1, 2, 4, 16})
		for(int gate : {0, 1, 2, 3})
			for(int numSlots : {2, 16})
			{
				if( (workers == 1) && (gate == 1 || gate == 2) )
					continue;
				unlink(path.c_str() );
				const int fd = open(path.c_str(), O_CREAT | O_RDWR, 0600);
				std::mutex gateMutex;
				TicketGate ticketGate;

				// dedicated writer: queue of (buf, offset, doneFlag)
				struct Item { char* buf; uint64_t off; std::atomic<int>* done; };
				std::mutex qMutex;
				std::condition_variable qCond;
				std::deque<Item> queue;
				std::atomic<bool> writerStop{false};
				std::thread writer;
				if(gate == 3)
					writer = std::thread([&]()
					{
						bindNode(0);
						for( ; ; )
						{
							Item item;
							{
								std::unique_lock<std::mutex> l(qMutex);
								qCond.wait(l, [&] { return !queue.empty() || writerStop; });
								if(queue.empty() )
									return;
								item = queue.front();
								queue.pop_front();
							}
							if(pwrite(fd, item.buf, MiB, item.off) != (ssize_t)MiB)
								exit(1);
							item.done->store(1, std::memory_order_release);
						}
					});

				double secs = runThreads(workers, 0, [&](int idx)
				{
					CK(cudaSetDevice(0) );
					char* host;
					char* dev;
					cudaStream_t stream;
					CK(cudaHostAlloc( (void**)&host, numSlots * MiB, cudaHostAllocDefault) );
					memset(host, 0, numSlots * MiB);
					CK(cudaMalloc( (void**)&dev, numSlots * MiB) );
					CK(cudaMemset(dev, idx + 1, numSlots * MiB) );
					CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) );
					std::vector<cudaEvent_t> events(numSlots);
					for(auto& e : events)
						CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming) );
					std::vector<std::atomic<int> > slotFree(numSlots);
					for(auto& s : slotFree)
						s = 1;
					const uint64_t share = fileSize / workers;
					const uint64_t numBlocks = share / MiB;
					const uint64_t ahead = numSlots / 2; // D2H copies issued ahead of the write
					uint64_t issued = 0;
					for(uint64_t b = 0; b < numBlocks; b++)
					{
						// keep `ahead` D2H copies in flight (the GPU stage of later blocks)
						while( (issued < numBlocks) && (issued < b + std::max<uint64_t>(1, ahead) ) )
						{
							const int s = issued % numSlots;
							while(!slotFree[s].load(std::memory_order_acquire) )
								__builtin_ia32_pause();
							CK(cudaMemcpyAsync(host + s * MiB, dev + s * MiB, MiB,
								cudaMemcpyDeviceToHost, stream) );
							CK(cudaEventRecord(events[s], stream) );
							issued++;
						}
						const int slot = b % numSlots;
						const uint64_t off = idx * share + b * MiB;
						CK(cudaEventSynchronize(events[slot]) );
						if(gate == 3)
						{
							slotFree[slot].store(0, std::memory_order_relaxed);
							{
								std::unique_lock<std::mutex> l(qMutex);
								queue.push_back(Item{host + slot * MiB, off, &slotFree[slot]});
							}
							qCond.notify_one();
							continue;
						}
						if(gate == 1)
							gateMutex.lock();
						uint64_t ticket = 0;
						if(gate == 2)
						{
							ticket = ticketGate.take();
							ticketGate.wait(ticket);
						}
						if(pwrite(fd, host + slot * MiB, MiB, off) != (ssize_t)MiB)
							exit(1);
						if(gate == 1)
							gateMutex.unlock();
						if(gate == 2)
							ticketGate.release();
					}
					for(auto& s : slotFree)
						while(!s.load(std::memory_order_acquire) )
							__builtin_ia32_pause();
					for(auto& e : events)
						cudaEventDestroy(e);
					cudaStreamDestroy(stream);
					cudaFreeHost(host);
					cudaFree(dev);
				});
				if(gate == 3)
				{
					{
						std::unique_lock<std::mutex> l(qMutex);
						writerStop = true;
					}
					qCond.notify_all();
					writer.join();
				}
				close(fd);
				static const char* gateNames[] = {"none", "mutex", "ticket_spin", "dedicated_writer"};
				printf("{\"test\":\"wpipe\",\"workers\":%d,\"gate\":\"%s\",\"slots\":%d,\"gib_s\":%.2f}\n",
					workers, gateNames[gate], numSlots, fileSize / (double)GiB / secs);
				fflush(stdout);
			}
	unlink(path.c_str() );
}

/* ---------------------------------------------------------------------------------------------- */
/* SM-driven staging: kernels that read / write pinned host memory directly over PCIe */

struct alignas(32) V32 { unsigned long long a, b, c, d; };

__device__ __forceinline__ V32 ldNc256(const void* p)
{
	V32 v;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
		: "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p) );
	return v;
}

__device__ __forceinline__ void st256(void* p, const V32& v)
{
	asm volatile("st.global.L1::no_allocate.v4.u64 [%0], {%1,%2,%3,%4};"
		:: "l"(p), "l"(v.a), "l"(v.b), "l"(v.c), "l"(v.d) : "memory");
}

/* one 32 KiB tile per CTA (256 threads x 4 x 32 B): src -> dst, xor-sum as a stand-in for verify */
__global__ void __launch_bounds__(256, 4) stageKernel(const char* src, char* dst,
	unsigned long long* sink)
{
	const size_t tile = (size_t)blockIdx.x * 32768;
	V32 v[4];
	#pragma unroll
	for(int u = 0; u < 4; u++)
		v[u] = ldNc256(src + tile + (size_t)(u * 256 + threadIdx.x) * 32);
	unsigned long long x = 0;
	#pragma unroll
	for(int u = 0; u < 4; u++)
	{
		st256(dst + tile + (size_t)(u * 256 + threadIdx.x) * 32, v[u]);
		x ^= v[u].a ^ v[u].b ^ v[u].c ^ v[u].d;
	}
	if(x == 0x1234567890abcdefULL)
		*sink = x;
}

static void runZcopy()
{
	for(int node = 0; node <= 1; node++)
	{
		if(nodeCPUs(node).empty() )
			continue;
		bindNode(node);
		const uint64_t bufLen = 256 * MiB;
		char* host;
		char* dev;
		unsigned long long* sink;
		CK(cudaHostAlloc( (void**)&host, bufLen, cudaHostAllocDefault) );
		memset(host, 1, bufLen);
		CK(cudaMalloc( (void**)&dev, bufLen) );
		CK(cudaMalloc( (void**)&sink, 8) );
		for(int dir = 0; dir < 2; dir++)
			for(uint64_t chunk : {(uint64_t)65536, 1 * MiB, 16 * MiB})
				for(int nstreams : {1, 4, 16})
				{
					std::vector<cudaStream_t> streams(nstreams);
					for(auto& s : streams)
						CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) );
					const uint64_t total = (chunk < MiB) ? (1 * GiB) : (4 * GiB);
					CK(cudaDeviceSynchronize() );
					Clock::time_point t0 = Clock::now();
					uint64_t off = 0;
					for(uint64_t done = 0, i = 0; done < total; done += chunk, i++)
					{
						if(dir == 0)
							stageKernel<<<(unsigned)(chunk / 32768), 256, 0, streams[i % nstreams]>>>(
								host + off, dev + off, sink);
						else
							stageKernel<<<(unsigned)(chunk / 32768), 256, 0, streams[i % nstreams]>>>(
								dev + off, host + off, sink);
						off = (off + chunk) % bufLen;
					}
					CK(cudaDeviceSynchronize() );
					double secs = secsSince(t0);
					printf("{\"test\":\"zcopy_kernel\",\"host_node\":%d,\"dir\":\"%s\",\"chunk_kib\":%llu,"
						"\"streams\":%d,\"gib_s\":%.2f,\"launches_per_s\":%.0f}\n", node,
						dir ? "d2h" : "h2d", (unsigned long long)(chunk / 1024), nstreams,
						total / (double)GiB / secs, (total / chunk) / secs);
					fflush(stdout);
					for(auto& s : streams)
						cudaStreamDestroy(s);
				}
		cudaFreeHost(host);
		cudaFree(dev);
		cudaFree(sink);
	}
	bindNode(-1);
}

/* pread -> fused stage kernel per block (no copy engine) through a small ring */
static void runChaseKernel(uint64_t fileSize)
{
	const std::string path = g_dir + "/xp_chasek.bin";
	writeFiles({path}, fileSize, 4, 0);
	for(int readers : {8, 16, 32})
		for(int node : {0, -1})
			for(int numSlots : {2, 4})
			{
				const int fd = open(path.c_str(), O_RDONLY);
				double secs = runThreads(readers, node, [&](int idx)
				{
					CK(cudaSetDevice(0) );
					char* host;
					char* dev;
					unsigned long long* sink;
					cudaStream_t stream;
					CK(cudaHostAlloc( (void**)&host, numSlots * MiB, cudaHostAllocDefault) );
					memset(host, 0, numSlots * MiB);
					CK(cudaMalloc( (void**)&dev, numSlots * MiB) );
					CK(cudaMalloc( (void**)&sink, 8) );
					CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) );
					std::vector<cudaEvent_t> events(numSlots);
					for(auto& e : events)
						CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming) );
					const uint64_t share = fileSize / readers;
					uint64_t i = 0;
					for(uint64_t off = idx * share; off < (idx + 1) * share; off += MiB, i++)
					{
						const int slot = i % numSlots;
						if(i >= (uint64_t)numSlots)
							CK(cudaEventSynchronize(events[slot]) );
						if(pread(fd, host + slot * MiB, MiB, off) != (ssize_t)MiB)
							exit(1);
						stageKernel<<<32, 256, 0, stream>>>(host + slot * MiB, dev + slot * MiB, sink);
						CK(cudaEventRecord(events[slot], stream) );
					}
					CK(cudaStreamSynchronize(stream) );
					for(auto& e : events)
						cudaEventDestroy(e);
					cudaStreamDestroy(stream);
					cudaFreeHost(host);
					cudaFree(dev);
					cudaFree(sink);
				});
				close(fd);
				printf("{\"test\":\"chase_read_kernel\",\"readers\":%d,\"node\":%d,\"slots\":%d,"
					"\"gib_s\":%.2f}\n", readers, node, numSlots, fileSize / (double)GiB / secs);
				fflush(stdout);
			}
	unlink(path.c_str() );
}

/* memory system ceiling of "CPU copy into a pinned slot, then the GPU reads the slot": no file
 * system involved. Per thread: stream through a 256 MiB source (DRAM resident, stands for the page
 * cache) with memcpy into a ring of numSlots x 1 MiB, then move the slot with the copy engine or
 * the stage kernel. Small ring = slot still in cache when the device reads it. */
static void runDmaSrc()
{
	for(int threads : {8, 16})
		for(int mode : {0, 1, 2}) // 0 memcpy only, 1 + copy engine, 2 + stage kernel
			for(int numSlots : {2, 4, 64})
			{
				const uint64_t perThread = 2 * GiB;
				double secs = runThreads(threads, 0, [&](int idx)
				{
					CK(cudaSetDevice(0) );
					const uint64_t srcLen = 256 * MiB;
					char* src = (char*)aligned_alloc(4096, srcLen);
					memset(src, idx + 1, srcLen);
					char* host;
					char* dev;
					unsigned long long* sink;
					cudaStream_t stream;
					CK(cudaHostAlloc( (void**)&host, numSlots * MiB, cudaHostAllocDefault) );
					memset(host, 0, numSlots * MiB);
					CK(cudaMalloc( (void**)&dev, numSlots * MiB) );
					CK(cudaMalloc( (void**)&sink, 8) );
					CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) );
					std::vector<cudaEvent_t> events(numSlots);
					for(auto& e : events)
						CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming) );
					uint64_t i = 0;
					for(uint64_t done = 0; done < perThread; done += MiB, i++)
					{
						const int slot = i % numSlots;
						if(mode && (i >= (uint64_t)numSlots) )
							CK(cudaEventSynchronize(events[slot]) );
						memcpy(host + slot * MiB, src + (done % srcLen), MiB);
						if(mode == 1)
							CK(cudaMemcpyAsync(dev + slot * MiB, host + slot * MiB, MiB,
								cudaMemcpyHostToDevice, stream) );
						if(mode == 2)
							stageKernel<<<32, 256, 0, stream>>>(host + slot * MiB, dev + slot * MiB,
								sink);
						if(mode)
							CK(cudaEventRecord(events[slot], stream) );
					}
					CK(cudaStreamSynchronize(stream) );
					for(auto& e : events)
						cudaEventDestroy(e);
					cudaStreamDestroy(stream);
					cudaFreeHost(host);
					cudaFree(dev);
					cudaFree(sink);
					free(src);
				});
				static const char* modeNames[] = {"memcpy_only", "copy_engine", "stage_kernel"};
				printf("{\"test\":\"dmasrc\",\"threads\":%d,\"mode\":\"%s\",\"slots\":%d,"
					"\"gib_s\":%.2f}\n", threads, modeNames[mode], numSlots,
					threads * perThread / (double)GiB / secs);
				fflush(stdout);
			}
}

/* is the FIRST read of freshly written tmpfs pages slower than later reads? 16 writers write their
 * own range of a fresh file, then three read passes with 16 readers (same ranges), then one with
 * shifted ranges; everything on node 0, 1 MiB private buffers */
static void runFirstRead(uint64_t fileSize)
{
	const std::string path = g_dir + "/xp_firstread.bin";
	for(int rep = 0; rep < 2; rep++)
	{
		double wsecs = writeFiles({path}, fileSize, 16, 0);
		printf("{\"test\":\"firstread_write\",\"rep\":%d,\"writers\":16,\"gib_s\":%.2f}\n", rep,
			fileSize / (double)GiB / wsecs);
		for(int pass = 0; pass < 3; pass++)
		{
			double secs = readFile(path, fileSize, 16, 0, MiB);
			printf("{\"test\":\"firstread_read\",\"rep\":%d,\"pass\":%d,\"readers\":16,\"gib_s\":%.2f}\n",
				rep, pass, fileSize / (double)GiB / secs);
			fflush(stdout);
		}
		double secs8 = readFile(path, fileSize, 8, 0, MiB);
		printf("{\"test\":\"firstread_read\",\"rep\":%d,\"pass\":3,\"readers\":8,\"gib_s\":%.2f}\n", rep,
			fileSize / (double)GiB / secs8);
		fflush(stdout);
		unlink(path.c_str() );
	}
}

int main(int argc, char** argv)
{
	std::string tests = (argc > 1) ? argv[1] : "pcie,tmpfs,wfiles,chase,wpipe";
	uint64_t gib = (argc > 2) ? strtoull(argv[2], NULL, 10) : 8;
	if(argc > 3)
		g_dir = argv[3];
	const bool withStacks = (tests.find("stacks") != std::string::npos);
	CK(cudaSetDevice(0) );
	CK(cudaFree(0) );
	char busId[64] = {};
	cudaDeviceGetPCIBusId(busId, sizeof(busId), 0);
	for(char* c = busId; *c; c++)
		*c = tolower(*c);
	std::ifstream numaFile(std::string("/sys/bus/pci/devices/") + busId + "/numa_node");
	std::string gpuNode = "?";
	if(numaFile)
		std::getline(numaFile, gpuNode);
	printf("{\"test\":\"topo\",\"gpu0_bus\":\"%s\",\"gpu0_numa_node\":\"%s\",\"node0_cpus\":%zu,"
		"\"node1_cpus\":%zu}\n", busId, gpuNode.c_str(), nodeCPUs(0).size(), nodeCPUs(1).size() );
	fflush(stdout);
	if(tests.find("pcie") != std::string::npos)
		runPcie();
	if(tests.find("tmpfs") != std::string::npos)
		runTmpfs(gib * GiB, withStacks);
	if(tests.find("wfiles") != std::string::npos)
		runWFiles(2 * GiB);
	if(tests.find("zcopy") != std::string::npos)
		runZcopy();
	if(tests.find("chase,") != std::string::npos || tests.rfind("chase") == tests.size() - 5)
		runChase(gib * GiB);
	if(tests.find("chasek") != std::string::npos)
		runChaseKernel(gib * GiB);
	if(tests.find("firstread") != std::string::npos)
		runFirstRead(gib * GiB);
	if(tests.find("dmasrc") != std::string::npos)
		runDmaSrc();
	if(tests.find("wpipe") != std::string::npos)
		runWPipe( (gib / 2) * GiB);
	return 0;
}
