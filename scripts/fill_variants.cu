// Exploration: write-only HBM bandwidth of store variants for K1 (pattern fill), aligned fast path.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fill_variants scripts/fill_variants.cu
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <stdint.h>
#include <stdio.h>

#define CHECK(x) do { cudaError_t e = (x); if(e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while(0)

struct __align__(32) u64x4 { uint64_t a, b, c, d; };

__device__ __forceinline__ void st256_na(void* p, uint64_t w)
{ asm volatile("st.global.L1::no_allocate.v4.u64 [%0], {%1,%2,%3,%4};" :: "l"(p), "l"(w), "l"(w+8), "l"(w+16), "l"(w+24) : "memory"); }
__device__ __forceinline__ void st256_cs(void* p, uint64_t w)
{ asm volatile("st.global.cs.v4.u64 [%0], {%1,%2,%3,%4};" :: "l"(p), "l"(w), "l"(w+8), "l"(w+16), "l"(w+24) : "memory"); }
__device__ __forceinline__ void st256_plain(void* p, uint64_t w)
{ asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" :: "l"(p), "l"(w), "l"(w+8), "l"(w+16), "l"(w+24) : "memory"); }
__device__ __forceinline__ void st128_na(void* p, uint64_t w)
{ asm volatile("st.global.L1::no_allocate.v2.u64 [%0], {%1,%2};" :: "l"(p), "l"(w), "l"(w+8) : "memory"); }

// MODE 0: 256b na, 1: 256b cs, 2: 256b plain, 3: 128b na
template<int MODE, int UNROLL, int MINB>
__global__ void __launch_bounds__(256, MINB) fillk(uint8_t* base, uint64_t bytes, uint64_t salt)
{
	constexpr int VEC = (MODE == 3) ? 16 : 32;
	const uint64_t tileBytes = 256ull * VEC * UNROLL;
	const uint64_t numTiles = bytes / tileBytes;
	const uint64_t chunk = (numTiles + gridDim.x - 1) / gridDim.x;
	uint64_t t0 = blockIdx.x * chunk, t1 = min(numTiles, t0 + chunk);
	for(uint64_t t = t0; t < t1; t++)
	{
		#pragma unroll
		for(int u = 0; u < UNROLL; u++)
		{
			uint64_t off = t * tileBytes + (uint64_t)(u * 256 + threadIdx.x) * VEC;
			if(MODE == 0) st256_na(base + off, off + salt);
			else if(MODE == 1) st256_cs(base + off, off + salt);
			else if(MODE == 2) st256_plain(base + off, off + salt);
			else st128_na(base + off, off + salt);
		}
	}
}

// grid-stride (tile round-robin) instead of contiguous chunks
template<int UNROLL, int MINB>
__global__ void __launch_bounds__(256, MINB) fillk_rr(uint8_t* base, uint64_t bytes, uint64_t salt)
{
	const uint64_t tileBytes = 256ull * 32 * UNROLL;
	const uint64_t numTiles = bytes / tileBytes;
	for(uint64_t t = blockIdx.x; t < numTiles; t += gridDim.x)
	{
		#pragma unroll
		for(int u = 0; u < UNROLL; u++)
		{
			uint64_t off = t * tileBytes + (uint64_t)(u * 256 + threadIdx.x) * 32;
			st256_na(base + off, off + salt);
		}
	}
}

// TMA bulk store: fill a smem tile, one thread issues cp.async.bulk shared->global; 2 stages
template<int TILE_KB>
__global__ void __launch_bounds__(256) fillk_tma(uint8_t* base, uint64_t bytes, uint64_t salt)
{
	extern __shared__ __align__(128) uint8_t smem[];
	constexpr uint64_t tileBytes = TILE_KB * 1024ull;
	constexpr int STAGES = 2;
	const uint64_t numTiles = bytes / tileBytes;
	const uint64_t chunk = (numTiles + gridDim.x - 1) / gridDim.x;
	uint64_t t0 = blockIdx.x * chunk, t1 = min(numTiles, t0 + chunk);
	int stage = 0;
	for(uint64_t t = t0; t < t1; t++)
	{
		uint8_t* buf = smem + stage * tileBytes;
		// make sure the bulk store that last read this stage is done (keep <= STAGES-1 pending)
		if(threadIdx.x == 0)
			asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(STAGES - 1) : "memory");
		__syncthreads();
		for(uint64_t i = threadIdx.x * 32ull; i < tileBytes; i += 256 * 32)
		{
			uint64_t w = t * tileBytes + i + salt;
			u64x4 v{w, w + 8, w + 16, w + 24};
			*reinterpret_cast<u64x4*>(buf + i) = v;
		}
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		__syncthreads();
		if(threadIdx.x == 0)
		{
			uint32_t saddr = (uint32_t)__cvta_generic_to_shared(buf);
			asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
				:: "l"(base + t * tileBytes), "r"(saddr), "r"((uint32_t)tileBytes) : "memory");
			asm volatile("cp.async.bulk.commit_group;" ::: "memory");
		}
		stage = (stage + 1) % STAGES;
	}
	if(threadIdx.x == 0)
		asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template<typename F> static void timeit(const char* name, F launch, uint64_t bytes)
{
	cudaEvent_t a, b; CHECK(cudaEventCreate(&a)); CHECK(cudaEventCreate(&b));
	for(int i = 0; i < 3; i++) launch();
	CHECK(cudaDeviceSynchronize());
	float best = 1e9, sum = 0; const int N = 20;
	for(int i = 0; i < N; i++)
	{
		CHECK(cudaEventRecord(a)); launch(); CHECK(cudaEventRecord(b)); CHECK(cudaEventSynchronize(b));
		float ms; CHECK(cudaEventElapsedTime(&ms, a, b)); best = fminf(best, ms); sum += ms;
	}
	CHECK(cudaGetLastError());
	printf("%-38s avg %8.1f GB/s  best %8.1f GB/s\n", name, bytes / (sum / N * 1e-3) / 1e9, bytes / (best * 1e-3) / 1e9);
}

int main()
{
	const uint64_t bytes = 4ull << 30;
	uint8_t* buf; CHECK(cudaMalloc(&buf, bytes));
	int sms; CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
	printf("SMs %d\n", sms);
	timeit("cudaMemset", [&]{ cudaMemsetAsync(buf, 0x5a, bytes); }, bytes);
	timeit("256b na  U4 4cta/SM chunk", [&]{ fillk<0,4,4><<<sms*4,256>>>(buf, bytes, 1); }, bytes);
	timeit("256b na  U4 8cta/SM chunk", [&]{ fillk<0,4,8><<<sms*8,256>>>(buf, bytes, 1); }, bytes);
	timeit("256b na  U8 4cta/SM chunk", [&]{ fillk<0,8,4><<<sms*4,256>>>(buf, bytes, 1); }, bytes);
	timeit("256b na  U2 8cta/SM chunk", [&]{ fillk<0,2,8><<<sms*8,256>>>(buf, bytes, 1); }, bytes);
	timeit("256b na  U1 8cta/SM chunk", [&]{ fillk<0,1,8><<<sms*8,256>>>(buf, bytes, 1); }, bytes);
	timeit("256b cs  U4 4cta/SM chunk", [&]{ fillk<1,4,4><<<sms*4,256>>>(buf, bytes, 1); }, bytes);
	timeit("256b cs  U4 8cta/SM chunk", [&]{ fillk<1,4,8><<<sms*8,256>>>(buf, bytes, 1); }, bytes);
	timeit("256b pl  U4 4cta/SM chunk", [&]{ fillk<2,4,4><<<sms*4,256>>>(buf, bytes, 1); }, bytes);
	timeit("128b na  U4 4cta/SM chunk", [&]{ fillk<3,4,4><<<sms*4,256>>>(buf, bytes, 1); }, bytes);
	timeit("128b na  U8 8cta/SM chunk", [&]{ fillk<3,8,8><<<sms*8,256>>>(buf, bytes, 1); }, bytes);
	timeit("256b na  U4 4cta/SM round-robin", [&]{ fillk_rr<4,4><<<sms*4,256>>>(buf, bytes, 1); }, bytes);
	timeit("256b na  U4 8cta/SM round-robin", [&]{ fillk_rr<4,8><<<sms*8,256>>>(buf, bytes, 1); }, bytes);
	timeit("256b na  U4 2cta/SM chunk", [&]{ fillk<0,4,2><<<sms*2,256>>>(buf, bytes, 1); }, bytes);
	timeit("256b na  U4 1cta/SM chunk", [&]{ fillk<0,4,1><<<sms*1,256>>>(buf, bytes, 1); }, bytes);
	CHECK(cudaFuncSetAttribute(fillk_tma<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
	timeit("TMA bulk store 32K x2 stages 1cta/SM", [&]{ fillk_tma<32><<<sms,256,64*1024>>>(buf, bytes, 1); }, bytes);
	timeit("TMA bulk store 32K x2 stages 2cta/SM", [&]{ fillk_tma<32><<<sms*2,256,64*1024>>>(buf, bytes, 1); }, bytes);
	timeit("TMA bulk store 32K x2 stages 3cta/SM", [&]{ fillk_tma<32><<<sms*3,256,64*1024>>>(buf, bytes, 1); }, bytes);
	CHECK(cudaFuncSetAttribute(fillk_tma<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024));
	timeit("TMA bulk store 16K x2 stages 4cta/SM", [&]{ fillk_tma<16><<<sms*4,256,32*1024>>>(buf, bytes, 1); }, bytes);
	timeit("TMA bulk store 16K x2 stages 6cta/SM", [&]{ fillk_tma<16><<<sms*6,256,32*1024>>>(buf, bytes, 1); }, bytes);
	// read-only reference for context: none here
	return 0;
}
