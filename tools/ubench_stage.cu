// ubench_stage.cu -- what does it cost to bring the 132 KiB NCO correction table into every SM's
// shared memory at kernel start?  148 CTAs x 512 threads x 227 KB, back-to-back launches.
//   A empty kernel   B ld.global/st.shared loop   C bulk copy (TMA) unicast
//   D/E bulk copy multicast in clusters of 2 / 4 (each CTA fetches 1/csz of the table for all)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_stage tools/ubench_stage.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
constexpr unsigned kBytes = 135296;
constexpr unsigned kSmem = 228000;

__global__ void __launch_bounds__(512, 1) k_empty(const uint4 *, unsigned *sink) { extern __shared__ uint4 sm[]; if (threadIdx.x == 9999) sink[0] = sm[0].x; }

__global__ void __launch_bounds__(512, 1) k_ldst(const uint4 *g, unsigned *sink)
{
	extern __shared__ uint4 sm[];
	#pragma unroll 4
	for (unsigned i = threadIdx.x; i < kBytes / 16; i += 512) sm[i] = __ldg(g + i);
	__syncthreads();
	if (sm[threadIdx.x].x == 0x12345678u) sink[0] = 1;
}

__device__ __forceinline__ void wait0(uint32_t bar)
{
	uint32_t done;
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar) : "memory");
	} while (!done);
}

__global__ void __launch_bounds__(512, 1) k_tma(const uint4 *g, unsigned *sink)
{
	extern __shared__ uint4 sm[];
	__shared__ __align__(8) unsigned long long bar;
	const uint32_t bar32 = (uint32_t)__cvta_generic_to_shared(&bar), dst32 = (uint32_t)__cvta_generic_to_shared(sm);
	if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar32)); asm volatile("fence.mbarrier_init.release.cluster;"); }
	__syncthreads();
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar32), "r"(kBytes) : "memory");
		for (unsigned off = 0; off < kBytes; off += 16384) {
			unsigned nb = min(16384u, kBytes - off);
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
					:: "r"(dst32 + off), "l"(reinterpret_cast<const char*>(g) + off), "r"(nb), "r"(bar32) : "memory");
		}
	}
	wait0(bar32);
	if (sm[threadIdx.x].x == 0x12345678u) sink[0] = 1;
}

template <int CS>
__global__ void __launch_bounds__(512, 1) k_mc(const uint4 *g, unsigned *sink)
{
	extern __shared__ uint4 sm[];
	__shared__ __align__(8) unsigned long long bar;
	cg::cluster_group cl = cg::this_cluster();
	const unsigned rank = cl.block_rank();
	const uint32_t bar32 = (uint32_t)__cvta_generic_to_shared(&bar), dst32 = (uint32_t)__cvta_generic_to_shared(sm);
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar32));
		asm volatile("fence.mbarrier_init.release.cluster;");
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar32), "r"(kBytes) : "memory");
	}
	cl.sync();   // every CTA's barrier is armed before anybody multicasts into it
	if (threadIdx.x == 0) {
		constexpr unsigned per = ((kBytes / CS) + 15u) & ~15u;
		const unsigned lo = rank * per, hi = min(kBytes, lo + per);
		for (unsigned off = lo; off < hi; off += 16384) {
			unsigned nb = min(16384u, hi - off);
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
					:: "r"(dst32 + off), "l"(reinterpret_cast<const char*>(g) + off), "r"(nb), "r"(bar32), "h"((unsigned short)((1u << CS) - 1)) : "memory");
		}
	}
	wait0(bar32);
	if (sm[threadIdx.x].x == 0x12345678u) sink[0] = 1;
	cl.sync();   // nobody exits while a peer may still multicast into its shared memory
}

template <typename K>
int timeit(const char *name, K kernel, int cs, const uint4 *g, unsigned *sink)
{
	CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
	cudaLaunchConfig_t cfg = {};
	cudaLaunchAttribute at[1];
	cfg.gridDim = dim3(148 / cs * cs); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = kSmem;
	if (cs > 1) {
		at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
		cfg.attrs = at; cfg.numAttrs = 1;
		int nc = 0;
		CK(cudaOccupancyMaxActiveClusters(&nc, kernel, &cfg));
		printf("  [%s: max active clusters of %d = %d -> %d CTAs]\n", name, cs, nc, nc * cs);
		cfg.gridDim = dim3(nc * cs < 148 ? nc * cs : 148 / cs * cs);
	}
	cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	cudaStream_t st; CK(cudaStreamCreate(&st));
	cfg.stream = st;
	for (int i = 0; i < 5; i++) CK(cudaLaunchKernelEx(&cfg, kernel, g, sink));
	CK(cudaDeviceSynchronize());
	// a graph of N launches: the host's launch rate is out of the picture
	const int N = 200;
	cudaGraph_t graph; cudaGraphExec_t exec;
	CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
	for (int i = 0; i < N; i++) CK(cudaLaunchKernelEx(&cfg, kernel, g, sink));
	CK(cudaStreamEndCapture(st, &graph));
	CK(cudaGraphInstantiate(&exec, graph, 0));
	CK(cudaGraphLaunch(exec, st)); CK(cudaStreamSynchronize(st));
	CK(cudaEventRecord(e0, st));
	CK(cudaGraphLaunch(exec, st));
	CK(cudaEventRecord(e1, st)); CK(cudaEventSynchronize(e1));
	float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
	printf("%-28s grid %3u: %.2f us per launch (graph of 200)\n", name, cfg.gridDim.x, 1e3 * ms / N);
	return 0;
}

int main()
{
	uint4 *g; unsigned *sink;
	CK(cudaMalloc(&g, kBytes)); CK(cudaMemset(g, 1, kBytes)); CK(cudaMalloc(&sink, 4));
	if (timeit("A empty", k_empty, 1, g, sink)) return 1;
	if (timeit("B ld.global/st.shared", k_ldst, 1, g, sink)) return 1;
	if (timeit("C bulk copy unicast", k_tma, 1, g, sink)) return 1;
	if (timeit("D bulk multicast cluster 2", k_mc<2>, 2, g, sink)) return 1;
	if (timeit("E bulk multicast cluster 4", k_mc<4>, 4, g, sink)) return 1;
	return 0;
}
