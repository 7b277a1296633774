// ubench_dsmem.cu -- can a full fp32 NCO table (65536 entries, 256 KB) live in the shared memory of a
// 2-CTA cluster, half per SM, and be gathered from at the rate the channel kernel needs?
// The kernel needs two random 4-byte lookups per receiver-frame and processes ~1.7 receiver-frames per
// cycle and SM today (2.3 if the table reconstruction's instructions went away): 3.5 - 4.6 lookups per
// cycle and SM, half of them in the peer's shared memory.
// Measures random gathers per cycle and SM: plain ld.shared, ld.shared::cluster on the own CTA, half
// remote, all remote; lanes scattered (the kernel's pattern) and lanes contiguous.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_dsmem tools/ubench_dsmem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int kEntries = 32768;          // per CTA: 128 KB
constexpr int ITER = 512, UN = 16;

enum Mode { PLAIN, OWN, HALF, REMOTE, REMOTE_COALESCED, HALF_COALESCED };
static const char *names[] = { "ld.shared (own, plain)", "ld.shared::cluster own CTA", "ld.shared::cluster half remote",
		"ld.shared::cluster all remote", "all remote, lanes contiguous", "half remote, lanes contiguous" };

template <int MODE>
__global__ void __launch_bounds__(256, 1) gather(unsigned *out, unsigned long long *cycles)
{
	extern __shared__ __align__(16) unsigned tbl[];
	const unsigned tid = threadIdx.x;
	for (unsigned i = tid; i < kEntries; i += blockDim.x)
		tbl[i] = i * 2654435761u;
	unsigned rank;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
	const uint32_t own = (uint32_t)__cvta_generic_to_shared(tbl);
	uint32_t a0, a1;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a0) : "r"(own), "r"(rank));
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a1) : "r"(own), "r"(rank ^ 1u));
	asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
	unsigned x = tid * 747796405u + blockIdx.x * 2891336453u + 1u, acc = 0;
	const unsigned long long t0 = clock64();
	for (int it = 0; it < ITER; it++) {
		unsigned v[UN];
		#pragma unroll
		for (int k = 0; k < UN; k++) {
			x = x * 1664525u + 1013904223u;
			unsigned idx = x >> 16;                                  // 16 bits: the table index
			if (MODE == REMOTE_COALESCED || MODE == HALF_COALESCED)
				idx = (idx & 0xFFE0u) | (tid & 31u);                 // a warp reads 32 consecutive entries
			const unsigned off = (idx & 0x7FFFu) * 4u;
			if (MODE == PLAIN) {
				asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[k]) : "r"(own + off));
			} else {
				uint32_t base = a0;
				if (MODE == REMOTE || MODE == REMOTE_COALESCED) base = a1;
				if (MODE == HALF) base = (idx & 0x8000u) ? a1 : a0;
				if (MODE == HALF_COALESCED) base = (__shfl_sync(0xffffffffu, idx, 0) & 0x8000u) ? a1 : a0;
				asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v[k]) : "r"(base + off));
			}
		}
		#pragma unroll
		for (int k = 0; k < UN; k++)
			acc += v[k];
	}
	const unsigned long long t1 = clock64();
	// nobody leaves while a peer may still read its table
	asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
	out[blockIdx.x * blockDim.x + tid] = acc;
	if (tid == 0)
		cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
static int run(int nsm, unsigned *d_out, unsigned long long *d_cyc)
{
	const size_t smem = sizeof(unsigned) * kEntries;
	CK(cudaFuncSetAttribute(gather<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(nsm & ~1);
	cfg.blockDim = dim3(256);
	cfg.dynamicSmemBytes = smem;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeClusterDimension;
	at[0].val.clusterDim.x = 2;
	at[0].val.clusterDim.y = 1;
	at[0].val.clusterDim.z = 1;
	cfg.attrs = at;
	cfg.numAttrs = 1;
	for (int rep = 0; rep < 2; rep++)
		CK(cudaLaunchKernelEx(&cfg, gather<MODE>, d_out, d_cyc));
	CK(cudaDeviceSynchronize());
	unsigned long long cyc[256];
	CK(cudaMemcpy(cyc, d_cyc, sizeof(unsigned long long) * (nsm & ~1), cudaMemcpyDeviceToHost));
	double worst = 0, sum = 0;
	for (int i = 0; i < (nsm & ~1); i++) {
		sum += (double)cyc[i];
		if ((double)cyc[i] > worst) worst = (double)cyc[i];
	}
	const double lookups = 256.0 * ITER * UN;
	printf("%-36s %7.2f lookups / cycle / SM (mean CTA), %7.2f (slowest CTA); %6.1f cycles per warp-instruction\n", names[MODE],
			lookups / (sum / (nsm & ~1)), lookups / worst, (sum / (nsm & ~1)) / (8.0 * ITER * UN) * 8.0 / 8.0);
	return 0;
}

int main()
{
	cudaDeviceProp p;
	CK(cudaGetDeviceProperties(&p, 0));
	printf("%s, %d SMs; 2-CTA clusters, 256 threads and a 128 KB table per CTA, %d gathers per thread\n", p.name, p.multiProcessorCount, ITER * UN);
	unsigned *d_out;
	unsigned long long *d_cyc;
	CK(cudaMalloc(&d_out, sizeof(unsigned) * 256 * 256));
	CK(cudaMalloc(&d_cyc, sizeof(unsigned long long) * 256));
	if (run<PLAIN>(p.multiProcessorCount, d_out, d_cyc)) return 1;
	if (run<OWN>(p.multiProcessorCount, d_out, d_cyc)) return 1;
	if (run<HALF>(p.multiProcessorCount, d_out, d_cyc)) return 1;
	if (run<REMOTE>(p.multiProcessorCount, d_out, d_cyc)) return 1;
	if (run<REMOTE_COALESCED>(p.multiProcessorCount, d_out, d_cyc)) return 1;
	if (run<HALF_COALESCED>(p.multiProcessorCount, d_out, d_cyc)) return 1;
	return 0;
}
