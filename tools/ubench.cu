// ubench.cu -- sm_100a micro-benchmarks behind the design of the fused channel kernel:
// dependent-issue latency and per-SMSP throughput of the instructions on its critical paths,
// and the warp scheduler's priority between a latency-bound and a throughput-bound warp.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench tools/ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int ITER = 256;

enum Op { FADD, FMUL, FFMA, FADD2, FFMA2, IADD3, LOP3, SHF, I2FP, FMNMX, IMAD, IMADHI, LDS16, LDS64, LDS128, MIXAF, P2P2, P2S1, P2S2, P2I1, S1S1, NOPS };
static const char *names[] = { "FADD", "FMUL", "FFMA", "FADD2", "FFMA2", "IADD3", "LOP3", "SHF", "I2FP", "FMNMX", "IMAD", "IMAD.HI", "LDS.S16", "LDS.64", "LDS.128", "IADD3+FFMA", "FFMA2+FADD2", "FFMA2+FFMA", "FFMA2+2xFFMA", "FFMA2+IADD3", "FFMA+FADD" };

template <int OP, int CH>
__device__ __forceinline__ void body(float (&f)[CH], unsigned (&u)[CH], unsigned long long (&p)[CH], float c, unsigned ci, uint32_t sbase)
{
	#pragma unroll
	for (int k = 0; k < CH; k++) {
		if (OP == FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[k]) : "f"(c));
		if (OP == FMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[k]) : "f"(c));
		if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[k]) : "f"(c));
		if (OP == FADD2) asm volatile("add.rn.f32x2 %0, %0, %0;" : "+l"(p[k]));
		if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[k]));
		if (OP == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[k]) : "r"(ci));
		if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %1, 0x6a;" : "+r"(u[k]) : "r"(ci));
		if (OP == SHF) asm volatile("shf.r.wrap.b32 %0, %0, %0, %1;" : "+r"(u[k]) : "r"(ci));
		if (OP == I2FP) { float t; asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(t) : "r"(u[k])); u[k] = __float_as_uint(t); }
		if (OP == FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[k]) : "f"(c));
		if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(u[k]) : "r"(ci));
		if (OP == IMADHI) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(u[k]) : "r"(ci));
		if (OP == LDS16) { int t; asm volatile("ld.shared.s16 %0, [%1];" : "=r"(t) : "r"(sbase + (u[k] & 0xFFFEu))); u[k] += t; }
		if (OP == LDS64) { unsigned a, b; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "r"(sbase + (u[k] & 0xFFF8u))); u[k] += a + b; }
		if (OP == LDS128) { unsigned a, b, cc, d; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(cc), "=r"(d) : "r"(sbase + (u[k] & 0xFFF0u))); u[k] += a + b + cc + d; }
		if (OP == MIXAF) { asm volatile("add.u32 %0, %0, %1;" : "+r"(u[k]) : "r"(ci)); asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[k]) : "f"(c)); }
		// what binds the channel kernels: packed f32x2 next to packed, scalar and integer work (independent chains)
		if (OP == P2P2) { asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[k])); unsigned long long q = (unsigned long long)u[k] << 32 | __float_as_uint(f[k]); asm volatile("add.rn.f32x2 %0, %0, %0;" : "+l"(q)); u[k] = (unsigned)(q >> 32); f[k] = __uint_as_float((unsigned)q); }
		if (OP == P2S1) { asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[k])); asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[k]) : "f"(c)); }
		if (OP == P2S2) { asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[k])); asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[k]) : "f"(c)); float g = __uint_as_float(u[k]); asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(g) : "f"(c)); u[k] = __float_as_uint(g); }
		if (OP == P2I1) { asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[k])); asm volatile("add.u32 %0, %0, %1;" : "+r"(u[k]) : "r"(ci)); }
		if (OP == S1S1) { asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[k]) : "f"(c)); float g = __uint_as_float(u[k]); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(g) : "f"(c)); u[k] = __float_as_uint(g); }
	}
}

// one CTA, nw warps; every warp runs ITER x CH ops; cycles measured per warp, max reported
template <int OP, int CH>
__global__ void k_tput(float c, unsigned ci, long long *out, unsigned *sink)
{
	extern __shared__ unsigned char sm[];
	for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<unsigned*>(sm)[i] = (i * 2654435761u) & 0x7u;
	__syncthreads();
	const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
	float f[CH]; unsigned u[CH]; unsigned long long p[CH];
	#pragma unroll
	for (int k = 0; k < CH; k++) { f[k] = threadIdx.x * 0.001f + k; u[k] = threadIdx.x * 2654435761u + k * 40503u; p[k] = ((unsigned long long)__float_as_uint(f[k]) << 32) | __float_as_uint(f[k]); }
	__syncthreads();
	long long t0 = clock64();
	#pragma unroll 1
	for (int it = 0; it < ITER; it++)
		body<OP, CH>(f, u, p, c, ci, sbase);
	long long t1 = clock64();
	unsigned acc = 0;
	#pragma unroll
	for (int k = 0; k < CH; k++) acc += __float_as_uint(f[k]) + u[k] + (unsigned)p[k] + (unsigned)(p[k] >> 32);
	if (acc == 0x12345678u) sink[0] = acc;
	if ((threadIdx.x & 31) == 0) out[threadIdx.x >> 5] = t1 - t0;
}

// priority: warps with wid%4==0 only do work (others exit). `chainw` = which of those runs the
// dependent FADD2 chain; the rest run an ALU/FMA throughput loop until the chain warp is done.
__global__ void k_prio(int chainw, int nwork, long long *out, unsigned *sink, float c)
{
	__shared__ volatile int done;
	const int wid = threadIdx.x >> 5;
	if (threadIdx.x == 0) done = 0;
	__syncthreads();
	if (wid % 4 != 0 || wid / 4 >= nwork) return;
	const int w = wid / 4;
	if (w == chainw) {
		unsigned long long p = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
		long long t0 = clock64();
		#pragma unroll 1
		for (int it = 0; it < 64; it++) {
			#pragma unroll
			for (int k = 0; k < 32; k++) asm volatile("add.rn.f32x2 %0, %0, %0;" : "+l"(p));
		}
		long long t1 = clock64();
		if ((threadIdx.x & 31) == 0) { out[0] = t1 - t0; done = 1; }
		if ((unsigned)p == 0x12345678u) sink[0] = 1;
	} else {
		float f[8]; unsigned u[8];
		#pragma unroll
		for (int k = 0; k < 8; k++) { f[k] = threadIdx.x + k; u[k] = threadIdx.x * 77u + k; }
		long long n = 0;
		while (!done) {
			#pragma unroll
			for (int k = 0; k < 8; k++) { asm volatile("add.u32 %0, %0, %1;" : "+r"(u[k]) : "r"(77u)); asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[k]) : "f"(c)); }
			n++;
		}
		unsigned acc = 0;
		#pragma unroll
		for (int k = 0; k < 8; k++) acc += __float_as_uint(f[k]) + u[k];
		if (acc == 0x12345678u) sink[0] = acc;
		if ((threadIdx.x & 31) == 0) out[1 + w] = n;
	}
}

template <int OP, int CH>
int run(const char *what, int nwarps, long long *d_out, unsigned *d_sink)
{
	CK(cudaFuncSetAttribute(k_tput<OP, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
	k_tput<OP, CH><<<1, nwarps * 32, 65536>>>(1.0000001f, 3u, d_out, d_sink);
	CK(cudaDeviceSynchronize());
	k_tput<OP, CH><<<1, nwarps * 32, 65536>>>(1.0000001f, 3u, d_out, d_sink);
	CK(cudaDeviceSynchronize());
	long long h[32];
	CK(cudaMemcpy(h, d_out, sizeof(long long) * nwarps, cudaMemcpyDeviceToHost));
	long long mx = 0;
	for (int i = 0; i < nwarps; i++) mx = h[i] > mx ? h[i] : mx;
	const double ops = (double)ITER * CH * ((OP == MIXAF || OP == P2P2 || OP == P2S1 || OP == P2I1 || OP == S1S1) ? 2 : OP == P2S2 ? 3 : 1);
	// per SMSP: nwarps/4 warps share one scheduler
	printf("%-12s %-10s warps=%2d chains=%d  cycles/op/warp=%7.3f  cycles per warp-instr per SMSP=%6.3f\n",
			names[OP], what, nwarps, CH, mx / ops, mx / (ops * (nwarps >= 4 ? nwarps / 4 : 1)));
	return 0;
}

int main()
{
	long long *d_out; unsigned *d_sink;
	CK(cudaMalloc(&d_out, sizeof(long long) * 64));
	CK(cudaMalloc(&d_sink, 16));
	printf("== dependent-issue latency (1 warp, 1 chain)\n");
	run<FADD, 1>("latency", 1, d_out, d_sink);
	run<FMUL, 1>("latency", 1, d_out, d_sink);
	run<FFMA, 1>("latency", 1, d_out, d_sink);
	run<FADD2, 1>("latency", 1, d_out, d_sink);
	run<FFMA2, 1>("latency", 1, d_out, d_sink);
	run<IADD3, 1>("latency", 1, d_out, d_sink);
	run<I2FP, 1>("latency", 1, d_out, d_sink);
	run<IMADHI, 1>("latency", 1, d_out, d_sink);
	run<LDS16, 1>("latency", 1, d_out, d_sink);
	run<LDS128, 1>("latency", 1, d_out, d_sink);
	printf("== throughput (32 warps = 8 per SMSP, 8 independent chains each)\n");
	run<FADD, 8>("tput", 32, d_out, d_sink);
	run<FMUL, 8>("tput", 32, d_out, d_sink);
	run<FFMA, 8>("tput", 32, d_out, d_sink);
	run<FADD2, 8>("tput", 32, d_out, d_sink);
	run<FFMA2, 8>("tput", 32, d_out, d_sink);
	run<IADD3, 8>("tput", 32, d_out, d_sink);
	run<LOP3, 8>("tput", 32, d_out, d_sink);
	run<SHF, 8>("tput", 32, d_out, d_sink);
	run<I2FP, 8>("tput", 32, d_out, d_sink);
	run<FMNMX, 8>("tput", 32, d_out, d_sink);
	run<IMAD, 8>("tput", 32, d_out, d_sink);
	run<IMADHI, 8>("tput", 32, d_out, d_sink);
	run<LDS16, 8>("tput", 32, d_out, d_sink);
	run<LDS64, 8>("tput", 32, d_out, d_sink);
	run<LDS128, 8>("tput", 32, d_out, d_sink);
	run<MIXAF, 8>("tput", 32, d_out, d_sink);
	printf("== packed f32x2 next to other work (per warp-INSTRUCTION; a pair/triple counts 2/3)\n");
	for (int nw : {32, 16, 8}) {
		run<FFMA2, 8>("tput", nw, d_out, d_sink);
		run<FFMA, 8>("tput", nw, d_out, d_sink);
		run<P2P2, 8>("tput", nw, d_out, d_sink);
		run<P2S1, 8>("tput", nw, d_out, d_sink);
		run<P2S2, 8>("tput", nw, d_out, d_sink);
		run<P2I1, 8>("tput", nw, d_out, d_sink);
		run<S1S1, 8>("tput", nw, d_out, d_sink);
	}
	printf("== scheduler priority: one FADD2-chain warp (2048 dependent ops) vs throughput warps on the same SMSP\n");
	for (int nwork : {1, 2, 6}) {
		for (int chainw : {0, nwork - 1}) {
			k_prio<<<1, 1024>>>(chainw, nwork, d_out, d_sink, 1.0000001f);
			CK(cudaDeviceSynchronize());
			long long h[16];
			CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
			printf("workers=%d chain warp slot=%d (wid %2d): chain cycles/op = %.2f\n", nwork, chainw, chainw * 4, h[0] / 2048.0);
			if (nwork == 1) break;
		}
	}
	return 0;
}
