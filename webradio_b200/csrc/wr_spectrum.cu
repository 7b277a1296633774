// wr_spectrum.cu -- SpectrumSink on the GPU (K5 + K6 of SURVEY.md 2a): Hamming window fused
// into the load, Stockham autosort FFT held entirely in shared memory, and the
// 10*log10(|X|^2) - 20*log10(N) conversion with the fft-shift fused into the store
// (reference src/io/spectrumsink.cxx:88-142).
//
// One CTA transforms one FFT frame of one stream.  A frame may straddle the carry buffer
// (frames left over from the previous call) and the new block, so the kernel reads through a
// two-segment view instead of first concatenating them in HBM.
#include "wr_common.h"
#include "wr_fft.cuh"
#include "wr_device.cuh"
#include "wr_upload.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace {

struct SpecArgs {
	const float2 *carry;     // [T][carry_stride] frames left from earlier calls
	size_t carry_stride;
	unsigned ncarry;         // valid frames in carry (same for every stream)
	const float2 *in;        // [T][in_stride] this call's frames
	size_t in_stride;
	const float *window;     // [N]
	const float2 *twiddle;   // [N] exp(-2*pi*i*k/N)
	float *rows;             // may be null
	size_t row_stride;       // floats between streams
	float *last;             // [T][N] most recent row per stream
	unsigned N, logN, hop;
	unsigned first_row;      // blockIdx.x + first_row = frame index within this call
	unsigned nrows;          // rows completed by this call
	float scaledb;
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
	return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__device__ __forceinline__ unsigned long long pack2(float lo, float hi)
{
	unsigned long long r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
	return r;
}

__device__ __forceinline__ float2 view(const SpecArgs &a, unsigned t, size_t pos)
{
	if (pos < a.ncarry)
		return a.carry[(size_t)t * a.carry_stride + pos];
	return __ldg(a.in + (size_t)t * a.in_stride + (pos - a.ncarry));
}

// Stockham autosort passes (natural order in, natural order out).  sa/sb ping-pong.
__global__ void __launch_bounds__(1024) spectrum_kernel_v1(const SpecArgs a)
{
	extern __shared__ float2 wr_fft_smem[];
	const unsigned N = a.N;
	float2 *sa = wr_fft_smem;
	float2 *sb = wr_fft_smem + N;
	const unsigned t = blockIdx.y;
	const unsigned m = blockIdx.x + a.first_row;
	const size_t start = (size_t)m * a.hop;
	const unsigned tid = threadIdx.x, nt = blockDim.x;

	// window fused into the load: inbuf[n] *= window[n] (spectrumsink.cxx:110-113)
	for (unsigned i = tid; i < N; i += nt) {
		float2 x = view(a, t, start + i);
		float w = a.window[i];
		sa[i] = make_float2(__fmul_rn(x.x, w), __fmul_rn(x.y, w));
	}
	__syncthreads();

	unsigned Ns = 1;
	if (a.logN & 1) { // one radix-2 pass first (Ns = 1: no twiddles)
		for (unsigned j = tid; j < N / 2; j += nt) {
			float2 u = sa[j], v = sa[j + N / 2];
			sb[2 * j] = make_float2(u.x + v.x, u.y + v.y);
			sb[2 * j + 1] = make_float2(u.x - v.x, u.y - v.y);
		}
		__syncthreads();
		float2 *tmp = sa; sa = sb; sb = tmp;
		Ns = 2;
	}
	const unsigned Q = N / 4;
	for (; Ns < N; Ns *= 4) {
		const unsigned tw_stride = N / (Ns * 4);
		for (unsigned j = tid; j < Q; j += nt) {
			const unsigned k = j & (Ns - 1);
			float2 v0 = sa[j], v1 = sa[j + Q], v2 = sa[j + 2 * Q], v3 = sa[j + 3 * Q];
			if (Ns > 1) {
				const unsigned q = k * tw_stride;
				v1 = cmul(v1, __ldg(a.twiddle + q));
				v2 = cmul(v2, __ldg(a.twiddle + 2 * q));
				v3 = cmul(v3, __ldg(a.twiddle + 3 * q));
			}
			float2 s02 = make_float2(v0.x + v2.x, v0.y + v2.y);
			float2 d02 = make_float2(v0.x - v2.x, v0.y - v2.y);
			float2 s13 = make_float2(v1.x + v3.x, v1.y + v3.y);
			float2 d13 = make_float2(v1.y - v3.y, v3.x - v1.x); // (v1 - v3) * (-i)
			const unsigned base = (j - k) * 4 + k;
			sb[base] = make_float2(s02.x + s13.x, s02.y + s13.y);
			sb[base + Ns] = make_float2(d02.x + d13.x, d02.y + d13.y);
			sb[base + 2 * Ns] = make_float2(s02.x - s13.x, s02.y - s13.y);
			sb[base + 3 * Ns] = make_float2(d02.x - d13.x, d02.y - d13.y);
		}
		__syncthreads();
		float2 *tmp = sa; sa = sb; sb = tmp;
	}

	// dB + fft-shift fused into the store (spectrumsink.cxx:127-140)
	float *row = a.rows ? a.rows + (size_t)t * a.row_stride + (size_t)m * N : nullptr;
	float *last = (m == a.nrows - 1) ? a.last + (size_t)t * N : nullptr;
	const unsigned half = N / 2;
	for (unsigned o = tid; o < N; o += nt) {
		const unsigned n = (o + half) & (N - 1); // output slot o holds bin n
		float2 X = sa[n];
		float p = __fadd_rn(__fmul_rn(X.x, X.x), __fmul_rn(X.y, X.y));
		float db = __fsub_rn(__fmul_rn(10.0f, log10f(p)), a.scaledb);
		if (row) row[o] = db;
		if (last) last[o] = db;
	}
}

// Register-blocked transform, N = R1 * 256 (see wr_fft.cuh): one 256-thread CTA per FFT frame.
constexpr int kRowPitch = 257; // float2 per 256-point row: odd pitch keeps pass 3's column reads conflict-free

template <int R1>
__global__ void __launch_bounds__(256, 2) spectrum_kernel_v2(const SpecArgs a)
{
	extern __shared__ float2 wr_fft_smem[];
	constexpr unsigned N = R1 * 256;
	float2 *sm = wr_fft_smem;
	float2 *tw2 = wr_fft_smem + R1 * kRowPitch;     // pass-2 twiddles W_256^(b*ka), laid out [ka][b]
	const unsigned t = blockIdx.y;
	const unsigned m = blockIdx.x + a.first_row;
	const size_t start = (size_t)m * a.hop;
	const unsigned tid = threadIdx.x;

	tw2[tid] = __ldg(a.twiddle + ((((tid & 15) * (tid >> 4)) & 255) * R1));
	// the frame usually lies entirely inside this call's block: one pointer, no per-sample test
	const bool contiguous = start >= a.ncarry;
	const float2 *__restrict__ src = a.in + (size_t)t * a.in_stride + (start - a.ncarry);

	// ---- pass 1: window fused into the load, radix-R1 over the stride-256 index ----
	{
		float2 v[R1];
		if (contiguous) {
			// all R1 loads issued back to back (one HBM round trip per thread)
			#pragma unroll
			for (int j = 0; j < R1; j++)
				v[j] = __ldg(src + tid + 256u * j);
		} else {
			#pragma unroll
			for (int j = 0; j < R1; j++)
				v[j] = view(a, t, start + tid + 256u * j);
		}
		#pragma unroll
		for (int j = 0; j < R1; j++) {
			const float w = __ldg(a.window + tid + 256u * j);
			// inbuf[n] *= window[n] (spectrumsink.cxx:110-113): both components in one packed multiply
			asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&v[j]))
					: "l"(*reinterpret_cast<const unsigned long long*>(&v[j])), "l"(pack2(w, w)));
		}
		wrfft::RegDft<R1>::run(v);
		// twiddle W_N^(tid * k1), built up from W_N^tid by repeated multiplication
		const float2 w1 = __ldg(a.twiddle + tid);
		float2 cur = w1;
		sm[tid] = v[0];
		#pragma unroll
		for (int k1 = 1; k1 < R1; k1++) {
			sm[k1 * kRowPitch + tid] = wrfft::cmul(v[k1], cur);
			cur = wrfft::cmul(cur, w1);
		}
	}
	__syncthreads();

	// ---- pass 2: radix-16 over the stride-16 index of every 256-point row ----
	#pragma unroll
	for (unsigned i = 0; i < (R1 * 16 + 255) / 256; i++) {
		const unsigned bf = tid + 256 * i;
		if (bf < R1 * 16) {
			const unsigned b = bf & 15, k1 = bf >> 4;
			float2 *row = sm + k1 * kRowPitch + b;
			float2 u[16];
			#pragma unroll
			for (int q = 0; q < 16; q++)
				u[q] = row[16 * q];
			wrfft::RegDft<16>::run(u);
			#pragma unroll
			for (int ka = 0; ka < 16; ka++)
				row[16 * ka] = ka ? wrfft::cmul(u[ka], tw2[16 * ka + b]) : u[0];
		}
	}
	__syncthreads();

	// ---- pass 3: radix-16 over the contiguous index; dB + fft-shift fused into the store ----
	float *rowout = a.rows ? a.rows + (size_t)t * a.row_stride + (size_t)m * N : nullptr;
	float *last = (m == a.nrows - 1) ? a.last + (size_t)t * N : nullptr;
	#pragma unroll
	for (unsigned i = 0; i < (R1 * 16 + 255) / 256; i++) {
		const unsigned bf = tid + 256 * i;
		if (bf < R1 * 16) {
			const unsigned k1 = bf % R1, ka = bf / R1;
			const float2 *row = sm + k1 * kRowPitch + 16 * ka;
			float2 u[16];
			#pragma unroll
			for (int q = 0; q < 16; q++)
				u[q] = row[q];
			wrfft::RegDft<16>::run(u);
			#pragma unroll
			for (int kb = 0; kb < 16; kb++) {
				const unsigned k = bf + R1 * 16 * kb;           // = k1 + R1 * (ka + 16 * kb)
				const unsigned o = (k + N / 2) & (N - 1);        // fft-shift (spectrumsink.cxx:139)
				const float p = fmaf(u[kb].x, u[kb].x, u[kb].y * u[kb].y);
				// 10*log10(p) = (10*log10(2)) * log2(p); the hardware log2 is good to ~1e-6 dB here,
				// far inside the 1e-5 magnitude tolerance, and keeps log(0) = -inf.  The .ftz form is
				// one MUFU without the denormal rescue around it: a power below 1e-38 (-380 dB) reads
				// as -inf, which the waterfall handler maps to its floor anyway.
				float l2;
				asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(p));
				const float db = fmaf(3.01029995663981195f, l2, -a.scaledb);
				if (rowout) rowout[o] = db;
				if (last) last[o] = db;
			}
		}
	}
}

// ---- v3: persistent, the frames arrive by bulk copies (TMA) behind mbarriers ---------------------
// One 256-thread CTA per SM walks a run of consecutive FFT frames of one stream.  The frames'
// samples are fetched by cp.async.bulk, hop frames (a "chunk") at a time, into a ring of N/hop + 1
// chunks in shared memory: while frame m is transformed, the chunk that completes frame m+1 is
// already landing, and as soon as pass 1 has consumed the oldest chunk its slot is handed to the
// copy for frame m+2.  Overlapping frames (hop = N/2, BASELINE config 4) share their common half
// in shared memory instead of fetching it twice.  The transform is v2's (radix-R1 in registers,
// radix-16, radix-16; wr_fft.cuh), so are window and dB epilogue.
struct SpecArgs3 {
	SpecArgs a;
	unsigned row0;           // first row this launch computes (rows in front of it straddle the carry buffer: v2)
	unsigned rowsPerRun;     // consecutive rows of one stream a CTA takes at a time
	unsigned runsPerStream;
	unsigned nStreams;
};

__device__ __forceinline__ void spec_mbar_init(uint32_t bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void spec_mbar_wait(uint32_t bar, unsigned parity)
{
	asm volatile("{\n\t.reg .pred p;\n"
			"WR_SPEC_WAIT%=:\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
			"@!p bra WR_SPEC_WAIT%=;\n\t}" :: "r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ void spec_bulk_load(uint32_t dst, const void *src, unsigned bytes, uint32_t bar)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
	// (a single copy may carry at most 2^20 - 16 bytes; chunks here are at most 64 KiB)
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
			:: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int R1, int C>     // C = N / hop: chunks per frame (1: hop = N, 2: hop = N/2)
__global__ void __launch_bounds__(256) spectrum_kernel_v3(const SpecArgs3 g)
{
	extern __shared__ __align__(16) float2 wr_fft_smem[];
	const SpecArgs &a = g.a;
	constexpr unsigned N = R1 * 256, HOP = N / C, NSLOT = C + 1;
	float2 *sm = wr_fft_smem;                               // work buffer, R1 rows of kRowPitch
	float2 *tw2 = wr_fft_smem + R1 * kRowPitch;             // pass-2 twiddles
	float2 *ring = tw2 + 256;                               // NSLOT chunks of HOP samples
	__shared__ __align__(8) unsigned long long wr_spec_bars[NSLOT];
	const unsigned tid = threadIdx.x;
	const uint32_t bar32 = (uint32_t)__cvta_generic_to_shared(wr_spec_bars);
	const uint32_t ring32 = (uint32_t)__cvta_generic_to_shared(ring);
	if (tid == 0) {
		for (unsigned i = 0; i < NSLOT; i++)
			spec_mbar_init(bar32 + 8u * i, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	tw2[tid] = __ldg(a.twiddle + ((((tid & 15) * (tid >> 4)) & 255) * R1));
	// the window of this thread's R1 samples and its pass-1 twiddle base are the same for every frame
	float win[R1];
	#pragma unroll
	for (int j = 0; j < R1; j++)
		win[j] = __ldg(a.window + tid + 256u * j);
	const float2 w1 = __ldg(a.twiddle + tid);
	__syncthreads();

	const unsigned nrows3 = a.nrows - g.row0;
	const unsigned totalRuns = g.runsPerStream * g.nStreams;
	unsigned phases = 0;                                    // bit s: parity of the next completion of slot s to wait for
	for (unsigned run = blockIdx.x; run < totalRuns; run += gridDim.x) {
		const unsigned t = run / g.runsPerStream;
		const unsigned r0 = (run - t * g.runsPerStream) * g.rowsPerRun;
		const unsigned nr = min(g.rowsPerRun, nrows3 - r0);
		const unsigned mFirst = g.row0 + r0;
		// chunk c of this run starts at frame (mFirst + c) * HOP of [carry | in]; all of it lies in `in`
		const float2 *src = a.in + (size_t)t * a.in_stride + ((size_t)mFirst * HOP - a.ncarry);
		const unsigned nchunks = nr + C - 1;
		// prologue: the chunks of the first frame and the one that completes the second
		if (tid == 0) {
			for (unsigned c = 0; c < NSLOT && c < nchunks; c++)
				spec_bulk_load(ring32 + (c % NSLOT) * HOP * 8u, src + (size_t)c * HOP, HOP * 8u, bar32 + 8u * (c % NSLOT));
		}
		for (unsigned f = 0; f < nr; f++) {
			const unsigned m = mFirst + f;
			// ---- pass 1: the frame's chunks have landed; window fused into the read ----
			float2 v[R1];
			#pragma unroll
			for (unsigned c = 0; c < (unsigned)C; c++) {
				const unsigned slot = (f + c) % NSLOT;
				// each slot is waited for once per chunk it receives; chunk f + c is new to this frame
				// only if c == C - 1 (or this is the run's first frame)
				if (c == (unsigned)C - 1 || f == 0) {
					spec_mbar_wait(bar32 + 8u * slot, (phases >> slot) & 1u);
					phases ^= 1u << slot;
				}
				const float2 *chunk = ring + slot * HOP;
				#pragma unroll
				for (int j = 0; j < R1 / C; j++)
					v[c * (R1 / C) + j] = chunk[tid + 256u * j];
			}
			#pragma unroll
			for (int j = 0; j < R1; j++) {
				// inbuf[n] *= window[n] (spectrumsink.cxx:110-113): both components in one packed multiply
				asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&v[j]))
						: "l"(*reinterpret_cast<const unsigned long long*>(&v[j])), "l"(pack2(win[j], win[j])));
			}
			wrfft::RegDft<R1>::run(v);
			// every thread is done with the previous frame's pass 3 (work buffer) and with this
			// frame's oldest chunk (ring)
			__syncthreads();
			if (tid == 0 && f + NSLOT < nchunks) {
				const unsigned c = f + NSLOT;               // the chunk that completes frame f + 2
				spec_bulk_load(ring32 + (c % NSLOT) * HOP * 8u, src + (size_t)c * HOP, HOP * 8u, bar32 + 8u * (c % NSLOT));
			}
			{
				float2 cur = w1;
				sm[tid] = v[0];
				#pragma unroll
				for (int k1 = 1; k1 < R1; k1++) {
					sm[k1 * kRowPitch + tid] = wrfft::cmul(v[k1], cur);
					cur = wrfft::cmul(cur, w1);
				}
			}
			__syncthreads();
			// ---- pass 2: radix-16 over the stride-16 index of every 256-point row ----
			// (a thread's butterflies are loaded TOGETHER before the first is computed: with one CTA of
			// eight warps per SM nobody else hides the shared-memory latency)
			constexpr unsigned NB = (R1 * 16 + 255) / 256;
			{
				float2 u[NB][16];
				#pragma unroll
				for (unsigned i = 0; i < NB; i++) {
					const unsigned bf = tid + 256 * i;
					if (bf < R1 * 16) {
						const float2 *row = sm + (bf >> 4) * kRowPitch + (bf & 15);
						#pragma unroll
						for (int q = 0; q < 16; q++)
							u[i][q] = row[16 * q];
					}
				}
				#pragma unroll
				for (unsigned i = 0; i < NB; i++) {
					const unsigned bf = tid + 256 * i;
					if (bf < R1 * 16) {
						const unsigned b = bf & 15, k1 = bf >> 4;
						float2 *row = sm + k1 * kRowPitch + b;
						wrfft::RegDft<16>::run(u[i]);
						#pragma unroll
						for (int ka = 0; ka < 16; ka++)
							row[16 * ka] = ka ? wrfft::cmul(u[i][ka], tw2[16 * ka + b]) : u[i][0];
					}
				}
			}
			__syncthreads();
			// ---- pass 3: radix-16 over the contiguous index; dB + fft-shift fused into the store ----
			float *rowout = a.rows ? a.rows + (size_t)t * a.row_stride + (size_t)m * N : nullptr;
			float *last = (m == a.nrows - 1) ? a.last + (size_t)t * N : nullptr;
			{
				float2 u[NB][16];
				#pragma unroll
				for (unsigned i = 0; i < NB; i++) {
					const unsigned bf = tid + 256 * i;
					if (bf < R1 * 16) {
						const float2 *row = sm + (bf % R1) * kRowPitch + 16 * (bf / R1);
						#pragma unroll
						for (int q = 0; q < 16; q++)
							u[i][q] = row[q];
					}
				}
				#pragma unroll
				for (unsigned i = 0; i < NB; i++) {
					const unsigned bf = tid + 256 * i;
					if (bf < R1 * 16) {
						wrfft::RegDft<16>::run(u[i]);
						#pragma unroll
						for (int kb = 0; kb < 16; kb++) {
							const unsigned k = bf + R1 * 16 * kb;
							const unsigned o = (k + N / 2) & (N - 1);
							const float p = fmaf(u[i][kb].x, u[i][kb].x, u[i][kb].y * u[i][kb].y);
							float l2;
							asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(p));
							const float db = fmaf(3.01029995663981195f, l2, -a.scaledb);
							if (rowout) rowout[o] = db;
							if (last) last[o] = db;
						}
					}
				}
			}
		}
		// the next run reuses ring and work buffer
		__syncthreads();
	}
}

// ---- v4: 8192-point transforms, TWO CTAs per SM ------------------------------------------------
// v3 keeps one CTA of eight warps per SM (64 KB work buffer + 96 KB ring): two warps per scheduler
// that meet at four block-wide barriers per frame, the schedulers 73 % busy by the cost model of
// DESIGN.md 5.0.  Here the 32 rows the radix-32 pass produces go through passes 2 and 3 in two halves
// of 16 rows (they are independent 256-point transforms), so the work buffer is 33 KB, and the
// ring holds just the frame's own chunks (the slot of the oldest chunk is handed to the copy for
// the next frame as soon as pass 1 has read it; it lands under passes 2 and 3): 100 KB per CTA,
// two CTAs per SM, four warps per scheduler -- one CTA computes while the other waits at a barrier
// or for its chunk.  The second half's rows wait in registers; the window is re-read (L1/L2) per
// frame instead of living in 32 registers.
template <int C>     // C = N / hop: 1 or 2
__global__ void __launch_bounds__(256, 2) spectrum_kernel_v4(const SpecArgs3 g)
{
	extern __shared__ __align__(16) float2 wr_fft_smem[];
	const SpecArgs &a = g.a;
	constexpr int R1 = 32, RH = 16;
	constexpr unsigned N = R1 * 256, HOP = N / C, NSLOT = C;
	float2 *sm = wr_fft_smem;                               // work buffer, RH rows of kRowPitch
	float2 *tw2 = wr_fft_smem + RH * kRowPitch;             // pass-2 twiddles
	float2 *ring = tw2 + 256;                               // NSLOT chunks of HOP samples
	__shared__ __align__(8) unsigned long long wr_spec_bars4[NSLOT];
	const unsigned tid = threadIdx.x;
	const uint32_t bar32 = (uint32_t)__cvta_generic_to_shared(wr_spec_bars4);
	const uint32_t ring32 = (uint32_t)__cvta_generic_to_shared(ring);
	if (tid == 0) {
		for (unsigned i = 0; i < NSLOT; i++)
			spec_mbar_init(bar32 + 8u * i, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	tw2[tid] = __ldg(a.twiddle + ((((tid & 15) * (tid >> 4)) & 255) * R1));
	const float2 w1 = __ldg(a.twiddle + tid);               // W_N^tid: row k1 of pass 1 is scaled by w1^k1
	const float2 w1h = __ldg(a.twiddle + RH * tid);         // w1^16, where the second half begins
	__syncthreads();

	const unsigned nrows3 = a.nrows - g.row0;
	const unsigned totalRuns = g.runsPerStream * g.nStreams;
	unsigned phases = 0;                                    // bit s: parity of the next completion of slot s to wait for
	for (unsigned run = blockIdx.x; run < totalRuns; run += gridDim.x) {
		const unsigned t = run / g.runsPerStream;
		const unsigned r0 = (run - t * g.runsPerStream) * g.rowsPerRun;
		const unsigned nr = min(g.rowsPerRun, nrows3 - r0);
		const unsigned mFirst = g.row0 + r0;
		const float2 *src = a.in + (size_t)t * a.in_stride + ((size_t)mFirst * HOP - a.ncarry);
		const unsigned nchunks = nr + C - 1;
		if (tid == 0) {
			for (unsigned c = 0; c < NSLOT && c < nchunks; c++)
				spec_bulk_load(ring32 + c * HOP * 8u, src + (size_t)c * HOP, HOP * 8u, bar32 + 8u * c);
		}
		for (unsigned f = 0; f < nr; f++) {
			const unsigned m = mFirst + f;
			// ---- pass 1: window (fetched before the wait), the frame's chunks, radix-32 ----
			float2 v[R1];
			#pragma unroll
			for (unsigned c = 0; c < (unsigned)C; c++) {
				// the window of this chunk's points is fetched before the wait for the chunk (a chunk at a
				// time: all 32 window values at once cost a spill, 0.692 -> 0.678 ms on cfg4)
				float win[R1 / C];
				#pragma unroll
				for (int j = 0; j < R1 / C; j++)
					win[j] = __ldg(a.window + tid + 256u * (c * (R1 / C) + j));
				const unsigned slot = (f + c) % NSLOT;
				if (c == (unsigned)C - 1 || f == 0) {
					spec_mbar_wait(bar32 + 8u * slot, (phases >> slot) & 1u);
					phases ^= 1u << slot;
				}
				const float2 *chunk = ring + slot * HOP;
				#pragma unroll
				for (int j = 0; j < R1 / C; j++) {
					v[c * (R1 / C) + j] = chunk[tid + 256u * j];
					// inbuf[n] *= window[n] (spectrumsink.cxx:110-113): both components in one packed multiply
					asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&v[c * (R1 / C) + j]))
							: "l"(*reinterpret_cast<const unsigned long long*>(&v[c * (R1 / C) + j])), "l"(pack2(win[j], win[j])));
				}
			}
			wrfft::RegDft<R1>::run(v);
			// every thread is done with the previous frame's pass 3 (work buffer) and with this
			// frame's oldest chunk (ring)
			__syncthreads();
			if (tid == 0 && f + NSLOT < nchunks) {
				const unsigned c = f + NSLOT;               // the chunk that completes frame f + 1
				spec_bulk_load(ring32 + (c % NSLOT) * HOP * 8u, src + (size_t)c * HOP, HOP * 8u, bar32 + 8u * (c % NSLOT));
			}
			float *rowout = a.rows + (size_t)t * a.row_stride + (size_t)m * N;
			float *last = (m == a.nrows - 1) ? a.last + (size_t)t * N : nullptr;
			#pragma unroll
			for (int h = 0; h < 2; h++) {
				if (h)
					__syncthreads();                        // the first half's pass 3 has read the work buffer
				{
					float2 cur = h ? w1h : w1;
					#pragma unroll
					for (int l = 0; l < RH; l++) {
						const int k1 = h * RH + l;
						if (k1 == 0) {
							sm[tid] = v[0];
						} else {
							sm[l * kRowPitch + tid] = wrfft::cmul(v[k1], cur);
							cur = wrfft::cmul(cur, w1);
						}
					}
				}
				__syncthreads();
				// ---- pass 2: radix-16 over the stride-16 index of every 256-point row (one butterfly per thread) ----
				{
					float2 u[16];
					const unsigned b = tid & 15, l = tid >> 4;
					float2 *row = sm + l * kRowPitch + b;
					#pragma unroll
					for (int q = 0; q < 16; q++)
						u[q] = row[16 * q];
					wrfft::RegDft<16>::run(u);
					#pragma unroll
					for (int ka = 0; ka < 16; ka++)
						row[16 * ka] = ka ? wrfft::cmul(u[ka], tw2[16 * ka + b]) : u[0];
				}
				__syncthreads();
				// ---- pass 3: radix-16 over the contiguous index; dB + fft-shift fused into the store ----
				{
					float2 u[16];
					const unsigned l = tid & (RH - 1), b2 = tid >> 4;
					const float2 *row = sm + l * kRowPitch + 16 * b2;
					#pragma unroll
					for (int q = 0; q < 16; q++)
						u[q] = row[q];
					wrfft::RegDft<16>::run(u);
					// bin k = k1 + 32 (b2 + 16 kb), k1 = 16 h + l; fft-shift: o = (k + N/2) mod N -- k < 512 + 512 kb,
					// so the wrap is a compile-time matter of kb
					const unsigned k0 = (unsigned)(h * RH) + l + R1 * b2;
					float dbv[16];
					#pragma unroll
					for (int kb = 0; kb < 16; kb++) {
						const float p = fmaf(u[kb].x, u[kb].x, u[kb].y * u[kb].y);
						float l2;
						asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(p));
						dbv[kb] = fmaf(3.01029995663981195f, l2, -a.scaledb);
					}
					#pragma unroll
					for (int kb = 0; kb < 16; kb++)
						rowout[k0 + (kb < 8 ? N / 2 + 512 * kb : 512 * (kb - 8))] = dbv[kb];
					if (last) {
						#pragma unroll
						for (int kb = 0; kb < 16; kb++)
							last[k0 + (kb < 8 ? N / 2 + 512 * kb : 512 * (kb - 8))] = dbv[kb];
					}
				}
			}
		}
		// the next run reuses ring and work buffer
		__syncthreads();
	}
}

// the browser's palette index for one row (wr_device.cuh: waterfall_index)
__global__ void spectrum_palette_kernel(const float *__restrict__ db, unsigned char *__restrict__ out, unsigned n)
{
	for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
		out[k] = wrd::waterfall_index(db[k]);
}

// leftover frames of [carry | in] starting at `from` become the next call's carry
__global__ void spectrum_carry_kernel(const SpecArgs a, float2 *next, size_t from, unsigned count)
{
	const unsigned t = blockIdx.y;
	for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
		next[(size_t)t * a.carry_stride + i] = view(a, t, from + i);
}

} // namespace

struct wr_spectrum {
	int device = 0;
	unsigned N = 0, logN = 0, hop = 0, T = 0, maxF = 0;
	cudaStream_t st = nullptr;
	float *d_window = nullptr;
	float2 *d_twiddle = nullptr;
	float2 *d_carry[2] = { nullptr, nullptr };
	int cur = 0;
	unsigned ncarry = 0;
	float *d_in = nullptr;     // host-path staging [T][maxF][2]
	float *d_rows = nullptr;   // host-path staging [T][maxRows][N]
	unsigned maxRows = 0;
	float *d_last = nullptr;   // [T][N]
	unsigned char *d_palette = nullptr;   // [N] scratch of wr_spectrum_get_palette
	bool haveLast = false;
	cudaStream_t lastStream = nullptr; // stream of the most recent launch
	// host path of large multi-stream blocks: groups of streams go through copy-in / transforms / copy-out pipelined
	static constexpr int kGroups = 8;
	cudaStream_t h2d = nullptr, d2h = nullptr;
	cudaEvent_t inReady[kGroups] = {}, rowsReady[kGroups] = {}, outDone = nullptr;
	bool noPipe = false;               // env WR_FFT_PIPE=0: one copy in, the kernels, one copy out
	bool forceV1 = false;              // env WR_FFT_V1=1: radix-4 shared-memory kernel for every size
	bool noV3 = false;                 // env WR_FFT_V3=0: never the persistent bulk-copy kernel
	bool noV4 = false;                 // env WR_FFT_V4=0: 8192-point transforms stay with v3 (one CTA per SM)
	unsigned runsPerCta = 0;           // env WR_FFT_RUNS: runs of consecutive rows per CTA of the persistent grid (0 = 7)
	int numSMs = 148;
	unsigned long long launches = 0;
};

namespace {

void free_spectrum(wr_spectrum *s)
{
	if (!s)
		return;
	cudaSetDevice(s->device);
	if (s->st)
		cudaStreamSynchronize(s->st);
	cudaFree(s->d_window);
	cudaFree(s->d_twiddle);
	cudaFree(s->d_carry[0]);
	cudaFree(s->d_carry[1]);
	cudaFree(s->d_in);
	cudaFree(s->d_rows);
	cudaFree(s->d_last);
	cudaFree(s->d_palette);
	if (s->st)
		cudaStreamDestroy(s->st);
	if (s->h2d)
		cudaStreamDestroy(s->h2d);
	if (s->d2h)
		cudaStreamDestroy(s->d2h);
	for (int i = 0; i < wr_spectrum::kGroups; i++) {
		if (s->inReady[i]) cudaEventDestroy(s->inReady[i]);
		if (s->rowsReady[i]) cudaEventDestroy(s->rowsReady[i]);
	}
	if (s->outDone)
		cudaEventDestroy(s->outDone);
	cudaGetLastError();
	delete s;
}

// The kernels of one call for streams [t0, t0 + nT): iq_dev and rows_dev point at stream t0.  The
// handle's state (carry side, carried frames) is advanced by the caller once every stream is through.
long run_part(wr_spectrum *s, unsigned t0, unsigned nT, const float *iq_dev, size_t in_stride, unsigned nframes,
		float *rows_dev, size_t row_stride, cudaStream_t st)
{
	const size_t avail = (size_t)s->ncarry + nframes;
	const unsigned nrows = avail >= s->N ? (unsigned)((avail - s->N) / s->hop + 1) : 0;
	SpecArgs a;
	a.carry = s->d_carry[s->cur] + (size_t)t0 * s->N;
	a.carry_stride = s->N;
	a.ncarry = s->ncarry;
	a.in = reinterpret_cast<const float2*>(iq_dev);
	a.in_stride = in_stride;
	a.window = s->d_window;
	a.twiddle = s->d_twiddle;
	a.rows = rows_dev;
	a.row_stride = row_stride;
	a.last = s->d_last + (size_t)t0 * s->N;
	a.N = s->N;
	a.logN = s->logN;
	a.hop = s->hop;
	a.nrows = nrows;
	// scaledb = 20 * log10f((float)N) with the host libm, as the reference computes it
	a.scaledb = 20 * log10f((float)s->N);
	if (nrows) {
		// without a row buffer only the newest transform is observable (getSpectrum), so only
		// that one is computed; the reference transforms every frame and discards all but the last
		a.first_row = rows_dev ? 0 : nrows - 1;
		unsigned nrows2 = nrows - a.first_row;      // rows the one-frame-per-CTA kernels take
		// Many rows per stream (a waterfall): the persistent kernel that streams frames through
		// shared memory by bulk copies takes every row that lies entirely in this call's block; rows
		// that begin in the carry buffer stay with the one-frame-per-CTA kernel.
		const unsigned row0 = (s->ncarry + s->hop - 1) / s->hop;
		const bool aligned = !((uintptr_t)iq_dev & 15u) && !(in_stride & 1u) && !(s->ncarry & 1u);
		if (rows_dev && !s->forceV1 && !s->noV3 && s->N >= 512 && (s->hop == s->N || 2 * s->hop == s->N)
				&& aligned && nrows > row0 + 1) {
			SpecArgs3 g;
			g.a = a;
			g.row0 = row0;
			const unsigned nrows3 = nrows - row0;
			const unsigned long long total = (unsigned long long)nrows3 * nT;
			const unsigned C = s->N / s->hop, R1 = s->N / 256;
			// 8192-point transforms: the two-CTAs-per-SM kernel (half the work buffer, the frame's own chunks only)
			const bool v4 = R1 == 32 && !s->noV4;
			// about seven runs per CTA of the persistent grid
			const unsigned long long ctas = (unsigned long long)s->numSMs * (v4 ? 2 : 1);
			const unsigned long long per = s->runsPerCta ? s->runsPerCta : 7ull;
			const unsigned want = (unsigned)std::max<unsigned long long>(4, (total + per * ctas - 1) / (per * ctas));
			g.rowsPerRun = std::min(want, nrows3);
			g.runsPerStream = (nrows3 + g.rowsPerRun - 1) / g.rowsPerRun;
			const size_t smem = v4 ? sizeof(float2) * ((size_t)16 * kRowPitch + 256 + (size_t)C * s->hop)
					: sizeof(float2) * ((size_t)R1 * kRowPitch + 256 + (size_t)(C + 1) * s->hop);
			const unsigned long long runs = (unsigned long long)g.runsPerStream * nT;
			g.nStreams = nT;
			// persistent: one CTA per SM (two when two fit), each walking the run list with stride gridDim.x
			const dim3 grid3((unsigned)std::min<unsigned long long>(runs, (unsigned long long)s->numSMs * (smem <= 100 * 1024 ? 2 : 1)));
			void (*k3)(const SpecArgs3) = nullptr;
			switch (v4 ? 1000 + C : R1 * 10 + C) {
			case 1001: k3 = spectrum_kernel_v4<1>; break;
			case 1002: k3 = spectrum_kernel_v4<2>; break;
			case 21: k3 = spectrum_kernel_v3<2, 1>; break;
			case 22: k3 = spectrum_kernel_v3<2, 2>; break;
			case 41: k3 = spectrum_kernel_v3<4, 1>; break;
			case 42: k3 = spectrum_kernel_v3<4, 2>; break;
			case 81: k3 = spectrum_kernel_v3<8, 1>; break;
			case 82: k3 = spectrum_kernel_v3<8, 2>; break;
			case 161: k3 = spectrum_kernel_v3<16, 1>; break;
			case 162: k3 = spectrum_kernel_v3<16, 2>; break;
			case 321: k3 = spectrum_kernel_v3<32, 1>; break;
			default: k3 = spectrum_kernel_v3<32, 2>; break;
			}
			WR_CUDA(cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
			k3<<<grid3, 256, smem, st>>>(g);
			s->launches++;
			WR_CUDA(cudaGetLastError());
			nrows2 = row0;                          // what is left for the kernels below
		}
		dim3 grid(nrows2, nT);
		if (nrows2 == 0) {
			// nothing left
		} else if (s->N >= 512 && !s->forceV1) {
			const size_t smem = sizeof(float2) * ((size_t)(s->N / 256) * kRowPitch + 256);
			switch (s->N / 256) {
			case 2: spectrum_kernel_v2<2><<<grid, 256, smem, st>>>(a); break;
			case 4: spectrum_kernel_v2<4><<<grid, 256, smem, st>>>(a); break;
			case 8: spectrum_kernel_v2<8><<<grid, 256, smem, st>>>(a); break;
			case 16: spectrum_kernel_v2<16><<<grid, 256, smem, st>>>(a); break;
			default: spectrum_kernel_v2<32><<<grid, 256, smem, st>>>(a); break;
			}
		} else {
			unsigned threads = std::min(1024u, std::max(32u, s->N / 4));
			size_t smem = sizeof(float2) * 2 * (size_t)s->N;
			spectrum_kernel_v1<<<grid, threads, smem, st>>>(a);
		}
		if (nrows2) {
			s->launches++;
			WR_CUDA(cudaGetLastError());
		}
		s->haveLast = true;
		s->lastStream = st;
	}
	const size_t consumed = (size_t)nrows * s->hop;
	const unsigned left = (unsigned)(avail - consumed);
	if (left) {
		dim3 grid(std::max(1u, std::min(8u, (left + 255) / 256)), nT);
		spectrum_carry_kernel<<<grid, 256, 0, st>>>(a, s->d_carry[s->cur ^ 1] + (size_t)t0 * s->N, consumed, left);
		s->launches++;
		WR_CUDA(cudaGetLastError());
	}
	return (long)nrows;
}

// every stream went through run_part: the carried frames are on the other side now
void advance(wr_spectrum *s, unsigned nframes, long nrows)
{
	const size_t avail = (size_t)s->ncarry + nframes;
	s->cur ^= 1;
	s->ncarry = (unsigned)(avail - (size_t)nrows * s->hop);
}

long run(wr_spectrum *s, const float *iq_dev, size_t in_stride, unsigned nframes,
		float *rows_dev, size_t row_stride, cudaStream_t st)
{
	const long nrows = run_part(s, 0, s->T, iq_dev, in_stride, nframes, rows_dev, row_stride, st);
	if (nrows >= 0)
		advance(s, nframes, nrows);
	return nrows;
}

} // namespace

extern "C" {

wr_spectrum *wr_spectrum_create(int device, unsigned fft_size, unsigned hop, unsigned n_streams, unsigned max_frames)
{
	// power of two only, as SpectrumSink::setFftSize enforces (spectrumsink.cxx:53-56)
	if (fft_size < 8 || (fft_size & (fft_size - 1)) || fft_size > 8192 || hop == 0 || hop > fft_size
			|| n_streams == 0 || n_streams > 65535 || max_frames == 0) {
		wr::set_error("wr_spectrum_create: fft_size must be a power of two in [8, 8192], 0 < hop <= fft_size "
				"(got N=%u hop=%u streams=%u)", fft_size, hop, n_streams);
		return nullptr;
	}
	if (!wr::check_device(device))
		return nullptr;
	wr_spectrum *s = new wr_spectrum();
	s->device = device;
	s->N = fft_size;
	s->hop = hop;
	s->T = n_streams;
	s->maxF = max_frames;
	while ((1u << s->logN) < fft_size)
		s->logN++;
	s->maxRows = (unsigned)(((size_t)fft_size - 1 + max_frames - fft_size) / hop + 1);

	std::vector<float> win(fft_size);
	for (unsigned n = 0; n < fft_size; n++) // spectrumsink.cxx:71-74
		win[n] = (float)(0.54 - 0.46 * cosf((float)(2 * M_PI * (float)n / (float)(fft_size - 1))));
	std::vector<float2> tw(fft_size);
	for (unsigned k = 0; k < fft_size; k++) {
		double ang = -2.0 * M_PI * (double)k / (double)fft_size;
		tw[k] = make_float2((float)cos(ang), (float)sin(ang));
	}
#define WR_SPEC_ALLOC(expr)                                               \
	do {                                                                  \
		cudaError_t e_ = (expr);                                          \
		if (e_ != cudaSuccess) {                                          \
			wr::set_error("%s: %s", #expr, cudaGetErrorString(e_));       \
			free_spectrum(s);                                             \
			return nullptr;                                               \
		}                                                                 \
	} while (0)
	WR_SPEC_ALLOC(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking));
	WR_SPEC_ALLOC(cudaMalloc(&s->d_window, sizeof(float) * fft_size));
	WR_SPEC_ALLOC(cudaMemcpy(s->d_window, win.data(), sizeof(float) * fft_size, cudaMemcpyHostToDevice));
	WR_SPEC_ALLOC(cudaMalloc(&s->d_twiddle, sizeof(float2) * fft_size));
	WR_SPEC_ALLOC(cudaMemcpy(s->d_twiddle, tw.data(), sizeof(float2) * fft_size, cudaMemcpyHostToDevice));
	for (int i = 0; i < 2; i++) {
		WR_SPEC_ALLOC(cudaMalloc(&s->d_carry[i], sizeof(float2) * (size_t)n_streams * fft_size));
		WR_SPEC_ALLOC(cudaMemset(s->d_carry[i], 0, sizeof(float2) * (size_t)n_streams * fft_size));
	}
	WR_SPEC_ALLOC(cudaMalloc(&s->d_last, sizeof(float) * (size_t)n_streams * fft_size));
	WR_SPEC_ALLOC(cudaMemset(s->d_last, 0, sizeof(float) * (size_t)n_streams * fft_size));
	WR_SPEC_ALLOC(cudaFuncSetAttribute(spectrum_kernel_v1, cudaFuncAttributeMaxDynamicSharedMemorySize,
			(int)(sizeof(float2) * 2 * 8192)));
	WR_SPEC_ALLOC(cudaFuncSetAttribute(spectrum_kernel_v2<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
			(int)(sizeof(float2) * (16 * kRowPitch + 256))));
	WR_SPEC_ALLOC(cudaFuncSetAttribute(spectrum_kernel_v2<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
			(int)(sizeof(float2) * (32 * kRowPitch + 256))));
	if (const char *e = getenv("WR_FFT_V1"))
		s->forceV1 = atoi(e) != 0;
	if (const char *e = getenv("WR_FFT_V3"))
		s->noV3 = atoi(e) == 0;
	if (const char *e = getenv("WR_FFT_V4"))
		s->noV4 = atoi(e) == 0;
	if (const char *e = getenv("WR_FFT_PIPE"))
		s->noPipe = atoi(e) == 0;
	if (const char *e = getenv("WR_FFT_RUNS"))
		s->runsPerCta = (unsigned)std::max(0, atoi(e));
	WR_SPEC_ALLOC(cudaDeviceGetAttribute(&s->numSMs, cudaDevAttrMultiProcessorCount, device));
#undef WR_SPEC_ALLOC
	return s;
}

void wr_spectrum_destroy(wr_spectrum *s) { free_spectrum(s); }

long wr_spectrum_process_device(wr_spectrum *s, const float *iq_dev, size_t stride_frames, unsigned nframes,
		float *rows_dev, size_t row_stride, void *cuda_stream)
{
	WR_REQUIRE(s && (iq_dev || nframes == 0), WR_EINVAL, "wr_spectrum_process_device: null argument");
	WR_REQUIRE(nframes <= s->maxF, WR_EINVAL, "wr_spectrum_process_device: %u frames > max_frames %u", nframes, s->maxF);
	if (!wr::use_device(s->device))
		return WR_ENODEV;
	return run(s, iq_dev, stride_frames, nframes, rows_dev, row_stride, cuda_stream ? (cudaStream_t)cuda_stream : s->st);
}

long wr_spectrum_process(wr_spectrum *s, const float *iq_host, unsigned nframes, float *rows_host, size_t row_stride)
{
	WR_REQUIRE(s && (iq_host || nframes == 0), WR_EINVAL, "wr_spectrum_process: null argument");
	WR_REQUIRE(nframes <= s->maxF, WR_EINVAL, "wr_spectrum_process: %u frames > max_frames %u", nframes, s->maxF);
	if (!wr::use_device(s->device))
		return WR_ENODEV;
	if (!s->d_in)
		WR_CUDA(cudaMalloc(&s->d_in, sizeof(float) * 2 * (size_t)s->T * s->maxF));
	if (rows_host && !s->d_rows)
		WR_CUDA(cudaMalloc(&s->d_rows, sizeof(float) * (size_t)s->T * s->maxRows * s->N));
	// Large multi-stream blocks with a row buffer (a waterfall of many tuners: BASELINE config 4 moves 1 GiB in
	// and 1 GiB out per call): the streams go through in groups, so that a group's copy-in runs under the
	// transforms of the group before it and under the copy-out of the one before that (the link is full duplex).
	if (rows_host && nframes && !s->noPipe && s->T >= 2 * wr_spectrum::kGroups
			&& sizeof(float) * 2 * (size_t)nframes * s->T >= ((size_t)32 << 20)) {
		constexpr int G = wr_spectrum::kGroups;
		if (!s->h2d) {
			WR_CUDA(cudaStreamCreateWithFlags(&s->h2d, cudaStreamNonBlocking));
			WR_CUDA(cudaStreamCreateWithFlags(&s->d2h, cudaStreamNonBlocking));
			for (int i = 0; i < G; i++) {
				WR_CUDA(cudaEventCreateWithFlags(&s->inReady[i], cudaEventDisableTiming));
				WR_CUDA(cudaEventCreateWithFlags(&s->rowsReady[i], cudaEventDisableTiming));
			}
			WR_CUDA(cudaEventCreateWithFlags(&s->outDone, cudaEventDisableTiming));
		}
		// (a failure half-way must not leave copies in flight on the caller's buffers)
		auto groups = [&]() -> long {
			long nrows = 0;
			for (int gi = 0; gi < G; gi++) {
				const unsigned t0 = (unsigned)((unsigned long long)s->T * gi / G), t1 = (unsigned)((unsigned long long)s->T * (gi + 1) / G);
				const unsigned nT = t1 - t0;
				float *din = s->d_in + (size_t)t0 * 2 * s->maxF;
				float *drows = s->d_rows + (size_t)t0 * s->maxRows * s->N;
				WR_CUDA(cudaMemcpy2DAsync(din, sizeof(float) * 2 * (size_t)s->maxF, iq_host + (size_t)t0 * 2 * nframes,
						sizeof(float) * 2 * (size_t)nframes, sizeof(float) * 2 * (size_t)nframes, nT, cudaMemcpyHostToDevice, s->h2d));
				WR_CUDA(cudaEventRecord(s->inReady[gi], s->h2d));
				WR_CUDA(cudaStreamWaitEvent(s->st, s->inReady[gi], 0));
				nrows = run_part(s, t0, nT, din, s->maxF, nframes, drows, (size_t)s->maxRows * s->N, s->st);
				if (nrows < 0)
					return nrows;
				WR_CUDA(cudaEventRecord(s->rowsReady[gi], s->st));
				if (nrows > 0) {
					WR_CUDA(cudaStreamWaitEvent(s->d2h, s->rowsReady[gi], 0));
					WR_CUDA(cudaMemcpy2DAsync(rows_host + (size_t)t0 * row_stride, sizeof(float) * row_stride, drows,
							sizeof(float) * (size_t)s->maxRows * s->N, sizeof(float) * (size_t)nrows * s->N, nT, cudaMemcpyDeviceToHost, s->d2h));
				}
			}
			return nrows;
		};
		const long nrows = groups();
		if (nrows < 0) {
			cudaStreamSynchronize(s->h2d);
			cudaStreamSynchronize(s->st);
			cudaStreamSynchronize(s->d2h);
			return nrows;
		}
		advance(s, nframes, nrows);
		WR_CUDA(cudaStreamSynchronize(s->st));
		WR_CUDA(cudaStreamSynchronize(s->d2h));
		return nrows;
	}
	if (nframes)
		WR_CUDA(cudaMemcpy2DAsync(s->d_in, sizeof(float) * 2 * (size_t)s->maxF, iq_host, sizeof(float) * 2 * (size_t)nframes,
				sizeof(float) * 2 * (size_t)nframes, s->T, cudaMemcpyHostToDevice, s->st));
	long nrows = run(s, s->d_in, s->maxF, nframes, rows_host ? s->d_rows : nullptr, (size_t)s->maxRows * s->N, s->st);
	if (nrows < 0)
		return nrows;
	if (rows_host && nrows > 0)
		WR_CUDA(cudaMemcpy2DAsync(rows_host, sizeof(float) * row_stride, s->d_rows, sizeof(float) * (size_t)s->maxRows * s->N,
				sizeof(float) * (size_t)nrows * s->N, s->T, cudaMemcpyDeviceToHost, s->st));
	WR_CUDA(cudaStreamSynchronize(s->st));
	return nrows;
}

long wr_spectrum_process_upload(wr_spectrum *s, wr_upload *u, unsigned nframes)
{
	WR_REQUIRE(s && u, WR_EINVAL, "wr_spectrum_process_upload: null argument");
	WR_REQUIRE(s->T == 1, WR_EINVAL, "wr_spectrum_process_upload: the sink has %u streams, an upload carries one", s->T);
	WR_REQUIRE(u->device == s->device, WR_EINVAL, "wr_spectrum_process_upload: upload on device %d, sink on %d", u->device, s->device);
	WR_REQUIRE(nframes == u->nframes && nframes <= s->maxF, WR_EINVAL, "wr_spectrum_process_upload: %u frames asked, %u uploaded, sink holds %u",
			nframes, u->nframes, s->maxF);
	if (!wr::use_device(s->device))
		return WR_ENODEV;
	// Asynchronous: the transform of the newest frame runs behind the upload's last piece on the
	// sink's own stream; wr_spectrum_get synchronises.  The upload is told, so that it does not
	// overwrite this side before the kernels have read it.
	if (nframes)
		WR_CUDA(cudaStreamWaitEvent(s->st, u->ready(nframes), 0));
	long nrows = run(s, u->dev(), u->maxFrames, nframes, nullptr, 0, s->st);
	if (nrows < 0)
		return nrows;
	WR_CUDA(cudaEventRecord(u->readDone[u->cur], s->st));
	u->readPending[u->cur] = true;
	return nrows;
}

int wr_spectrum_reserve(wr_spectrum *s, unsigned max_frames)
{
	WR_REQUIRE(s && max_frames > 0, WR_EINVAL, "wr_spectrum_reserve: bad argument");
	if (max_frames <= s->maxF)
		return WR_OK;
	if (!wr::use_device(s->device))
		return WR_ENODEV;
	// only the staging buffers of the host path depend on the block length; the carried partial
	// frame (SpectrumSink's inoffset, reference spectrumsink.h:62) and the last row stay
	WR_CUDA(cudaStreamSynchronize(s->st));
	cudaFree(s->d_in);
	cudaFree(s->d_rows);
	s->d_in = nullptr;
	s->d_rows = nullptr;
	s->maxF = max_frames;
	s->maxRows = (unsigned)(((size_t)s->N - 1 + max_frames - s->N) / s->hop + 1);
	return WR_OK;
}

int wr_spectrum_get(wr_spectrum *s, unsigned stream, float *db_host)
{
	WR_REQUIRE(s && db_host && stream < s->T, WR_EINVAL, "wr_spectrum_get: bad argument");
	if (!wr::use_device(s->device))
		return WR_ENODEV;
	WR_CUDA(cudaStreamSynchronize(s->st));
	if (s->lastStream && s->lastStream != s->st)
		WR_CUDA(cudaStreamSynchronize(s->lastStream));
	if (!s->haveLast) {
		// no transform yet: the reference's outbuf is still zero -> 10*log10f(0) - scaledb = -inf
		float scaledb = 20 * log10f((float)s->N);
		for (unsigned n = 0; n < s->N; n++)
			db_host[n] = 10 * log10f(0.0f) - scaledb;
		return WR_OK;
	}
	WR_CUDA(cudaMemcpy(db_host, s->d_last + (size_t)stream * s->N, sizeof(float) * s->N, cudaMemcpyDeviceToHost));
	return WR_OK;
}

int wr_spectrum_get_palette(wr_spectrum *s, unsigned stream, uint8_t *index_host)
{
	WR_REQUIRE(s && index_host && stream < s->T, WR_EINVAL, "wr_spectrum_get_palette: bad argument");
	if (!wr::use_device(s->device))
		return WR_ENODEV;
	WR_CUDA(cudaStreamSynchronize(s->st));
	if (s->lastStream && s->lastStream != s->st)
		WR_CUDA(cudaStreamSynchronize(s->lastStream));
	if (!s->haveLast) {
		// no transform yet: every bin is -inf, which the handler sends as -10000.0 -> index 0
		memset(index_host, 0, s->N);
		return WR_OK;
	}
	if (!s->d_palette)
		WR_CUDA(cudaMalloc(&s->d_palette, s->N));
	spectrum_palette_kernel<<<(s->N + 255) / 256, 256, 0, s->st>>>(s->d_last + (size_t)stream * s->N, s->d_palette, s->N);
	s->launches++;
	WR_CUDA(cudaGetLastError());
	WR_CUDA(cudaMemcpyAsync(index_host, s->d_palette, s->N, cudaMemcpyDeviceToHost, s->st));
	WR_CUDA(cudaStreamSynchronize(s->st));
	return WR_OK;
}

unsigned long long wr_spectrum_launch_count(const wr_spectrum *s) { return s ? s->launches : 0; }

int wr_spectrum_sync(wr_spectrum *s)
{
	WR_REQUIRE(s, WR_EINVAL, "wr_spectrum_sync: null handle");
	if (!wr::use_device(s->device))
		return WR_ENODEV;
	WR_CUDA(cudaStreamSynchronize(s->st));
	return WR_OK;
}

} // extern "C"
