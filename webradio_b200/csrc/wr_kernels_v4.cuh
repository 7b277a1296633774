// wr_kernels_v4.cuh -- fused NCO mix + decimating channel FIR (K1+K2 of SURVEY.md 2a), fourth
// generation: a STREAMING FIR for banks of independent tuner streams (BASELINE cfg3).
//
// What v3 measured on cfg3 (profiles/r01_ncu_full_chan_cfg3.txt): L1/shared-memory throughput 85 %,
// issue slots 54 %.  Its FIR warps compute one output per thread by GATHERING the output's N1 mixed
// samples from a shared-memory ring -- neighbouring outputs overlap in all but D1 of them, so every
// mixed sample is read N1/D1 (5.1) times, 49 of the kernel's 58 M shared-memory wavefronts per
// launch -- and its mixer warps wait for the raw loads (long-scoreboard stalls, 2.2 per issue).
//
// Shape of v4.  No warp specialisation and no mixed ring: ONE THREAD OWNS A CONTIGUOUS RUN OF K
// OUTPUTS of one receiver and streams over that run's frames in time order.  It mixes a frame and
// applies it at once to the A'+1 outputs whose windows cover it, each output in its own register
// accumulator; for a fixed output the frames arrive in increasing tap index, so every accumulator
// sees exactly the reference's sequence of separately rounded products and sums
// (lowpass.cxx:151-159).  A mixed sample is produced once, used from registers and never stored.
//   * A warp = 32 consecutive runs of one receiver (lane L owns outputs [L*K, (L+1)*K)); a
//     receiver takes G warps, K = ceil(M1 / (32 G)).  Warps are independent of each other: no
//     block-wide barrier after start-up, no slot hand-over, no spinning.
//   * Raw IQ: each warp keeps an NS-deep ring of stages in shared memory, a stage = SFR frames of
//     each of the warp's 32 runs (32 rows of 80 bytes).  The rows are 25.6 KB apart in HBM, so
//     the stage is fetched by 16-byte cp.async copies, 8-byte granules apart within a row being
//     contiguous: 5 lanes cover a row's 80 bytes, a warp-instruction touches 7 rows.  The copies
//     are issued NS-1 stages ahead and never block.  A lane then reads ITS row with five 128-bit
//     loads; 80 bytes is an odd number of 16-byte chunks, so 8 consecutive rows cover all banks.
//   * NCO: the sine and cosine entries of wr_lo3.h, five frames at a time in packed f32x2 (the
//     v3 code); the phase of a frame is closed-form, so runs need no hand-over either.
//   * Taps: the receiver's taps are staged once per warp-unit in shared memory as (A'+1) rows of
//     D1, row a holding taps a*D1 ... a*D1+D1-1, i.e. what the output of AGE a (periods since its
//     window began) needs at each position of a period; all lanes of a warp are at the same
//     position, so a tap load is a broadcast.  In one period every tap is used exactly once.
//   * What the runs overlap: a run's last N1-1 frames are also the first frames of the next run
//     (its outputs' windows begin there).  They are mixed twice -- (N1-1)/(K*D1) = 6.6 % at
//     cfg3's K = 64 -- which buys independence; keeping them instead would take 65 KB per
//     receiver in flight (32 x 254 frames).
//   * The first outputs of a block, whose windows reach into the carried history, are computed by
//     a short gather prologue from [history | first frames]; the carried history of the NEXT
//     block (the last N1-1 mixed frames) is produced by an epilogue.  The steady state therefore
//     has no special cases: frames outside the block read as zero (cp.async src-size).
// Geometry: N1 odd (so that a run starts on a 16-byte boundary), D1 even and a multiple of the
// stage length.  Shared-tuner banks (cfg2, cfg5) stay with v3, whose mixers share the raw
// registers among the receivers of a stream.
#pragma once

#include "wr_kernels_v3.cuh"

namespace wrd {

constexpr int kV4MaxWarps = 16;

template <int N1, int D1>
struct V4Geo {
	static constexpr int SFR = 10;                          // frames per run and stage
	static constexpr int S = D1 / SFR;                      // stages per period
	static constexpr int AP = (N1 - 1) / D1;                // full periods a window spans beyond its first
	static constexpr int REM = (N1 - 1) % D1;               // last tap's position in the window's last period
	static constexpr int ROWS = AP + 1;                     // tap rows / live accumulators
	static constexpr int TS = (SFR + 3) & ~3;               // a stage's taps of one row, padded to 128-bit loads
	static constexpr int SFIN = REM / SFR;                  // stage in which the oldest output completes
	static constexpr int CFIN = REM % SFR + 1;              // ... after this many of its frames
	static constexpr int KSKIP = (N1 - 1 + D1 - 1) / D1;    // outputs whose windows reach into the history
	static constexpr unsigned kRowBytes = SFR * 8;          // a run's frames of one stage
	static constexpr unsigned kStageBytes = 32 * kRowBytes;
	static constexpr unsigned kTapBytes = (unsigned)ROWS * S * TS * 4;
	static constexpr unsigned kScratchFrames = (unsigned)(KSKIP - 1) * D1 + N1;   // [history | first frames] of the prologue
	static_assert(D1 % SFR == 0, "a period must be a whole number of stages");
	static_assert(N1 % 2 == 1 && D1 % 2 == 0, "v4 needs an odd tap count and an even decimation (16-byte aligned runs)");
	static_assert((kRowBytes / 16) % 2 == 1, "a row must be an odd number of 16-byte chunks (bank spread)");
	static_assert(AP >= 1, "the window must span more than one period");
};

struct V4Args {
	const int16_t *delta;     // padded corrections (wr_lo3.h), staged to shared memory per CTA
	float eps;
	float negzero;            // -0.0f, opaque to the compiler (mul2_rn_exact)
	unsigned prmtHi;          // 0x4B00, opaque to the compiler
	unsigned G;               // warps per receiver
	unsigned K;               // outputs per run
	unsigned NS;              // stages in a warp's ring
	unsigned nUnits;          // R * G
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}

__device__ __forceinline__ void cp_async16z(uint32_t dst, const void *src, unsigned srcBytes)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(srcBytes) : "memory");
}

__device__ __forceinline__ void cp_async_commit()
{
	asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int N>
__device__ __forceinline__ void cp_async_wait()
{
	asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory");
}

// One stage of a run: SFR frames mixed and applied to the live outputs.  CX = how many of the
// stage's leading frames still carry the OLDEST output (tap row AP); FIN = that output completes
// in this stage (after frame CX-1), and `fin` receives it.
template <int N1, int D1, int CX, bool FIN>
__device__ __forceinline__ void v4_stage(const f2_t (&raw)[V4Geo<N1, D1>::SFR], f2_t (&acc)[V4Geo<N1, D1>::ROWS],
		uint32_t q0, uint32_t qs, const Lo3Regs &lo, uint32_t tap32, f2_t nz, f2_t &fin)
{
	using G = V4Geo<N1, D1>;
	constexpr int SFR = G::SFR, AP = G::AP, TS = G::TS, S = G::S;
	constexpr int H = SFR / 2;
	#pragma unroll
	for (int h = 0; h < 2; h++) {
		uint32_t q[H];
		float sn[H], cs[H];
		#pragma unroll
		for (int i = 0; i < H; i++)
			q[i] = q0 + (uint32_t)(h * H + i) * qs;
		lo3_sincos_n<H>(q, lo, sn, cs);
		#pragma unroll
		for (int i = 0; i < H; i++) {
			const int fi = h * H + i;
			// downconverter.cxx:109-110:  I' = i*cos + q*sin ;  Q' = q*cos - i*sin
			float ic, qc, is, qq;
			f2_unpack(f2_fma(raw[fi], f2_pack(cs[i], cs[i]), nz), ic, qc);
			f2_unpack(f2_fma(raw[fi], f2_pack(sn[i], sn[i]), nz), is, qq);
			const f2_t m = f2_pack(__fadd_rn(ic, qq), __fsub_rn(qc, is));
			// lowpass.cxx:155-156 for every output whose window covers this frame: the output of
			// age a takes tap a*D1 + (position in the period)
			#pragma unroll
			for (int a = 0; a < AP; a++)
				acc[a] = tap3(acc[a], lds32(tap32 + 4u * (unsigned)(a * S * TS + fi)), m, nz);
			if (fi < CX) {
				acc[AP] = tap3(acc[AP], lds32(tap32 + 4u * (unsigned)(AP * S * TS + fi)), m, nz);
				if (FIN && fi == CX - 1)
					fin = acc[AP];
			}
		}
	}
}

template <int N1, int D1>
__global__ void __launch_bounds__(kV4MaxWarps * 32, 1) chan_kernel_v4(const ChanArgs a, const V4Args v)
{
	using G = V4Geo<N1, D1>;
	constexpr int SFR = G::SFR, S = G::S, AP = G::AP, ROWS = G::ROWS, TS = G::TS;
	extern __shared__ __align__(16) unsigned char wr_smem_v4[];
	const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nWarps = blockDim.x >> 5;
	// let the demodulator kernel behind this one be scheduled as SMs drain (it waits for this grid)
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

	// the NCO correction table, once per CTA, by bulk copies that complete on an mbarrier
	__shared__ __align__(8) unsigned long long wr_bar_v4;
	const uint32_t bar32 = (uint32_t)__cvta_generic_to_shared(&wr_bar_v4);
	if (tid == 0) {
		mbar_init(bar32, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	uint32_t smem32 = (uint32_t)__cvta_generic_to_shared(wr_smem_v4);
	if (tid == 0) {
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar32), "r"(kV3TableBytes) : "memory");
		constexpr unsigned kChunk = 16384;
		for (unsigned off = 0; off < kV3TableBytes; off += kChunk) {
			const unsigned nb = min(kChunk, kV3TableBytes - off);
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
					:: "r"(smem32 + off), "l"(reinterpret_cast<const char*>(v.delta) + off), "r"(nb), "r"(bar32) : "memory");
		}
	}
	if (a.ts && blockIdx.x == 0 && tid == 0)
		a.ts[kTsChanStart] = global_ns();
	if (a.cta_ts && tid == 0)
		a.cta_ts[2 * blockIdx.x] = global_ns();
	if (a.in_flag) {
		// pipelined host path: the copy-in stream raises the flag behind the tuner block (see v3)
		if (warp == 0) {
			const unsigned long long t0 = global_ns();
			for (;;) {
				unsigned seen;
				asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(a.in_flag) : "memory");
				if (__any_sync(0xFFFFFFFFu, (int)(seen - a.in_seq) >= 0))
					break;
				if (__any_sync(0xFFFFFFFFu, global_ns() - t0 > kSpinNs)) {
					if (lane == 0)
						atomicOr(a.err, kSyncTimeout);
					break;
				}
				__nanosleep(a.poll_ns);
			}
		}
		__syncthreads();
	}
	if (a.ts && blockIdx.x == 0 && tid == 0)
		a.ts[kTsChanInput] = global_ns();

	const Lo3Regs lo = lo3_regs(v.eps, smem32 + kV3MidOffset, v.prmtHi);
	const f2_t nz = f2_pack(v.negzero, v.negzero);
	const unsigned NS = v.NS;
	const unsigned perWarp = NS * G::kStageBytes + G::kTapBytes;
	const uint32_t ring32 = smem32 + kV3TableBytes + warp * perWarp;
	const uint32_t taps32 = ring32 + NS * G::kStageBytes;
	const unsigned K = v.K;
	bool tableReady = false;

	// cp.async work of this lane per stage: chunks lane, lane+32, ... of the stage's 160 (5 per row);
	// chunk c belongs to row c / 5 and lands at byte 16 c of the stage
	constexpr int CPR = (int)(G::kRowBytes / 16);             // chunks per row
	constexpr int NCH = CPR;                                  // chunks per lane: 32 rows * CPR / 32 lanes
	unsigned crow[NCH], ccol[NCH];
	#pragma unroll
	for (int i = 0; i < NCH; i++) {
		const unsigned c = lane + 32u * (unsigned)i;
		crow[i] = c / CPR;
		ccol[i] = c % CPR;
	}

	// units of this warp: receiver-major, spread over the CTAs
	const unsigned W = gridDim.x * nWarps;
	for (unsigned unit = warp * gridDim.x + blockIdx.x; unit < v.nUnits; unit += W) {
		const unsigned r = unit / v.G, g = unit - r * v.G;
		const RxConf cf = a.conf[r];
		const uint32_t ph0 = a.st_in[r].phase;
		const int32_t step = cf.step;
		const char *__restrict__ src = reinterpret_cast<const char*>(a.iq) + (size_t)cf.stream * a.stream_stride * 8u;
		const unsigned L = g * 32u + lane;                       // this lane's run
		const unsigned k0 = L * K;                               // its first output
		__syncwarp();
		// ---- the receiver's taps, rows of one period: row a, stage s, frame i <- tap a*D1 + s*SFR + i ----
		for (unsigned e = lane; e < (unsigned)(ROWS * S * TS); e += 32) {
			const unsigned aa = e / (S * TS), rest = e - aa * (S * TS), ss = rest / TS, ii = rest - ss * TS;
			const unsigned j = aa * D1 + ss * SFR + ii;
			sts32(taps32 + 4u * e, (ii < (unsigned)SFR && j < (unsigned)N1) ? a.taps1[(size_t)r * N1 + j] : 0.0f);
		}
		if (!tableReady) {
			mbar_wait(bar32, 0);      // first use of the NCO table: the bulk copies must have landed
			tableReady = true;
		}
		// ---- prologue (first warp of a receiver): the outputs whose windows reach into the history ----
		if (g == 0) {
			const uint32_t scr32 = ring32;                       // [history (N1-1) | mixed frames 0 ...], float2 each
			for (unsigned c = lane; c < G::kScratchFrames; c += 32) {
				float2 m;
				if (c < (unsigned)(N1 - 1)) {
					m = a.hist_in[(size_t)r * (N1 - 1) + c];
				} else {
					const unsigned f = c - (unsigned)(N1 - 1);
					float sn, cs;
					lo3_sincos(((ph0 + f * (uint32_t)step) << 1) + 0x80000000u, lo, sn, cs);
					const float2 x = f < a.F ? __ldg(reinterpret_cast<const float2*>(src) + f) : make_float2(0.0f, 0.0f);
					m = mix(x, cs, sn);
				}
				sts64(scr32 + 8u * c, m);
			}
			__syncwarp();
			if (lane < (unsigned)G::KSKIP && lane < a.M1) {
				f2_t acc = 0ull;
				const float *tp = a.taps1 + (size_t)r * N1;
				#pragma unroll 4
				for (int j = 0; j < N1; j++)
					acc = tap3(acc, __ldg(tp + j), lds64p(scr32 + 8u * (lane * (unsigned)D1 + (unsigned)j)), nz);
				float2 y;
				f2_unpack(acc, y.x, y.y);
				a.chan[(size_t)r * a.chan_stride + lane] = y;
			}
			__syncwarp();
		}

		// ---- the stream: block coordinate c0 = k0*D1 is input frame c0 - (N1-1) ----
		const long long f0 = (long long)k0 * D1 - (N1 - 1);      // first frame of this lane's run (negative: history, reads as zero)
		const long long f0w = (long long)(g * 32u) * K * D1 - (N1 - 1);            // ... of the warp's first run
		const long long f0l = (long long)(g * 32u + 31u) * K * D1 - (N1 - 1);      // ... of its last run
		const unsigned nPeriods = K + (unsigned)AP;              // the last one only up to stage SFIN
		const unsigned nStages = (nPeriods - 1) * S + (unsigned)G::SFIN + 1;
		const size_t rowPitch = (size_t)K * D1 * 8u;             // bytes between the runs of neighbouring lanes
		const char *cbase[NCH];
		long long cf0[NCH];
		#pragma unroll
		for (int i = 0; i < NCH; i++) {
			cf0[i] = f0w + (long long)crow[i] * K * D1 + 2 * (long long)ccol[i];    // first frame of the chunk in stage 0
			cbase[i] = src + cf0[i] * 8;
		}
		auto fetch = [&](unsigned n) {
			// stage n of every run of the warp into ring slot n % NS
			const uint32_t dst = ring32 + (n % NS) * G::kStageBytes + 16u * lane;
			const long long adv = (long long)n * SFR;
			if (f0w + adv >= 0 && f0l + adv + SFR <= (long long)a.F) {
				#pragma unroll
				for (int i = 0; i < NCH; i++)
					cp_async16(dst + 512u * (unsigned)i, cbase[i] + adv * 8);
			} else {
				#pragma unroll
				for (int i = 0; i < NCH; i++) {
					const long long f = cf0[i] + adv;             // even: a chunk never straddles frame 0
					const long long left = (long long)a.F - f;
					const unsigned nb = (f < 0 || left <= 0) ? 0u : (left >= 2 ? 16u : 8u);
					cp_async16z(dst + 512u * (unsigned)i, nb ? cbase[i] + adv * 8 : src, nb);
				}
			}
		};
		// fill the ring
		for (unsigned n = 0; n + 1 < NS; n++) {
			if (n < nStages)
				fetch(n);
			cp_async_commit();
		}
		f2_t acc[ROWS];
		#pragma unroll
		for (int i = 0; i < ROWS; i++)
			acc[i] = 0ull;
		// biased doubled phase of the run's first frame, and of a frame's step
		uint32_t q = ((ph0 + (uint32_t)(int32_t)f0 * (uint32_t)step) << 1) + 0x80000000u;
		const uint32_t qs = 2u * (uint32_t)step;
		const uint32_t row32 = lane * G::kRowBytes;
		float2 *out = a.chan + (size_t)r * a.chan_stride;
		unsigned n = 0;
		for (unsigned p = 0; p < nPeriods; p++) {
			// a period begins: every live output is one period older, a new one starts
			#pragma unroll
			for (int i = ROWS - 1; i > 0; i--)
				acc[i] = acc[i - 1];
			acc[0] = 0ull;
			const unsigned sEnd = (p + 1 == nPeriods) ? (unsigned)G::SFIN + 1 : (unsigned)S;
			for (unsigned s = 0; s < sEnd; s++, n++) {
				// stage n has landed (all but the NS-2 youngest groups are complete) ...
				if (NS == 2) cp_async_wait<0>(); else if (NS == 3) cp_async_wait<1>(); else cp_async_wait<2>();
				__syncwarp();
				// ... and slot (n-1) % NS is free: every lane is past its reads of stage n-1
				if (n + NS - 1 < nStages)
					fetch(n + NS - 1);
				cp_async_commit();
				const uint32_t st32 = ring32 + (n % NS) * G::kStageBytes + row32;
				f2_t raw[SFR];
				#pragma unroll
				for (int i = 0; i < SFR; i += 2)
					lds128p(st32 + 8u * (unsigned)i, raw[i], raw[i + 1]);
				const uint32_t tap32 = taps32 + 4u * s * (unsigned)TS;
				f2_t fin = 0ull;
				if (s == (unsigned)G::SFIN) {
					v4_stage<N1, D1, G::CFIN, true>(raw, acc, q, qs, lo, tap32, nz, fin);
					// the output that began AP periods ago is complete
					const unsigned k = k0 + p - (unsigned)AP;
					if (p >= (unsigned)AP && k < a.M1 && k >= (unsigned)G::KSKIP) {
						float2 y;
						f2_unpack(fin, y.x, y.y);
						out[k] = y;
					}
				} else if (s < (unsigned)G::SFIN) {
					v4_stage<N1, D1, SFR, false>(raw, acc, q, qs, lo, tap32, nz, fin);
				} else {
					v4_stage<N1, D1, 0, false>(raw, acc, q, qs, lo, tap32, nz, fin);
				}
				q += (uint32_t)SFR * qs;
			}
		}
		cp_async_wait<0>();
		// ---- epilogue (first warp of a receiver): the carried state of the next block ----
		if (g == 0) {
			for (unsigned i = lane; i < (unsigned)(N1 - 1); i += 32) {
				// the last N1-1 mixed frames of [history | block]
				const long long f = (long long)a.F - (N1 - 1) + i;
				float2 m;
				if (f < 0) {
					m = a.hist_in[(size_t)r * (N1 - 1) + (unsigned)(f + (N1 - 1))];
				} else {
					float sn, cs;
					lo3_sincos(((ph0 + (uint32_t)f * (uint32_t)step) << 1) + 0x80000000u, lo, sn, cs);
					m = mix(__ldg(reinterpret_cast<const float2*>(src) + f), cs, sn);
				}
				a.hist_out[(size_t)r * (N1 - 1) + i] = m;
			}
			if (lane == 0)
				a.st_out[r].phase = phase_at(ph0, step, a.F);
		}
	}
	if (a.ts && blockIdx.x == 0 && tid == 0)
		a.ts[kTsChanEnd] = global_ns();
	if (a.cta_ts && tid == 0)
		a.cta_ts[2 * blockIdx.x + 1] = global_ns();
	// programmatic dependent launch: see the end of chan_body_v3
	asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ------------------------------------------------------------------ host side ----

typedef void (*V4Kernel)(const ChanArgs, const V4Args);

struct V4Plan {
	bool ok = false;
	V4Kernel kernel = nullptr;
	int regs = 0;
	int numSMs = 0;
	unsigned n1 = 0, d1 = 0;
	unsigned kskip = 0;          // outputs the prologue computes
	unsigned stageBytes = 0, tapBytes = 0, scratchBytes = 0;
	size_t smemMax = 0;
	unsigned maxPerStream = 1;   // most receivers that share one tuner stream (v4_set_groups)
	unsigned G = 0;              // WR_V4_G: warps per receiver (0 = by bank size)
	unsigned NS = 0;             // WR_V4_NS: ring depth (0 = what fits, at most 4)
	unsigned maxWarps = kV4MaxWarps;   // WR_V4_WARPS
	bool pdl = true;
};

template <int N1, int D1>
inline void v4_fill(V4Plan &p)
{
	using G = V4Geo<N1, D1>;
	p.kernel = chan_kernel_v4<N1, D1>;
	p.kskip = G::KSKIP;
	p.stageBytes = G::kStageBytes;
	p.tapBytes = G::kTapBytes;
	p.scratchBytes = G::kScratchFrames * 8u;
}

// v4 serves the long-filter geometries of the BASELINE configs; it needs the v3 plan's table.
inline int v4_init(V4Plan &p, const V3Plan &v3, int device, unsigned n1, unsigned d1)
{
	p.ok = false;
	p.n1 = n1;
	p.d1 = d1;
	if (n1 == 255 && d1 == 50) v4_fill<255, 50>(p);
	else if (n1 == 127 && d1 == 50) v4_fill<127, 50>(p);
	else if (n1 == 127 && d1 == 40) v4_fill<127, 40>(p);
	else return WR_OK;
	if (!v3.d_delta)
		return WR_OK;
	if (const char *e = getenv("WR_V4_G"))
		p.G = (unsigned)std::max(0, atoi(e));
	if (const char *e = getenv("WR_V4_NS"))
		p.NS = (unsigned)std::max(0, atoi(e));
	if (const char *e = getenv("WR_V4_WARPS"))
		p.maxWarps = (unsigned)std::min(kV4MaxWarps, std::max(1, atoi(e)));
	if (const char *e = getenv("WR_V3_PDL"))
		p.pdl = atoi(e) != 0;
	cudaDeviceProp prop;
	WR_CUDA(cudaGetDeviceProperties(&prop, device));
	p.numSMs = prop.multiProcessorCount;
	p.smemMax = prop.sharedMemPerBlockOptin;
	WR_CUDA(cudaFuncSetAttribute(p.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smemMax));
	cudaFuncAttributes fa;
	WR_CUDA(cudaFuncGetAttributes(&fa, p.kernel));
	p.regs = fa.numRegs;
	p.ok = true;
	return WR_OK;
}

inline void v4_set_groups(V4Plan &p, const RxConf *h_conf, unsigned R, unsigned T)
{
	std::vector<unsigned> count(T, 0);
	unsigned most = 0;
	for (unsigned r = 0; r < R; r++)
		most = std::max(most, ++count[h_conf[r].stream % T]);
	p.maxPerStream = std::max(1u, most);
}

struct V4Launch {
	unsigned G, K, NS, warps;
};

// How a block of F frames would be cut; false if v4 does not serve it.
inline bool v4_shape(const V4Plan &p, const V3Plan &v3, unsigned R, unsigned F, const void *iq, size_t stream_stride, V4Launch *out)
{
	if (!p.ok || !v3.ok)          // (v3.ok: the table survived the compression)
		return false;
	// independent streams only: a run re-reads its tuner stream from L2/HBM, which a bank of 64
	// receivers per tuner cannot afford; runs must start on 16-byte boundaries
	if (p.maxPerStream > 2 || ((uintptr_t)iq & 15u) || (stream_stride & 1u))
		return false;
	const unsigned M1 = F / p.d1;
	if (F < p.n1 - 1 || M1 < 32u * 2u * p.kskip)
		return false;
	// warps per receiver: one, unless the bank is too small to give every SM a few warps that way
	unsigned G = p.G;
	if (G == 0) {
		G = 1;
		while ((unsigned long long)R * G < 4ull * (unsigned)p.numSMs && (M1 + 32u * 2u * G - 1) / (32u * 2u * G) >= 4u * p.kskip)
			G *= 2;
	}
	unsigned K = (M1 + 32u * G - 1) / (32u * G);
	if (K < p.kskip + 1)
		return false;
	const unsigned long long units = (unsigned long long)R * G;
	if (units < (unsigned long long)p.numSMs)     // not enough work for a persistent grid of independent warps
		return false;
	const size_t table = kV3TableBytes;
	unsigned warps = (unsigned)std::min<unsigned long long>(p.maxWarps, (units + p.numSMs - 1) / p.numSMs);
	const unsigned regCap = 65536u / (32u * (unsigned)((p.regs + 7) & ~7));
	warps = std::max(1u, std::min(warps, regCap));
	unsigned NS = p.NS;
	for (;;) {
		// ring depth: what fits beside the table, at most 4 stages, at least 2; the prologue's scratch must fit the ring
		const size_t avail = (p.smemMax - table - 64) / warps;
		unsigned fit = avail > p.tapBytes ? (unsigned)((avail - p.tapBytes) / p.stageBytes) : 0;
		const unsigned need = std::max(2u, (p.scratchBytes + p.stageBytes - 1) / p.stageBytes);
		if (fit >= need) {
			NS = NS ? std::min(std::max(NS, need), fit) : std::min(fit, std::max(4u, need));
			break;
		}
		if (warps == 1)
			return false;
		warps--;
	}
	out->G = G; out->K = K; out->NS = NS; out->warps = warps;
	return true;
}

inline int v4_launch_chan(V4Plan &p, const V3Plan &v3, const V4Launch &L, ChanArgs &ca, unsigned R, cudaStream_t st, unsigned long long *launches)
{
	V4Args v;
	v.delta = v3.d_delta;
	v.eps = v3.coef.eps;
	v.negzero = -0.0f;
	v.prmtHi = 0x4B00u;
	v.G = L.G;
	v.K = L.K;
	v.NS = L.NS;
	v.nUnits = R * L.G;
	const size_t smem = kV3TableBytes + (size_t)L.warps * ((size_t)L.NS * p.stageBytes + p.tapBytes);
	const unsigned grid = (unsigned)std::min<unsigned long long>((v.nUnits + L.warps - 1) / L.warps, (unsigned long long)p.numSMs);
	cudaLaunchConfig_t cfg = {};
	cudaLaunchAttribute attr[1];
	cfg.gridDim = dim3(grid);
	cfg.blockDim = dim3(L.warps * 32);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = st;
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = p.pdl ? 1 : 0;
	cudaError_t e = cudaLaunchKernelEx(&cfg, p.kernel, (const ChanArgs)ca, (const V4Args)v);
	(*launches)++;
	if (e == cudaSuccess)
		e = cudaGetLastError();
	if (e != cudaSuccess) {
		wr::set_error("chan_kernel_v4 launch (grid %u, %u warps, %zu bytes of shared memory): %s",
				grid, L.warps, smem, cudaGetErrorString(e));
		return WR_ECUDA;
	}
	return WR_OK;
}

} // namespace wrd
