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
	unsigned nUnits;          // R * G
	unsigned scratchBytes;    // per-warp scratch of the prologue behind the taps (0: the ring serves, and is filled after the prologue)
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
__device__ __forceinline__ void v4_stage(uint32_t st32, f2_t (&acc)[V4Geo<N1, D1>::ROWS],
		uint32_t q0, uint32_t qs, const Lo3Regs &lo, uint32_t tap32, f2_t nz, f2_t &fin)
{
	using G = V4Geo<N1, D1>;
	constexpr int SFR = G::SFR, AP = G::AP, TS = G::TS, S = G::S;
	constexpr int H = SFR / 2;
	f2_t raw[SFR];             // this lane's frames of the stage: its row of the ring slot
	#pragma unroll
	for (int i = 0; i < SFR; i += 2)
		lds128p(st32 + 8u * (unsigned)i, raw[i], raw[i + 1]);
	float4 t[G::ROWS];         // the taps of four consecutive frames, one 128-bit broadcast load per live row
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
			if (fi % 4 == 0) {
				#pragma unroll
				for (int a = 0; a < AP; a++)
					t[a] = lds128f(tap32 + 4u * (unsigned)(a * S * TS + fi));
				if (fi < CX)
					t[AP] = lds128f(tap32 + 4u * (unsigned)(AP * S * TS + fi));
			}
			// downconverter.cxx:109-110:  I' = i*cos + q*sin ;  Q' = q*cos - i*sin
			float ic, qc, is, qq;
			f2_unpack(f2_fma(raw[fi], f2_pack(cs[i], cs[i]), nz), ic, qc);
			f2_unpack(f2_fma(raw[fi], f2_pack(sn[i], sn[i]), nz), is, qq);
			const f2_t m = f2_pack(__fadd_rn(ic, qq), __fsub_rn(qc, is));
			// lowpass.cxx:155-156 for every output whose window covers this frame: the output of
			// age a takes tap a*D1 + (position in the period)
			#pragma unroll
			for (int a = 0; a < AP; a++) {
				const float c = (fi % 4 == 0) ? t[a].x : (fi % 4 == 1) ? t[a].y : (fi % 4 == 2) ? t[a].z : t[a].w;
				acc[a] = tap3(acc[a], c, m, nz);
			}
			if (fi < CX) {
				const float c = (fi % 4 == 0) ? t[AP].x : (fi % 4 == 1) ? t[AP].y : (fi % 4 == 2) ? t[AP].z : t[AP].w;
				acc[AP] = tap3(acc[AP], c, m, nz);
				if (FIN && fi == CX - 1)
					fin = acc[AP];
			}
		}
	}
}

// What a run carries from stage to stage (registers once everything is inlined).
template <int N1, int D1>
struct V4Run {
	using G = V4Geo<N1, D1>;
	static constexpr int NCH = (int)(G::kRowBytes / 16);    // 16-byte chunks a lane copies per stage (32 rows * chunks per row / 32 lanes)
	f2_t acc[G::ROWS];         // live outputs by age
	uint32_t q;                // biased doubled phase of the next frame
	uint32_t qs;               // ... and its step per frame
	uint32_t st32;             // this lane's row in the ring slot of the stage being computed
	uint32_t stEnd, ringBytes;
	unsigned nf, nStages;      // next stage to fetch / stages of the run
	unsigned nLo, nHi;         // stages [nLo, nHi) lie inside the block for every run of the warp
	const char *cptr[NCH];     // this lane's chunks of stage 0 of the period being computed
	long long cf0[NCH];        // first frame of those chunks in stage 0 of the run
	const char *src;           // the stream
	unsigned F;
	unsigned laneOff;          // 16 * lane - lane * kRowBytes: from st32 to this lane's first chunk of the slot
};

// The rare stage that reaches outside the block (history in front of it, nothing behind it):
// frames outside read as zero through the copies' source size.
template <int N1, int D1>
__device__ __noinline__ void v4_fetch_edge(uint32_t dst, const char *c0, const char *c1, const char *c2, const char *c3, const char *c4,
		long long f0, long long f1, long long f2, long long f3, long long f4, long long F, const char *src)
{
	const char *c[5] = { c0, c1, c2, c3, c4 };
	const long long f[5] = { f0, f1, f2, f3, f4 };
	#pragma unroll
	for (int i = 0; i < V4Run<N1, D1>::NCH; i++) {
		const long long left = F - f[i];                     // (a chunk starts on an even frame: it never straddles frame 0)
		const unsigned nb = (f[i] < 0 || left <= 0) ? 0u : (left >= 2 ? 16u : 8u);
		cp_async16z(dst + 512u * (unsigned)i, nb ? c[i] : src, nb);
	}
}

// Copies stage (current + NS - 1) of every run of the warp into the ring slot the stage before the
// current one has just left.  SOFF = that stage's position counted from the start of the period
// being computed, so its source is an immediate offset from the period's chunk pointers.
template <int N1, int D1, int SOFF>
__device__ __forceinline__ void v4_fetch(V4Run<N1, D1> &r, uint32_t slot32)
{
	using G = V4Geo<N1, D1>;
	constexpr int NCH = V4Run<N1, D1>::NCH;
	static_assert(NCH == 5, "v4_fetch_edge takes five chunks");
	if (r.nf < r.nStages) {
		const uint32_t dst = slot32 + r.laneOff;
		if (r.nf >= r.nLo && r.nf < r.nHi) {
			#pragma unroll
			for (int i = 0; i < NCH; i++)
				cp_async16(dst + 512u * (unsigned)i, r.cptr[i] + SOFF * (int)G::kRowBytes);
		} else {
			const long long adv = (long long)r.nf * G::SFR;
			v4_fetch_edge<N1, D1>(dst, r.cptr[0] + SOFF * (int)G::kRowBytes, r.cptr[1] + SOFF * (int)G::kRowBytes,
					r.cptr[2] + SOFF * (int)G::kRowBytes, r.cptr[3] + SOFF * (int)G::kRowBytes, r.cptr[4] + SOFF * (int)G::kRowBytes,
					r.cf0[0] + adv, r.cf0[1] + adv, r.cf0[2] + adv, r.cf0[3] + adv, r.cf0[4] + adv, (long long)r.F, r.src);
		}
		r.nf++;
	}
	cp_async_commit();
}

// Stages [SB, SE) of one period, unrolled: every stage index is a compile-time constant, so tap
// offsets, the variant of the stage body and the sources of the copies are immediates.
template <int N1, int D1, int NS, int SB, int SE>
__device__ __forceinline__ void v4_period(V4Run<N1, D1> &r, const Lo3Regs &lo, uint32_t taps32, f2_t nz, f2_t &fin)
{
	using G = V4Geo<N1, D1>;
	if constexpr (SB < SE) {
		// stage n has landed (all but the NS-2 youngest groups are complete) ...
		cp_async_wait<NS - 2>();
		__syncwarp();
		// ... and the slot of stage n-1 is free: every lane is past its reads of it
		const uint32_t prev32 = (r.st32 == r.stEnd - r.ringBytes + 0u) ? r.stEnd - G::kStageBytes : r.st32 - G::kStageBytes;
		v4_fetch<N1, D1, SB + NS - 1>(r, prev32);
		constexpr int CX = (G::REM + 1 - SB * G::SFR) < 0 ? 0 : ((G::REM + 1 - SB * G::SFR) > G::SFR ? G::SFR : (G::REM + 1 - SB * G::SFR));
		v4_stage<N1, D1, CX, SB == G::SFIN>(r.st32, r.acc, r.q, r.qs, lo, taps32 + 4u * (unsigned)(SB * G::TS), nz, fin);
		r.q += (uint32_t)G::SFR * r.qs;
		r.st32 += G::kStageBytes;
		if (r.st32 == r.stEnd)
			r.st32 -= r.ringBytes;
		v4_period<N1, D1, NS, SB + 1, SE>(r, lo, taps32, nz, fin);
	}
}

template <int N1, int D1, int WMAX, int NS>
__global__ void __launch_bounds__(WMAX * 32, 1) chan_kernel_v4(const ChanArgs a, const V4Args v)
{
	using G = V4Geo<N1, D1>;
	using Run = V4Run<N1, D1>;
	constexpr int SFR = G::SFR, S = G::S, AP = G::AP, ROWS = G::ROWS, TS = G::TS, NCH = Run::NCH;
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
	// per warp: [ring: NS stages | taps | scratch of the prologue (absent when it does not fit: the ring serves)]
	const unsigned perWarp = NS * G::kStageBytes + G::kTapBytes + v.scratchBytes;
	const uint32_t ring32 = smem32 + kV3TableBytes + warp * perWarp;
	const uint32_t taps32 = ring32 + NS * G::kStageBytes;
	const uint32_t scr32 = v.scratchBytes ? taps32 + G::kTapBytes : ring32;
	const unsigned K = v.K;
	bool tableReady = false;

	// units of this warp: receiver-major, spread over the CTAs
	const unsigned W = gridDim.x * nWarps;
	for (unsigned unit = warp * gridDim.x + blockIdx.x; unit < v.nUnits; unit += W) {
		const unsigned rx = unit / v.G, g = unit - rx * v.G;
		const RxConf cf = a.conf[rx];
		const uint32_t ph0 = a.st_in[rx].phase;
		const int32_t step = cf.step;
		const char *__restrict__ src = reinterpret_cast<const char*>(a.iq) + (size_t)cf.stream * a.stream_stride * 8u;
		const unsigned L = g * 32u + lane;                       // this lane's run
		const unsigned k0 = L * K;                               // its first output
		__syncwarp();

		// ---- the stream: block coordinate c0 = k0*D1 is input frame c0 - (N1-1) ----
		Run r;
		const long long f0 = (long long)k0 * D1 - (N1 - 1);      // first frame of this lane's run (negative: history, reads as zero)
		const long long f0w = (long long)(g * 32u) * K * D1 - (N1 - 1);            // ... of the warp's first run
		const long long f0l = (long long)(g * 32u + 31u) * K * D1 - (N1 - 1);      // ... of its last run
		const unsigned nPeriods = K + (unsigned)AP;              // the last one only up to stage SFIN
		r.nStages = (nPeriods - 1) * S + (unsigned)G::SFIN + 1;
		r.nLo = f0w < 0 ? (unsigned)((-f0w + SFR - 1) / SFR) : 0u;
		const long long room = ((long long)a.F - f0l) / SFR;
		r.nHi = room <= 0 ? 0u : (unsigned)(room < (long long)r.nStages ? room : (long long)r.nStages);
		r.src = src;
		r.F = a.F;
		#pragma unroll
		for (int i = 0; i < NCH; i++) {
			// chunk lane + 32 i of a stage's 160: row c / 5, 16-byte column c % 5; it lands at byte 16 c of the slot
			const unsigned c = lane + 32u * (unsigned)i;
			r.cf0[i] = f0w + (long long)(c / NCH) * K * D1 + 2 * (long long)(c % NCH);
			r.cptr[i] = src + r.cf0[i] * 8;
		}
		r.nf = 0;
		r.ringBytes = NS * G::kStageBytes;
		r.st32 = ring32 + lane * G::kRowBytes;
		r.stEnd = r.st32 + r.ringBytes;
		r.laneOff = 16u * lane - lane * G::kRowBytes;
		const bool ringIsScratch = (v.scratchBytes == 0);
		if (!ringIsScratch || g != 0) {
			// fill the ring now: the copies fly while the taps are staged and the prologue runs
			v4_fetch<N1, D1, 0>(r, ring32 + lane * G::kRowBytes);
			if (NS > 2) v4_fetch<N1, D1, 1>(r, ring32 + G::kStageBytes + lane * G::kRowBytes);
			if (NS > 3) v4_fetch<N1, D1, 2>(r, ring32 + 2 * G::kStageBytes + lane * G::kRowBytes);
		}

		// ---- the receiver's taps, rows of one period: row a, stage s, frame i <- tap a*D1 + s*SFR + i ----
		{
			constexpr int NE = ROWS * S * TS, PER = (NE + 31) / 32;
			float tv[PER];
			#pragma unroll
			for (int e = 0; e < PER; e++) {
				const unsigned idx = lane + 32u * (unsigned)e;
				const unsigned aa = idx / (S * TS), rest = idx - aa * (S * TS), ss = rest / TS, ii = rest - ss * TS;
				const unsigned j = aa * D1 + ss * SFR + ii;
				tv[e] = (idx < (unsigned)NE && ii < (unsigned)SFR && j < (unsigned)N1) ? __ldg(a.taps1 + (size_t)rx * N1 + j) : 0.0f;
			}
			#pragma unroll
			for (int e = 0; e < PER; e++)
				if (lane + 32u * (unsigned)e < (unsigned)NE)
					sts32(taps32 + 4u * (lane + 32u * (unsigned)e), tv[e]);
		}
		if (!tableReady) {
			mbar_wait(bar32, 0);      // first use of the NCO table: the bulk copies must have landed
			tableReady = true;
		}
		__syncwarp();
		// ---- prologue (first warp of a receiver): the outputs whose windows reach into the history ----
		if (g == 0) {
			// [history (N1-1) | mixed frames 0 ...], float2 each; four entries per lane and round, loads first
			for (unsigned c0 = 0; c0 < G::kScratchFrames; c0 += 128) {
				float2 x[4];
				#pragma unroll
				for (int u = 0; u < 4; u++) {
					const unsigned c = c0 + lane + 32u * (unsigned)u;
					if (c < (unsigned)(N1 - 1))
						x[u] = a.hist_in[(size_t)rx * (N1 - 1) + c];
					else if (c < G::kScratchFrames && c - (unsigned)(N1 - 1) < a.F)
						x[u] = __ldg(reinterpret_cast<const float2*>(src) + (c - (unsigned)(N1 - 1)));
					else
						x[u] = make_float2(0.0f, 0.0f);
				}
				#pragma unroll
				for (int u = 0; u < 4; u++) {
					const unsigned c = c0 + lane + 32u * (unsigned)u;
					if (c >= (unsigned)(N1 - 1) && c < G::kScratchFrames) {
						const unsigned f = c - (unsigned)(N1 - 1);
						float sn, cs;
						lo3_sincos(((ph0 + f * (uint32_t)step) << 1) + 0x80000000u, lo, sn, cs);
						x[u] = mix(x[u], cs, sn);
					}
					if (c < G::kScratchFrames)
						sts64(scr32 + 8u * c, x[u]);
				}
			}
			__syncwarp();
			if (lane < (unsigned)G::KSKIP && lane < a.M1) {
				f2_t acc = 0ull;
				uint32_t x32 = scr32 + 8u * lane * (unsigned)D1;
				#pragma unroll 1
				for (int aa = 0; aa < ROWS; aa++) {
					#pragma unroll 1
					for (int ss = 0; ss < S; ss++) {
						const uint32_t t32 = taps32 + 4u * (unsigned)((aa * S + ss) * TS);
						#pragma unroll
						for (int ii = 0; ii < SFR; ii++)
							if (aa * D1 + ss * SFR + ii < N1)
								acc = tap3(acc, lds32(t32 + 4u * (unsigned)ii), lds64p(x32 + 8u * (unsigned)ii), nz);
						x32 += 8u * (unsigned)SFR;
					}
				}
				float2 y;
				f2_unpack(acc, y.x, y.y);
				a.chan[(size_t)rx * a.chan_stride + lane] = y;
			}
			__syncwarp();
			if (ringIsScratch) {
				v4_fetch<N1, D1, 0>(r, ring32 + lane * G::kRowBytes);
				if (NS > 2) v4_fetch<N1, D1, 1>(r, ring32 + G::kStageBytes + lane * G::kRowBytes);
				if (NS > 3) v4_fetch<N1, D1, 2>(r, ring32 + 2 * G::kStageBytes + lane * G::kRowBytes);
			}
		}
		// v4_fetch takes its sources relative to the period being computed: the ring fill above read
		// stages 0 .. NS-2 of period 0 with SOFF = 0 .. NS-2, exactly where cptr points

		#pragma unroll
		for (int i = 0; i < ROWS; i++)
			r.acc[i] = 0ull;
		r.q = ((ph0 + (uint32_t)(int32_t)f0 * (uint32_t)step) << 1) + 0x80000000u;
		r.qs = 2u * (uint32_t)step;
		float2 *out = a.chan + (size_t)rx * a.chan_stride;
		unsigned kdone = k0 - (unsigned)AP;                     // the output that completes in the current period (wraps below zero at first)
		for (unsigned p = 0; p < nPeriods; p++) {
			// a period begins: every live output is one period older, a new one starts
			#pragma unroll
			for (int i = ROWS - 1; i > 0; i--)
				r.acc[i] = r.acc[i - 1];
			r.acc[0] = 0ull;
			f2_t fin = 0ull;
			if (p + 1 < nPeriods)
				v4_period<N1, D1, NS, 0, S>(r, lo, taps32, nz, fin);
			else
				v4_period<N1, D1, NS, 0, G::SFIN + 1>(r, lo, taps32, nz, fin);
			// the output that began AP periods ago is complete
			if (kdone - k0 < K && kdone < a.M1 && kdone >= (unsigned)G::KSKIP) {
				float2 y;
				f2_unpack(fin, y.x, y.y);
				out[kdone] = y;
			}
			kdone++;
			#pragma unroll
			for (int i = 0; i < NCH; i++)
				r.cptr[i] += D1 * 8;
		}
		cp_async_wait<0>();
		// ---- epilogue (first warp of a receiver): the carried state of the next block ----
		if (g == 0) {
			for (unsigned i0 = 0; i0 < (unsigned)(N1 - 1); i0 += 64) {
				// the last N1-1 mixed frames of [history | block]
				float2 x[2];
				#pragma unroll
				for (int u = 0; u < 2; u++) {
					const unsigned i = i0 + lane + 32u * (unsigned)u;
					const long long f = (long long)a.F - (N1 - 1) + i;
					if (i >= (unsigned)(N1 - 1))
						x[u] = make_float2(0.0f, 0.0f);
					else if (f < 0)
						x[u] = a.hist_in[(size_t)rx * (N1 - 1) + (unsigned)(f + (N1 - 1))];
					else
						x[u] = __ldg(reinterpret_cast<const float2*>(src) + f);
				}
				#pragma unroll
				for (int u = 0; u < 2; u++) {
					const unsigned i = i0 + lane + 32u * (unsigned)u;
					const long long f = (long long)a.F - (N1 - 1) + i;
					if (i < (unsigned)(N1 - 1)) {
						if (f >= 0) {
							float sn, cs;
							lo3_sincos(((ph0 + (uint32_t)f * (uint32_t)step) << 1) + 0x80000000u, lo, sn, cs);
							x[u] = mix(x[u], cs, sn);
						}
						a.hist_out[(size_t)rx * (N1 - 1) + i] = x[u];
					}
				}
			}
			if (lane == 0)
				a.st_out[rx].phase = phase_at(ph0, step, a.F);
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
	V4Kernel kernelW8 = nullptr;   // up to 8 warps per CTA (255 registers per thread), ring of three stages
	V4Kernel kernelW16 = nullptr;  // up to kV4MaxWarps warps (128 registers), ring of two stages
	int regsW8 = 0, regsW16 = 0;
	int numSMs = 0;
	unsigned n1 = 0, d1 = 0;
	unsigned kskip = 0;          // outputs the prologue computes
	unsigned stageBytes = 0, tapBytes = 0, scratchBytes = 0;
	size_t smemMax = 0;
	unsigned maxPerStream = 1;   // most receivers that share one tuner stream (v4_set_groups)
	unsigned G = 0;              // WR_V4_G: warps per receiver (0 = by bank size)
	bool ownScratch = true;      // WR_V4_SCRATCH=0: the prologue always borrows the ring (28 KB less shared memory on cfg3)
	bool pdl = true;
};

template <int N1, int D1>
inline void v4_fill(V4Plan &p)
{
	using G = V4Geo<N1, D1>;
	p.kernelW8 = chan_kernel_v4<N1, D1, 8, 3>;
	p.kernelW16 = chan_kernel_v4<N1, D1, kV4MaxWarps, 2>;
	p.kskip = G::KSKIP;
	p.stageBytes = G::kStageBytes;
	p.tapBytes = G::kTapBytes;
	p.scratchBytes = (G::kScratchFrames * 8u + 15u) & ~15u;
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
	if (const char *e = getenv("WR_V3_PDL"))
		p.pdl = atoi(e) != 0;
	if (const char *e = getenv("WR_V4_SCRATCH"))
		p.ownScratch = atoi(e) != 0;
	cudaDeviceProp prop;
	WR_CUDA(cudaGetDeviceProperties(&prop, device));
	p.numSMs = prop.multiProcessorCount;
	cudaFuncAttributes fa;
	WR_CUDA(cudaFuncGetAttributes(&fa, p.kernelW8));
	p.regsW8 = fa.numRegs;
	// the opt-in limit covers static and dynamic shared memory together (the kernel's mbarrier is static)
	p.smemMax = prop.sharedMemPerBlockOptin - ((fa.sharedSizeBytes + 255) & ~(size_t)255);
	WR_CUDA(cudaFuncSetAttribute(p.kernelW8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smemMax));
	WR_CUDA(cudaFuncGetAttributes(&fa, p.kernelW16));
	p.regsW16 = fa.numRegs;
	WR_CUDA(cudaFuncSetAttribute(p.kernelW16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smemMax));
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
	unsigned G, K, NS, warps, scratch;
	bool w8;
};

// How a block of F frames would be cut; false if v4 does not serve it.
inline bool v4_shape(const V4Plan &p, const V3Plan &v3, unsigned R, unsigned F, const void *iq, size_t stream_stride, bool forced, V4Launch *out)
{
	if (!p.ok || !v3.ok)          // (v3.ok: the table survived the compression)
		return false;
	// independent streams only: a run re-reads its tuner stream from L2/HBM, which a bank of 64
	// receivers per tuner cannot afford; runs must start on 16-byte boundaries
	if ((p.maxPerStream > 2 && !forced) || ((uintptr_t)iq & 15u) || (stream_stride & 1u))
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
	if (units < (unsigned long long)p.numSMs && !forced)     // not enough work for a persistent grid of independent warps
		return false;
	unsigned warps = (unsigned)std::min<unsigned long long>(kV4MaxWarps, (units + p.numSMs - 1) / p.numSMs);
	const bool w8 = warps <= 8;
	const unsigned NS = w8 ? 3u : 2u;
	const size_t avail = p.smemMax - kV3TableBytes;
	// with its own scratch the prologue runs while the first stages are already in flight; without,
	// the ring serves as scratch (it must be large enough) and is filled afterwards
	size_t perWarp = (size_t)NS * p.stageBytes + p.tapBytes + p.scratchBytes;
	unsigned scratch = p.scratchBytes;
	if (perWarp * warps > avail || (!p.ownScratch && (size_t)NS * p.stageBytes >= p.scratchBytes)) {
		scratch = 0;
		perWarp = (size_t)NS * p.stageBytes + p.tapBytes;
		if ((size_t)NS * p.stageBytes < p.scratchBytes)
			return false;
		while (warps > 1 && perWarp * warps > avail)
			warps--;
		if (perWarp * warps > avail)
			return false;
	}
	out->G = G; out->K = K; out->NS = NS; out->warps = warps; out->scratch = scratch; out->w8 = w8;
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
	v.nUnits = R * L.G;
	v.scratchBytes = L.scratch;
	const size_t smem = kV3TableBytes + (size_t)L.warps * ((size_t)L.NS * p.stageBytes + p.tapBytes + L.scratch);
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
	cudaError_t e = cudaLaunchKernelEx(&cfg, L.w8 ? p.kernelW8 : p.kernelW16, (const ChanArgs)ca, (const V4Args)v);
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
