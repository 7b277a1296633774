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
//   * Every receiver's M1 outputs are cut into the same number n_r of runs (of K or K+1 outputs), and
//     the R * n_r runs of the bank are dealt to the lanes of the grid in order: lane L of warp slot s
//     owns run 32 s + L.  n_r is chosen by the host so that ONE round of runs fills every lane
//     of the persistent grid (v4_shape) -- cfg3: 1024 receivers x 37 runs = 148 CTAs x 8 warps x
//     32 lanes exactly, K = 55/56 -- instead of whole receivers per warp, which left a quarter
//     of the schedulers half empty (1024 receivers over 148 SMs = 6.9 warps per SM).  n_r >= 32,
//     so the 32 consecutive runs of a warp belong to at most TWO receivers: both tap sets are
//     staged, a lane reads the one of its receiver.  Warps are independent of each other: no
//     block-wide barrier after start-up, no slot hand-over, no spinning.
//   * (Raw RTL-SDR bytes -- template parameter U8 -- take the same road: a stage's row is 20 bytes,
//     fetched by 4-byte copies, five lanes to a row as well; a lane reads its row with five 32-bit
//     loads and converts two frames per word, (2^23 + b) * 2^-7 - 65537 = (b - 128) / 128 exactly.)
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
//     (its outputs' windows begin there).  They are mixed twice -- (N1-1)/(K*D1) = 9 % at
//     cfg3's K = 56 -- which buys independence; keeping them instead would take 65 KB per
//     warp in flight (32 x 254 frames).
//   * The first outputs of a block, whose windows reach into the carried history, are computed by
//     a short gather prologue from [history | first frames] (staged in the warp's ring before it
//     is filled); the carried history of the NEXT block (the last N1-1 mixed frames) is produced
//     by an epilogue.  Both belong to the warp that owns the receiver's first run.  The steady state therefore
//     has no special cases: frames outside the block read as zero (cp.async src-size).
// Geometry: N1 odd (so that a run starts on a 16-byte boundary), D1 even and a multiple of the
// stage length.  Small shared-tuner banks (cfg2) stay with v3, whose mixers share the raw
// registers among the receivers of a stream; large ones (cfg5) are faster here (v4_shape).
#pragma once

#include "wr_kernels_v3.cuh"

#ifndef WR_V4_BODY_UNROLL
#define WR_V4_BODY_UNROLL 1    // body stages per iteration of the period's loop (A/B knob)
#endif

namespace wrd {

template <int N1, int D1, bool U8 = false>
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
	static constexpr unsigned FB = U8 ? 2 : 8;              // bytes per frame: raw RTL-SDR bytes or float IQ
	static constexpr unsigned CH = U8 ? 4 : 16;             // bytes per cp.async copy (two frames)
	static constexpr unsigned kRowBytes = SFR * FB;         // a run's frames of one stage
	static constexpr unsigned kStageBytes = 32 * kRowBytes;
	static constexpr unsigned kTapBytes = (unsigned)ROWS * S * TS * 4;
	static constexpr unsigned kScratchFrames = (unsigned)(KSKIP - 1) * D1 + N1;   // [history | first frames] of the prologue
	static_assert(D1 % SFR == 0, "a period must be a whole number of stages");
	static_assert(N1 % 2 == 1 && D1 % 2 == 0, "v4 needs an odd tap count and an even decimation (16-byte aligned runs)");
	static_assert((kRowBytes / CH) % 2 == 1, "a row must be an odd number of chunks (bank spread)");
	static_assert(AP >= 1, "the window must span more than one period");
};

// bytes of a warp's ring region: the ring, or the prologue's scratch where that is larger (16-byte multiple)
__host__ __device__ constexpr unsigned v4_ring_region(unsigned ring, unsigned scratch)
{
	return ring >= scratch ? ring : ((scratch + 15u) & ~15u);
}

struct V4Args {
	const int16_t *delta;     // padded corrections (wr_lo3.h), staged to shared memory per CTA
	float eps;
	float negzero;            // -0.0f, opaque to the compiler (mul2_rn_exact)
	unsigned prmtHi;          // 0x4B00, opaque to the compiler
	unsigned runsPerRx;       // n_r >= 32: runs a receiver's outputs are cut into
	unsigned runLen;          // K: outputs of a short run ...
	unsigned longRuns;        // ... the first `longRuns` runs of a receiver have K + 1
	unsigned totalRuns;       // R * n_r
	unsigned R;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}

__device__ __forceinline__ void cp_async16z(uint32_t dst, const void *src, unsigned srcBytes)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(srcBytes) : "memory");
}

__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst), "l"(src) : "memory");
}

__device__ __forceinline__ void cp_async4z(uint32_t dst, const void *src, unsigned srcBytes)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(dst), "l"(src), "r"(srcBytes) : "memory");
}

// Two frames of raw RTL-SDR bytes (i0 q0 i1 q1 in one word) as float IQ: (2^23 + b) * 2^-7 - 65537 =
// (b - 128) / 128 exactly, rtlsdrtuner.cxx:106 (RawIO<true>::cvt of the v3 kernel, for both halves)
__device__ __forceinline__ void v4_cvt_bytes(uint32_t w, uint32_t hi, f2_t &a, f2_t &b)
{
	uint32_t i0, q0, i1, q1;
	asm("prmt.b32 %0, %1, %2, 0x5440;" : "=r"(i0) : "r"(w), "r"(hi));
	asm("prmt.b32 %0, %1, %2, 0x5441;" : "=r"(q0) : "r"(w), "r"(hi));
	asm("prmt.b32 %0, %1, %2, 0x5442;" : "=r"(i1) : "r"(w), "r"(hi));
	asm("prmt.b32 %0, %1, %2, 0x5443;" : "=r"(q1) : "r"(w), "r"(hi));
	const f2_t k = f2_pack(0.0078125f, 0.0078125f), m = f2_pack(-65537.0f, -65537.0f);
	a = f2_fma(f2_pack(__uint_as_float(i0), __uint_as_float(q0)), k, m);
	b = f2_fma(f2_pack(__uint_as_float(i1), __uint_as_float(q1)), k, m);
}

// frame f of a tuner stream as float IQ (the prologue's and the epilogue's gathers)
template <bool U8>
__device__ __forceinline__ float2 v4_load_frame(const char *src, long long f, uint32_t hi)
{
	if constexpr (U8) {
		float2 x;
		f2_unpack(RawIO<true>::cvt(RawIO<true>::load(src + f * 2), hi), x.x, x.y);
		return x;
	} else {
		return __ldg(reinterpret_cast<const float2*>(src) + f);
	}
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
#ifndef WR_V4_AHEAD
#define WR_V4_AHEAD 0          // 1: the NCO of stage n+1 is computed during stage n, its table gathers issued in front of the FIR
#endif
template <int N1, int D1, int CX, bool FIN, bool U8 = false>
__device__ __forceinline__ void v4_stage(uint32_t st32, f2_t (&acc)[V4Geo<N1, D1>::ROWS],
		uint32_t q0, uint32_t qs, const Lo3Regs &lo, uint32_t tap32, f2_t nz, f2_t &fin,
		float (&snc)[V4Geo<N1, D1>::SFR], float (&csc)[V4Geo<N1, D1>::SFR])
{
	using G = V4Geo<N1, D1>;
	constexpr int SFR = G::SFR, AP = G::AP, TS = G::TS, S = G::S;
	// the NCO of a whole stage is evaluated in one batch (ten independent chains; in two batches of five
	// the same code ran 4 % slower on cfg3 and 6 % on cfg5, measured on one box: WR_V4_HB=2)
#ifndef WR_V4_HB
#define WR_V4_HB 1
#endif
	constexpr int H = SFR / WR_V4_HB;
	f2_t raw[SFR];             // this lane's frames of the stage: its row of the ring slot
	if constexpr (U8) {
		uint32_t w[SFR / 2];
		#pragma unroll
		for (int i = 0; i < SFR / 2; i++)
			asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[i]) : "r"(st32 + 4u * (unsigned)i));
		#pragma unroll
		for (int i = 0; i < SFR / 2; i++)
			v4_cvt_bytes(w[i], lo.hi, raw[2 * i], raw[2 * i + 1]);
	} else {
		#pragma unroll
		for (int i = 0; i < SFR; i += 2)
			lds128p(st32 + 8u * (unsigned)i, raw[i], raw[i + 1]);
	}
	float4 t[G::ROWS];         // the taps of four consecutive frames, one 128-bit broadcast load per live row
#if WR_V4_AHEAD
	static_assert(WR_V4_HB == 1, "WR_V4_AHEAD works on whole stages");
	// the NEXT stage's table gathers go out now; its polynomial follows this stage's FIR
	f2_t Fn[SFR];
	int dsn[SFR], dcn[SFR];
	{
		uint32_t q[SFR];
		#pragma unroll
		for (int i = 0; i < SFR; i++)
			q[i] = q0 + (uint32_t)(SFR + i) * qs;
		lo3_issue_n<SFR>(q, lo, Fn, dsn, dcn);
	}
#endif
	#pragma unroll
	for (int h = 0; h < WR_V4_HB; h++) {
#if WR_V4_AHEAD
		float (&sn)[SFR] = snc;
		float (&cs)[SFR] = csc;
#else
		uint32_t q[H];
		float sn[H], cs[H];
		#pragma unroll
		for (int i = 0; i < H; i++)
			q[i] = q0 + (uint32_t)(h * H + i) * qs;
		lo3_sincos_n<H>(q, lo, sn, cs);
#endif
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
#if WR_V4_AHEAD
	lo3_finish_n<SFR>(Fn, dsn, dcn, lo, snc, csc);
#endif
}

// What a run carries from stage to stage (registers once everything is inlined).
template <int N1, int D1, bool U8 = false>
struct V4Run {
	using G = V4Geo<N1, D1, U8>;
	static constexpr int NCH = (int)(G::kRowBytes / G::CH); // chunks a lane copies per stage (32 rows * chunks per row / 32 lanes)
	f2_t acc[G::ROWS];         // live outputs by age
	float sn[G::SFR], cs[G::SFR];   // (WR_V4_AHEAD) the NCO of the stage about to be computed
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
	unsigned laneOff;          // CH * lane - lane * kRowBytes: from st32 to this lane's first chunk of the slot
};

// The rare stage that reaches outside the block (history in front of it, nothing behind it):
// frames outside read as zero through the copies' source size.
template <int N1, int D1, bool U8 = false>
__device__ __noinline__ void v4_fetch_edge(uint32_t dst, const char *c0, const char *c1, const char *c2, const char *c3, const char *c4,
		long long f0, long long f1, long long f2, long long f3, long long f4, long long F, const char *src)
{
	using G = V4Geo<N1, D1, U8>;
	const char *c[5] = { c0, c1, c2, c3, c4 };
	const long long f[5] = { f0, f1, f2, f3, f4 };
	#pragma unroll
	for (int i = 0; i < V4Run<N1, D1, U8>::NCH; i++) {
		const long long left = F - f[i];                     // (a chunk starts on an even frame: it never straddles frame 0)
		const unsigned nb = (f[i] < 0 || left <= 0) ? 0u : (left >= 2 ? G::CH : G::CH / 2);
		if constexpr (U8)
			cp_async4z(dst + 32u * G::CH * (unsigned)i, nb ? c[i] : src, nb);
		else
			cp_async16z(dst + 32u * G::CH * (unsigned)i, nb ? c[i] : src, nb);
	}
}

// Copies stage (current + NS - 1) of every run of the warp into the ring slot the stage before the
// current one has just left.  SOFF = that stage's position counted from the stage cptr points at
// (the one being computed; the ring fill: stage 0), so its source is an immediate offset.
template <int N1, int D1, int SOFF, bool U8 = false>
__device__ __forceinline__ void v4_fetch(V4Run<N1, D1, U8> &r, uint32_t slot32)
{
	using G = V4Geo<N1, D1, U8>;
	constexpr int NCH = V4Run<N1, D1, U8>::NCH;
	static_assert(NCH == 5, "v4_fetch_edge takes five chunks");
	if (r.nf < r.nStages) {
		const uint32_t dst = slot32 + r.laneOff;
		if (r.nf >= r.nLo && r.nf < r.nHi) {
			#pragma unroll
			for (int i = 0; i < NCH; i++) {
				if constexpr (U8)
					cp_async4(dst + 32u * G::CH * (unsigned)i, r.cptr[i] + SOFF * (int)G::kRowBytes);
				else
					cp_async16(dst + 32u * G::CH * (unsigned)i, r.cptr[i] + SOFF * (int)G::kRowBytes);
			}
		} else {
			const long long adv = (long long)r.nf * G::SFR;
			v4_fetch_edge<N1, D1, U8>(dst, r.cptr[0] + SOFF * (int)G::kRowBytes, r.cptr[1] + SOFF * (int)G::kRowBytes,
					r.cptr[2] + SOFF * (int)G::kRowBytes, r.cptr[3] + SOFF * (int)G::kRowBytes, r.cptr[4] + SOFF * (int)G::kRowBytes,
					r.cf0[0] + adv, r.cf0[1] + adv, r.cf0[2] + adv, r.cf0[3] + adv, r.cf0[4] + adv, (long long)r.F, r.src);
		}
		r.nf++;
	}
	cp_async_commit();
}

// One stage of the steady state: stage n has landed, the slot of stage n-1 is handed to the copy of
// stage n + NS - 1, the stage's frames are mixed and applied.  The copies' sources advance with the
// stages (cptr points at the stage being computed), so the offset of the fetch is the constant NS - 1.
template <int N1, int D1, int NS, int CX, bool FIN, bool U8 = false>
__device__ __forceinline__ void v4_step(V4Run<N1, D1, U8> &r, const Lo3Regs &lo, uint32_t tapsStage32, f2_t nz, f2_t &fin)
{
	using G = V4Geo<N1, D1, U8>;
	// stage n has landed (all but the NS-2 youngest groups are complete) ...
	cp_async_wait<NS - 2>();
	__syncwarp();
	// ... and the slot of stage n-1 is free: every lane is past its reads of it
	const uint32_t prev32 = (r.st32 == r.stEnd - r.ringBytes + 0u) ? r.stEnd - G::kStageBytes : r.st32 - G::kStageBytes;
	v4_fetch<N1, D1, NS - 1, U8>(r, prev32);
	v4_stage<N1, D1, CX, FIN, U8>(r.st32, r.acc, r.q, r.qs, lo, tapsStage32, nz, fin, r.sn, r.cs);
	r.q += (uint32_t)G::SFR * r.qs;
	r.st32 += G::kStageBytes;
	if (r.st32 == r.stEnd)
		r.st32 -= r.ringBytes;
	#pragma unroll
	for (int i = 0; i < V4Run<N1, D1, U8>::NCH; i++)
		r.cptr[i] += G::kRowBytes;
}

// The HEAD of a period: stages [SB, SFIN], in which the oldest output still collects taps (and, in
// stage SFIN, completes).  Unrolled: which frames carry the oldest output is a compile-time matter.
template <int N1, int D1, int NS, int SB, bool U8 = false>
__device__ __forceinline__ void v4_head(V4Run<N1, D1, U8> &r, const Lo3Regs &lo, uint32_t taps32, f2_t nz, f2_t &fin)
{
	using G = V4Geo<N1, D1, U8>;
	if constexpr (SB <= G::SFIN) {
		constexpr int CX = (G::REM + 1 - SB * G::SFR) > G::SFR ? G::SFR : (G::REM + 1 - SB * G::SFR);
		v4_step<N1, D1, NS, CX, SB == G::SFIN, U8>(r, lo, taps32 + 4u * (unsigned)(SB * G::TS), nz, fin);
		v4_head<N1, D1, NS, SB + 1, U8>(r, lo, taps32, nz, fin);
	}
}

// The BODY of a period: stages (SFIN, S), all alike but for their taps -- ONE copy of the stage's
// code in a loop (the fully unrolled period was 34 KB of instructions; rolled it is 14 KB at the same
// speed, measured on one box against the unrolled build: 256.6 vs 256.3 us per cfg3 step).
template <int N1, int D1, int NS, bool U8 = false>
__device__ __forceinline__ void v4_body(V4Run<N1, D1, U8> &r, const Lo3Regs &lo, uint32_t taps32, f2_t nz)
{
	using G = V4Geo<N1, D1, U8>;
	f2_t none = 0ull;
	uint32_t t32 = taps32 + 4u * (unsigned)((G::SFIN + 1) * G::TS);
	constexpr int kBodyUnroll = WR_V4_BODY_UNROLL;
	#pragma unroll kBodyUnroll
	for (int sb = G::SFIN + 1; sb < G::S; sb++) {
		v4_step<N1, D1, NS, 0, false, U8>(r, lo, t32, nz, none);
		t32 += 4u * (unsigned)G::TS;
	}
}

template <int N1, int D1, int WMAX, int NS, bool U8 = false>
__global__ void __launch_bounds__(WMAX * 32, 1) chan_kernel_v4(const ChanArgs a, const V4Args v)
{
	using G = V4Geo<N1, D1, U8>;
	using Run = V4Run<N1, D1, U8>;
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
	// per warp: [ring: NS stages (also the prologue's scratch, before it is filled) | taps of two receivers]
	// (raw bytes: the ring proper is a quarter of the size, the region is as large as the scratch needs)
	constexpr unsigned kRingRegion = v4_ring_region(NS * G::kStageBytes, G::kScratchFrames * 8u);
	const unsigned perWarp = kRingRegion + 2u * G::kTapBytes;
	const uint32_t ring32 = smem32 + kV3TableBytes + warp * perWarp;
	const uint32_t tapsA32 = ring32 + kRingRegion;
	bool tableReady = false;
	const unsigned Kmax = v.runLen + (v.longRuns ? 1u : 0u);

	// rounds of this warp slot: 32 consecutive runs each, slots spread over the CTAs
	const unsigned stride = gridDim.x * nWarps * 32u;
	for (unsigned gbase = (warp * gridDim.x + blockIdx.x) * 32u; gbase < v.totalRuns; gbase += stride) {
		// ---- this lane's run: receiver rx, run j of n_r (lanes past the last run shadow lane 0 and store nothing) ----
		const bool live = gbase + lane < v.totalRuns;
		const unsigned rxA = gbase / v.runsPerRx, jA = gbase - rxA * v.runsPerRx;   // lane 0's
		unsigned rx = rxA, j = jA + (live ? lane : 0u);
		if (j >= v.runsPerRx) {                                 // (n_r >= 32: at most one step)
			rx++;
			j -= v.runsPerRx;
		}
		const unsigned lastLive = min(gbase + 31u, v.totalRuns - 1u) - gbase;
		const bool hasB = jA + lastLive >= v.runsPerRx;         // the warp's runs reach into receiver rxA + 1
		const unsigned k0 = j * v.runLen + min(j, v.longRuns);   // first output of the run
		const unsigned Kl = live ? v.runLen + (j < v.longRuns ? 1u : 0u) : 0u;
		const RxConf cf = a.conf[rx];
		const uint32_t ph0 = a.st_in[rx].phase;
		const int32_t step = cf.step;
		const char *__restrict__ src = reinterpret_cast<const char*>(a.iq) + (size_t)cf.stream * a.stream_stride * G::FB;
		const uint32_t taps32 = tapsA32 + (rx - rxA) * G::kTapBytes;
		__syncwarp();

		// ---- the stream: block coordinate c0 = k0*D1 is input frame c0 - (N1-1) ----
		Run r;
		const long long f0 = (long long)k0 * D1 - (N1 - 1);      // first frame of this lane's run (negative: history, reads as zero)
		const unsigned nPeriods = Kmax + (unsigned)AP;           // the last one only up to stage SFIN
		r.nStages = (nPeriods - 1) * S + (unsigned)G::SFIN + 1;
		{
			// stages [nLo, nHi) lie inside the block for EVERY run of the warp
			const unsigned loL = f0 < 0 ? (unsigned)((-f0 + SFR - 1) / SFR) : 0u;
			const long long room = ((long long)a.F - f0) / SFR;
			const unsigned hiL = room <= 0 ? 0u : (unsigned)(room < (long long)r.nStages ? room : (long long)r.nStages);
			r.nLo = __reduce_max_sync(0xFFFFFFFFu, loL);
			r.nHi = __reduce_min_sync(0xFFFFFFFFu, hiL);
		}
		r.src = src;
		r.F = a.F;
		#pragma unroll
		for (int i = 0; i < NCH; i++) {
			// chunk lane + 32 i of a stage's 160: row c / 5, 16-byte column c % 5; it lands at byte 16 c of the slot
			const unsigned c = lane + 32u * (unsigned)i;
			const long long f0r = __shfl_sync(0xFFFFFFFFu, f0, c / NCH);
			const char *srcr = reinterpret_cast<const char*>(__shfl_sync(0xFFFFFFFFu, (unsigned long long)(uintptr_t)src, c / NCH));
			r.cf0[i] = f0r + 2 * (long long)(c % NCH);
			r.cptr[i] = srcr + r.cf0[i] * (long long)G::FB;
		}
		r.nf = 0;
		r.ringBytes = NS * G::kStageBytes;
		r.st32 = ring32 + lane * G::kRowBytes;
		r.stEnd = r.st32 + r.ringBytes;
		r.laneOff = G::CH * lane - lane * G::kRowBytes;

		// ---- the taps of the warp's receiver(s), rows of one period: row a, stage s, frame i <- tap a*D1 + s*SFR + i ----
		#pragma unroll 1
		for (unsigned x = 0; x < (hasB ? 2u : 1u); x++) {
			constexpr int NE = ROWS * S * TS, PER = (NE + 31) / 32;
			const float *tp = a.taps1 + (size_t)(rxA + x) * N1;
			float tv[PER];
			#pragma unroll
			for (int e = 0; e < PER; e++) {
				const unsigned idx = lane + 32u * (unsigned)e;
				const unsigned aa = idx / (S * TS), rest = idx - aa * (S * TS), ss = rest / TS, ii = rest - ss * TS;
				const unsigned jt = aa * D1 + ss * SFR + ii;
				tv[e] = (idx < (unsigned)NE && ii < (unsigned)SFR && jt < (unsigned)N1) ? __ldg(tp + jt) : 0.0f;
			}
			#pragma unroll
			for (int e = 0; e < PER; e++)
				if (lane + 32u * (unsigned)e < (unsigned)NE)
					sts32(tapsA32 + x * G::kTapBytes + 4u * (lane + 32u * (unsigned)e), tv[e]);
		}
		if (!tableReady) {
			mbar_wait(bar32, 0);      // first use of the NCO table: the bulk copies must have landed
			tableReady = true;
		}
		__syncwarp();
		// ---- prologue (the warp that owns a receiver's first run): the outputs whose windows reach into the history ----
		#pragma unroll 1
		for (unsigned x = (jA == 0 ? 0u : 1u); x < (hasB ? 2u : 1u); x++) {
			const unsigned rxp = rxA + x;
			const uint32_t tapsP32 = tapsA32 + x * G::kTapBytes;
			const uint32_t php = a.st_in[rxp].phase;
			const RxConf cfp = a.conf[rxp];
			const char *__restrict__ srcp = reinterpret_cast<const char*>(a.iq) + (size_t)cfp.stream * a.stream_stride * G::FB;
			// [history (N1-1) | mixed frames 0 ...], float2 each; four entries per lane and round, loads first
			for (unsigned c0 = 0; c0 < G::kScratchFrames; c0 += 128) {
				float2 xx[4];
				#pragma unroll
				for (int u = 0; u < 4; u++) {
					const unsigned c = c0 + lane + 32u * (unsigned)u;
					if (c < (unsigned)(N1 - 1))
						xx[u] = a.hist_in[(size_t)rxp * (N1 - 1) + c];
					else if (c < G::kScratchFrames && c - (unsigned)(N1 - 1) < a.F)
						xx[u] = v4_load_frame<U8>(srcp, (long long)(c - (unsigned)(N1 - 1)), lo.hi);
					else
						xx[u] = make_float2(0.0f, 0.0f);
				}
				#pragma unroll
				for (int u = 0; u < 4; u++) {
					const unsigned c = c0 + lane + 32u * (unsigned)u;
					if (c >= (unsigned)(N1 - 1) && c < G::kScratchFrames) {
						const unsigned f = c - (unsigned)(N1 - 1);
						float sn, cs;
						lo3_sincos(((php + f * (uint32_t)cfp.step) << 1) + 0x80000000u, lo, sn, cs);
						xx[u] = mix(xx[u], cs, sn);
					}
					if (c < G::kScratchFrames)
						sts64(ring32 + 8u * c, xx[u]);
				}
			}
			__syncwarp();
			if (lane < (unsigned)G::KSKIP && lane < a.M1) {
				f2_t acc = 0ull;
				uint32_t x32 = ring32 + 8u * lane * (unsigned)D1;
				#pragma unroll 1
				for (int aa = 0; aa < ROWS; aa++) {
					#pragma unroll 1
					for (int ss = 0; ss < S; ss++) {
						const uint32_t t32 = tapsP32 + 4u * (unsigned)((aa * S + ss) * TS);
						#pragma unroll
						for (int ii = 0; ii < SFR; ii++)
							if (aa * D1 + ss * SFR + ii < N1)
								acc = tap3(acc, lds32(t32 + 4u * (unsigned)ii), lds64p(x32 + 8u * (unsigned)ii), nz);
						x32 += 8u * (unsigned)SFR;
					}
				}
				float2 y;
				f2_unpack(acc, y.x, y.y);
				a.chan[(size_t)rxp * a.chan_stride + lane] = y;
			}
			__syncwarp();
		}
		// ---- fill the ring: stages 0 .. NS-2 of period 0 (v4_fetch takes its sources relative to the
		// period being computed, SOFF = 0 .. NS-2 is exactly where cptr points) ----
		v4_fetch<N1, D1, 0, U8>(r, ring32 + lane * G::kRowBytes);
		if (NS > 2) v4_fetch<N1, D1, 1, U8>(r, ring32 + G::kStageBytes + lane * G::kRowBytes);
		if (NS > 3) v4_fetch<N1, D1, 2, U8>(r, ring32 + 2 * G::kStageBytes + lane * G::kRowBytes);

		#pragma unroll
		for (int i = 0; i < ROWS; i++)
			r.acc[i] = 0ull;
		r.q = ((ph0 + (uint32_t)(int32_t)f0 * (uint32_t)step) << 1) + 0x80000000u;
		r.qs = 2u * (uint32_t)step;
#if WR_V4_AHEAD
		{
			uint32_t q[SFR];
			#pragma unroll
			for (int i = 0; i < SFR; i++)
				q[i] = r.q + (uint32_t)i * r.qs;
			lo3_sincos_n<SFR>(q, lo, r.sn, r.cs);
		}
#endif
		float2 *out = a.chan + (size_t)rx * a.chan_stride;
		unsigned kdone = k0 - (unsigned)AP;                     // the output that completes in the current period (wraps below zero at first)
		for (unsigned p = 0; p < nPeriods; p++) {
			// a period begins: every live output is one period older, a new one starts
			#pragma unroll
			for (int i = ROWS - 1; i > 0; i--)
				r.acc[i] = r.acc[i - 1];
			r.acc[0] = 0ull;
			f2_t fin = 0ull;
			v4_head<N1, D1, NS, 0, U8>(r, lo, taps32, nz, fin);
			if (p + 1 < nPeriods)
				v4_body<N1, D1, NS, U8>(r, lo, taps32, nz);
			// the output that began AP periods ago is complete (the first KSKIP of a receiver are the prologue's)
			if (kdone - k0 < Kl && kdone >= (unsigned)G::KSKIP) {
				float2 y;
				f2_unpack(fin, y.x, y.y);
				out[kdone] = y;
			}
			kdone++;
		}
		cp_async_wait<0>();
		// ---- epilogue (the warp that owns a receiver's first run): the carried state of the next block ----
		#pragma unroll 1
		for (unsigned x = (jA == 0 ? 0u : 1u); x < (hasB ? 2u : 1u); x++) {
			const unsigned rxp = rxA + x;
			const uint32_t php = a.st_in[rxp].phase;
			const RxConf cfp = a.conf[rxp];
			const char *__restrict__ srcp = reinterpret_cast<const char*>(a.iq) + (size_t)cfp.stream * a.stream_stride * G::FB;
			for (unsigned i0 = 0; i0 < (unsigned)(N1 - 1); i0 += 64) {
				// the last N1-1 mixed frames of [history | block]
				float2 xx[2];
				#pragma unroll
				for (int u = 0; u < 2; u++) {
					const unsigned i = i0 + lane + 32u * (unsigned)u;
					const long long f = (long long)a.F - (N1 - 1) + i;
					if (i >= (unsigned)(N1 - 1))
						xx[u] = make_float2(0.0f, 0.0f);
					else if (f < 0)
						xx[u] = a.hist_in[(size_t)rxp * (N1 - 1) + (unsigned)(f + (N1 - 1))];
					else
						xx[u] = v4_load_frame<U8>(srcp, f, lo.hi);
				}
				#pragma unroll
				for (int u = 0; u < 2; u++) {
					const unsigned i = i0 + lane + 32u * (unsigned)u;
					const long long f = (long long)a.F - (N1 - 1) + i;
					if (i < (unsigned)(N1 - 1)) {
						if (f >= 0) {
							float sn, cs;
							lo3_sincos(((php + (uint32_t)f * (uint32_t)cfp.step) << 1) + 0x80000000u, lo, sn, cs);
							xx[u] = mix(xx[u], cs, sn);
						}
						a.hist_out[(size_t)rxp * (N1 - 1) + i] = xx[u];
					}
				}
			}
			if (lane == 0)
				a.st_out[rxp].phase = phase_at(php, cfp.step, a.F);
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

#ifndef WR_V4_NS
#define WR_V4_NS 3
#endif
constexpr int kV4Stages = WR_V4_NS;   // depth of a warp's ring of stages: the copies run kV4Stages - 1 stages ahead
constexpr int kV4Warps = 8;           // warps per CTA: two per scheduler (up to 255 registers per thread), ring of three stages

struct V4Plan {
	bool ok = false;
	V4Kernel kernel = nullptr;
	int regs = 0;
	int numSMs = 0;
	unsigned n1 = 0, d1 = 0;
	unsigned kskip = 0;          // outputs the prologue computes
	unsigned ap = 0;             // periods a run spends beyond its own outputs
	V4Kernel kernelU8 = nullptr; // the same kernel fed raw RTL-SDR bytes
	unsigned ringBytes = 0, ringBytesU8 = 0, tapBytes = 0;   // a warp's ring region (float / bytes), one tap set
	size_t smemMax = 0;
	unsigned maxPerStream = 1;   // most receivers that share one tuner stream (v4_set_groups)
	unsigned runs = 0;           // WR_V4_RUNS: runs per receiver (0 = fill the grid; 32 = one receiver per warp)
	unsigned kmin = 0;           // WR_V4_KMIN: shortest run (0 = default)
	bool pdl = true;
};

template <int N1, int D1>
inline void v4_fill(V4Plan &p)
{
	using G = V4Geo<N1, D1>;
	p.kernel = chan_kernel_v4<N1, D1, kV4Warps, kV4Stages>;
	p.kernelU8 = chan_kernel_v4<N1, D1, kV4Warps, kV4Stages, true>;
	p.kskip = G::KSKIP;
	p.ap = G::AP;
	p.ringBytes = v4_ring_region(kV4Stages * G::kStageBytes, G::kScratchFrames * 8u);
	p.ringBytesU8 = v4_ring_region(kV4Stages * V4Geo<N1, D1, true>::kStageBytes, G::kScratchFrames * 8u);
	p.tapBytes = G::kTapBytes;
}

// the geometries v4 is instantiated for (the long filters of the BASELINE configs)
inline bool v4_pick(V4Plan &p, unsigned n1, unsigned d1)
{
	p.n1 = n1;
	p.d1 = d1;
	if (n1 == 255 && d1 == 50) v4_fill<255, 50>(p);
	else if (n1 == 127 && d1 == 50) v4_fill<127, 50>(p);
	else if (n1 == 127 && d1 == 40) v4_fill<127, 40>(p);
	else return false;
	if (const char *e = getenv("WR_V4_RUNS"))
		p.runs = (unsigned)std::max(0, atoi(e));
	if (const char *e = getenv("WR_V4_KMIN"))
		p.kmin = (unsigned)std::max(0, atoi(e));
	return true;
}

// v4 needs the v3 plan's table.
inline int v4_init(V4Plan &p, const V3Plan &v3, int device, unsigned n1, unsigned d1)
{
	p.ok = false;
	if (!v4_pick(p, n1, d1) || !v3.d_delta)
		return WR_OK;
	if (const char *e = getenv("WR_V3_PDL"))
		p.pdl = atoi(e) != 0;
	cudaDeviceProp prop;
	WR_CUDA(cudaGetDeviceProperties(&prop, device));
	p.numSMs = prop.multiProcessorCount;
	cudaFuncAttributes fa;
	WR_CUDA(cudaFuncGetAttributes(&fa, p.kernel));
	p.regs = fa.numRegs;
	// the opt-in limit covers static and dynamic shared memory together (the kernel's mbarrier is static)
	p.smemMax = prop.sharedMemPerBlockOptin - ((fa.sharedSizeBytes + 255) & ~(size_t)255);
	WR_CUDA(cudaFuncSetAttribute(p.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smemMax));
	WR_CUDA(cudaFuncSetAttribute(p.kernelU8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smemMax));
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
	unsigned warps;
	unsigned runsPerRx, runLen, longRuns;   // n_r runs per receiver: `longRuns` of runLen + 1 outputs, then runLen
	unsigned rounds;                        // rounds of 32 runs the busiest warp slot works through
	unsigned grid;
};

// How R receivers x M1 outputs are cut into runs for a grid of numSMs x warps x 32 lanes: as
// many runs per receiver as ONE round of the grid holds (more rounds only when the bank has more
// receivers than that leaves room for 32 runs each), never shorter than kmin outputs.
// A run of K outputs costs K + ap periods, so fewer, longer runs are cheaper per output -- but
// a grid whose lanes are not all busy wastes more: the kernel's time is rounds * (K + ap).
inline bool v4_cut(const V4Plan &p, unsigned R, unsigned M1, unsigned warps, V4Launch *out)
{
	const unsigned kmin = std::max(p.kskip + 1, p.kmin ? p.kmin : 2 * (p.ap + 1));
	const unsigned nrMax = M1 / kmin;
	if (!R || nrMax < 32u)
		return false;
	const unsigned long long lanes = (unsigned long long)p.numSMs * warps * 32u;
	unsigned nr = p.runs;
	if (nr == 0) {
		unsigned long long q = 1;
		while ((q * lanes) / R < 32u)
			q++;
		nr = (unsigned)std::min<unsigned long long>(nrMax, (q * lanes) / R);
	}
	nr = std::min(std::max(nr, 32u), nrMax);
	const unsigned long long total = (unsigned long long)R * nr;
	if (total > 0x7FFFFFFFull)
		return false;
	const unsigned long long warpRounds = (total + 31) / 32;
	out->warps = warps;
	out->runsPerRx = nr;
	out->runLen = M1 / nr;
	out->longRuns = M1 % nr;
	out->grid = (unsigned)std::min<unsigned long long>((unsigned long long)p.numSMs, (warpRounds + warps - 1) / warps);
	out->rounds = (unsigned)((warpRounds + (unsigned long long)out->grid * warps - 1) / ((unsigned long long)out->grid * warps));
	return true;
}

// How a block of F frames would be launched; false if v4 does not serve it.
inline bool v4_shape(const V4Plan &p, const V3Plan &v3, unsigned R, unsigned F, const void *iq, bool u8, size_t stream_stride, bool forced, V4Launch *out)
{
	if (!p.ok || !v3.ok)          // (v3.ok: the table survived the compression)
		return false;
	// runs must start on the copies' boundaries: 16 bytes (two float frames), 4 bytes (two frames of raw bytes)
	if (((uintptr_t)iq & (u8 ? 3u : 15u)) || (stream_stride & 1u))
		return false;
	const unsigned M1 = F / p.d1;
	if (F < p.n1 - 1 || M1 < 32u * 2u * p.kskip)
		return false;
	unsigned warps = kV4Warps;
	const size_t avail = p.smemMax - kV3TableBytes;
	const size_t perWarp = (size_t)(u8 ? p.ringBytesU8 : p.ringBytes) + 2 * (size_t)p.tapBytes;
	while (warps > 1 && perWarp * warps > avail)
		warps--;
	if (perWarp * warps > avail || !v4_cut(p, R, M1, warps, out))
		return false;
	// not enough work for a persistent grid of independent warps: a warp per SM at least
	if ((unsigned long long)R * out->runsPerRx < 32ull * (unsigned)p.numSMs && !forced)
		return false;
	// Shared tuners (64 receivers per stream: cfg2, cfg5): every run re-reads its tuner stream -- from
	// L2, the receivers of a stream walk it together -- where v3's mixers share the raw registers among
	// the receivers of a stream.  v4 wins once the bank fills the WHOLE grid with long runs (cfg5:
	// 1024 receivers x 10240 outputs, runs of 277, 0.84 -> 0.73 ms per block); a small bank (cfg2:
	// 64 x 2048 outputs = 3.5 per lane) stays with v3.
	if (p.maxPerStream > 2 && !forced && (out->grid < (unsigned)p.numSMs || out->runLen < 8u * p.ap))
		return false;
	return true;
}

inline int v4_launch_chan(V4Plan &p, const V3Plan &v3, const V4Launch &L, ChanArgs &ca, unsigned R, bool u8, cudaStream_t st, unsigned long long *launches)
{
	V4Args v;
	v.delta = v3.d_delta;
	v.eps = v3.coef.eps;
	v.negzero = -0.0f;
	v.prmtHi = 0x4B00u;
	v.runsPerRx = L.runsPerRx;
	v.runLen = L.runLen;
	v.longRuns = L.longRuns;
	v.totalRuns = R * L.runsPerRx;
	v.R = R;
	const size_t smem = kV3TableBytes + (size_t)L.warps * ((size_t)(u8 ? p.ringBytesU8 : p.ringBytes) + 2 * (size_t)p.tapBytes);
	cudaLaunchConfig_t cfg = {};
	cudaLaunchAttribute attr[1];
	cfg.gridDim = dim3(L.grid);
	cfg.blockDim = dim3(L.warps * 32);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = st;
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = p.pdl ? 1 : 0;
	cudaError_t e = cudaLaunchKernelEx(&cfg, u8 ? p.kernelU8 : p.kernel, (const ChanArgs)ca, (const V4Args)v);
	(*launches)++;
	if (e == cudaSuccess)
		e = cudaGetLastError();
	if (e != cudaSuccess) {
		wr::set_error("chan_kernel_v4 launch (grid %u, %u warps, %zu bytes of shared memory): %s",
				L.grid, L.warps, smem, cudaGetErrorString(e));
		return WR_ECUDA;
	}
	return WR_OK;
}

} // namespace wrd
