// wr_kernels_v3.cuh -- fused NCO mix + decimating channel FIR (K1+K2 of SURVEY.md 2a), third
// generation: a STREAMING RING of mixed slots instead of one tile per work item.
//
// What v2 measured: its per-item set-up was 21% of all instructions, the mixers spent 21% of
// their time waiting for a tile buffer and 18% waiting for raw loads, and the consumer's single
// dependent FADD2 chain per tile took ~5100 clk because the warp scheduler is fair -- a
// latency-bound warp gets one issue slot per round like everyone else, so a role that needs 5x
// the instructions of the others per hand-over is the critical path.
//
// Shape of v3.  Persistent grid, one 512-thread CTA per SM.  A CTA owns a contiguous range of
// "units" (receiver group, pass): a pass is SF = 32*NG*d1 consecutive frames of the block
// (~3200), i.e. GO = 32*NG channel-rate outputs per receiver.
//   * The NCO correction table (wr_lo3.h, 132 KiB) is staged into shared memory by bulk copies
//     (TMA) that complete on an mbarrier, behind the first group's set-up.
//   * 10 MIXER warps (320 threads, J = SF/320 frames each).  The raw IQ of a pass is held in
//     registers and shared by the <= RB receivers of the group that listen to that stream; a
//     pass is mixed in two halves, and while the last receiver is mixed each half's registers
//     are refilled with the next pass's frames.  Per (receiver, pass) the mixers fill one of S
//     ring slots in shared memory: [A periods of halo | GO periods], period-major.  The halo --
//     the last A*d1 mixed frames of the receiver's previous pass -- never leaves the mixer
//     threads that produced it: they keep it in registers and store it again into the next
//     slot, so nothing is mixed twice and slots are self-contained.
//   * S*2 FIR warps, two per slot, each taking every second group of 32 outputs: one thread per
//     output over the taps in the reference's order (packed f32x2, products and sums rounded
//     separately), samples two per 128-bit load, taps four per 128-bit load and applied through
//     the broadcast operand form.
//   * Slots are handed over with mbarriers (full/empty per slot); a small descriptor per slot
//     tells the FIR warps which receiver and which pass it holds.
//   * Tuner blocks arrive as float IQ or as raw RTL-SDR bytes (converted in the load path).
//   * Launched with programmatic stream serialization: the whole kernel may run under the previous
//     kernel in the stream (the demodulator of the block before, which reads the other side of the
//     double-buffered channel-rate stream); it only waits for it before it exits.
// NCO: the sine and cosine table entries of a frame are reconstructed TOGETHER in the two halves
// of packed f32x2 registers (wr_lo3.h): one byte permute per entry builds the float index, ten
// packed instructions produce both bases and both padded table positions; written stage by stage
// across the frames of a half pass so that the scheduler sees independent chains.
#pragma once

#include "wr_bank.cuh"
#include "wr_device.cuh"
#include "wr_kernels_v2.cuh"   // bar_sync / bar_arrive / mul2_rn_exact
#include "wr_lo3.h"

#include <algorithm>
#include <vector>

namespace wrd {

constexpr int kV3Mixers = 320;                 // mixer threads (10 warps)
constexpr int kV3Slots = 3;                    // ring slots
constexpr int kV3FirPerSlot = 2;               // FIR warps that share a slot (each takes every second group of 32 outputs)
constexpr int kV3Fir = kV3Slots * kV3FirPerSlot;
constexpr int kV3Threads = kV3Mixers + 32 * kV3Fir;
constexpr unsigned kV3TableBytes = (((unsigned)(WR_LO3_SLOT_MAX - WR_LO3_SLOT_MIN + 1) * 2u) + 15u) & ~15u;
constexpr unsigned kV3MidOffset = (unsigned)(-(WR_LO3_SLOT_MIN)) * 2u;   // byte offset of slot 0

constexpr int kV3BarMix = 1;   // named barrier of the mixer warps (0 is __syncthreads)

constexpr int v3_pad_period(int d)
{
	// period-major layout: the FIR's lanes step DP float2 apart and load 16 bytes each, so a
	// quarter-warp covers all 32 banks iff 2*DP = 4 (mod 8) words, i.e. DP = 2 (mod 4)
	int dp = d;
	while (dp % 4 != 2)
		dp += 1;
	return dp;
}

// Output groups (of 32) per slot: the largest even count that keeps a pass at or below ~3200
// frames AND gives every mixer thread an even number of frames (a pass is mixed in two halves).
constexpr int v3_groups(int d)
{
	int best = 0;
	for (int ng = 2; 32 * ng * d <= 3200; ng += 2)
		if ((32 * ng * d) % (2 * kV3Mixers) == 0)
			best = ng;
	return best;
}

template <int N1, int D1>
struct V3Geo {
	static constexpr int A = (N1 - 1 + D1 - 1) / D1;       // periods of halo
	static constexpr int OFF = A * D1 - (N1 - 1);          // offset of an output's first tap in its period
	static constexpr int DP = v3_pad_period(D1);
	// output groups (of 32) per slot: passes of ~3200 frames, so that a mixer thread owns 8-10
	// frames per pass and the per-pass bookkeeping is spread over twice as many frames as with
	// one group per slot
	static constexpr int NG = v3_groups(D1);
	static_assert(NG >= 2, "no pass length fits this decimation");
	static constexpr int GO = 32 * NG;                     // outputs per slot
	static constexpr int SF = GO * D1;                     // frames per pass
	static constexpr int J = SF / kV3Mixers;               // frames per mixer thread
	static constexpr int HF = A * D1;                      // halo frames
	static constexpr int SLOT = (A + GO) * DP;             // float2 entries of one ring slot
	static constexpr int TAPOFF = (4 - (OFF & 1)) & 3;     // taps stored shifted so that groups of four are 16-byte aligned
	static constexpr unsigned kTapsStride = (((unsigned)(N1 + 3) * 4u) + 15u) & ~15u;
	static_assert(D1 % 2 == 0, "v3 needs an even decimation (128-bit sample loads)");
	static_assert(SF % kV3Mixers == 0, "pass length must be a multiple of the mixer thread count");
	static_assert(HF <= kV3Mixers, "halo must fit the last frame of each mixer thread");
	static_assert(HF <= SF, "a pass must be at least as long as the halo");
	static_assert(N1 <= kV3Mixers, "taps are staged one per mixer thread");
	static_assert(J % 2 == 0, "a pass is mixed in two halves");
	static_assert(NG % kV3FirPerSlot == 0, "the FIR warps of a slot split its groups evenly");
};

struct V3Args {
	const int16_t *delta;     // padded corrections (wr_lo3.h), staged to shared memory per CTA
	float eps;
	const unsigned *order;    // receivers sorted by stream
	const int4 *groups;       // {first index into order, count, stream, unused}
	unsigned nGroups;
	unsigned P;               // passes per block: ceil(F / SF)
	float negzero;            // -0.0f, opaque to the compiler (mul2_rn_exact)
	unsigned prmtHi;          // 0x4B00, opaque to the compiler so that it stays in a register (lo3_sincos)
};

// ---- packed f32x2 helpers -------------------------------------------------------------------
typedef unsigned long long f2_t;

__device__ __forceinline__ f2_t f2_pack(float lo, float hi)
{
	f2_t r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
	return r;
}

__device__ __forceinline__ void f2_unpack(f2_t v, float &lo, float &hi)
{
	asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

__device__ __forceinline__ f2_t f2_fma(f2_t a, f2_t b, f2_t c)
{
	f2_t r;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
	return r;
}

__device__ __forceinline__ f2_t f2_mul(f2_t a, f2_t b)
{
	f2_t r;
	asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
	return r;
}

// -a in both halves; ptxas folds it into the operand modifier of the packed instruction that uses it
__device__ __forceinline__ f2_t f2_neg(f2_t a)
{
	float lo, hi;
	asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a));
	lo = -lo;
	hi = -hi;
	f2_t d;
	asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
	return d;
}

__device__ __forceinline__ f2_t f2_add(f2_t a, f2_t b)
{
	f2_t r;
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
	return r;
}

// ---- mbarrier hand-over ---------------------------------------------------------------------
// Slots change hands through mbarriers rather than named barriers: a named barrier makes every
// mixer warp wait for the slowest one at each slot, an mbarrier lets each warp run up to a ring
// ahead of the others (a phase completes when all expected arrivals are in; waiting does not
// arrive).  Arrive is a release, a successful wait an acquire (PTX defaults, CTA scope).
__device__ __forceinline__ void mbar_init(uint32_t bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, unsigned parity)
{
	asm volatile("{\n\t.reg .pred p;\n"
			"WR_MBAR_WAIT%=:\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
			"@!p bra WR_MBAR_WAIT%=;\n\t}" :: "r"(bar), "r"(parity) : "memory");
}

// The FIR warps' wait for a full slot.  They are the role with slack, and they spin on schedulers
// the mixers share: on cfg3 the loop runs 4.8 M times per launch, 7 % of the kernel's executed
// warp-instructions (profiles/r01_ncu_hotspots_chan_cfg3.txt).  Experiment for the next GPU visit
// (make lib-exp, then WEBRADIO_B200_LIB=webradio_b200/libwebradio_b200_exp.so): back off with nanosleep between attempts.  Without the
// macro this is mbar_wait itself -- the default build's code is unchanged.
#ifdef WR_EXP_FIR_SLEEP
__device__ __forceinline__ void mbar_wait_fir(uint32_t bar, unsigned parity)
{
	asm volatile("{\n\t.reg .pred p;\n"
			"WR_MBAR_WAITS%=:\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
			"@p bra WR_MBAR_DONE%=;\n\t"
			"nanosleep.u32 %2;\n\t"
			"bra WR_MBAR_WAITS%=;\n"
			"WR_MBAR_DONE%=:\n\t}" :: "r"(bar), "r"(parity), "n"(WR_EXP_FIR_SLEEP) : "memory");
}
#else
__device__ __forceinline__ void mbar_wait_fir(uint32_t bar, unsigned parity) { mbar_wait(bar, parity); }
#endif

// Packed constants of the table reconstruction, built once per thread.
struct Lo3Regs {
	f2_t tscale, tbias, slotk, slotm, eps, c0, c1, c2;
#if WR_LO3_DEGREE == 3
	f2_t c3;
#endif
	uint32_t cbase;   // shared-space address of slot 0 minus 2 * WR_LO3_SLOTBITS (mod 2^32)
	uint32_t hi;      // 0x4B00: exponent bytes of the float index
};

__device__ __forceinline__ Lo3Regs lo3_regs(float eps, uint32_t dmid32, uint32_t hi)
{
	Lo3Regs k;
	k.tscale = f2_pack(WR_LO3_TSCALE, WR_LO3_TSCALE);
	k.tbias = f2_pack(WR_LO3_TBIAS, WR_LO3_TBIAS);
	k.slotk = f2_pack(WR_LO3_SLOTK, WR_LO3_SLOTK);
	k.slotm = f2_pack(WR_LO3_SLOTM, WR_LO3_SLOTM);
	k.eps = f2_pack(eps, eps);
	k.c0 = f2_pack(WR_LO3_C0, WR_LO3_C0);
	k.c1 = f2_pack(WR_LO3_C1, WR_LO3_C1);
	k.c2 = f2_pack(WR_LO3_C2, WR_LO3_C2);
#if WR_LO3_DEGREE == 3
	k.c3 = f2_pack(WR_LO3_C3, WR_LO3_C3);
#endif
	k.cbase = dmid32 - 2u * (uint32_t)WR_LO3_SLOTBITS;
	k.hi = hi;
	return k;
}

// sin/cos of the NCO for the BIASED doubled phase qb = (phase << 1) + 0x80000000: exactly the
// reference's sinTable[sinidx], sinTable[cosidx] (downconverter.cxx:100-102).  The bias turns
// the signed table index into the low 16 bits of a float's mantissa with one byte permute.
__device__ __forceinline__ void lo3_sincos(uint32_t qb, const Lo3Regs &k, float &sn, float &cs)
{
	uint32_t fs, fc;
	const uint32_t qc = qb + 0x40000000u;      // + a quarter turn
	asm("prmt.b32 %0, %1, %2, 0x5432;" : "=r"(fs) : "r"(qb), "r"(k.hi));
	asm("prmt.b32 %0, %1, %2, 0x5432;" : "=r"(fc) : "r"(qc), "r"(k.hi));
	const f2_t F = f2_pack(__uint_as_float(fs), __uint_as_float(fc));
	const f2_t SL = f2_fma(F, k.slotk, k.slotm);
	const f2_t T = f2_fma(F, k.tscale, k.tbias);
	const f2_t Y = f2_mul(T, T);
	const f2_t Z = f2_fma(f2_neg(Y), T, T);        // t - t^3
	const f2_t U = f2_fma(T, k.eps, Z);
#if WR_LO3_DEGREE == 3
	f2_t P = f2_fma(Y, k.c3, k.c2);
	P = f2_fma(Y, P, k.c1);
#else
	f2_t P = f2_fma(Y, k.c2, k.c1);
#endif
	P = f2_fma(Y, P, k.c0);
	const f2_t B = f2_mul(U, P);
	float sls, slc, bs, bc;
	f2_unpack(SL, sls, slc);
	f2_unpack(B, bs, bc);
	int ds, dc;
	asm("ld.shared.s16 %0, [%1];" : "=r"(ds) : "r"(k.cbase + 2u * __float_as_uint(sls)));
	asm("ld.shared.s16 %0, [%1];" : "=r"(dc) : "r"(k.cbase + 2u * __float_as_uint(slc)));
	sn = __int_as_float(__float_as_int(bs) + ds);
	cs = __int_as_float(__float_as_int(bc) + dc);
}

// The same for J independent phases, written stage by stage across the J frames so that the
// instruction stream offers J independent dependency chains to the scheduler (a mixer SMSP holds
// only three or four warps; per-frame code interleaves two chains at best).
template <int J>
__device__ __forceinline__ void lo3_sincos_n(const uint32_t (&qb)[J], const Lo3Regs &k, float (&sn)[J], float (&cs)[J])
{
	f2_t F[J], T[J], Y[J], W[J], P[J];             // (W: t - t^3, then the eps term)
	uint32_t as[J], ac[J];
	int ds[J], dc[J];
	#pragma unroll
	for (int j = 0; j < J; j++) {
		uint32_t fs, fc;
		const uint32_t qc = qb[j] + 0x40000000u;      // + a quarter turn
		asm("prmt.b32 %0, %1, %2, 0x5432;" : "=r"(fs) : "r"(qb[j]), "r"(k.hi));
		asm("prmt.b32 %0, %1, %2, 0x5432;" : "=r"(fc) : "r"(qc), "r"(k.hi));
		F[j] = f2_pack(__uint_as_float(fs), __uint_as_float(fc));
	}
	#pragma unroll
	for (int j = 0; j < J; j++) {
		float sls, slc;
		f2_unpack(f2_fma(F[j], k.slotk, k.slotm), sls, slc);
		as[j] = k.cbase + 2u * __float_as_uint(sls);
		ac[j] = k.cbase + 2u * __float_as_uint(slc);
	}
	#pragma unroll
	for (int j = 0; j < J; j++) {
		asm volatile("ld.shared.s16 %0, [%1];" : "=r"(ds[j]) : "r"(as[j]));
		asm volatile("ld.shared.s16 %0, [%1];" : "=r"(dc[j]) : "r"(ac[j]));
	}
	#pragma unroll
	for (int j = 0; j < J; j++)
		T[j] = f2_fma(F[j], k.tscale, k.tbias);
	#pragma unroll
	for (int j = 0; j < J; j++)
		Y[j] = f2_mul(T[j], T[j]);
#if WR_LO3_DEGREE == 3
	#pragma unroll
	for (int j = 0; j < J; j++) {
		W[j] = f2_fma(f2_neg(Y[j]), T[j], T[j]);
		P[j] = f2_fma(Y[j], k.c3, k.c2);
	}
	#pragma unroll
	for (int j = 0; j < J; j++) {
		W[j] = f2_fma(T[j], k.eps, W[j]);
		P[j] = f2_fma(Y[j], P[j], k.c1);
	}
	#pragma unroll
	for (int j = 0; j < J; j++)
		P[j] = f2_fma(Y[j], P[j], k.c0);
#else
	#pragma unroll
	for (int j = 0; j < J; j++) {
		W[j] = f2_fma(f2_neg(Y[j]), T[j], T[j]);
		P[j] = f2_fma(Y[j], k.c2, k.c1);
	}
	#pragma unroll
	for (int j = 0; j < J; j++) {
		W[j] = f2_fma(T[j], k.eps, W[j]);
		P[j] = f2_fma(Y[j], P[j], k.c0);
	}
#endif
	#pragma unroll
	for (int j = 0; j < J; j++) {
		float bs, bc;
		f2_unpack(f2_mul(W[j], P[j]), bs, bc);
		sn[j] = __int_as_float(__float_as_int(bs) + ds[j]);
		cs[j] = __int_as_float(__float_as_int(bc) + dc[j]);
	}
}

// lo3_sincos_n in two halves, for a caller that puts other work between them: the table lookups are
// random 16-bit gathers (two per frame, ~2.6 cycles of the SM's one shared-memory pipe each), and with
// eight warps gathering they queue -- issued a whole FIR phase ahead of their use, nobody waits for them.
template <int J>
__device__ __forceinline__ void lo3_issue_n(const uint32_t (&qb)[J], const Lo3Regs &k, f2_t (&F)[J], int (&ds)[J], int (&dc)[J])
{
	uint32_t as[J], ac[J];
	#pragma unroll
	for (int j = 0; j < J; j++) {
		uint32_t fs, fc;
		const uint32_t qc = qb[j] + 0x40000000u;      // + a quarter turn
		asm("prmt.b32 %0, %1, %2, 0x5432;" : "=r"(fs) : "r"(qb[j]), "r"(k.hi));
		asm("prmt.b32 %0, %1, %2, 0x5432;" : "=r"(fc) : "r"(qc), "r"(k.hi));
		F[j] = f2_pack(__uint_as_float(fs), __uint_as_float(fc));
	}
	#pragma unroll
	for (int j = 0; j < J; j++) {
		float sls, slc;
		f2_unpack(f2_fma(F[j], k.slotk, k.slotm), sls, slc);
		as[j] = k.cbase + 2u * __float_as_uint(sls);
		ac[j] = k.cbase + 2u * __float_as_uint(slc);
	}
	#pragma unroll
	for (int j = 0; j < J; j++) {
		asm volatile("ld.shared.s16 %0, [%1];" : "=r"(ds[j]) : "r"(as[j]));
		asm volatile("ld.shared.s16 %0, [%1];" : "=r"(dc[j]) : "r"(ac[j]));
	}
}

template <int J>
__device__ __forceinline__ void lo3_finish_n(const f2_t (&F)[J], const int (&ds)[J], const int (&dc)[J], const Lo3Regs &k,
		float (&sn)[J], float (&cs)[J])
{
	f2_t T[J], Y[J], W[J], P[J];
	#pragma unroll
	for (int j = 0; j < J; j++)
		T[j] = f2_fma(F[j], k.tscale, k.tbias);
	#pragma unroll
	for (int j = 0; j < J; j++)
		Y[j] = f2_mul(T[j], T[j]);
#if WR_LO3_DEGREE == 3
	#pragma unroll
	for (int j = 0; j < J; j++) {
		W[j] = f2_fma(f2_neg(Y[j]), T[j], T[j]);
		P[j] = f2_fma(Y[j], k.c3, k.c2);
	}
	#pragma unroll
	for (int j = 0; j < J; j++) {
		W[j] = f2_fma(T[j], k.eps, W[j]);
		P[j] = f2_fma(Y[j], P[j], k.c1);
	}
	#pragma unroll
	for (int j = 0; j < J; j++)
		P[j] = f2_fma(Y[j], P[j], k.c0);
#else
	#pragma unroll
	for (int j = 0; j < J; j++) {
		W[j] = f2_fma(f2_neg(Y[j]), T[j], T[j]);
		P[j] = f2_fma(Y[j], k.c2, k.c1);
	}
	#pragma unroll
	for (int j = 0; j < J; j++) {
		W[j] = f2_fma(T[j], k.eps, W[j]);
		P[j] = f2_fma(Y[j], P[j], k.c0);
	}
#endif
	#pragma unroll
	for (int j = 0; j < J; j++) {
		float bs, bc;
		f2_unpack(f2_mul(W[j], P[j]), bs, bc);
		sn[j] = __int_as_float(__float_as_int(bs) + ds[j]);
		cs[j] = __int_as_float(__float_as_int(bc) + dc[j]);
	}
}

__device__ __forceinline__ void sts64(uint32_t addr, float2 v)
{
	asm volatile("st.shared.v2.f32 [%0], {%1, %2};" :: "r"(addr), "f"(v.x), "f"(v.y) : "memory");
}

__device__ __forceinline__ void sts32(uint32_t addr, float v)
{
	asm volatile("st.shared.f32 [%0], %1;" :: "r"(addr), "f"(v) : "memory");
}

__device__ __forceinline__ float2 lds64(uint32_t addr)
{
	float2 v;
	asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
	return v;
}

__device__ __forceinline__ float lds32(uint32_t addr)
{
	float v;
	asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
	return v;
}

__device__ __forceinline__ f2_t lds64p(uint32_t addr)
{
	f2_t v;
	asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr));
	return v;
}

__device__ __forceinline__ void lds128p(uint32_t addr, f2_t &a, f2_t &b)
{
	asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}

__device__ __forceinline__ float4 lds128f(uint32_t addr)
{
	float4 v;
	asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
	return v;
}

__device__ __forceinline__ f2_t ldg64p(const float2 *p)
{
	f2_t v;
	asm volatile("ld.global.nc.b64 %0, [%1];" : "=l"(v) : "l"(p));
	return v;
}

// position (in float2 entries from the slot base) of the frame v frames after the start of the halo
template <int D1, int DP>
__device__ __forceinline__ unsigned v3_pos(unsigned v)
{
	return (v / D1) * DP + (v % D1);
}

// One tap of LowPass::process for the I/Q pair (lowpass.cxx:155-156): the product is rounded on
// its own (fma(c, x, -0) == RN(c * x), see mul2_rn_exact), then added.  {c, c} is the broadcast
// operand form of the packed instruction, so a tap is one 32-bit register.
__device__ __forceinline__ f2_t tap3(f2_t acc, float c, f2_t x, f2_t nz)
{
	return f2_add(acc, f2_fma(x, f2_pack(c, c), nz));
}

// One output of the channel FIR: taps in the reference's order (lowpass.cxx:151-159), every
// offset an immediate.  base32 = shared address of the first frame of the output's first period,
// taps32 = shared address of the receiver's taps (plain floats, four per 128-bit load).
template <int N1, int D1, int DP>
__device__ __forceinline__ f2_t fir3(uint32_t base32, uint32_t taps32, f2_t nz)
{
	using G = V3Geo<N1, D1>;
	f2_t acc = 0ull;
	int j = 0;
	if (G::OFF & 1) {
		acc = tap3(acc, lds32(taps32 + 4u * (G::TAPOFF + 0)), lds64p(base32 + 8u * v3_pos<D1, DP>(G::OFF)), nz);
		j = 1;
	}
	#pragma unroll
	for (; j + 3 < N1; j += 4) {
		f2_t x0, x1, x2, x3;
		const float4 c = lds128f(taps32 + 4u * (G::TAPOFF + j));
		lds128p(base32 + 8u * v3_pos<D1, DP>(G::OFF + j), x0, x1);
		lds128p(base32 + 8u * v3_pos<D1, DP>(G::OFF + j + 2), x2, x3);
		acc = tap3(acc, c.x, x0, nz);
		acc = tap3(acc, c.y, x1, nz);
		acc = tap3(acc, c.z, x2, nz);
		acc = tap3(acc, c.w, x3, nz);
	}
	if (j + 1 < N1) {
		f2_t x0, x1;
		lds128p(base32 + 8u * v3_pos<D1, DP>(G::OFF + j), x0, x1);
		acc = tap3(acc, lds32(taps32 + 4u * (G::TAPOFF + j)), x0, nz);
		acc = tap3(acc, lds32(taps32 + 4u * (G::TAPOFF + j + 1)), x1, nz);
		j += 2;
	}
	if (j < N1)
		acc = tap3(acc, lds32(taps32 + 4u * (G::TAPOFF + j)), lds64p(base32 + 8u * v3_pos<D1, DP>(G::OFF + j)), nz);
	return acc;
}

// ---- tuner-block formats ---------------------------------------------------------------------
// F32: interleaved float IQ, as DspBlock buffers carry it (dspblock.h:45).
// U8 : raw RTL-SDR bytes; the tuner's conversion ((float)b - 128.0) / 128.0 (reference
//      src/io/rtlsdrtuner.cxx:106) runs here, in the load path, so a frame costs 2 bytes of HBM
//      and PCIe traffic instead of 8.  Built as 2^23 + b in the mantissa (one byte permute per
//      component), then one packed fma: (2^23 + b) * 2^-7 - 65537 = (b - 128) / 128, exact.
template <bool U8> struct RawIO;

template <> struct RawIO<false> {
	typedef f2_t T;
	static constexpr unsigned FB = 8;    // bytes per frame
	__device__ __forceinline__ static T zero() { return 0ull; }
	__device__ __forceinline__ static T load(const char *p) { return ldg64p(reinterpret_cast<const float2*>(p)); }
	__device__ __forceinline__ static f2_t cvt(T v, uint32_t) { return v; }
};

template <> struct RawIO<true> {
	typedef uint32_t T;
	static constexpr unsigned FB = 2;
	__device__ __forceinline__ static T zero() { return 0x8080u; }   // (128, 128) -> (0.0, 0.0)
	__device__ __forceinline__ static T load(const char *p)
	{
		unsigned short v;
		asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(v) : "l"(p));
		return v;
	}
	__device__ __forceinline__ static f2_t cvt(T v, uint32_t hi)     // hi = 0x4B00
	{
		uint32_t fi, fq;
		asm("prmt.b32 %0, %1, %2, 0x5440;" : "=r"(fi) : "r"(v), "r"(hi));
		asm("prmt.b32 %0, %1, %2, 0x5441;" : "=r"(fq) : "r"(v), "r"(hi));
		return f2_fma(f2_pack(__uint_as_float(fi), __uint_as_float(fq)), f2_pack(0.0078125f, 0.0078125f),
				f2_pack(-65537.0f, -65537.0f));
	}
};

template <int N1, int D1, int RB, bool U8>
__device__ __forceinline__ void chan_body_v3(const ChanArgs &a, const V3Args &v)
{
	using G = V3Geo<N1, D1>;
	using IO = RawIO<U8>;
	typedef typename IO::T raw_t;
	constexpr unsigned FB = IO::FB;
	constexpr int NMT = kV3Mixers, J = G::J, JH = G::J / 2, S = kV3Slots, FW = kV3FirPerSlot;
	constexpr unsigned kSlotBytes = (unsigned)G::SLOT * 8u;
	extern __shared__ __align__(16) unsigned char wr_smem_v3[];
	const unsigned tid = threadIdx.x;
	// let the demodulator kernel behind this one be scheduled as SMs drain (it waits for this grid)
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

	// Stage the correction table once per CTA: one thread issues bulk copies (TMA, no tensor map
	// needed for a contiguous range) that complete on an mbarrier, so the 132 KiB transfer runs
	// behind the first group's set-up and raw loads instead of in front of them.
	__shared__ __align__(8) unsigned long long wr_bars[1 + 2 * S];   // table | full[S] | empty[S]
	const uint32_t bar32 = (uint32_t)__cvta_generic_to_shared(wr_bars);
	const uint32_t full32 = bar32 + 8u, empty32 = bar32 + 8u + 8u * S;
	if (tid == 0) {
		mbar_init(bar32, 1);
		for (int i = 0; i < S; i++) {
			mbar_init(full32 + 8u * i, NMT / 32);    // one arrival per mixer warp
			mbar_init(empty32 + 8u * i, FW);         // one arrival per FIR warp of the slot
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (tid == 0) {
		const uint32_t dst32 = (uint32_t)__cvta_generic_to_shared(wr_smem_v3);
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar32), "r"(kV3TableBytes) : "memory");
		constexpr unsigned kChunk = 16384;
		for (unsigned off = 0; off < kV3TableBytes; off += kChunk) {
			const unsigned nb = min(kChunk, kV3TableBytes - off);
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
					:: "r"(dst32 + off), "l"(reinterpret_cast<const char*>(v.delta) + off), "r"(nb), "r"(bar32) : "memory");
		}
	}

	uint32_t smem32 = (uint32_t)__cvta_generic_to_shared(wr_smem_v3);
	asm volatile("" : "+r"(smem32));   // opaque: keeps the window base in a register instead of re-deriving it per pass
	const uint32_t ring32 = smem32 + kV3TableBytes;
	const uint32_t taps32 = ring32 + S * kSlotBytes;
	const uint32_t desc32 = taps32 + RB * G::kTapsStride;

	const unsigned P = v.P;
	const unsigned long long U = (unsigned long long)v.nGroups * P;
	const unsigned u0 = (unsigned)((unsigned long long)blockIdx.x * U / gridDim.x);
	const unsigned u1 = (unsigned)((unsigned long long)(blockIdx.x + 1) * U / gridDim.x);

	if (tid < NMT) {
		// ================================= MIXER warps =================================
		const unsigned mt = tid;
		const Lo3Regs lo = lo3_regs(v.eps, smem32 + kV3MidOffset, v.prmtHi);
		const f2_t nzp = f2_pack(v.negzero, v.negzero);
		uint32_t posMain[J];
		#pragma unroll
		for (int j = 0; j < J; j++)
			posMain[j] = 8u * v3_pos<D1, G::DP>((unsigned)G::HF + mt + (unsigned)j * NMT);
		const bool isTail = mt >= (unsigned)(NMT - G::HF);
		const unsigned ti = mt - (unsigned)(NMT - G::HF);          // halo frame this thread carries
		const uint32_t posHalo = 8u * v3_pos<D1, G::DP>(isTail ? ti : 0u);
		const unsigned Pfull = a.F / (unsigned)G::SF;              // passes that lie entirely inside the block
		const bool pfLine = mt < (unsigned)G::SF * FB / 128u;      // this thread prefetches line mt of a pass
		const unsigned pfOff = 4u * (unsigned)G::SF * FB + (128u - FB) * mt;   // bytes from this thread's frame j = 0

		raw_t raw[J];             // raw IQ of the current pass: {i, q} packed, or the two bytes
		float2 tail[RB];
		uint32_t qb[RB];          // biased doubled phase of this thread's frame j = 0 of the current pass
		uint32_t qstep[RB];       // 2 * NMT * step
		unsigned rx[RB];
		#pragma unroll
		for (int rl = 0; rl < RB; rl++) {
			tail[rl] = make_float2(0.0f, 0.0f);
			qb[rl] = 0; qstep[rl] = 0; rx[rl] = 0;
		}

		const bool lane0 = (mt & 31) == 0;
		unsigned n = 0, slot = 0;    // slot uses so far; n = kuse * S + slot
		unsigned kuse = 0;           // how many times the ring has wrapped
		unsigned unit = u0;
		bool tableReady = false;
		const bool tracer = a.ts && blockIdx.x == 0 && mt == 0;
		if (tracer)
			a.ts[kTsChanStart] = global_ns();
		if (a.cta_ts && mt == 0)
			a.cta_ts[2 * blockIdx.x] = global_ns();
		if (a.in_flag) {
			// Pipelined host path: the copy-in stream raises the flag behind the tuner block.  One
			// warp per CTA polls, with warp-uniform control flow (a lane spinning on its own leaves
			// the warp diverged for the rest of the kernel: measured, the mixing ran at half
			// speed), and sparsely (148 CTAs polling one L2 line back to back slowed the very copy
			// they were waiting for); the other mixer warps sleep in the named barrier.
			if (mt < 32) {
				const unsigned long long t0 = global_ns();
				for (;;) {
					unsigned seen;
					asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(a.in_flag) : "memory");
					if (__any_sync(0xFFFFFFFFu, (int)(seen - a.in_seq) >= 0))
						break;
					if (__any_sync(0xFFFFFFFFu, global_ns() - t0 > kSpinNs)) {
						if (mt == 0)
							atomicOr(a.err, kSyncTimeout);   // never hang the GPU on a copy that did not come
						break;
					}
					__nanosleep(a.poll_ns);
				}
			}
			bar_sync(kV3BarMix, NMT);
		}
		if (tracer)
			a.ts[kTsChanInput] = global_ns();
		while (unit < u1) {
			// ---- a run of consecutive passes [p0, pend) of one receiver group ----
			const unsigned rg = unit / P, p0 = unit - rg * P;
			const unsigned pend = min(P, p0 + (u1 - unit));
			// the FIR warps may still read the previous group's taps: wait until every slot filled so
			// far has been handed back (all mixer warps had arrived on it, so nobody is behind)
			for (unsigned m = (n > S ? n - S : 0u); m < n; m++)
				mbar_wait(empty32 + 8u * (m % S), (m / S) & 1u);
			const int4 grp = __ldg(v.groups + rg);
			const int cnt = grp.y;
			const char *__restrict__ src = reinterpret_cast<const char*>(a.iq) + (size_t)(unsigned)grp.z * a.stream_stride * FB;
			// this thread's frame j = 0 of the current pass; frames past the block read as zero
			const char *rawp = src + ((size_t)p0 * G::SF + mt) * FB;
			#pragma unroll
			for (int j = 0; j < J; j++) {
				const unsigned f = p0 * (unsigned)G::SF + mt + (unsigned)j * NMT;
				raw[j] = f < a.F ? IO::load(src + (size_t)f * FB) : IO::zero();
			}
			uint32_t ph0[RB];
			int32_t step[RB];
			#pragma unroll
			for (int rl = 0; rl < RB; rl++) {
				ph0[rl] = 0; step[rl] = 0;
				if (rl < cnt) {
					const unsigned r = __ldg(v.order + grp.x + rl);
					step[rl] = a.conf[r].step;
					ph0[rl] = a.st_in[r].phase;
					rx[rl] = r;
					qstep[rl] = 2u * (uint32_t)NMT * (uint32_t)step[rl];
					qb[rl] = ((ph0[rl] + (p0 * (unsigned)G::SF + mt) * (uint32_t)step[rl]) << 1) + 0x80000000u;
					if (mt < (unsigned)N1)
						sts32(taps32 + rl * G::kTapsStride + 4u * (G::TAPOFF + mt), a.taps1[(size_t)r * N1 + mt]);
					if (isTail && p0 == 0) {
						// frames before the block: the carried history (zeros beyond it)
						const int hidx = (int)(N1 - 1) - G::HF + (int)ti;
						tail[rl] = hidx >= 0 ? a.hist_in[(size_t)r * (N1 - 1) + hidx] : make_float2(0.0f, 0.0f);
					}
				}
			}
			if (!tableReady) {
				// first use of the NCO table: the bulk copies must have landed
				mbar_wait(bar32, 0);
				tableReady = true;
			}
			if (isTail && p0 != 0) {
				// the run starts inside the block: mix the halo frames once
				#pragma unroll
				for (int rl = 0; rl < RB; rl++) {
					if (rl < cnt) {
						const unsigned f = p0 * (unsigned)G::SF - (unsigned)G::HF + ti;
						float sn, cs;
						lo3_sincos(((ph0[rl] + f * (uint32_t)step[rl]) << 1) + 0x80000000u, lo, sn, cs);
						float2 x;
						f2_unpack(IO::cvt(IO::load(src + (size_t)f * FB), lo.hi), x.x, x.y);
						tail[rl] = mix(x, cs, sn);
					}
				}
			}

			// One pass per iteration.  A pass is mixed in two halves of JH frames per thread; while
			// the LAST receiver of the group is mixed, each half's raw registers are refilled with
			// the next pass's frames as soon as the half is done with them, so the loads fly behind
			// the mixing without a second register set.
			for (unsigned p = p0; p < pend; p++) {
				const bool lastPass = (p + 1 == P);
				const bool more = (p + 1 < pend);
				const bool nextFull = (p + 1 < Pfull);
				// pull the pass four ahead into L2 while this one is mixed: one 128-byte line per thread
				if (pfLine && p + 4 < Pfull)
					asm volatile("prefetch.global.L2 [%0];" :: "l"(rawp + pfOff));
				// raw bytes are converted once per pass; float blocks are used where they are
				f2_t cur8[U8 ? J : 1];
				if (U8) {
					#pragma unroll
					for (int j = 0; j < J; j++)
						cur8[U8 ? j : 0] = IO::cvt(raw[j], lo.hi);
				}
				#pragma unroll
				for (int rl = 0; rl < RB; rl++) {
					if (rl < cnt) {
						if (kuse)
							mbar_wait(empty32 + 8u * slot, (kuse - 1u) & 1u);  // the FIR warps are done with this slot
						const uint32_t slot32 = ring32 + slot * kSlotBytes;
						if (isTail)
							sts64(slot32 + posHalo, tail[rl]);
						const bool refill = (rl == cnt - 1) && more;
						#pragma unroll
						for (int h = 0; h < 2; h++) {
							uint32_t q[JH];
							float sn[JH], cs[JH];
							#pragma unroll
							for (int jj = 0; jj < JH; jj++)
								q[jj] = qb[rl] + (uint32_t)(h * JH + jj) * qstep[rl];
							lo3_sincos_n<JH>(q, lo, sn, cs);
							#pragma unroll
							for (int jj = 0; jj < JH; jj++) {
								const int j = h * JH + jj;
								f2_t x;
								if constexpr (U8)
									x = cur8[j];
								else
									x = raw[j];
								// downconverter.cxx:109-110:  I' = i*cos + q*sin ;  Q' = q*cos - i*sin
								float ic, qc, is, qs;
								f2_unpack(f2_fma(x, f2_pack(cs[jj], cs[jj]), nzp), ic, qc);
								f2_unpack(f2_fma(x, f2_pack(sn[jj], sn[jj]), nzp), is, qs);
								const float2 m = make_float2(__fadd_rn(ic, qs), __fsub_rn(qc, is));
								sts64(slot32 + posMain[j], m);
								if (j == J - 1)
									tail[rl] = m;      // only meaningful (and only used) in the tail threads
							}
							if (refill) {
								// this half's raw registers are free: fetch the same frames of the next pass
								if (nextFull) {
									#pragma unroll
									for (int jj = 0; jj < JH; jj++)
										raw[h * JH + jj] = IO::load(rawp + (G::SF + (h * JH + jj) * NMT) * FB);
								} else {
									#pragma unroll
									for (int jj = 0; jj < JH; jj++) {
										const unsigned f = (p + 1) * (unsigned)G::SF + mt + (unsigned)(h * JH + jj) * NMT;
										raw[h * JH + jj] = f < a.F ? IO::load(src + (size_t)f * FB) : IO::zero();
									}
								}
							}
						}
						qb[rl] += (uint32_t)J * qstep[rl];   // 2 * SF * step further: frame j = 0 of the next pass
						if (mt == 0)
							asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};"
									:: "r"(desc32 + slot * 16u), "r"(rx[rl]), "r"(p), "r"(0u), "r"((unsigned)rl) : "memory");
						if (lastPass) {
							// carried state: the last N1-1 mixed frames of [history | block], NCO phase
							bar_sync(kV3BarMix, NMT);
							const unsigned fp = a.F - p * (unsigned)G::SF;          // frames of this pass, 1..SF
							if (mt < (unsigned)(N1 - 1)) {
								const unsigned vv = (unsigned)G::HF + fp - (unsigned)(N1 - 1) + mt;
								a.hist_out[(size_t)rx[rl] * (N1 - 1) + mt] = lds64(slot32 + 8u * v3_pos<D1, G::DP>(vv));
							}
							if (mt == 0)
								a.st_out[rx[rl]].phase = phase_at(a.st_in[rx[rl]].phase, a.conf[rx[rl]].step, a.F);
						}
						__syncwarp();
						if (lane0)
							mbar_arrive(full32 + 8u * slot);
						n++;
						if (++slot == S) {
							slot = 0;
							kuse++;
						}
					}
				}
				rawp += G::SF * FB;
			}
			unit += pend - p0;
		}
		// tell every FIR warp to stop (one descriptor per slot reaches both of its warps)
		for (int i = 0; i < S; i++) {
			if (kuse)
				mbar_wait(empty32 + 8u * slot, (kuse - 1u) & 1u);
			if (mt == 0)
				asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};"
						:: "r"(desc32 + slot * 16u), "r"(0u), "r"(0u), "r"(0u), "r"(0xFFFFFFFFu) : "memory");
			__syncwarp();
			if (lane0)
				mbar_arrive(full32 + 8u * slot);
			n++;
			if (++slot == S) {
				slot = 0;
				kuse++;
			}
		}
		if (tracer)
			a.ts[kTsChanEnd] = global_ns();
		if (a.cta_ts && mt == 0)
			a.cta_ts[2 * blockIdx.x + 1] = global_ns();
	} else {
		// ================================== FIR warps ==================================
		const unsigned fw = (tid - NMT) >> 5, lane = tid & 31;
		const f2_t nz = f2_pack(v.negzero, v.negzero);
		// FIR warp fw serves slot fw % S and, of its NG groups of 32 outputs, every FW-th one
		const unsigned slot = fw % S, sub = fw / S;
		// the demodulator of the previous block may still be reading the channel-rate buffer
		if (!a.wait_late)
			asm volatile("griddepcontrol.wait;" ::: "memory");
		for (unsigned kuse = 0; ; kuse++) {
			mbar_wait_fir(full32 + 8u * slot, kuse & 1u);
			unsigned r, p, unused, rl;
			asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r), "=r"(p), "=r"(unused), "=r"(rl) : "r"(desc32 + slot * 16u));
			if (rl == 0xFFFFFFFFu)
				break;
			const unsigned k0 = p * (unsigned)G::GO;
			const unsigned nout = a.M1 > k0 ? min(a.M1 - k0, (unsigned)G::GO) : 0u;
			const uint32_t slot32 = ring32 + slot * kSlotBytes;
			const uint32_t t32 = taps32 + rl * G::kTapsStride;
			#pragma unroll 1
			for (int g = (int)sub; g < G::NG; g += FW) {
				const unsigned o = (unsigned)g * 32u + lane;
				if ((unsigned)g * 32u < nout) {     // warp-uniform
					const f2_t acc = fir3<N1, D1, G::DP>(slot32 + 8u * o * (unsigned)G::DP, t32, nz);
					if (o < nout) {
						float2 y;
						f2_unpack(acc, y.x, y.y);
						a.chan[(size_t)r * a.chan_stride + k0 + o] = y;
					}
				}
			}
			__syncwarp();
			if (lane == 0)
				mbar_arrive(empty32 + 8u * slot);
		}
	}
	// Programmatic dependent launch: this grid may have run entirely under the previous kernel in
	// the stream, the demodulator of the block before.  Nothing here reads what that kernel writes
	// and nothing it reads is written here (the channel-rate buffer and the carried state alternate
	// between two sides), but the NEXT demodulator waits only for this grid -- so this grid must not
	// complete before its predecessor has, or two demodulator kernels could overlap.
	asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <int N1, int D1, int RB, bool U8>
__global__ void __launch_bounds__(kV3Threads, 1) chan_kernel_v3(const ChanArgs a, const V3Args v)
{
	chan_body_v3<N1, D1, RB, U8>(a, v);
}

// The same kernel held to 96 registers per thread: 16 K of an SM's registers stay free, enough for
// two of the demodulator kernel's CTAs beside this one (wr_bank.cu: persistent demodulator grid).
template <int N1, int D1, int RB, bool U8>
__global__ void __maxnreg__(96) chan_kernel_v3_r96(const ChanArgs a, const V3Args v)
{
	chan_body_v3<N1, D1, RB, U8>(a, v);
}

// ------------------------------------------------------------------ host side ----

typedef void (*V3Kernel)(const ChanArgs, const V3Args);

struct V3Plan {
	bool ok = false;
	V3Kernel kernel = nullptr;     // float tuner blocks
	V3Kernel kernel8 = nullptr;    // raw RTL-SDR bytes
	V3Kernel kernelH = nullptr;    // the same two, instantiated for groups of at most RB/2 receivers:
	V3Kernel kernel8H = nullptr;   // half the unrolled mixer code when the groups are small anyway
	int regs[4] = { 0, 0, 0, 0 };  // registers per thread of kernel, kernel8, kernelH, kernel8H
	unsigned groupCap = 0;         // largest group v3_set_groups built
	int device = 0;
	int numSMs = 0;
	unsigned n1 = 0, d1 = 0;
	unsigned SF = 0, RB = 1;
	unsigned unitsPerSM = 6;       // receiver groups are halved until there are this many (group, pass) units per SM (WR_V3_UNITS_PER_SM)
	size_t smemBytes = 0;
	int16_t *d_delta = nullptr;
	wr::Lo3Coef coef = {};
	unsigned *d_order = nullptr;
	int4 *d_groups = nullptr;
	unsigned nGroups = 0;
	unsigned capR = 0;
	bool pdl = true;          // WR_V3_PDL=0 turns programmatic dependent launch off
	unsigned maxF = 0;        // block length the bank was created for (sizes the receiver groups)
};

template <int N1, int D1, int RB>
inline void v3_fill(V3Plan &p)
{
	using G = V3Geo<N1, D1>;
	p.kernel = chan_kernel_v3<N1, D1, RB, false>;
	p.kernel8 = chan_kernel_v3<N1, D1, RB, true>;
	p.kernelH = chan_kernel_v3<N1, D1, (RB > 1 ? RB / 2 : 1), false>;
	// raw bytes, small groups: the 96-register build (no spills; 105 otherwise), so that small
	// banks fed raw bytes get the persistent demodulator grid too (cfg2 fed bytes: 22.5 -> 21.6 us).
	// The full-size kernels lose more to the cap than the grid brings (cfg5: 842 -> 864 us).
	p.kernel8H = chan_kernel_v3_r96<N1, D1, (RB > 1 ? RB / 2 : 1), true>;
	p.SF = G::SF;
	p.RB = RB;
	p.smemBytes = kV3TableBytes + (size_t)kV3Slots * G::SLOT * 8 + (size_t)RB * G::kTapsStride + (size_t)kV3Slots * 16;
}

inline void v3_destroy(V3Plan &p)
{
	cudaFree(p.d_delta);
	cudaFree(p.d_order);
	cudaFree(p.d_groups);
	p.d_delta = nullptr;
	p.d_order = nullptr;
	p.d_groups = nullptr;
	p.ok = false;
}

// v3 is built for the geometries of the BASELINE configs and the reference's shipped point;
// everything else stays with v2 / v1.
inline int v3_init(V3Plan &p, int device, unsigned n1, unsigned d1, unsigned maxF)
{
	p.device = device;
	p.maxF = maxF;
	p.n1 = n1;
	p.d1 = d1;
	p.ok = false;
	p.kernel = nullptr;
	if (const char *e = getenv("WR_V3_PDL"))
		p.pdl = atoi(e) != 0;
	if (const char *e = getenv("WR_V3_UNITS_PER_SM"))
		p.unitsPerSM = (unsigned)std::max(1, atoi(e));
	if (n1 == 64 && d1 == 10) v3_fill<64, 10, 4>(p);
	else if (n1 == 64 && d1 == 8) v3_fill<64, 8, 4>(p);     // 2.048 MSPS -> 256 k
	else if (n1 == 127 && d1 == 50) v3_fill<127, 50, 4>(p);
	else if (n1 == 255 && d1 == 50) v3_fill<255, 50, 2>(p);
	else if (n1 == 127 && d1 == 40) v3_fill<127, 40, 4>(p);
	else return WR_OK;
	cudaDeviceProp prop;
	WR_CUDA(cudaGetDeviceProperties(&prop, device));
	p.numSMs = prop.multiProcessorCount;
	if (p.smemBytes > (size_t)prop.sharedMemPerBlockOptin)
		return WR_OK;
	WR_CUDA(cudaFuncSetAttribute(p.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smemBytes));
	WR_CUDA(cudaFuncSetAttribute(p.kernel8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smemBytes));
	WR_CUDA(cudaFuncSetAttribute(p.kernelH, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smemBytes));
	WR_CUDA(cudaFuncSetAttribute(p.kernel8H, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smemBytes));
	{
		const V3Kernel ks[4] = { p.kernel, p.kernel8, p.kernelH, p.kernel8H };
		for (int i = 0; i < 4; i++) {
			cudaFuncAttributes fa;
			WR_CUDA(cudaFuncGetAttributes(&fa, ks[i]));
			p.regs[i] = fa.numRegs;
		}
	}
	WR_CUDA(cudaMalloc(&p.d_delta, kV3TableBytes));
	p.ok = true;
	return WR_OK;
}

inline bool v3_supported(const V3Plan &p, unsigned F) { return p.ok && F >= p.SF; }

// (Re)compress the NCO table after it changed; v3 turns itself off if it is not representable.
inline int v3_set_table(V3Plan &p, const float *h_table, cudaStream_t st)
{
	if (!p.d_delta)
		return WR_OK;
	std::vector<int16_t> delta(WR_SINTABLE_SIZE);
	if (!wr::lo3_compress(h_table, delta.data(), &p.coef)) {
		p.ok = false;
		return WR_OK;
	}
	std::vector<int16_t> padded(kV3TableBytes / 2, 0);
	for (int s = -32768; s < 32768; s++)
		padded[(size_t)(wr::lo3_slot_host(s) - WR_LO3_SLOT_MIN)] = delta[(uint16_t)s];
	WR_CUDA(cudaMemcpyAsync(p.d_delta, padded.data(), kV3TableBytes, cudaMemcpyHostToDevice, st));
	WR_CUDA(cudaStreamSynchronize(st)); // `padded` is a local
	return WR_OK;
}

// Receivers sorted by stream and cut into groups of <= RB that share one.
inline int v3_set_groups(V3Plan &p, const RxConf *h_conf, unsigned R, cudaStream_t st)
{
	if (!p.d_delta)
		return WR_OK;
	std::vector<unsigned> order(R);
	for (unsigned r = 0; r < R; r++)
		order[r] = r;
	std::stable_sort(order.begin(), order.end(),
			[&](unsigned x, unsigned y) { return h_conf[x].stream < h_conf[y].stream; });
	// Receivers that share a stream are mixed from one set of raw registers, up to RB at a time.
	// A work unit is (group, pass); with few receivers and short blocks large groups leave the
	// 148 CTAs with 3 or 4 units each (a 15% imbalance), so the group size is halved until there
	// are enough units to spread evenly.
	const unsigned passes = std::max(1u, (p.maxF + p.SF - 1) / p.SF);
	unsigned cap = p.RB;
	std::vector<int4> groups;
	for (;; cap /= 2) {
		groups.clear();
		for (unsigned i = 0; i < R;) {
			unsigned n = 1;
			while (i + n < R && n < cap && h_conf[order[i + n]].stream == h_conf[order[i]].stream)
				n++;
			groups.push_back(make_int4((int)i, (int)n, (int)h_conf[order[i]].stream, 0));
			i += n;
		}
		if (cap == 1 || (unsigned long long)groups.size() * passes >= (unsigned long long)p.unitsPerSM * (unsigned)p.numSMs)
			break;
	}
	p.groupCap = cap;
	if (R > p.capR) {
		cudaFree(p.d_order);
		cudaFree(p.d_groups);
		p.d_order = nullptr;
		p.d_groups = nullptr;
		WR_CUDA(cudaMalloc(&p.d_order, sizeof(unsigned) * R));
		WR_CUDA(cudaMalloc(&p.d_groups, sizeof(int4) * R));
		p.capR = R;
	}
	WR_CUDA(cudaMemcpyAsync(p.d_order, order.data(), sizeof(unsigned) * R, cudaMemcpyHostToDevice, st));
	WR_CUDA(cudaMemcpyAsync(p.d_groups, groups.data(), sizeof(int4) * groups.size(), cudaMemcpyHostToDevice, st));
	WR_CUDA(cudaStreamSynchronize(st)); // locals
	p.nGroups = (unsigned)groups.size();
	return WR_OK;
}

// Registers of an SM that the channel kernel's CTA leaves to other kernels' CTAs (what decides how
// many demodulator CTAs run beside it).
inline int v3_spare_regs(const V3Plan &p, bool u8)
{
	const bool half = p.RB > 1 && p.groupCap <= p.RB / 2;
	const int r = p.regs[(half ? 2 : 0) + (u8 ? 1 : 0)];
	return 65536 - ((r + 7) & ~7) * kV3Threads;
}

inline int v3_launch_chan(V3Plan &p, ChanArgs &ca, bool u8, cudaStream_t st, unsigned long long *launches)
{
	V3Args v;
	v.delta = p.d_delta;
	v.eps = p.coef.eps;
	v.order = p.d_order;
	v.groups = p.d_groups;
	v.nGroups = p.nGroups;
	v.P = (ca.F + p.SF - 1) / p.SF;
	v.negzero = -0.0f;
	v.prmtHi = 0x4B00u;
	const unsigned long long units = (unsigned long long)v.nGroups * v.P;
	const unsigned grid = (unsigned)std::min<unsigned long long>(units, (unsigned long long)p.numSMs);
	// Programmatic dependent launch: this grid may run while the previous kernel in the stream (the
	// demodulator of the block before) is still at work; it executes griddepcontrol.wait only
	// before it exits (see the end of the kernel).
	cudaLaunchConfig_t cfg = {};
	cudaLaunchAttribute attr[1];
	cfg.gridDim = dim3(grid);
	cfg.blockDim = dim3(kV3Threads);
	cfg.dynamicSmemBytes = p.smemBytes;
	cfg.stream = st;
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = p.pdl ? 1 : 0;
	const bool half = p.RB > 1 && p.groupCap <= p.RB / 2;
	const V3Kernel k = half ? (u8 ? p.kernel8H : p.kernelH) : (u8 ? p.kernel8 : p.kernel);
	cudaError_t e = cudaLaunchKernelEx(&cfg, k, (const ChanArgs)ca, (const V3Args)v);
	(*launches)++;
	if (e == cudaSuccess)
		e = cudaGetLastError();
	if (e != cudaSuccess) {
		wr::set_error("chan_kernel_v3 launch (grid %u, %d threads, %zu bytes of shared memory): %s",
				grid, kV3Threads, p.smemBytes, cudaGetErrorString(e));
		return WR_ECUDA;
	}
	return WR_OK;
}

} // namespace wrd
