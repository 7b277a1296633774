// wr_kernels_v2.cuh -- fused NCO mix + decimating channel FIR + demod with the NCO table
// RESIDENT IN SHARED MEMORY (K1+K2+K3 of SURVEY.md 2a, second generation).
//
// Why: v1 gathers sin/cos from the 256 KiB float table through L1/L2 -- two scattered 4-byte
// loads per receiver-frame, which bound the kernel at ~70% L1 utilisation (profiles/r01_v1_*).
// The table cannot fit an SM as floats, but it can as 16-bit corrections to a closed-form base
// (wr_lo.h): 128 KiB per CTA, bit-exact by construction and verified on the host.
//
// Shape: persistent grid, one 1024-thread CTA per SM, warp-specialised.  Each CTA loops over
//   item = (group of <= RB receivers listening to the same tuner stream, tile of TK outputs).
//   * MIXER warps load the raw IQ the tile touches ONCE into registers (coalesced float2 loads,
//     J frames per thread), and for every receiver of the group mix it (LO from the
//     shared-memory table) into one of two shared-memory tiles, laid out period-major with an
//     odd padded period so that the FIR's lanes hit distinct banks;
//   * CONSUMER warps run one thread per output over the taps in the reference's order (packed
//     f32x2: I and Q in one instruction, each product and each sum still rounded separately;
//     fully unrolled for the common (taps, decimation) pairs) and write the channel-rate IQ
//     (8 bytes per output); on the last tile of a receiver they also write the carried state
//     (mixed history, NCO phase).  The demodulator runs fused into the audio-FIR kernel over the
//     channel-rate stream (demod_audio_kernel_v2): its double-precision atan2 is latency-bound
//     and would otherwise sit on the consumers' critical path.
// The two roles hand tiles over through named barriers (full/empty per buffer), so the
// latency-bound tap chains of receiver g overlap the mixing of receiver g+1.
#pragma once

#include "wr_bank.cuh"
#include "wr_device.cuh"
#include "wr_lo.h"

#include <algorithm>
#include <cstdlib>
#include <vector>

namespace wrd {

constexpr int kV2Threads = 1024;
constexpr int kV2J = 8;                           // raw frames held in registers per mixer thread
// Correction table in shared memory, indexed by the SIGNED 16-bit table index s from its middle:
//     entry slot(s) = (s * 1057) >> 10        (arithmetic shift; = s + floor(33 s / 1024))
// i.e. one padding entry after every ~31 entries.  A warp's 32 lookups form an arithmetic
// progression in s (stride = IF step >> 15); without padding, strides that are multiples of a
// power of two pile onto a few banks (measured 6.5x wavefront excess on cfg2, 32-way for IFs that
// are multiples of Fs/32).  The slot map is strictly increasing (injective) and moves the bank
// by ~1 per 64, per 2048 and per 1024*odd entries of stride alike.
constexpr int kLoPadMul = 1057, kLoPadShift = 10;
constexpr int kLoSlotMin = (-32768 * kLoPadMul) >> kLoPadShift;   // -33824
constexpr int kLoSlotMax = (32767 * kLoPadMul) >> kLoPadShift;    //  33822
constexpr unsigned kV2TableBytes = ((unsigned)(kLoSlotMax - kLoSlotMin + 1) * 2u + 15u) & ~15u;
constexpr unsigned kLoMidOffset = (unsigned)(-kLoSlotMin) * 2u;   // byte offset of the entry of s = 0

// named barriers (0 is __syncthreads)
constexpr int kBarFull = 1;    // +buffer: tile written, consumers may read
constexpr int kBarEmpty = 3;   // +buffer: tile consumed, mixers may overwrite
constexpr int kBarCons = 5;    // among the consumer warps only

struct V2Args {
	const int16_t *delta;     // [65536] corrections (HBM copy, staged to shared memory per CTA)
	float eps;                // table-dependent clamp of the base polynomial (wr_lo.h)
	const unsigned *order;    // receivers sorted by stream
	const int4 *groups;       // {first index into order, count, stream, unused}
	unsigned nGroups;
	unsigned TK, ntiles, nItems;
	unsigned NC;              // consumer warps (the other kV2Threads/32 - NC warps mix)
	unsigned A;               // periods of history an output reaches back: ceil((n1-1)/d1)
	unsigned off;             // A*d1 - (n1-1): offset of an output's first tap in its period
	unsigned Dp;              // padded period (d1 or d1+1, odd)
	unsigned magicD;          // ceil(2^32 / d1): u / d1 == umulhi(u, magicD) over a tile
	unsigned Ucap;            // frames one tile buffer holds
	unsigned Lcap;            // float2 slots of one mixed tile (Ucap plus period padding)
	float negzero;            // -0.0f, deliberately opaque to the compiler (see mul2_rn_exact)
};

__device__ __forceinline__ void bar_sync(int id, int count)
{
	asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory");
}

__device__ __forceinline__ void bar_arrive(int id, int count)
{
	asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(count) : "memory");
}

// Base of the compressed table; the host twin is wr::lo_base_host (same IEEE operations, same
// constants: these literals must stay identical to wr::lo_coef() in wr_lo.h).
constexpr float kLoA0 = 2.9261695289051204e-09f;
constexpr float kLoA1 = 2.714416547945065e-18f;
constexpr float kLoA2 = 9.777997364797523e-28f;

__device__ __forceinline__ float lo_base(int s, float eps)
{
	const float sf = __int2float_rn(s);
	const float w = fmaxf(__fsub_rn(32768.0f, fabsf(sf)), eps);
	const float u = __fmul_rn(sf, w);
	const float au = fabsf(u);
	float p = __fmaf_rn(au, kLoA2, kLoA1);
	p = __fmaf_rn(au, p, kLoA0);
	return __fmul_rn(u, p);
}

// dmid32 = shared-space byte address of the table entry of s = 0; q = phase << 1 (so that the
// signed table index is q >> 16).
__device__ __forceinline__ float lo_value(int q, uint32_t dmid32, float eps)
{
	const int s = q >> 16;
	const uint32_t addr = dmid32 + 2u * (uint32_t)((s * kLoPadMul) >> kLoPadShift);
	int d;
	asm("ld.shared.s16 %0, [%1];" : "=r"(d) : "r"(addr));
	return __int_as_float(__float_as_int(lo_base(s, eps)) + d);
}

// sin/cos of the NCO for a 31-bit phase (bit 31 of `p` is garbage and ignored), exactly the
// reference's sinTable[sinidx], sinTable[cosidx] (downconverter.cxx:100-102).
__device__ __forceinline__ void lo_sincos(uint32_t p, uint32_t dmid32, float eps, float &sn, float &cs)
{
	const int qs = (int)(p << 1);
	const int qc = (int)((p << 1) + 0x40000000u);   // + a quarter turn
	sn = lo_value(qs, dmid32, eps);
	cs = lo_value(qc, dmid32, eps);
}

// Packed (I,Q) product, each half rounded exactly like a scalar multiply.
// ptxas 12.9 contracts mul.rn.f32x2 followed by add.rn.f32x2 into one FFMA2 -- a single rounding,
// which breaks bit parity with the reference's separately rounded `out += coeff * sample`
// (lowpass.cxx:156) -- even with -fmad=false and explicit .rn.  Writing the product as
// fma(a, b, -0.0) with the -0.0 coming from a kernel argument keeps it a distinct instruction:
// fma(a, b, -0) == RN(a*b) for every input, including signed zeros and denormals.
__device__ __forceinline__ float2 mul2_rn_exact(float2 a, float2 b, float2 nz)
{
	float2 r;
	asm("fma.rn.f32x2 %0, %1, %2, %3;"
			: "=l"(*reinterpret_cast<unsigned long long*>(&r))
			: "l"(*reinterpret_cast<const unsigned long long*>(&a)),
			  "l"(*reinterpret_cast<const unsigned long long*>(&b)),
			  "l"(*reinterpret_cast<const unsigned long long*>(&nz)));
	return r;
}

// Tile geometry of one work item, identical for both roles and every receiver of the group.
struct TileGeo {
	bool last;
	unsigned kstart, nout;
	int fb0;          // frame index (relative to the block) of local coordinate u = 0
	unsigned U;       // frames to mix
	unsigned nhist;   // leading frames that lie before the block (carried history)
};

__device__ __forceinline__ TileGeo tile_geo(const ChanArgs &a, const V2Args &v, unsigned tile)
{
	TileGeo g;
	g.last = (tile == v.ntiles - 1);
	g.kstart = tile * v.TK;
	const unsigned kend = min(g.kstart + v.TK, a.M1);
	g.nout = kend > g.kstart ? kend - g.kstart : 0;
	g.fb0 = ((int)g.kstart - (int)v.A) * (int)a.d1;   // aligned to a decimation period
	// frames the outputs need: up to (kend-1)*d1; the last tile runs on to F-1 (history)
	g.U = g.nout ? (unsigned)((int)((kend - 1) * a.d1) - g.fb0 + 1) : 0;
	if (g.last)
		g.U = (unsigned)((int)a.F - g.fb0);
	g.nhist = g.fb0 < 0 ? (unsigned)(-g.fb0) : 0;
	return g;
}

// One output of the channel FIR with taps and decimation known at compile time: straight-line
// code, every shared-memory offset an immediate.  Same tap order and rounding as the generic loop.
template <int N1, int D1, int DP>
__device__ __forceinline__ float2 fir_unrolled(const float2 *__restrict__ base, const float2 *__restrict__ rt2, float2 nz)
{
	constexpr int A = (N1 - 1 + D1 - 1) / D1;
	constexpr int OFF = A * D1 - (N1 - 1);
	float2 acc = make_float2(0.0f, 0.0f);
	#pragma unroll
	for (int j = 0; j < N1; j++) {
		const int u = OFF + j;
		const int pos = (u / D1) * DP + (u % D1);
		acc = __fadd2_rn(acc, mul2_rn_exact(rt2[j], base[pos], nz));
		// keep the scheduler from hoisting every load of the chain to the top (64 registers per
		// thread at 1024 threads per CTA): loads may only run ~8 taps ahead of the adds
		if ((j & 7) == 7)
			asm volatile("" ::: "memory");
	}
	return acc;
}

template <int NT, int J, bool kPad, int N1C, int D1C>
__global__ void __launch_bounds__(NT, 1) chan_kernel_v2(const ChanArgs a, const V2Args v)
{
	extern __shared__ __align__(16) unsigned char wr_smem_v2[];
	const unsigned tid = threadIdx.x;
	const unsigned n1 = a.n1, d1 = a.d1;
	const unsigned tileBytes = v.Lcap * 8u;
	float2 *rt2 = reinterpret_cast<float2*>(wr_smem_v2 + kV2TableBytes + 2 * (size_t)tileBytes); // taps {c, c}

	// stage the correction table (128 KiB) once per CTA
	{
		const uint4 *g = reinterpret_cast<const uint4*>(v.delta);
		uint4 *d = reinterpret_cast<uint4*>(wr_smem_v2);
		#pragma unroll 4
		for (unsigned i = tid; i < kV2TableBytes / 16; i += NT)
			d[i] = __ldg(g + i);
	}
	__syncthreads();

	// items are numbered tile-fastest; a CTA steps through them with stride gridDim.x
	const unsigned stepT = gridDim.x % v.ntiles, stepG = gridDim.x / v.ntiles;
	const unsigned NCT = v.NC * 32;        // consumer threads
	const unsigned NMT = NT - NCT;         // mixer threads
	// Shared-space addresses and loop constants, bounced through a shuffle so that the compiler
	// keeps them in registers instead of re-deriving them in every unrolled mix body.
	const uint32_t smem32 = __shfl_sync(0xffffffffu, (uint32_t)__cvta_generic_to_shared(wr_smem_v2), 0);

	if (tid >= NCT) {
		// =============================== MIXER warps ===============================
		const unsigned mt = tid - NCT;
		const uint32_t dmid32 = smem32 + kLoMidOffset;                   // table entry of s = 0
		const float eps = __shfl_sync(0xffffffffu, v.eps, 0);
		const unsigned padMagic = kPad ? __shfl_sync(0xffffffffu, v.magicD, 0) : 0u; // pos(u) = u + u / d1
		unsigned n = 0;
		unsigned tile = blockIdx.x % v.ntiles, gidx = blockIdx.x / v.ntiles;
		for (unsigned item = blockIdx.x; item < v.nItems; item += gridDim.x) {
			const TileGeo g = tile_geo(a, v, tile);
			const int4 grp = __ldg(v.groups + gidx);
			tile += stepT; gidx += stepG;
			if (tile >= v.ntiles) { tile -= v.ntiles; gidx++; }
			// parameters of the group's first receiver, fetched alongside the raw tile
			unsigned r = __ldg(v.order + grp.x);
			int32_t step = a.conf[r].step;
			uint32_t phase0 = a.st_in[r].phase;

			// raw IQ of the tile, once, into registers (zeros before the block and past U)
			const float2 *__restrict__ src = a.iq + (size_t)(unsigned)grp.z * a.stream_stride + (g.fb0 + (int)mt);
			const unsigned ulim = g.U - g.nhist;
			float2 raw[J];
			#pragma unroll
			for (int j = 0; j < J; j++) {
				const unsigned u = mt + j * NMT;
				raw[j] = (u - g.nhist < ulim) ? __ldg(src + j * NMT) : make_float2(0.0f, 0.0f);
			}

			for (int gi = 0; gi < grp.y; gi++, n++) {
				// this receiver's parameters were fetched one iteration ago; fetch the next one's
				const unsigned rcur = r;
				const int32_t stepcur = step;
				const uint32_t phasecur = phase0;
				if (gi + 1 < grp.y) {
					r = __ldg(v.order + grp.x + gi + 1);
					step = a.conf[r].step;
					phase0 = a.st_in[r].phase;
				}
				const uint32_t tile32 = smem32 + kV2TableBytes + (n & 1) * tileBytes;
				if (n >= 2)
					bar_sync(kBarEmpty + (n & 1), NT);       // consumers are done with this buffer

				// mix (slots past U or before the block get zeros: raw is zero there)
				const uint32_t pstep = (uint32_t)stepcur * NMT;
				uint32_t p = phasecur + (uint32_t)(g.fb0 + (int)mt) * (uint32_t)stepcur;
				#pragma unroll
				for (int j = 0; j < J; j += 2) {
					// two bodies per (uniform) trip test: independent chains for the scheduler
					if (j * NMT < g.U) {
						#pragma unroll
						for (int jj = j; jj < j + 2; jj++) {
							const unsigned u = mt + jj * NMT;
							float sn, cs;
							lo_sincos(p, dmid32, eps, sn, cs);
							const float2 m = mix(raw[jj], cs, sn);
							const unsigned pos = kPad ? u + __umulhi(u, padMagic) : u;
							if (u < g.U) // slots past the tile's last frame may lie outside the buffer
								asm volatile("st.shared.v2.f32 [%0], {%1, %2};" :: "r"(tile32 + 8u * pos), "f"(m.x), "f"(m.y) : "memory");
							p += pstep;
						}
					}
				}
				if (g.nhist) { // first tile: frames before the block come from the carried history
					for (unsigned u = mt; u < g.nhist; u += NMT) {
						const unsigned pos = kPad ? u + __umulhi(u, padMagic) : u;
						const int hidx = (int)(n1 - 1) + g.fb0 + (int)u;
						const float2 h = hidx >= 0 ? a.hist_in[(size_t)rcur * (n1 - 1) + hidx] : make_float2(0.0f, 0.0f);
						asm volatile("st.shared.v2.f32 [%0], {%1, %2};" :: "r"(tile32 + 8u * pos), "f"(h.x), "f"(h.y) : "memory");
					}
				}
				bar_arrive(kBarFull + (n & 1), NT);
			}
		}
		// drain the (up to two) hand-backs nobody waited for, so no barrier is left half-arrived
		for (unsigned m = n >= 2 ? n - 2 : 0; m < n; m++)
			bar_sync(kBarEmpty + (m & 1), NT);
	} else {
		// ============================== CONSUMER warps ==============================
		const float2 nz = make_float2(v.negzero, v.negzero);
		const unsigned padMagic = kPad ? v.magicD : 0u;
		unsigned n = 0;
		unsigned tile = blockIdx.x % v.ntiles, gidx = blockIdx.x / v.ntiles;
		for (unsigned item = blockIdx.x; item < v.nItems; item += gridDim.x) {
			const TileGeo g = tile_geo(a, v, tile);
			const int4 grp = __ldg(v.groups + gidx);
			tile += stepT; gidx += stepG;
			if (tile >= v.ntiles) { tile -= v.ntiles; gidx++; }
			// one otherwise idle consumer warp pulls the NEXT item's raw tile into L2 while this
			// one is processed, so the mixers' loads see L2 rather than HBM latency
			if (tid >= NCT - 32 && item + gridDim.x < v.nItems) {
				const TileGeo gn = tile_geo(a, v, tile);
				const unsigned nstream = (unsigned)__ldg(v.groups + gidx).z;
				const int f0 = max(gn.fb0, 0);
				const int f1 = gn.fb0 + (int)gn.U;
				const char *base = reinterpret_cast<const char*>(a.iq + (size_t)nstream * a.stream_stride + f0);
				const int bytes = (f1 - f0) * 8;
				for (int off = (int)(tid & 31) * 128; off < bytes; off += 32 * 128)
					asm volatile("prefetch.global.L2 [%0];" :: "l"(base + off));
			}
			for (int gi = 0; gi < grp.y; gi++, n++) {
				const unsigned r = v.order[grp.x + gi];
				const float2 *s = reinterpret_cast<const float2*>(wr_smem_v2 + kV2TableBytes + (n & 1) * (size_t)tileBytes);

				for (unsigned i = tid; i < n1; i += NCT) {
					const float c = a.taps1[(size_t)r * n1 + i];
					rt2[i] = make_float2(c, c);
				}
				bar_sync(kBarCons, NCT);
				bar_sync(kBarFull + (n & 1), NT);            // the mixers finished this tile

				// FIR: one thread per output, taps in the reference's order
				for (unsigned o = tid; o < g.nout; o += NCT) {
					float2 acc;
					if constexpr (N1C > 0) {
						acc = fir_unrolled<N1C, D1C, kPad ? D1C + 1 : D1C>(s + (size_t)o * v.Dp, rt2, nz);
					} else {
						acc = make_float2(0.0f, 0.0f);
						unsigned j = 0, q = v.off;
						const float2 *base = s + (size_t)o * v.Dp;
						while (j < n1) {
							const unsigned lim = min(d1 - q, n1 - j);
							const float2 *x = base + q;
							const float2 *c = rt2 + j;
							unsigned t = 0;
							for (; t + 4 <= lim; t += 4) {
								const float2 x0 = x[t], x1 = x[t + 1], x2 = x[t + 2], x3 = x[t + 3];
								const float2 c0 = c[t], c1 = c[t + 1], c2 = c[t + 2], c3 = c[t + 3];
								acc = __fadd2_rn(acc, mul2_rn_exact(c0, x0, nz));
								acc = __fadd2_rn(acc, mul2_rn_exact(c1, x1, nz));
								acc = __fadd2_rn(acc, mul2_rn_exact(c2, x2, nz));
								acc = __fadd2_rn(acc, mul2_rn_exact(c3, x3, nz));
							}
							for (; t < lim; t++)
								acc = __fadd2_rn(acc, mul2_rn_exact(c[t], x[t], nz));
							j += lim;
							base += v.Dp;
							q = 0;
						}
					}
					a.chan[(size_t)r * a.chan_stride + g.kstart + o] = acc;
				}
				if (g.last) {
					// carried state: the last n1-1 mixed frames of [history | block], NCO phase
					for (unsigned i = tid; i + 1 < n1; i += NCT) {
						const unsigned u = (unsigned)((int)a.F - (int)(n1 - 1) + (int)i - g.fb0);
						const unsigned pos = kPad ? u + __umulhi(u, padMagic) : u;
						a.hist_out[(size_t)r * (n1 - 1) + i] = s[pos];
					}
					if (tid == 0)
						a.st_out[r].phase = phase_at(a.st_in[r].phase, a.conf[r].step, a.F);
				}
				bar_sync(kBarCons, NCT);                     // every consumer is done reading the tile
				bar_arrive(kBarEmpty + (n & 1), NT);         // hand it back to the mixers
			}
		}
	}
}

// Demodulator + audio FIR (K3 + K4) over the channel-rate stream the kernel above wrote.
// Grid (tiles + 1, receivers).  A tile stages [history | demodulated samples] for its outputs in
// shared memory -- demodulating the n2-1 samples of overlap with the previous tile again rather
// than waiting for another CTA -- and runs one thread per audio output over the taps in the
// reference's order.  The extra CTA per receiver writes the carried state: audio-FIR history,
// prev_i / prev_q (reference demodulator.cxx:110-111), and the demodulated samples past the last
// audio output.
struct DemodAudioArgs {
	const float2 *chan;      // [R][chan_stride] channel-rate IQ of this block
	size_t chan_stride;
	const RxConf *conf;
	const RxState *st_in;
	RxState *st_out;
	float *x;                // demod side `cur`: [0,n2-1) history (read), then this block's samples (written)
	float *x_next;           // demod side `cur^1`: receives the next block's history
	size_t dstride;
	const float *taps2;      // [R][n2] reversed
	float *audio;
	size_t audio_stride;
	unsigned M1, M2, n2, d2;
	unsigned TK, ntiles;
	unsigned items;          // (ntiles + 1) * receivers
	float out_scale;         // 1, or 32768 for the encoder's sample format (reference mp3encoder.cxx:66-73)
	float negzero;           // -0.0f, opaque to the compiler (mul2_rn_exact)
	// Pipelined host path: the last CTA to finish publishes done_seq in done_flag (device memory), which
	// the copy-out stream waits on -- no event between the kernels of consecutive blocks, so the
	// programmatic overlap of this grid with the next block's channel kernel survives.
	unsigned *done_count;    // nullptr = nobody waits
	unsigned *done_flag;     // in HBM (a copy-out stream waits) ...
	unsigned *host_done;     // ... or in mapped host memory, when `audio` itself is the caller's pinned buffer
	unsigned done_seq;
	unsigned long long *ts;  // optional trace record (wr_bank.cuh)
	unsigned long long *cta_ts;  // optional per-CTA {start, end} of this block (WR_TRACE_CTA)
};

// sample i of [history | demod(chan)] for receiver r
__device__ __forceinline__ float demod_at(const DemodAudioArgs &a, const float *xr, const float2 *ch,
		int mode, float2 prev0, unsigned i)
{
	if (i < a.n2 - 1)
		return xr[i];
	const unsigned k = i - (a.n2 - 1);
	return demod(mode, ch[k], k ? ch[k - 1] : prev0);
}

// End of a CTA: count it; the last one of the grid raises the flag (release: every CTA fenced its
// stores before it counted itself).
__device__ __forceinline__ void demod_audio_done(const DemodAudioArgs &a, unsigned tid)
{
	if (a.cta_ts && tid == 0)
		a.cta_ts[2 * blockIdx.x + 1] = global_ns();
	if (!a.done_count)
		return;
	__syncthreads();
	if (tid == 0) {
		if (a.host_done)
			__threadfence_system();   // this CTA's audio went over PCIe
		else
			__threadfence();
		const unsigned total = gridDim.x * gridDim.y;
		if (atomicAdd(a.done_count, 1u) == total - 1u) {
			*a.done_count = 0;      // the next launch of this kernel starts after this grid completed
			if (a.ts)
				a.ts[kTsDemodEnd] = global_ns();
			if (a.host_done) {
				__threadfence_system();
				*reinterpret_cast<volatile unsigned*>(a.host_done) = a.done_seq;
			} else if (a.done_flag) {
				__threadfence();
				atomicExch(a.done_flag, a.done_seq);
			}
		}
	}
}

// Staging of a tile for the modes that need one channel-rate sample per output sample (AM, USB,
// LSB -- the mode is the receiver's, i.e. uniform over the CTA, so it is a template parameter of the
// loop instead of a switch per sample): four samples per thread and round, loads first.  The
// general loop of the kernel spends ~70 instructions per sample on the mode switch, the FM
// discriminator's second operand and its bounds; on cfg3 that was half of the kernel's instructions.
template <int MODE, int kThreads>
__device__ __forceinline__ void demod_stage_simple(const float2 *__restrict__ ch, float *__restrict__ xr, float *__restrict__ s,
		unsigned tid, unsigned i0, unsigned L, unsigned nh)
{
	for (unsigned ib = tid; ib < L; ib += 4 * kThreads) {
		float2 c[4];
		float h[4];
		#pragma unroll
		for (int u = 0; u < 4; u++) {
			const unsigned il = ib + u * kThreads, i = i0 + il;
			c[u] = make_float2(0.0f, 0.0f);
			h[u] = 0.0f;
			if (il < L) {
				if (i < nh)
					h[u] = xr[i];
				else
					c[u] = ch[i - nh];
			}
		}
		#pragma unroll
		for (int u = 0; u < 4; u++) {
			const unsigned il = ib + u * kThreads, i = i0 + il;
			if (il < L) {
				float v = h[u];
				if (i >= nh) {
					v = demod(MODE, c[u], c[u]);
					xr[i] = v;
				}
				s[il] = v;
			}
		}
	}
}

// This kernel runs under the NEXT block's channel kernel (programmatic dependent launch), whose
// CTA needs almost a whole SM: it starts on an SM only when that SM's demodulator CTAs have
// drained, so the life time of a CTA here is on the critical path of the block.  It is all
// latency -- hence every global load of a tile is issued before anything waits for one, and the
// CTA is wide enough to stage a tile of the common geometry in a single round.
constexpr int kDemodThreads = 192;
constexpr int kDemodRounds = 2;

template <int kThreads>
__global__ void __launch_bounds__(kThreads, 7) demod_audio_kernel_v2(const DemodAudioArgs a)
{
	extern __shared__ float4 wr_smem_da[];
	const unsigned tid = threadIdx.x;
	const unsigned n2 = a.n2, d2 = a.d2;
	// Programmatic dependent launch, both ways: this grid may have been launched while the channel
	// kernel that feeds it was still running (wait for it), and the next block's channel kernel may
	// start its prologue now.  Both are no-ops for a plain launch.
	asm volatile("griddepcontrol.wait;" ::: "memory");
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
	if (a.ts && tid == 0) {
		if (blockIdx.x == 0)
			a.ts[kTsDemodStart] = global_ns();
		if (a.cta_ts)
			a.cta_ts[2 * blockIdx.x] = global_ns();
	}
	// work items (receiver, tile): one per CTA, or -- persistent grid -- every gridDim.x-th one
	const unsigned perRx = a.ntiles + 1;
	for (unsigned w = blockIdx.x; w < a.items; w += gridDim.x) {
	const unsigned r = w / perRx, tile = w - r * perRx;
	const RxConf cf = a.conf[r];
	const RxState st = a.st_in[r];
	const float2 prev0 = make_float2(st.prev_i, st.prev_q);
	const float2 *__restrict__ ch = a.chan + (size_t)r * a.chan_stride;
	float *xr = a.x + (size_t)r * a.dstride;

	if (tile == a.ntiles) {
		// next block's history: the last n2-1 samples of [history | demod]
		for (unsigned i = tid; i + 1 < n2; i += kThreads)
			a.x_next[(size_t)r * a.dstride + i] = demod_at(a, xr, ch, cf.mode, prev0, a.M1 + i);
		// demodulated samples no audio output of this block consumed
		for (unsigned k = a.M2 * d2 + tid; k < a.M1; k += kThreads)
			xr[(n2 - 1) + k] = demod(cf.mode, ch[k], k ? ch[k - 1] : prev0);
		if (tid == 0) {
			const float2 lastc = a.M1 ? ch[a.M1 - 1] : prev0;
			a.st_out[r].prev_i = lastc.x;
			a.st_out[r].prev_q = lastc.y;
		}
		continue;
	}

	const unsigned Lmax = a.TK * d2 + n2 - 1;
	float *s = reinterpret_cast<float*>(wr_smem_da);
	float *rt = s + ((Lmax + 3) & ~3u);
	const unsigned m0 = tile * a.TK;
	const unsigned mend = min(m0 + a.TK, a.M2);
	const unsigned nout = mend - m0;
	// staged window: what the outputs read ((nout-1)*d2 + n2 samples) extended to the end of
	// the tile's own decimation periods, so that every sample k in [m0*d2, mend*d2) is produced
	// (and written to the demod stream in HBM) by exactly one tile
	const unsigned L = nout * d2 + n2 - 1;
	const unsigned i0 = m0 * d2;
	const float tap0 = tid < n2 ? a.taps2[(size_t)r * n2 + tid] : 0.0f;
	if (cf.mode != WR_MODE_FM) {
		// one-sample modes: the specialised staging loops
		if (tid < n2)
			rt[tid] = tap0;
		for (unsigned i = tid + kThreads; i < n2; i += kThreads)
			rt[i] = a.taps2[(size_t)r * n2 + i];
		if (cf.mode == WR_MODE_AM)
			demod_stage_simple<WR_MODE_AM, kThreads>(ch, xr, s, tid, i0, L, n2 - 1);
		else if (cf.mode == WR_MODE_USB)
			demod_stage_simple<WR_MODE_USB, kThreads>(ch, xr, s, tid, i0, L, n2 - 1);
		else
			demod_stage_simple<WR_MODE_LSB, kThreads>(ch, xr, s, tid, i0, L, n2 - 1);
	} else
	for (unsigned ib = tid; ib < L; ib += kDemodRounds * kThreads) {
		// operands of up to kDemodRounds samples first: their loads are in flight together
		float2 cur[kDemodRounds], prv[kDemodRounds];
		float hist[kDemodRounds];
		#pragma unroll
		for (int u = 0; u < kDemodRounds; u++) {
			const unsigned il = ib + u * kThreads, i = i0 + il;
			cur[u] = prv[u] = make_float2(0.0f, 0.0f);
			hist[u] = 0.0f;
			if (il < L) {
				if (i < n2 - 1) {
					hist[u] = xr[i];
				} else {
					const unsigned k = i - (n2 - 1);
					cur[u] = ch[k];
					if (k)
						prv[u] = ch[k - 1];
				}
			}
		}
		if (ib == tid) {
			if (tid < n2)
				rt[tid] = tap0;
			for (unsigned i = tid + kThreads; i < n2; i += kThreads)
				rt[i] = a.taps2[(size_t)r * n2 + i];
		}
		#pragma unroll
		for (int u = 0; u < kDemodRounds; u++) {
			const unsigned il = ib + u * kThreads, i = i0 + il;
			if (il < L) {
				float v = hist[u];
				if (i >= n2 - 1) {
					v = demod(cf.mode, cur[u], i == n2 - 1 ? prev0 : prv[u]);
					xr[i] = v;
				}
				s[il] = v;
			}
		}
	}
	__syncthreads();
	// Audio FIR, two outputs per thread (o and o + H) in the two halves of packed f32x2 registers:
	// each output's taps in the reference's order (lowpass.cxx:151-159), products and sums rounded
	// separately (mul2_rn_exact), taps four per 128-bit load.
	const float2 nz = make_float2(a.negzero, a.negzero);
	if (d2 == 1 && (n2 & 3u) == 0 && a.TK >= 8u * 96u) {
		// Large tiles at the demodulator's own rate (cfg3: 1024 receivers x 2048 samples): EIGHT
		// consecutive outputs per thread over a window of samples that slides through registers --
		// per four taps one 128-bit load of samples and one of taps serve 32 output-taps (the
		// two-outputs path below: 2.25 loads per 2), each output in its own accumulator and the
		// reference's order (taps ascending, products and sums rounded separately).  Outputs past
		// the tile's end read the tap buffer behind the window and are not stored.
		for (unsigned o = tid * 8u; o < nout; o += kThreads * 8u) {
			const float *p = s + o;
			float acc[8], w[12];
			#pragma unroll
			for (int i = 0; i < 8; i++)
				acc[i] = 0.0f;
			{
				const float4 w0 = *reinterpret_cast<const float4*>(p), w1 = *reinterpret_cast<const float4*>(p + 4);
				w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w;
				w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
			}
			#pragma unroll 4
			for (unsigned j = 0; j < n2; j += 4) {
				const float4 wn = *reinterpret_cast<const float4*>(p + j + 8);
				const float4 c = *reinterpret_cast<const float4*>(rt + j);
				w[8] = wn.x; w[9] = wn.y; w[10] = wn.z; w[11] = wn.w;
				const float cc[4] = { c.x, c.y, c.z, c.w };
				#pragma unroll
				for (int t = 0; t < 4; t++) {
					#pragma unroll
					for (int i = 0; i < 8; i++)
						acc[i] = __fadd_rn(acc[i], __fmul_rn(cc[t], w[i + t]));
				}
				#pragma unroll
				for (int i = 0; i < 8; i++)
					w[i] = w[i + 4];
			}
			float *dst = a.audio + (size_t)r * a.audio_stride + m0 + o;
			#pragma unroll
			for (int i = 0; i < 8; i++)
				if (o + (unsigned)i < nout)
					dst[i] = __fmul_rn(acc[i], a.out_scale);
		}
		__syncthreads();
		continue;
	}
	const unsigned H = (nout + 1) / 2;
	for (unsigned o = tid; o < H; o += kThreads) {
		const bool two = o + H < nout;
		const float *p0 = s + (size_t)o * d2;
		const float *p1 = s + (size_t)(two ? o + H : o) * d2;
		float2 acc = make_float2(0.0f, 0.0f);
		unsigned j = 0;
		for (; j + 3 < n2; j += 4) {
			const float4 c = *reinterpret_cast<const float4*>(rt + j);
			acc = __fadd2_rn(acc, mul2_rn_exact(make_float2(c.x, c.x), make_float2(p0[j], p1[j]), nz));
			acc = __fadd2_rn(acc, mul2_rn_exact(make_float2(c.y, c.y), make_float2(p0[j + 1], p1[j + 1]), nz));
			acc = __fadd2_rn(acc, mul2_rn_exact(make_float2(c.z, c.z), make_float2(p0[j + 2], p1[j + 2]), nz));
			acc = __fadd2_rn(acc, mul2_rn_exact(make_float2(c.w, c.w), make_float2(p0[j + 3], p1[j + 3]), nz));
		}
		for (; j < n2; j++)
			acc = __fadd2_rn(acc, mul2_rn_exact(make_float2(rt[j], rt[j]), make_float2(p0[j], p1[j]), nz));
		a.audio[(size_t)r * a.audio_stride + m0 + o] = __fmul_rn(acc.x, a.out_scale);
		if (two)
			a.audio[(size_t)r * a.audio_stride + m0 + o + H] = __fmul_rn(acc.y, a.out_scale);
	}
	__syncthreads();    // the staged window is rewritten by the next item
	}
	demod_audio_done(a, tid);
}

// ------------------------------------------------------------------ host side ----

typedef void (*V2Kernel)(const ChanArgs, const V2Args);

// Fully unrolled FIR for the geometries of the BASELINE configs and the reference's shipped
// point; anything else takes the generic tap loop.
inline V2Kernel v2_pick_kernel(unsigned n1, unsigned d1)
{
	if (n1 == 64 && d1 == 10) return chan_kernel_v2<kV2Threads, kV2J, true, 64, 10>;
	if (n1 == 127 && d1 == 50) return chan_kernel_v2<kV2Threads, kV2J, true, 127, 50>;
	if (n1 == 255 && d1 == 50) return chan_kernel_v2<kV2Threads, kV2J, true, 255, 50>;
	if (n1 == 127 && d1 == 40) return chan_kernel_v2<kV2Threads, kV2J, true, 127, 40>;
	if (d1 % 2 == 0) return chan_kernel_v2<kV2Threads, kV2J, true, 0, 0>;
	return chan_kernel_v2<kV2Threads, kV2J, false, 0, 0>;
}

struct V2Plan {
	bool ok = false;
	V2Kernel kernel = nullptr;
	bool tableStale = true;
	bool groupsStale = true;
	int device = 0;
	int numSMs = 0;
	unsigned n1 = 0, d1 = 0;
	unsigned TK = 0, RB = 2, NC = 0, Ucap = 0, A = 0, off = 0, Dp = 0, magicD = 0, Lcap = 0;
	size_t smemBytes = 0;
	int16_t *d_delta = nullptr;
	wr::LoCoef coef = {};
	unsigned *d_order = nullptr;
	int4 *d_groups = nullptr;
	unsigned nGroups = 0;
	unsigned capR = 0;
};

inline bool v2_supported(const V2Plan &p) { return p.ok; }

inline void v2_destroy(V2Plan &p)
{
	cudaFree(p.d_delta);
	cudaFree(p.d_order);
	cudaFree(p.d_groups);
	p.d_delta = nullptr;
	p.d_order = nullptr;
	p.d_groups = nullptr;
	p.ok = false;
}

// Geometry of the tile for (n1, d1); leaves p.ok false when v2 cannot serve it (v1 is used).
inline int v2_init(V2Plan &p, int device, unsigned n1, unsigned d1)
{
	p.device = device;
	p.n1 = n1;
	p.d1 = d1;
	p.ok = false;
	cudaDeviceProp prop;
	WR_CUDA(cudaGetDeviceProperties(&prop, device));
	p.numSMs = prop.multiProcessorCount;
	if (const char *e = getenv("WR_V2_RB"))
		p.RB = std::max(1, atoi(e));
	p.A = (n1 - 1 + d1 - 1) / d1;
	p.off = p.A * d1 - (n1 - 1);
	p.Dp = (d1 % 2 == 0) ? d1 + 1 : d1;    // odd period: FIR lanes (stride Dp float2) spread over all banks
	p.magicD = (unsigned)((0x100000000ull + d1 - 1) / d1);
	// Two tile buffers next to the table: the tile holds as many frames as shared memory allows
	// (a multiple of 256, at most kV2J frames per mixer thread).
	const unsigned warps = kV2Threads / 32;
	for (unsigned ucap = kV2J * 32 * (warps - 4); ucap >= 1024; ucap -= 256) {
		const long periods = (long)(ucap / d1) - 2 - (long)p.A;
		if (periods < 1)
			return WR_OK;                  // decimation too large for one tile: v1 serves it
		const unsigned tk = (unsigned)std::min<long>(periods, 1024);
		const unsigned lcap = ucap + ucap / d1 + 2;
		const size_t smem = kV2TableBytes + sizeof(float2) * (2 * (size_t)lcap + n1 + 2);
		if (smem > (size_t)prop.sharedMemPerBlockOptin)
			continue;
		// mixers: exactly as many warps as cover the tile with kV2J frames per thread (no idle
		// slots in the unrolled mix loop); the remaining warps (>= 4, one per scheduler) consume
		const unsigned nm = (ucap + kV2J * 32 - 1) / (kV2J * 32);
		if (nm + 4 > warps)
			continue;
		unsigned nc = warps - nm;
		if (const char *e = getenv("WR_V2_NC"))
			nc = (unsigned)std::min<int>(std::max(4, atoi(e)), (int)(warps - nm));
		p.NC = nc; p.Ucap = ucap; p.TK = tk; p.Lcap = lcap; p.smemBytes = smem;
		break;
	}
	if (!p.NC)
		return WR_OK;
	for (unsigned u = 0; u < p.Ucap + d1; u++) // the reciprocal must be exact over the tile
		if ((unsigned)(((unsigned long long)u * p.magicD) >> 32) != u / d1)
			return WR_OK;
	p.kernel = v2_pick_kernel(n1, d1);
	WR_CUDA(cudaFuncSetAttribute(p.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smemBytes));
	WR_CUDA(cudaMalloc(&p.d_delta, kV2TableBytes));
	p.tableStale = true;
	p.groupsStale = true;
	p.ok = true;
	return WR_OK;
}

// (Re)compress the NCO table after it changed; v2 turns itself off if it is not representable.
inline int v2_set_table(V2Plan &p, const float *h_table, cudaStream_t st)
{
	if (!p.d_delta)
		return WR_OK;
	std::vector<int16_t> delta(WR_SINTABLE_SIZE);
	if (!wr::lo_compress(h_table, delta.data(), &p.coef)) {
		p.ok = false;
		return WR_OK;
	}
	// lay the corrections out the way the kernels address them (signed index, padded rows)
	std::vector<int16_t> padded(kV2TableBytes / 2, 0);
	for (int sidx = -32768; sidx < 32768; sidx++)
		padded[(size_t)(((sidx * kLoPadMul) >> kLoPadShift) - kLoSlotMin)] = delta[(uint16_t)sidx];
	WR_CUDA(cudaMemcpyAsync(p.d_delta, padded.data(), kV2TableBytes, cudaMemcpyHostToDevice, st));
	WR_CUDA(cudaStreamSynchronize(st)); // `padded` is a local
	p.tableStale = false;
	return WR_OK;
}

// Receivers sorted by stream and cut into groups of <= RB that share one.
inline int v2_set_groups(V2Plan &p, const RxConf *h_conf, unsigned R, cudaStream_t st)
{
	if (!p.ok)
		return WR_OK;
	std::vector<unsigned> order(R);
	for (unsigned r = 0; r < R; r++)
		order[r] = r;
	std::stable_sort(order.begin(), order.end(),
			[&](unsigned x, unsigned y) { return h_conf[x].stream < h_conf[y].stream; });
	std::vector<int4> groups;
	for (unsigned i = 0; i < R;) {
		unsigned n = 1;
		while (i + n < R && n < p.RB && h_conf[order[i + n]].stream == h_conf[order[i]].stream)
			n++;
		groups.push_back(make_int4((int)i, (int)n, (int)h_conf[order[i]].stream, 0));
		i += n;
	}
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
	p.groupsStale = false;
	return WR_OK;
}

inline int v2_launch_chan(V2Plan &p, ChanArgs &ca, unsigned R, cudaStream_t st, unsigned long long *launches)
{
	(void)R;
	V2Args v;
	const wr::LoCoef &k = p.coef;
	v.delta = p.d_delta;
	v.eps = k.eps;
	v.order = p.d_order;
	v.groups = p.d_groups;
	v.nGroups = p.nGroups;
	v.TK = p.TK;
	v.NC = p.NC;
	v.ntiles = std::max(1u, (ca.M1 + p.TK - 1) / p.TK);
	v.nItems = v.ntiles * p.nGroups;
	v.A = p.A;
	v.off = p.off;
	v.Dp = p.Dp;
	v.magicD = p.magicD;
	v.Ucap = p.Ucap;
	v.Lcap = p.Lcap;
	v.negzero = -0.0f;
	ca.TK = p.TK;
	ca.ntiles = v.ntiles;
	const unsigned grid = std::min<unsigned>(v.nItems, (unsigned)p.numSMs);
	p.kernel<<<grid, kV2Threads, p.smemBytes, st>>>(ca, v);
	(*launches)++;
	if (cudaGetLastError() != cudaSuccess)
		return WR_ECUDA;
	return WR_OK;
}

} // namespace wrd
