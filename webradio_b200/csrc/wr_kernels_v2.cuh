// wr_kernels_v2.cuh -- fused NCO mix + decimating channel FIR + demod with the NCO table
// RESIDENT IN SHARED MEMORY (K1+K2+K3 of SURVEY.md 2a, second generation).
//
// Why: v1 gathers sin/cos from the 256 KiB float table through L1/L2 -- two scattered 4-byte
// loads per receiver-frame, which bound the kernel at ~70% L1 utilisation (profiles/r01_v1_*).
// The table cannot fit an SM as floats, but it can as 16-bit corrections to a closed-form base
// (wr_lo.h): 128 KiB per CTA, bit-exact by construction and verified on the host.
//
// Shape: persistent grid, one 512-thread CTA per SM, each CTA loops over work items
//   item = (group of <= RB receivers listening to the same tuner stream, tile of TK outputs).
// Per item the raw IQ the tile touches is loaded ONCE into registers (coalesced float2 loads,
// J frames per thread) and re-used for every receiver of the group; per receiver the CTA
//   1. mixes its frames (LO from the shared-memory table) into a shared-memory tile laid out
//      period-major with an odd padded period, so that the FIR's lanes hit distinct banks,
//   2. runs one thread per output over the taps in the reference's order (packed f32x2
//      mul / add: I and Q in one instruction, each product and sum still rounded separately),
//   3. demodulates in the epilogue and writes 4 bytes per channel-rate sample.
// The last tile of a receiver also produces the carried state (mixed history, phase, prev I/Q).
#pragma once

#include "wr_bank.cuh"
#include "wr_device.cuh"
#include "wr_lo.h"

#include <algorithm>
#include <cstdlib>
#include <vector>

namespace wrd {

constexpr int kV2Threads = 512;
constexpr int kV2J = 16;                          // raw frames held in registers per thread
constexpr unsigned kV2Ucap = kV2Threads * kV2J;   // frames one tile may span
constexpr unsigned kV2TableBytes = WR_SINTABLE_SIZE * 2;

struct V2Args {
	const int16_t *delta;     // [65536] corrections (HBM copy, staged to shared memory per CTA)
	float eps;                // table-dependent clamp of the base polynomial (wr_lo.h)
	const unsigned *order;    // receivers sorted by stream
	const int2 *groups;       // {first index into order, count}
	unsigned nGroups;
	unsigned TK, ntiles, nItems;
	unsigned A;               // periods of history an output reaches back: ceil((n1-1)/d1)
	unsigned off;             // A*d1 - (n1-1): offset of an output's first tap in its period
	unsigned Dp;              // padded period (d1 or d1+1, odd)
	unsigned magicD;          // ceil(2^32 / d1): u / d1 == umulhi(u, magicD) for u < kV2Ucap
	unsigned Lcap;            // float2 slots of the mixed tile
	float negzero;            // -0.0f, deliberately opaque to the compiler (see mul2_rn_exact)
};

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

// dmid32 = shared-space byte address of the MIDDLE of the correction table (entry for s = 0)
__device__ __forceinline__ float lo_value(int s, uint32_t dmid32, float eps)
{
	int d;
	asm("ld.shared.s16 %0, [%1];" : "=r"(d) : "r"(dmid32 + 2u * (uint32_t)s));
	return __int_as_float(__float_as_int(lo_base(s, eps)) + d);
}

// sin/cos of the NCO for a 31-bit phase (bit 31 of `p` is garbage and ignored), exactly the
// reference's sinTable[sinidx], sinTable[cosidx] (downconverter.cxx:100-102).
__device__ __forceinline__ void lo_sincos(uint32_t p, uint32_t dmid32, float eps, float &sn, float &cs)
{
	const int ss = (int)(p << 1) >> 16;                    // table index as signed 16 bit
	const int sc = (int)((p + 0x20000000u) << 1) >> 16;    // + a quarter turn
	sn = lo_value(ss, dmid32, eps);
	cs = lo_value(sc, dmid32, eps);
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

template <int NT, int J, bool kPad>
__global__ void __launch_bounds__(NT, 1) chan_kernel_v2(const ChanArgs a, const V2Args v)
{
	extern __shared__ __align__(16) unsigned char wr_smem_v2[];
	int16_t *dtab = reinterpret_cast<int16_t*>(wr_smem_v2);
	float2 *s = reinterpret_cast<float2*>(wr_smem_v2 + kV2TableBytes);
	float2 *rt2 = s + v.Lcap;                 // taps, each duplicated {c, c} for the packed multiply
	float2 *co = rt2 + a.n1;                  // channel outputs of the tile

	const unsigned tid = threadIdx.x;
	const unsigned n1 = a.n1, d1 = a.d1;
	const float2 nz = make_float2(v.negzero, v.negzero);
	const float eps = v.eps;
	const uint32_t smem32 = (uint32_t)__cvta_generic_to_shared(wr_smem_v2);
	const uint32_t dmid32 = smem32 + 65536u;          // entry of s = 0
	const uint32_t tile32 = smem32 + kV2TableBytes;   // mixed tile
	const unsigned padMagic = kPad ? v.magicD : 0u;   // pos(u) = u + u / d1 when the period is padded

	// stage the correction table (128 KiB) once per CTA
	{
		const uint4 *g = reinterpret_cast<const uint4*>(v.delta);
		uint4 *d = reinterpret_cast<uint4*>(dtab);
		#pragma unroll 4
		for (unsigned i = tid; i < kV2TableBytes / 16; i += NT)
			d[i] = __ldg(g + i);
	}
	__syncthreads();

	for (unsigned item = blockIdx.x; item < v.nItems; item += gridDim.x) {
		const unsigned tile = item % v.ntiles;
		const int2 grp = v.groups[item / v.ntiles];
		const unsigned stream = a.conf[v.order[grp.x]].stream;
		const float2 *__restrict__ in = a.iq + (size_t)stream * a.stream_stride;

		// ---- tile geometry (identical for every receiver of the group) ----
		const bool last = (tile == v.ntiles - 1);
		const unsigned k0 = tile * v.TK;
		const unsigned kend = min(k0 + v.TK, a.M1);
		const unsigned kstart = k0 ? k0 - 1 : 0;   // one extra output: the FM look-back sample
		const unsigned extra = k0 - kstart;
		const unsigned nout = kend > kstart ? kend - kstart : 0;
		// local frame coordinate u = f - fb0, fb0 aligned to a decimation period
		const int fb0 = ((int)kstart - (int)v.A) * (int)d1;
		// frames the outputs need: up to (kend-1)*d1; the last tile runs on to F-1 (history)
		unsigned U = nout ? (unsigned)((int)((kend - 1) * d1) - fb0 + 1) : 0;
		if (last)
			U = (unsigned)((int)a.F - fb0);
		const int jeff = (int)((U + NT - 1) / NT);          // uniform trip count of the mix loop
		const unsigned nhist = fb0 < 0 ? (unsigned)(-fb0) : 0; // leading frames that are history

		// ---- raw IQ of the tile, once, into registers ----
		float2 raw[J];
		#pragma unroll
		for (int j = 0; j < J; j++) {
			const unsigned u = tid + j * NT;
			const int f = fb0 + (int)u;
			raw[j] = (u < U && f >= 0) ? __ldg(in + f) : make_float2(0.0f, 0.0f);
		}

		for (int gi = 0; gi < grp.y; gi++) {
			const unsigned r = v.order[grp.x + gi];
			const RxConf cf = a.conf[r];
			const RxState st = a.st_in[r];

			// ---- 1. mix into the shared tile (slots past U or before the block hold zeros) ----
			const uint32_t pstep = (uint32_t)cf.step * NT;
			uint32_t p = st.phase + (uint32_t)(fb0 + (int)tid) * (uint32_t)cf.step;
			#pragma unroll
			for (int j = 0; j < J; j++) {
				if (j < jeff) {
					const unsigned u = tid + j * NT;
					float sn, cs;
					lo_sincos(p, dmid32, eps, sn, cs);
					const float2 m = mix(raw[j], cs, sn);
					const unsigned pos = kPad ? u + __umulhi(u, padMagic) : u;
					asm volatile("st.shared.v2.f32 [%0], {%1, %2};" :: "r"(tile32 + 8u * pos), "f"(m.x), "f"(m.y) : "memory");
					p += pstep;
				}
			}
			if (nhist) { // first tile: frames before the block come from the carried history
				for (unsigned u = tid; u < nhist; u += NT) {
					const unsigned pos = kPad ? u + __umulhi(u, padMagic) : u;
					const int hidx = (int)(n1 - 1) + fb0 + (int)u;
					s[pos] = hidx >= 0 ? a.hist_in[(size_t)r * (n1 - 1) + hidx] : make_float2(0.0f, 0.0f);
				}
			}
			for (unsigned i = tid; i < n1; i += NT) {
				const float c = a.taps1[(size_t)r * n1 + i];
				rt2[i] = make_float2(c, c);
			}
			__syncthreads();

			// ---- 2. FIR: one thread per output, taps in the reference's order ----
			for (unsigned o = tid; o < nout; o += NT) {
				float2 acc = make_float2(0.0f, 0.0f);
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
				co[o] = acc;
			}
			if (last) {
				// carried state: the last n1-1 mixed frames of [history | block]
				for (unsigned i = tid; i + 1 < n1; i += NT) {
					const unsigned u = (unsigned)((int)a.F - (int)(n1 - 1) + (int)i - fb0);
					const unsigned pos = kPad ? u + __umulhi(u, padMagic) : u;
					a.hist_out[(size_t)r * (n1 - 1) + i] = s[pos];
				}
				if (tid == 0) {
					a.st_out[r].phase = phase_at(st.phase, cf.step, a.F);
					if (a.M1 == 0) {
						a.st_out[r].prev_i = st.prev_i;
						a.st_out[r].prev_q = st.prev_q;
					}
				}
			}
			__syncthreads();

			// ---- 3. demodulator epilogue ----
			for (unsigned o = tid + extra; o < nout; o += NT) {
				const unsigned k = kstart + o;
				const float2 cur = co[o];
				const float2 prev = o ? co[o - 1] : make_float2(st.prev_i, st.prev_q);
				a.demod[(size_t)r * a.dstride + a.demod_off + k] = demod(cf.mode, cur, prev);
				if (a.chan)
					a.chan[(size_t)r * a.chan_stride + k] = cur;
				if (k == a.M1 - 1) {
					a.st_out[r].prev_i = cur.x;
					a.st_out[r].prev_q = cur.y;
				}
			}
			// the next receiver's mix overwrites s / rt2 only after every thread passed the
			// barrier above; co is rewritten only after the next receiver's first barrier
		}
		__syncthreads(); // co of the last receiver is read above while the next item refills s
	}
}

// ------------------------------------------------------------------ host side ----

struct V2Plan {
	bool ok = false;
	bool tableStale = true;
	bool groupsStale = true;
	int device = 0;
	int numSMs = 0;
	unsigned n1 = 0, d1 = 0;
	unsigned TK = 0, RB = 2, A = 0, off = 0, Dp = 0, magicD = 0, Lcap = 0;
	size_t smemBytes = 0;
	int16_t *d_delta = nullptr;
	wr::LoCoef coef = {};
	unsigned *d_order = nullptr;
	int2 *d_groups = nullptr;
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
	for (unsigned u = 0; u < kV2Ucap + d1; u++) // the reciprocal must be exact over the tile
		if ((unsigned)(((unsigned long long)u * p.magicD) >> 32) != u / d1)
			return WR_OK;
	// frames a tile spans: (TK + 1 + A) periods, plus up to two more on the last tile
	const long periods = (long)(kV2Ucap / d1) - 3 - (long)p.A;
	if (periods < 1)
		return WR_OK;
	unsigned tk = (unsigned)std::min<long>(periods, 1023);
	if (tk + 1 >= 32)
		tk = ((tk + 1) / 32) * 32 - 1;     // TK + 1 outputs (with the FM helper) fill whole warps
	p.TK = tk;
	p.Lcap = kV2Ucap + kV2Ucap / d1 + 2;
	p.smemBytes = kV2TableBytes + sizeof(float2) * ((size_t)p.Lcap + n1 + tk + 2);
	if (p.smemBytes > (size_t)prop.sharedMemPerBlockOptin)
		return WR_OK;
	if (p.Dp != d1)
		WR_CUDA(cudaFuncSetAttribute(chan_kernel_v2<kV2Threads, kV2J, true>,
				cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smemBytes));
	else
		WR_CUDA(cudaFuncSetAttribute(chan_kernel_v2<kV2Threads, kV2J, false>,
				cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smemBytes));
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
	// the kernels index the table by the SIGNED 16-bit index from its middle (entry s lives at
	// slot s + 32768): rotate the unsigned-indexed host array by half a turn
	std::rotate(delta.begin(), delta.begin() + 32768, delta.end());
	WR_CUDA(cudaMemcpyAsync(p.d_delta, delta.data(), kV2TableBytes, cudaMemcpyHostToDevice, st));
	WR_CUDA(cudaStreamSynchronize(st)); // `delta` is a local
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
	std::vector<int2> groups;
	for (unsigned i = 0; i < R;) {
		unsigned n = 1;
		while (i + n < R && n < p.RB && h_conf[order[i + n]].stream == h_conf[order[i]].stream)
			n++;
		groups.push_back(make_int2((int)i, (int)n));
		i += n;
	}
	if (R > p.capR) {
		cudaFree(p.d_order);
		cudaFree(p.d_groups);
		p.d_order = nullptr;
		p.d_groups = nullptr;
		WR_CUDA(cudaMalloc(&p.d_order, sizeof(unsigned) * R));
		WR_CUDA(cudaMalloc(&p.d_groups, sizeof(int2) * R));
		p.capR = R;
	}
	WR_CUDA(cudaMemcpyAsync(p.d_order, order.data(), sizeof(unsigned) * R, cudaMemcpyHostToDevice, st));
	WR_CUDA(cudaMemcpyAsync(p.d_groups, groups.data(), sizeof(int2) * groups.size(), cudaMemcpyHostToDevice, st));
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
	v.ntiles = std::max(1u, (ca.M1 + p.TK - 1) / p.TK);
	v.nItems = v.ntiles * p.nGroups;
	v.A = p.A;
	v.off = p.off;
	v.Dp = p.Dp;
	v.magicD = p.magicD;
	v.Lcap = p.Lcap;
	v.negzero = -0.0f;
	ca.TK = p.TK;
	ca.ntiles = v.ntiles;
	const unsigned grid = std::min<unsigned>(v.nItems, (unsigned)p.numSMs);
	if (p.Dp != p.d1)
		chan_kernel_v2<kV2Threads, kV2J, true><<<grid, kV2Threads, p.smemBytes, st>>>(ca, v);
	else
		chan_kernel_v2<kV2Threads, kV2J, false><<<grid, kV2Threads, p.smemBytes, st>>>(ca, v);
	(*launches)++;
	return WR_OK;
}

} // namespace wrd
