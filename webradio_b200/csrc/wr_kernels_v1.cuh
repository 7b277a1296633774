// wr_kernels_v1.cuh -- generic (any n1/d1/n2/d2) kernels of the receiver bank.
//
//   chan_kernel_v1 : K1+K2+K3 of SURVEY.md 2a fused -- NCO mix (never materialised in HBM),
//                    decimating channel FIR, demodulator epilogue.
//   audio_kernel_v1: K4 -- decimating audio FIR over the demodulated stream.
//
// Grid: (tiles + 1, receivers).  CTA (t, r), t < tiles, produces TK channel-rate outputs of
// receiver r: it mixes the (TK-1)*d1 + n1 input frames those outputs touch into shared memory
// (frames before the block start come from the mixed-history buffer), then one thread per
// output walks the taps in the reference's order.  CTA (tiles, r) is the "tail": it writes the
// state the next block needs (mixed history, NCO phase).
//
// Roofline note: per receiver-frame this kernel does 2 table gathers from L2/L1 (the 256 KiB
// table does not fit shared memory); that gather, not HBM, bounds v1.  v2 (wr_kernels_v2.cuh)
// removes it.
#pragma once

#include "wr_bank.cuh"
#include "wr_device.cuh"

namespace wrd {

__device__ __forceinline__ float2 mixed_frame_v1(const ChanArgs &a, const float2 *__restrict__ in,
		const RxConf &cf, uint32_t phase0, long f, unsigned r)
{
	if (f < 0) // before this block: mixed by the previous launch
		return a.hist_in[(size_t)r * (a.n1 - 1) + (size_t)((long)(a.n1 - 1) + f)];
	uint32_t si, ci;
	lo_indices(phase_at(phase0, cf.step, (uint32_t)f), si, ci);
	float s = __ldg(a.table + si);
	float c = __ldg(a.table + ci);
	float2 x = __ldg(in + f);
	return mix(x, c, s);
}

template <int kThreads>
__global__ void __launch_bounds__(kThreads) chan_kernel_v1(const ChanArgs a)
{
	extern __shared__ float4 wr_smem_v1[];
	const unsigned r = blockIdx.y;
	const unsigned tile = blockIdx.x;
	const unsigned tid = threadIdx.x;
	const RxConf cf = a.conf[r];
	const RxState st = a.st_in[r];
	const float2 *__restrict__ in = a.iq + (size_t)cf.stream * a.stream_stride;
	const unsigned n1 = a.n1, d1 = a.d1;

	if (tile == a.ntiles) {
		// ---- tail CTA: state for the next block ----
		// new history = last n1-1 frames of [old history | this block], mixed
		for (unsigned i = tid; i < n1 - 1; i += kThreads) {
			long f = (long)a.F - (long)(n1 - 1) + (long)i;
			a.hist_out[(size_t)r * (n1 - 1) + i] = mixed_frame_v1(a, in, cf, st.phase, f, r);
		}
		if (tid == 0) {
			a.st_out[r].phase = phase_at(st.phase, cf.step, a.F);
			if (a.M1 == 0) { // no output this block: the FM look-back sample is unchanged
				a.st_out[r].prev_i = st.prev_i;
				a.st_out[r].prev_q = st.prev_q;
			}
		}
		return;
	}

	// ---- shared memory carve-up: mixed frames | reversed taps | channel outputs ----
	const unsigned Lmax = a.TK * d1 + n1; // (TK+1 outputs - 1) * d1 + n1
	float2 *s = reinterpret_cast<float2*>(wr_smem_v1);
	float *rt = reinterpret_cast<float*>(s + Lmax);
	float2 *co = reinterpret_cast<float2*>(rt + ((n1 + 3) & ~3u));

	const unsigned k0 = tile * a.TK;
	const unsigned kend = min(k0 + a.TK, a.M1);
	// FM needs the channel sample before the tile's first one: recompute it (one extra output)
	const unsigned extra = (cf.mode == WR_MODE_FM && k0 > 0) ? 1u : 0u;
	const unsigned kstart = k0 - extra;
	const unsigned nout = kend - kstart;
	const long fbeg = (long)kstart * d1 - (long)(n1 - 1);
	const unsigned L = (nout - 1) * d1 + n1;

	for (unsigned i = tid; i < L; i += kThreads)
		s[i] = mixed_frame_v1(a, in, cf, st.phase, fbeg + (long)i, r);
	for (unsigned i = tid; i < n1; i += kThreads)
		rt[i] = a.taps1[(size_t)r * n1 + i];
	__syncthreads();

	for (unsigned o = tid; o < nout; o += kThreads) {
		float2 acc = make_float2(0.0f, 0.0f);
		const float2 *p = s + (size_t)o * d1;
		for (unsigned j = 0; j < n1; j++)
			tap2(acc, rt[j], p[j]);
		co[o] = acc;
	}
	__syncthreads();

	for (unsigned o = tid + extra; o < nout; o += kThreads) {
		const unsigned k = kstart + o;
		const float2 cur = co[o];
		float2 prev;
		if (o > 0)
			prev = co[o - 1];
		else // k == k0: only reached with k0 == 0 in FM mode, else prev is unused
			prev = make_float2(st.prev_i, st.prev_q);
		a.demod[(size_t)r * a.dstride + a.demod_off + k] = demod(cf.mode, cur, prev);
		if (a.chan)
			a.chan[(size_t)r * a.chan_stride + k] = cur;
		if (k == a.M1 - 1) { // reference demodulator.cxx:110-111: prev persists across blocks
			a.st_out[r].prev_i = cur.x;
			a.st_out[r].prev_q = cur.y;
		}
	}
}

template <int kThreads>
__global__ void __launch_bounds__(kThreads) audio_kernel_v1(const AudioArgs a)
{
	extern __shared__ float4 wr_smem_v1[];
	const unsigned r = blockIdx.y;
	const unsigned tile = blockIdx.x;
	const unsigned tid = threadIdx.x;
	const unsigned n2 = a.n2, d2 = a.d2;
	const float *__restrict__ x = a.x + (size_t)r * a.dstride;

	if (tile == a.ntiles) {
		// tail: the last n2-1 samples of [history | new] become the next block's history
		for (unsigned i = tid; i < n2 - 1; i += kThreads)
			a.x_next[(size_t)r * a.dstride + i] = x[(size_t)a.M1 + i];
		return;
	}

	const unsigned Lmax = (a.TK - 1) * d2 + n2;
	float *s = reinterpret_cast<float*>(wr_smem_v1);
	float *rt = s + ((Lmax + 3) & ~3u);

	const unsigned m0 = tile * a.TK;
	const unsigned mend = min(m0 + a.TK, a.M2);
	const unsigned nout = mend - m0;
	const unsigned L = (nout - 1) * d2 + n2;
	for (unsigned i = tid; i < L; i += kThreads)
		s[i] = x[(size_t)m0 * d2 + i];
	for (unsigned i = tid; i < n2; i += kThreads)
		rt[i] = a.taps2[(size_t)r * n2 + i];
	__syncthreads();

	for (unsigned o = tid; o < nout; o += kThreads) {
		float acc = 0.0f;
		const float *p = s + (size_t)o * d2;
		for (unsigned j = 0; j < n2; j++)
			tap1(acc, rt[j], p[j]);
		a.audio[(size_t)r * a.audio_stride + m0 + o] = __fmul_rn(acc, a.out_scale);
	}
}

} // namespace wrd
