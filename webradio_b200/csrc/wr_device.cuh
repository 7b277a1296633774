// wr_device.cuh -- per-sample device arithmetic of the receiver chain, written so that every
// float operation is rounded exactly where the reference's scalar CPU loops round it.
//
// Parity rules ("hard parts"):
//   * no FMA contraction anywhere on the sample path: __fmul_rn/__fadd_rn/__fsub_rn (nvcc
//     would otherwise fuse a*b+c), matching g++ -O2 without -mfma on the reference;
//   * FIR taps are accumulated strictly in the reference's order, one accumulator per output;
//   * the NCO uses the host-built 65536-entry table (each entry an independently rounded
//     sinf) and the 31-bit integer phase accumulator.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "wr_atan2f.h"

namespace wrd {

constexpr uint32_t kPhaseMask = 0x7FFFFFFFu;   // PHASE_BITS = 31 (reference downconverter.cxx:35)
constexpr int kLookupShift = 15;               // PHASE_BITS - LOOKUP_BITS
constexpr uint32_t kLookupMask = 0xFFFFu;
constexpr uint32_t kQuarter = 16384u;          // (1 << LOOKUP_BITS) / 4

// Phase of frame f of a block that starts at phase0: the reference adds phaseStep once per
// frame and masks to 31 bits (downconverter.cxx:103); since 2^31 divides 2^32 the closed form
// (phase0 + f*step) mod 2^31 is the same integer.
__device__ __forceinline__ uint32_t phase_at(uint32_t phase0, int32_t step, uint32_t f)
{
	return (phase0 + f * (uint32_t)step) & kPhaseMask;
}

// reference downconverter.cxx:100-102: table indices from the pre-increment phase
__device__ __forceinline__ void lo_indices(uint32_t phase, uint32_t &sinidx, uint32_t &cosidx)
{
	sinidx = phase >> kLookupShift;
	cosidx = (sinidx + kQuarter) & kLookupMask;
}

// reference downconverter.cxx:109-110: multiply by the conjugate of the local oscillator.
//   I' = i*cos + q*sin ;  Q' = q*cos - i*sin     (4 rounded products, 2 rounded sums)
__device__ __forceinline__ float2 mix(float2 x, float c, float s)
{
	float2 y;
	y.x = __fadd_rn(__fmul_rn(x.x, c), __fmul_rn(x.y, s));
	y.y = __fsub_rn(__fmul_rn(x.y, c), __fmul_rn(x.x, s));
	return y;
}

// One tap of LowPass::process (reference lowpass.cxx:155-156): out[c] += coeff * sample,
// product and sum rounded separately, for the I/Q pair.
__device__ __forceinline__ void tap2(float2 &acc, float c, float2 x)
{
	acc.x = __fadd_rn(acc.x, __fmul_rn(c, x.x));
	acc.y = __fadd_rn(acc.y, __fmul_rn(c, x.y));
}

__device__ __forceinline__ void tap1(float &acc, float c, float x)
{
	acc = __fadd_rn(acc, __fmul_rn(c, x));
}

// The reference calls the host libm's atan2f (demodulator.cxx:97).  wr_atan2f.h restates the
// glibc routine operation for operation (same constants, same rounding order), so the FM
// discriminator is bit-identical to the reference on a glibc box; tests/test_atan2f.py pins the
// restatement against the installed libm.
__device__ __forceinline__ float atan2f_ref(float y, float x)
{
	return wrat::atan2f_glibc(y, x);
}

// reference demodulator.cxx:83-112.  prev is the previous channel-rate sample.
__device__ __forceinline__ float demod(int mode, float2 cur, float2 prev)
{
	switch (mode) {
	case WR_MODE_AM:
		// sqrt(i*i + q*q): float overload, IEEE correctly rounded
		return __fsqrt_rn(__fadd_rn(__fmul_rn(cur.x, cur.x), __fmul_rn(cur.y, cur.y)));
	case WR_MODE_FM: {
		// conjugate product with the previous sample, then atan2f(ii, qq) -- the real part
		// is passed as y (sic) -- and a divide chain evaluated in double:
		//   (float)((double)atan2f(ii, qq) / M_PI / 2.0)
		float ii = __fadd_rn(__fmul_rn(cur.x, prev.x), __fmul_rn(cur.y, prev.y));
		float qq = __fsub_rn(__fmul_rn(cur.y, prev.x), __fmul_rn(cur.x, prev.y));
		const double a = (double)atan2f_ref(ii, qq);
		// x / 2.0 == x * 0.5 exactly (a power of two, no double underflow in this range)
		return (float)__dmul_rn(__ddiv_rn(a, 3.14159265358979323846), 0.5);
	}
	case WR_MODE_USB:
		return __fadd_rn(cur.x, cur.y);
	default: // WR_MODE_LSB
		return __fsub_rn(cur.x, cur.y);
	}
}

// What the browser does with one waterfall bin: WaterfallHandler::doGet sends non-finite values as
// -10000.0 (reference src/web/waterfallhandler.cxx:62-68), Waterfall.update computes
// (dB + 50.0) / 25.0, times 255.0, floor, clamp to [0, 255] -- JavaScript numbers, i.e. doubles
// (reference html/waterfall.js:92-109).
__device__ __forceinline__ unsigned char waterfall_index(float db)
{
	const double v = isfinite(db) ? (double)db : -10000.0;
	const double val = floor(__dmul_rn(__ddiv_rn(__dadd_rn(v, 50.0), 25.0), 255.0));
	return (unsigned char)(val < 0.0 ? 0.0 : (val > 255.0 ? 255.0 : val));
}

} // namespace wrd
