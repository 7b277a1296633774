// wr_lo.h -- exact compression of the NCO sine table so that it fits in shared memory.
//
// The reference's 65536-entry float table (256 KiB, each entry an independently rounded sinf,
// reference downconverter.cxx:49-51) does not fit one SM's shared memory, and gathering from
// L1/L2 bounds the whole chain (profiles/r01_v1_*).  The v2 kernels therefore hold, per entry,
// only a 16-bit correction (128 KiB) to a cheap closed-form base:
//
//     s  = index as signed 16 bit  (angle = pi * s / 32768)
//     u  = s * max(32768 - |s|, eps)            (odd; zero crossings at s = 0 and s = -32768)
//     B  = u * (a0 + |u| * (a1 + |u| * a2))     (fmaf, ~1e-4 relative everywhere)
//     table[index] == int_as_float(float_as_int(B) + delta[index])      BIT-EXACT
//
// B is evaluated with the same IEEE operations on host (std::fmaf) and device (__fmaf_rn), so
// the host computes and VERIFIES every delta at start-up.  eps only matters at index 32768,
// where the reference's entry is sinf((float)pi) = -8.74e-8 rather than 0: it is chosen so that
// B(-32768) lands within 16 bits of that value.  A table that cannot be reproduced exactly
// (never the reference's) simply disables the v2 kernels.
#pragma once

#include <stdint.h>
#include <vector>

namespace wr {

struct LoCoef {
	float a0, a1, a2;
	float eps;
};

// Fixed fit of sin(pi t) ~ u (a0 + a1|u| + a2 u^2), u = 32768^2 t(1-|t|)
inline LoCoef lo_coef()
{
	LoCoef k;
	k.a0 = 2.9261695289051204e-09f;
	k.a1 = 2.714416547945065e-18f;
	k.a2 = 9.777997364797523e-28f;
	k.eps = 0.0f; // table dependent, filled in by lo_compress
	return k;
}

// Host evaluation of the base (bit-identical to wrd::lo_base on the device).
float lo_base_host(int s, const LoCoef &k);

// Computes eps and delta[65536] for `table`.  Returns true iff EVERY entry is then reproduced
// bit for bit by int_as_float(float_as_int(B) + delta).
bool lo_compress(const float *table, int16_t *delta, LoCoef *coef);

} // namespace wr
