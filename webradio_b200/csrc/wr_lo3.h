// wr_lo3.h -- second exact compression of the NCO sine table, shaped for packed f32x2 arithmetic
// (the v3 kernels evaluate the sine and the cosine entry of a frame in the two halves of one
// register pair).  Same idea as wr_lo.h -- a 16-bit correction per entry to a closed-form base
// that host and device evaluate with identical IEEE operations -- but the base avoids abs/max
// (neither exists as a packed instruction) and integer-to-float conversions:
//
//     F    = int_as_float(0x4B000000 | (index + 32768) & 0xFFFF)   = 2^23 + (s + 32768), exact;
//                                                                     s = index as signed 16 bit
//     t    = fma(F, 2^-15, -257)                = s / 32768, exact           (angle = pi t)
//     y    = t * t
//     u    = fma(t, eps, fma(-y, t, t))         (t - t^3: zero crossings at t = 0, +-1; eps: see below)
//     base = u * (c0 + y (c1 + y (c2 + y c3)))              fit of sin(pi t) / (t (1 - t^2)); or a quadratic, below
//     table[index] == int_as_float(float_as_int(base) + delta[index])        BIT-EXACT
//
// and the padded position of an entry in shared memory comes out of the same register pair:
//
//     slot = float_as_int(fma(F, 1057/1024, 3890144)) - 0x4B400000 = RN(s * 1057 / 1024)
//
// (strictly increasing in s; one padding entry per ~31 keeps a warp's arithmetic progression of
// lookups off a single bank, see wr_kernels_v2.cuh).  eps makes the one entry that is not ~sin
// representable: sinf((float)pi) = -8.74e-8 at index 32768, where t - t^3 = 0.
// The host computes eps and every delta at start-up and VERIFIES all 65536 entries; a table that
// cannot be reproduced exactly disables the v3 kernels (v2/v1 take over).
#pragma once

#include <stdint.h>

// WR_LO3_DEGREE: degree of the base polynomial in y.
// 3 (default): the minimax cubic, corrections up to 28 019 ULP.
// 2: a QUADRATIC in y is enough for 16-bit corrections -- if it is the right one.  The entries that
// need the largest corrections are not the polynomial's fault: the reference's table is sinf of an
// angle ROUNDED TO FLOAT, so just below the zero crossings (index 32767, 65531...) the entries carry
// half an ULP of pi or 2 pi of absolute error, up to 28 000 ULP of their own tiny values.  The plain
// minimax quadratic (8.7e-4 relative, 7 300 ULP) adds to that at t -> 1 and index 32767 would need 17
// bits; pinning P(1) = pi/2 (1 - 3.4e-4) and minimising the rest leaves 1.1e-3 (18 500 ULP) in the bulk
// and 23 801 ULP at worst, for one packed fma per frame less.  (Measured: DESIGN.md 5.1a.)
#ifndef WR_LO3_DEGREE
#define WR_LO3_DEGREE 3
#endif
#if WR_LO3_DEGREE == 3
#define WR_LO3_C0     3.141528367996216f
#define WR_LO3_C1    (-2.024317979812622f)
#define WR_LO3_C2     0.5156225562095642f
#define WR_LO3_C3    (-0.06206892058253288f)
#else
#define WR_LO3_C0     3.1380698680877686f
#define WR_LO3_C1    (-1.9768874645233154f)
#define WR_LO3_C2     0.4090798795223236f
#endif
#define WR_LO3_TSCALE 3.0517578125e-05f      /* 2^-15 */
#define WR_LO3_TBIAS  (-257.0f)
#define WR_LO3_SLOTK  1.0322265625f          /* 1057/1024 */
#define WR_LO3_SLOTM  3890144.0f             /* 1.5*2^23 - (2^23 + 32768) * 1057/1024 */
#define WR_LO3_SLOTBITS 0x4B400000           /* float_as_int(1.5*2^23) */
#define WR_LO3_SLOT_MIN (-33824)
#define WR_LO3_SLOT_MAX 33823

namespace wr {

struct Lo3Coef {
	float eps;
};

// Host evaluation of the base / the slot map (bit-identical to the device code in
// wr_kernels_v3.cuh: same operations, same constants).
float lo3_base_host(int s, const Lo3Coef &k);
int lo3_slot_host(int s);

// Computes eps and delta[65536] (indexed by table index) for `table`.  Returns true iff every
// entry is then reproduced bit for bit and the slot map is the expected injection.
bool lo3_compress(const float *table, int16_t *delta, Lo3Coef *coef);

} // namespace wr
