// wr_atan2f.h -- the host C library's atan2f, restated operation for operation.
//
// The reference's FM discriminator calls atan2f from the box's libm (demodulator.cxx:97).  On
// the glibc this image ships (2.39, x86-64) that is the fdlibm-derived e_atan2f.c + s_atanf.c:
// an argument reduction at 7/16, 11/16, 19/16, 39/16 and a degree-11 odd/even split polynomial,
// all in float, every product, sum and quotient rounded on its own.  Written here once, with
// the arithmetic spelled through WR_F* so that
//   * the device build uses the _rn intrinsics (no FMA contraction, IEEE division), and
//   * the host build (wr_host.cpp, -ffp-contract=off) is plain C,
// the two are the same function bit for bit, and tests/test_atan2f.py pins the host twin against
// the libm actually installed on the box (hundreds of millions of arguments, every branch
// threshold).  If a box ever ships a different atan2f that test fails and says so; the FM
// parity tests then fall back to their ULP tolerance.
//
// Constants are the decimal literals of the fdlibm sources (their hex comments are off by one
// unit for aT[0]: the compiler's reading of the decimal is what libm contains).
#pragma once

#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define WR_AT_FN __device__ __forceinline__
#define WR_FMUL(a, b) __fmul_rn((a), (b))
#define WR_FADD(a, b) __fadd_rn((a), (b))
#define WR_FSUB(a, b) __fsub_rn((a), (b))
#define WR_FDIV(a, b) __fdiv_rn((a), (b))
#define WR_F2U(f) __float_as_uint(f)
#define WR_U2F(u) __uint_as_float(u)
#else
#include <string.h>
#define WR_AT_FN static inline
#define WR_FMUL(a, b) ((a) * (b))
#define WR_FADD(a, b) ((a) + (b))
#define WR_FSUB(a, b) ((a) - (b))
#define WR_FDIV(a, b) ((a) / (b))
static inline uint32_t wr_at_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float wr_at_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#define WR_F2U(f) wr_at_f2u(f)
#define WR_U2F(u) wr_at_u2f(u)
#endif

namespace wrat {

// s_atanf.c
WR_AT_FN float atanf_glibc(float x)
{
	const float hi0 = 4.6364760399e-01f, hi1 = 7.8539812565e-01f, hi2 = 9.8279368877e-01f, hi3 = 1.5707962513e+00f;
	const float lo0 = 5.0121582440e-09f, lo1 = 3.7748947079e-08f, lo2 = 3.4473217170e-08f, lo3 = 7.5497894159e-08f;
	const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f, aT3 = -1.1111110449e-01f,
			aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f, aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f,
			aT8 = 4.9768779427e-02f, aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
	const int32_t hx = (int32_t)WR_F2U(x);
	const int32_t ix = hx & 0x7fffffff;
	float hi, lo;
	int id;
	if (ix >= 0x4c000000) {                 // |x| >= 2^25
		if (ix > 0x7f800000)
			return WR_FADD(x, x);           // NaN
		const float r = WR_FADD(hi3, lo3);
		return hx > 0 ? r : WR_FSUB(-hi3, lo3);
	}
	if (ix < 0x3ee00000) {                  // |x| < 7/16
		if (ix < 0x31000000)                // |x| < 2^-29
			return x;
		id = -1;
		hi = lo = 0.0f;
	} else {
		x = WR_U2F((uint32_t)ix);           // fabsf
		if (ix < 0x3f980000) {              // |x| < 19/16
			if (ix < 0x3f300000) {          // 7/16 <= |x| < 11/16
				id = 0; hi = hi0; lo = lo0;
				x = WR_FDIV(WR_FSUB(WR_FMUL(2.0f, x), 1.0f), WR_FADD(2.0f, x));
			} else {                        // 11/16 <= |x| < 19/16
				id = 1; hi = hi1; lo = lo1;
				x = WR_FDIV(WR_FSUB(x, 1.0f), WR_FADD(x, 1.0f));
			}
		} else {
			if (ix < 0x401c0000) {          // |x| < 39/16
				id = 2; hi = hi2; lo = lo2;
				x = WR_FDIV(WR_FSUB(x, 1.5f), WR_FADD(1.0f, WR_FMUL(1.5f, x)));
			} else {                        // 39/16 <= |x| < 2^25
				id = 3; hi = hi3; lo = lo3;
				x = WR_FDIV(-1.0f, x);
			}
		}
	}
	const float z = WR_FMUL(x, x);
	const float w = WR_FMUL(z, z);
	// the sum of aT[i] z^(i+1), split into odd and even halves
	float s1 = WR_FADD(aT8, WR_FMUL(w, aT10));
	s1 = WR_FADD(aT6, WR_FMUL(w, s1));
	s1 = WR_FADD(aT4, WR_FMUL(w, s1));
	s1 = WR_FADD(aT2, WR_FMUL(w, s1));
	s1 = WR_FADD(aT0, WR_FMUL(w, s1));
	s1 = WR_FMUL(z, s1);
	float s2 = WR_FADD(aT7, WR_FMUL(w, aT9));
	s2 = WR_FADD(aT5, WR_FMUL(w, s2));
	s2 = WR_FADD(aT3, WR_FMUL(w, s2));
	s2 = WR_FADD(aT1, WR_FMUL(w, s2));
	s2 = WR_FMUL(w, s2);
	const float xs = WR_FMUL(x, WR_FADD(s1, s2));
	if (id < 0)
		return WR_FSUB(x, xs);
	const float r = WR_FSUB(hi, WR_FSUB(WR_FSUB(xs, lo), x));
	return hx < 0 ? -r : r;
}

// e_atan2f.c
WR_AT_FN float atan2f_glibc(float y, float x)
{
	const float tiny = 1.0e-30f;
	const float pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
	const int32_t hx = (int32_t)WR_F2U(x), hy = (int32_t)WR_F2U(y);
	const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
	if (ix > 0x7f800000 || iy > 0x7f800000)
		return WR_FADD(x, y);                                   // NaN
	if (hx == 0x3f800000)
		return atanf_glibc(y);                                  // x = 1
	const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);          // 2 * sign(x) + sign(y)
	if (iy == 0) {
		if (m < 2)
			return y;                                           // atan(+-0, +anything) = +-0
		return m == 2 ? WR_FADD(pi, tiny) : WR_FSUB(-pi, tiny); // atan(+-0, -anything) = +-pi
	}
	if (ix == 0)
		return hy < 0 ? WR_FSUB(-pi_o_2, tiny) : WR_FADD(pi_o_2, tiny);
	if (ix == 0x7f800000) {
		if (iy == 0x7f800000) {
			switch (m) {
			case 0: return WR_FADD(pi_o_4, tiny);
			case 1: return WR_FSUB(-pi_o_4, tiny);
			case 2: return WR_FADD(WR_FMUL(3.0f, pi_o_4), tiny);
			default: return WR_FSUB(WR_FMUL(-3.0f, pi_o_4), tiny);
			}
		}
		switch (m) {
		case 0: return 0.0f;
		case 1: return -0.0f;
		case 2: return WR_FADD(pi, tiny);
		default: return WR_FSUB(-pi, tiny);
		}
	}
	if (iy == 0x7f800000)
		return hy < 0 ? WR_FSUB(-pi_o_2, tiny) : WR_FADD(pi_o_2, tiny);
	const int32_t k = (iy - ix) >> 23;
	float z;
	if (k > 60)
		z = WR_FADD(pi_o_2, WR_FMUL(0.5f, pi_lo));              // |y/x| > 2^60
	else if (hx < 0 && k < -60)
		z = 0.0f;                                               // |y|/x < -2^60
	else
		z = atanf_glibc(WR_U2F(WR_F2U(WR_FDIV(y, x)) & 0x7fffffffu));
	switch (m) {
	case 0: return z;                                           // atan(+, +)
	case 1: return WR_U2F(WR_F2U(z) ^ 0x80000000u);             // atan(-, +)
	case 2: return WR_FSUB(pi, WR_FSUB(z, pi_lo));              // atan(+, -)
	default: return WR_FSUB(WR_FSUB(z, pi_lo), pi);             // atan(-, -)
	}
}

} // namespace wrat
