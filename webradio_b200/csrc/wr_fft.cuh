// wr_fft.cuh -- register-blocked FFT for the spectrum kernel (K5).
//
// N = R1 * 16 * 16 points (R1 = 2 ... 32, i.e. N = 512 ... 8192), 256 threads per transform,
// three passes with ONE shared-memory buffer of N complex values:
//   pass 1  radix-R1 over the stride-256 index, in registers, straight from HBM (window fused
//           into the load), twiddled by W_N^(t*k1), written to shared memory as R1 rows of 256;
//   pass 2  radix-16 over the stride-16 index of every 256-point row, twiddled by W_256^(b*ka);
//   pass 3  radix-16 over the contiguous index; results leave the registers straight to HBM
//           (dB + fft-shift fused into the store, coalesced).
// So a transform costs one HBM read, one HBM write and two shared-memory round trips
// (algorithmic 8*hop + 4*N bytes per frame), instead of log4(N) round trips.
// The small in-register DFTs are radix-4 Stockham butterflies (two radix-4 stages = radix-16).
#pragma once

#include <cuda_runtime.h>

namespace wrfft {

// Complex add / subtract as ONE packed f32x2 instruction each (FADD2): the butterflies are 40 % of
// the spectrum kernels' instructions, and those kernels are bound by issue slots, not by the FMA pipe
// (a packed operation occupies the pipe as long as the two scalar ones it replaces).
__device__ __forceinline__ float2 cadd(float2 a, float2 b)
{
	float2 r;
	asm("{\n\t.reg .b64 x, y, z;\n\t"
			"mov.b64 x, {%2, %3};\n\tmov.b64 y, {%4, %5};\n\t"
			"add.rn.f32x2 z, x, y;\n\t"
			"mov.b64 {%0, %1}, z;\n\t}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
	return r;
}
__device__ __forceinline__ float2 csub(float2 a, float2 b)
{
	float2 r;
	asm("{\n\t.reg .b64 x, y, z;\n\t"
			"mov.b64 x, {%2, %3};\n\tmov.b64 y, {%4, %5};\n\t"
			"sub.rn.f32x2 z, x, y;\n\t"
			"mov.b64 {%0, %1}, z;\n\t}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
	return r;
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
	return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// multiply by -i (forward transform quarter turn)
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }

// forward radix-2 / radix-4 butterflies, natural order in and out
__device__ __forceinline__ void dft2(float2 &a, float2 &b)
{
	const float2 t = csub(a, b);
	a = cadd(a, b);
	b = t;
}

__device__ __forceinline__ void dft4(float2 &a, float2 &b, float2 &c, float2 &d)
{
	const float2 s02 = cadd(a, c), d02 = csub(a, c);
	const float2 s13 = cadd(b, d), t = csub(b, d);
	a = cadd(s02, s13);
	c = csub(s02, s13);
	// d02 +- (-i) t, the quarter turn folded into which component meets which
	b = make_float2(d02.x + t.y, d02.y - t.x);
	d = make_float2(d02.x - t.y, d02.y + t.x);
}

// exp(-2*pi*i*k/16), k = 0..15 (compile-time constants)
__device__ __forceinline__ float2 w16(int k)
{
	const float c1 = 0.92387953251128673848f, s1 = 0.38268343236508978178f, r = 0.70710678118654752440f;
	switch (k & 15) {
	case 0: return make_float2(1.0f, 0.0f);
	case 1: return make_float2(c1, -s1);
	case 2: return make_float2(r, -r);
	case 3: return make_float2(s1, -c1);
	case 4: return make_float2(0.0f, -1.0f);
	case 5: return make_float2(-s1, -c1);
	case 6: return make_float2(-r, -r);
	case 7: return make_float2(-c1, -s1);
	case 8: return make_float2(-1.0f, 0.0f);
	case 9: return make_float2(-c1, s1);
	case 10: return make_float2(-r, r);
	case 11: return make_float2(-s1, c1);
	case 12: return make_float2(0.0f, 1.0f);
	case 13: return make_float2(s1, c1);
	case 14: return make_float2(r, r);
	default: return make_float2(c1, s1);
	}
}

// exp(-2*pi*i*k/32)
__device__ __forceinline__ float2 w32(int k)
{
	const float c[9] = { 1.0f, 0.98078528040323044913f, 0.92387953251128673848f, 0.83146961230254523708f,
		0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508978178f, 0.19509032201612826785f, 0.0f };
	k &= 31;
	// cos(pi k/16), -sin(pi k/16) folded into the first quadrant
	const int q = k >> 3, r = k & 7;
	float cs, sn;
	if (q == 0) { cs = c[r]; sn = c[8 - r]; }
	else if (q == 1) { cs = -c[8 - r]; sn = c[r]; }
	else if (q == 2) { cs = -c[r]; sn = -c[8 - r]; }
	else { cs = c[8 - r]; sn = -c[r]; }
	return make_float2(cs, -sn);
}

// In-register forward DFT of R points, natural order in and out.  v[j] -> V[k].
template <int R> struct RegDft;

template <> struct RegDft<2> {
	__device__ __forceinline__ static void run(float2 (&v)[2]) { dft2(v[0], v[1]); }
};

template <> struct RegDft<4> {
	__device__ __forceinline__ static void run(float2 (&v)[4]) { dft4(v[0], v[1], v[2], v[3]); }
};

template <> struct RegDft<8> {
	__device__ __forceinline__ static void run(float2 (&v)[8])
	{
		// 8 = 2 x 4: radix-2 over j = j0 + 4*j1 (j1), twiddle W_8^(j0*k1), radix-4 over j0
		#pragma unroll
		for (int j0 = 0; j0 < 4; j0++)
			dft2(v[j0], v[j0 + 4]);
		v[5] = cmul(v[5], w16(2));
		v[6] = mul_mi(v[6]);
		v[7] = cmul(v[7], w16(6));
		dft4(v[0], v[1], v[2], v[3]);   // k1 = 0: outputs k = 2*k0
		dft4(v[4], v[5], v[6], v[7]);   // k1 = 1: outputs k = 2*k0 + 1
		// natural order: V[2*k0 + k1] = v[4*k1 + k0]
		const float2 t1 = v[1], t2 = v[2], t3 = v[3], t4 = v[4], t5 = v[5], t6 = v[6];
		v[1] = t4; v[2] = t1; v[3] = t5; v[4] = t2; v[5] = t6; v[6] = t3;
	}
};

template <> struct RegDft<16> {
	__device__ __forceinline__ static void run(float2 (&v)[16])
	{
		// 16 = 4 x 4: j = j0 + 4*j1; radix-4 over j1, twiddle W_16^(j0*k1), radix-4 over j0
		#pragma unroll
		for (int j0 = 0; j0 < 4; j0++)
			dft4(v[j0], v[j0 + 4], v[j0 + 8], v[j0 + 12]);      // v[j0 + 4*k1]
		#pragma unroll
		for (int k1 = 1; k1 < 4; k1++)
			#pragma unroll
			for (int j0 = 1; j0 < 4; j0++)
				v[j0 + 4 * k1] = cmul(v[j0 + 4 * k1], w16(j0 * k1));
		#pragma unroll
		for (int k1 = 0; k1 < 4; k1++)
			dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]); // v[4*k1 + k0], k = k1 + 4*k0
		// natural order: V[k1 + 4*k0] = v[4*k1 + k0]  (a 4x4 transpose)
		#pragma unroll
		for (int a = 0; a < 4; a++)
			#pragma unroll
			for (int b = a + 1; b < 4; b++) {
				const float2 t = v[4 * a + b];
				v[4 * a + b] = v[4 * b + a];
				v[4 * b + a] = t;
			}
	}
};

template <> struct RegDft<32> {
	__device__ __forceinline__ static void run(float2 (&v)[32])
	{
		// 32 = 2 x 16: j = j0 + 16*j1; radix-2 over j1, twiddle W_32^(j0*k1), radix-16 over j0
		#pragma unroll
		for (int j0 = 0; j0 < 16; j0++)
			dft2(v[j0], v[j0 + 16]);
		#pragma unroll
		for (int j0 = 1; j0 < 16; j0++)
			v[j0 + 16] = cmul(v[j0 + 16], w32(j0));
		float2 lo[16], hi[16];
		#pragma unroll
		for (int j0 = 0; j0 < 16; j0++) { lo[j0] = v[j0]; hi[j0] = v[j0 + 16]; }
		RegDft<16>::run(lo);   // k1 = 0: outputs k = 2*k0
		RegDft<16>::run(hi);   // k1 = 1: outputs k = 2*k0 + 1
		#pragma unroll
		for (int k0 = 0; k0 < 16; k0++) { v[2 * k0] = lo[k0]; v[2 * k0 + 1] = hi[k0]; }
	}
};

} // namespace wrfft
