/*
 * oracle/shim/fftw3.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Stand-in for FFTW3 single precision, which the reference links
 * (reference configure.ac:32, src/Makefile.am:36) but which is not installed
 * in this image.  It supplies exactly the symbols the reference's hot path
 * uses (reference src/dsp/lowpass.cxx:98-100,120-123,180 and
 * src/io/spectrumsink.cxx:65-68,81-84,115) so that the reference sources can
 * be compiled UNMODIFIED, in place, into oracle/_ref/.
 *
 * Arithmetic: every transform is evaluated in float64 (direct DFT with exactly
 * reduced twiddle indices for n <= 256, iterative radix-2 FFT above that) and
 * rounded to float once.  This is at least as accurate as FFTW3f itself, so
 * (i) filter-design coefficients are the correctly rounded version of what the
 * reference intends and (ii) spectrum parity is judged at the north_star
 * tolerance (1e-5 relative), not bit-wise.  PARITY UNPINNED at the FFTW
 * boundary: the reference holds no golden vectors for it.
 */
#ifndef WR_ORACLE_FFTW3_SHIM_H
#define WR_ORACLE_FFTW3_SHIM_H

#include <stdlib.h>
#include <math.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef float fftwf_complex[2];

#define FFTW_FORWARD  (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE  (0U)
#define FFTW_ESTIMATE (1U << 6)

struct wr_shim_plan {
	int n;
	int sign;
	fftwf_complex *in;
	fftwf_complex *out;
	double *work; /* 2*n doubles */
};
typedef struct wr_shim_plan *fftwf_plan;

static inline void *fftwf_malloc(size_t bytes)
{
	void *p = NULL;
	if (posix_memalign(&p, 64, bytes ? bytes : 64) != 0)
		return NULL;
	return p;
}

static inline void fftwf_free(void *p) { free(p); }

static inline fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out,
		int sign, unsigned flags)
{
	(void)flags;
	struct wr_shim_plan *p = (struct wr_shim_plan*)malloc(sizeof(*p));
	if (!p)
		return NULL;
	p->n = n;
	p->sign = sign;
	p->in = in;
	p->out = out;
	p->work = (double*)malloc(sizeof(double) * 2 * (size_t)(n > 0 ? n : 1));
	return p;
}

static inline void fftwf_destroy_plan(fftwf_plan p)
{
	if (p) {
		free(p->work);
		free(p);
	}
}

static inline void fftwf_cleanup(void) {}

/* Unnormalised DFT, out[k] = sum_n in[n] * exp(sign * 2*pi*i * k*n / N). */
static inline void fftwf_execute(const fftwf_plan p)
{
	const int n = p->n;
	const double sgn = (p->sign < 0) ? -1.0 : 1.0;
	double *w = p->work;
	const double two_pi = 6.283185307179586476925286766559;

	if (n <= 256 || (n & (n - 1))) {
		/* direct evaluation; k*n reduced modulo N exactly in integers */
		for (int k = 0; k < n; k++) {
			double re = 0.0, im = 0.0;
			for (int m = 0; m < n; m++) {
				long long idx = ((long long)k * m) % n;
				double a = two_pi * (double)idx / (double)n;
				double c = cos(a), s = sgn * sin(a);
				double xr = p->in[m][0], xi = p->in[m][1];
				re += xr * c - xi * s;
				im += xr * s + xi * c;
			}
			w[2 * k] = re;
			w[2 * k + 1] = im;
		}
	} else {
		/* iterative radix-2 decimation-in-time in float64 */
		int bits = 0;
		while ((1 << bits) < n)
			bits++;
		for (int i = 0; i < n; i++) {
			unsigned r = 0;
			for (int b = 0; b < bits; b++)
				if (i & (1 << b))
					r |= 1u << (bits - 1 - b);
			w[2 * r] = p->in[i][0];
			w[2 * r + 1] = p->in[i][1];
		}
		for (int len = 2; len <= n; len <<= 1) {
			int half = len >> 1;
			for (int j = 0; j < half; j++) {
				double a = two_pi * (double)j / (double)len;
				double c = cos(a), s = sgn * sin(a);
				for (int base = 0; base < n; base += len) {
					double *u = &w[2 * (base + j)];
					double *v = &w[2 * (base + j + half)];
					double tr = v[0] * c - v[1] * s;
					double ti = v[0] * s + v[1] * c;
					v[0] = u[0] - tr;
					v[1] = u[1] - ti;
					u[0] += tr;
					u[1] += ti;
				}
			}
		}
	}
	for (int k = 0; k < n; k++) {
		p->out[k][0] = (float)w[2 * k];
		p->out[k][1] = (float)w[2 * k + 1];
	}
}

#ifdef __cplusplus
}
#endif

#endif /* WR_ORACLE_FFTW3_SHIM_H */
