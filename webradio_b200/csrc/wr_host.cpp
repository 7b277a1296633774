// wr_host.cpp -- cold-path host code of libwebradio_b200.so: error plumbing, the NCO table,
// the phase step and the filter design (K0 in SURVEY.md 2a stays on the host, as in the
// reference).  None of this runs per sample.
#include "wr_common.h"
#include "wr_lo.h"
#include "wr_lo3.h"
#include "wr_atan2f.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace wr {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}

const char *get_error() { return g_err; }

bool check_device(int device)
{
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n <= 0) {
		set_error("no CUDA device available (%s); libwebradio_b200 has no CPU fallback",
				e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
		cudaGetLastError();
		return false;
	}
	if (device < 0 || device >= n) {
		set_error("device %d out of range (have %d)", device, n);
		return false;
	}
	return use_device(device);
}

bool use_device(int device)
{
	int cur = -1;
	if (cudaGetDevice(&cur) == cudaSuccess && cur == device)
		return true;
	cudaError_t e = cudaSetDevice(device);
	if (e != cudaSuccess) {
		set_error("cudaSetDevice(%d): %s; libwebradio_b200 has no CPU fallback", device, cudaGetErrorString(e));
		cudaGetLastError();
		return false;
	}
	return true;
}

float lo_base_host(int s, const LoCoef &k)
{
	const float sf = (float)s;
	const float w = fmaxf(32768.0f - fabsf(sf), k.eps);
	const float u = sf * w;
	const float au = fabsf(u);
	float p = fmaf(au, k.a2, k.a1);
	p = fmaf(au, p, k.a0);
	return u * p;
}

bool lo_compress(const float *table, int16_t *delta, LoCoef *coef)
{
	LoCoef k = lo_coef();
	// index 32768 (s = -32768): 32768 - |s| = 0, so B = -32768 * eps * a0 (higher terms vanish)
	const float t = table[32768];
	k.eps = (t < 0.0f) ? (-t / (32768.0f * k.a0)) : 0.0f;
	if (!(k.eps < 1.0f))
		return false;
	for (uint32_t idx = 0; idx < WR_SINTABLE_SIZE; idx++) {
		const float b = lo_base_host((int)(int16_t)(uint16_t)idx, k);
		int32_t bb, tb;
		memcpy(&bb, &b, 4);
		memcpy(&tb, &table[idx], 4);
		const int64_t d = (int64_t)tb - (int64_t)bb;
		if (d < -32768 || d > 32767)
			return false;
		delta[idx] = (int16_t)d;
		int32_t back = bb + delta[idx];
		if (back != tb)
			return false;
	}
	if (coef)
		*coef = k;
	return true;
}


static inline float lo3_F(int s)
{
	// 2^23 + (s + 32768): what the device builds with one byte permute of the biased phase
	const uint32_t bits = 0x4B000000u | ((uint32_t)(s + 32768) & 0xFFFFu);
	float f;
	memcpy(&f, &bits, 4);
	return f;
}

static inline float lo3_poly(float y)
{
#if WR_LO3_DEGREE == 3
	float p = fmaf(y, WR_LO3_C3, WR_LO3_C2);
	p = fmaf(y, p, WR_LO3_C1);
#else
	const float p = fmaf(y, WR_LO3_C2, WR_LO3_C1);
#endif
	return fmaf(y, p, WR_LO3_C0);
}

float lo3_base_host(int s, const Lo3Coef &k)
{
	const float t = fmaf(lo3_F(s), WR_LO3_TSCALE, WR_LO3_TBIAS);
	const float y = t * t;
	const float z = fmaf(-y, t, t);                // t - t^3: zero at t = 0 and +-1 with a single rounding
	const float u = fmaf(t, k.eps, z);
	return u * lo3_poly(y);
}

int lo3_slot_host(int s)
{
	const float sl = fmaf(lo3_F(s), WR_LO3_SLOTK, WR_LO3_SLOTM);
	int32_t bits;
	memcpy(&bits, &sl, 4);
	return bits - WR_LO3_SLOTBITS;
}

bool lo3_compress(const float *table, int16_t *delta, Lo3Coef *coef)
{
	Lo3Coef k;
	// index 32768 (s = -32768, t = -1): t - t^3 = 0, so base = (-1 * eps) * poly(1)
	const float t = table[32768];
	k.eps = (t < 0.0f) ? (t / (-1.0f * lo3_poly(1.0f))) : 0.0f;
	if (!(k.eps < 1e-3f))
		return false;
	int prev = WR_LO3_SLOT_MIN - 1;
	for (int s = -32768; s < 32768; s++) {
		const int slot = lo3_slot_host(s);
		if (slot <= prev || slot > WR_LO3_SLOT_MAX)
			return false;
		prev = slot;
		const uint32_t idx = (uint32_t)s & 0xFFFFu;
		const float b = lo3_base_host(s, k);
		int32_t bb, tb;
		memcpy(&bb, &b, 4);
		memcpy(&tb, &table[idx], 4);
		const int64_t d = (int64_t)tb - (int64_t)bb;
		if (d < -32768 || d > 32767)
			return false;
		delta[idx] = (int16_t)d;
		if (bb + (int32_t)delta[idx] != tb)
			return false;
	}
	if (coef)
		*coef = k;
	return true;
}

} // namespace wr

extern "C" {

const char *wr_version(void) { return "webradio_b200 0.1 (sm_100a)"; }

const char *wr_last_error(void) { return wr::get_error(); }

void wr_atan2f_host(const float *y, const float *x, size_t n, float *out)
{
	for (size_t i = 0; i < n; i++)
		out[i] = wrat::atan2f_glibc(y[i], x[i]);
}

int wr_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

// Replaces the expression at reference downconverter.cxx:65 and :80: a 64-bit product of the
// IF and 2^31 divided (truncating toward zero) by the input sample rate.
int32_t wr_phase_step(int if_hz, unsigned sample_rate)
{
	if (sample_rate == 0)
		return 0;
	const int64_t full_turn = (int64_t)1 << 31;
	return (int32_t)((int64_t)if_hz * full_turn / (int64_t)sample_rate);
}

// Replaces the loop at reference downconverter.cxx:49-51.  The angle is formed in double
// ((float)n*2 is a float, the product with pi promotes) and narrowed to float by sinf; the
// table must come from the host libm so that it is bit-identical to the reference's on the
// same machine -- the GPU only ever gathers from it.
void wr_build_sintable(float *out)
{
	const float scale = (float)WR_SINTABLE_SIZE;
	for (unsigned n = 0; n < WR_SINTABLE_SIZE; n++) {
		double angle = (double)((float)n * 2) * M_PI / (double)scale;
		out[n] = sinf((float)angle);
	}
}

// Diagnostic for the shared-memory NCO table (wr_lo.h): compresses `table` (NULL = the default
// table), reconstructs every entry the way the v2 kernels do and compares bit for bit.
// Returns 0 if all 65536 entries are reproduced exactly, -1 if the table cannot be represented.
int wr_lo_compress_check(const float *table)
{
	std::vector<float> def;
	if (!table) {
		def.resize(WR_SINTABLE_SIZE);
		wr_build_sintable(def.data());
		table = def.data();
	}
	std::vector<int16_t> delta(WR_SINTABLE_SIZE);
	wr::LoCoef k;
	if (!wr::lo_compress(table, delta.data(), &k))
		return -1;
	for (uint32_t idx = 0; idx < WR_SINTABLE_SIZE; idx++) {
		const float b = wr::lo_base_host((int)(int16_t)(uint16_t)idx, k);
		int32_t bb;
		memcpy(&bb, &b, 4);
		bb += delta[idx];
		if (memcmp(&bb, &table[idx], 4) != 0)
			return -1;
	}
	return 0;
}

// Same diagnostic for the packed-arithmetic compression of the v3 kernels (wr_lo3.h).
int wr_lo3_compress_check(const float *table)
{
	std::vector<float> def;
	if (!table) {
		def.resize(WR_SINTABLE_SIZE);
		wr_build_sintable(def.data());
		table = def.data();
	}
	std::vector<int16_t> delta(WR_SINTABLE_SIZE);
	wr::Lo3Coef k;
	if (!wr::lo3_compress(table, delta.data(), &k))
		return -1;
	for (int s = -32768; s < 32768; s++) {
		const uint32_t idx = (uint32_t)s & 0xFFFFu;
		const float b = wr::lo3_base_host(s, k);
		int32_t bb;
		memcpy(&bb, &b, 4);
		bb += delta[idx];
		if (memcmp(&bb, &table[idx], 4) != 0)
			return -1;
	}
	return 0;
}

// Replaces LowPass::init's window (reference lowpass.cxx:102-110) and LowPass::recalculate
// (lowpass.cxx:164-189).  The reference runs an n-point inverse FFTW transform of a real,
// symmetric 0/1 mask; that transform is evaluated here directly in float64 (the mask is real
// and even, so only the cosine terms survive) and rounded once, which is the correctly
// rounded version of what the reference's float FFT approximates.
int wr_lowpass_design(unsigned n, unsigned passband_hz, unsigned sample_rate, float *coeff)
{
	WR_REQUIRE(n >= 2 && coeff && sample_rate > 0, WR_EINVAL, "wr_lowpass_design: bad arguments");
	std::vector<double> mask(n, 0.0);
	// same unsigned arithmetic as lowpass.cxx:167
	unsigned maxbin = n * passband_hz / sample_rate / 2;
	for (unsigned k = 0; k < n / 2 + 1; k++) {
		double v = (k < maxbin) ? 1.0 : 0.0;
		mask[k] = v;
		mask[(n - k) % n] = v;
	}
	const double two_pi = 6.283185307179586476925286766559;
	for (unsigned k = 0; k < n; k++) {
		unsigned bin = (k + n / 2) % n; // lowpass.cxx:184 re-ordering
		double re = 0.0;
		for (unsigned m = 0; m < n; m++) {
			unsigned long long idx = ((unsigned long long)bin * m) % n;
			double a = two_pi * (double)idx / (double)n;
			re += mask[m] * cos(a);
		}
		float impulse = (float)re;
		// lowpass.cxx:108-109: Hamming window, then the 1/N of the unnormalised transform
		float w = (float)(0.54 - 0.46 * cosf((float)(2 * M_PI * (float)k / (float)(n - 1))));
		w /= (float)n;
		coeff[k] = impulse * w;
	}
	return WR_OK;
}

} // extern "C"
