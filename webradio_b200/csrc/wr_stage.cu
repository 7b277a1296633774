// wr_stage.cu -- strict single-stage blocks: one kernel per process() call on host buffers,
// i.e. exactly what one reference DspBlock does on its own.  The DspBlock drop-ins use these
// whenever a chain cannot be handed to a fused receiver bank (an unexpected consumer is
// attached between stages), and the stage-by-stage parity tests use them directly.
#include "wr_common.h"
#include "wr_device.cuh"

#include <algorithm>
#include <vector>

namespace {

// DownConverter::process (reference downconverter.cxx:91-114)
__global__ void stage_mix_kernel(const float2 *__restrict__ in, float2 *__restrict__ out,
		const float *__restrict__ table, uint32_t phase0, int32_t step, unsigned nframes)
{
	for (unsigned f = blockIdx.x * blockDim.x + threadIdx.x; f < nframes; f += gridDim.x * blockDim.x) {
		uint32_t si, ci;
		wrd::lo_indices(wrd::phase_at(phase0, step, f), si, ci);
		out[f] = wrd::mix(in[f], __ldg(table + ci), __ldg(table + si));
	}
}

// LowPass::process (reference lowpass.cxx:145-159) over blk = [history | input]
template <int kCh>
__global__ void stage_fir_kernel(const float *__restrict__ blk, const float *__restrict__ rtaps,
		float *__restrict__ out, unsigned nout, unsigned ntaps, unsigned decim)
{
	extern __shared__ float s_taps[];
	for (unsigned i = threadIdx.x; i < ntaps; i += blockDim.x)
		s_taps[i] = rtaps[i];
	__syncthreads();
	for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < nout; k += gridDim.x * blockDim.x) {
		const float *p = blk + (size_t)k * decim * kCh;
		if (kCh == 2) {
			float2 acc = make_float2(0.0f, 0.0f);
			const float2 *p2 = reinterpret_cast<const float2*>(p);
			for (unsigned j = 0; j < ntaps; j++)
				wrd::tap2(acc, s_taps[j], p2[j]);
			reinterpret_cast<float2*>(out)[k] = acc;
		} else {
			float acc = 0.0f;
			for (unsigned j = 0; j < ntaps; j++)
				wrd::tap1(acc, s_taps[j], p[j]);
			out[k] = acc;
		}
	}
}

// Demodulator::process (reference demodulator.cxx:77-115)
__global__ void stage_demod_kernel(const float2 *__restrict__ in, float *__restrict__ out,
		int mode, float2 prev0, unsigned nframes)
{
	for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < nframes; k += gridDim.x * blockDim.x)
		out[k] = wrd::demod(mode, in[k], k ? in[k - 1] : prev0);
}

__global__ void stage_palette_kernel(const float *__restrict__ db, unsigned char *__restrict__ out, unsigned n)
{
	for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
		out[k] = wrd::waterfall_index(db[k]);
}

// test hook: the device's atan2f (wr_atan2f.h) over arrays
__global__ void stage_atan2f_kernel(const float *__restrict__ y, const float *__restrict__ x, float *__restrict__ out, unsigned n)
{
	for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
		out[k] = wrd::atan2f_ref(y[k], x[k]);
}

} // namespace

struct wr_stage {
	int device = 0;
	cudaStream_t st = nullptr;
	float *d_in = nullptr, *d_out = nullptr;
	size_t capIn = 0, capOut = 0;
	float *d_table = nullptr;      // default table
	float *d_user_table = nullptr; // caller-supplied table (re-uploaded per call)
	// FIR
	unsigned ch = 0, ntaps = 0;
	float *d_taps = nullptr;
	float *d_blk[2] = { nullptr, nullptr };
	size_t blkCap = 0;
	int cur = 0;
	unsigned long long launches = 0;
};

namespace {

int grow(float **p, size_t *cap, size_t need)
{
	if (need <= *cap)
		return WR_OK;
	cudaFree(*p);
	*p = nullptr;
	*cap = 0;
	WR_CUDA(cudaMalloc(p, sizeof(float) * need));
	*cap = need;
	return WR_OK;
}

unsigned grid_for(unsigned n, unsigned threads)
{
	unsigned g = (n + threads - 1) / threads;
	return std::max(1u, std::min(g, 148u * 16u));
}

} // namespace

extern "C" {

wr_stage *wr_stage_create(int device)
{
	if (!wr::check_device(device))
		return nullptr;
	wr_stage *s = new wr_stage();
	s->device = device;
	if (cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking) != cudaSuccess) {
		wr::set_error("wr_stage_create: cudaStreamCreate failed");
		delete s;
		return nullptr;
	}
	return s;
}

void wr_stage_destroy(wr_stage *s)
{
	if (!s)
		return;
	cudaSetDevice(s->device);
	if (s->st)
		cudaStreamSynchronize(s->st);
	cudaFree(s->d_in);
	cudaFree(s->d_out);
	cudaFree(s->d_table);
	cudaFree(s->d_user_table);
	cudaFree(s->d_taps);
	cudaFree(s->d_blk[0]);
	cudaFree(s->d_blk[1]);
	if (s->st)
		cudaStreamDestroy(s->st);
	cudaGetLastError();
	delete s;
}

int wr_stage_mix(wr_stage *s, const float *table, uint32_t *phase, int32_t step,
		const float *iq_host, unsigned nframes, float *out_host)
{
	WR_REQUIRE(s && phase && (nframes == 0 || (iq_host && out_host)), WR_EINVAL, "wr_stage_mix: null argument");
	if (!wr::use_device(s->device))
		return WR_ENODEV;
	const float *d_tab;
	if (table) {
		if (!s->d_user_table)
			WR_CUDA(cudaMalloc(&s->d_user_table, sizeof(float) * WR_SINTABLE_SIZE));
		WR_CUDA(cudaMemcpyAsync(s->d_user_table, table, sizeof(float) * WR_SINTABLE_SIZE, cudaMemcpyHostToDevice, s->st));
		d_tab = s->d_user_table;
	} else {
		if (!s->d_table) {
			std::vector<float> t(WR_SINTABLE_SIZE);
			wr_build_sintable(t.data());
			WR_CUDA(cudaMalloc(&s->d_table, sizeof(float) * WR_SINTABLE_SIZE));
			WR_CUDA(cudaMemcpy(s->d_table, t.data(), sizeof(float) * WR_SINTABLE_SIZE, cudaMemcpyHostToDevice));
		}
		d_tab = s->d_table;
	}
	if (nframes) {
		int rc;
		if ((rc = grow(&s->d_in, &s->capIn, 2 * (size_t)nframes)) != WR_OK) return rc;
		if ((rc = grow(&s->d_out, &s->capOut, 2 * (size_t)nframes)) != WR_OK) return rc;
		WR_CUDA(cudaMemcpyAsync(s->d_in, iq_host, sizeof(float) * 2 * nframes, cudaMemcpyHostToDevice, s->st));
		stage_mix_kernel<<<grid_for(nframes, 256), 256, 0, s->st>>>(
				reinterpret_cast<const float2*>(s->d_in), reinterpret_cast<float2*>(s->d_out),
				d_tab, *phase, step, nframes);
		s->launches++;
		WR_CUDA(cudaGetLastError());
		WR_CUDA(cudaMemcpyAsync(out_host, s->d_out, sizeof(float) * 2 * nframes, cudaMemcpyDeviceToHost, s->st));
	}
	WR_CUDA(cudaStreamSynchronize(s->st));
	// closed form of nframes increments of (phase + step) & mask (downconverter.cxx:103)
	*phase = (*phase + nframes * (uint32_t)step) & 0x7FFFFFFFu;
	return WR_OK;
}

int wr_stage_fir_config(wr_stage *s, unsigned channels, const float *coeff, unsigned ntaps)
{
	WR_REQUIRE(s && coeff && (channels == 1 || channels == 2) && ntaps >= 1 && ntaps <= 8192, WR_EINVAL,
			"wr_stage_fir_config: bad argument (channels=%u ntaps=%u)", channels, ntaps);
	if (!wr::use_device(s->device))
		return WR_ENODEV;
	const bool geometryChanged = channels != s->ch || ntaps != s->ntaps;
	if (ntaps != s->ntaps) {
		cudaFree(s->d_taps);
		s->d_taps = nullptr;
		WR_CUDA(cudaMalloc(&s->d_taps, sizeof(float) * ntaps));
	}
	std::vector<float> rev(ntaps);
	for (unsigned j = 0; j < ntaps; j++)
		rev[j] = coeff[ntaps - 1 - j];
	WR_CUDA(cudaStreamSynchronize(s->st));
	WR_CUDA(cudaMemcpy(s->d_taps, rev.data(), sizeof(float) * ntaps, cudaMemcpyHostToDevice));
	s->ch = channels;
	s->ntaps = ntaps;
	if (geometryChanged)
		return wr_stage_fir_reset(s);
	return WR_OK;
}

int wr_stage_fir_reset(wr_stage *s)
{
	WR_REQUIRE(s, WR_EINVAL, "wr_stage_fir_reset: null stage");
	if (!wr::use_device(s->device))
		return WR_ENODEV;
	const size_t hist = (size_t)s->ch * (s->ntaps ? s->ntaps - 1 : 0);
	for (int i = 0; i < 2; i++)
		if (s->d_blk[i] && hist)
			WR_CUDA(cudaMemsetAsync(s->d_blk[i], 0, sizeof(float) * std::min(hist, s->blkCap), s->st));
	return WR_OK;
}

int wr_stage_fir(wr_stage *s, const float *in_host, unsigned nframes, unsigned decim, float *out_host)
{
	WR_REQUIRE(s && s->ntaps && decim >= 1, WR_EINVAL, "wr_stage_fir: not configured or bad decimation");
	WR_REQUIRE(nframes == 0 || (in_host && out_host), WR_EINVAL, "wr_stage_fir: null buffer");
	if (!wr::use_device(s->device))
		return WR_ENODEV;
	const unsigned ch = s->ch;
	const size_t hist = (size_t)ch * (s->ntaps - 1);
	const size_t need = hist + (size_t)ch * nframes;
	if (need > s->blkCap) {
		// grow both sides, preserving the current history
		float *nb[2] = { nullptr, nullptr };
		for (int i = 0; i < 2; i++) {
			WR_CUDA(cudaMalloc(&nb[i], sizeof(float) * need));
			WR_CUDA(cudaMemsetAsync(nb[i], 0, sizeof(float) * std::max<size_t>(hist, 1), s->st));
		}
		if (s->d_blk[s->cur] && hist)
			WR_CUDA(cudaMemcpyAsync(nb[s->cur], s->d_blk[s->cur], sizeof(float) * hist, cudaMemcpyDeviceToDevice, s->st));
		WR_CUDA(cudaStreamSynchronize(s->st));
		cudaFree(s->d_blk[0]);
		cudaFree(s->d_blk[1]);
		s->d_blk[0] = nb[0];
		s->d_blk[1] = nb[1];
		s->blkCap = need;
	}
	float *blk = s->d_blk[s->cur];
	float *nxt = s->d_blk[s->cur ^ 1];
	const unsigned nout = nframes / decim;
	int rc;
	if ((rc = grow(&s->d_out, &s->capOut, std::max<size_t>(1, (size_t)nout * ch))) != WR_OK) return rc;
	if (nframes)
		WR_CUDA(cudaMemcpyAsync(blk + hist, in_host, sizeof(float) * ch * nframes, cudaMemcpyHostToDevice, s->st));
	if (nout) {
		const unsigned threads = 128;
		// (whole 16-byte groups: the compiler fetches the last one to three taps with one 128-bit load)
		const size_t tapBytes = sizeof(float) * ((s->ntaps + 3u) & ~3u);
		if (ch == 2)
			stage_fir_kernel<2><<<grid_for(nout, threads), threads, tapBytes, s->st>>>(
					blk, s->d_taps, s->d_out, nout, s->ntaps, decim);
		else
			stage_fir_kernel<1><<<grid_for(nout, threads), threads, tapBytes, s->st>>>(
					blk, s->d_taps, s->d_out, nout, s->ntaps, decim);
		s->launches++;
		WR_CUDA(cudaGetLastError());
		WR_CUDA(cudaMemcpyAsync(out_host, s->d_out, sizeof(float) * ch * nout, cudaMemcpyDeviceToHost, s->st));
	}
	// lowpass.cxx:140-142: the last ntaps-1 frames of [history | input] seed the next call
	if (hist)
		WR_CUDA(cudaMemcpyAsync(nxt, blk + (size_t)ch * nframes, sizeof(float) * hist, cudaMemcpyDeviceToDevice, s->st));
	WR_CUDA(cudaStreamSynchronize(s->st));
	s->cur ^= 1;
	return WR_OK;
}

int wr_stage_demod(wr_stage *s, int mode, float *prev, const float *iq_host, unsigned nframes, float *out_host)
{
	WR_REQUIRE(s && prev, WR_EINVAL, "wr_stage_demod: null argument");
	WR_REQUIRE(mode >= WR_MODE_AM && mode <= WR_MODE_LSB, WR_EINVAL, "wr_stage_demod: bad mode %d", mode);
	if (nframes == 0)
		return WR_OK;
	WR_REQUIRE(iq_host && out_host, WR_EINVAL, "wr_stage_demod: null buffer");
	if (!wr::use_device(s->device))
		return WR_ENODEV;
	int rc;
	if ((rc = grow(&s->d_in, &s->capIn, 2 * (size_t)nframes)) != WR_OK) return rc;
	if ((rc = grow(&s->d_out, &s->capOut, 2 * (size_t)nframes)) != WR_OK) return rc;
	WR_CUDA(cudaMemcpyAsync(s->d_in, iq_host, sizeof(float) * 2 * nframes, cudaMemcpyHostToDevice, s->st));
	stage_demod_kernel<<<grid_for(nframes, 256), 256, 0, s->st>>>(
			reinterpret_cast<const float2*>(s->d_in), s->d_out, mode, make_float2(prev[0], prev[1]), nframes);
	s->launches++;
	WR_CUDA(cudaGetLastError());
	WR_CUDA(cudaMemcpyAsync(out_host, s->d_out, sizeof(float) * nframes, cudaMemcpyDeviceToHost, s->st));
	WR_CUDA(cudaStreamSynchronize(s->st));
	prev[0] = iq_host[2 * (size_t)(nframes - 1)];     // demodulator.cxx:110-111
	prev[1] = iq_host[2 * (size_t)(nframes - 1) + 1];
	return WR_OK;
}

int wr_stage_palette(wr_stage *s, const float *db_host, unsigned n, uint8_t *index_host)
{
	WR_REQUIRE(s, WR_EINVAL, "wr_stage_palette: null stage");
	if (n == 0)
		return WR_OK;
	WR_REQUIRE(db_host && index_host, WR_EINVAL, "wr_stage_palette: null buffer");
	if (!wr::use_device(s->device))
		return WR_ENODEV;
	int rc;
	if ((rc = grow(&s->d_in, &s->capIn, (size_t)n)) != WR_OK) return rc;
	if ((rc = grow(&s->d_out, &s->capOut, ((size_t)n + 3) / 4)) != WR_OK) return rc;
	WR_CUDA(cudaMemcpyAsync(s->d_in, db_host, sizeof(float) * n, cudaMemcpyHostToDevice, s->st));
	stage_palette_kernel<<<grid_for(n, 256), 256, 0, s->st>>>(s->d_in, reinterpret_cast<unsigned char*>(s->d_out), n);
	s->launches++;
	WR_CUDA(cudaGetLastError());
	WR_CUDA(cudaMemcpyAsync(index_host, s->d_out, n, cudaMemcpyDeviceToHost, s->st));
	WR_CUDA(cudaStreamSynchronize(s->st));
	return WR_OK;
}

int wr_stage_atan2f(wr_stage *s, const float *y_host, const float *x_host, unsigned n, float *out_host)
{
	WR_REQUIRE(s, WR_EINVAL, "wr_stage_atan2f: null stage");
	if (n == 0)
		return WR_OK;
	WR_REQUIRE(y_host && x_host && out_host, WR_EINVAL, "wr_stage_atan2f: null buffer");
	if (!wr::use_device(s->device))
		return WR_ENODEV;
	int rc;
	if ((rc = grow(&s->d_in, &s->capIn, 2 * (size_t)n)) != WR_OK) return rc;
	if ((rc = grow(&s->d_out, &s->capOut, (size_t)n)) != WR_OK) return rc;
	WR_CUDA(cudaMemcpyAsync(s->d_in, y_host, sizeof(float) * n, cudaMemcpyHostToDevice, s->st));
	WR_CUDA(cudaMemcpyAsync(s->d_in + n, x_host, sizeof(float) * n, cudaMemcpyHostToDevice, s->st));
	stage_atan2f_kernel<<<grid_for(n, 256), 256, 0, s->st>>>(s->d_in, s->d_in + n, s->d_out, n);
	s->launches++;
	WR_CUDA(cudaGetLastError());
	WR_CUDA(cudaMemcpyAsync(out_host, s->d_out, sizeof(float) * n, cudaMemcpyDeviceToHost, s->st));
	WR_CUDA(cudaStreamSynchronize(s->st));
	return WR_OK;
}

} // extern "C"
