/* mock_capi.cxx -- TEST INFRASTRUCTURE ONLY: a CPU stand-in for the device entry points of
 * include/webradio_b200.h, so that the HOST logic of the drop-in blocks (webradio_b200/dsp,
 * webradio_b200/io: graph plumbing, lazy batching of receiver chains into one bank, setters that
 * reach the bank at block boundaries, the strict one-call-per-block fallback) can be exercised
 * without a GPU in the `-m "not gpu"` suite.
 *
 * It is linked ONLY into tests/harness/libwr_blocks_harness_mock.so (make harness-mock) together
 * with the C++ drop-in classes and the product's own host cold path (build/wr_host.o:
 * wr_phase_step, wr_lowpass_design, wr_last_error ...).  The arithmetic behind the entry points
 * below is the oracle's (oracle/wr_oracle.c, stage functions) -- so what a test on this library
 * checks is which calls the blocks make, with which arguments, in which order and on which
 * slices, NOT the CUDA kernels; those are checked by tests/test_blocks_gpu.py and
 * tests/test_parity_gpu.py on a B200.  Nothing under webradio_b200/ links or loads this file.
 *
 * Semantics mirrored from the header: settings take effect at the next block (trivially true: the
 * blocks call the setters between blocks, on the DSP thread); receiver r's audio starts at
 * r * audio_stride; carried state per receiver = NCO phase, both FIR histories, prev I/Q.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "webradio_b200.h"
#include "../../oracle/wr_oracle.h"

namespace {

const std::vector<float> &sintable()
{
	static std::vector<float> t;
	if (t.empty()) {
		t.resize(WRO_TABLE_SIZE);
		wr_build_sintable(t.data());   // the product's own host routine (wr_host.cpp)
	}
	return t;
}

struct Rx {
	uint32_t phase;
	int32_t step;
	int mode;
	float prev[2];
	wro_fir *fir1, *fir2;
	Rx() : phase(0), step(0), mode(WR_MODE_AM), fir1(NULL), fir2(NULL) { prev[0] = prev[1] = 0.0f; }
};

} // namespace

struct wr_bank {
	unsigned T, R, maxF, n1, d1, n2, d2;
	std::vector<Rx> rx;
	std::vector<unsigned> stream;
	std::vector<float> mixed, chan, demod;
	unsigned long long calls;
};

struct wr_stage {
	wro_fir *fir;
	unsigned channels, ntaps, decim;
	std::vector<float> taps;
};

struct wr_spectrum {
	wro_spectrum *s;
	unsigned n;
};

struct wr_upload {
	int device;
	size_t maxFrames;
	std::vector<float> data;   /* the "device copy" */
	unsigned nframes;
};

/* counters a test can read to see HOW the blocks drove the library */
static unsigned long long g_bank_process_calls, g_stage_calls, g_banks_created, g_uploads, g_upload_frames;

extern "C" {

unsigned long long wr_mock_bank_process_calls(void) { return g_bank_process_calls; }
unsigned long long wr_mock_stage_calls(void) { return g_stage_calls; }
unsigned long long wr_mock_banks_created(void) { return g_banks_created; }
unsigned long long wr_mock_uploads(void) { return g_uploads; }
unsigned long long wr_mock_upload_frames(void) { return g_upload_frames; }
void wr_mock_reset_counters(void) { g_bank_process_calls = g_stage_calls = g_banks_created = g_uploads = g_upload_frames = 0; }
/* how many devices the stand-in pretends to have (the blocks spread producers over them) */
extern "C" void wrhost_set_device_count_for_test(int n);
static int g_devices = 1;
void wr_mock_set_devices(int n) { g_devices = n > 0 ? n : 1; wrhost_set_device_count_for_test(n > 1 ? n : 0); }
static std::vector<int> g_bank_devices;
int wr_mock_bank_device(unsigned i) { return i < g_bank_devices.size() ? g_bank_devices[i] : -1; }
void wr_mock_clear_bank_devices(void) { g_bank_devices.clear(); }
/* failure injection: the next `n` bank / stage compute calls report a CUDA error (WR_ECUDA) */
static int g_fail_bank, g_fail_stage;
void wr_mock_fail_next(int bank_calls, int stage_calls) { g_fail_bank = bank_calls; g_fail_stage = stage_calls; }

wr_bank *wr_bank_create(int device, unsigned n_streams, unsigned n_receivers, unsigned max_frames,
		unsigned n1, unsigned d1, unsigned n2, unsigned d2)
{
	if (!n_streams || !n_receivers || !n1 || !d1 || !n2 || !d2 || device < 0 || device >= g_devices)
		return NULL;
	g_bank_devices.push_back(device);
	wr_bank *b = new wr_bank();
	b->T = n_streams; b->R = n_receivers; b->maxF = max_frames;
	b->n1 = n1; b->d1 = d1; b->n2 = n2; b->d2 = d2;
	b->rx.resize(n_receivers);
	b->stream.resize(n_receivers);
	std::vector<float> z1(n1, 0.0f), z2(n2, 0.0f);
	for (unsigned r = 0; r < n_receivers; r++) {
		b->stream[r] = r % n_streams;
		b->rx[r].fir1 = wro_fir_create(2, z1.data(), n1, d1);
		b->rx[r].fir2 = wro_fir_create(1, z2.data(), n2, d2);
	}
	b->calls = 0;
	g_banks_created++;
	return b;
}

void wr_bank_destroy(wr_bank *b)
{
	if (!b)
		return;
	for (size_t r = 0; r < b->rx.size(); r++) {
		wro_fir_destroy(b->rx[r].fir1);
		wro_fir_destroy(b->rx[r].fir2);
	}
	delete b;
}

int wr_rx_set_stream(wr_bank *b, unsigned rx, unsigned stream)
{
	if (!b || rx >= b->R || stream >= b->T)
		return WR_EINVAL;
	b->stream[rx] = stream;
	return WR_OK;
}

int wr_rx_set_phase_step(wr_bank *b, unsigned rx, int32_t step)
{
	if (!b || rx >= b->R)
		return WR_EINVAL;
	b->rx[rx].step = step;
	return WR_OK;
}

int wr_rx_set_taps(wr_bank *b, unsigned rx, int stage, const float *coeff, unsigned ntaps)
{
	if (!b || rx >= b->R || !coeff || ntaps != (stage ? b->n2 : b->n1))
		return WR_EINVAL;
	wro_fir_set_taps(stage ? b->rx[rx].fir2 : b->rx[rx].fir1, coeff, ntaps);
	return WR_OK;
}

int wr_rx_set_mode(wr_bank *b, unsigned rx, int mode)
{
	if (!b || rx >= b->R || mode < WR_MODE_AM || mode > WR_MODE_LSB)
		return WR_EINVAL;
	b->rx[rx].mode = mode;
	return WR_OK;
}

int wr_rx_set_phase(wr_bank *b, unsigned rx, uint32_t phase)
{
	if (!b || rx >= b->R)
		return WR_EINVAL;
	b->rx[rx].phase = phase & 0x7FFFFFFFu;
	return WR_OK;
}

int wr_rx_get_phase(wr_bank *b, unsigned rx, uint32_t *phase)
{
	if (!b || rx >= b->R || !phase)
		return WR_EINVAL;
	*phase = b->rx[rx].phase;
	return WR_OK;
}

int wr_rx_reset(wr_bank *b, unsigned rx, unsigned flags)
{
	if (!b || rx >= b->R)
		return WR_EINVAL;
	Rx &x = b->rx[rx];
	if (flags & WR_RESET_PHASE)
		x.phase = 0;
	if (flags & WR_RESET_DEMOD)
		x.prev[0] = x.prev[1] = 0.0f;
	if (flags & (WR_RESET_CHANNEL | WR_RESET_AUDIO))
		return WR_EINVAL;   /* not needed by the blocks; the oracle's FIR has no history reset */
	return WR_OK;
}

int wr_rx_set_lookback(wr_bank *b, unsigned rx, const float *prev_iq)
{
	if (!b || rx >= b->R || !prev_iq)
		return WR_EINVAL;
	b->rx[rx].prev[0] = prev_iq[0];
	b->rx[rx].prev[1] = prev_iq[1];
	return WR_OK;
}

int wr_rx_get_lookback(wr_bank *b, unsigned rx, float *prev_iq)
{
	if (!b || rx >= b->R || !prev_iq)
		return WR_EINVAL;
	prev_iq[0] = b->rx[rx].prev[0];
	prev_iq[1] = b->rx[rx].prev[1];
	return WR_OK;
}

int wr_bank_process(wr_bank *b, const float *iq_host, unsigned nframes, float *audio_host, size_t audio_stride)
{
	if (!b || !iq_host || !audio_host || nframes > b->maxF)
		return WR_EINVAL;
	g_bank_process_calls++;
	b->calls++;
	if (g_fail_bank > 0) {
		g_fail_bank--;
		return WR_ECUDA;
	}
	const unsigned m1 = nframes / b->d1;
	b->mixed.resize(2 * (size_t)nframes + 2);
	b->chan.resize(2 * (size_t)m1 + 2);
	b->demod.resize((size_t)m1 + 1);
	for (unsigned r = 0; r < b->R; r++) {
		Rx &x = b->rx[r];
		const float *iq = iq_host + 2 * (size_t)nframes * b->stream[r];
		wro_mix(sintable().data(), &x.phase, x.step, iq, nframes, b->mixed.data());
		wro_fir_process(x.fir1, b->mixed.data(), nframes, b->chan.data());
		wro_demod(x.mode, x.prev, b->chan.data(), m1, b->demod.data());
		wro_fir_process(x.fir2, b->demod.data(), m1, audio_host + (size_t)r * audio_stride);
	}
	return WR_OK;
}

int wr_rx_get_history(wr_bank *b, unsigned rx, int stage, float *out, unsigned nfloats)
{
	if (!b || rx >= b->R || !out || nfloats != (stage ? b->n2 - 1 : 2 * (b->n1 - 1)))
		return WR_EINVAL;
	wro_fir_get_history(stage ? b->rx[rx].fir2 : b->rx[rx].fir1, out);
	return WR_OK;
}

int wr_rx_set_history(wr_bank *b, unsigned rx, int stage, const float *in, unsigned nfloats)
{
	if (!b || rx >= b->R || !in || nfloats != (stage ? b->n2 - 1 : 2 * (b->n1 - 1)))
		return WR_EINVAL;
	wro_fir_set_history(stage ? b->rx[rx].fir2 : b->rx[rx].fir1, in);
	return WR_OK;
}

/* the shared upload: a host copy stands in for the device copy */
wr_upload *wr_upload_create(int device, size_t max_frames)
{
	if (!max_frames || device < 0 || device >= g_devices)
		return NULL;
	wr_upload *u = new wr_upload();
	u->device = device;
	u->maxFrames = max_frames;
	u->nframes = 0;
	return u;
}

void wr_upload_destroy(wr_upload *u) { delete u; }
size_t wr_upload_capacity(const wr_upload *u) { return u ? u->maxFrames : 0; }
int wr_upload_device(const wr_upload *u) { return u ? u->device : -1; }

int wr_upload_begin(wr_upload *u, const float *iq_host, unsigned nframes)
{
	if (!u || (!iq_host && nframes) || nframes > u->maxFrames)
		return WR_EINVAL;
	g_uploads++;
	g_upload_frames += nframes;
	u->data.assign(iq_host, iq_host + 2 * (size_t)nframes);
	u->nframes = nframes;
	return WR_OK;
}

int wr_upload_finish(wr_upload *u) { return u ? WR_OK : WR_EINVAL; }

int wr_bank_process_upload(wr_bank *b, wr_upload *u, unsigned nframes, float *audio_host, size_t audio_stride)
{
	if (!b || !u || b->T != 1 || nframes != u->nframes)
		return WR_EINVAL;
	return wr_bank_process(b, u->data.data(), nframes, audio_host, audio_stride);
}

void *wr_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void wr_host_free(void *p) { free(p); }

wr_stage *wr_stage_create(int)
{
	wr_stage *s = new wr_stage();
	s->fir = NULL;
	s->channels = s->ntaps = s->decim = 0;
	return s;
}

void wr_stage_destroy(wr_stage *s)
{
	if (!s)
		return;
	if (s->fir)
		wro_fir_destroy(s->fir);
	delete s;
}

int wr_stage_mix(wr_stage *s, const float *table_or_null, uint32_t *phase, int32_t step,
		const float *iq_host, unsigned nframes, float *out_host)
{
	if (!s || !phase || (nframes && (!iq_host || !out_host)))
		return WR_EINVAL;
	g_stage_calls++;
	if (g_fail_stage > 0) {
		g_fail_stage--;
		return WR_ECUDA;
	}
	wro_mix(table_or_null ? table_or_null : sintable().data(), phase, step, iq_host, nframes, out_host);
	return WR_OK;
}

int wr_stage_fir_config(wr_stage *s, unsigned channels, const float *coeff, unsigned ntaps)
{
	if (!s || !coeff || !ntaps || (channels != 1 && channels != 2))
		return WR_EINVAL;
	s->taps.assign(coeff, coeff + ntaps);
	if (s->fir && (channels != s->channels || ntaps != s->ntaps)) {
		wro_fir_destroy(s->fir);   // a different filter shape starts from an empty history
		s->fir = NULL;
	}
	s->channels = channels;
	s->ntaps = ntaps;
	if (s->fir)
		wro_fir_set_taps(s->fir, coeff, ntaps);   // new coefficients, history kept (LowPass::setPassband)
	return WR_OK;
}

int wr_stage_fir(wr_stage *s, const float *in_host, unsigned nframes, unsigned decim, float *out_host)
{
	if (!s || !s->ntaps || !decim || (nframes && (!in_host || !out_host)))
		return WR_EINVAL;
	g_stage_calls++;
	if (s->fir && decim != s->decim) {
		wro_fir_destroy(s->fir);
		s->fir = NULL;
	}
	if (!s->fir) {
		s->fir = wro_fir_create(s->channels, s->taps.data(), s->ntaps, decim);
		s->decim = decim;
	}
	wro_fir_process(s->fir, in_host, nframes, out_host);
	return WR_OK;
}

int wr_stage_fir_reset(wr_stage *s)
{
	if (!s)
		return WR_EINVAL;
	if (s->fir) {
		wro_fir_destroy(s->fir);
		s->fir = NULL;
	}
	return WR_OK;
}

int wr_stage_demod(wr_stage *s, int mode, float *prev, const float *iq_host, unsigned nframes, float *out_host)
{
	if (!s || !prev)
		return WR_EINVAL;
	g_stage_calls++;
	if (nframes == 0)
		return WR_OK;
	return wro_demod(mode, prev, iq_host, nframes, out_host) == 0 ? WR_OK : WR_EINVAL;
}

wr_spectrum *wr_spectrum_create(int, unsigned fft_size, unsigned hop, unsigned n_streams, unsigned)
{
	if (n_streams != 1)
		return NULL;   // the SpectrumSink block has one input stream
	wro_spectrum *o = wro_spectrum_create(fft_size, hop);
	if (!o)
		return NULL;
	wr_spectrum *s = new wr_spectrum();
	s->s = o;
	s->n = fft_size;
	return s;
}

void wr_spectrum_destroy(wr_spectrum *s)
{
	if (!s)
		return;
	wro_spectrum_destroy(s->s);
	delete s;
}

long wr_spectrum_process(wr_spectrum *s, const float *iq_host, unsigned nframes, float *rows_host, size_t)
{
	if (!s || (nframes && !iq_host))
		return WR_EINVAL;
	return (long)wro_spectrum_process(s->s, iq_host, nframes, rows_host, rows_host ? (size_t)nframes / s->n + 2 : 0);
}

long wr_spectrum_process_upload(wr_spectrum *s, wr_upload *u, unsigned nframes)
{
	if (!s || !u || nframes != u->nframes)
		return WR_EINVAL;
	return (long)wro_spectrum_process(s->s, u->data.data(), nframes, NULL, 0);
}

int wr_spectrum_reserve(wr_spectrum *s, unsigned max_frames) { return (s && max_frames) ? WR_OK : WR_EINVAL; }

int wr_spectrum_get(wr_spectrum *s, unsigned stream, float *db_host)
{
	if (!s || stream != 0 || !db_host)
		return WR_EINVAL;
	wro_spectrum_get(s->s, db_host);
	return WR_OK;
}

} /* extern "C" */
