/*
 * oracle/wr_oracle.c -- TEST INFRASTRUCTURE ONLY (see wr_oracle.h).
 *
 * CPU restatement, in plain C, of the arithmetic of the reference's hot path.
 * Every function cites the reference lines it follows.  Compile with
 * -O2 -ffp-contract=off and no -march (oracle/Makefile) so that every float
 * multiply and add is rounded separately, as in the reference's stock build.
 *
 * Pinning: tests/test_oracle_vs_ref.py checks this file bit-for-bit against
 * oracle/_ref/libwr_ref.so (the unmodified reference sources) and
 * tests/test_oracle_golden.py against tests/golden/ (vectors generated from
 * the reference by scripts/make_golden.py).  The FFT itself (FFTW3f in the
 * reference) is "parity unpinned": evaluated in float64 via oracle/shim.
 */
#define _GNU_SOURCE
#include "wr_oracle.h"
#include "shim/fftw3.h"

#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include <pthread.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define PHASE_BITS 31
#define LOOKUP_BITS 16
#define PHASE_MASK 0x7FFFFFFFu
#define LOOKUP_MASK 0xFFFFu
#define LOOKUP_SHIFT (PHASE_BITS - LOOKUP_BITS)

/* reference downconverter.cxx:49-51: the argument is evaluated in double
 * ((float)n * 2 is float, * M_PI promotes) and narrowed to float by sinf. */
void wro_sintable(float *out)
{
	for (unsigned n = 0; n < WRO_TABLE_SIZE; n++)
		out[n] = sinf((float)n * 2 * M_PI / (float)(1UL << LOOKUP_BITS));
}

/* reference downconverter.cxx:80 (and :65): 64-bit product, truncating division */
int32_t wro_phase_step(int if_hz, unsigned fs)
{
	return (int)((int64_t)if_hz * (int64_t)(1UL << PHASE_BITS) / (int64_t)fs);
}

/* reference lowpass.cxx:102-110 (window) and :164-189 (recalculate) */
void wro_lowpass_design(unsigned n, unsigned passband_hz, unsigned fs, float *coeff)
{
	fftwf_complex *spec = (fftwf_complex*)fftwf_malloc(sizeof(fftwf_complex) * n);
	fftwf_complex *impulse = (fftwf_complex*)fftwf_malloc(sizeof(fftwf_complex) * n);
	fftwf_plan p = fftwf_plan_dft_1d((int)n, spec, impulse, FFTW_BACKWARD, FFTW_ESTIMATE);
	float *window = (float*)malloc(sizeof(float) * n);

	for (unsigned k = 0; k < n; k++) {
		window[k] = 0.54 - 0.46 * cosf(2 * M_PI * (float)k / (float)(n - 1));
		window[k] /= (float)n;
	}

	/* lowpass.cxx:167: all-unsigned arithmetic */
	unsigned maxbin = n * passband_hz / fs / 2;
	for (unsigned k = 0; k < n / 2 + 1; k++) {
		unsigned mirror = (n - k) % n; /* reference: & (n-1), n a power of two */
		spec[k][0] = spec[mirror][0] = (k < maxbin) ? 1.0 : 0.0;
		spec[k][1] = spec[mirror][1] = 0.0;
	}
	fftwf_execute(p);
	for (unsigned k = 0; k < n; k++) {
		unsigned bin = (k + n / 2) % n;
		coeff[k] = impulse[bin][0] * window[k];
	}

	free(window);
	fftwf_destroy_plan(p);
	fftwf_free(spec);
	fftwf_free(impulse);
}

/* reference spectrumsink.cxx:71-74 */
void wro_spectrum_window(unsigned n, float *window)
{
	for (unsigned k = 0; k < n; k++)
		window[k] = 0.54 - 0.46 * cosf(2 * M_PI * (float)k / (float)(n - 1));
}

/* reference downconverter.cxx:91-114 */
void wro_mix(const float *table, uint32_t *phase_io, int32_t step,
		const float *in, size_t nframes, float *out)
{
	uint32_t phase = *phase_io;
	while (nframes--) {
		uint32_t sinidx = phase >> LOOKUP_SHIFT;
		uint32_t cosidx = (sinidx + (1u << LOOKUP_BITS) / 4) & LOOKUP_MASK;
		phase = (phase + (uint32_t)step) & PHASE_MASK;
		float i = *in++;
		float q = *in++;
		*out++ = i * table[cosidx] + q * table[sinidx];
		*out++ = q * table[cosidx] - i * table[sinidx];
	}
	*phase_io = phase;
}

/* ---- LowPass::process (reference lowpass.cxx:131-162) ---- */
struct wro_fir {
	unsigned channels, ntaps, decim;
	float *coeff;
	float *block;     /* [history | current input], as the reference's `block` vector */
	size_t block_len; /* floats */
	float *preset;    /* history handed over by wro_fir_set_history, consumed by the next call (test plumbing) */
};

wro_fir *wro_fir_create(unsigned channels, const float *coeff, unsigned ntaps, unsigned decim)
{
	wro_fir *f = (wro_fir*)calloc(1, sizeof(*f));
	f->channels = channels;
	f->decim = decim;
	wro_fir_set_taps(f, coeff, ntaps);
	return f;
}

void wro_fir_set_taps(wro_fir *f, const float *coeff, unsigned ntaps)
{
	if (ntaps != f->ntaps) {
		/* a new length restarts the history from zeros (the reference cannot
		 * change length at run time: lowpass.cxx:39) */
		free(f->block);
		f->block = NULL;
		f->block_len = 0;
	}
	free(f->coeff);
	f->coeff = (float*)malloc(sizeof(float) * ntaps);
	memcpy(f->coeff, coeff, sizeof(float) * ntaps);
	f->ntaps = ntaps;
}

size_t wro_fir_process(wro_fir *f, const float *in, size_t nframes, float *out)
{
	const unsigned ch = f->channels;
	const size_t hist = (size_t)ch * (f->ntaps - 1);
	const size_t in_len = nframes * ch;
	const size_t want = in_len + hist;

	/* lowpass.cxx:138-139: vector::resize (grow = zero fill, shrink = truncate) */
	if (f->block_len != want) {
		f->block = (float*)realloc(f->block, sizeof(float) * (want ? want : 1));
		if (want > f->block_len)
			memset(f->block + f->block_len, 0, sizeof(float) * (want - f->block_len));
		f->block_len = want;
	}
	/* lowpass.cxx:140-142: keep the last ntaps-1 frames, append the new block */
	if (f->preset) {
		memcpy(f->block, f->preset, sizeof(float) * hist);
		free(f->preset);
		f->preset = NULL;
	} else {
		memmove(f->block, f->block + f->block_len - hist, sizeof(float) * hist);
	}
	memcpy(f->block + hist, in, sizeof(float) * in_len);

	/* lowpass.cxx:145-159: nout = floor(nframes / decim); taps walked last-to-first
	 * against samples walked oldest-to-newest; mul and add rounded separately */
	const size_t nout = nframes / f->decim;
	const size_t instep = (size_t)ch * f->decim;
	const float *src = f->block;
	for (size_t k = 0; k < nout; k++) {
		for (unsigned c = 0; c < ch; c++)
			out[c] = 0.0f;
		const float *p = src;
		for (unsigned j = 0; j < f->ntaps; j++) {
			const float cf = f->coeff[f->ntaps - 1 - j];
			for (unsigned c = 0; c < ch; c++)
				out[c] += cf * (*p++);
		}
		out += ch;
		src += instep;
	}
	return nout;
}

/* Test plumbing (no reference counterpart): the history a filter carries -- the last ntaps-1 frames
 * of `block` (lowpass.cxx:140-142) -- read out, or handed to a filter that has not run yet, so that
 * the CPU stand-in of the C ABI can move a receiver between banks as the CUDA library does. */
size_t wro_fir_get_history(const wro_fir *f, float *out)
{
	const size_t hist = (size_t)f->channels * (f->ntaps - 1);
	if (f->preset)
		memcpy(out, f->preset, sizeof(float) * hist);
	else if (f->block && f->block_len >= hist)
		memcpy(out, f->block + f->block_len - hist, sizeof(float) * hist);
	else
		memset(out, 0, sizeof(float) * hist);
	return hist;
}

void wro_fir_set_history(wro_fir *f, const float *in)
{
	const size_t hist = (size_t)f->channels * (f->ntaps - 1);
	free(f->preset);
	f->preset = (float*)malloc(sizeof(float) * (hist ? hist : 1));
	memcpy(f->preset, in, sizeof(float) * hist);
}

void wro_fir_destroy(wro_fir *f)
{
	if (!f)
		return;
	free(f->preset);
	free(f->coeff);
	free(f->block);
	free(f);
}

/* What a waterfall bin becomes on the way to the screen.  WaterfallHandler::doGet (reference
 * src/web/waterfallhandler.cxx:62-68) sends finite values as they are and anything else as
 * -10000.0; Waterfall.update (reference html/waterfall.js:92-109) computes, in JavaScript numbers
 * (doubles): val = (series[bin] + 50.0) / 25.0; val = val * 255.0; floor; clamp to [0, 255]. */
void wro_waterfall_index(const float *db, size_t n, unsigned char *out)
{
	for (size_t i = 0; i < n; i++) {
		double v = isfinite(db[i]) ? (double)db[i] : -10000.0;
		double val = (v + 50.0) / 25.0;
		val = val * 255.0;
		val = floor(val);
		if (val < 0) val = 0;
		if (val > 255) val = 255;
		out[i] = (unsigned char)val;
	}
}

/* reference src/web/mp3encoder.cxx:66-73: left[n] = (*ptr++) * 32768.0 (double product, stored to float) */
void wro_lame_scale(const float *x, size_t n, float *out)
{
	for (size_t i = 0; i < n; i++)
		out[i] = x[i] * 32768.0;
}

/* reference src/io/rtlsdrtuner.cxx:106: buffer->push_back(((float)(*buf++) - 128.0) / 128.0) */
void wro_rtlsdr_convert(const unsigned char *buf, size_t n, float *out)
{
	for (size_t i = 0; i < n; i++)
		out[i] = ((float)buf[i] - 128.0) / 128.0;
}

/* the libm routine reference demodulator.cxx:97 calls, exposed so that tests can pin the
 * product's restatement of it against the C library installed on the box */
void wro_libm_atan2f(const float *y, const float *x, size_t n, float *out)
{
	for (size_t i = 0; i < n; i++)
		out[i] = atan2f(y[i], x[i]);
}

/* reference demodulator.cxx:77-115 */
int wro_demod(int mode, float *prev, const float *in, size_t nframes, float *out)
{
	float prev_i = prev[0], prev_q = prev[1];
	while (nframes--) {
		float i = *in++;
		float q = *in++;
		switch (mode) {
		case WRO_AM:
			*out++ = sqrtf(i * i + q * q);
			break;
		case WRO_FM: {
			float ii = i * prev_i + q * prev_q;
			float qq = q * prev_i - i * prev_q;
			/* note the argument order (ii, qq) and the double divide chain */
			*out++ = atan2f(ii, qq) / M_PI / 2.0;
			break;
		}
		case WRO_USB:
			*out++ = i + q;
			break;
		case WRO_LSB:
			*out++ = i - q;
			break;
		default:
			return -1;
		}
		prev_i = i;
		prev_q = q;
	}
	prev[0] = prev_i;
	prev[1] = prev_q;
	return 0;
}

/* ---- whole receiver: reference radio.cxx:62-90 ---- */
struct wro_rx {
	unsigned fs;
	float *table;
	uint32_t phase; /* downconverter.h:58, initialised once (downconverter.cxx:46) */
	int32_t step;
	wro_fir *chan, *audio;
	int mode;
	float prev[2];  /* demodulator.h:60-61 */
	float *t_mixed, *t_chan, *t_demod;
	size_t cap;
};

wro_rx *wro_rx_create(unsigned fs, int if_hz,
		const float *taps1, unsigned n1, unsigned d1, int mode,
		const float *taps2, unsigned n2, unsigned d2)
{
	wro_rx *r = (wro_rx*)calloc(1, sizeof(*r));
	r->fs = fs;
	r->table = (float*)malloc(sizeof(float) * WRO_TABLE_SIZE);
	wro_sintable(r->table);
	r->phase = 0;
	r->step = wro_phase_step(if_hz, fs);
	r->chan = wro_fir_create(2, taps1, n1, d1);
	r->audio = wro_fir_create(1, taps2, n2, d2);
	r->mode = mode;
	return r;
}

void wro_rx_set_if(wro_rx *r, int if_hz) { r->step = wro_phase_step(if_hz, r->fs); }
void wro_rx_set_mode(wro_rx *r, int mode) { r->mode = mode; }
void wro_rx_set_taps(wro_rx *r, int which, const float *taps, unsigned n)
{
	wro_fir_set_taps(which ? r->audio : r->chan, taps, n);
}

size_t wro_rx_process(wro_rx *r, const float *iq, size_t nframes,
		float *mixed, float *chan, float *demod, float *audio)
{
	if (r->cap < nframes) {
		r->t_mixed = (float*)realloc(r->t_mixed, sizeof(float) * 2 * nframes);
		r->t_chan = (float*)realloc(r->t_chan, sizeof(float) * 2 * nframes);
		r->t_demod = (float*)realloc(r->t_demod, sizeof(float) * nframes);
		r->cap = nframes;
	}
	float *m = mixed ? mixed : r->t_mixed;
	float *c = chan ? chan : r->t_chan;
	float *d = demod ? demod : r->t_demod;
	wro_mix(r->table, &r->phase, r->step, iq, nframes, m);
	size_t n1 = wro_fir_process(r->chan, m, nframes, c);
	if (wro_demod(r->mode, r->prev, c, n1, d) != 0)
		return 0;
	return wro_fir_process(r->audio, d, n1, audio);
}

void wro_rx_destroy(wro_rx *r)
{
	if (!r)
		return;
	wro_fir_destroy(r->chan);
	wro_fir_destroy(r->audio);
	free(r->table);
	free(r->t_mixed);
	free(r->t_chan);
	free(r->t_demod);
	free(r);
}

/* ---- SpectrumSink: reference spectrumsink.cxx:60-142 ---- */
struct wro_spectrum {
	unsigned n, hop;
	float *window;
	float *raw;            /* last n frames, unwindowed (needed for hop < n) */
	fftwf_complex *inbuf, *outbuf;
	fftwf_plan p;
	unsigned inoffset;
};

wro_spectrum *wro_spectrum_create(unsigned n, unsigned hop)
{
	if (n == 0 || (n & (n - 1)) || hop == 0 || hop > n)
		return NULL; /* spectrumsink.cxx:53-56: power of two only */
	wro_spectrum *s = (wro_spectrum*)calloc(1, sizeof(*s));
	s->n = n;
	s->hop = hop;
	s->window = (float*)malloc(sizeof(float) * n);
	wro_spectrum_window(n, s->window);
	s->raw = (float*)calloc(2 * (size_t)n, sizeof(float));
	s->inbuf = (fftwf_complex*)fftwf_malloc(sizeof(fftwf_complex) * n);
	s->outbuf = (fftwf_complex*)fftwf_malloc(sizeof(fftwf_complex) * n);
	memset(s->outbuf, 0, sizeof(fftwf_complex) * n);
	s->p = fftwf_plan_dft_1d((int)n, s->inbuf, s->outbuf, FFTW_FORWARD, FFTW_ESTIMATE);
	return s;
}

static void spectrum_db(const wro_spectrum *s, float *db)
{
	/* spectrumsink.cxx:127-140 */
	const unsigned n = s->n;
	float scaledb = 20 * log10f((float)n);
	for (unsigned k = 0; k < n; k++) {
		float v = 10 * log10f(s->outbuf[k][0] * s->outbuf[k][0] + s->outbuf[k][1] * s->outbuf[k][1]);
		db[(k < n / 2) ? (k + n / 2) : (k - n / 2)] = v - scaledb;
	}
}

size_t wro_spectrum_process(wro_spectrum *s, const float *in, size_t nframes,
		float *rows, size_t max_rows)
{
	size_t done = 0;
	const unsigned n = s->n;
	/* spectrumsink.cxx:101-121, with the frame advancing by hop instead of n */
	while (nframes) {
		size_t blocksize = n - s->inoffset;
		if (blocksize > nframes)
			blocksize = nframes;
		memcpy(s->raw + 2 * (size_t)s->inoffset, in, blocksize * 2 * sizeof(float));
		s->inoffset += (unsigned)blocksize;
		if (s->inoffset == n) {
			for (unsigned k = 0; k < n; k++) {
				/* spectrumsink.cxx:110-113: in-place float multiply */
				s->inbuf[k][0] = s->raw[2 * k] * s->window[k];
				s->inbuf[k][1] = s->raw[2 * k + 1] * s->window[k];
			}
			fftwf_execute(s->p);
			if (rows && done < max_rows)
				spectrum_db(s, rows + done * n);
			done++;
			memmove(s->raw, s->raw + 2 * (size_t)s->hop, sizeof(float) * 2 * (size_t)(n - s->hop));
			s->inoffset = n - s->hop;
		}
		nframes -= blocksize;
		in += blocksize * 2;
	}
	return done;
}

void wro_spectrum_get(const wro_spectrum *s, float *db) { spectrum_db(s, db); }

void wro_spectrum_get_bins(const wro_spectrum *s, float *bins)
{
	memcpy(bins, s->outbuf, sizeof(fftwf_complex) * s->n);
}

void wro_spectrum_destroy(wro_spectrum *s)
{
	if (!s)
		return;
	fftwf_destroy_plan(s->p);
	fftwf_free(s->inbuf);
	fftwf_free(s->outbuf);
	free(s->raw);
	free(s->window);
	free(s);
}

/* ---- timed CPU baseline ---- */
struct bench_job {
	unsigned fs; size_t nframes; unsigned n_streams;
	unsigned rx_lo, rx_hi;
	const int *if_hz, *modes;
	const float *taps1; unsigned n1, d1;
	const float *taps2; unsigned n2, d2;
	const float *iq;
	unsigned warmup, blocks;
	pthread_barrier_t *bar;
	double seconds;
	float checksum;
};

static double now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void *bench_worker(void *arg)
{
	struct bench_job *j = (struct bench_job*)arg;
	unsigned nrx = j->rx_hi - j->rx_lo;
	wro_rx **rx = (wro_rx**)calloc(nrx ? nrx : 1, sizeof(*rx));
	size_t naudio = j->nframes / j->d1 / j->d2;
	float *audio = (float*)malloc(sizeof(float) * (naudio ? naudio : 1));
	for (unsigned r = 0; r < nrx; r++)
		rx[r] = wro_rx_create(j->fs, j->if_hz[j->rx_lo + r], j->taps1, j->n1, j->d1,
				j->modes[j->rx_lo + r], j->taps2, j->n2, j->d2);
	float acc = 0.0f;
	double t0 = 0.0;
	for (unsigned b = 0; b < j->warmup + j->blocks; b++) {
		if (b == j->warmup) {
			pthread_barrier_wait(j->bar);
			t0 = now_s();
		}
		for (unsigned r = 0; r < nrx; r++) {
			const float *src = j->iq + (size_t)((j->rx_lo + r) % j->n_streams) * j->nframes * 2;
			size_t n = wro_rx_process(rx[r], src, j->nframes, NULL, NULL, NULL, audio);
			if (n)
				acc += audio[n - 1];
		}
	}
	j->seconds = now_s() - t0;
	pthread_barrier_wait(j->bar);
	j->checksum = acc;
	for (unsigned r = 0; r < nrx; r++)
		wro_rx_destroy(rx[r]);
	free(rx);
	free(audio);
	return NULL;
}

double wro_bench(unsigned fs, size_t nframes, unsigned n_rx, unsigned n_streams,
		const int *if_hz, const int *modes,
		const float *taps1, unsigned n1, unsigned d1,
		const float *taps2, unsigned n2, unsigned d2,
		const float *iq, unsigned nthreads, unsigned warmup, unsigned blocks,
		float *audio_checksum)
{
	if (nthreads == 0)
		nthreads = 1;
	if (nthreads > n_rx)
		nthreads = n_rx;
	pthread_t *th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
	struct bench_job *jobs = (struct bench_job*)calloc(nthreads, sizeof(*jobs));
	pthread_barrier_t bar;
	pthread_barrier_init(&bar, NULL, nthreads);
	for (unsigned t = 0; t < nthreads; t++) {
		struct bench_job *j = &jobs[t];
		j->fs = fs; j->nframes = nframes; j->n_streams = n_streams ? n_streams : 1;
		j->rx_lo = (unsigned)((uint64_t)n_rx * t / nthreads);
		j->rx_hi = (unsigned)((uint64_t)n_rx * (t + 1) / nthreads);
		j->if_hz = if_hz; j->modes = modes;
		j->taps1 = taps1; j->n1 = n1; j->d1 = d1;
		j->taps2 = taps2; j->n2 = n2; j->d2 = d2;
		j->iq = iq; j->warmup = warmup; j->blocks = blocks; j->bar = &bar;
		pthread_create(&th[t], NULL, bench_worker, j);
	}
	double worst = 0.0;
	float sum = 0.0f;
	for (unsigned t = 0; t < nthreads; t++) {
		pthread_join(th[t], NULL);
		if (jobs[t].seconds > worst)
			worst = jobs[t].seconds;
		sum += jobs[t].checksum;
	}
	if (audio_checksum)
		*audio_checksum = sum;
	pthread_barrier_destroy(&bar);
	free(jobs);
	free(th);
	return worst;
}
