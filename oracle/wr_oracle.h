/*
 * oracle/wr_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's per-receiver DSP path
 * (mikestir/webradio src/dsp + src/io/spectrumsink.cxx).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it, and only as the checker or the timed CPU baseline -- never as
 * a product path.  Pinned against oracle/_ref (the unmodified reference
 * compiled in place) and the fixtures in tests/golden/ generated from it.
 */
#ifndef WR_ORACLE_H
#define WR_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { WRO_AM = 0, WRO_FM = 1, WRO_USB = 2, WRO_LSB = 3 };

#define WRO_TABLE_SIZE 65536u

/* NCO sine table: reference downconverter.cxx:49-51 */
void wro_sintable(float *out65536);
/* phase step for an IF: reference downconverter.cxx:59-67,80 */
int32_t wro_phase_step(int if_hz, unsigned fs);
/* windowed frequency-sampling design: reference lowpass.cxx:102-110,164-189.
 * n must be a power of two to be the reference's design; other n use the
 * same formulas with "mod n" in place of the reference's "& (n-1)". */
void wro_lowpass_design(unsigned n, unsigned passband_hz, unsigned fs, float *coeff);
/* spectrum window: reference spectrumsink.cxx:71-74 */
void wro_spectrum_window(unsigned n, float *window);

/* ---- stage kernels with explicit state (each mirrors one process()) ---- */

/* reference downconverter.cxx:91-114; phase is read and updated */
void wro_mix(const float *table, uint32_t *phase, int32_t step,
		const float *iq, size_t nframes, float *out);

typedef struct wro_fir wro_fir; /* LowPass::process state: reference lowpass.h:60-64 */
wro_fir *wro_fir_create(unsigned channels, const float *coeff, unsigned ntaps, unsigned decim);
void wro_fir_set_taps(wro_fir *f, const float *coeff, unsigned ntaps);
/* reference lowpass.cxx:131-162; returns output frames = floor(nframes/decim) */
size_t wro_fir_process(wro_fir *f, const float *in, size_t nframes, float *out);
size_t wro_fir_get_history(const wro_fir *f, float *out);
void wro_fir_set_history(wro_fir *f, const float *in);
void wro_fir_destroy(wro_fir *f);

/* reference demodulator.cxx:77-115; prev[2] = {prev_i, prev_q} is read and updated */
int wro_demod(int mode, float *prev, const float *iq, size_t nframes, float *out);
/* reference src/web/waterfallhandler.cxx:62-68 + html/waterfall.js:92-109: palette index of a dB row */
void wro_waterfall_index(const float *db, size_t n, unsigned char *out);
/* reference src/web/mp3encoder.cxx:66-73: the encoder's +/-32768 sample format */
void wro_lame_scale(const float *x, size_t n, float *out);
/* reference src/io/rtlsdrtuner.cxx:104-108: raw bytes to samples */
void wro_rtlsdr_convert(const unsigned char *buf, size_t n, float *out);
/* the host libm's atan2f over arrays: what reference demodulator.cxx:97 calls on this box */
void wro_libm_atan2f(const float *y, const float *x, size_t n, float *out);

/* ---- one whole receiver (reference radio.cxx:62-90 chain) ---- */
typedef struct wro_rx wro_rx;
wro_rx *wro_rx_create(unsigned fs, int if_hz,
		const float *taps1, unsigned n1, unsigned d1, int mode,
		const float *taps2, unsigned n2, unsigned d2);
void wro_rx_set_if(wro_rx *r, int if_hz);
void wro_rx_set_mode(wro_rx *r, int mode);
void wro_rx_set_taps(wro_rx *r, int which, const float *taps, unsigned n);
/* Any of mixed/chan/demod may be NULL.  Returns audio frames written. */
size_t wro_rx_process(wro_rx *r, const float *iq, size_t nframes,
		float *mixed, float *chan, float *demod, float *audio);
void wro_rx_destroy(wro_rx *r);

/* ---- SpectrumSink: reference spectrumsink.cxx:60-142 ---- */
typedef struct wro_spectrum wro_spectrum;
/* hop == n is the reference behaviour; hop < n (overlap) is the cfg4 extension */
wro_spectrum *wro_spectrum_create(unsigned n, unsigned hop);
/* Feeds frames; every completed FFT frame's dB row is appended to rows (if
 * non-NULL, capacity max_rows*n floats).  Returns the number of FFT frames
 * completed by this call. */
size_t wro_spectrum_process(wro_spectrum *s, const float *iq, size_t nframes,
		float *rows, size_t max_rows);
/* reference spectrumsink.cxx:125-142 on the most recent transform */
void wro_spectrum_get(const wro_spectrum *s, float *db);
/* raw (unshifted) complex bins of the most recent transform, float[2n] */
void wro_spectrum_get_bins(const wro_spectrum *s, float *bins);
void wro_spectrum_destroy(wro_spectrum *s);

/* ---- timed CPU baseline (bench.py cpu_baseline "port") ----
 * nthreads worker threads, each owning n_rx/nthreads receivers, all fed the
 * same n_streams==1 block (shared tuner) or their own stream (iq holds
 * n_streams blocks back to back, receiver r reads stream r % n_streams).
 * Returns seconds for `blocks` timed blocks after `warmup` untimed ones. */
double wro_bench(unsigned fs, size_t nframes, unsigned n_rx, unsigned n_streams,
		const int *if_hz, const int *modes,
		const float *taps1, unsigned n1, unsigned d1,
		const float *taps2, unsigned n2, unsigned d2,
		const float *iq, unsigned nthreads, unsigned warmup, unsigned blocks,
		float *audio_checksum);

#ifdef __cplusplus
}
#endif
#endif
