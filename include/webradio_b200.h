/*
 * webradio_b200.h -- C ABI of libwebradio_b200.so, the B200 (sm_100a) implementation of
 * WebRadio's per-receiver DSP hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types, no
 * exceptions.  Every entry point names the reference interface it replaces
 * (mikestir/webradio @ 6500296d, paths relative to the reference root).  The C++ classes in
 * webradio_b200/dsp and webradio_b200/io (same names and methods as the reference's DspBlock
 * subclasses) are thin callers of these functions; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - samples are float; IQ is interleaved [I0,Q0,I1,Q1,...] (reference src/dsp/dspblock.h:45)
 *   - every function returns 0 on success and a negative WR_E* code on failure; the message is
 *     available from wr_last_error() (thread-local).  A CUDA error never throws or aborts:
 *     it maps to WR_ECUDA, as DspBlock::process() maps failure to `false`
 *     (reference src/dsp/dspblock.cxx:192-195).
 *   - there is NO CPU fallback: without a CUDA device every create call fails with WR_ENODEV.
 *   - one caller thread per handle for the process calls; the per-receiver setters may be
 *     called from any thread and take effect at the next block boundary (the reference applies
 *     them unlocked from HTTP threads: src/web/receiverhandler.cxx:125-140).
 */
#ifndef WEBRADIO_B200_H
#define WEBRADIO_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WR_OK        0
#define WR_EINVAL   (-1)  /* bad argument */
#define WR_ENODEV   (-2)  /* no usable CUDA device */
#define WR_ECUDA    (-3)  /* CUDA runtime error (see wr_last_error) */
#define WR_ENOMEM   (-4)
#define WR_ESTATE   (-5)  /* call not valid in the handle's current state */

/* Demodulator::Mode, same numbering (reference src/dsp/demodulator.h:41-47) */
#define WR_MODE_AM   0
#define WR_MODE_FM   1
#define WR_MODE_USB  2
#define WR_MODE_LSB  3

#define WR_SINTABLE_SIZE 65536u  /* 1 << LOOKUP_BITS, reference src/dsp/downconverter.cxx:36 */

/* wr_rx_reset flags */
#define WR_RESET_PHASE     1u  /* NCO phase          (reference downconverter.h:58) */
#define WR_RESET_CHANNEL   2u  /* channel FIR history (reference lowpass.h:64)      */
#define WR_RESET_DEMOD     4u  /* prev_i / prev_q    (reference demodulator.h:60-61) */
#define WR_RESET_AUDIO     8u  /* audio FIR history                                  */

/* stage ids for wr_bank_read_stage / the strict stage blocks */
#define WR_STAGE_CHANNEL 1     /* channel-filtered IQ, float[2*M1] */
#define WR_STAGE_DEMOD   2     /* demodulated,         float[M1]   */

/* ---------------------------------------------------------------- misc ---- */
const char *wr_version(void);
const char *wr_last_error(void);
int wr_device_count(void);

/* phaseStep for an IF: replaces the expression in DownConverter::setIF / ::init
 * (reference src/dsp/downconverter.cxx:65,80). */
int32_t wr_phase_step(int if_hz, unsigned sample_rate);

/* The NCO table exactly as DownConverter's constructor builds it, with the host libm
 * (reference src/dsp/downconverter.cxx:49-51).  out has WR_SINTABLE_SIZE entries. */
void wr_build_sintable(float *out);

/* Diagnostic: checks that `table` (NULL = the default table) survives the exact 16-bit
 * compression the shared-memory NCO kernels use (webradio_b200/csrc/wr_lo.h).  Returns 0 if
 * every entry is reproduced bit for bit, -1 if the table cannot be represented (v1 kernels). */
int wr_lo_compress_check(const float *table);
/* Same check for the packed-arithmetic compression the v3 kernels use
 * (webradio_b200/csrc/wr_lo3.h).  -1 means the v3 kernels stand aside for v2/v1. */
int wr_lo3_compress_check(const float *table);
/* Diagnostic (host only, no device needed): how the streaming channel kernel (v4, banks of
 * independent tuner streams) cuts n_receivers x n_outputs channel-rate outputs into runs for a
 * persistent grid on n_sms SMs: every receiver into runs_per_receiver runs, the first long_runs
 * of run_len + 1 outputs and the rest of run_len, dealt 32 to a warp in order; `rounds` is what
 * the busiest warp works through, `grid` the CTAs launched, `warps` the warps of a CTA.
 * Returns 1, or 0 if v4 does not serve the geometry / block. */
int wr_plan_runs(unsigned ntaps, unsigned decimation, unsigned n_receivers, unsigned n_outputs, unsigned n_sms,
		unsigned *runs_per_receiver, unsigned *run_len, unsigned *long_runs, unsigned *rounds, unsigned *grid, unsigned *warps);

/* Frequency-sampling low-pass design: replaces LowPass::init (window) + LowPass::recalculate
 * (reference src/dsp/lowpass.cxx:102-110,164-189).  Host code (cold path, K0 in SURVEY.md 2a).
 * ntaps a power of two reproduces the reference; other lengths use the same formulas mod ntaps. */
int wr_lowpass_design(unsigned ntaps, unsigned passband_hz, unsigned sample_rate, float *coeff);

/* ------------------------------------------------- receiver bank (fused) ---- */
/*
 * A bank is the batched form of N `Receiver` chains (reference src/radio.cxx:62-90):
 *   DownConverter::process -> LowPass::process (IQ) -> Demodulator::process -> LowPass::process
 *   (reference src/dsp/downconverter.cxx:91-114, lowpass.cxx:131-162, demodulator.cxx:77-115)
 * for n_receivers receivers fed by n_streams tuner streams on one GPU, with all carried state
 * (NCO phase, both FIR histories, prev I/Q) resident in HBM between calls.
 * All receivers of a bank share the filter geometry (n1,d1,n2,d2); taps, IF, mode and the
 * stream each receiver listens to are per receiver.
 */
typedef struct wr_bank wr_bank;

wr_bank *wr_bank_create(int device, unsigned n_streams, unsigned n_receivers, unsigned max_frames,
		unsigned n1, unsigned d1, unsigned n2, unsigned d2);
void wr_bank_destroy(wr_bank *b);

/* Replace the NCO table (default: wr_build_sintable at create). */
int wr_bank_set_sintable(wr_bank *b, const float *table);

/* Receiver::setFrontEnd analogue: which stream feeds receiver rx (default rx % n_streams). */
int wr_rx_set_stream(wr_bank *b, unsigned rx, unsigned stream);
/* DownConverter::setIF (reference downconverter.cxx:59-67) with the step precomputed. */
int wr_rx_set_phase_step(wr_bank *b, unsigned rx, int32_t step);
/* Coefficients as LowPass::recalculate leaves them in `coeff` (reference lowpass.cxx:182-189),
 * i.e. coeff[0] multiplies the NEWEST sample.  stage 0 = channel filter, 1 = audio filter;
 * ntaps must equal the bank geometry. */
int wr_rx_set_taps(wr_bank *b, unsigned rx, int stage, const float *coeff, unsigned ntaps);
/* LowPass::setPassband for EVERY receiver of the bank at once: the design of
 * wr_lowpass_design (LowPass::recalculate, reference lowpass.cxx:164-189) runs on the device,
 * one CTA per receiver, and leaves the taps where the kernels read them; passband_hz has
 * n_receivers entries.  Bit-identical to calling wr_lowpass_design + wr_rx_set_taps per receiver.
 * wr_rx_get_taps reads a receiver's current coefficients back (reference order, coeff[0] newest). */
int wr_bank_design_taps(wr_bank *b, int stage, const unsigned *passband_hz, unsigned sample_rate);
int wr_rx_get_taps(wr_bank *b, unsigned rx, int stage, float *coeff, unsigned ntaps);
/* Demodulator::setMode (reference demodulator.h:49). */
int wr_rx_set_mode(wr_bank *b, unsigned rx, int mode);
/* Clear carried state (OR of WR_RESET_*); mirrors what ctor/deinit do in the reference. */
int wr_rx_reset(wr_bank *b, unsigned rx, unsigned flags);
/* Overwrite the NCO phase accumulator (31 bits), e.g. to carry a receiver over from another
 * bank; takes effect at the next block boundary. */
int wr_rx_set_phase(wr_bank *b, unsigned rx, uint32_t phase);
/* Read back NCO phase (reference downconverter.h:58) as of the last completed block. */
int wr_rx_get_phase(wr_bank *b, unsigned rx, uint32_t *phase);
/* The FM discriminator's look-back sample {prev_i, prev_q} (reference demodulator.h:60-61; set in
 * the constructor only, so the reference's Demodulator carries it across stop()/start(), e.g.
 * Receiver::setFrontEnd on a live radio, radio.cxx:109-117): read it as of the last completed
 * block, or overwrite it (takes effect at the next block boundary) -- lets a receiver take it
 * along from one bank to another.  prev_iq has two floats. */
int wr_rx_get_lookback(wr_bank *b, unsigned rx, float *prev_iq);
int wr_rx_set_lookback(wr_bank *b, unsigned rx, const float *prev_iq);

/* One block through every receiver, HOST buffers (the DspBlock::process data convention,
 * reference src/dsp/dspblock.cxx:177-195: synchronous, results host-visible on return).
 *   iq_host    : [n_streams][nframes][2] float
 *   audio_host : receiver r's floor(floor(nframes/d1)/d2) output frames start at r*audio_stride
 * Inside the one synchronous call a block of more than ~0.5 MB is cut into up to four
 * consecutive sub-blocks, so that the copy in of one runs under the kernels of the one before and
 * under the copy out of the one before that; the carried state makes the samples the same
 * (env WR_SYNC_SPLIT=n forces n pieces, 1 = none).  Pinned buffers copy at link speed; pageable
 * ones work and are staged by the driver. */
int wr_bank_process(wr_bank *b, const float *iq_host, unsigned nframes,
		float *audio_host, size_t audio_stride);

/* ------------------------------------------------ shared tuner-block upload ---- */
/* One DspBlock::run of a tuner pushes the SAME host buffer to every consumer (the SpectrumSink
 * first, then each receiver: reference src/dsp/dspblock.cxx:207-209, src/radio.cxx:126-128,151-156).
 * A wr_upload carries that block to one device ONCE: wr_upload_begin page-locks the caller's buffer
 * where it lies (first sight only; a DspBlock keeps its output vector from block to block) and
 * starts asynchronous copies in up to four pieces; the *_process_upload calls of the bank and of
 * the spectrum sink then read the device copy, each starting as soon as the pieces it needs have
 * landed.  wr_upload_finish returns once the host buffer is no longer needed -- the producer calls
 * it before it touches its buffer again (end of DspBlock::run).  One caller thread per handle. */
typedef struct wr_upload wr_upload;
wr_upload *wr_upload_create(int device, size_t max_frames);
void wr_upload_destroy(wr_upload *u);
size_t wr_upload_capacity(const wr_upload *u);
int wr_upload_device(const wr_upload *u);
int wr_upload_begin(wr_upload *u, const float *iq_host, unsigned nframes);   /* [nframes][2] floats */
int wr_upload_finish(wr_upload *u);
/* wr_bank_process on the upload's device copy (one stream; same results, same sub-block overlap:
 * the kernels of a sub-block start when its part of the upload has landed, its audio leaves while
 * the next one is computed); synchronous -- audio_host is complete on return. */
int wr_bank_process_upload(wr_bank *b, wr_upload *u, unsigned nframes, float *audio_host, size_t audio_stride);
/* Page-locked host memory for audio_host and friends (copies to pageable memory are staged by the
 * driver at a fraction of the link speed). */
void *wr_host_alloc(size_t bytes);
void wr_host_free(void *p);

/* The FIR histories of one receiver as of the last completed block -- what LowPass keeps in `block`
 * between calls (reference lowpass.h:64, lowpass.cxx:133-142): stage 0 = the last n1-1 MIXED frames
 * (2*(n1-1) floats, interleaved IQ), stage 1 = the last n2-1 demodulated samples.  With
 * wr_rx_get/set_phase and wr_rx_get/set_lookback this is all the state a receiver carries, so a
 * host can move it from one bank to another (a bank rebuilt because receivers joined or left a
 * running front-end) without a glitch.  Call between blocks, from the processing thread. */
int wr_rx_get_history(wr_bank *b, unsigned rx, int stage, float *out, unsigned nfloats);
int wr_rx_set_history(wr_bank *b, unsigned rx, int stage, const float *in, unsigned nfloats);

/* Same, with input and output already in HBM on the bank's device; asynchronous on
 * cuda_stream (a cudaStream_t; NULL = the bank's own stream, wr_bank_stream).  The bank's stream
 * is non-blocking: it does NOT synchronise with the legacy default stream, whose handle is also
 * NULL and which therefore cannot be named here -- a caller that produces iq_dev or consumes
 * audio_dev with its own kernels passes the stream those kernels run in, or works in
 * wr_bank_stream(b). */
int wr_bank_process_device(wr_bank *b, const float *iq_dev, size_t stream_stride_frames,
		unsigned nframes, float *audio_dev, size_t audio_stride, void *cuda_stream);

/* Pipelined host path: enqueue one block (H2D copy, kernels, D2H copy all asynchronous, from
 * / into caller-owned PINNED buffers) and wait for the oldest outstanding one.  Up to
 * wr_bank_pipeline_depth() blocks may be in flight; a caller must not touch a block's buffers
 * between its wr_bank_submit and the wr_bank_wait that returns it. */
int wr_bank_submit(wr_bank *b, const float *iq_pinned, unsigned nframes,
		float *audio_pinned, size_t audio_stride);
int wr_bank_wait(wr_bank *b);
int wr_bank_pipeline_depth(const wr_bank *b);
/* How a block travels between the copy-in stream, the two kernels and the host on the pipelined
 * path (v3 channel kernel; the v1/v2 kernels always use events):
 *   WR_HANDOVER_FLAGS  (default) the copy-in stream raises a counter in HBM behind the tuner block
 *                      (a stream memory operation; a 4-byte copy on drivers without them) and the
 *                      channel kernel's loaders wait for it.  The launch stream holds nothing but
 *                      kernels and event records, so consecutive blocks overlap as they do on the
 *                      device-resident path (a cudaStreamWaitEvent in front of the channel kernel
 *                      costs that overlap: measured, ~6 us per 22 us block).  The audio leaves
 *                      through an event and a copy on the copy-out stream;
 *   WR_HANDOVER_EVENTS CUDA events in both directions;
 *   WR_HANDOVER_DIRECT flags in; out, the demodulator kernel stores the audio straight into the
 *                      caller's pinned buffer and its last CTA raises a counter in mapped host
 *                      memory that wr_bank_wait polls (no copy-out stream at all: lowest latency
 *                      for one block at a time, but the kernel then runs at PCIe write speed).
 * Returns the scheme, or a negative WR_E* code.  Call it with no blocks in flight. */
#define WR_HANDOVER_EVENTS 0
#define WR_HANDOVER_FLAGS  1
#define WR_HANDOVER_DIRECT 2
int wr_bank_set_handover(wr_bank *b, int scheme);

/* Host-language loop helpers (what a C++ caller would write itself; they keep interpreter
 * overhead out of bench.py's timed regions).  Step i uses iq[(first + i) % n_iq] and
 * audio[(first + i) % n_audio].
 *   wr_bank_run_device_steps: `steps` calls of wr_bank_process_device on the bank's stream;
 *   wr_bank_run_host_steps  : `steps` blocks through wr_bank_submit / wr_bank_wait with the
 *                             pipeline kept full (pipelined != 0) or strictly one at a time. */
int wr_bank_run_device_steps(wr_bank *b, const float *const *iq_dev, unsigned n_iq, size_t stream_stride_frames,
		unsigned nframes, float *const *audio_dev, unsigned n_audio, size_t audio_stride,
		unsigned first, unsigned steps);
int wr_bank_run_host_steps(wr_bank *b, const float *const *iq_pinned, unsigned n_iq, unsigned nframes,
		float *const *audio_pinned, unsigned n_audio, size_t audio_stride,
		unsigned first, unsigned steps, int pipelined);
/* Raw RTL-SDR bytes instead of floats: the same five calls with the tuner
 * block as interleaved unsigned 8-bit I/Q, [n_streams][nframes][2] bytes.  The tuner's sample
 * conversion ((float)b - 128.0) / 128.0 (reference src/io/rtlsdrtuner.cxx:104-108) runs inside
 * the channel kernel's load, so a frame costs 2 bytes of PCIe and HBM traffic instead of 8;
 * results are bit-identical to converting on the host and calling the float entry points.
 * stream_stride_frames is in FRAMES (2 bytes each). */
int wr_bank_process_u8(wr_bank *b, const uint8_t *iq_host, unsigned nframes,
		float *audio_host, size_t audio_stride);
int wr_bank_process_device_u8(wr_bank *b, const uint8_t *iq_dev, size_t stream_stride_frames,
		unsigned nframes, float *audio_dev, size_t audio_stride, void *cuda_stream);
int wr_bank_submit_u8(wr_bank *b, const uint8_t *iq_pinned, unsigned nframes,
		float *audio_pinned, size_t audio_stride);
int wr_bank_run_device_steps_u8(wr_bank *b, const uint8_t *const *iq_dev, unsigned n_iq, size_t stream_stride_frames,
		unsigned nframes, float *const *audio_dev, unsigned n_audio, size_t audio_stride,
		unsigned first, unsigned steps);
int wr_bank_run_host_steps_u8(wr_bank *b, const uint8_t *const *iq_pinned, unsigned n_iq, unsigned nframes,
		float *const *audio_pinned, unsigned n_audio, size_t audio_stride,
		unsigned first, unsigned steps, int pipelined);
void *wr_bank_stream(wr_bank *b);   /* cudaStream_t of the bank */
int wr_bank_sync(wr_bank *b);

/* Keep (1) or drop (0, default) the intermediate channel-IQ stream of the fused kernel so
 * that wr_bank_read_stage(WR_STAGE_CHANNEL) can return it; the demod stream is always kept. */
int wr_bank_keep_channel(wr_bank *b, int keep);
/* Copy an intermediate stream of the LAST block of receiver rx to the host.  Returns the
 * number of floats written, or a negative error. */
long wr_bank_read_stage(wr_bank *b, unsigned rx, int stage, float *out_host, size_t cap_floats);

/* Audio sample format of every later block: WR_AUDIO_FLOAT (default) is the
 * DspBlock convention, WR_AUDIO_LAME multiplies by 32768 in the audio kernel's store -- the
 * "LAME wants +/-32768" conversion MP3Encoder::encode does per sample before
 * lame_encode_buffer_float (reference src/web/mp3encoder.cxx:66-73). */
#define WR_AUDIO_FLOAT 0
#define WR_AUDIO_LAME  1
int wr_bank_set_audio_format(wr_bank *b, int format);

/* Selects the kernel family: 0 = auto (default: the newest one that supports the geometry and
 * the block length), 1 = v1 generic kernels (NCO table read from L2), 2 = v2 kernels (NCO table
 * resident in shared memory, tile per work item), 3 = v3 kernels (streaming ring of mixed
 * slots, packed NCO arithmetic), 4 = v4 kernels (streaming FIR, one thread per run of outputs:
 * banks large enough to fill the GPU with long runs, float or raw bytes).  For tests and profiling. */
int wr_bank_set_variant(wr_bank *b, int variant);
/* Which family ran the last block: 1 ... 4 (0 before the first block). */
int wr_bank_variant_in_use(const wr_bank *b);
/* Kernel launches issued by this bank since creation (for bench.py's gpu_launches). */
unsigned long long wr_bank_launch_count(const wr_bank *b);
/* Per-launch device timing with CUDA events on the launch stream.  While enabled every block
 * records events around [0] the fused mix+FIR+demod kernel and [1] the audio FIR kernel;
 * wr_bank_kernel_times returns the summed milliseconds and the number of blocks since timing
 * was enabled.  This is what fills DspBlock's profile counters
 * (reference src/dsp/dspblock.cxx:186-204) and bench.py's roofline. */
int wr_bank_set_timing(wr_bank *b, int on);
int wr_bank_kernel_times(wr_bank *b, double *ms2_total, unsigned long long *nblocks);

/* --------------------------------------------- strict single-stage blocks ---- */
/* One kernel per process() call on host buffers: what each reference block does on its own.
 * Used by the DspBlock drop-ins when a chain cannot be fused, and by stage-by-stage parity. */
/* The FM discriminator's atan2f (reference src/dsp/demodulator.cxx:97 calls the host libm):
 * host twin of the routine the kernels run (webradio_b200/csrc/wr_atan2f.h, a restatement of
 * glibc's e_atan2f.c / s_atanf.c).  Exists so that tests can pin it against the installed libm. */
void wr_atan2f_host(const float *y, const float *x, size_t n, float *out);

typedef struct wr_stage wr_stage;

wr_stage *wr_stage_create(int device);
void wr_stage_destroy(wr_stage *s);
/* DownConverter::process (reference downconverter.cxx:91-114). *phase is read and advanced. */
int wr_stage_mix(wr_stage *s, const float *table_or_null, uint32_t *phase, int32_t step,
		const float *iq_host, unsigned nframes, float *out_host);
/* LowPass::process (reference lowpass.cxx:131-162): history lives in the stage handle. */
int wr_stage_fir_config(wr_stage *s, unsigned channels, const float *coeff, unsigned ntaps);
int wr_stage_fir(wr_stage *s, const float *in_host, unsigned nframes, unsigned decim, float *out_host);
int wr_stage_fir_reset(wr_stage *s);
/* Demodulator::process (reference demodulator.cxx:77-115). prev[2] read and updated. */
int wr_stage_demod(wr_stage *s, int mode, float *prev, const float *iq_host, unsigned nframes,
		float *out_host);
/* The waterfall palette map of wr_spectrum_get_palette over an arbitrary host array of dB values. */
int wr_stage_palette(wr_stage *s, const float *db_host, unsigned n, uint8_t *index_host);
/* Test hook: the device's atan2f over host arrays (one kernel). */
int wr_stage_atan2f(wr_stage *s, const float *y_host, const float *x_host, unsigned n, float *out_host);

/* ----------------------------------------------------------- spectrum ---- */
/* SpectrumSink (reference src/io/spectrumsink.cxx:60-142): Hamming window, forward complex
 * FFT, 10*log10(re^2+im^2) - 20*log10(N), fft-shifted.  n_streams independent streams are
 * transformed in one launch; hop == fft_size is the reference behaviour, hop < fft_size is
 * the overlapped waterfall of BASELINE config 4. */
typedef struct wr_spectrum wr_spectrum;

wr_spectrum *wr_spectrum_create(int device, unsigned fft_size, unsigned hop, unsigned n_streams,
		unsigned max_frames);
void wr_spectrum_destroy(wr_spectrum *s);
/* SpectrumSink::process: feed nframes per stream ([n_streams][nframes][2] host floats).
 * If rows_host is non-NULL every completed FFT frame's dB row is returned:
 * stream t's rows start at rows_host + t*row_stride_floats, fft_size floats each.
 * Returns the number of rows completed per stream by this call (>= 0) or a negative error. */
long wr_spectrum_process(wr_spectrum *s, const float *iq_host, unsigned nframes,
		float *rows_host, size_t row_stride_floats);
/* SpectrumSink::process on a shared upload (see wr_upload): asynchronous -- the newest complete
 * frame is transformed behind the upload on the sink's stream, wr_spectrum_get synchronises.
 * Returns the rows completed (only the last one is kept). */
long wr_spectrum_process_upload(wr_spectrum *s, wr_upload *u, unsigned nframes);
/* Raise max_frames (a tuner whose block length grew) WITHOUT losing the carried partial frame or
 * the last row, which destroying and re-creating the handle would. */
int wr_spectrum_reserve(wr_spectrum *s, unsigned max_frames);
/* Device-resident variant (asynchronous on cuda_stream). */
long wr_spectrum_process_device(wr_spectrum *s, const float *iq_dev, size_t stream_stride_frames,
		unsigned nframes, float *rows_dev, size_t row_stride_floats, void *cuda_stream);
/* SpectrumSink::getSpectrum (reference spectrumsink.cxx:125-142): dB of the most recent
 * transform of one stream; fft_size floats. */
int wr_spectrum_get(wr_spectrum *s, unsigned stream, float *db_host);
/* The same row as the 256-entry palette index the browser computes for it:
 * WaterfallHandler::doGet replaces non-finite bins by -10000.0 (reference
 * src/web/waterfallhandler.cxx:59-69) and Waterfall.update maps
 *   floor(((dB + 50.0) / 25.0) * 255.0) clamped to [0, 255]     (html/waterfall.js:92-109)
 * in double arithmetic; one byte per bin leaves the GPU instead of a float. */
int wr_spectrum_get_palette(wr_spectrum *s, unsigned stream, uint8_t *index_host);
unsigned long long wr_spectrum_launch_count(const wr_spectrum *s);
int wr_spectrum_sync(wr_spectrum *s);

#ifdef __cplusplus
}
#endif
#endif /* WEBRADIO_B200_H */
