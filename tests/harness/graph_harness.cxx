/*
 * tests/harness/graph_harness.cxx -- TEST DRIVER (not product code).
 *
 * Builds the receiver graph the reference's Radio glue builds
 * (reference src/radio.cxx:62-90,120-133: tuner -> SpectrumSink, and per
 * receiver DownConverter -> LowPass -> Demodulator -> LowPass) out of whatever
 * DspBlock implementation it is compiled against, using only the public
 * plugin surface (reference src/dsp/dspblock.h:60-79,132-137), and exposes it
 * through a tiny C ABI so pytest can drive it with ctypes.
 *
 * It is compiled twice from this one source:
 *   - oracle/Makefile: against the UNMODIFIED reference sources in
 *     /root/reference/src (-DWR_REFERENCE_BUILD) -> oracle/_ref/libwr_ref.so
 *     (the real reference, used as checker and as CPU baseline);
 *   - Makefile (repo root): against webradio_b200/dsp + webradio_b200/io, the
 *     GPU-backed drop-in blocks -> tests/harness/libwr_blocks_harness.so.
 * Same graph, same inputs, so the parity tests read like the reference's own
 * tests would.
 */
#include <mutex>
#include <vector>
#include <string>
#include <algorithm>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <map>
#include <unistd.h>
#include <fcntl.h>

#ifdef WR_REFERENCE_BUILD
/* The reference fixes the tap count at compile time (lowpass.cxx:39) and keeps
 * the coefficient vector private (lowpass.h:53,60).  To exercise the
 * unmodified LowPass::process loop with 127/255 injected taps, and to read the
 * NCO table (downconverter.h:57), this one test TU opens up access.  Class
 * layout is unaffected. */
#define private public
#endif
#include "dspblock.h"
#include "lowpass.h"
#include "downconverter.h"
#include "demodulator.h"
#include "spectrumsink.h"
#ifdef WR_REFERENCE_BUILD
#undef private
#endif

namespace {

/* Source that replays a caller-supplied interleaved IQ block. */
class ReplaySource : public DspSource {
public:
	ReplaySource() : DspSource("replay", "ReplaySource"), cur(NULL), curLen(0) {}
	void feed(const float *p, size_t n) { cur = p; curLen = n; }
protected:
	bool init() {
		_outputSampleRate = inputSampleRate();
		_outputChannels = inputChannels();
		return true;
	}
	void deinit() {}
	bool process(const vector<sample_t> &in, vector<sample_t> &out) {
		(void)in;
		if (!cur || curLen != out.size())
			return false;
		memcpy(out.data(), cur, curLen * sizeof(float));
		return true;
	}
private:
	const float *cur;
	size_t curLen;
};

/* Sink that keeps a copy of the last block pushed into it. */
class TapSink : public DspBlock {
public:
	TapSink(const string &name) : DspBlock(name, "TapSink") {}
	vector<float> last;
protected:
	bool init() { return true; }
	void deinit() {}
	bool process(const vector<sample_t> &in, vector<sample_t> &out) {
		(void)out;
		last.assign(in.begin(), in.end());
		return true;
	}
};

struct Rx {
	DownConverter *dc;
	LowPass *chan;
	Demodulator *demod;
	LowPass *audio;
	TapSink *tap[4];
};

struct Graph {
	ReplaySource *src;
	unsigned fs;
	unsigned blockFrames;
	vector<Rx> rx;
	SpectrumSink *spectrum;
	bool started;
	unsigned long runs;
};

// Silences the blocks' LOG_DEBUG chatter by pointing fd 2 at /dev/null for the duration of a call.
// fd 2 is process-wide and bench.py runs one graph per host thread, so the redirection is counted:
// the first thread in redirects, the last one out restores (two threads saving and restoring on
// their own left stderr on /dev/null for good -- and swallowed the caller's own error messages).
struct QuietStderr {
	bool on;
	static std::mutex &mu() { static std::mutex m; return m; }
	static int &depth() { static int d = 0; return d; }
	static int &saved() { static int s = -1; return s; }
	QuietStderr(bool want) : on(want) {
		if (!on) return;
		std::lock_guard<std::mutex> lk(mu());
		if (depth()++ == 0) {
			fflush(stderr);
			saved() = dup(2);
			int nul = open("/dev/null", O_WRONLY);
			if (nul >= 0) { dup2(nul, 2); close(nul); }
		}
	}
	~QuietStderr() {
		if (!on) return;
		std::lock_guard<std::mutex> lk(mu());
		if (--depth() == 0 && saved() >= 0) { fflush(stderr); dup2(saved(), 2); close(saved()); saved() = -1; }
	}
};

bool g_quiet = true;

} // namespace

extern "C" {

/* 0 = the reference itself, 1 = the GPU-backed drop-in blocks */
int wrh_flavour(void)
{
#ifdef WR_REFERENCE_BUILD
	return 0;
#else
	return 1;
#endif
}

void wrh_set_quiet(int quiet) { g_quiet = quiet != 0; }

void *wrh_graph_create(unsigned fs, unsigned block_frames)
{
	Graph *g = new Graph();
	g->src = new ReplaySource();
	g->fs = fs;
	g->blockFrames = block_frames;
	g->spectrum = NULL;
	g->started = false;
	g->runs = 0;
	return g;
}

/* capture_mask bit s: attach a TapSink after stage s
 * (0 mixed IQ, 1 channel-filtered IQ, 2 demodulated, 3 audio).
 * ch_rate/au_rate: requested output rate, or 0 to use the *_decim value. */
int wrh_graph_add_receiver(void *h, int if_hz, unsigned ch_passband, unsigned ch_rate,
		unsigned ch_decim, int mode, unsigned au_passband, unsigned au_rate,
		unsigned au_decim, unsigned capture_mask)
{
	Graph *g = (Graph*)h;
#ifdef WR_REFERENCE_BUILD
	/* A receiver connected to a live tuner is started with the default 48 kHz input rate
	 * (dspblock.cxx:59-62 does not cascade rates), which makes the reference's LowPass interpolate
	 * by 5 and read far beyond its block (lowpass.cxx:147-159): not something to run. */
	if (g->started)
		return -1;
#endif
	QuietStderr quiet(g_quiet);
	char nm[32];
	snprintf(nm, sizeof(nm), "%04X", (unsigned)g->rx.size());
	Rx r;
	r.dc = new DownConverter(nm);
	r.chan = new LowPass(nm);
	r.demod = new Demodulator(nm);
	r.audio = new LowPass(nm);
	DspBlock *stage[4] = { r.dc, r.chan, r.demod, r.audio };
	for (int s = 0; s < 4; s++) {
		r.tap[s] = NULL;
		if (capture_mask & (1u << s)) {
			r.tap[s] = new TapSink(nm);
			stage[s]->connect(r.tap[s]);
		}
	}
	r.dc->connect(r.chan);
	r.chan->connect(r.demod);
	r.demod->connect(r.audio);

	r.dc->setIF(if_hz);
	r.chan->setPassband(ch_passband);
	if (ch_rate) r.chan->setOutputSampleRate(ch_rate); else r.chan->setDecimation(ch_decim);
	r.audio->setPassband(au_passband);
	if (au_rate) r.audio->setOutputSampleRate(au_rate); else r.audio->setDecimation(au_decim);
	r.demod->setMode((Demodulator::Mode)mode);

	g->src->connect(r.dc);
	g->rx.push_back(r);
	return (int)g->rx.size() - 1;
}

int wrh_graph_add_spectrum(void *h, unsigned fft_size)
{
	Graph *g = (Graph*)h;
	if (g->started || g->spectrum)
		return -1;
	QuietStderr quiet(g_quiet);
	g->spectrum = new SpectrumSink("0000");
	g->spectrum->setFftSize(fft_size);
	g->src->connect(g->spectrum);
	return 0;
}

int wrh_graph_start(void *h)
{
	Graph *g = (Graph*)h;
	QuietStderr q(g_quiet);
	g->src->setSampleRate(g->fs);
	g->src->setChannels(2);
	g->src->setBlockSize(g->blockFrames * 2);
	g->started = g->src->start();
	return g->started ? 0 : -1;
}

int wrh_graph_run(void *h, const float *iq)
{
	Graph *g = (Graph*)h;
	// only the first block resizes buffers (and logs it); later blocks run without the two dup2 calls
	QuietStderr q(g_quiet && g->runs == 0);
	g->runs++;
	g->src->feed(iq, (size_t)g->blockFrames * 2);
	return g->src->run() ? 0 : -1;
}

/* Copies the last captured block of (rx, stage) into out; returns the number
 * of floats available (may exceed cap), or -1 if that tap was not attached. */
long wrh_graph_get(void *h, int rx, int stage, float *out, long cap)
{
	Graph *g = (Graph*)h;
	if (rx < 0 || rx >= (int)g->rx.size() || stage < 0 || stage > 3 || !g->rx[rx].tap[stage])
		return -1;
	const vector<float> &v = g->rx[rx].tap[stage]->last;
	long n = (long)v.size();
	if (out && cap > 0)
		memcpy(out, v.data(), sizeof(float) * (size_t)std::min(n, cap));
	return n;
}

/* Receiver::setFrontEnd(NULL) / setFrontEnd(frontEnd) on a live pipeline (reference
 * radio.cxx:109-117,151-163): the tuner disconnects the chain's first block, which stops the
 * chain (dspblock.cxx:78-91), or connects it again, which starts it on the spot with the rates it
 * was last given (dspblock.cxx:57-76). */
int wrh_graph_detach(void *h, int rx)
{
	Graph *g = (Graph*)h;
	if (rx < 0 || rx >= (int)g->rx.size())
		return -1;
	QuietStderr q(g_quiet);
	g->runs = 0; // buffers may be resized (and logged) again
	g->src->disconnect(g->rx[rx].dc);
	return 0;
}

int wrh_graph_attach(void *h, int rx)
{
	Graph *g = (Graph*)h;
	if (rx < 0 || rx >= (int)g->rx.size())
		return -1;
	QuietStderr q(g_quiet);
	g->runs = 0;
	g->src->connect(g->rx[rx].dc);
	return g->rx[rx].dc->isRunning() ? 0 : -1;
}

/* DspSource::stop() then start() (what a tuner restart does, reference dspblock.cxx:106-167): every
 * block is deinitialised and initialised again. */
int wrh_graph_restart(void *h)
{
	Graph *g = (Graph*)h;
	QuietStderr q(g_quiet);
	g->runs = 0;
	g->src->stop();
	g->started = g->src->start();
	return g->started ? 0 : -1;
}

int wrh_graph_set_if(void *h, int rx, int hz)
{
	Graph *g = (Graph*)h;
	g->rx[rx].dc->setIF(hz);
	return g->rx[rx].dc->IF();
}

int wrh_graph_set_mode(void *h, int rx, const char *mode)
{
	Graph *g = (Graph*)h;
	return g->rx[rx].demod->setModeString(mode) ? 0 : -1;
}

int wrh_graph_get_mode(void *h, int rx)
{
	Graph *g = (Graph*)h;
	return (int)g->rx[rx].demod->mode();
}

int wrh_graph_set_passband(void *h, int rx, int which, unsigned hz)
{
	Graph *g = (Graph*)h;
	LowPass *lp = which ? g->rx[rx].audio : g->rx[rx].chan;
	lp->setPassband(hz);
	return (int)lp->passband();
}

/* Replace the designed taps of a running filter (which: 0 channel, 1 audio). */
int wrh_graph_set_taps(void *h, int rx, int which, const float *taps, unsigned n)
{
	Graph *g = (Graph*)h;
	if (!g->started)
		return -1;
	LowPass *lp = which ? g->rx[rx].audio : g->rx[rx].chan;
#ifdef WR_REFERENCE_BUILD
	lp->_firLength = n;
	lp->coeff.assign(taps, taps + n);
	vector<sample_t>().swap(lp->block); /* history restarts from zeros at the new length */
	return 0;
#else
	return lp->setCoefficients(taps, n) ? 0 : -1;
#endif
}

/* Read back the taps currently in use. */
int wrh_graph_get_taps(void *h, int rx, int which, float *out, unsigned cap)
{
	Graph *g = (Graph*)h;
	LowPass *lp = which ? g->rx[rx].audio : g->rx[rx].chan;
#ifdef WR_REFERENCE_BUILD
	unsigned n = (unsigned)lp->coeff.size();
	if (out)
		memcpy(out, lp->coeff.data(), sizeof(float) * std::min(n, cap));
	return (int)n;
#else
	return (int)lp->coefficients(out, cap);
#endif
}

/* Block rates as negotiated by DspBlock::start (reference dspblock.cxx:106-151). */
int wrh_graph_rates(void *h, int rx, unsigned *out8)
{
	Graph *g = (Graph*)h;
	const Rx &r = g->rx[rx];
	out8[0] = r.dc->outputSampleRate();
	out8[1] = r.chan->outputSampleRate();
	out8[2] = r.chan->DspBlock::decimation();
	out8[3] = r.demod->outputSampleRate();
	out8[4] = r.demod->outputChannels();
	out8[5] = r.audio->outputSampleRate();
	out8[6] = r.audio->DspBlock::decimation();
	out8[7] = r.audio->outputChannels();
	return 0;
}

int wrh_graph_spectrum(void *h, float *db)
{
	Graph *g = (Graph*)h;
	if (!g->spectrum)
		return -1;
	g->spectrum->getSpectrum(db);
	return (int)g->spectrum->fftSize();
}

/* Reference profile counters (reference dspblock.h:69-75): ns spent in
 * process() and frames consumed, per stage of one receiver. */
int wrh_graph_profile(void *h, int rx, unsigned long long *ns4, unsigned long long *frames4)
{
	Graph *g = (Graph*)h;
	const Rx &r = g->rx[rx];
	DspBlock *stage[4] = { r.dc, r.chan, r.demod, r.audio };
	for (int s = 0; s < 4; s++) {
		ns4[s] = stage[s]->totalNanoseconds();
		frames4[s] = stage[s]->totalIn();
	}
	return 0;
}

void wrh_graph_destroy(void *h)
{
	Graph *g = (Graph*)h;
	QuietStderr q(g_quiet);
	/* stop the pipeline first (the source's destructor would do it too late:
	 * consumers must still be alive when stop() cascades) */
	if (g->started)
		g->src->stop();
	for (size_t i = 0; i < g->rx.size(); i++) {
		Rx &r = g->rx[i];
		delete r.dc; delete r.chan; delete r.demod; delete r.audio;
		for (int s = 0; s < 4; s++) delete r.tap[s];
	}
	delete g->spectrum;
	delete g->src;
	delete g;
}

#ifdef WR_REFERENCE_BUILD
/* The NCO table exactly as the reference builds it (downconverter.cxx:49-51). */
int wrh_ref_sintable(float *out65536)
{
	QuietStderr q(g_quiet);
	DownConverter dc("tbl");
	memcpy(out65536, dc.sinTable.data(), sizeof(float) * dc.sinTable.size());
	return (int)dc.sinTable.size();
}
#endif

} // extern "C"
