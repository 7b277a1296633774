/*
 * tests/harness/radio_dropin.cxx -- TEST DRIVER for the drop-in claim.
 *
 * Links the reference's UNMODIFIED graph glue (/root/reference/src/radio.cxx, compiled in place)
 * against webradio_b200's DspBlock drop-in classes and drives it exactly as the reference's
 * main() does (src/main.cxx:71-115): FrontEnd(factory) -> Receivers -> tuner()->start() ->
 * Radio::run().  Only AudioStreamManager is a stub (tests/harness/stubs/audiostream.h).
 * Built into tests/harness/libwr_radio_dropin.so where the reference tree is mounted; the
 * prebuilt library travels to the GPU box.
 */
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <unistd.h>
#include <fcntl.h>

#include "radio.h"
#include "gpubank.h"

namespace {

class ReplayTuner : public Tuner {
public:
	ReplayTuner(const string &name) : Tuner(name, "ReplayTuner"), cur(NULL), len(0), head(0) {}
	void feed(const float *p, size_t n) { cur = p; len = n; }
	// RtlSdrTuner's hand-over (reference src/io/rtlsdrtuner.cxx:265-285): a ring of block-sized vectors
	// filled elsewhere, each SWAPPED into the block's output buffer -- no copy, and the buffer the
	// consumers see changes from block to block
	void ringFill(const float *p, size_t n) { ring.push_back(vector<sample_t>(p, p + n)); }
	bool ringMode() const { return !ring.empty(); }
private:
	bool init() { _outputSampleRate = inputSampleRate(); _outputChannels = inputChannels(); return true; }
	void deinit() {}
	bool process(const vector<sample_t> &in, vector<sample_t> &out)
	{
		(void)in;
		if (!ring.empty()) {
			if (ring[head].size() != out.size())
				return false;
			out.swap(ring[head]);
			head = (head + 1) % ring.size();
			return true;
		}
		if (!cur || len != out.size())
			return false;
		memcpy(out.data(), cur, len * sizeof(float));
		return true;
	}
	const float *cur;
	size_t len;
	vector<vector<sample_t> > ring;
	size_t head;
};

ReplayTuner *g_lastTuner = NULL;
Tuner *replayFactory(const string &name) { return g_lastTuner = new ReplayTuner(name); }

struct Rig {
	FrontEnd *fe;
	ReplayTuner *tuner;
	std::vector<Receiver*> rx;
	unsigned frames;
};

struct Quiet {
	int saved;
	Quiet() { fflush(stderr); saved = dup(2); int n = open("/dev/null", O_WRONLY); if (n >= 0) { dup2(n, 2); close(n); } }
	~Quiet() { fflush(stderr); dup2(saved, 2); close(saved); }
};

} // namespace

extern "C" {

void *wrr_create(unsigned fs, unsigned block_frames, unsigned fft_size)
{
	Quiet q;
	Rig *r = new Rig();
	r->fe = new FrontEnd(replayFactory);
	r->tuner = g_lastTuner;
	r->frames = block_frames;
	r->fe->tuner()->setSampleRate(fs);
	r->fe->tuner()->setBlockSize(block_frames * 2);
	r->fe->spectrum()->setFftSize(fft_size);
	return r;
}

/* new Receiver() with the reference's defaults (80 kHz -> 240 k, 8 kHz -> 48 k, radio.cxx:78-82) */
int wrr_add_receiver(void *h, int if_hz, const char *mode)
{
	Quiet q;
	Rig *r = (Rig*)h;
	Receiver *rx = new Receiver();
	rx->setFrontEnd(r->fe);
	rx->downconverter()->setIF(if_hz);
	if (!rx->demodulator()->setModeString(mode))
		return -1;
	r->rx.push_back(rx);
	return (int)r->rx.size() - 1;
}

/* The same, at another geometry: tap counts and decimations are set through the drop-in blocks' own
 * additions (LowPass::setFirLength; the reference fixes 64 taps at compile time, lowpass.cxx:39)
 * before the pipeline starts.  0 keeps the reference default. */
int wrr_add_receiver_geo(void *h, int if_hz, const char *mode, unsigned n1, unsigned d1, unsigned n2, unsigned d2)
{
	Quiet q;
	Rig *r = (Rig*)h;
	Receiver *rx = new Receiver();
	if (n1) rx->channelFilter()->setFirLength(n1);
	if (d1) rx->channelFilter()->setDecimation(d1);
	if (n2) rx->audioFilter()->setFirLength(n2);
	if (d2) rx->audioFilter()->setDecimation(d2);
	rx->setFrontEnd(r->fe);
	rx->downconverter()->setIF(if_hz);
	if (!rx->demodulator()->setModeString(mode))
		return -1;
	r->rx.push_back(rx);
	return (int)r->rx.size() - 1;
}

/* injected coefficients (stage 0 = channel filter, 1 = audio filter), on a started pipeline */
int wrr_set_taps(void *h, int rx, int stage, const float *taps, unsigned n)
{
	Rig *r = (Rig*)h;
	LowPass *lp = stage ? r->rx[rx]->audioFilter() : r->rx[rx]->channelFilter();
	return lp->setCoefficients(taps, n) ? 0 : -1;
}

int wrr_start(void *h)
{
	Quiet q;
	return ((Rig*)h)->fe->tuner()->start() ? 0 : -1;
}

/* one Radio::run() (radio.cxx:56-59) over a caller-supplied tuner block */
int wrr_run(void *h, const float *iq)
{
	Rig *r = (Rig*)h;
	r->tuner->feed(iq, (size_t)r->frames * 2);
	Radio::run();
	return 0;
}

/* `steps` calls of Radio::run() over a rotating set of caller-supplied tuner blocks (bench.py's
 * plug-in leg: the whole loop stays on the C++ side, as the reference's main() has it) */
int wrr_run_steps(void *h, const float *const *iq, unsigned n_iq, unsigned first, unsigned steps)
{
	Rig *r = (Rig*)h;
	for (unsigned i = 0; i < steps; i++) {
		r->tuner->feed(iq[(first + i) % n_iq], (size_t)r->frames * 2);
		Radio::run();
	}
	return 0;
}

/* Several front-ends at once, as Radio::run() sees them: every rig's tuner is handed its block, then
 * ONE Radio::run() visits all front-ends in turn (radio.cxx:56-59). */
int wrr_run_many(void *const *h, unsigned n, const float *const *iq)
{
	for (unsigned i = 0; i < n; i++) {
		Rig *r = (Rig*)h[i];
		r->tuner->feed(iq[i], (size_t)r->frames * 2);
	}
	Radio::run();
	return 0;
}

/* the CUDA device the drop-in blocks placed this front-end on (wrhost::deviceFor) */
int wrr_device(void *h)
{
	Rig *r = (Rig*)h;
	return wrhost::deviceFor(r->tuner);
}

/* the tuner's ring of blocks (see ReplayTuner::ringFill); wrr_run_ring runs `steps` Radio::run() over it */
int wrr_ring_fill(void *h, const float *iq)
{
	Rig *r = (Rig*)h;
	r->tuner->ringFill(iq, (size_t)r->frames * 2);
	return 0;
}

int wrr_run_ring(void *h, unsigned steps)
{
	Rig *r = (Rig*)h;
	if (!r->tuner->ringMode())
		return -1;
	for (unsigned i = 0; i < steps; i++)
		Radio::run();
	return 0;
}

/* AudioStreamManager stand-in: keep the last audio block of every receiver (tests) or return at once
 * as the reference does without a client (timed runs) */
void wrr_capture(int on) { AudioStreamManager::capture() = on != 0; }

/* what a web handler would do (receiverhandler.cxx:125-140) */
int wrr_retune(void *h, int rx, int if_hz, const char *mode, unsigned chan_passband)
{
	Rig *r = (Rig*)h;
	r->rx[rx]->downconverter()->setIF(if_hz);
	if (chan_passband)
		r->rx[rx]->channelFilter()->setPassband(chan_passband);
	return r->rx[rx]->demodulator()->setModeString(mode) ? 0 : -1;
}

long wrr_audio(void *h, int rx, float *out, long cap)
{
	Rig *r = (Rig*)h;
	const vector<float> &v = r->rx[rx]->stream()->last;
	long n = (long)v.size();
	if (out && cap > 0)
		memcpy(out, v.data(), sizeof(float) * (size_t)(n < cap ? n : cap));
	return n;
}

int wrr_spectrum(void *h, float *db)
{
	Rig *r = (Rig*)h;
	r->fe->spectrum()->getSpectrum(db);
	return (int)r->fe->spectrum()->fftSize();
}

unsigned wrr_counts(void *h, unsigned *n_frontends, unsigned *n_receivers)
{
	(void)h;
	*n_frontends = (unsigned)Radio::frontEnds().size();
	*n_receivers = (unsigned)Radio::receivers().size();
	return 0;
}

void wrr_destroy(void *h)
{
	Quiet q;
	Rig *r = (Rig*)h;
	r->fe->tuner()->stop();
	for (size_t i = 0; i < r->rx.size(); i++)
		delete r->rx[i];
	delete r->fe;
	delete r;
}

} // extern "C"
