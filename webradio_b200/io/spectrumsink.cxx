// spectrumsink.cxx -- SpectrumSink over libwebradio_b200.
// Behavioural contract: WebRadio src/io/spectrumsink.cxx:33-142.
//
// The reference transforms every complete frame of every block and keeps only the last one
// (spectrumsink.cxx:101-121); only that last transform is observable through getSpectrum(), so
// wr_spectrum_process is asked for no rows and computes just the newest frame per block.
#include "spectrumsink.h"

#include <cmath>

#include "debug.h"
#include "gpubank.h"
#include "webradio_b200.h"

SpectrumSink::SpectrumSink(const string &name) :
	SampleSink(name, "SpectrumSink"),
	_fftSize(DEFAULT_FFT_SIZE), spectrum(NULL), capacityFrames(0)
{
}

SpectrumSink::~SpectrumSink()
{
	if (spectrum)
		wr_spectrum_destroy(spectrum);
}

// reference spectrumsink.cxx:48-58: ignored while running; powers of two only
void SpectrumSink::setFftSize(unsigned int size)
{
	if (isRunning())
		return;
	if (size == 0 || (size & (size - 1))) {
		LOG_ERROR("size must be a power of 2\n");
		return;
	}
	_fftSize = size;
}

// reference spectrumsink.cxx:60-77.  The device handle is created on the first block, when
// the block length is known.
bool SpectrumSink::init()
{
	if (inputChannels() != 2) {
		LOG_ERROR("SpectrumSink expects IQ input\n");
		return false;
	}
	std::lock_guard<std::mutex> lk(lock);
	if (spectrum) {
		wr_spectrum_destroy(spectrum); // restart: inoffset = 0, no transform yet
		spectrum = NULL;
	}
	return true;
}

void SpectrumSink::deinit()
{
	std::lock_guard<std::mutex> lk(lock);
	if (spectrum) {
		wr_spectrum_destroy(spectrum);
		spectrum = NULL;
	}
}

// reference spectrumsink.cxx:88-123.  The block is read from the device copy its producer's
// consumers share (wrhost::uploadFor: the receiver bank of the same front-end uses it too), on the
// producer's device; the transform runs behind the copy and getSpectrum() waits for it.
bool SpectrumSink::process(const vector<sample_t> &inBuffer, vector<sample_t> &outBuffer)
{
	(void)outBuffer;
	const unsigned int nframes = (unsigned int)(inBuffer.size() / inputChannels());
	std::lock_guard<std::mutex> lk(lock);
	DspBlock *src = upstream();
	if (!spectrum) {
		capacityFrames = nframes > 0 ? nframes : 1;
		spectrum = wr_spectrum_create(wrhost::deviceFor(src), _fftSize, _fftSize, 1, capacityFrames);
		if (!spectrum) {
			LOG_ERROR("SpectrumSink: %s\n", wr_last_error());
			return false;
		}
	} else if (nframes > capacityFrames) {
		// a longer block than before: more room, the partial frame and the last spectrum stay
		if (wr_spectrum_reserve(spectrum, nframes) != WR_OK) {
			LOG_ERROR("SpectrumSink: %s\n", wr_last_error());
			return false;
		}
		capacityFrames = nframes;
	}
	if (src) {
		const uint64_t t0 = wrhost::profOn() ? wrhost::profNow() : 0;
		wr_upload *up = wrhost::uploadFor(src, this, src->runSerial(), inBuffer.data(), nframes);
		const bool ok = up && wr_spectrum_process_upload(spectrum, up, nframes) >= 0;
		if (t0)
			wrhost::profAdd(wrhost::kProfSpectrum, wrhost::profNow() - t0);
		if (ok)
			return true;
		LOG_ERROR("SpectrumSink: %s\n", wr_last_error());
		return false;
	}
	if (wr_spectrum_process(spectrum, inBuffer.data(), nframes, NULL, 0) < 0) {
		LOG_ERROR("SpectrumSink: %s\n", wr_last_error());
		return false;
	}
	return true;
}

// reference spectrumsink.cxx:125-142
void SpectrumSink::getSpectrum(float *magnitudes)
{
	std::lock_guard<std::mutex> lk(lock);
	if (spectrum && wr_spectrum_get(spectrum, 0, magnitudes) == WR_OK)
		return;
	// not started yet: the reference's output buffer is all zero -> 10*log10f(0) - 20*log10f(N)
	const float floorDb = 10 * log10f(0.0f) - 20 * log10f((float)_fftSize);
	for (unsigned int n = 0; n < _fftSize; n++)
		magnitudes[n] = floorDb;
}
