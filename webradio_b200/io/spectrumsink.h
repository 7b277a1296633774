// spectrumsink.h -- windowed FFT sink behind the waterfall display, CUDA-backed.  Same public
// interface as WebRadio's src/io/spectrumsink.h:44-70.
#ifndef SPECTRUMSINK_H_
#define SPECTRUMSINK_H_

#include <mutex>
#include <string>
#include <vector>

#include "samplesink.h"

#define DEFAULT_FFT_SIZE 512

struct wr_spectrum;

using namespace std;

class SpectrumSink : public SampleSink
{
public:
	SpectrumSink(const string &name = "<undefined>");
	virtual ~SpectrumSink();

	unsigned int fftSize() const { return _fftSize; }
	void setFftSize(unsigned int size);

	// dB magnitudes of the most recent transform, ascending frequency (fft-shifted);
	// `magnitudes` holds fftSize() floats.  Callable from any thread.
	void getSpectrum(float *magnitudes);

private:
	bool init();
	void deinit();
	bool process(const vector<sample_t> &inBuffer, vector<sample_t> &outBuffer);

	unsigned int _fftSize;
	wr_spectrum *spectrum;
	unsigned int capacityFrames;
	std::mutex lock; // guards `spectrum` against getSpectrum() from HTTP threads
};

#endif /* SPECTRUMSINK_H_ */
