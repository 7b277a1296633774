// lowpass.h -- decimating low-pass FIR block, CUDA-backed.  Same public interface as WebRadio's
// src/dsp/lowpass.h:35-68, plus setCoefficients()/coefficients() for run-time tap injection
// (the reference fixes the length at 64 at compile time, lowpass.cxx:39).
#ifndef FILTER_H_
#define FILTER_H_

#include <atomic>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "dspblock.h"

struct wr_stage;
namespace wrhost { class FusedBank; }

using namespace std;

class LowPass : public DspBlock
{
public:
	LowPass(const string &name = "<undefined>");
	virtual ~LowPass();

	unsigned int passband() const { return _passband; }
	void setPassband(unsigned int hz);
	void setDecimation(unsigned int n);
	void setOutputSampleRate(unsigned int hz);

	// ---- additions ----
	// number of taps used by the next start() (default 64, the reference's FIR_LENGTH)
	void setFirLength(unsigned int ntaps);
	unsigned int firLength() const { return _firLength; }
	// replace the designed taps of a running filter (coeff[0] multiplies the newest sample)
	bool setCoefficients(const float *coeff, unsigned int ntaps);
	unsigned int coefficients(float *out, unsigned int cap) const;
	// bumped whenever the taps change; read by the fused bank at block boundaries
	uint64_t tapsVersion() const { return _tapsVersion; }
	void snapshotTaps(vector<float> &out) const;
	// fused-bank hand-off: while attached, process() copies (audio role) or skips (channel role)
	void attachBank(wrhost::FusedBank *b, int slot, bool audioRole);
	void detachBank();

private:
	bool init();
	void deinit();
	bool process(const vector<sample_t> &inBuffer, vector<sample_t> &outBuffer);
	void recalculate();

	unsigned int _firLength;
	vector<float> coeff;
	mutable std::mutex tapsLock;
	std::atomic<uint64_t> _tapsVersion;   // bumped under tapsLock, polled by the DSP thread without it
	uint64_t stageTapsVersion;
	std::atomic<unsigned int> _passband;

	unsigned int _reqDecimation;
	unsigned int _reqOutputRate;

	wr_stage *stage;
	wrhost::FusedBank *bank;
	int bankSlot;
	bool bankAudioRole;
};

#endif /* FILTER_H_ */
