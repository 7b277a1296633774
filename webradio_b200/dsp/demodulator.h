// demodulator.h -- AM / FM / USB / LSB demodulator block, CUDA-backed.  Same public interface as
// WebRadio's src/dsp/demodulator.h:35-63.
#ifndef DEMODULATOR_H_
#define DEMODULATOR_H_

#include <atomic>
#include <string>
#include <vector>

#include "dspblock.h"

struct wr_stage;

using namespace std;

class Demodulator : public DspBlock
{
public:
	Demodulator(const string &name = "<undefined>");
	virtual ~Demodulator();

	enum Mode {
		AM,
		FM,
		USB,
		LSB,
		MAX_MODE
	};

	const Mode mode() const { return (Mode)_mode.load(std::memory_order_relaxed); }
	void setMode(const Mode mode) { _mode.store((int)mode, std::memory_order_relaxed); }
	const string &modeString() const { return _modeStrings[mode()]; }
	bool setModeString(const string &mode);

	// ---- fused-bank hand-off ----
	void setFused(bool fused) { _fused = fused; }
	// the FM look-back sample travels with the block: into a bank when the chain joins one, back
	// out when it leaves (the reference's prev_i/prev_q live in the object, demodulator.h:60-61)
	const float *lookback() const { return prev; }
	void setLookback(const float *iq) { prev[0] = iq[0]; prev[1] = iq[1]; }

private:
	bool init();
	void deinit();
	bool process(const vector<sample_t> &inBuffer, vector<sample_t> &outBuffer);

	// written by HTTP threads (setMode/setModeString), read by the DSP thread once per block
	std::atomic<int> _mode;
	vector<string> _modeStrings;
	float prev[2]; // prev_i, prev_q (reference demodulator.h:60-61)
	wr_stage *stage;
	bool _fused;
};

#endif /* DEMODULATOR_H_ */
