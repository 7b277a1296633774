// downconverter.h -- NCO + complex mixer block, CUDA-backed.  Same public interface as
// WebRadio's src/dsp/downconverter.h:34-60.
#ifndef DOWNCONVERTER_H_
#define DOWNCONVERTER_H_

#include <atomic>
#include <stdint.h>

#include <string>
#include <vector>

#include "dspblock.h"
#include "lowpass.h"

struct wr_stage;
namespace wrhost { class FusedBank; }

using namespace std;

class DownConverter : public DspBlock
{
public:
	DownConverter(const string &name = "<undefined>");
	virtual ~DownConverter();

	// The reference keeps an inner LowPass that is never connected or run; these four
	// accessors only forward to it (reference downconverter.h:41-44, downconverter.cxx:44).
	unsigned int bandwidth() const { return filter->passband(); }
	void setBandwidth(unsigned int hz) { filter->setPassband(hz); }
	unsigned int decimation() const { return filter->decimation(); }
	void setDecimation(unsigned int n) { filter->setDecimation(n); }

	int IF() const { return _if; }
	void setIF(int hz);

	// ---- used by the fused receiver bank ----
	int32_t phaseStepNow() const { return phaseStep; }
	uint32_t phaseNow() const { return phase; }
	wrhost::FusedBank *currentBank() const { return bank; }
	void adoptBank(wrhost::FusedBank *b, int slot) { bank = b; bankSlot = slot; }

private:
	bool init();
	void deinit();
	bool process(const vector<sample_t> &inBuffer, vector<sample_t> &outBuffer);

	LowPass *filter;
	// _if and phaseStep are written by HTTP threads (setIF) and read by the DSP thread once per block
	std::atomic<int> _if;

	// NCO state (reference downconverter.h:57-59); the sine table itself lives in HBM
	uint32_t phase;
	std::atomic<int32_t> phaseStep;

	// execution back-ends
	wrhost::FusedBank *bank;
	int bankSlot;
	uint64_t plannedTopology;
	wr_stage *stage;
};

#endif /* DOWNCONVERTER_H_ */
