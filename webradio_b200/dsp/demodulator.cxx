// demodulator.cxx -- Demodulator block over libwebradio_b200.
// Behavioural contract: WebRadio src/dsp/demodulator.cxx:31-115.
#include "demodulator.h"

#include "debug.h"
#include "gpubank.h"
#include "webradio_b200.h"

// The reference registers the block under the type string "AMDemod" (demodulator.cxx:32); the
// web UI shows it, so it is kept.
Demodulator::Demodulator(const string &name) :
	DspBlock(name, "AMDemod"),
	_mode(AM), stage(NULL), _fused(false)
{
	prev[0] = prev[1] = 0.0f;
	const char *names[] = { "AM", "FM", "USB", "LSB" }; // enum order
	for (int i = 0; i < MAX_MODE; i++)
		_modeStrings.push_back(names[i]);
}

Demodulator::~Demodulator()
{
	if (stage)
		wr_stage_destroy(stage);
}

// reference demodulator.cxx:47-56
bool Demodulator::setModeString(const string &mode)
{
	for (size_t i = 0; i < _modeStrings.size(); i++)
		if (_modeStrings[i] == mode) {
			setMode((Mode)i);
			return true;
		}
	return false;
}

// reference demodulator.cxx:58-68
bool Demodulator::init()
{
	if (inputChannels() != 2) {
		LOG_ERROR("Expect IQ input\n");
		return false;
	}
	_outputSampleRate = inputSampleRate();
	_outputChannels = 1;
	return true;
}

void Demodulator::deinit()
{
}

// reference demodulator.cxx:77-115
bool Demodulator::process(const vector<sample_t> &inBuffer, vector<sample_t> &outBuffer)
{
	if (_fused)
		return true; // demodulated inside the bank kernel's epilogue
	const Mode m = mode();
	if (m < AM || m >= MAX_MODE) {
		LOG_ERROR("Bad mode\n");
		return false;
	}
	if (!stage) {
		stage = wr_stage_create(wrhost::defaultDevice());
		if (!stage) {
			LOG_ERROR("Demodulator: %s\n", wr_last_error());
			return false;
		}
	}
	const unsigned int nframes = (unsigned int)(inBuffer.size() / inputChannels());
	if (wr_stage_demod(stage, (int)m, prev, inBuffer.data(), nframes, outBuffer.data()) != WR_OK) {
		LOG_ERROR("Demodulator: %s\n", wr_last_error());
		return false;
	}
	return true;
}
