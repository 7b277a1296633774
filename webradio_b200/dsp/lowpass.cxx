// lowpass.cxx -- LowPass block over libwebradio_b200.
// Behavioural contract: WebRadio src/dsp/lowpass.cxx:41-189.
#include "lowpass.h"

#include <algorithm>
#include <cstring>

#include "debug.h"
#include "gpubank.h"
#include "webradio_b200.h"

#define DEFAULT_FIR_LENGTH 64 /* the reference's compile-time FIR_LENGTH (lowpass.cxx:39) */

LowPass::LowPass(const string &name) :
	DspBlock(name, "LowPass"),
	_firLength(DEFAULT_FIR_LENGTH),
	_tapsVersion(1), stageTapsVersion(0),
	_passband(0),
	_reqDecimation(0), _reqOutputRate(DEFAULT_SAMPLE_RATE),
	stage(NULL), bank(NULL), bankSlot(-1), bankAudioRole(false)
{
}

LowPass::~LowPass()
{
	if (stage)
		wr_stage_destroy(stage);
}

// reference lowpass.cxx:55-61: a running filter is redesigned immediately (HTTP thread); the
// new taps are swapped in under a lock and reach the GPU at the next block boundary
void LowPass::setPassband(unsigned int hz)
{
	_passband = hz;
	if (isRunning())
		recalculate();
}

// reference lowpass.cxx:63-79: the two ways of asking for a rate are mutually exclusive and
// ignored while running
void LowPass::setDecimation(unsigned int n)
{
	if (isRunning())
		return;
	_reqDecimation = n;
	_reqOutputRate = 0;
}

void LowPass::setOutputSampleRate(unsigned int hz)
{
	if (isRunning())
		return;
	_reqOutputRate = hz;
	_reqDecimation = 0;
}

void LowPass::setFirLength(unsigned int ntaps)
{
	if (isRunning() || ntaps == 0)
		return;
	_firLength = ntaps;
}

// reference lowpass.cxx:81-116
bool LowPass::init()
{
	if (_reqOutputRate > 0) {
		_outputSampleRate = _reqOutputRate;
	} else if (_reqDecimation > 0) {
		_outputSampleRate = inputSampleRate() / _reqDecimation;
	} else {
		LOG_ERROR("Must specify either decimation or output rate\n");
		return false;
	}
	_outputChannels = inputChannels();
	if (inputChannels() != 1 && inputChannels() != 2) {
		LOG_ERROR("LowPass: %u channels not supported\n", inputChannels());
		return false;
	}
	recalculate();
	return true;
}

// reference lowpass.cxx:118-129.  The history goes HERE, with `block` -- not in init(): a block
// that is started twice without a stop in between (DspBlock::connect of an already connected
// consumer on a live pipeline, dspblock.cxx:59-62) keeps filtering where it was.
void LowPass::deinit()
{
	if (stage)
		wr_stage_fir_reset(stage);
	std::lock_guard<std::mutex> lk(tapsLock);
	vector<float>().swap(coeff);
}

// reference lowpass.cxx:164-189, evaluated by wr_lowpass_design (host, cold path)
void LowPass::recalculate()
{
	vector<float> fresh(_firLength);
	if (wr_lowpass_design(_firLength, _passband, inputSampleRate(), fresh.data()) != WR_OK) {
		LOG_ERROR("LowPass: %s\n", wr_last_error());
		return;
	}
	std::lock_guard<std::mutex> lk(tapsLock);
	coeff.swap(fresh);
	_tapsVersion++;
}

bool LowPass::setCoefficients(const float *c, unsigned int ntaps)
{
	if (!c || ntaps == 0)
		return false;
	std::lock_guard<std::mutex> lk(tapsLock);
	if (ntaps != _firLength && bank)
		return false; // the fused bank's geometry is fixed once it streams
	_firLength = ntaps;
	coeff.assign(c, c + ntaps);
	_tapsVersion++;
	return true;
}

unsigned int LowPass::coefficients(float *out, unsigned int cap) const
{
	std::lock_guard<std::mutex> lk(tapsLock);
	if (out)
		memcpy(out, coeff.data(), sizeof(float) * std::min<size_t>(cap, coeff.size()));
	return (unsigned int)coeff.size();
}

void LowPass::snapshotTaps(vector<float> &out) const
{
	std::lock_guard<std::mutex> lk(tapsLock);
	out = coeff;
}

void LowPass::attachBank(wrhost::FusedBank *b, int slot, bool audioRole)
{
	bank = b;
	bankSlot = slot;
	bankAudioRole = audioRole;
}

void LowPass::detachBank()
{
	bank = NULL;
	bankSlot = -1;
}

// reference lowpass.cxx:131-162
bool LowPass::process(const vector<sample_t> &inBuffer, vector<sample_t> &outBuffer)
{
	if (bank) {
		if (!bankAudioRole)
			return true; // channel filter of a fused chain: computed inside the bank kernel
		unsigned int n = 0;
		const float *a = bank->audio(bankSlot, &n);
		if (!a)
			return false;
		const size_t want = outBuffer.size();
		const uint64_t t0 = wrhost::profOn() ? wrhost::profNow() : 0;
		memcpy(outBuffer.data(), a, sizeof(float) * std::min<size_t>(want, n));
		if (want > n)
			std::fill(outBuffer.begin() + n, outBuffer.end(), 0.0f);
		if (t0)
			wrhost::profAdd(wrhost::kProfAudioCopy, wrhost::profNow() - t0);
		return true;
	}

	// strict path: this block alone
	if (!stage) {
		stage = wr_stage_create(wrhost::defaultDevice());
		if (!stage) {
			LOG_ERROR("LowPass: %s\n", wr_last_error());
			return false;
		}
		stageTapsVersion = 0;
	}
	if (stageTapsVersion != _tapsVersion) {
		vector<float> taps;
		uint64_t v;
		{
			std::lock_guard<std::mutex> lk(tapsLock);
			taps = coeff;
			v = _tapsVersion;
		}
		if (taps.empty() || wr_stage_fir_config(stage, inputChannels(), taps.data(), (unsigned)taps.size()) != WR_OK) {
			LOG_ERROR("LowPass: %s\n", wr_last_error());
			return false;
		}
		stageTapsVersion = v;
	}
	const unsigned int nframes = (unsigned int)(inBuffer.size() / inputChannels());
	if (wr_stage_fir(stage, inBuffer.data(), nframes, DspBlock::decimation(), outBuffer.data()) != WR_OK) {
		LOG_ERROR("LowPass: %s\n", wr_last_error());
		return false;
	}
	return true;
}
