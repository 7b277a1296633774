// gpubank.cxx -- see gpubank.h.
#include "gpubank.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "debug.h"
#include "demodulator.h"
#include "downconverter.h"
#include "lowpass.h"

namespace wrhost {

int defaultDevice()
{
	static int dev = -1;
	if (dev < 0) {
		const char *e = getenv("WEBRADIO_B200_DEVICE");
		dev = e ? atoi(e) : 0;
	}
	return dev;
}

namespace {

std::mutex g_lock;
std::vector<FusedBank*> g_banks;

// DownConverter -> LowPass(2 ch) -> Demodulator -> LowPass(1 ch), single consumer at each of
// the first three hops: nothing else may observe the intermediate streams.
bool fusable(DownConverter *dc, Chain *out)
{
	if (!dc->isRunning() || dc->downstream().size() != 1)
		return false;
	LowPass *chan = dynamic_cast<LowPass*>(dc->downstream()[0]);
	if (!chan || !chan->isRunning() || chan->inputChannels() != 2 || chan->downstream().size() != 1)
		return false;
	Demodulator *dm = dynamic_cast<Demodulator*>(chan->downstream()[0]);
	if (!dm || !dm->isRunning() || dm->downstream().size() != 1)
		return false;
	LowPass *audio = dynamic_cast<LowPass*>(dm->downstream()[0]);
	if (!audio || !audio->isRunning() || audio->inputChannels() != 1)
		return false;
	if (chan->firLength() == 0 || audio->firLength() == 0)
		return false;
	out->dc = dc;
	out->chan = chan;
	out->demod = dm;
	out->audio = audio;
	out->step = 0;
	out->phase0 = 0;
	out->prev0[0] = out->prev0[1] = 0.0f;
	out->mode = -1;
	out->chanTapsVersion = out->audioTapsVersion = 0;
	out->active = true;
	return true;
}

} // namespace

FusedBank::FusedBank(DspBlock *producer, unsigned n1, unsigned d1, unsigned n2, unsigned d2) :
	_producer(producer), _n1(n1), _d1(d1), _n2(n2), _d2(d2), _bank(NULL), _maxFrames(0),
	_lastSerial(0), _lastOk(false), _audioStride(0), _audioFrames(0)
{
}

FusedBank::~FusedBank()
{
	if (_bank)
		wr_bank_destroy(_bank);
}

bool FusedBank::matches(DspBlock *producer, unsigned n1, unsigned d1, unsigned n2, unsigned d2) const
{
	return producer == _producer && n1 == _n1 && d1 == _d1 && n2 == _n2 && d2 == _d2;
}

int FusedBank::add(const Chain &c)
{
	_chains.push_back(c);
	return (int)_chains.size() - 1;
}

void FusedBank::deactivate(int slot)
{
	if (slot >= 0 && slot < (int)_chains.size())
		_chains[slot].active = false;
}

bool FusedBank::empty() const
{
	for (size_t i = 0; i < _chains.size(); i++)
		if (_chains[i].active)
			return false;
	return true;
}

bool FusedBank::seal(unsigned nframes)
{
	_maxFrames = nframes;
	_bank = wr_bank_create(defaultDevice(), 1, (unsigned)_chains.size(), nframes, _n1, _d1, _n2, _d2);
	if (!_bank) {
		LOG_ERROR("receiver bank: %s\n", wr_last_error());
		return false;
	}
	for (size_t i = 0; i < _chains.size(); i++) {
		wr_rx_set_phase(_bank, (unsigned)i, _chains[i].phase0);
		wr_rx_set_lookback(_bank, (unsigned)i, _chains[i].prev0);
	}
	_audioStride = std::max(1u, nframes / _d1 / _d2);
	_audio.assign((size_t)_audioStride * _chains.size(), 0.0f);
	LOG_DEBUG("receiver bank: %u chains fused on device %d (taps %u/%u, decimation %u/%u)\n",
			(unsigned)_chains.size(), defaultDevice(), _n1, _n2, _d1, _d2);
	return true;
}

// Apply whatever the setters changed since the previous block (they run on other threads and
// only touch the blocks' own fields; this is the block boundary where the GPU sees them).
void FusedBank::pushSettings()
{
	std::vector<float> taps;
	for (size_t i = 0; i < _chains.size(); i++) {
		Chain &c = _chains[i];
		if (!c.active)
			continue;
		const int32_t step = c.dc->phaseStepNow();
		if (step != c.step || c.mode < 0) {
			wr_rx_set_phase_step(_bank, (unsigned)i, step);
			c.step = step;
		}
		const int mode = (int)c.demod->mode();
		if (mode != c.mode) {
			wr_rx_set_mode(_bank, (unsigned)i, mode);
			c.mode = mode;
		}
		if (c.chan->tapsVersion() != c.chanTapsVersion) {
			c.chanTapsVersion = c.chan->tapsVersion();
			c.chan->snapshotTaps(taps);
			if (taps.size() == _n1)
				wr_rx_set_taps(_bank, (unsigned)i, 0, taps.data(), _n1);
		}
		if (c.audio->tapsVersion() != c.audioTapsVersion) {
			c.audioTapsVersion = c.audio->tapsVersion();
			c.audio->snapshotTaps(taps);
			if (taps.size() == _n2)
				wr_rx_set_taps(_bank, (unsigned)i, 1, taps.data(), _n2);
		}
	}
}

bool FusedBank::ensureProcessed(uint64_t serial, const float *iq, unsigned nframes)
{
	if (_bank && serial == _lastSerial)
		return _lastOk;
	if (!_bank && !seal(nframes))
		return false;
	_lastSerial = serial;
	_lastOk = false;
	if (nframes > _maxFrames) {
		LOG_ERROR("receiver bank: block of %u frames exceeds the %u it was sized for\n", nframes, _maxFrames);
		return false;
	}
	pushSettings();
	if (wr_bank_process(_bank, iq, nframes, _audio.data(), _audioStride) != WR_OK) {
		LOG_ERROR("receiver bank: %s\n", wr_last_error());
		return false;
	}
	_audioFrames = nframes / _d1 / _d2;
	_lastOk = true;
	return true;
}

const float *FusedBank::audio(int slot, unsigned *nframes) const
{
	if (!_lastOk || slot < 0 || slot >= (int)_chains.size())
		return NULL;
	*nframes = _audioFrames;
	return _audio.data() + (size_t)slot * _audioStride;
}

uint32_t FusedBank::phaseOf(int slot)
{
	uint32_t p = 0;
	if (_bank && slot >= 0)
		wr_rx_get_phase(_bank, (unsigned)slot, &p);
	return p;
}

FusedBank *planFor(DownConverter *dc, int *slot)
{
	std::lock_guard<std::mutex> lk(g_lock);
	Chain c;
	if (!fusable(dc, &c))
		return NULL;
	// still the member it was (re-plan after an unrelated topology change)?
	if (FusedBank *cur = dc->currentBank()) {
		int s = cur->slotOf(dc);
		if (s >= 0 && cur->sameChain(s, c)) {
			*slot = s;
			return cur;
		}
	}
	DspBlock *producer = dc->upstream();
	const unsigned n1 = c.chan->firLength(), d1 = c.chan->DspBlock::decimation();
	const unsigned n2 = c.audio->firLength(), d2 = c.audio->DspBlock::decimation();
	FusedBank *target = NULL;
	for (size_t b = 0; b < g_banks.size(); b++)
		if (!g_banks[b]->sealed() && g_banks[b]->matches(producer, n1, d1, n2, d2))
			target = g_banks[b];
	if (!target) {
		target = new FusedBank(producer, n1, d1, n2, d2);
		g_banks.push_back(target);
	}
	// Enlist every sibling chain of the same producer and geometry now: the first receiver to
	// be run triggers the kernels for all of them, so they must be in the bank before it seals.
	std::vector<DspBlock*> sibs;
	if (producer)
		sibs = producer->downstream();
	else
		sibs.push_back(dc);
	*slot = -1;
	for (size_t i = 0; i < sibs.size(); i++) {
		DownConverter *d = dynamic_cast<DownConverter*>(sibs[i]);
		if (!d || (d != dc && d->currentBank()))
			continue;
		Chain sc;
		if (!fusable(d, &sc))
			continue;
		if (sc.chan->firLength() != n1 || sc.chan->DspBlock::decimation() != d1 ||
				sc.audio->firLength() != n2 || sc.audio->DspBlock::decimation() != d2)
			continue;
		sc.phase0 = d->phaseNow();
		sc.prev0[0] = sc.demod->lookback()[0];
		sc.prev0[1] = sc.demod->lookback()[1];
		int s = target->add(sc);
		sc.chan->attachBank(target, s, false);
		sc.demod->setFused(true);
		sc.audio->attachBank(target, s, true);
		if (d == dc)
			*slot = s;
		else
			d->adoptBank(target, s);
	}
	return target;
}

void release(FusedBank *bank, int slot)
{
	std::lock_guard<std::mutex> lk(g_lock);
	if (!bank)
		return;
	bank->detach(slot);
	if (bank->empty()) {
		g_banks.erase(std::remove(g_banks.begin(), g_banks.end(), bank), g_banks.end());
		delete bank;
	}
}

int FusedBank::slotOf(const DownConverter *dc) const
{
	for (size_t i = 0; i < _chains.size(); i++)
		if (_chains[i].active && _chains[i].dc == dc)
			return (int)i;
	return -1;
}

bool FusedBank::sameChain(int slot, const Chain &c) const
{
	const Chain &m = _chains[slot];
	return m.chan == c.chan && m.demod == c.demod && m.audio == c.audio;
}

void FusedBank::detach(int slot)
{
	if (slot < 0 || slot >= (int)_chains.size() || !_chains[slot].active)
		return;
	Chain &c = _chains[slot];
	c.chan->detachBank();
	// the discriminator's look-back sample goes back into the block it belongs to
	float prev[2];
	if (_bank && wr_rx_get_lookback(_bank, (unsigned)slot, prev) == WR_OK)
		c.demod->setLookback(prev);
	c.demod->setFused(false);
	c.audio->detachBank();
	c.active = false;
}

} // namespace wrhost
