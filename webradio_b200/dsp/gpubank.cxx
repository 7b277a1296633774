// gpubank.cxx -- see gpubank.h.
#include "gpubank.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <ctime>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

#include "debug.h"
#include "demodulator.h"
#include "downconverter.h"
#include "lowpass.h"

namespace wrhost {

// ---- optional host-side profile of the plug-in path ----
namespace {
std::atomic<uint64_t> g_profNs[kProfSlots];
std::atomic<uint64_t> g_profN[kProfSlots];
void profDump()
{
	static const char *names[kProfSlots] = { "tuner block upload (begin)", "settings push", "bank process (kernels + audio copy-out)",
			"audio slice copies (LowPass)", "spectrum process" };
	for (int i = 0; i < kProfSlots; i++)
		if (g_profN[i])
			fprintf(stderr, "webradio_b200 profile: %-42s %10.1f us total  %9llu calls  %8.2f us per call\n", names[i],
					g_profNs[i] / 1e3, (unsigned long long)g_profN[i], g_profNs[i] / 1e3 / (double)g_profN[i]);
}
}
bool profOn()
{
	static const bool on = [] {
		const char *e = getenv("WEBRADIO_B200_PROFILE");
		const bool v = e && atoi(e) != 0;
		if (v)
			atexit(profDump);
		return v;
	}();
	return on;
}
uint64_t profNow()
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}
void profAdd(int slot, uint64_t ns)
{
	g_profNs[slot] += ns;
	g_profN[slot]++;
}

namespace {

std::mutex g_lock;
std::vector<FusedBank*> g_banks;

// ---- devices ------------------------------------------------------------------------------
// One entry per producer ever seen; a tuner keeps its device for life (its bank, its spectrum
// sink and its upload all live there).
std::map<const void*, int> g_deviceOf;
unsigned g_nextDevice = 0;

int g_devicesOverride = 0;   // test hook: wrhost_set_device_count_for_test

int devicesInUse()
{
	if (g_devicesOverride > 0)
		return g_devicesOverride;
	static int n = -1;
	if (n < 0) {
		n = wr_device_count();
		if (const char *e = getenv("WEBRADIO_B200_DEVICES"))
			n = std::min(n, std::max(1, atoi(e)));
		if (getenv("WEBRADIO_B200_DEVICE"))
			n = 1;
		if (n < 1)
			n = 1;   // no device at all: the create calls will say so (WR_ENODEV)
	}
	return n;
}

int firstDevice()
{
	static int dev = -1;
	if (dev < 0) {
		const char *e = getenv("WEBRADIO_B200_DEVICE");
		dev = e ? atoi(e) : 0;
	}
	return dev;
}

int deviceForLocked(const void *producer)
{
	if (!producer)
		return firstDevice();
	std::map<const void*, int>::iterator it = g_deviceOf.find(producer);
	if (it != g_deviceOf.end())
		return it->second;
	const int dev = firstDevice() + (int)(g_nextDevice++ % (unsigned)devicesInUse());
	g_deviceOf[producer] = dev;
	return dev;
}

// ---- uploads ------------------------------------------------------------------------------
struct TunerBlock {
	wr_upload *up;
	uint64_t serial;
	unsigned nframes;
	bool begun;
	TunerBlock() : up(NULL), serial(0), nframes(0), begun(false) {}
};
std::map<const void*, TunerBlock> g_uploads;

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
	out->row = -1;
	return true;
}

} // namespace

int deviceFor(const DspBlock *producer)
{
	std::lock_guard<std::mutex> lk(g_lock);
	return deviceForLocked(producer);
}

int defaultDevice() { return firstDevice(); }

wr_upload *uploadFor(DspBlock *producer, const void *owner, uint64_t serial, const float *host, unsigned nframes)
{
	std::lock_guard<std::mutex> lk(g_lock);
	const void *key = producer ? (const void*)producer : owner;
	TunerBlock &tb = g_uploads[key];
	if (tb.up && tb.begun && tb.serial == serial && tb.nframes == nframes)
		return tb.up;
	if (tb.up && wr_upload_capacity(tb.up) < nframes) {
		wr_upload_destroy(tb.up);   // a longer block than ever before: a larger device copy
		tb.up = NULL;
	}
	if (!tb.up) {
		tb.up = wr_upload_create(deviceForLocked(producer), nframes ? nframes : 1);
		if (!tb.up) {
			LOG_ERROR("tuner block upload: %s\n", wr_last_error());
			return NULL;
		}
	}
	if (wr_upload_begin(tb.up, host, nframes) != WR_OK) {
		LOG_ERROR("tuner block upload: %s\n", wr_last_error());
		tb.begun = false;
		return NULL;
	}
	tb.serial = serial;
	tb.nframes = nframes;
	tb.begun = true;
	if (producer)
		producer->noteUpload();
	return tb.up;
}

void blockDone(DspBlock *producer)
{
	std::lock_guard<std::mutex> lk(g_lock);
	std::map<const void*, TunerBlock>::iterator it = g_uploads.find(producer);
	if (it != g_uploads.end() && it->second.up && it->second.begun)
		wr_upload_finish(it->second.up);
}

void forget(const void *key)
{
	std::lock_guard<std::mutex> lk(g_lock);
	std::map<const void*, TunerBlock>::iterator it = g_uploads.find(key);
	if (it == g_uploads.end())
		return;
	if (it->second.up)
		wr_upload_destroy(it->second.up);
	g_uploads.erase(it);
}

FusedBank::FusedBank(DspBlock *producer, unsigned n1, unsigned d1, unsigned n2, unsigned d2) :
	_producer(producer), _n1(n1), _d1(d1), _n2(n2), _d2(d2), _device(deviceForLocked(producer)), _bank(NULL),
	_maxFrames(0), _dirty(true), _lastSerial(0), _lastOk(false), _audio(NULL), _audioStride(0), _audioFrames(0), _rows(0)
{
}

FusedBank::~FusedBank()
{
	if (_bank)
		wr_bank_destroy(_bank);
	if (_audio)
		wr_host_free(_audio);
}

bool FusedBank::matches(DspBlock *producer, unsigned n1, unsigned d1, unsigned n2, unsigned d2) const
{
	return producer == _producer && n1 == _n1 && d1 == _d1 && n2 == _n2 && d2 == _d2;
}

int FusedBank::add(const Chain &c)
{
	_dirty = true;
	for (size_t i = 0; i < _chains.size(); i++)
		if (!_chains[i].active) {
			_chains[i] = c;   // a member that left earlier: its place is free
			return (int)i;
		}
	_chains.push_back(c);
	return (int)_chains.size() - 1;
}

bool FusedBank::empty() const
{
	for (size_t i = 0; i < _chains.size(); i++)
		if (_chains[i].active)
			return false;
	return true;
}

// (Re)build the wr_bank for exactly the chains that are alive.  Chains that were in the old bank
// keep everything a reference Receiver keeps between blocks -- NCO phase (downconverter.h:58), FM
// look-back sample (demodulator.h:60-61), both FIR histories (lowpass.h:64); a chain that joins
// brings its phase and look-back sample (they live in the block objects across stop()/start()) and
// starts with empty histories (lowpass.cxx:118-129 releases them on deinit).
bool FusedBank::rebuild(unsigned nframes)
{
	const unsigned maxFrames = std::max(nframes, _maxFrames);
	unsigned rows = 0;
	for (size_t i = 0; i < _chains.size(); i++)
		if (_chains[i].active)
			rows++;
	if (rows == 0)
		return false;
	wr_bank *nb = wr_bank_create(_device, 1, rows, maxFrames, _n1, _d1, _n2, _d2);
	if (!nb) {
		LOG_ERROR("receiver bank: %s\n", wr_last_error());
		return false;
	}
	std::vector<float> h1(2 * (size_t)(_n1 - 1) + 1), h2((size_t)(_n2 - 1) + 1);
	unsigned r = 0;
	bool ok = true;
	for (size_t i = 0; i < _chains.size(); i++) {
		Chain &c = _chains[i];
		if (!c.active)
			continue;
		if (_bank && c.row >= 0) {
			uint32_t phase = 0;
			float prev[2] = { 0.0f, 0.0f };
			ok = ok && wr_rx_get_phase(_bank, (unsigned)c.row, &phase) == WR_OK
					&& wr_rx_get_lookback(_bank, (unsigned)c.row, prev) == WR_OK
					&& wr_rx_get_history(_bank, (unsigned)c.row, 0, h1.data(), 2 * (_n1 - 1)) == WR_OK
					&& wr_rx_get_history(_bank, (unsigned)c.row, 1, h2.data(), _n2 - 1) == WR_OK;
			ok = ok && wr_rx_set_phase(nb, r, phase) == WR_OK && wr_rx_set_lookback(nb, r, prev) == WR_OK
					&& wr_rx_set_history(nb, r, 0, h1.data(), 2 * (_n1 - 1)) == WR_OK
					&& wr_rx_set_history(nb, r, 1, h2.data(), _n2 - 1) == WR_OK;
		} else {
			ok = ok && wr_rx_set_phase(nb, r, c.phase0) == WR_OK && wr_rx_set_lookback(nb, r, c.prev0) == WR_OK;
		}
		c.row = (int)r++;
		// everything else is pushed again by pushSettings()
		c.mode = -1;
		c.chanTapsVersion = c.audioTapsVersion = 0;
	}
	if (!ok) {
		LOG_ERROR("receiver bank: state hand-over failed: %s\n", wr_last_error());
		wr_bank_destroy(nb);
		for (size_t i = 0; i < _chains.size(); i++)
			_chains[i].row = -1;
		if (_bank)
			wr_bank_destroy(_bank);
		_bank = NULL;
		return false;
	}
	if (_bank)
		wr_bank_destroy(_bank);
	_bank = nb;
	_rows = rows;
	_maxFrames = maxFrames;
	const unsigned stride = std::max(1u, maxFrames / _d1 / _d2);
	if (_audio)
		wr_host_free(_audio);
	_audio = static_cast<float*>(wr_host_alloc(sizeof(float) * (size_t)stride * rows));
	if (!_audio) {
		LOG_ERROR("receiver bank: %s\n", wr_last_error());
		return false;
	}
	memset(_audio, 0, sizeof(float) * (size_t)stride * rows);
	_audioStride = stride;
	_dirty = false;
	LOG_DEBUG("receiver bank: %u chains fused on device %d (taps %u/%u, decimation %u/%u)\n",
			rows, _device, _n1, _n2, _d1, _d2);
	return true;
}

// Apply whatever the setters changed since the previous block (they run on other threads and
// only touch the blocks' own fields; this is the block boundary where the GPU sees them).
void FusedBank::pushSettings()
{
	std::vector<float> taps;
	for (size_t i = 0; i < _chains.size(); i++) {
		Chain &c = _chains[i];
		if (!c.active || c.row < 0)
			continue;
		const unsigned row = (unsigned)c.row;
		const int32_t step = c.dc->phaseStepNow();
		if (step != c.step || c.mode < 0) {
			wr_rx_set_phase_step(_bank, row, step);
			c.step = step;
		}
		const int mode = (int)c.demod->mode();
		if (mode != c.mode) {
			wr_rx_set_mode(_bank, row, mode);
			c.mode = mode;
		}
		if (c.chan->tapsVersion() != c.chanTapsVersion) {
			c.chanTapsVersion = c.chan->tapsVersion();
			c.chan->snapshotTaps(taps);
			if (taps.size() == _n1)
				wr_rx_set_taps(_bank, row, 0, taps.data(), _n1);
		}
		if (c.audio->tapsVersion() != c.audioTapsVersion) {
			c.audioTapsVersion = c.audio->tapsVersion();
			c.audio->snapshotTaps(taps);
			if (taps.size() == _n2)
				wr_rx_set_taps(_bank, row, 1, taps.data(), _n2);
		}
	}
}

bool FusedBank::ensureProcessed(uint64_t serial, const float *iq, unsigned nframes, const void *owner)
{
	if (_bank && serial == _lastSerial && (!_dirty || _lastOk))
		return _lastOk;   // (a chain that joined while this very block was being pushed waits for the next one)
	_lastSerial = serial;
	_lastOk = false;
	// membership changed, or a block longer than the bank was sized for: a new bank, state carried
	if ((!_bank || _dirty || nframes > _maxFrames) && !rebuild(nframes))
		return false;
	const bool prof = profOn();
	uint64_t t0 = prof ? profNow() : 0;
	wr_upload *up = uploadFor(_producer, owner, serial, iq, nframes);
	if (!up)
		return false;
	if (prof) { const uint64_t t = profNow(); profAdd(kProfUpload, t - t0); t0 = t; }
	pushSettings();
	if (prof) { const uint64_t t = profNow(); profAdd(kProfSettings, t - t0); t0 = t; }
	const int rc = wr_bank_process_upload(_bank, up, nframes, _audio, _audioStride);
	if (prof) profAdd(kProfBank, profNow() - t0);
	if (!_producer)
		wr_upload_finish(up);   // nobody will call blockDone for a chain without a producer
	if (rc != WR_OK) {
		LOG_ERROR("receiver bank: %s\n", wr_last_error());
		return false;
	}
	_audioFrames = nframes / _d1 / _d2;
	_lastOk = true;
	return true;
}

const float *FusedBank::audio(int member, unsigned *nframes) const
{
	if (!_lastOk || member < 0 || member >= (int)_chains.size() || _chains[member].row < 0)
		return NULL;
	*nframes = _audioFrames;
	return _audio + (size_t)_chains[member].row * _audioStride;
}

uint32_t FusedBank::phaseOf(int member)
{
	if (member < 0 || member >= (int)_chains.size())
		return 0;
	uint32_t p = _chains[member].phase0;
	if (_bank && _chains[member].row >= 0)
		wr_rx_get_phase(_bank, (unsigned)_chains[member].row, &p);
	return p;
}

FusedBank *planFor(DownConverter *dc, int *member)
{
	std::lock_guard<std::mutex> lk(g_lock);
	Chain c;
	if (!fusable(dc, &c))
		return NULL;
	DspBlock *producer = dc->upstream();
	const unsigned n1 = c.chan->firLength(), d1 = c.chan->DspBlock::decimation();
	const unsigned n2 = c.audio->firLength(), d2 = c.audio->DspBlock::decimation();
	FusedBank *target = NULL;
	*member = -1;
	// still the member it was (re-plan after a topology change somewhere)?
	if (FusedBank *cur = dc->currentBank()) {
		int s = cur->memberOf(dc);
		if (s >= 0 && cur->sameChain(s, c) && cur->matches(producer, n1, d1, n2, d2)) {
			*member = s;
			target = cur;
		}
	}
	if (!target)
		for (size_t b = 0; b < g_banks.size(); b++)
			if (g_banks[b]->matches(producer, n1, d1, n2, d2))
				target = g_banks[b];
	if (!target) {
		target = new FusedBank(producer, n1, d1, n2, d2);
		g_banks.push_back(target);
	}
	// Enlist every sibling chain of the same producer and geometry that is not in a bank yet --
	// also when this chain was a member already: the first receiver of a bank to run a block
	// triggers the kernels for all of them, so a receiver that joined the front-end since the last
	// block must be a member BEFORE that happens (or the others would run the block twice).
	std::vector<DspBlock*> sibs;
	if (producer)
		sibs = producer->downstream();
	else
		sibs.push_back(dc);
	for (size_t i = 0; i < sibs.size(); i++) {
		DownConverter *d = dynamic_cast<DownConverter*>(sibs[i]);
		if (!d || (d == dc && *member >= 0) || (d != dc && d->currentBank()))
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
			*member = s;
		else
			d->adoptBank(target, s);
	}
	return target;
}

void release(FusedBank *bank, int member)
{
	std::lock_guard<std::mutex> lk(g_lock);
	if (!bank)
		return;
	bank->detach(member);
	if (bank->empty()) {
		g_banks.erase(std::remove(g_banks.begin(), g_banks.end(), bank), g_banks.end());
		delete bank;
	}
}

int FusedBank::memberOf(const DownConverter *dc) const
{
	for (size_t i = 0; i < _chains.size(); i++)
		if (_chains[i].active && _chains[i].dc == dc)
			return (int)i;
	return -1;
}

bool FusedBank::sameChain(int member, const Chain &c) const
{
	const Chain &m = _chains[member];
	return m.chan == c.chan && m.demod == c.demod && m.audio == c.audio;
}

void FusedBank::detach(int member)
{
	if (member < 0 || member >= (int)_chains.size() || !_chains[member].active)
		return;
	Chain &c = _chains[member];
	c.chan->detachBank();
	// the discriminator's look-back sample goes back into the block it belongs to
	float prev[2];
	if (_bank && c.row >= 0 && wr_rx_get_lookback(_bank, (unsigned)c.row, prev) == WR_OK)
		c.demod->setLookback(prev);
	c.demod->setFused(false);
	c.audio->detachBank();
	c.active = false;
	_dirty = true;   // the next block runs without this chain
}

} // namespace wrhost

// Test hook (the CPU stand-in of the device entry points pretends to have several devices): deal
// producers over `n` devices from now on, whatever wr_device_count() says.  0 = back to normal.
extern "C" void wrhost_set_device_count_for_test(int n)
{
	std::lock_guard<std::mutex> lk(wrhost::g_lock);
	wrhost::g_devicesOverride = n;
	wrhost::g_deviceOf.clear();
	wrhost::g_nextDevice = 0;
}
