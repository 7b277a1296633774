// gpubank.h -- batches whole receiver chains behind the per-block DspBlock::process() calls.
//
// The Radio scheduler visits receivers one after another (reference src/radio.cxx:56-59), and
// every process() must return host-visible data.  To keep the GPU busy anyway, all receiver
// chains hanging off the same producer block (DownConverter -> LowPass(IQ) -> Demodulator ->
// LowPass(audio), with nothing else attached in between) are registered in one wr_bank; the first
// DownConverter::process() that sees a NEW producer buffer runs the fused kernels for ALL of
// them, and the audio LowPass of each chain then just copies its slice out (SURVEY.md 8b,
// "batching behind a per-receiver API").  Chains that cannot be fused fall back to the strict
// one-kernel-per-block stage calls (wr_stage_*), never to CPU arithmetic.
//
// Membership follows the graph: a receiver that joins or leaves a RUNNING front-end
// (Receiver::setFrontEnd, reference radio.cxx:109-117,151-163) marks the bank of that front-end
// and geometry dirty, and the bank is rebuilt at the next block boundary with exactly the chains
// that are alive -- those that stay keep their NCO phase, FM look-back sample and both FIR
// histories (wr_rx_get/set_*), a newcomer starts with empty histories as a freshly started
// LowPass does.  So a front-end's receivers are always ONE batched call per block, however they
// got there.
//
// The tuner block itself goes to the device once per DspBlock::run of its producer
// (wrhost::uploadFor): the SpectrumSink, which FrontEnd connects first (radio.cxx:126-128), and
// the bank read the same device copy.  Producers are dealt round-robin over the visible GPUs
// (wrhost::deviceFor): every receiver follows its tuner, no data crosses between devices.
#ifndef WEBRADIO_B200_GPUBANK_H
#define WEBRADIO_B200_GPUBANK_H

#include <stdint.h>

#include <vector>

#include "webradio_b200.h"

class DspBlock;
class DownConverter;
class LowPass;
class Demodulator;

namespace wrhost {

// CUDA device of everything fed by `producer` (NULL: of a chain without one).  Producers are
// dealt round-robin over the devices in use, at first sight; env WEBRADIO_B200_DEVICE pins
// everything to one device, WEBRADIO_B200_DEVICES=n uses the first n.
int deviceFor(const DspBlock *producer);
// the device of chains and sinks that have no producer
int defaultDevice();

// The device copy of the block `producer` is pushing (run serial `serial`): begun by the first
// consumer that asks, shared by the others.  `owner` keys the copy when there is no producer.
wr_upload *uploadFor(DspBlock *producer, const void *owner, uint64_t serial, const float *host, unsigned nframes);
// end of the producer's run(): its buffer may change from here on
void blockDone(DspBlock *producer);
// the producer stops or goes away: page-locking of its buffer is undone, the device copy freed
void forget(const void *producerOrOwner);

// WEBRADIO_B200_PROFILE=1: where the host thread spends a block (printed at exit; off: one branch per call)
enum ProfSlot { kProfUpload, kProfSettings, kProfBank, kProfAudioCopy, kProfSpectrum, kProfSlots };
bool profOn();
uint64_t profNow();
void profAdd(int slot, uint64_t ns);

struct Chain {
	DownConverter *dc;
	LowPass *chan;
	Demodulator *demod;
	LowPass *audio;
	// what was last pushed to the bank
	int32_t step;
	uint32_t phase0; // NCO phase the chain brings along when it joins
	float prev0[2];  // ... and the FM look-back sample (prev_i, prev_q)
	int mode;
	uint64_t chanTapsVersion, audioTapsVersion;
	bool active;
	int row;         // receiver index in the wr_bank, -1 until the bank has been (re)built with this chain
};

class FusedBank {
public:
	FusedBank(DspBlock *producer, unsigned n1, unsigned d1, unsigned n2, unsigned d2);
	~FusedBank();

	DspBlock *producer() const { return _producer; }
	bool matches(DspBlock *producer, unsigned n1, unsigned d1, unsigned n2, unsigned d2) const;
	int device() const { return _device; }

	// add a chain (any time; the bank is rebuilt at the next block); returns its member id
	int add(const Chain &c);
	// un-hook the chain's blocks from the bank (they fall back to the strict stage path)
	void detach(int member);
	bool empty() const;
	int memberOf(const DownConverter *dc) const;
	bool sameChain(int member, const Chain &c) const;

	// Runs the fused kernels for every chain if `serial` (the producer's run serial) is new.
	bool ensureProcessed(uint64_t serial, const float *iq, unsigned nframes, const void *owner);
	// audio of one chain for the block last processed
	const float *audio(int member, unsigned *nframes) const;
	uint32_t phaseOf(int member);

private:
	bool rebuild(unsigned nframes);
	void pushSettings();

	DspBlock *_producer;
	unsigned _n1, _d1, _n2, _d2;
	int _device;
	std::vector<Chain> _chains;
	wr_bank *_bank;
	unsigned _maxFrames;
	bool _dirty;
	uint64_t _lastSerial;
	bool _lastOk;
	float *_audio;   // page-locked, [row][stride]
	unsigned _audioStride, _audioFrames, _rows;
};

// Finds (or plans) the bank membership for a DownConverter; returns NULL if its chain is not fusable.
FusedBank *planFor(DownConverter *dc, int *member);
void release(FusedBank *bank, int member);

} // namespace wrhost

#endif
