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

// CUDA device used by the drop-in blocks (env WEBRADIO_B200_DEVICE, default 0)
int defaultDevice();

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
};

class FusedBank {
public:
	FusedBank(DspBlock *producer, unsigned n1, unsigned d1, unsigned n2, unsigned d2);
	~FusedBank();

	DspBlock *producer() const { return _producer; }
	bool matches(DspBlock *producer, unsigned n1, unsigned d1, unsigned n2, unsigned d2) const;
	bool sealed() const { return _bank != NULL; }

	// before the first block: add a chain; returns its slot
	int add(const Chain &c);
	void deactivate(int slot);
	// un-hook the chain's blocks from the bank (they fall back to the strict stage path)
	void detach(int slot);
	bool empty() const;
	int slotOf(const DownConverter *dc) const;
	bool sameChain(int slot, const Chain &c) const;

	// Runs the fused kernels for every chain if `serial` (the producer's run serial) is new.
	bool ensureProcessed(uint64_t serial, const float *iq, unsigned nframes);
	// audio of one chain for the block last processed
	const float *audio(int slot, unsigned *nframes) const;
	uint32_t phaseOf(int slot);

private:
	bool seal(unsigned nframes);
	void pushSettings();

	DspBlock *_producer;
	unsigned _n1, _d1, _n2, _d2;
	std::vector<Chain> _chains;
	wr_bank *_bank;
	unsigned _maxFrames;
	uint64_t _lastSerial;
	bool _lastOk;
	std::vector<float> _audio; // [slot][stride]
	unsigned _audioStride, _audioFrames;
};

// Finds (or plans) the bank slot for a DownConverter; returns NULL if its chain is not fusable.
FusedBank *planFor(DownConverter *dc, int *slot);
void release(FusedBank *bank, int slot);

} // namespace wrhost

#endif
