// wr_upload.cuh -- one tuner block in HBM, shared by every consumer of the producer that made it
// (the receiver bank and the spectrum sink of a front-end, reference src/radio.cxx:126-128,151-156).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>

struct wr_upload {
	static constexpr int kPieces = 4;
	int device = 0;
	size_t maxFrames = 0;
	cudaStream_t st = nullptr;
	float *d_iq[2] = { nullptr, nullptr };     // [maxFrames][2], two blocks alternate
	cudaEvent_t landed[2][kPieces] = {};       // piece j of side b is in HBM
	cudaEvent_t readDone[2] = { nullptr, nullptr };   // the last asynchronous reader of side b is done with it
	bool readPending[2] = { false, false };
	int cur = 0;                               // side of the block begun last
	unsigned nframes = 0, npieces = 0;
	unsigned pieceEnd[kPieces] = {};           // frames [0, pieceEnd[j]) are covered by pieces 0..j
	// host ranges page-locked in place: a tuner that swaps ring buffers into its output (reference
	// rtlsdrtuner.cxx:265-285) shows a handful of different buffers in turn
	static constexpr int kRegs = 16;
	struct Reg { const void *ptr; size_t bytes; unsigned long used; } reg[kRegs] = {};
	unsigned long regClock = 0;
	unsigned regFailures = 0;
	// buffers seen once: a buffer is page-locked only when it comes back (a transient caller buffer,
	// e.g. a test's array, is copied as it is -- registering it would cost more than the copy and
	// leave a registration behind that its owner knows nothing about)
	struct Seen { const void *ptr; size_t bytes; } seen[kRegs] = {};
	unsigned seenNext = 0;

	const float *dev() const { return d_iq[cur]; }
	// the event after which frames [0, upto) of the current block are in HBM
	cudaEvent_t ready(unsigned upto) const
	{
		for (unsigned j = 0; j < npieces; j++)
			if (pieceEnd[j] >= upto)
				return landed[cur][j];
		return landed[cur][npieces ? npieces - 1 : 0];
	}
};
