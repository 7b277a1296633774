// wr_bank.cuh -- data layout of a receiver bank in HBM and the kernel argument blocks.
//
// HBM layout (R receivers, T tuner streams, geometry n1/d1/n2/d2, blocks of F frames):
//   table   float[65536]                 NCO sine table (read-only, L2 resident)
//   taps1   float[R][n1]  (REVERSED)     taps1[r][j] multiplies block sample k*d1 + j,
//   taps2   float[R][n2]  (REVERSED)       i.e. taps[j] = coeff[n-1-j] of the reference
//   conf    RxConf[R]                    phase step, mode, stream index
//   state   RxState[2][R]   ping-pong    NCO phase + previous channel sample (FM look-back)
//   hist1   float2[2][R][n1-1] ping-pong last n1-1 MIXED frames of the previous block
//   demod   float[2][R][n2-1 + M1max]    ping-pong; [0,n2-1) = audio-FIR history,
//                                        [n2-1, n2-1+M1) = this block's demodulated samples
//   iq      float2[T][F]                 the tuner block(s)
//   audio   float[R][M2]
// "ping-pong": a block reads side `cur` and writes side `cur^1`, because different CTAs of
// one launch read the old history while another writes the new one.
#pragma once

#include <stdint.h>
#include <cuda_runtime.h>

namespace wrd {

struct RxConf {
	int32_t step;     // phaseStep (reference downconverter.h:59)
	int32_t mode;     // Demodulator::Mode
	uint32_t stream;  // which tuner stream feeds this receiver
	uint32_t pad;
};

struct RxState {
	uint32_t phase;   // reference downconverter.h:58
	float prev_i;     // reference demodulator.h:60
	float prev_q;     // reference demodulator.h:61
	uint32_t pad;
};

struct ChanArgs {
	const float2 *iq;        // [T][stream_stride]
	size_t stream_stride;    // frames
	const float *table;
	const float *taps1;      // [R][n1] reversed
	const RxConf *conf;
	const RxState *st_in;
	RxState *st_out;
	const float2 *hist_in;   // [R][n1-1]
	float2 *hist_out;
	float *demod;            // side `cur`, row stride dstride, new samples at +demod_off
	size_t dstride;
	unsigned demod_off;      // n2-1
	float2 *chan;            // optional [R][chan_stride] channel-rate IQ (nullptr = dropped)
	size_t chan_stride;
	unsigned F, M1, n1, d1;
	unsigned TK, ntiles;
	// Pipelined host path (wr_bank_submit): the tuner block is being copied in by another stream.
	// in_flag counts the blocks that have landed; the kernel's loaders wait until it reaches in_seq.
	// nullptr = the block is already resident.  err (mapped host memory) receives WR_SYNC_TIMEOUT
	// if that never happens.
	const unsigned *in_flag;
	unsigned in_seq;
	unsigned *err;
	unsigned long long *ts;  // optional trace record of this block (WR_TRACE): kTs* device timestamps
	unsigned long long *cta_ts;  // optional per-CTA {start, end} of this block (WR_TRACE_CTA)
	unsigned poll_ns;        // pause between two looks at in_flag
	int wait_late;           // 1: the kernel waits for its predecessor (programmatic dependent launch) only before it exits
};

// trace record of one block: %globaltimer nanoseconds written by the kernels
enum { kTsChanStart = 0, kTsChanInput = 1, kTsChanEnd = 2, kTsDemodStart = 3, kTsDemodEnd = 4, kTsWords = 8 };

__device__ __forceinline__ unsigned long long global_ns()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

constexpr unsigned kSyncTimeout = 1u;   // bit in ChanArgs::err: the input flag was not raised within kSpinNs
constexpr unsigned long long kSpinNs = 2000000000ull;

struct AudioArgs {
	const float *x;          // demod side `cur`: [R][dstride] = history | new
	float *x_next;           // demod side `cur^1` (receives the new history)
	size_t dstride;
	const float *taps2;      // [R][n2] reversed
	float *audio;            // [R][audio_stride]
	size_t audio_stride;
	unsigned M1, M2, n2, d2;
	unsigned TK, ntiles;
	float out_scale;         // 1, or 32768 for the encoder's sample format (reference mp3encoder.cxx:66-73)
};

} // namespace wrd
