// wr_bank.cu -- the receiver bank: host-side state machine + kernel launches behind the
// wr_bank_* / wr_rx_* entry points of include/webradio_b200.h.
#include "wr_common.h"
#include "wr_bank.cuh"
#include "wr_kernels_v1.cuh"
#include "wr_kernels_v2.cuh"
#include "wr_kernels_v3.cuh"
#include "wr_kernels_v4.cuh"
#include "wr_upload.cuh"

#include <cuda.h>        // types of the two stream memory operations; the entry points are looked up at run time

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <cstdio>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

using wrd::RxConf;
using wrd::RxState;

namespace {

constexpr int kSlots = 6;        // most blocks the submit/wait path keeps in flight
constexpr int kThreadsV1 = 128;

constexpr unsigned kSetPhase = 0x100u; // internal flag: overwrite the phase with a given value
constexpr unsigned kSetPrev = 0x200u;  // internal flag: overwrite the FM look-back sample with a given value

__global__ void reset_kernel(RxState *st, float2 *hist1, float *demod_hist, unsigned rx,
		unsigned n1m1, unsigned n2m1, size_t dstride, unsigned flags, uint32_t phaseValue, float2 prevValue)
{
	const unsigned tid = threadIdx.x;
	if (tid == 0) {
		if (flags & WR_RESET_PHASE)
			st[rx].phase = 0;
		if (flags & kSetPhase)
			st[rx].phase = phaseValue & 0x7FFFFFFFu;
		if (flags & WR_RESET_DEMOD) {
			st[rx].prev_i = 0.0f;
			st[rx].prev_q = 0.0f;
		}
		if (flags & kSetPrev) {
			st[rx].prev_i = prevValue.x;
			st[rx].prev_q = prevValue.y;
		}
	}
	if (flags & WR_RESET_CHANNEL)
		for (unsigned i = tid; i < n1m1; i += blockDim.x)
			hist1[(size_t)rx * n1m1 + i] = make_float2(0.0f, 0.0f);
	if (flags & WR_RESET_AUDIO)
		for (unsigned i = tid; i < n2m1; i += blockDim.x)
			demod_hist[(size_t)rx * dstride + i] = 0.0f;
}

// RtlSdrTuner's sample conversion (reference src/io/rtlsdrtuner.cxx:106) as a kernel of its own:
// only for geometries the v3 channel kernel (which converts in its load path) does not serve.
__global__ void u8_to_f32_kernel(const unsigned char *__restrict__ in, float *__restrict__ out,
		size_t in_stride, size_t out_stride, unsigned nfloats)
{
	const unsigned char *src = in + (size_t)blockIdx.y * in_stride;
	float *dst = out + (size_t)blockIdx.y * out_stride;
	for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < nfloats; i += gridDim.x * blockDim.x)
		dst[i] = __fdiv_rn(__fsub_rn((float)src[i], 128.0f), 128.0f);
}

// LowPass::recalculate for every receiver of a bank at once (reference lowpass.cxx:164-189; the
// arithmetic is wr_lowpass_design's, term for term: same mask, same summation order in double,
// one rounding to float, same window).  cos(2 pi idx / n) and the window come from the HOST libm
// in two small tables, so host and device designs are bit-identical.
__global__ void design_kernel(const double *__restrict__ costab, const float *__restrict__ window,
		const unsigned *__restrict__ passband, unsigned fs, unsigned n, float *__restrict__ taps_rev)
{
	const unsigned r = blockIdx.x;
	const unsigned maxbin = n * passband[r] / fs / 2;          // lowpass.cxx:167, unsigned arithmetic
	for (unsigned k = threadIdx.x; k < n; k += blockDim.x) {
		const unsigned bin = (k + n / 2) % n;                  // lowpass.cxx:184
		double re = 0.0;
		for (unsigned m = 0; m < n; m++) {
			const double mask = (min(m, n - m) < maxbin) ? 1.0 : 0.0;
			const unsigned idx = (unsigned)(((unsigned long long)bin * m) % n);
			re = __dadd_rn(re, __dmul_rn(mask, costab[idx]));
		}
		taps_rev[(size_t)r * n + (n - 1 - k)] = __fmul_rn((float)re, window[k]);
	}
}

// Stream memory operations (cuStreamWriteValue32 / cuStreamWaitValue32), fetched through the
// runtime so that the library does not link against libcuda: the copy streams of the pipelined
// host path signal and wait through two words in HBM instead of events on the launch stream.
typedef CUresult (*StreamValueFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

struct StreamOps {
	StreamValueFn write = nullptr, wait = nullptr;
	bool ok = false;
};

const StreamOps &stream_ops()
{
	static const StreamOps ops = [] {
		StreamOps o;
		const char *e = getenv("WR_FLAG_SYNC");
		if (e && atoi(e) == 0)
			return o;
		cudaDriverEntryPointQueryResult q1, q2;
		void *w = nullptr, *a = nullptr;
		if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &w, cudaEnableDefault, &q1) == cudaSuccess
				&& cudaGetDriverEntryPoint("cuStreamWaitValue32", &a, cudaEnableDefault, &q2) == cudaSuccess
				&& q1 == cudaDriverEntryPointSuccess && q2 == cudaDriverEntryPointSuccess && w && a) {
			o.write = reinterpret_cast<StreamValueFn>(w);
			o.wait = reinterpret_cast<StreamValueFn>(a);
			o.ok = true;
		}
		cudaGetLastError();
		return o;
	}();
	return ops;
}

// How a block is handed from the copy-in stream to the channel kernel and from the demodulator
// kernel to the host.
enum { IN_EVENT = 0, IN_FLAG = 1, IN_COPYFLAG = 2 };     // event | stream write-value | a 4-byte copy behind the block
enum { OUT_EVENT = 0, OUT_FLAG = 1, OUT_DIRECT = 2 };    // event + copy | stream wait-value + copy | the kernel stores into the pinned buffer

struct Slot {
	float *d_iq = nullptr;      // [T][maxF][2] (float blocks) or the same bytes holding raw u8 blocks
	float *d_audio = nullptr;   // [R][maxM2]
	cudaEvent_t in_ready = nullptr, done = nullptr, out_ready = nullptr;
	bool busy = false;
	bool used = false;
	unsigned seq = 0;
	int outMode = OUT_EVENT;
};

inline unsigned long long host_ns()
{
	timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}

constexpr unsigned kTraceBlocks = 4096;
// WR_TRACE_CTA=<file>: {start, end} of every CTA of both kernels for a few blocks
constexpr unsigned kCtaTraceBlocks = 6, kCtaTraceChan = 256, kCtaTraceDemod = 4096;
constexpr int kSeqRing = 64;

} // namespace

struct wr_bank {
	int device = 0;
	int numSMs = 148;
	unsigned T = 0, R = 0, maxF = 0, n1 = 0, d1 = 0, n2 = 0, d2 = 0;
	unsigned pitchF = 0;        // frames between the streams of the bank's own device blocks: maxF rounded up to even, so
	                            // that every stream starts on the copies' boundary of the streaming channel kernel
	unsigned maxM1 = 0, maxM2 = 0;
	size_t dstride = 0;
	cudaStream_t compute = nullptr, h2d = nullptr, d2h = nullptr;

	float *d_table = nullptr;
	float *d_taps1 = nullptr, *d_taps2 = nullptr;
	RxConf *d_conf = nullptr;
	RxState *d_state[2] = { nullptr, nullptr };
	float2 *d_hist1[2] = { nullptr, nullptr };
	float *d_demod[2] = { nullptr, nullptr };
	float2 *d_chan = nullptr;   // [2][R][maxM1]: consecutive blocks alternate, so that a block's channel kernel can run under the previous block's demodulator
	int chanSide = 0;           // side the last block wrote
	float *d_iqf = nullptr;     // scratch: a u8 tuner block converted for the v1/v2 kernels
	int cur = 0;
	unsigned lastM1 = 0, lastM2 = 0;
	bool keepChan = false;
	float outScale = 1.0f;      // audio sample format: 1 = DspBlock floats, 32768 = what the MP3 encoder feeds LAME

	// host shadows of the per-receiver configuration; setters touch only these (any thread)
	std::mutex mu;
	std::vector<RxConf> h_conf;
	std::vector<float> h_taps1, h_taps2;   // reversed, as the kernels read them
	std::vector<unsigned> h_reset;         // pending WR_RESET_* per receiver
	std::vector<uint32_t> h_phase;         // pending wr_rx_set_phase values
	std::vector<float2> h_prev;            // pending wr_rx_set_lookback values
	bool confDirty = true, taps1Dirty = true, taps2Dirty = true, resetDirty = false;
	bool tableDirty = false;
	bool streamsDirty = true;              // receiver -> stream map changed (v2 regroups)
	std::vector<float> h_table;
	// pinned staging for the uploads
	RxConf *p_conf = nullptr;
	float *p_taps1 = nullptr, *p_taps2 = nullptr, *p_table = nullptr;
	cudaEvent_t stagingFree = nullptr;

	Slot slot[kSlots];
	int depth = kSlots;         // slots in use: as many as fit a few GB of HBM
	int head = 0, tail = 0, inflight = 0;
	// hand-over of the pipelined host path: d_sync = {blocks copied in, blocks demodulated, CTA count}
	unsigned *d_sync = nullptr;
	unsigned *h_err = nullptr;  // mapped host words: [0] a missed hand-over reported by the kernels, [1] blocks whose audio the demodulator kernel stored into host memory
	unsigned *p_seq = nullptr;  // pinned ring of sequence numbers (source of the IN_COPYFLAG copies)
	unsigned seq = 0;           // blocks submitted through wr_bank_submit
	int handIn = IN_EVENT, handOut = OUT_EVENT;
	unsigned pollNs = 1000;     // WR_POLL_NS: pause between two looks of the channel kernel at the copy-in counter
	int demodRegs = 0;
	bool anyFM = true;          // any receiver in FM mode (as of the last configuration upload)
	int demodPerSM = -1;        // WR_DEMOD_PER_SM: > 0 = persistent demodulator grid of that many CTAs per SM, 0 = one CTA per tile, < 0 = by bank size
	unsigned syncSplit = 0;     // WR_SYNC_SPLIT: pieces a synchronous wr_bank_process call is cut into (0 = by size)
	int waitLate = 1;           // WR_WAIT_LATE: the channel kernel runs under the previous block's demodulator kernel
	bool warnedSlow = false;    // the one-line notice about an un-instantiated geometry has been printed
	bool bigTiles = true;       // WR_DEMOD_BIG_TILES=0: never the 1024-output tiles of the sliding-window audio FIR
	// the device address of a pinned output buffer (OUT_DIRECT): small cache of the runtime's answer
	struct HostMap { const void *host; void *dev; size_t bytes; } hostMap[8] = {};
	int hostMapNext = 0;
	// WR_TRACE=<file>: per-block device and host timestamps, dumped when the bank is destroyed
	const char *tracePath = nullptr;
	unsigned long long *d_ts = nullptr;
	unsigned long long *d_cta = nullptr;
	const char *ctaPath = nullptr;
	unsigned ctaFirst = 300;    // WR_TRACE_CTA_FIRST: first traced block
	std::vector<unsigned long long> hostTs;   // {submit entered, submit returned, wait returned} per block

	int variant = 0;
	int variantInUse = 0;
	wrd::V2Plan v2;
	wrd::V3Plan v3;
	wrd::V4Plan v4;
	unsigned long long launches = 0;
	// optional per-launch device timing: a ring of event triples drained into accumulators
	bool timing = false;
	static constexpr int kTimeRing = 128;
	cudaEvent_t tev[kTimeRing][3] = {};
	int tHead = 0, tCount = 0;
	double accMs[2] = { 0.0, 0.0 };
	unsigned long long accBlocks = 0;
};

namespace {

// fold the oldest timed block into the accumulators (blocks until its events completed)
int drain_one(wr_bank *b)
{
	const int idx = (b->tHead - b->tCount + wr_bank::kTimeRing) % wr_bank::kTimeRing;
	float a = 0.0f, c = 0.0f;
	WR_CUDA(cudaEventSynchronize(b->tev[idx][2]));
	WR_CUDA(cudaEventElapsedTime(&a, b->tev[idx][0], b->tev[idx][1]));
	WR_CUDA(cudaEventElapsedTime(&c, b->tev[idx][1], b->tev[idx][2]));
	b->accMs[0] += a;
	b->accMs[1] += c;
	b->accBlocks++;
	b->tCount--;
	return WR_OK;
}

int apply_pending(wr_bank *b, cudaStream_t st)
{
	std::lock_guard<std::mutex> lk(b->mu);
	if (!(b->confDirty || b->taps1Dirty || b->taps2Dirty || b->resetDirty || b->tableDirty))
		return WR_OK;
	// the pinned staging buffers may still be the source of an earlier async upload
	WR_CUDA(cudaEventSynchronize(b->stagingFree));
	if (b->tableDirty) {
		memcpy(b->p_table, b->h_table.data(), sizeof(float) * WR_SINTABLE_SIZE);
		WR_CUDA(cudaMemcpyAsync(b->d_table, b->p_table, sizeof(float) * WR_SINTABLE_SIZE,
				cudaMemcpyHostToDevice, st));
		b->tableDirty = false;
		int rc = wrd::v2_set_table(b->v2, b->h_table.data(), st);
		if (rc != WR_OK)
			return rc;
		rc = wrd::v3_set_table(b->v3, b->h_table.data(), st);
		if (rc != WR_OK)
			return rc;
	}
	if (b->confDirty) {
		b->anyFM = false;
		for (unsigned r = 0; r < b->R; r++)
			b->anyFM = b->anyFM || b->h_conf[r].mode == WR_MODE_FM;
		memcpy(b->p_conf, b->h_conf.data(), sizeof(RxConf) * b->R);
		WR_CUDA(cudaMemcpyAsync(b->d_conf, b->p_conf, sizeof(RxConf) * b->R, cudaMemcpyHostToDevice, st));
		b->confDirty = false;
		if (b->streamsDirty) {
			int rc = wrd::v2_set_groups(b->v2, b->h_conf.data(), b->R, st);
			if (rc != WR_OK)
				return rc;
			rc = wrd::v3_set_groups(b->v3, b->h_conf.data(), b->R, st);
			if (rc != WR_OK)
				return rc;
			wrd::v4_set_groups(b->v4, b->h_conf.data(), b->R, b->T);
			b->streamsDirty = false;
		}
	}
	if (b->taps1Dirty) {
		memcpy(b->p_taps1, b->h_taps1.data(), sizeof(float) * b->h_taps1.size());
		WR_CUDA(cudaMemcpyAsync(b->d_taps1, b->p_taps1, sizeof(float) * b->h_taps1.size(),
				cudaMemcpyHostToDevice, st));
		b->taps1Dirty = false;
	}
	if (b->taps2Dirty) {
		memcpy(b->p_taps2, b->h_taps2.data(), sizeof(float) * b->h_taps2.size());
		WR_CUDA(cudaMemcpyAsync(b->d_taps2, b->p_taps2, sizeof(float) * b->h_taps2.size(),
				cudaMemcpyHostToDevice, st));
		b->taps2Dirty = false;
	}
	WR_CUDA(cudaEventRecord(b->stagingFree, st));
	if (b->resetDirty) {
		for (unsigned r = 0; r < b->R; r++) {
			if (!b->h_reset[r])
				continue;
			reset_kernel<<<1, 128, 0, st>>>(b->d_state[b->cur], b->d_hist1[b->cur], b->d_demod[b->cur],
					r, b->n1 - 1, b->n2 - 1, b->dstride, b->h_reset[r], b->h_phase[r], b->h_prev[r]);
			b->launches++;
			b->h_reset[r] = 0;
		}
		WR_CUDA(cudaGetLastError());
		b->resetDirty = false;
	}
	return WR_OK;
}

// Tile sizes of the v1 kernels: as many outputs per CTA as keep the mixed tile under ~40 KiB.
unsigned pick_tk_v1(unsigned n1, unsigned d1)
{
	const unsigned budgetFrames = 40 * 1024 / 8;
	unsigned tk = 128;
	while (tk > 8 && (size_t)tk * d1 + n1 > budgetFrames)
		tk /= 2;
	return tk;
}

// The hand-over scheme a bank starts with (WR_HAND_IN / WR_HAND_OUT override it, for experiments).
void default_handover(wr_bank *b)
{
	b->handIn = stream_ops().ok ? IN_FLAG : IN_COPYFLAG;
	b->handOut = OUT_EVENT;
	if (const char *e = getenv("WR_HAND_IN"))
		b->handIn = std::min(std::max(atoi(e), 0), 2);
	if (const char *e = getenv("WR_HAND_OUT"))
		b->handOut = std::min(std::max(atoi(e), 0), 2);
	if (!stream_ops().ok) {
		if (b->handIn == IN_FLAG) b->handIn = IN_COPYFLAG;
		if (b->handOut == OUT_FLAG) b->handOut = OUT_EVENT;
	}
}

// device address of a caller's pinned buffer, or nullptr if the device cannot address it
void *device_view(wr_bank *b, const void *host, size_t bytes)
{
	for (auto &m : b->hostMap)
		if (m.host == host && m.bytes >= bytes)
			return m.dev;
	cudaPointerAttributes at;
	if (cudaPointerGetAttributes(&at, host) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer) {
		cudaGetLastError();
		return nullptr;
	}
	// the whole range must be pinned: ask about its last byte too
	cudaPointerAttributes at2;
	if (bytes > 1 && (cudaPointerGetAttributes(&at2, static_cast<const char*>(host) + bytes - 1) != cudaSuccess
			|| at2.type != cudaMemoryTypeHost || !at2.devicePointer)) {
		cudaGetLastError();
		return nullptr;
	}
	auto &m = b->hostMap[b->hostMapNext];
	b->hostMapNext = (b->hostMapNext + 1) % 8;
	m.host = host; m.dev = at.devicePointer; m.bytes = bytes;
	return m.dev;
}

// can this block go through the v4 channel kernel (streaming FIR; banks that fill the grid with long runs)?
bool block_uses_v4(const wr_bank *b, unsigned F, bool u8, const void *iq, size_t stream_stride, wrd::V4Launch *shape)
{
	wrd::V4Launch tmp;
	return (b->variant == 4 || b->variant == 0)
			&& wrd::v4_shape(b->v4, b->v3, b->R, F, iq, u8, stream_stride, b->variant == 4, shape ? shape : &tmp);
}

// can this block go through the v3 channel kernel?  (v3 and v4 know the flag hand-over of the
// pipelined host path)
bool block_uses_v3(const wr_bank *b, unsigned F)
{
	return (b->variant == 3 || b->variant == 0) && wrd::v3_supported(b->v3, F);
}

int launch_block(wr_bank *b, const void *iq_dev, bool u8, size_t stream_stride, unsigned F,
		float *audio_dev, size_t audio_stride, cudaStream_t st, unsigned seq = 0, int handIn = IN_EVENT, int handOut = OUT_EVENT)
{
	int rc = apply_pending(b, st);
	if (rc != WR_OK)
		return rc;
	const unsigned M1 = F / b->d1;
	const unsigned M2 = M1 / b->d2;
	const int cur = b->cur, nxt = cur ^ 1;

	cudaEvent_t *tev = nullptr;
	if (b->timing) {
		if (b->tCount == wr_bank::kTimeRing && (rc = drain_one(b)) != WR_OK)
			return rc;
		tev = b->tev[b->tHead];
		WR_CUDA(cudaEventRecord(tev[0], st));
	}

	wrd::V4Launch shape4;
	const bool useV4 = block_uses_v4(b, F, u8, iq_dev, stream_stride, &shape4);
	if (b->variant == 4 && !useV4) {
		wr::set_error("v4 kernels do not serve this block (n1=%u d1=%u, %u frames, %s input, up to %u receivers per stream)",
				b->n1, b->d1, F, u8 ? "u8" : "float", b->v4.maxPerStream);
		return WR_EINVAL;
	}
	const bool useV3 = !useV4 && block_uses_v3(b, F);
	if (b->variant == 3 && !useV3) {
		wr::set_error("v3 kernels do not support this geometry (n1=%u d1=%u, %u frames)", b->n1, b->d1, F);
		return WR_EINVAL;
	}
	bool useV2 = useV4 || useV3 || (b->variant == 2) || (b->variant == 0 && wrd::v2_supported(b->v2));
	if (b->variant == 2 && !wrd::v2_supported(b->v2)) {
		wr::set_error("v2 kernels do not support this geometry (n1=%u d1=%u)", b->n1, b->d1);
		return WR_EINVAL;
	}
	if (!useV4 && !useV3 && !b->v3.ok && !b->warnedSlow && b->variant == 0) {
		// Say so ONCE when a geometry lands on the older kernel families: v3/v4 are instantiated for
		// the BASELINE geometries only (profiles/README.md has the per-geometry throughput table).
		// (v3.ok is also false for an NCO table the 16-bit compression cannot represent.)
		b->warnedSlow = true;
		if (!getenv("WEBRADIO_B200_QUIET"))
			fprintf(stderr, "webradio_b200: channel filter of %u taps, decimation %u has no fused v3/v4 kernel: using the %s "
					"kernels (1.8-5x slower; instantiated geometries: 64/10, 64/8, 127/50, 127/40, 255/50)\n",
					b->n1, b->d1, useV2 ? "v2 (tiled)" : "v1 (generic)");
	}
	if (useV2 && !b->d_chan)
		WR_CUDA(cudaMalloc(&b->d_chan, 2 * sizeof(float2) * (size_t)b->R * std::max(1u, b->maxM1)));
	float2 *const chanBuf = b->d_chan ? b->d_chan + (size_t)cur * b->R * std::max(1u, b->maxM1) : nullptr;
	if (u8 && !useV3 && !useV4) {
		// the older kernel families read float blocks: convert once into a scratch block
		if (!b->d_iqf)
			WR_CUDA(cudaMalloc(&b->d_iqf, sizeof(float) * 2 * (size_t)b->T * b->pitchF));
		dim3 grid(std::max(1u, std::min(64u, (2 * F + 1023) / 1024)), b->T);
		u8_to_f32_kernel<<<grid, 256, 0, st>>>(static_cast<const unsigned char*>(iq_dev), b->d_iqf,
				2 * stream_stride, 2 * (size_t)b->pitchF, 2 * F);
		b->launches++;
		WR_CUDA(cudaGetLastError());
		iq_dev = b->d_iqf;
		stream_stride = b->pitchF;
		u8 = false;
	}

	wrd::ChanArgs ca;
	ca.iq = reinterpret_cast<const float2*>(iq_dev);
	ca.stream_stride = stream_stride;
	ca.table = b->d_table;
	ca.taps1 = b->d_taps1;
	ca.conf = b->d_conf;
	ca.st_in = b->d_state[cur];
	ca.st_out = b->d_state[nxt];
	ca.hist_in = b->d_hist1[cur];
	ca.hist_out = b->d_hist1[nxt];
	ca.demod = b->d_demod[cur];
	ca.dstride = b->dstride;
	ca.demod_off = b->n2 - 1;
	ca.chan = (b->keepChan || useV2) ? chanBuf : nullptr; // v2 always materialises the channel stream
	ca.chan_stride = b->maxM1;
	ca.F = F;
	ca.M1 = M1;
	ca.n1 = b->n1;
	ca.d1 = b->d1;
	ca.in_flag = handIn != IN_EVENT ? b->d_sync + 0 : nullptr;
	ca.in_seq = seq;
	ca.err = b->h_err;
	unsigned long long *ts = (b->d_ts && seq && seq <= kTraceBlocks) ? b->d_ts + (size_t)(seq - 1) * wrd::kTsWords : nullptr;
	ca.ts = ts;
	unsigned long long *cta_ts = (b->d_cta && seq >= b->ctaFirst && seq < b->ctaFirst + kCtaTraceBlocks)
			? b->d_cta + (size_t)(seq - b->ctaFirst) * 2 * (kCtaTraceChan + kCtaTraceDemod) : nullptr;
	ca.cta_ts = cta_ts;
	ca.wait_late = b->waitLate;
	ca.poll_ns = b->pollNs;

	if (useV4) {
		rc = wrd::v4_launch_chan(b->v4, b->v3, shape4, ca, b->R, u8, st, &b->launches);
		if (rc != WR_OK)
			return rc;
	} else if (useV3) {
		rc = wrd::v3_launch_chan(b->v3, ca, u8, st, &b->launches);
		if (rc != WR_OK)
			return rc;
	} else if (useV2) {
		rc = wrd::v2_launch_chan(b->v2, ca, b->R, st, &b->launches);
		if (rc != WR_OK)
			return rc;
	} else {
		ca.TK = pick_tk_v1(b->n1, b->d1);
		ca.ntiles = (M1 + ca.TK - 1) / ca.TK;
		size_t smem = sizeof(float2) * ((size_t)ca.TK * b->d1 + b->n1)
				+ sizeof(float) * ((b->n1 + 3) & ~3u) + sizeof(float2) * (ca.TK + 1);
		dim3 grid(ca.ntiles + 1, b->R);
		wrd::chan_kernel_v1<kThreadsV1><<<grid, kThreadsV1, smem, st>>>(ca);
		b->launches++;
	}
	WR_CUDA(cudaGetLastError());
	if (tev)
		WR_CUDA(cudaEventRecord(tev[1], st));

	if (useV2) {
		// demodulator + audio FIR over the channel-rate IQ the v2 kernel wrote
		wrd::DemodAudioArgs da;
		da.chan = chanBuf;
		da.chan_stride = b->maxM1;
		da.conf = b->d_conf;
		da.st_in = b->d_state[cur];
		da.st_out = b->d_state[nxt];
		da.x = b->d_demod[cur];
		da.x_next = b->d_demod[nxt];
		da.dstride = b->dstride;
		da.taps2 = b->d_taps2;
		da.audio = audio_dev;
		da.audio_stride = audio_stride;
		da.M1 = M1;
		da.M2 = M2;
		da.n2 = b->n2;
		da.d2 = b->d2;
		// Audio outputs per CTA.  A tile demodulates the n2-1 samples in front of it again, so large
		// tiles do less work -- but the NEXT block's channel kernel starts on an SM only when this
		// grid's CTAs have drained from it, so when the whole grid is resident at once (a few CTAs
		// per SM) what counts is the life time of a CTA: 128 outputs, staged in one round.  Fewer
		// still when that would leave most SMs without a CTA (a single receiver's block is 16
		// tiles of 128).
		da.TK = (unsigned long long)b->R * ((M2 + 255) / 256) >= 16ull * (unsigned)b->numSMs ? 256 : 128;
		while (da.TK > 16 && (unsigned long long)b->R * ((M2 + da.TK - 1) / da.TK) < 296ull)
			da.TK /= 2;
		// A large bank at the demodulator's own rate (cfg3): tiles of 1024 outputs, eight per thread
		// through registers (demod_audio_kernel_v2's sliding-window path) -- a sixteenth of the
		// samples is demodulated twice instead of a quarter, and the grid is one wave.
		if (b->d2 == 1 && (b->n2 & 3u) == 0 && b->bigTiles && M2 >= 1024
				&& (unsigned long long)b->R * ((M2 + 1023) / 1024) >= 8ull * (unsigned)b->numSMs)
			da.TK = 1024;
		da.ntiles = (M2 + da.TK - 1) / da.TK;
		da.out_scale = b->outScale;
		da.negzero = -0.0f;
		da.done_count = (handOut != OUT_EVENT || ts) ? b->d_sync + 2 : nullptr;
		da.done_flag = handOut == OUT_FLAG ? b->d_sync + 1 : nullptr;
		da.host_done = handOut == OUT_DIRECT ? b->h_err + 1 : nullptr;
		da.done_seq = seq;
		da.ts = ts;
		da.cta_ts = cta_ts ? cta_ts + 2 * kCtaTraceChan : nullptr;
		size_t lmax = (size_t)da.TK * b->d2 + b->n2 - 1;
		size_t smem = sizeof(float) * (((lmax + 3) & ~(size_t)3) + b->n2);
		const unsigned long long items = (unsigned long long)(da.ntiles + 1) * b->R;
		WR_REQUIRE(items <= 0x7FFFFFFFull, WR_EINVAL, "receiver bank too large for one launch");
		da.items = (unsigned)items;
		// Grid: one CTA per item -- or, for a small bank under the v3 channel kernel, a PERSISTENT
		// grid of two CTAs per SM that walk the items.  Two of these CTAs fit beside the channel
		// kernel's CTA (registers), so the next block's channel kernel starts on every SM at once
		// instead of waiting for seven short-lived CTAs per SM to drain (cfg2: 21.8 -> 20.7 us per
		// block); the grid then runs for most of that kernel's life time, which a large bank
		// cannot afford (cfg5, 17408 items: 0.91 -> 1.14 ms).
		int perSM = b->demodPerSM;
		if (perSM < 0) {
			// ... and only if two of them really fit beside the channel kernel's CTA: a second
			// one that does not would hold that CTA back for its whole life time (cfg2 fed raw
			// bytes, whose channel kernel needs 8 more registers per thread: 21 -> 33 us)
			if (b->demodRegs == 0) {
				cudaFuncAttributes fa;
				WR_CUDA(cudaFuncGetAttributes(&fa, wrd::demod_audio_kernel_v2<wrd::kDemodThreads>));
				b->demodRegs = fa.numRegs;
			}
			const int fit = useV3 ? wrd::v3_spare_regs(b->v3, u8) / (((b->demodRegs + 7) & ~7) * wrd::kDemodThreads) : 0;
			// banks without an FM receiver demodulate for next to nothing (a square root or a sum
			// per sample), so they can afford the persistent grid at several times the size
			// (cfg3, 9216 items: 291.7 -> 287.1 us)
			const unsigned long long cap = (b->anyFM ? 10ull : 64ull) * (unsigned)b->numSMs;
			perSM = (useV3 && b->v3.pdl && fit >= 2 && items <= cap) ? 2 : 0;   // (v4's CTA leaves no room: one CTA per item)
		}
		dim3 grid(perSM > 0 ? (unsigned)std::min<unsigned long long>(items, (unsigned long long)perSM * (unsigned)b->numSMs) : (unsigned)items);
		if (cta_ts && grid.x > kCtaTraceDemod)
			da.cta_ts = nullptr;
		cudaLaunchConfig_t cfg = {};
		cudaLaunchAttribute attr[1];
		cfg.gridDim = grid;
		cfg.blockDim = dim3(wrd::kDemodThreads);
		cfg.dynamicSmemBytes = smem;
		cfg.stream = st;
		attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		attr[0].val.programmaticStreamSerializationAllowed = 1;
		cfg.attrs = attr;
		cfg.numAttrs = ((useV3 && b->v3.pdl) || (useV4 && b->v4.pdl)) ? 1 : 0;   // the v3/v4 channel kernels release their dependents early
		WR_CUDA(cudaLaunchKernelEx(&cfg, wrd::demod_audio_kernel_v2<wrd::kDemodThreads>, (const wrd::DemodAudioArgs)da));
		b->launches++;
	} else {
		wrd::AudioArgs aa;
		aa.x = b->d_demod[cur];
		aa.x_next = b->d_demod[nxt];
		aa.dstride = b->dstride;
		aa.taps2 = b->d_taps2;
		aa.audio = audio_dev;
		aa.audio_stride = audio_stride;
		aa.M1 = M1;
		aa.M2 = M2;
		aa.n2 = b->n2;
		aa.d2 = b->d2;
		aa.TK = 128;
		aa.ntiles = (M2 + aa.TK - 1) / aa.TK;
		aa.out_scale = b->outScale;
		size_t lmax = (size_t)(aa.TK - 1) * b->d2 + b->n2;
		size_t smem = sizeof(float) * (((lmax + 3) & ~(size_t)3) + b->n2);
		dim3 grid(aa.ntiles + 1, b->R);
		wrd::audio_kernel_v1<kThreadsV1><<<grid, kThreadsV1, smem, st>>>(aa);
		b->launches++;
	}
	WR_CUDA(cudaGetLastError());
	if (tev) {
		WR_CUDA(cudaEventRecord(tev[2], st));
		b->tHead = (b->tHead + 1) % wr_bank::kTimeRing;
		b->tCount++;
	}

	b->variantInUse = useV4 ? 4 : useV3 ? 3 : useV2 ? 2 : 1;
	b->chanSide = cur;
	b->cur = nxt;
	b->lastM1 = M1;
	b->lastM2 = M2;
	return WR_OK;
}

} // namespace

extern "C" int wr_plan_runs(unsigned ntaps, unsigned decimation, unsigned n_receivers, unsigned n_outputs, unsigned n_sms,
		unsigned *runs_per_receiver, unsigned *run_len, unsigned *long_runs, unsigned *rounds, unsigned *grid, unsigned *warps)
{
	wrd::V4Plan p;
	wrd::V4Launch L;
	if (!n_sms || !wrd::v4_pick(p, ntaps, decimation))
		return 0;
	p.numSMs = (int)n_sms;
	if (!wrd::v4_cut(p, n_receivers, n_outputs, wrd::kV4Warps, &L))
		return 0;
	if (runs_per_receiver) *runs_per_receiver = L.runsPerRx;
	if (run_len) *run_len = L.runLen;
	if (long_runs) *long_runs = L.longRuns;
	if (rounds) *rounds = L.rounds;
	if (grid) *grid = L.grid;
	if (warps) *warps = L.warps;
	return 1;
}

namespace {

void free_bank(wr_bank *b)
{
	if (!b)
		return;
	cudaSetDevice(b->device);
	if (b->compute)
		cudaStreamSynchronize(b->compute);
	if (b->h2d)
		cudaStreamSynchronize(b->h2d);
	if (b->d2h)
		cudaStreamSynchronize(b->d2h);
	wrd::v2_destroy(b->v2);
	wrd::v3_destroy(b->v3);
	cudaFree(b->d_table);
	cudaFree(b->d_taps1);
	cudaFree(b->d_taps2);
	cudaFree(b->d_conf);
	for (int i = 0; i < 2; i++) {
		cudaFree(b->d_state[i]);
		cudaFree(b->d_hist1[i]);
		cudaFree(b->d_demod[i]);
	}
	cudaFree(b->d_chan);
	cudaFree(b->d_iqf);
	if (b->tracePath && b->d_ts) {
		const unsigned n = std::min(b->seq, kTraceBlocks);
		std::vector<unsigned long long> dev((size_t)n * wrd::kTsWords);
		if (n && cudaMemcpy(dev.data(), b->d_ts, dev.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess) {
			if (FILE *f = fopen(b->tracePath, "a")) {
				fprintf(f, "# bank R=%u T=%u F=%u in=%d out=%d depth=%d\n", b->R, b->T, b->maxF, b->handIn, b->handOut, b->depth);
				fprintf(f, "seq,host_submit,host_submitted,host_waited,chan_start,chan_input,chan_end,demod_start,demod_end\n");
				for (unsigned i = 0; i < n; i++) {
					const unsigned long long *d = dev.data() + (size_t)i * wrd::kTsWords;
					fprintf(f, "%u,%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu\n", i + 1, b->hostTs[3 * i], b->hostTs[3 * i + 1], b->hostTs[3 * i + 2],
							d[wrd::kTsChanStart], d[wrd::kTsChanInput], d[wrd::kTsChanEnd], d[wrd::kTsDemodStart], d[wrd::kTsDemodEnd]);
				}
				fclose(f);
			}
		}
	}
	if (b->ctaPath && b->d_cta) {
		const size_t words = (size_t)2 * (kCtaTraceChan + kCtaTraceDemod) * kCtaTraceBlocks;
		std::vector<unsigned long long> dev(words);
		if (cudaMemcpy(dev.data(), b->d_cta, words * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess) {
			if (FILE *f = fopen(b->ctaPath, "a")) {
				fprintf(f, "# bank R=%u T=%u F=%u\nblock,kernel,cta,start,end\n", b->R, b->T, b->maxF);
				for (unsigned blk = 0; blk < kCtaTraceBlocks; blk++) {
					const unsigned long long *d = dev.data() + (size_t)blk * 2 * (kCtaTraceChan + kCtaTraceDemod);
					for (unsigned c = 0; c < kCtaTraceChan + kCtaTraceDemod; c++)
						if (d[2 * c])
							fprintf(f, "%u,%s,%u,%llu,%llu\n", b->ctaFirst + blk, c < kCtaTraceChan ? "chan" : "demod",
									c < kCtaTraceChan ? c : c - kCtaTraceChan, d[2 * c], d[2 * c + 1]);
				}
				fclose(f);
			}
		}
	}
	cudaFree(b->d_cta);
	cudaFree(b->d_ts);
	cudaFree(b->d_sync);
	cudaFreeHost(b->h_err);
	cudaFreeHost(b->p_seq);
	cudaFreeHost(b->p_conf);
	cudaFreeHost(b->p_taps1);
	cudaFreeHost(b->p_taps2);
	cudaFreeHost(b->p_table);
	for (int i = 0; i < kSlots; i++) {
		cudaFree(b->slot[i].d_iq);
		cudaFree(b->slot[i].d_audio);
		if (b->slot[i].in_ready) cudaEventDestroy(b->slot[i].in_ready);
		if (b->slot[i].done) cudaEventDestroy(b->slot[i].done);
		if (b->slot[i].out_ready) cudaEventDestroy(b->slot[i].out_ready);
	}
	if (b->stagingFree) cudaEventDestroy(b->stagingFree);
	for (int i = 0; i < wr_bank::kTimeRing; i++)
		for (int j = 0; j < 3; j++)
			if (b->tev[i][j]) cudaEventDestroy(b->tev[i][j]);
	if (b->compute) cudaStreamDestroy(b->compute);
	if (b->h2d) cudaStreamDestroy(b->h2d);
	if (b->d2h) cudaStreamDestroy(b->d2h);
	cudaGetLastError();
	delete b;
}

#define WR_BANK_ALLOC(expr)                                                             \
	do {                                                                                \
		cudaError_t e_ = (expr);                                                        \
		if (e_ != cudaSuccess) {                                                        \
			wr::set_error("%s: %s", #expr, cudaGetErrorString(e_));                     \
			free_bank(b);                                                               \
			return nullptr;                                                             \
		}                                                                               \
	} while (0)

} // namespace

extern "C" {

wr_bank *wr_bank_create(int device, unsigned n_streams, unsigned n_receivers, unsigned max_frames,
		unsigned n1, unsigned d1, unsigned n2, unsigned d2)
{
	if (n_streams == 0 || n_receivers == 0 || max_frames == 0 || n1 == 0 || n2 == 0 || d1 == 0 || d2 == 0
			|| n_receivers > 65535 || n1 > 4096 || n2 > 4096) {
		wr::set_error("wr_bank_create: bad geometry (streams=%u receivers=%u frames=%u n1=%u d1=%u n2=%u d2=%u)",
				n_streams, n_receivers, max_frames, n1, d1, n2, d2);
		return nullptr;
	}
	if (!wr::check_device(device))
		return nullptr;
	wr_bank *b = new wr_bank();
	b->device = device;
	b->T = n_streams; b->R = n_receivers; b->maxF = max_frames;
	b->pitchF = (max_frames + 1u) & ~1u;
	b->n1 = n1; b->d1 = d1; b->n2 = n2; b->d2 = d2;
	b->maxM1 = max_frames / d1;
	b->maxM2 = b->maxM1 / d2;
	b->dstride = ((size_t)(n2 - 1) + b->maxM1 + 3) & ~(size_t)3;
	const unsigned R = n_receivers;

	WR_BANK_ALLOC(cudaDeviceGetAttribute(&b->numSMs, cudaDevAttrMultiProcessorCount, device));
	WR_BANK_ALLOC(cudaStreamCreateWithFlags(&b->compute, cudaStreamNonBlocking));
	WR_BANK_ALLOC(cudaStreamCreateWithFlags(&b->h2d, cudaStreamNonBlocking));
	WR_BANK_ALLOC(cudaStreamCreateWithFlags(&b->d2h, cudaStreamNonBlocking));
	WR_BANK_ALLOC(cudaEventCreateWithFlags(&b->stagingFree, cudaEventDisableTiming));

	WR_BANK_ALLOC(cudaMalloc(&b->d_table, sizeof(float) * WR_SINTABLE_SIZE));
	WR_BANK_ALLOC(cudaMalloc(&b->d_taps1, sizeof(float) * (size_t)R * n1));
	WR_BANK_ALLOC(cudaMalloc(&b->d_taps2, sizeof(float) * (size_t)R * n2));
	WR_BANK_ALLOC(cudaMalloc(&b->d_conf, sizeof(RxConf) * R));
	const size_t h1 = std::max<size_t>(1, (size_t)R * (n1 - 1));
	for (int i = 0; i < 2; i++) {
		WR_BANK_ALLOC(cudaMalloc(&b->d_state[i], sizeof(RxState) * R));
		WR_BANK_ALLOC(cudaMemset(b->d_state[i], 0, sizeof(RxState) * R));
		WR_BANK_ALLOC(cudaMalloc(&b->d_hist1[i], sizeof(float2) * h1));
		WR_BANK_ALLOC(cudaMemset(b->d_hist1[i], 0, sizeof(float2) * h1));
		WR_BANK_ALLOC(cudaMalloc(&b->d_demod[i], sizeof(float) * (size_t)R * b->dstride));
		WR_BANK_ALLOC(cudaMemset(b->d_demod[i], 0, sizeof(float) * (size_t)R * b->dstride));
	}
	WR_BANK_ALLOC(cudaMallocHost(&b->p_conf, sizeof(RxConf) * R));
	WR_BANK_ALLOC(cudaMallocHost(&b->p_taps1, sizeof(float) * (size_t)R * n1));
	WR_BANK_ALLOC(cudaMallocHost(&b->p_taps2, sizeof(float) * (size_t)R * n2));
	WR_BANK_ALLOC(cudaMallocHost(&b->p_table, sizeof(float) * WR_SINTABLE_SIZE));
	WR_BANK_ALLOC(cudaMalloc(&b->d_sync, sizeof(unsigned) * 4));
	WR_BANK_ALLOC(cudaMemset(b->d_sync, 0, sizeof(unsigned) * 4));
	WR_BANK_ALLOC(cudaHostAlloc(&b->h_err, sizeof(unsigned) * 2, cudaHostAllocMapped));
	b->h_err[0] = b->h_err[1] = 0;
	WR_BANK_ALLOC(cudaMallocHost(&b->p_seq, sizeof(unsigned) * kSeqRing));
	default_handover(b);
	if (const char *e = getenv("WR_POLL_NS"))
		b->pollNs = (unsigned)atoi(e);
	if (const char *e = getenv("WR_DEMOD_PER_SM"))
		b->demodPerSM = atoi(e);
	if (const char *e = getenv("WR_WAIT_LATE"))
		b->waitLate = atoi(e) != 0;
	if (const char *e = getenv("WR_DEMOD_BIG_TILES"))
		b->bigTiles = atoi(e) != 0;
	if (const char *e = getenv("WR_SYNC_SPLIT"))
		b->syncSplit = (unsigned)std::max(0, atoi(e));
	if ((b->tracePath = getenv("WR_TRACE")) != nullptr) {
		WR_BANK_ALLOC(cudaMalloc(&b->d_ts, sizeof(unsigned long long) * wrd::kTsWords * kTraceBlocks));
		WR_BANK_ALLOC(cudaMemset(b->d_ts, 0, sizeof(unsigned long long) * wrd::kTsWords * kTraceBlocks));
		b->hostTs.assign((size_t)3 * kTraceBlocks, 0);
		if (const char *e = getenv("WR_TRACE_CTA_FIRST"))
			b->ctaFirst = (unsigned)atoi(e);
		if ((b->ctaPath = getenv("WR_TRACE_CTA")) != nullptr) {
			const size_t n = sizeof(unsigned long long) * 2 * (kCtaTraceChan + kCtaTraceDemod) * kCtaTraceBlocks;
			WR_BANK_ALLOC(cudaMalloc(&b->d_cta, n));
			WR_BANK_ALLOC(cudaMemset(b->d_cta, 0, n));
		}
	}
	{
		// pipeline depth: what the chain copy-in -> two kernels -> copy-out needs to stay full on
		// small blocks, bounded by ~6 GB of staging for large ones
		const size_t slotBytes = sizeof(float) * 2 * (size_t)n_streams * max_frames + sizeof(float) * (size_t)R * std::max(1u, b->maxM2);
		const size_t fit = ((size_t)6 << 30) / std::max<size_t>(1, slotBytes);
		b->depth = (int)std::max<size_t>(2, std::min<size_t>(kSlots, fit));
	}
	for (int i = 0; i < kSlots; i++) {
		WR_BANK_ALLOC(cudaEventCreateWithFlags(&b->slot[i].in_ready, cudaEventDisableTiming));
		WR_BANK_ALLOC(cudaEventCreateWithFlags(&b->slot[i].done, cudaEventDisableTiming));
		WR_BANK_ALLOC(cudaEventCreateWithFlags(&b->slot[i].out_ready, cudaEventDisableTiming));
	}

	b->h_conf.resize(R);
	for (unsigned r = 0; r < R; r++) {
		b->h_conf[r].step = 0;
		b->h_conf[r].mode = WR_MODE_AM;  // Demodulator's constructor default (demodulator.cxx:33)
		b->h_conf[r].stream = r % n_streams;
		b->h_conf[r].pad = 0;
	}
	b->h_taps1.assign((size_t)R * n1, 0.0f);
	b->h_taps2.assign((size_t)R * n2, 0.0f);
	b->h_reset.assign(R, 0);
	b->h_phase.assign(R, 0);
	b->h_prev.assign(R, make_float2(0.0f, 0.0f));
	b->h_table.resize(WR_SINTABLE_SIZE);
	wr_build_sintable(b->h_table.data());
	b->tableDirty = true;

	if (wrd::v2_init(b->v2, device, n1, d1) != WR_OK || wrd::v3_init(b->v3, device, n1, d1, max_frames) != WR_OK
			|| wrd::v4_init(b->v4, b->v3, device, n1, d1) != WR_OK) {
		free_bank(b);
		return nullptr;
	}
	return b;
}

void wr_bank_destroy(wr_bank *b) { free_bank(b); }

int wr_bank_set_sintable(wr_bank *b, const float *table)
{
	WR_REQUIRE(b && table, WR_EINVAL, "wr_bank_set_sintable: null argument");
	std::lock_guard<std::mutex> lk(b->mu);
	memcpy(b->h_table.data(), table, sizeof(float) * WR_SINTABLE_SIZE);
	b->tableDirty = true;
	return WR_OK;
}

int wr_rx_set_stream(wr_bank *b, unsigned rx, unsigned stream)
{
	WR_REQUIRE(b && rx < b->R && stream < b->T, WR_EINVAL, "wr_rx_set_stream: rx %u / stream %u out of range", rx, stream);
	std::lock_guard<std::mutex> lk(b->mu);
	b->h_conf[rx].stream = stream;
	b->confDirty = true;
	b->streamsDirty = true;
	return WR_OK;
}

int wr_rx_set_phase_step(wr_bank *b, unsigned rx, int32_t step)
{
	WR_REQUIRE(b && rx < b->R, WR_EINVAL, "wr_rx_set_phase_step: rx %u out of range", rx);
	std::lock_guard<std::mutex> lk(b->mu);
	b->h_conf[rx].step = step;
	b->confDirty = true;
	return WR_OK;
}

int wr_rx_set_taps(wr_bank *b, unsigned rx, int stage, const float *coeff, unsigned ntaps)
{
	WR_REQUIRE(b && coeff && rx < b->R && (stage == 0 || stage == 1), WR_EINVAL, "wr_rx_set_taps: bad argument");
	const unsigned n = stage ? b->n2 : b->n1;
	WR_REQUIRE(ntaps == n, WR_EINVAL, "wr_rx_set_taps: %u taps given, bank geometry has %u", ntaps, n);
	std::lock_guard<std::mutex> lk(b->mu);
	float *dst = (stage ? b->h_taps2.data() : b->h_taps1.data()) + (size_t)rx * n;
	for (unsigned j = 0; j < n; j++)
		dst[j] = coeff[n - 1 - j]; // the kernels walk taps in sample order (lowpass.cxx:152-156)
	(stage ? b->taps2Dirty : b->taps1Dirty) = true;
	return WR_OK;
}

int wr_bank_design_taps(wr_bank *b, int stage, const unsigned *passband_hz, unsigned sample_rate)
{
	WR_REQUIRE(b && passband_hz && sample_rate > 0 && (stage == 0 || stage == 1), WR_EINVAL, "wr_bank_design_taps: bad argument");
	if (!wr::use_device(b->device))
		return WR_ENODEV;
	const unsigned n = stage ? b->n2 : b->n1;
	WR_REQUIRE(n >= 2, WR_EINVAL, "wr_bank_design_taps: %u taps", n);
	// pending setter changes first, so that the copy back below leaves host and device identical
	int rc = apply_pending(b, b->compute);
	if (rc != WR_OK)
		return rc;
	std::vector<double> costab(n);
	std::vector<float> window(n);
	const double two_pi = 6.283185307179586476925286766559;
	for (unsigned i = 0; i < n; i++) {
		costab[i] = cos(two_pi * (double)i / (double)n);
		// lowpass.cxx:108-109: Hamming window, then the 1/N of the unnormalised transform
		float w = (float)(0.54 - 0.46 * cosf((float)(2 * M_PI * (float)i / (float)(n - 1))));
		w /= (float)n;
		window[i] = w;
	}
	double *d_cos = nullptr;
	float *d_win = nullptr;
	unsigned *d_pb = nullptr;
	cudaError_t e = cudaMalloc(&d_cos, sizeof(double) * n);
	if (e == cudaSuccess) e = cudaMalloc(&d_win, sizeof(float) * n);
	if (e == cudaSuccess) e = cudaMalloc(&d_pb, sizeof(unsigned) * b->R);
	if (e == cudaSuccess) e = cudaMemcpyAsync(d_cos, costab.data(), sizeof(double) * n, cudaMemcpyHostToDevice, b->compute);
	if (e == cudaSuccess) e = cudaMemcpyAsync(d_win, window.data(), sizeof(float) * n, cudaMemcpyHostToDevice, b->compute);
	if (e == cudaSuccess) e = cudaMemcpyAsync(d_pb, passband_hz, sizeof(unsigned) * b->R, cudaMemcpyHostToDevice, b->compute);
	float *d_taps = stage ? b->d_taps2 : b->d_taps1;
	std::vector<float> &h_taps = stage ? b->h_taps2 : b->h_taps1;
	if (e == cudaSuccess) {
		design_kernel<<<b->R, 128, 0, b->compute>>>(d_cos, d_win, d_pb, sample_rate, n, d_taps);
		b->launches++;
		e = cudaGetLastError();
	}
	{
		std::lock_guard<std::mutex> lk(b->mu);
		if (e == cudaSuccess)
			e = cudaMemcpyAsync(h_taps.data(), d_taps, sizeof(float) * h_taps.size(), cudaMemcpyDeviceToHost, b->compute);
		if (e == cudaSuccess)
			e = cudaStreamSynchronize(b->compute);
		(stage ? b->taps2Dirty : b->taps1Dirty) = false;
	}
	cudaFree(d_cos);
	cudaFree(d_win);
	cudaFree(d_pb);
	if (e != cudaSuccess) {
		wr::set_error("wr_bank_design_taps: %s", cudaGetErrorString(e));
		cudaGetLastError();
		return WR_ECUDA;
	}
	return WR_OK;
}

int wr_rx_get_taps(wr_bank *b, unsigned rx, int stage, float *coeff, unsigned ntaps)
{
	WR_REQUIRE(b && coeff && rx < b->R && (stage == 0 || stage == 1), WR_EINVAL, "wr_rx_get_taps: bad argument");
	const unsigned n = stage ? b->n2 : b->n1;
	WR_REQUIRE(ntaps == n, WR_EINVAL, "wr_rx_get_taps: %u taps asked, bank geometry has %u", ntaps, n);
	std::lock_guard<std::mutex> lk(b->mu);
	const float *src = (stage ? b->h_taps2.data() : b->h_taps1.data()) + (size_t)rx * n;
	for (unsigned j = 0; j < n; j++)
		coeff[j] = src[n - 1 - j];
	return WR_OK;
}

int wr_rx_set_mode(wr_bank *b, unsigned rx, int mode)
{
	WR_REQUIRE(b && rx < b->R, WR_EINVAL, "wr_rx_set_mode: rx %u out of range", rx);
	WR_REQUIRE(mode >= WR_MODE_AM && mode <= WR_MODE_LSB, WR_EINVAL, "wr_rx_set_mode: bad mode %d", mode);
	std::lock_guard<std::mutex> lk(b->mu);
	b->h_conf[rx].mode = mode;
	b->confDirty = true;
	return WR_OK;
}

int wr_rx_reset(wr_bank *b, unsigned rx, unsigned flags)
{
	WR_REQUIRE(b && rx < b->R, WR_EINVAL, "wr_rx_reset: rx %u out of range", rx);
	std::lock_guard<std::mutex> lk(b->mu);
	b->h_reset[rx] |= flags;
	if (flags & WR_RESET_DEMOD)
		b->h_reset[rx] &= ~kSetPrev;   // a reset after wr_rx_set_lookback wins
	b->resetDirty = true;
	return WR_OK;
}

int wr_rx_set_phase(wr_bank *b, unsigned rx, uint32_t phase)
{
	WR_REQUIRE(b && rx < b->R, WR_EINVAL, "wr_rx_set_phase: rx %u out of range", rx);
	std::lock_guard<std::mutex> lk(b->mu);
	b->h_reset[rx] = (b->h_reset[rx] & ~WR_RESET_PHASE) | kSetPhase;
	b->h_phase[rx] = phase;
	b->resetDirty = true;
	return WR_OK;
}

int wr_rx_get_phase(wr_bank *b, unsigned rx, uint32_t *phase)
{
	WR_REQUIRE(b && phase && rx < b->R, WR_EINVAL, "wr_rx_get_phase: bad argument");
	if (!wr::use_device(b->device))
		return WR_ENODEV;
	WR_CUDA(cudaStreamSynchronize(b->compute));
	RxState st;
	WR_CUDA(cudaMemcpy(&st, b->d_state[b->cur] + rx, sizeof(st), cudaMemcpyDeviceToHost));
	*phase = st.phase;
	return WR_OK;
}

int wr_rx_set_lookback(wr_bank *b, unsigned rx, const float *prev_iq)
{
	WR_REQUIRE(b && prev_iq && rx < b->R, WR_EINVAL, "wr_rx_set_lookback: bad argument");
	std::lock_guard<std::mutex> lk(b->mu);
	b->h_reset[rx] = (b->h_reset[rx] & ~WR_RESET_DEMOD) | kSetPrev;
	b->h_prev[rx] = make_float2(prev_iq[0], prev_iq[1]);
	b->resetDirty = true;
	return WR_OK;
}

int wr_rx_get_lookback(wr_bank *b, unsigned rx, float *prev_iq)
{
	WR_REQUIRE(b && prev_iq && rx < b->R, WR_EINVAL, "wr_rx_get_lookback: bad argument");
	if (!wr::use_device(b->device))
		return WR_ENODEV;
	WR_CUDA(cudaStreamSynchronize(b->compute));
	RxState st;
	WR_CUDA(cudaMemcpy(&st, b->d_state[b->cur] + rx, sizeof(st), cudaMemcpyDeviceToHost));
	prev_iq[0] = st.prev_i;
	prev_iq[1] = st.prev_q;
	return WR_OK;
}

static int process_device_any(wr_bank *b, const void *iq_dev, bool u8, size_t stream_stride_frames,
		unsigned nframes, float *audio_dev, size_t audio_stride, void *cuda_stream)
{
	WR_REQUIRE(b && iq_dev && audio_dev, WR_EINVAL, "wr_bank_process_device: null argument");
	WR_REQUIRE(nframes <= b->maxF, WR_EINVAL, "wr_bank_process_device: %u frames > max_frames %u", nframes, b->maxF);
	if (!wr::use_device(b->device))
		return WR_ENODEV;
	cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : b->compute;
	return launch_block(b, iq_dev, u8, stream_stride_frames, nframes, audio_dev, audio_stride, st, b->d_ts ? ++b->seq : 0);
}

int wr_bank_process_device(wr_bank *b, const float *iq_dev, size_t stream_stride_frames,
		unsigned nframes, float *audio_dev, size_t audio_stride, void *cuda_stream)
{
	return process_device_any(b, iq_dev, false, stream_stride_frames, nframes, audio_dev, audio_stride, cuda_stream);
}

int wr_bank_process_device_u8(wr_bank *b, const uint8_t *iq_dev, size_t stream_stride_frames,
		unsigned nframes, float *audio_dev, size_t audio_stride, void *cuda_stream)
{
	return process_device_any(b, iq_dev, true, stream_stride_frames, nframes, audio_dev, audio_stride, cuda_stream);
}

// One piece of work for the copy-in / kernels / copy-out pipeline: `nframes` frames of every stream,
// taken from a host block whose rows (streams) lie `host_pitch` frames apart.  wr_bank_submit hands
// over whole blocks (host_pitch = nframes); wr_bank_process cuts one block into consecutive
// sub-blocks, which the carried state turns into exactly the same samples.
// `iq_dev` != nullptr: the frames are (or will be, once `dev_ready` has fired) in HBM already --
// the shared upload of wr_bank_process_upload; nothing is copied in.
static int submit_any(wr_bank *b, const void *iq_host, bool u8, unsigned nframes, float *audio_host, size_t audio_stride,
		size_t host_pitch = 0, const float *iq_dev = nullptr, cudaEvent_t dev_ready = nullptr)
{
	if (host_pitch == 0)
		host_pitch = nframes;
	WR_REQUIRE(b && (iq_host || iq_dev) && audio_host, WR_EINVAL, "wr_bank_submit: null argument");
	WR_REQUIRE(nframes <= b->maxF, WR_EINVAL, "wr_bank_submit: %u frames > max_frames %u", nframes, b->maxF);
	WR_REQUIRE(b->inflight < b->depth, WR_ESTATE, "wr_bank_submit: %d blocks already in flight", b->inflight);
	if (!wr::use_device(b->device))
		return WR_ENODEV;
	const unsigned long long t_enter = b->d_ts ? host_ns() : 0;
	Slot &s = b->slot[b->head];
	if (!s.d_iq && !iq_dev)
		WR_CUDA(cudaMalloc(&s.d_iq, sizeof(float) * 2 * (size_t)b->T * b->pitchF));
	if (!s.d_audio)
		WR_CUDA(cudaMalloc(&s.d_audio, sizeof(float) * (size_t)b->R * std::max(1u, b->maxM2)));
	const unsigned M2 = nframes / b->d1 / b->d2;
	const unsigned maxM2 = std::max(1u, b->maxM2);
	const size_t fb = u8 ? 2 : sizeof(float) * 2;   // bytes per frame on the wire and in HBM
	// A slot is reused only after wr_bank_wait returned for the block that held it (inflight <
	// depth = the number of slots), i.e. after its kernels and its copy out completed: the host's
	// own order protects the slot buffers under every hand-over scheme below.
	//
	// Hand-over without events (v3 channel kernel): nothing but the two kernels goes into the
	// launch stream, so consecutive blocks keep their programmatic overlap.
	//   in : the copy-in stream raises a counter in HBM behind the tuner block (a stream memory
	//        operation, or a 4-byte copy) and the channel kernel's loaders wait for it;
	//   out: the demodulator kernel stores the audio straight into the caller's pinned buffer and
	//        its last CTA raises a counter in mapped host memory that wr_bank_wait polls -- or it
	//        raises a counter in HBM that the copy-out stream waits for.
	// Hand-over by events (other kernel families): events between the three streams.
	const bool v3 = block_uses_v4(b, nframes, u8, iq_dev ? (const void*)iq_dev : (const void*)s.d_iq, b->pitchF, nullptr) || block_uses_v3(b, nframes);
	int handIn = (v3 && !iq_dev) ? b->handIn : IN_EVENT, handOut = v3 ? b->handOut : OUT_EVENT;
	float *audio_dev = s.d_audio;
	size_t audio_dev_stride = maxM2;
	if (handOut == OUT_DIRECT) {
		void *dv = M2 ? device_view(b, audio_host, sizeof(float) * ((size_t)(b->R - 1) * audio_stride + M2)) : nullptr;
		if (dv) {
			audio_dev = static_cast<float*>(dv);
			audio_dev_stride = audio_stride;
		} else {
			handOut = stream_ops().ok ? OUT_FLAG : OUT_EVENT;   // pageable or foreign memory: copy it out
		}
	}
	const unsigned seq = ++b->seq;
	if (iq_dev) {
		if (dev_ready)
			WR_CUDA(cudaStreamWaitEvent(b->compute, dev_ready, 0));
	} else if (b->T == 1 || (nframes == b->pitchF && host_pitch == nframes))
		WR_CUDA(cudaMemcpyAsync(s.d_iq, iq_host, fb * (size_t)nframes * b->T, cudaMemcpyHostToDevice, b->h2d));
	else
		WR_CUDA(cudaMemcpy2DAsync(s.d_iq, fb * (size_t)b->pitchF, iq_host, fb * host_pitch,
				fb * (size_t)nframes, b->T, cudaMemcpyHostToDevice, b->h2d));
	if (iq_dev) {
		// (ordered by the event above)
	} else if (handIn == IN_FLAG) {
		if (stream_ops().write((CUstream)b->h2d, (CUdeviceptr)(uintptr_t)(b->d_sync + 0), seq, 0) != CUDA_SUCCESS) {
			wr::set_error("wr_bank_submit: cuStreamWriteValue32 failed");
			return WR_ECUDA;
		}
	} else if (handIn == IN_COPYFLAG) {
		// the ring entry is free again: depth < kSeqRing blocks can be in flight
		b->p_seq[seq % kSeqRing] = seq;
		WR_CUDA(cudaMemcpyAsync(b->d_sync + 0, b->p_seq + seq % kSeqRing, sizeof(unsigned), cudaMemcpyHostToDevice, b->h2d));
	} else {
		WR_CUDA(cudaEventRecord(s.in_ready, b->h2d));
		WR_CUDA(cudaStreamWaitEvent(b->compute, s.in_ready, 0));
	}
	int rc = launch_block(b, iq_dev ? (const void*)iq_dev : (const void*)s.d_iq, u8, b->pitchF, nframes, audio_dev, audio_dev_stride,
			b->compute, seq, handIn, handOut);
	if (rc != WR_OK)
		return rc;
	if (handOut != OUT_DIRECT) {
		if (handOut == OUT_FLAG) {
			if (stream_ops().wait((CUstream)b->d2h, (CUdeviceptr)(uintptr_t)(b->d_sync + 1), seq, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS) {
				wr::set_error("wr_bank_submit: cuStreamWaitValue32 failed");
				return WR_ECUDA;
			}
		} else {
			WR_CUDA(cudaEventRecord(s.done, b->compute));
			WR_CUDA(cudaStreamWaitEvent(b->d2h, s.done, 0));
		}
		if (M2 > 0) {
			if (audio_stride == M2 && maxM2 == M2)
				WR_CUDA(cudaMemcpyAsync(audio_host, s.d_audio, sizeof(float) * (size_t)M2 * b->R, cudaMemcpyDeviceToHost, b->d2h));
			else
				WR_CUDA(cudaMemcpy2DAsync(audio_host, sizeof(float) * audio_stride, s.d_audio, sizeof(float) * maxM2,
						sizeof(float) * M2, b->R, cudaMemcpyDeviceToHost, b->d2h));
		}
		WR_CUDA(cudaEventRecord(s.out_ready, b->d2h));
	}
	s.busy = true;
	s.used = true;
	s.seq = seq;
	s.outMode = handOut;
	b->head = (b->head + 1) % b->depth;
	b->inflight++;
	if (b->d_ts && seq <= kTraceBlocks) {
		b->hostTs[3 * (size_t)(seq - 1)] = t_enter;
		b->hostTs[3 * (size_t)(seq - 1) + 1] = host_ns();
	}
	return WR_OK;
}

int wr_bank_submit(wr_bank *b, const float *iq_host, unsigned nframes, float *audio_host, size_t audio_stride)
{
	return submit_any(b, iq_host, false, nframes, audio_host, audio_stride);
}

int wr_bank_submit_u8(wr_bank *b, const uint8_t *iq_host, unsigned nframes, float *audio_host, size_t audio_stride)
{
	return submit_any(b, iq_host, true, nframes, audio_host, audio_stride);
}

int wr_bank_wait(wr_bank *b)
{
	WR_REQUIRE(b, WR_EINVAL, "wr_bank_wait: null bank");
	WR_REQUIRE(b->inflight > 0, WR_ESTATE, "wr_bank_wait: nothing in flight");
	if (!wr::use_device(b->device))
		return WR_ENODEV;
	Slot &s = b->slot[b->tail];
	if (s.outMode == OUT_DIRECT) {
		// the demodulator kernel's last CTA publishes the block's number behind its audio
		volatile unsigned *done = b->h_err + 1;
		const unsigned long long t0 = host_ns();
		for (unsigned spins = 0; (int)(*done - s.seq) < 0; spins++) {
			__builtin_ia32_pause();
			if ((spins & 0xFFFu) == 0xFFFu) {
				// a failed launch or a dead context must not spin forever
				cudaError_t e = cudaStreamQuery(b->compute);
				if (e != cudaSuccess && e != cudaErrorNotReady) {
					wr::set_error("wr_bank_wait: %s", cudaGetErrorString(e));
					return WR_ECUDA;
				}
				if (e == cudaSuccess && (int)(*done - s.seq) < 0 && host_ns() - t0 > 10000000000ull) {
					wr::set_error("wr_bank_wait: the launch stream drained without the block's completion word");
					return WR_ECUDA;
				}
			}
		}
		__atomic_thread_fence(__ATOMIC_ACQUIRE);
	} else {
		WR_CUDA(cudaEventSynchronize(s.out_ready));
	}
	s.busy = false;
	b->tail = (b->tail + 1) % b->depth;
	b->inflight--;
	if (b->d_ts && s.seq && s.seq <= kTraceBlocks)
		b->hostTs[3 * (size_t)(s.seq - 1) + 2] = host_ns();
	if (*static_cast<volatile unsigned*>(b->h_err) & wrd::kSyncTimeout) {
		*b->h_err = 0;
		wr::set_error("wr_bank_wait: the channel kernel gave up waiting for its tuner block (copy-in stream stalled)");
		return WR_ECUDA;
	}
	return WR_OK;
}

int wr_bank_pipeline_depth(const wr_bank *b) { return b ? b->depth : kSlots; }

int wr_bank_set_handover(wr_bank *b, int scheme)
{
	WR_REQUIRE(b && scheme >= WR_HANDOVER_EVENTS && scheme <= WR_HANDOVER_DIRECT, WR_EINVAL, "wr_bank_set_handover: bad argument");
	WR_REQUIRE(b->inflight == 0, WR_ESTATE, "wr_bank_set_handover: blocks in flight");
	if (scheme == WR_HANDOVER_EVENTS) {
		b->handIn = IN_EVENT;
		b->handOut = OUT_EVENT;
	} else {
		b->handIn = stream_ops().ok ? IN_FLAG : IN_COPYFLAG;
		b->handOut = scheme == WR_HANDOVER_DIRECT ? OUT_DIRECT : OUT_EVENT;
	}
	return scheme;
}

static int process_any(wr_bank *b, const void *iq_host, bool u8, unsigned nframes, float *audio_host, size_t audio_stride)
{
	WR_REQUIRE(b, WR_EINVAL, "wr_bank_process: null bank");
	WR_REQUIRE(b->inflight == 0, WR_ESTATE, "wr_bank_process: pipelined blocks still in flight");
	// One synchronous call, several pieces in flight inside it: the block is cut into consecutive
	// sub-blocks so that the copy-in of piece k+1 runs under the kernels of piece k and under the
	// copy-out of piece k-1.  A piece is a whole number of audio frames (and of the channel kernel's
	// passes' worth of frames, so that it stays on the same kernel family as a full block).
	unsigned pieces = b->keepChan ? 1u : b->syncSplit;     // (wr_bank_read_stage reads the last launch: keep it the whole block)
	const unsigned quantum = b->d1 * b->d2;
	if (pieces == 0) {
		// by size.  Every piece costs about ten runtime calls (~15 us of host time), so a block of
		// under a few MB is cut in two at most (measured on cfg2, 819 KB: 1 piece 88 k, 2 pieces
		// 96 k, 3 pieces 85 k, 4 pieces 79 k MS/s); large blocks in four
		const size_t bytes = (u8 ? 2 : 8) * (size_t)nframes * b->T;
		pieces = bytes >= (8u << 20) ? 4u : bytes >= (512u << 10) ? 2u : 1u;
	}
	pieces = std::min<unsigned>(pieces, (unsigned)b->depth);
	unsigned per = nframes / std::max(1u, pieces);
	per -= per % quantum;
	const unsigned minPiece = std::max(quantum, b->v3.ok ? b->v3.SF : 0u);
	if (pieces <= 1 || per < minPiece) {
		int rc = submit_any(b, iq_host, u8, nframes, audio_host, audio_stride);
		if (rc != WR_OK)
			return rc;
		return wr_bank_wait(b);
	}
	const size_t fb = u8 ? 2 : 8;
	int rc = WR_OK;
	unsigned done = 0, submitted = 0;
	for (unsigned k = 0; k < pieces && rc == WR_OK; k++) {
		const unsigned nf = (k + 1 == pieces) ? nframes - done : per;     // the last piece takes the remainder
		rc = submit_any(b, static_cast<const char*>(iq_host) + fb * done, u8, nf,
				audio_host + done / quantum, audio_stride, nframes);
		if (rc == WR_OK)
			submitted++;
		done += nf;
	}
	for (unsigned k = 0; k < submitted; k++) {
		const int rw = wr_bank_wait(b);
		if (rc == WR_OK)
			rc = rw;
	}
	return rc;
}

int wr_bank_process(wr_bank *b, const float *iq_host, unsigned nframes, float *audio_host, size_t audio_stride)
{
	return process_any(b, iq_host, false, nframes, audio_host, audio_stride);
}

int wr_bank_process_upload(wr_bank *b, wr_upload *u, unsigned nframes, float *audio_host, size_t audio_stride)
{
	WR_REQUIRE(b && u && audio_host, WR_EINVAL, "wr_bank_process_upload: null argument");
	WR_REQUIRE(b->T == 1, WR_EINVAL, "wr_bank_process_upload: the bank has %u streams, an upload carries one", b->T);
	WR_REQUIRE(u->device == b->device, WR_EINVAL, "wr_bank_process_upload: upload on device %d, bank on %d", u->device, b->device);
	WR_REQUIRE(nframes == u->nframes && nframes <= b->maxF, WR_EINVAL, "wr_bank_process_upload: %u frames asked, %u uploaded, bank holds %u",
			nframes, u->nframes, b->maxF);
	WR_REQUIRE(b->inflight == 0, WR_ESTATE, "wr_bank_process_upload: pipelined blocks still in flight");
	if (!wr::use_device(b->device))
		return WR_ENODEV;
	// the bank's own sub-blocks (whole audio frames), each started as soon as the upload's piece
	// that completes it has landed; the audio of one leaves while the next is computed
	const unsigned quantum = b->d1 * b->d2;
	unsigned pieces = b->keepChan ? 1u : b->syncSplit ? b->syncSplit : std::min<unsigned>(u->npieces ? u->npieces : 1u, (unsigned)b->depth);
	pieces = std::min<unsigned>(pieces, (unsigned)b->depth);
	unsigned per = nframes / std::max(1u, pieces);
	per -= per % quantum;
	const unsigned minPiece = std::max(quantum, b->v3.ok ? b->v3.SF : 0u);
	if (pieces <= 1 || per < minPiece) {
		pieces = 1;
		per = nframes;
	}
	int rc = WR_OK;
	unsigned done = 0, submitted = 0;
	for (unsigned k = 0; k < pieces && rc == WR_OK; k++) {
		const unsigned nf = (k + 1 == pieces) ? nframes - done : per;
		rc = submit_any(b, nullptr, false, nf, audio_host + done / quantum, audio_stride, nframes,
				u->dev() + 2 * (size_t)done, u->ready(done + nf));
		if (rc == WR_OK)
			submitted++;
		done += nf;
	}
	for (unsigned k = 0; k < submitted; k++) {
		const int rw = wr_bank_wait(b);
		if (rc == WR_OK)
			rc = rw;
	}
	return rc;
}

int wr_rx_get_history(wr_bank *b, unsigned rx, int stage, float *out, unsigned nfloats)
{
	WR_REQUIRE(b && out && rx < b->R && (stage == 0 || stage == 1), WR_EINVAL, "wr_rx_get_history: bad argument");
	const unsigned want = stage ? b->n2 - 1 : 2 * (b->n1 - 1);
	WR_REQUIRE(nfloats == want, WR_EINVAL, "wr_rx_get_history: %u floats asked, the stage carries %u", nfloats, want);
	if (!wr::use_device(b->device))
		return WR_ENODEV;
	WR_CUDA(cudaStreamSynchronize(b->compute));
	if (want == 0)
		return WR_OK;
	const void *src = stage ? (const void*)(b->d_demod[b->cur] + (size_t)rx * b->dstride)
			: (const void*)(b->d_hist1[b->cur] + (size_t)rx * (b->n1 - 1));
	WR_CUDA(cudaMemcpy(out, src, sizeof(float) * want, cudaMemcpyDeviceToHost));
	return WR_OK;
}

int wr_rx_set_history(wr_bank *b, unsigned rx, int stage, const float *in, unsigned nfloats)
{
	WR_REQUIRE(b && in && rx < b->R && (stage == 0 || stage == 1), WR_EINVAL, "wr_rx_set_history: bad argument");
	const unsigned want = stage ? b->n2 - 1 : 2 * (b->n1 - 1);
	WR_REQUIRE(nfloats == want, WR_EINVAL, "wr_rx_set_history: %u floats given, the stage carries %u", nfloats, want);
	WR_REQUIRE(b->inflight == 0, WR_ESTATE, "wr_rx_set_history: blocks in flight");
	if (!wr::use_device(b->device))
		return WR_ENODEV;
	// state earlier setters asked for (resets, phases) first, so that this one has the last word
	int rc = apply_pending(b, b->compute);
	if (rc != WR_OK)
		return rc;
	WR_CUDA(cudaStreamSynchronize(b->compute));
	if (want == 0)
		return WR_OK;
	void *dst = stage ? (void*)(b->d_demod[b->cur] + (size_t)rx * b->dstride)
			: (void*)(b->d_hist1[b->cur] + (size_t)rx * (b->n1 - 1));
	WR_CUDA(cudaMemcpy(dst, in, sizeof(float) * want, cudaMemcpyHostToDevice));
	return WR_OK;
}

int wr_bank_process_u8(wr_bank *b, const uint8_t *iq_host, unsigned nframes, float *audio_host, size_t audio_stride)
{
	return process_any(b, iq_host, true, nframes, audio_host, audio_stride);
}

static int run_device_steps_any(wr_bank *b, const void *const *iq_dev, bool u8, unsigned n_iq, size_t stream_stride_frames,
		unsigned nframes, float *const *audio_dev, unsigned n_audio, size_t audio_stride,
		unsigned first, unsigned steps)
{
	WR_REQUIRE(b && iq_dev && audio_dev && n_iq && n_audio, WR_EINVAL, "wr_bank_run_device_steps: bad argument");
	for (unsigned i = 0; i < steps; i++) {
		int rc = process_device_any(b, iq_dev[(first + i) % n_iq], u8, stream_stride_frames, nframes,
				audio_dev[(first + i) % n_audio], audio_stride, nullptr);
		if (rc != WR_OK)
			return rc;
	}
	return WR_OK;
}

int wr_bank_run_device_steps(wr_bank *b, const float *const *iq_dev, unsigned n_iq, size_t stream_stride_frames,
		unsigned nframes, float *const *audio_dev, unsigned n_audio, size_t audio_stride,
		unsigned first, unsigned steps)
{
	return run_device_steps_any(b, reinterpret_cast<const void *const *>(iq_dev), false, n_iq, stream_stride_frames,
			nframes, audio_dev, n_audio, audio_stride, first, steps);
}

int wr_bank_run_device_steps_u8(wr_bank *b, const uint8_t *const *iq_dev, unsigned n_iq, size_t stream_stride_frames,
		unsigned nframes, float *const *audio_dev, unsigned n_audio, size_t audio_stride,
		unsigned first, unsigned steps)
{
	return run_device_steps_any(b, reinterpret_cast<const void *const *>(iq_dev), true, n_iq, stream_stride_frames,
			nframes, audio_dev, n_audio, audio_stride, first, steps);
}

static int run_host_steps_any(wr_bank *b, const void *const *iq_pinned, bool u8, unsigned n_iq, unsigned nframes,
		float *const *audio_pinned, unsigned n_audio, size_t audio_stride,
		unsigned first, unsigned steps, int pipelined)
{
	WR_REQUIRE(b && iq_pinned && audio_pinned && n_iq && n_audio, WR_EINVAL, "wr_bank_run_host_steps: bad argument");
	WR_REQUIRE(b->inflight == 0, WR_ESTATE, "wr_bank_run_host_steps: blocks already in flight");
	const int depth = pipelined ? std::min<int>(b->depth, (int)n_audio) : 1;
	int rc;
	if (!pipelined) {
		// strictly one block per call: the synchronous entry point itself (wr_bank_process)
		for (unsigned i = 0; i < steps; i++)
			if ((rc = process_any(b, iq_pinned[(first + i) % n_iq], u8, nframes, audio_pinned[(first + i) % n_audio], audio_stride)) != WR_OK)
				return rc;
		return WR_OK;
	}
	for (unsigned i = 0; i < steps; i++) {
		if (b->inflight == depth && (rc = wr_bank_wait(b)) != WR_OK)
			return rc;
		rc = submit_any(b, iq_pinned[(first + i) % n_iq], u8, nframes, audio_pinned[(first + i) % n_audio], audio_stride);
		if (rc != WR_OK)
			return rc;
	}
	while (b->inflight > 0)
		if ((rc = wr_bank_wait(b)) != WR_OK)
			return rc;
	return WR_OK;
}

int wr_bank_run_host_steps(wr_bank *b, const float *const *iq_pinned, unsigned n_iq, unsigned nframes,
		float *const *audio_pinned, unsigned n_audio, size_t audio_stride,
		unsigned first, unsigned steps, int pipelined)
{
	return run_host_steps_any(b, reinterpret_cast<const void *const *>(iq_pinned), false, n_iq, nframes,
			audio_pinned, n_audio, audio_stride, first, steps, pipelined);
}

int wr_bank_run_host_steps_u8(wr_bank *b, const uint8_t *const *iq_pinned, unsigned n_iq, unsigned nframes,
		float *const *audio_pinned, unsigned n_audio, size_t audio_stride,
		unsigned first, unsigned steps, int pipelined)
{
	return run_host_steps_any(b, reinterpret_cast<const void *const *>(iq_pinned), true, n_iq, nframes,
			audio_pinned, n_audio, audio_stride, first, steps, pipelined);
}

void *wr_bank_stream(wr_bank *b) { return b ? (void*)b->compute : nullptr; }

int wr_bank_sync(wr_bank *b)
{
	WR_REQUIRE(b, WR_EINVAL, "wr_bank_sync: null bank");
	if (!wr::use_device(b->device))
		return WR_ENODEV;
	WR_CUDA(cudaStreamSynchronize(b->h2d));
	WR_CUDA(cudaStreamSynchronize(b->compute));
	WR_CUDA(cudaStreamSynchronize(b->d2h));
	return WR_OK;
}

int wr_bank_keep_channel(wr_bank *b, int keep)
{
	WR_REQUIRE(b, WR_EINVAL, "wr_bank_keep_channel: null bank");
	if (!wr::use_device(b->device))
		return WR_ENODEV;
	if (keep && !b->d_chan)
		WR_CUDA(cudaMalloc(&b->d_chan, 2 * sizeof(float2) * (size_t)b->R * std::max(1u, b->maxM1)));
	b->keepChan = keep != 0;
	return WR_OK;
}

long wr_bank_read_stage(wr_bank *b, unsigned rx, int stage, float *out, size_t cap)
{
	WR_REQUIRE(b && out && rx < b->R, WR_EINVAL, "wr_bank_read_stage: bad argument");
	if (!wr::use_device(b->device))
		return WR_ENODEV;
	WR_CUDA(cudaStreamSynchronize(b->compute));
	if (stage == WR_STAGE_DEMOD) {
		size_t n = std::min<size_t>(b->lastM1, cap);
		const float *src = b->d_demod[b->cur ^ 1] + (size_t)rx * b->dstride + (b->n2 - 1);
		WR_CUDA(cudaMemcpy(out, src, sizeof(float) * n, cudaMemcpyDeviceToHost));
		return (long)n;
	}
	if (stage == WR_STAGE_CHANNEL) {
		WR_REQUIRE(b->keepChan && b->d_chan, WR_ESTATE, "wr_bank_read_stage: channel stream not kept (wr_bank_keep_channel)");
		size_t n = std::min<size_t>((size_t)b->lastM1 * 2, cap);
		WR_CUDA(cudaMemcpy(out, b->d_chan + ((size_t)b->chanSide * b->R + rx) * std::max(1u, b->maxM1), sizeof(float) * n, cudaMemcpyDeviceToHost));
		return (long)n;
	}
	wr::set_error("wr_bank_read_stage: unknown stage %d", stage);
	return WR_EINVAL;
}

int wr_bank_set_audio_format(wr_bank *b, int format)
{
	WR_REQUIRE(b && (format == WR_AUDIO_FLOAT || format == WR_AUDIO_LAME), WR_EINVAL, "wr_bank_set_audio_format: bad format %d", format);
	b->outScale = format == WR_AUDIO_LAME ? 32768.0f : 1.0f;
	return WR_OK;
}

int wr_bank_set_variant(wr_bank *b, int variant)
{
	WR_REQUIRE(b && variant >= 0 && variant <= 4, WR_EINVAL, "wr_bank_set_variant: bad variant %d", variant);
	b->variant = variant;
	return WR_OK;
}

int wr_bank_variant_in_use(const wr_bank *b) { return b ? b->variantInUse : 0; }

unsigned long long wr_bank_launch_count(const wr_bank *b) { return b ? b->launches : 0; }

int wr_bank_set_timing(wr_bank *b, int on)
{
	WR_REQUIRE(b, WR_EINVAL, "wr_bank_set_timing: null bank");
	if (!wr::use_device(b->device))
		return WR_ENODEV;
	if (on && !b->tev[0][0])
		for (int i = 0; i < wr_bank::kTimeRing; i++)
			for (int j = 0; j < 3; j++)
				WR_CUDA(cudaEventCreate(&b->tev[i][j]));
	while (b->tCount > 0) {
		int rc = drain_one(b);
		if (rc != WR_OK)
			return rc;
	}
	b->timing = on != 0;
	b->accMs[0] = b->accMs[1] = 0.0;
	b->accBlocks = 0;
	return WR_OK;
}

int wr_bank_kernel_times(wr_bank *b, double *ms2_total, unsigned long long *nblocks)
{
	WR_REQUIRE(b && ms2_total && nblocks, WR_EINVAL, "wr_bank_kernel_times: null argument");
	if (!wr::use_device(b->device))
		return WR_ENODEV;
	while (b->tCount > 0) {
		int rc = drain_one(b);
		if (rc != WR_OK)
			return rc;
	}
	ms2_total[0] = b->accMs[0];
	ms2_total[1] = b->accMs[1];
	*nblocks = b->accBlocks;
	return WR_OK;
}

} // extern "C"
