// wr_upload.cu -- wr_upload_* and wr_host_* of include/webradio_b200.h: the tuner block goes to the
// device ONCE per DspBlock::run of its producer, straight from the producer's own std::vector
// (page-locked in place on first sight), in pieces, so that consumers start on the first piece while
// the rest is still on the wire.
#include "wr_common.h"
#include "wr_upload.cuh"

#include <algorithm>

namespace {

void free_upload(wr_upload *u)
{
	if (!u)
		return;
	cudaSetDevice(u->device);
	if (u->st)
		cudaStreamSynchronize(u->st);
	for (int b = 0; b < 2; b++) {
		if (u->readPending[b] && u->readDone[b])
			cudaEventSynchronize(u->readDone[b]);
		cudaFree(u->d_iq[b]);
		for (int j = 0; j < wr_upload::kPieces; j++)
			if (u->landed[b][j]) cudaEventDestroy(u->landed[b][j]);
		if (u->readDone[b]) cudaEventDestroy(u->readDone[b]);
	}
	for (int i = 0; i < wr_upload::kRegs; i++)
		if (u->reg[i].ptr)
			cudaHostUnregister(const_cast<void*>(u->reg[i].ptr));
	if (u->st)
		cudaStreamDestroy(u->st);
	cudaGetLastError();
	delete u;
}

} // namespace

extern "C" {

void *wr_host_alloc(size_t bytes)
{
	void *p = nullptr;
	if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
		wr::set_error("wr_host_alloc(%zu): %s", bytes, cudaGetErrorString(cudaGetLastError()));
		return nullptr;
	}
	return p;
}

void wr_host_free(void *p)
{
	if (p)
		cudaFreeHost(p);
	cudaGetLastError();
}

wr_upload *wr_upload_create(int device, size_t max_frames)
{
	if (max_frames == 0) {
		wr::set_error("wr_upload_create: max_frames is 0");
		return nullptr;
	}
	if (!wr::check_device(device))
		return nullptr;
	wr_upload *u = new wr_upload();
	u->device = device;
	u->maxFrames = max_frames;
	cudaError_t e = cudaStreamCreateWithFlags(&u->st, cudaStreamNonBlocking);
	for (int b = 0; b < 2 && e == cudaSuccess; b++) {
		e = cudaMalloc(&u->d_iq[b], sizeof(float) * 2 * max_frames);
		for (int j = 0; j < wr_upload::kPieces && e == cudaSuccess; j++)
			e = cudaEventCreateWithFlags(&u->landed[b][j], cudaEventDisableTiming);
		if (e == cudaSuccess)
			e = cudaEventCreateWithFlags(&u->readDone[b], cudaEventDisableTiming);
	}
	if (e != cudaSuccess) {
		wr::set_error("wr_upload_create: %s", cudaGetErrorString(e));
		free_upload(u);
		return nullptr;
	}
	return u;
}

void wr_upload_destroy(wr_upload *u) { free_upload(u); }

size_t wr_upload_capacity(const wr_upload *u) { return u ? u->maxFrames : 0; }
int wr_upload_device(const wr_upload *u) { return u ? u->device : -1; }

int wr_upload_begin(wr_upload *u, const float *iq_host, unsigned nframes)
{
	WR_REQUIRE(u && (iq_host || nframes == 0), WR_EINVAL, "wr_upload_begin: null argument");
	WR_REQUIRE(nframes <= u->maxFrames, WR_EINVAL, "wr_upload_begin: %u frames > capacity %zu", nframes, u->maxFrames);
	if (!wr::use_device(u->device))
		return WR_ENODEV;
	const size_t bytes = sizeof(float) * 2 * (size_t)nframes;
	// Page-lock the producer's buffer where it lies: a DspBlock keeps its output vector from block to
	// block (reference dspblock.cxx:177-184 resizes it only when the block length changes) or swaps a
	// few ring buffers through it (rtlsdrtuner.cxx:265-285), so each buffer is registered once.  If the
	// driver refuses, the copies below still work (staged by the driver, slower).
	if (nframes) {
		const char *lo = reinterpret_cast<const char*>(iq_host), *hi = lo + bytes;
		int hit = -1, lru = 0;
		for (int i = 0; i < wr_upload::kRegs; i++) {
			wr_upload::Reg &r = u->reg[i];
			if (r.ptr) {
				const char *rlo = reinterpret_cast<const char*>(r.ptr), *rhi = rlo + r.bytes;
				if (rlo == lo && rhi >= hi) {
					hit = i;
				} else if (rlo < hi && lo < rhi) {
					// overlaps a registered range without being it: that buffer is gone (freed, moved,
					// grown) -- its registration must go before the runtime sees a half-locked source
					cudaStreamSynchronize(u->st);
					cudaHostUnregister(const_cast<void*>(r.ptr));
					cudaGetLastError();
					r.ptr = nullptr;
					r.bytes = 0;
					r.used = 0;
				}
			}
			if (u->reg[i].used < u->reg[lru].used)
				lru = i;
		}
		if (hit < 0 && u->regFailures < 4) {
			bool again = false;
			for (int i = 0; i < wr_upload::kRegs; i++)
				again = again || (u->seen[i].ptr == iq_host && u->seen[i].bytes == bytes);
			if (!again) {
				u->seen[u->seenNext] = { iq_host, bytes };
				u->seenNext = (u->seenNext + 1) % wr_upload::kRegs;
			} else {
				wr_upload::Reg &r = u->reg[lru];
				if (r.ptr) {
					cudaStreamSynchronize(u->st);
					cudaHostUnregister(const_cast<void*>(r.ptr));
					cudaGetLastError();
					r.ptr = nullptr;
					r.bytes = 0;
				}
				if (cudaHostRegister(const_cast<float*>(iq_host), bytes, cudaHostRegisterPortable) == cudaSuccess) {
					r.ptr = iq_host;
					r.bytes = bytes;
					hit = lru;
				} else {
					cudaGetLastError();
					u->regFailures++;    // e.g. memory its owner page-locked already: fine, copied as it is
				}
			}
		}
		if (hit >= 0)
			u->reg[hit].used = ++u->regClock;
	}
	const int side = u->cur ^ 1;
	// an asynchronous reader (the spectrum kernel) may still be on this side from two blocks ago
	if (u->readPending[side]) {
		WR_CUDA(cudaStreamWaitEvent(u->st, u->readDone[side], 0));
		u->readPending[side] = false;
	}
	// two pieces from 512 KiB, four from 8 MiB (every piece costs runtime calls on both sides), whole
	// multiples of 256 frames
	unsigned np = bytes >= (8u << 20) ? 4u : bytes >= (512u << 10) ? 2u : 1u;
	unsigned per = np ? (nframes / np) & ~255u : 0;
	if (per == 0)
		np = 1;
	unsigned done = 0;
	for (unsigned j = 0; j < np; j++) {
		const unsigned nf = (j + 1 == np) ? nframes - done : per;
		if (nf)
			WR_CUDA(cudaMemcpyAsync(u->d_iq[side] + 2 * (size_t)done, iq_host + 2 * (size_t)done,
					sizeof(float) * 2 * (size_t)nf, cudaMemcpyHostToDevice, u->st));
		WR_CUDA(cudaEventRecord(u->landed[side][j], u->st));
		done += nf;
		u->pieceEnd[j] = done;
	}
	u->npieces = np;
	u->nframes = nframes;
	u->cur = side;
	return WR_OK;
}

int wr_upload_finish(wr_upload *u)
{
	WR_REQUIRE(u, WR_EINVAL, "wr_upload_finish: null handle");
	if (!wr::use_device(u->device))
		return WR_ENODEV;
	if (u->npieces)
		WR_CUDA(cudaEventSynchronize(u->landed[u->cur][u->npieces - 1]));
	return WR_OK;
}

} // extern "C"
