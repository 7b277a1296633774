// wr_common.h -- host-side plumbing shared by the translation units of libwebradio_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include <string>

#include "webradio_b200.h"

namespace wr {

// Thread-local error string behind wr_last_error().
void set_error(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
const char *get_error();

// Validates `device` (a CUDA device must exist: there is no CPU fallback) and selects it.
bool check_device(int device);
// Makes `device` current for the calling thread (cheap when it already is).
bool use_device(int device);

} // namespace wr

// Map a CUDA runtime failure to WR_ECUDA without throwing (reference process() -> false).
#define WR_CUDA(call)                                                                   \
	do {                                                                                \
		cudaError_t wr_e_ = (call);                                                     \
		if (wr_e_ != cudaSuccess) {                                                     \
			wr::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(wr_e_),    \
					__FILE__, __LINE__);                                                \
			return WR_ECUDA;                                                            \
		}                                                                               \
	} while (0)

#define WR_CUDA_PTR(call)                                                               \
	do {                                                                                \
		cudaError_t wr_e_ = (call);                                                     \
		if (wr_e_ != cudaSuccess) {                                                     \
			wr::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(wr_e_),    \
					__FILE__, __LINE__);                                                \
			return nullptr;                                                             \
		}                                                                               \
	} while (0)

#define WR_REQUIRE(cond, code, ...)                                                     \
	do {                                                                                \
		if (!(cond)) {                                                                  \
			wr::set_error(__VA_ARGS__);                                                 \
			return (code);                                                              \
		}                                                                               \
	} while (0)
