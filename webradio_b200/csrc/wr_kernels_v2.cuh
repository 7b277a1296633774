// wr_kernels_v2.cuh -- placeholder until the shared-memory-table kernels land.
#pragma once

#include "wr_bank.cuh"
#include "wr_device.cuh"

namespace wrd {

struct V2Plan {
	bool ok = false;
	bool tableStale = true;
};

inline bool v2_supported(const V2Plan &p) { return p.ok; }
inline int v2_init(V2Plan &, int, unsigned, unsigned) { return 0; }
inline void v2_destroy(V2Plan &) {}
inline int v2_launch_chan(V2Plan &, ChanArgs &, unsigned, cudaStream_t, unsigned long long *) { return -1; }

} // namespace wrd
