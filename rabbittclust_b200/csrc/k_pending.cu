// k_pending.cu -- entry points declared in include/tclust.h whose kernels are not built yet.
// They fail loudly (no CPU fallback).
#include "tc_common.cuh"
extern "C" int tc_sketch_kssd(tc_ctx *, const uint8_t *, const uint64_t *, const uint32_t *, uint64_t, uint32_t, int, int,
                              const int32_t *, uint64_t *, uint64_t, uint32_t *)
{ tc_set_error("tc_sketch_kssd: not implemented"); return 2; }
extern "C" int tc_greedy_scan(tc_ctx *, uint32_t, uint32_t, const uint32_t *, uint32_t, uint32_t *, uint32_t *)
{ tc_set_error("tc_greedy_scan: not implemented"); return 2; }
extern "C" int tc_greedy_cluster(tc_ctx *, int, int, double, int32_t *, uint32_t *)
{ tc_set_error("tc_greedy_cluster: not implemented"); return 2; }
