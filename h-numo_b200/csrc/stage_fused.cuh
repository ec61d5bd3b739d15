// stage_fused.cuh -- optimised fused barotropic stage kernel (thread-per-line sum factorisation, compile-time
// polynomial order).  Placeholder until the simple variant is parity-green on the GPU.
#pragma once
#include "btp_kernels.cuh"

namespace hn {
inline void upload_fused_ops(const Ops&, int, int) {}
inline bool stage_fused_supported(const Solver&) { return false; }
inline int launch_stage_fused(Solver&, const StageArgs&) { return -1; }
}  // namespace hn
