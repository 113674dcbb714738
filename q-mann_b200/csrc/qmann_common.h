// qmann_common.h -- shared host-side helpers of libqmann_b200.so (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>

// ATTENTION_CONST_SCALE is a compile-time constant of the reference (MemN2N/define.h:67) that the
// per-layer entry cuda_dot_mat_vec_fwd_appx has no parameter for; the batched API takes it in
// qmann_config::const_scale.
#define QMANN_ATTENTION_CONST_SCALE (-3)

namespace qmann {

// reference error convention (lib/layer_cuda.h:13-22): print "[*E] CUDA : <fn> : <msg>", exit(code)
void check_cuda(const char *fn, cudaError_t code);
// every kernel launch of this library is counted (bench.py reports it as gpu_launches)
void count_launch(unsigned n = 1);

}  // namespace qmann
