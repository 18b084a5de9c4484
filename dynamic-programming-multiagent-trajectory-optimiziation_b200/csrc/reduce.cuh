// reduce.cuh -- warp-shuffle / block reductions shared by the solver kernels.
#pragma once
#include <cuda_runtime.h>

namespace scvx {

// ---- small helpers ------------------------------------------------------------------------------
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double wmin(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide reduction of NV values per thread; op: 0 = sum, 1 = min, 2 = max (per slot).  Result in red[0..NV).
template <int NV>
__device__ __forceinline__ void block_reduce(double* vals, const int* ops, double* red /* [8][NV] + [NV] */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double v = vals[i];
    v = (ops[i] == 0) ? wsum(v) : (ops[i] == 1 ? wmin(v) : wmax(v));
    if (lane == 0) red[(wid + 1) * NV + i] = v;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    const int i = threadIdx.x;
    double v = red[NV + i];
    for (int w = 1; w < nw; ++w) {
      const double o = red[(w + 1) * NV + i];
      v = (ops[i] == 0) ? v + o : (ops[i] == 1 ? fmin(v, o) : fmax(v, o));
    }
    red[i] = v;
  }
  __syncthreads();
}


}  // namespace scvx
