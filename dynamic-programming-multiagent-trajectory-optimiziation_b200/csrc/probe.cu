// probe.cu -- measurement helpers (not on the product path): an fp64 FMA throughput probe used by bench.py as the
// roofline denominator for the FP64-pipe-bound kernels (MEASURED_PEAKS.json carries HBM and bf16 peaks only).
#include "common.cuh"

namespace scvx {
// 8 independent DFMA chains per thread, `iters` x 8 x 2 flops each; result written so nothing is optimised away.
__global__ void __launch_bounds__(256) fp64_fma_probe_kernel(int iters, double seed, double* out) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
// write-only sweep over a buffer larger than L2 (bench.py uses it between timed steps)
__global__ void __launch_bounds__(256) l2_flush_kernel(double* buf, size_t n, double v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) buf[i] = v;
}
}  // namespace scvx
using namespace scvx;

extern "C" int scvx_probe_fp64(int blocks, int iters, double* out, double* flops, void* stream) {
  if (blocks <= 0 || iters <= 0 || !out) return bad_arg("blocks/iters/out");
  fp64_fma_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 0.5, out);
  SCVX_CHECK_LAUNCH("scvx_probe_fp64");
  if (flops) *flops = (double)blocks * 256.0 * (double)iters * 16.0;
  return SCVX_OK;
}
extern "C" int scvx_l2_flush(double* buf, unsigned long long n_doubles, void* stream) {
  if (!buf || n_doubles == 0) return bad_arg("buf");
  l2_flush_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(buf, (size_t)n_doubles, 1.0);
  SCVX_CHECK_LAUNCH("scvx_l2_flush");
  return SCVX_OK;
}
