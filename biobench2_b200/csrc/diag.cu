// diag.cu — launch timing read-out and issue-peak microbenchmarks (roofline denominators for bench.py).
#include "p7gpu_internal.h"

namespace p7g {

// dependent-free chains of the packed DPX add-max the SSV/MSV kernels are made of
__global__ void __launch_bounds__(256) peak_int16x2_kernel(uint32_t g, int iters, uint32_t *out)
{
  uint32_t v[8];
#pragma unroll
  for (int q = 0; q < 8; q++) v[q] = threadIdx.x + q;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = __viaddmax_s16x2(v[q], g, 0u);
  }
  uint32_t acc = 0;
#pragma unroll
  for (int q = 0; q < 8; q++) acc ^= v[q];
  if (acc == 0x12345678u) out[0] = acc;
}

// fp32 multiply and add issued separately (the float kernels are compiled -fmad=false like the reference's SSE2)
__global__ void __launch_bounds__(256) peak_fp32_kernel(float a, float b, int iters, float *out)
{
  float v[8];
#pragma unroll
  for (int q = 0; q < 8; q++) v[q] = (float) (threadIdx.x + q);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = __fadd_rn(__fmul_rn(v[q], a), b);
  }
  float acc = 0;
#pragma unroll
  for (int q = 0; q < 8; q++) acc += v[q];
  if (acc == 1.2345f) out[0] = acc;
}

// conflict-free LDS.32 stream
__global__ void __launch_bounds__(256) peak_lds_kernel(int iters, uint32_t *out)
{
  __shared__ uint32_t s[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = i;
  __syncthreads();
  uint32_t acc = 0; int pos = threadIdx.x;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int q = 0; q < 8; q++) acc += s[(pos + q * 256) & 4095];
    pos = (pos + 32) & 4095;
  }
  if (acc == 0x12345678u) out[0] = acc;
}

static int fold_timing(p7gpu_ctx *ctx)
{
  P7G_CUDA(cudaStreamSynchronize(ctx->stream));
  for (auto &t : ctx->timed) {
    float ms = 0.f;
    P7G_CUDA(cudaEventElapsedTime(&ms, t.start, t.stop));
    ctx->kernel_ms[t.kind] += ms; ctx->kernel_launches[t.kind]++;
    cudaEventDestroy(t.start); cudaEventDestroy(t.stop);
  }
  ctx->timed.clear();
  return P7GPU_OK;
}

}  // namespace p7g
using namespace p7g;

extern "C" {

int p7gpu_ctx_set_timing(p7gpu_ctx *ctx, int on)
{
  if (!ctx) return P7GPU_EINVAL;
  P7G_CHECK(fold_timing(ctx));
  ctx->timing = on != 0;
  return P7GPU_OK;
}

int p7gpu_ctx_reset_timing(p7gpu_ctx *ctx)
{
  if (!ctx) return P7GPU_EINVAL;
  P7G_CHECK(fold_timing(ctx));
  for (int k = 0; k < P7GPU_NKERNELS; k++) { ctx->kernel_ms[k] = 0; ctx->kernel_launches[k] = 0; }
  return P7GPU_OK;
}

int p7gpu_ctx_kernel_time(p7gpu_ctx *ctx, int kind, double *ms_total, int64_t *nlaunches)
{
  if (!ctx || kind < 0 || kind >= P7GPU_NKERNELS) return P7GPU_EINVAL;
  P7G_CHECK(fold_timing(ctx));
  if (ms_total)  *ms_total  = ctx->kernel_ms[kind];
  if (nlaunches) *nlaunches = ctx->kernel_launches[kind];
  return P7GPU_OK;
}

int p7gpu_measure_peaks(p7gpu_ctx *ctx, double *int16x2_gops, double *fp32_gflops, double *lds_gbs)
{
  if (!ctx) return P7GPU_EINVAL;
  P7G_CUDA(cudaSetDevice(ctx->device));
  P7G_CHECK(ensure_scratch(ctx, 1024));
  const int blocks = ctx->sm_count * 8, threads = 256, iters = 4096;
  cudaEvent_t e0, e1; P7G_CUDA(cudaEventCreate(&e0)); P7G_CUDA(cudaEventCreate(&e1));
  double best[3] = {0, 0, 0};
  for (int which = 0; which < 3; which++)
    for (int rep = 0; rep < 4; rep++) {
      P7G_CUDA(cudaEventRecord(e0, ctx->stream));
      if (which == 0)      peak_int16x2_kernel<<<blocks, threads, 0, ctx->stream>>>(0xfffdfffeu, iters, (uint32_t *) ctx->d_scratch);
      else if (which == 1) peak_fp32_kernel<<<blocks, threads, 0, ctx->stream>>>(0.999f, 0.001f, iters, (float *) ctx->d_scratch);
      else                 peak_lds_kernel<<<blocks, threads, 0, ctx->stream>>>(iters, (uint32_t *) ctx->d_scratch);
      P7G_CUDA(cudaEventRecord(e1, ctx->stream));
      P7G_CUDA(cudaEventSynchronize(e1));
      P7G_CUDA(cudaGetLastError());
      float ms = 0; P7G_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      const double n = (double) blocks * threads * iters * 8;
      double rate = (which == 0) ? n * 2 / (ms * 1e-3) / 1e9          // 2 int16 element-ops per packed instruction
                  : (which == 1) ? n * 2 / (ms * 1e-3) / 1e9          // 1 mul + 1 add
                                 : n * 4 / (ms * 1e-3) / 1e9;         // bytes
      if (rep > 0 && rate > best[which]) best[which] = rate;
    }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (int16x2_gops) *int16x2_gops = best[0];
  if (fp32_gflops)  *fp32_gflops  = best[1];
  if (lds_gbs)      *lds_gbs      = best[2];
  return P7GPU_OK;
}

}  // extern "C"
