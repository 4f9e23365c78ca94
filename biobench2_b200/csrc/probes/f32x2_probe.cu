// f32x2_probe.cu — issue rate of sm_100's packed fp32 instructions (fma.rn.f32x2 -> FFMA2, mul -> FMUL2, add -> FADD2) against the
// scalar ones, every SM sub-partition busy, ILP chains per warp as a template parameter.  Prints one JSON line.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/f32x2_probe probes/f32x2_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{ unsigned long long d; asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b)
{ unsigned long long d; asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b)
{ unsigned long long d; asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

template <int MODE, int ILP>
__global__ void __launch_bounds__(128) k(float *out, int iters, float seed)
{
  if (MODE == 0) {                       // scalar FFMA
    float v[ILP]; const float a = seed, b = seed * 0.5f;
#pragma unroll
    for (int c = 0; c < ILP; c++) v[c] = threadIdx.x + c;
    for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < ILP; c++) v[c] = fmaf(v[c], a, b);
    }
    float s = 0; for (int c = 0; c < ILP; c++) s += v[c];
    if (s == 12345.f) out[0] = s;
  } else {
    unsigned long long v[ILP];
    const unsigned long long a = ((unsigned long long) __float_as_uint(seed) << 32) | __float_as_uint(seed);
    const unsigned long long b = ((unsigned long long) __float_as_uint(seed * 0.5f) << 32) | __float_as_uint(seed * 0.25f);
#pragma unroll
    for (int c = 0; c < ILP; c++) v[c] = ((unsigned long long) (threadIdx.x + c) << 32) | 0x3f800000u;
    for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < ILP; c++) v[c] = MODE == 1 ? fma2(v[c], a, b) : (MODE == 2 ? mul2(v[c], a) : add2(v[c], b));
    }
    unsigned long long s = 0; for (int c = 0; c < ILP; c++) s ^= v[c];
    if (s == 12345ull) out[0] = (float) s;
  }
}

template <int MODE, int ILP>
static double rate(int sms, float *d)
{
  const int iters = 4000, blocks = sms * 8;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE, ILP><<<blocks, 128>>>(d, 100, 1.0001f);
  cudaEventRecord(a);
  k<MODE, ILP><<<blocks, 128>>>(d, iters, 1.0001f);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double warp_instr = (double) blocks * 4 * iters * 8 * ILP;
  return warp_instr / (ms * 1e-3);      // warp instructions per second, whole device
}

int main()
{
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  float *d; CK(cudaMalloc(&d, 64));
  int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  const double per = 1.0 / ((double) pr.multiProcessorCount * 4 * clk * 1e3);   // -> warp instr per clk per sub-partition
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d", pr.name, pr.multiProcessorCount, clk);
  printf(", \"ffma_ilp8\": %.3f, \"ffma_ilp2\": %.3f, \"ffma_ilp1\": %.3f", rate<0, 8>(pr.multiProcessorCount, d) * per, rate<0, 2>(pr.multiProcessorCount, d) * per, rate<0, 1>(pr.multiProcessorCount, d) * per);
  printf(", \"ffma2_ilp8\": %.3f, \"ffma2_ilp2\": %.3f, \"ffma2_ilp1\": %.3f", rate<1, 8>(pr.multiProcessorCount, d) * per, rate<1, 2>(pr.multiProcessorCount, d) * per, rate<1, 1>(pr.multiProcessorCount, d) * per);
  printf(", \"fmul2_ilp8\": %.3f, \"fadd2_ilp8\": %.3f", rate<2, 8>(pr.multiProcessorCount, d) * per, rate<3, 8>(pr.multiProcessorCount, d) * per);
  printf(", \"unit\": \"warp instructions per clock per SM sub-partition (8 warps resident per sub-partition)\"}\n");
  CK(cudaDeviceSynchronize());
  return 0;
}
