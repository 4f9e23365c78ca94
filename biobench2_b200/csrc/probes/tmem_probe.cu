// tmem_probe.cu — is Tensor Memory a faster home for the SSV gain table than shared memory?
//
// ssv_kernel (ssv_msv.cu) is bound by shared-memory bandwidth: Q LDS.32 per DP row per warp against 128 B/clk/SM, with the
// integer pipe at ~84 %.  A table in TMEM would be read with one tcgen05.ld.32x32b.x8 per row (lane-private columns, the
// column offset = residue * 8 is a run-time operand).  This probe measures, on the device it runs on:
//   (1) tcgen05.ld.32x32b.x8 throughput with every SM sub-partition streaming;
//   (2) a DP-row mock-up — table fetch + Q packed add-max + Q/2 three-input max + one lane rotation — with the fetch from
//       shared memory (Q LDS.32, as ssv_kernel) and from TMEM (one tcgen05.ld), rows per second each.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tmem_probe probes/tmem_probe.cu ; prints one JSON line.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int Q = 7, NROWS = 30, TCOLS = 256;     // 30 table rows x 8 columns (Q padded to 8) = 240 of 256 allocated columns

__device__ __forceinline__ void tmem_alloc(uint32_t *smem_slot)
{
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "l"((uint64_t) __cvta_generic_to_shared(smem_slot)), "n"(TCOLS));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_free(uint32_t addr)
{
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(addr), "n"(TCOLS));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8])
{
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8])
{
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
}
// same, with a DP register threaded through as a fake in/out operand: the compiler may otherwise sink the (volatile, but
// data-independent) load below the arithmetic of the current rows, which would serialise fetch and compute
__device__ __forceinline__ void tmem_ld8_pinned(uint32_t taddr, uint32_t (&v)[8], uint32_t &pin)
{
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%9];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "+r"(pin) : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// every warp fills (warps 0-3) then reads its own quarter of the CTA's allocation; pattern = lane * 4096 + column
__device__ uint32_t setup_tmem(uint32_t *s_slot)
{
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(s_slot);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = *s_slot;
  const uint32_t mine = base + ((uint32_t) ((warp & 3) * 32) << 16);
  if (warp < 4) {
    for (int c = 0; c < TCOLS; c += 8) {
      uint32_t v[8];
#pragma unroll
      for (int q = 0; q < 8; q++) v[q] = (uint32_t) (((warp & 3) * 32 + lane) * 4096 + c + q);
      tmem_st8(mine + c, v);
    }
    tmem_wait_st();
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  return mine;
}

__global__ void __launch_bounds__(256) check_kernel(int *bad)
{
  __shared__ uint32_t s_slot;
  const uint32_t mine = setup_tmem(&s_slot);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = 0; c < TCOLS; c += 8) {
    uint32_t v[8];
    tmem_ld8(mine + c, v);
    tmem_wait_ld();
#pragma unroll
    for (int q = 0; q < 8; q++) if (v[q] != (uint32_t) (((warp & 3) * 32 + lane) * 4096 + c + q)) atomicAdd(bad, 1);
  }
  __syncthreads();
  if (warp == 0) tmem_free(s_slot);
}

// (1) the load alone: DEPTH loads in flight per wait
template <int DEPTH>
__global__ void __launch_bounds__(256) ldtm_stream_kernel(int iters, uint32_t *out)
{
  __shared__ uint32_t s_slot;
  const uint32_t mine = setup_tmem(&s_slot);
  uint32_t acc = 0, col = (threadIdx.x >> 5) * 8;
  for (int i = 0; i < iters; i++) {
    uint32_t v[DEPTH][8];
#pragma unroll
    for (int d = 0; d < DEPTH; d++) { tmem_ld8(mine + ((col + 8 * d) & (TCOLS - 8)), v[d]); }
    tmem_wait_ld();
#pragma unroll
    for (int d = 0; d < DEPTH; d++)
#pragma unroll
      for (int q = 0; q < 8; q++) acc ^= v[d][q];
    col = (col + 8 * DEPTH + (acc & 8)) & (TCOLS - 8);
  }
  if (acc == 0x12345678u) out[0] = acc;
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) tmem_free(s_slot);
}

// (2) DP-row mock-up.  The table row ("residue") sequence is a small LCG, warp-uniform as in the real kernel.
// R = DP rows per tcgen05.wait::ld (the wait covers every load in flight, so deeper prefetch does not help; more rows per wait does).
__device__ __forceinline__ uint32_t next_x(uint32_t x) { x += 7u; return x >= NROWS ? x - NROWS : x; }

template <int R>
__device__ __forceinline__ void dp_rows(uint32_t (&v)[Q], uint32_t &xE, const uint32_t (&g)[R][8], int prevl, uint32_t rmul)
{
#pragma unroll
  for (int r = 0; r < R; r++) {
    const uint32_t carry = __shfl_sync(0xffffffffu, v[Q - 1], prevl) * rmul;
#pragma unroll
    for (int q = Q - 1; q >= 1; q--) v[q] = __viaddmax_s16x2_relu(v[q - 1], g[r][q], g[r][q]);
    v[0] = __viaddmax_s16x2_relu(carry, g[r][0], g[r][0]);
    xE = __vimax3_s16x2(xE, v[0], v[1]); xE = __vimax3_s16x2(xE, v[2], v[3]); xE = __vimax3_s16x2(xE, v[4], v[5]); xE = __vmaxs2(xE, v[6]);
  }
}

template <bool TMEM, int R, int THREADS>
__global__ void __launch_bounds__(THREADS) row_kernel(int rows, uint32_t *out)
{
  __shared__ uint32_t s_slot;
  __shared__ uint32_t s_tab[NROWS * 8 * 32];                       // [row][q][lane]: conflict-free LDS.32
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < NROWS * 8 * 32; i += blockDim.x) s_tab[i] = 0xfffdfffeu + (i & 3);
  uint32_t mine = 0;
  if (TMEM) mine = setup_tmem(&s_slot); else __syncthreads();
  uint32_t v[Q], xE = 0u;
#pragma unroll
  for (int q = 0; q < Q; q++) v[q] = threadIdx.x + q;
  const int prevl = (lane + 31) & 31;
  uint32_t rmul = lane == 0 ? 0x10000u : 1u;
  asm("" : "+r"(rmul));
  uint32_t x = (threadIdx.x >> 5) % NROWS;
  uint32_t ga[R][8], gb[R][8];
  auto fetch = [&](uint32_t (&g)[R][8]) {
#pragma unroll
    for (int r = 0; r < R; r++) {
      if (TMEM) tmem_ld8_pinned(mine + x * 8, g[r], v[0]);
      else {
        const uint32_t *rowp = s_tab + x * (8 * 32) + lane;
#pragma unroll
        for (int q = 0; q < Q; q++) g[r][q] = rowp[q * 32];
      }
      x = next_x(x);
    }
  };
  fetch(ga);
  if (TMEM) tmem_wait_ld();
  for (int r0 = 0; r0 < rows; r0 += 2 * R) {
    fetch(gb);                                                     // in flight under group A's arithmetic
    dp_rows<R>(v, xE, ga, prevl, rmul);
    if (TMEM) tmem_wait_ld();
    fetch(ga);
    dp_rows<R>(v, xE, gb, prevl, rmul);
    if (TMEM) tmem_wait_ld();
  }
  if (xE == 0x12345678u) out[0] = xE;
  if (TMEM) { __syncthreads(); if ((threadIdx.x >> 5) == 0) tmem_free(s_slot); }
}


// (3) issue rates of the packed integer instructions the filters are made of, dependent-free chains, 16 warps per SM
template <int KIND>
__global__ void __launch_bounds__(256) dpx_rate_kernel(uint32_t g, int iters, uint32_t *out)
{
  uint32_t v[8];
#pragma unroll
  for (int q = 0; q < 8; q++) v[q] = threadIdx.x * 3 + q;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
      if (KIND == 0) v[q] = __viaddmax_s16x2(v[q], g, 0u);
      if (KIND == 1) v[q] = __vimax3_s16x2(v[q], g, v[(q + 1) & 7]);
      if (KIND == 2) v[q] = __vmaxs2(v[q], v[(q + 3) & 7] ^ g);
      if (KIND == 3) v[q] = __viaddmax_s16x2_relu(v[q], g, g);
      if (KIND == 4) v[q] = __byte_perm(v[q], g, 0x2103);
      if (KIND == 5) v[q] = (uint32_t) max((int) v[q] + (int) g, (int) v[(q + 1) & 7]);
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int q = 0; q < 8; q++) acc ^= v[q];
  if (acc == 0x12345678u) out[0] = acc;
}

int main()
{
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  const int sms = prop.multiProcessorCount;
  uint32_t *d_out; int *d_bad; CK(cudaMalloc(&d_out, 64)); CK(cudaMalloc(&d_bad, 4)); CK(cudaMemset(d_bad, 0, 4));
  check_kernel<<<sms * 2, 256>>>(d_bad);
  CK(cudaDeviceSynchronize());
  int bad = -1; CK(cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto timeit = [&](auto launch) -> double {
    double best = 1e30;
    for (int rep = 0; rep < 4; rep++) {
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    return best;
  };
  const int blocks = sms * 2, iters = 4096, rows = 32768;         // 2 CTAs x 256 columns = the SM's 512 TMEM columns; 16 warps per SM
  const double ms1 = timeit([&] { ldtm_stream_kernel<1><<<blocks, 256>>>(iters, d_out); });
  const double ms2 = timeit([&] { ldtm_stream_kernel<2><<<blocks, 256>>>(iters / 2, d_out); });
  const double ms4 = timeit([&] { ldtm_stream_kernel<4><<<blocks, 256>>>(iters / 4, d_out); });
  CK(cudaGetLastError());
  const double bytes = (double) blocks * 8 * iters * 1024.0;       // one x8 load = 32 lanes x 8 x 4 bytes
  struct Row { const char *name; double ms; int warps; };
  Row rw[6];
  rw[0] = {"smem_r1_256", timeit([&] { row_kernel<false, 1, 256><<<sms * 2, 256>>>(rows, d_out); }), 16};
  rw[1] = {"smem_r1_512", timeit([&] { row_kernel<false, 1, 512><<<sms * 2, 512>>>(rows, d_out); }), 32};
  rw[2] = {"tmem_r1_256", timeit([&] { row_kernel<true, 1, 256><<<sms * 2, 256>>>(rows, d_out); }), 16};
  rw[3] = {"tmem_r2_256", timeit([&] { row_kernel<true, 2, 256><<<sms * 2, 256>>>(rows, d_out); }), 16};
  rw[4] = {"tmem_r1_512", timeit([&] { row_kernel<true, 1, 512><<<sms * 2, 512>>>(rows, d_out); }), 32};
  rw[5] = {"tmem_r2_512", timeit([&] { row_kernel<true, 2, 512><<<sms * 2, 512>>>(rows, d_out); }), 32};
  CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
  double dpx[6];
  {
    const int it2 = 8192;
    dpx[0] = timeit([&] { dpx_rate_kernel<0><<<sms * 2, 256>>>(0xfffdfffeu, it2, d_out); });
    dpx[1] = timeit([&] { dpx_rate_kernel<1><<<sms * 2, 256>>>(0xfffdfffeu, it2, d_out); });
    dpx[2] = timeit([&] { dpx_rate_kernel<2><<<sms * 2, 256>>>(0xfffdfffeu, it2, d_out); });
    dpx[3] = timeit([&] { dpx_rate_kernel<3><<<sms * 2, 256>>>(0xfffdfffeu, it2, d_out); });
    dpx[4] = timeit([&] { dpx_rate_kernel<4><<<sms * 2, 256>>>(0xfffdfffeu, it2, d_out); });
    dpx[5] = timeit([&] { dpx_rate_kernel<5><<<sms * 2, 256>>>(0xfffdfffeu, it2, d_out); });
    CK(cudaGetLastError());
    for (int k = 0; k < 6; k++) dpx[k] = (double) sms * 16 * it2 * 8 / (dpx[k] * 1e-3);      // warp instructions per second
  }
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
  printf("{\"device\": \"%s\", \"sms\": %d, \"readback_mismatches\": %d, "
         "\"ldtm_x8_gbs\": {\"depth1\": %.1f, \"depth2\": %.1f, \"depth4\": %.1f}, \"ldtm_bytes_per_clk_per_sm_at_%d_mhz\": %.1f, "
         "\"row_mockup\": {\"q\": %d, \"cells_per_row\": %d",
         prop.name, sms, bad, bytes / (ms1 * 1e-3) / 1e9, bytes / (ms2 * 1e-3) / 1e9, bytes / (ms4 * 1e-3) / 1e9, clk / 1000,
         bytes / (ms4 * 1e-3) / ((double) clk * 1e3) / sms, Q, 64 * Q);
  for (int k = 0; k < 6; k++) {
    const double wrows = (double) sms * rw[k].warps * rows;        // warp-rows
    printf(", \"%s\": {\"warps_per_sm\": %d, \"tcups_equiv\": %.2f, \"clk_per_row_per_subpartition\": %.1f}", rw[k].name, rw[k].warps,
           wrows * 64 * Q / (rw[k].ms * 1e-3) / 1e12, (double) clk * 1e3 * rw[k].ms * 1e-3 / (wrows / (sms * 4.0)));
  }
  printf("}, \"warp_instr_per_clk_per_subpartition\": {\"viaddmnmx_s16x2\": %.3f, \"vimnmx3_s16x2\": %.3f, \"vimnmx_s16x2\": %.3f, \"viaddmnmx_s16x2_relu\": %.3f, \"prmt\": %.3f, \"viaddmnmx_s32\": %.3f}}\n",
         dpx[0] / ((double) clk * 1e3 * sms * 4), dpx[1] / ((double) clk * 1e3 * sms * 4), dpx[2] / ((double) clk * 1e3 * sms * 4),
         dpx[3] / ((double) clk * 1e3 * sms * 4), dpx[4] / ((double) clk * 1e3 * sms * 4), dpx[5] / ((double) clk * 1e3 * sms * 4));
  return 0;
}
