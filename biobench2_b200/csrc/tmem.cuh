// tmem.cuh — Tensor Memory (tcgen05) helpers for row-invariant, lane-private tables.
//
// TMEM is 128 lanes x 512 columns of 32 bits per SM; warp w of a CTA reads and writes lanes 32*(w%4) .. +31 only.  A table that
// every warp needs is therefore written once into each of the four quarters.  tcgen05.ld.32x32b.xN hands thread t of the warp N
// consecutive columns of TMEM lane 32*(w%4)+t; the address (lane << 16 | column) is warp-uniform.  Measured on B200
// (probes/tmem_probe.cu): 64 B/clk per SM sub-partition, i.e. twice the shared-memory bandwidth of the SM, off the LSU pipe.
#pragma once
#include <stdint.h>

namespace p7g {

template <int NX> struct Tmem;
template <> struct Tmem<1> {
  static __device__ __forceinline__ void st(uint32_t a, const uint32_t (&v)[1]) { asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" :: "r"(a), "r"(v[0])); }
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t (&v)[1]) { asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v[0]) : "r"(a)); }
};
template <> struct Tmem<2> {
  static __device__ __forceinline__ void st(uint32_t a, const uint32_t (&v)[2]) { asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" :: "r"(a), "r"(v[0]), "r"(v[1])); }
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t (&v)[2]) { asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(a)); }
};
template <> struct Tmem<4> {
  static __device__ __forceinline__ void st(uint32_t a, const uint32_t (&v)[4]) { asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" :: "r"(a), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3])); }
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t (&v)[4]) { asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(a)); }
};
template <> struct Tmem<8> {
  static __device__ __forceinline__ void st(uint32_t a, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(a), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])); }
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(a)); }
};

template <> struct Tmem<16> {
  static __device__ __forceinline__ void st(uint32_t a, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(a), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                    "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])); }
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(a)); }
};
template <> struct Tmem<32> {
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(a)); }
};

// one warp allocates NCOLS (a power of two >= 32) columns for the CTA and publishes the base address through shared memory
template <int NCOLS>
__device__ __forceinline__ uint32_t tmem_alloc_cta(uint32_t *s_slot)
{
  if ((threadIdx.x >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "l"((uint64_t) __cvta_generic_to_shared(s_slot)), "n"(NCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  return *s_slot;
}
template <int NCOLS>
__device__ __forceinline__ void tmem_free_cta(uint32_t base)
{
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "n"(NCOLS));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_publish()          // after the writers' tcgen05.wait::st: make the table visible to every warp
{
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
}

}  // namespace p7g
