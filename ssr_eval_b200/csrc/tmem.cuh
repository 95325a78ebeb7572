// tmem.cuh -- tensor memory (TMEM, 512 columns x 128 lanes x 32 bit per SM) used as a per-thread
// operand store.  With the 32x32b access shape thread i of warp w reads / writes lane 32 (w % 4) + i,
// N consecutive columns per instruction: private storage next to the register file that costs no
// registers and no LSU / shared-memory bandwidth (LDTM / STTM run on the tensor-memory datapath).
// The FFT kernels keep per-thread constants (twiddles) and per-thread sample rings there.
#pragma once
#include <stdint.h>

namespace ssr {

// Allocation: one warp allocates `COLS` columns (power of two >= 32) for the CTA and publishes the
// base address through shared memory; every thread then derives its warp's lane window.
template <int COLS>
__device__ __forceinline__ unsigned tmem_alloc(unsigned* slot_smem, int warp) {
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (unsigned)__cvta_generic_to_shared(slot_smem)),
                 "n"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  return *slot_smem + ((unsigned)((warp & 3) * 32) << 16);
}

// Call after a CTA-wide barrier that follows the last TMEM access of every warp.
template <int COLS>
__device__ __forceinline__ void tmem_free(const unsigned* slot_smem, int warp) {
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*slot_smem), "n"(COLS) : "memory");
}

__device__ __forceinline__ void tmem_ld16(unsigned addr, unsigned (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(addr)
      : "memory");
}
// Waits for every outstanding tcgen05.ld of this thread.  The registers are operands so that no use of
// them can be scheduled ahead of the wait.
__device__ __forceinline__ void tmem_wait_ld(unsigned (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
// Orders the uses of a second register group after a preceding tmem_wait_ld (no instruction).
__device__ __forceinline__ void tmem_pin(unsigned (&r)[16]) {
  asm volatile(""
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_st16(unsigned addr, const unsigned (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// four complex doubles <-> 16 columns
template <typename CD>
__device__ __forceinline__ void tmem_pack4(const CD* v, unsigned (&r)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r[4 * i + 0] = (unsigned)__double2loint(v[i].x);
    r[4 * i + 1] = (unsigned)__double2hiint(v[i].x);
    r[4 * i + 2] = (unsigned)__double2loint(v[i].y);
    r[4 * i + 3] = (unsigned)__double2hiint(v[i].y);
  }
}
template <typename CD>
__device__ __forceinline__ void tmem_unpack4(const unsigned (&r)[16], CD* v) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    v[i] = CD{__hiloint2double((int)r[4 * i + 1], (int)r[4 * i + 0]),
              __hiloint2double((int)r[4 * i + 3], (int)r[4 * i + 2])};
}

}  // namespace ssr
