// f32x2.cuh -- packed float32 pairs: sm_100's FADD2 / FMUL2 / FFMA2 (PTX add / mul / fma.rn.f32x2) apply one
// IEEE operation to both halves of a 64-bit register pair per instruction.  Each half is rounded exactly like
// the scalar instruction, so results are bit-identical to scalar code; what is saved is ISSUE SLOTS, which is
// what binds the float32 kernels here (K2: 57 % issue-slot utilisation at 29 % FMA-pipe utilisation).
#pragma once

namespace ssr {

__device__ __forceinline__ unsigned long long f2_pack(float2 a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ float2 f2_unpack(unsigned long long r) {
  float2 a;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r));
  return a;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(d);
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {  // a * b + c, one rounding per half
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)), "l"(f2_pack(c)));
  return f2_unpack(d);
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {  // a - b = fma(b, -1, a): exact product, one rounding
  return fma2(b, make_float2(-1.f, -1.f), a);
}

}  // namespace ssr
