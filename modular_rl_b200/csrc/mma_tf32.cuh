// Warp-level tensor-core pieces shared by the fused chain kernels: mma.sync m16n8k8 TF32 in split
// precision (3xTF32).  x = hi + lo with hi = rna_tf32(x), lo = x - hi (exact in fp32);
// D += lo.hi + hi.lo + hi.hi keeps FP32-class accuracy (north_star: 1e-5) at 1/3 of the TF32 rate.
#pragma once
#include <stdint.h>

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  // round-to-nearest TF32 by integer arithmetic on the sign-magnitude bits (cvt.rna.tf32 costs ~4
  // instructions on sm_100a): add half an ulp, then truncate.  hi must be exact for the subtraction;
  // for lo the tensor core's own truncation of the low 13 bits completes the rounding.  Rounding (not
  // truncating) hi matters: with a truncated hi the dropped lo.lo term is one-signed, a 2e-7 relative
  // bias that the CG solve amplifies (measured on the ill-conditioned golden case).
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi)) + 0x1000u;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
