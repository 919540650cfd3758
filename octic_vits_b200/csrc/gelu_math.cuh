// Packed fp32 pair arithmetic (sm_100 FFMA2 / FADD2 / FMUL2) and the erf-GELU built on it; shared by the D8 GELU
// kernels (pointwise.cu) and the GELU epilogues of the GEMM (gemm_sm100.cu).
// GELU maths: reference octic_vits/d8_layers.py:98-102 (nn.GELU, exact erf form), derivative octic_vits/d8_gelu.py:16-26.
#pragma once
#include <cuda_runtime.h>

namespace octic {

constexpr float kInvSqrt2 = 0.70710678118654752f;
constexpr float kInvSqrt2Pi = 0.39894228040143268f;

// ------------------------------------- packed fp32 pairs (FFMA2 / FADD2 / FMUL2) -------------------------------------
// sm_100 issues two fp32 operations per instruction on a register pair (PTX fma / add / sub / mul .f32x2).  The D8 GELU
// kernels are instruction-issue bound (ncu: issue slots 75-79 % active at 51-64 % of HBM peak), so two channels share
// every butterfly and polynomial instruction; MUFU (rcp, ex2) and |x| remain per element.  Same fp32 arithmetic.
struct f2 { unsigned long long v; };
__device__ __forceinline__ f2 mk2(float a, float b) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void un2(f2 p, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v)); }
__device__ __forceinline__ f2 sp2(float c) { return mk2(c, c); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r;
}
// A&S 7.1.26 pieces shared by value and derivative: ax = |x|, e = exp(-x^2/2), pe = (1 - erf(|x|/sqrt2)) = poly(t) t e
__device__ __forceinline__ void erfc_parts_2(f2 x, f2& ax, f2& e, f2& pe) {
  float x0, x1;
  un2(x, x0, x1);
  ax = mk2(fabsf(x0), fabsf(x1));
  const f2 den = fma2(ax, sp2(0.3275911f * kInvSqrt2), sp2(1.0f));
  float d0, d1;
  un2(den, d0, d1);
  const f2 t = mk2(__fdividef(1.0f, d0), __fdividef(1.0f, d1));      // MUFU.RCP x2
  f2 poly = fma2(sp2(1.061405429f), t, sp2(-1.453152027f));
  poly = fma2(poly, t, sp2(1.421413741f));
  poly = fma2(poly, t, sp2(-0.284496736f));
  poly = fma2(poly, t, sp2(0.254829592f));
  const f2 arg = mul2(mul2(x, x), sp2(-0.5f * 1.4426950408889634f));
  float a0, a1;
  un2(arg, a0, a1);
  e = mk2(exp2f(a0), exp2f(a1));                                       // MUFU.EX2 x2
  pe = mul2(mul2(poly, t), e);
}
// 2 * gelu(x) = x + |x| * erf(|x| / sqrt2)   (x erf(x/sqrt2) is even); the 0.5 goes into the R2I scale
__device__ __forceinline__ f2 gelu_x2_2(f2 x) {
  f2 ax, e, pe;
  erfc_parts_2(x, ax, e, pe);
  return fma2(ax, fma2(pe, sp2(-1.0f), sp2(1.0f)), x);
}
// gelu'(x) = Phi(x) + x phi(x),  Phi(x) = x >= 0 ? 1 - pe/2 : pe/2
__device__ __forceinline__ f2 gelu_grad_2(f2 x) {
  f2 ax, e, pe;
  erfc_parts_2(x, ax, e, pe);
  float x0, x1, p0, p1;
  un2(x, x0, x1);
  un2(pe, p0, p1);
  const f2 phi = mk2(x0 >= 0.f ? fmaf(p0, -0.5f, 1.0f) : 0.5f * p0, x1 >= 0.f ? fmaf(p1, -0.5f, 1.0f) : 0.5f * p1);
  return fma2(x, mul2(e, sp2(kInvSqrt2Pi)), phi);
}


}  // namespace octic
