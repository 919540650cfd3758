// Octic / dense multi-head attention, forward and backward, for short sequences (N = 197..261 tokens).
//
// Reference: AttentionD8.forward (octic_vits/d8_layers.py:623-656) builds q,k,v [B,H,N,hd] by concatenating, per
// head, slices of the five irrep tensors -- [A1(c_h) | A2 | B1 | B2 | E row0 (2c_h) | E row1 (2c_h)] -- calls
// F.scaled_dot_product_attention, and scatters the result back into the 5-tuple.  Here the gather happens while
// K/V (or Q/dO) are staged into shared memory and the scatter happens in the output stores, straight from/to the
// packed [T, 3D] / [T, D] rows, so the cat/permute/contiguous copies of the reference do not exist.  The dense
// layout (deit/vit.py:36-50) is the same kernel with a trivial column map.
//
// One CTA per (batch, head): the whole K and V (<= 272 x 80 bf16 each) sit in shared memory; each warp owns 16 query
// rows and runs an online-softmax loop over 64-key chunks with mma.sync.m16n8k16 (bf16 in, fp32 accumulate).
// Backward is two passes without atomics: pass A (warps own key rows) -> dK, dV; pass B (warps own query rows) -> dQ.
// These mma.sync kernels are the fallback for the packed-input layout and for sequences outside the envelope of the
// tcgen05 kernels in attention_tc.cu, which serve the model's hot path.
#include "attention_common.cuh"
#include <stdlib.h>

namespace octic {

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 b = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&b);
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// B fragments of two adjacent 8-wide n-tiles for  C[m, n] += A[m, k] * M[n, k]   (M row-major [n][k] in smem):
// b[0],b[1] -> n-tile n0, b[2],b[3] -> n-tile n0 + 8, k-step k0 (16 wide)
template <int STR>
__device__ __forceinline__ void load_b_nk(uint32_t (&b)[4], const __nv_bfloat16* M, int n0, int k0, int lane) {
  const int mat = lane >> 3, row = lane & 7;
  ldsm_x4(b, M + (n0 + (mat >> 1) * 8 + row) * STR + k0 + (mat & 1) * 8);
}
// B fragments for  C[m, n] += A[m, k] * M[k, n]   (M row-major [k][n] in smem): n-tiles n0 and n0 + 8, k-step k0
template <int STR>
__device__ __forceinline__ void load_b_kn(uint32_t (&b)[4], const __nv_bfloat16* M, int k0, int n0, int lane) {
  const int mat = lane >> 3, row = lane & 7;
  ldsm_x4_t(b, M + (k0 + (mat & 1) * 8 + row) * STR + n0 + (mat >> 1) * 8);
}


__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// Stage rows [0, Npad) of tensor s of (b, h) into smem (zero rows >= N) with 4-byte cp.async: the gather of the
// head vector from the packed row happens here.  Thread -> fixed column pair, rows strided, so the column map is read
// once and the loop body is a pointer bump + one cp.async.  Caller must cp_async_wait_all() + __syncthreads().
template <int HD>
__device__ __forceinline__ void stage_rows(__nv_bfloat16* dst, const __nv_bfloat16* src_rows, long ld, int s, int N,
                                           int Npad, const int* cb, const int* sm) {
  constexpr int STR = HD + 8, NP = HD / 2;
  const int rows_per_iter = blockDim.x / NP;
  const int jp = threadIdx.x % NP, r0 = threadIdx.x / NP;
  if (r0 >= rows_per_iter) return;
  const int col = cb[jp] + s * sm[jp];
  const __nv_bfloat16* src = src_rows + static_cast<long>(r0) * ld + col;
  __nv_bfloat16* d = dst + r0 * STR + jp * 2;
  const long sstep = static_cast<long>(rows_per_iter) * ld;
  const int dstep = rows_per_iter * STR;
  for (int row = r0; row < Npad; row += rows_per_iter, src += sstep, d += dstep) {
    if (row < N) cp_async4(d, src);
    else *reinterpret_cast<uint32_t*>(d) = 0u;
  }
}

// A fragments (16 rows x HD) of tensor s for rows r_lo = row0 + lane/4 and r_lo + 8, read straight from global.
template <int HD>
__device__ __forceinline__ void load_a_rows(uint32_t (&a)[HD / 16][4], const __nv_bfloat16* src_rows, long ld, int s,
                                            int row0, int N, const int* cb, const int* sm, int lane) {
  const int r_lo = row0 + (lane >> 2), r_hi = r_lo + 8;
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks) {
    const int jp0 = ks * 8 + (lane & 3), jp1 = jp0 + 4;
    const int c0 = cb[jp0] + s * sm[jp0], c1 = cb[jp1] + s * sm[jp1];
    a[ks][0] = r_lo < N ? *reinterpret_cast<const uint32_t*>(src_rows + r_lo * ld + c0) : 0u;
    a[ks][1] = r_hi < N ? *reinterpret_cast<const uint32_t*>(src_rows + r_hi * ld + c0) : 0u;
    a[ks][2] = r_lo < N ? *reinterpret_cast<const uint32_t*>(src_rows + r_lo * ld + c1) : 0u;
    a[ks][3] = r_hi < N ? *reinterpret_cast<const uint32_t*>(src_rows + r_hi * ld + c1) : 0u;
  }
}

// ------------------------------------------------------ tiles ------------------------------------------------------
// ldmatrix addressing: every lane keeps a constant byte offset inside a [rows][STR] bf16 tile, the rest of the address
// is (row block, column block) arithmetic that is compile-time inside the unrolled loops.
__device__ __forceinline__ void ldsm_x4_addr(uint32_t (&r)[4], uint32_t a) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t_addr(uint32_t (&r)[4], uint32_t a) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
template <int STR>
__device__ __forceinline__ uint32_t lane_off_nk(int lane) {   // operand stored [n][k], B fragment of A * M^T
  const int mat = lane >> 3, row = lane & 7;
  return static_cast<uint32_t>((((mat >> 1) * 8 + row) * STR + (mat & 1) * 8) * 2);
}
template <int STR>
__device__ __forceinline__ uint32_t lane_off_kn(int lane) {   // operand stored [k][n], B fragment of A * M
  const int mat = lane >> 3, row = lane & 7;
  return static_cast<uint32_t>((((mat & 1) * 8 + row) * STR + (mat >> 1) * 8) * 2);
}

// acc[NT][4] += A(16 x HD, registers) * M[n0 .. n0 + 8 NT, :]^T       (M stored [n][HD] in smem; NT even)
template <int HD, int NT>
__device__ __forceinline__ void mma_a_mt(float (&acc)[NT][4], const uint32_t (&a)[HD / 16][4], uint32_t m_nk, int n0) {
  constexpr int STR = HD + 8;
  const uint32_t base = m_nk + static_cast<uint32_t>(n0) * (STR * 2);
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks) {
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      uint32_t bf[4];
      ldsm_x4_addr(bf, base + (np * 16 * STR + ks * 16) * 2);
      mma16816(acc[2 * np], a[ks], bf[0], bf[1]);
      mma16816(acc[2 * np + 1], a[ks], bf[2], bf[3]);
    }
  }
}
// acc[HD/8][4] += P(16 x 8 NT, bf16 fragments built from p[NT][4]) * M[k0 .. k0 + 8 NT, :]   (M stored [k][HD])
template <int HD, int NT>
__device__ __forceinline__ void mma_p_m(float (&acc)[HD / 8][4], const float (&p)[NT][4], uint32_t m_kn, int k0) {
  constexpr int STR = HD + 8;
  const uint32_t base = m_kn + static_cast<uint32_t>(k0) * (STR * 2);
#pragma unroll
  for (int kk = 0; kk < NT / 2; ++kk) {
    uint32_t pa[4];
    pa[0] = pack_bf16(p[2 * kk][0], p[2 * kk][1]);
    pa[1] = pack_bf16(p[2 * kk][2], p[2 * kk][3]);
    pa[2] = pack_bf16(p[2 * kk + 1][0], p[2 * kk + 1][1]);
    pa[3] = pack_bf16(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
    for (int dp = 0; dp < HD / 16; ++dp) {
      uint32_t bf[4];
      ldsm_x4_t_addr(bf, base + (kk * 16 * STR + dp * 16) * 2);
      mma16816(acc[2 * dp], pa, bf[0], bf[1]);
      mma16816(acc[2 * dp + 1], pa, bf[2], bf[3]);
    }
  }
}
template <int NT>
__device__ __forceinline__ void zero_acc(float (&a)[NT][4]) {
#pragma unroll
  for (int i = 0; i < NT; ++i) a[i][0] = a[i][1] = a[i][2] = a[i][3] = 0.f;
}

// ------------------------------------------------------ forward ------------------------------------------------------
struct SoftmaxState { float m_lo, m_hi, l_lo, l_hi; };

// one KV chunk of 8*NT keys starting at kv0; MASK: the chunk may contain keys >= N
template <int HD, int NT, bool MASK>
__device__ __forceinline__ void fwd_chunk(const uint32_t (&qa)[HD / 16][4], float (&oacc)[HD / 8][4], SoftmaxState& st,
                                          uint32_t k_nk, uint32_t v_kn, int kv0, int N, float scale_log2, int lane) {
  float s[NT][4];
  zero_acc<NT>(s);
  mma_a_mt<HD, NT>(s, qa, k_nk, kv0);
  float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    if (MASK) {
      const int col = kv0 + nt * 8 + (lane & 3) * 2;
      if (col >= N) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
      if (col + 1 >= N) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
    }
    mx_lo = fmaxf(mx_lo, fmaxf(s[nt][0], s[nt][1]));
    mx_hi = fmaxf(mx_hi, fmaxf(s[nt][2], s[nt][3]));
  }
  mx_lo = quad_max(mx_lo); mx_hi = quad_max(mx_hi);
  const float mn_lo = fmaxf(st.m_lo, mx_lo), mn_hi = fmaxf(st.m_hi, mx_hi);
  const float corr_lo = exp2f((st.m_lo - mn_lo) * scale_log2), corr_hi = exp2f((st.m_hi - mn_hi) * scale_log2);
  st.m_lo = mn_lo; st.m_hi = mn_hi;
  const float off_lo = mn_lo * scale_log2, off_hi = mn_hi * scale_log2;
  float rs_lo = 0.f, rs_hi = 0.f;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    s[nt][0] = exp2f(fmaf(s[nt][0], scale_log2, -off_lo));
    s[nt][1] = exp2f(fmaf(s[nt][1], scale_log2, -off_lo));
    s[nt][2] = exp2f(fmaf(s[nt][2], scale_log2, -off_hi));
    s[nt][3] = exp2f(fmaf(s[nt][3], scale_log2, -off_hi));
    rs_lo += s[nt][0] + s[nt][1];
    rs_hi += s[nt][2] + s[nt][3];
  }
  st.l_lo = st.l_lo * corr_lo + rs_lo;
  st.l_hi = st.l_hi * corr_hi + rs_hi;
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    oacc[i][0] *= corr_lo; oacc[i][1] *= corr_lo; oacc[i][2] *= corr_hi; oacc[i][3] *= corr_hi;
  }
  mma_p_m<HD, NT>(oacc, s, v_kn, kv0);
}

template <int HD>
__global__ void __launch_bounds__(192, 2) attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ o,
                                                          float* __restrict__ lse, int N, int H, HeadMap m, float scale_log2) {
  constexpr int STR = HD + 8, KS = HD / 16, DT = HD / 8;
  extern __shared__ __align__(16) uint8_t smem[];
  const int Npad = (N + 15) & ~15;
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* Vs = Ks + Npad * STR;
  int* cb = reinterpret_cast<int*>(Vs + Npad * STR);
  int* sm = cb + HD / 2;
  int* ocb = sm + HD / 2;

  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const long ld3 = 3L * m.D;
  const __nv_bfloat16* rows = qkv + static_cast<long>(b) * N * ld3;

  if (threadIdx.x < HD / 2) {
    int base, smul;
    qkv_col(m, h, threadIdx.x * 2, base, smul);
    cb[threadIdx.x] = base; sm[threadIdx.x] = smul;
    ocb[threadIdx.x] = o_col(m, h, threadIdx.x * 2);
  }
  __syncthreads();
  stage_rows<HD>(Ks, rows, ld3, 1, N, Npad, cb, sm);
  stage_rows<HD>(Vs, rows, ld3, 2, N, Npad, cb, sm);
  cp_async_wait_all();
  __syncthreads();

  const uint32_t k_nk = static_cast<uint32_t>(__cvta_generic_to_shared(Ks)) + lane_off_nk<STR>(lane);
  const uint32_t v_kn = static_cast<uint32_t>(__cvta_generic_to_shared(Vs)) + lane_off_kn<STR>(lane);
  const int ntiles = Npad / 16;
  const int full_end = (N / 64) * 64;      // keys [0, full_end) need no masking
  for (int qt = warp; qt < ntiles; qt += nwarps) {
    const int q0 = qt * 16;
    uint32_t qa[KS][4];
    load_a_rows<HD>(qa, rows, ld3, 0, q0, N, cb, sm, lane);
    float oacc[DT][4];
    zero_acc<DT>(oacc);
    SoftmaxState st{-INFINITY, -INFINITY, 0.f, 0.f};
    int kv0 = 0;
    for (; kv0 < full_end; kv0 += 64) fwd_chunk<HD, 8, false>(qa, oacc, st, k_nk, v_kn, kv0, N, scale_log2, lane);
    switch ((Npad - kv0) >> 3) {
      case 2: fwd_chunk<HD, 2, true>(qa, oacc, st, k_nk, v_kn, kv0, N, scale_log2, lane); break;
      case 4: fwd_chunk<HD, 4, true>(qa, oacc, st, k_nk, v_kn, kv0, N, scale_log2, lane); break;
      case 6: fwd_chunk<HD, 6, true>(qa, oacc, st, k_nk, v_kn, kv0, N, scale_log2, lane); break;
      case 8: fwd_chunk<HD, 8, true>(qa, oacc, st, k_nk, v_kn, kv0, N, scale_log2, lane); break;
      default: break;
    }
    const float l_lo = quad_sum(st.l_lo), l_hi = quad_sum(st.l_hi);
    const float inv_lo = 1.0f / l_lo, inv_hi = 1.0f / l_hi;
    const int r_lo = q0 + (lane >> 2), r_hi = r_lo + 8;
    __nv_bfloat16* orow_lo = o + (static_cast<long>(b) * N + r_lo) * m.D;
    __nv_bfloat16* orow_hi = o + (static_cast<long>(b) * N + r_hi) * m.D;
#pragma unroll
    for (int dt = 0; dt < DT; ++dt) {
      const int oc = ocb[dt * 4 + (lane & 3)];
      if (r_lo < N) *reinterpret_cast<uint32_t*>(orow_lo + oc) = pack_bf16(oacc[dt][0] * inv_lo, oacc[dt][1] * inv_lo);
      if (r_hi < N) *reinterpret_cast<uint32_t*>(orow_hi + oc) = pack_bf16(oacc[dt][2] * inv_hi, oacc[dt][3] * inv_hi);
    }
    if (lse != nullptr && (lane & 3) == 0) {
      float* l = lse + (static_cast<long>(b) * H + h) * N;
      if (r_lo < N) l[r_lo] = (st.m_lo * scale_log2 + log2f(l_lo)) * kLn2;
      if (r_hi < N) l[r_hi] = (st.m_hi * scale_log2 + log2f(l_hi)) * kLn2;
    }
  }
}

// ------------------------------------------- backward: delta = rowsum(dO * O) -------------------------------------------
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ o,
                                                         const __nv_bfloat16* __restrict__ d_o, float* __restrict__ delta,
                                                         int B, int N, int H, HeadMap m, int do_head_major) {
  // one thread per (b, n, h); delta laid out [B, H, N]; o follows the map m, dO either m or head-major [H][hd]
  const long total = static_cast<long>(B) * N * H;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int h = static_cast<int>(idx % H);
    const long t = idx / H;
    const int b = static_cast<int>(t / N), n = static_cast<int>(t - static_cast<long>(b) * N);
    const __nv_bfloat16* orow = o + t * m.D;
    const __nv_bfloat16* drow = d_o + t * m.D;
    float acc = 0.f;
    for (int j = 0; j < m.hd; j += 2) {
      const int c = o_col(m, h, j);
      const int cd = do_head_major ? h * m.hd + j : c;
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(orow + c));
      const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(drow + cd));
      acc += a.x * g.x + a.y * g.y;
    }
    delta[(static_cast<long>(b) * H + h) * N + n] = acc;
  }
}

// Same quantity, one WARP per token row when H = 16: lane = (head, half of the head vector).  The head-major dO half
// is 16-byte loads, the o half is 4-byte loads at columns computed once per thread; the two halves meet in one
// shuffle.  (The per-(token, head) thread version above spends its time in the column map and reaches 2.2 TB/s.)
template <int HD>
__global__ void __launch_bounds__(256) attn_delta_h16_kernel(const __nv_bfloat16* __restrict__ o,
                                                             const __nv_bfloat16* __restrict__ d_o, float* __restrict__ delta,
                                                             long T, int N, HeadMap m, int do_head_major) {
  constexpr int NP = HD / 4;                          // bf16 pairs per lane (half a head vector)
  const int lane = threadIdx.x & 31, h = lane >> 1, half = lane & 1;
  int ocol[NP];
#pragma unroll
  for (int k = 0; k < NP; ++k) ocol[k] = o_col(m, h, half * (HD / 2) + 2 * k);
  const long warp0 = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const long nwarps = (gridDim.x * static_cast<long>(blockDim.x)) >> 5;
  for (long t = warp0; t < T; t += nwarps) {
    const __nv_bfloat16* orow = o + t * m.D;
    const __nv_bfloat16* drow = d_o + t * m.D;
    uint32_t ov[NP], dv[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) ov[k] = *reinterpret_cast<const uint32_t*>(orow + ocol[k]);
    if (do_head_major && (NP % 4) == 0) {
      const uint4* dp = reinterpret_cast<const uint4*>(drow + h * HD + half * (HD / 2));
#pragma unroll
      for (int k = 0; k < NP / 4; ++k) {
        const uint4 v = dp[k];
        dv[4 * k] = v.x; dv[4 * k + 1] = v.y; dv[4 * k + 2] = v.z; dv[4 * k + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < NP; ++k)
        dv[k] = *reinterpret_cast<const uint32_t*>(drow + (do_head_major ? h * HD + half * (HD / 2) + 2 * k : ocol[k]));
    }
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ov[k]));
      const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dv[k]));
      acc = fmaf(a.x, g.x, fmaf(a.y, g.y, acc));
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (half == 0) {
      const long b = t / N;
      delta[(b * 16 + h) * N + (t - b * N)] = acc;
    }
  }
}

static void launch_attn_delta(const __nv_bfloat16* o, const __nv_bfloat16* d_o, float* delta, int B, int N, int H,
                              const HeadMap& m, int do_head_major, cudaStream_t s) {
  const long T = static_cast<long>(B) * N;
  if (H == 16 && (m.hd == 80 || m.hd == 64 || m.hd == 128 || m.hd == 96 || m.hd == 32) &&
      (reinterpret_cast<uintptr_t>(d_o) & 15) == 0 && (m.D % 8) == 0) {
    long blocks = (T + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    const int g = static_cast<int>(blocks);
    switch (m.hd) {
      case 80: attn_delta_h16_kernel<80><<<g, 256, 0, s>>>(o, d_o, delta, T, N, m, do_head_major); break;
      case 64: attn_delta_h16_kernel<64><<<g, 256, 0, s>>>(o, d_o, delta, T, N, m, do_head_major); break;
      case 128: attn_delta_h16_kernel<128><<<g, 256, 0, s>>>(o, d_o, delta, T, N, m, do_head_major); break;
      case 96: attn_delta_h16_kernel<96><<<g, 256, 0, s>>>(o, d_o, delta, T, N, m, do_head_major); break;
      default: attn_delta_h16_kernel<32><<<g, 256, 0, s>>>(o, d_o, delta, T, N, m, do_head_major); break;
    }
    return;
  }
  const long total = T * H;
  int dgrid = static_cast<int>((total + 255) / 256);
  if (dgrid > 148 * 16) dgrid = 148 * 16;
  attn_delta_kernel<<<dgrid, 256, 0, s>>>(o, d_o, delta, B, N, H, m, do_head_major);
}

// stage rows of an attention-output-shaped tensor (dO) for head h
template <int HD>
__device__ __forceinline__ void stage_o_rows(__nv_bfloat16* dst, const __nv_bfloat16* src_rows, long ld, int N, int Npad,
                                             const int* ocb) {
  constexpr int STR = HD + 8, NP = HD / 2;
  const int rows_per_iter = blockDim.x / NP;
  const int jp = threadIdx.x % NP, r0 = threadIdx.x / NP;
  if (r0 >= rows_per_iter) return;
  const __nv_bfloat16* src = src_rows + static_cast<long>(r0) * ld + ocb[jp];
  __nv_bfloat16* d = dst + r0 * STR + jp * 2;
  const long sstep = static_cast<long>(rows_per_iter) * ld;
  const int dstep = rows_per_iter * STR;
  for (int row = r0; row < Npad; row += rows_per_iter, src += sstep, d += dstep) {
    if (row < N) cp_async4(d, src);
    else *reinterpret_cast<uint32_t*>(d) = 0u;
  }
}
template <int HD>
__device__ __forceinline__ void load_a_o_rows(uint32_t (&a)[HD / 16][4], const __nv_bfloat16* src_rows, long ld, int row0,
                                              int N, const int* ocb, int lane) {
  const int r_lo = row0 + (lane >> 2), r_hi = r_lo + 8;
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks) {
    const int c0 = ocb[ks * 8 + (lane & 3)], c1 = ocb[ks * 8 + 4 + (lane & 3)];
    a[ks][0] = r_lo < N ? *reinterpret_cast<const uint32_t*>(src_rows + r_lo * ld + c0) : 0u;
    a[ks][1] = r_hi < N ? *reinterpret_cast<const uint32_t*>(src_rows + r_hi * ld + c0) : 0u;
    a[ks][2] = r_lo < N ? *reinterpret_cast<const uint32_t*>(src_rows + r_lo * ld + c1) : 0u;
    a[ks][3] = r_hi < N ? *reinterpret_cast<const uint32_t*>(src_rows + r_hi * ld + c1) : 0u;
  }
}

// ------------------------------------------- backward pass A: dK, dV -------------------------------------------
// Warps own 16 key rows.  S^T = K Q^T, P^T = exp(S^T*scale - lse[q]), dV += P^T dO, dP^T = V dO^T,
// dS^T = P^T (dP^T - delta[q]), dK += scale * dS^T Q.   Padded queries have lse = +inf, hence P^T = 0.
template <int HD, int NT>
__device__ __forceinline__ void bwd_kv_chunk(const uint32_t (&ka)[HD / 16][4], const uint32_t (&va)[HD / 16][4],
                                             float (&dk)[HD / 8][4], float (&dv)[HD / 8][4], uint32_t q_nk, uint32_t q_kn,
                                             uint32_t do_nk, uint32_t do_kn, const float* lse_s, const float* del_s, int q0,
                                             float scale_log2, int lane) {
  float st[NT][4], dp[NT][4];
  zero_acc<NT>(st);
  zero_acc<NT>(dp);
  mma_a_mt<HD, NT>(st, ka, q_nk, q0);
  mma_a_mt<HD, NT>(dp, va, do_nk, q0);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int col = q0 + nt * 8 + (lane & 3) * 2;
    const float2 l = *reinterpret_cast<const float2*>(lse_s + col);
    const float2 d = *reinterpret_cast<const float2*>(del_s + col);
    const float p0 = exp2f(fmaf(st[nt][0], scale_log2, -l.x)), p1 = exp2f(fmaf(st[nt][1], scale_log2, -l.y));
    const float p2 = exp2f(fmaf(st[nt][2], scale_log2, -l.x)), p3 = exp2f(fmaf(st[nt][3], scale_log2, -l.y));
    st[nt][0] = p0; st[nt][1] = p1; st[nt][2] = p2; st[nt][3] = p3;
    dp[nt][0] = p0 * (dp[nt][0] - d.x); dp[nt][1] = p1 * (dp[nt][1] - d.y);
    dp[nt][2] = p2 * (dp[nt][2] - d.x); dp[nt][3] = p3 * (dp[nt][3] - d.y);
  }
  mma_p_m<HD, NT>(dv, st, do_kn, q0);
  mma_p_m<HD, NT>(dk, dp, q_kn, q0);
}

template <int HD>
__global__ void __launch_bounds__(192, (HD <= 80) ? 2 : 1) attn_bwd_kv_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                          const __nv_bfloat16* __restrict__ d_o,
                                                          const float* __restrict__ lse, const float* __restrict__ delta,
                                                          __nv_bfloat16* __restrict__ dqkv, int N, int H, HeadMap m,
                                                          float scale, float scale_log2) {
  constexpr int STR = HD + 8, KS = HD / 16, DT = HD / 8;
  extern __shared__ __align__(16) uint8_t smem[];
  const int Npad = (N + 15) & ~15;
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* dOs = Qs + Npad * STR;
  float* lse_s = reinterpret_cast<float*>(dOs + Npad * STR);
  float* del_s = lse_s + Npad;
  int* cb = reinterpret_cast<int*>(del_s + Npad);
  int* sm = cb + HD / 2;
  int* ocb = sm + HD / 2;

  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const long ld3 = 3L * m.D;
  const __nv_bfloat16* rows = qkv + static_cast<long>(b) * N * ld3;
  const __nv_bfloat16* dorows = d_o + static_cast<long>(b) * N * m.D;
  __nv_bfloat16* drows = dqkv + static_cast<long>(b) * N * ld3;

  if (threadIdx.x < HD / 2) {
    int base, smul;
    qkv_col(m, h, threadIdx.x * 2, base, smul);
    cb[threadIdx.x] = base; sm[threadIdx.x] = smul;
    ocb[threadIdx.x] = o_col(m, h, threadIdx.x * 2);
  }
  for (int i = threadIdx.x; i < Npad; i += blockDim.x) {
    const long off = (static_cast<long>(b) * H + h) * N + i;
    lse_s[i] = i < N ? lse[off] * kLog2e : INFINITY;   // +inf -> P = 0 for padded queries
    del_s[i] = i < N ? delta[off] : 0.f;
  }
  __syncthreads();
  stage_rows<HD>(Qs, rows, ld3, 0, N, Npad, cb, sm);
  stage_o_rows<HD>(dOs, dorows, m.D, N, Npad, ocb);
  cp_async_wait_all();
  __syncthreads();

  const uint32_t qs = static_cast<uint32_t>(__cvta_generic_to_shared(Qs));
  const uint32_t dos = static_cast<uint32_t>(__cvta_generic_to_shared(dOs));
  const uint32_t q_nk = qs + lane_off_nk<STR>(lane), q_kn = qs + lane_off_kn<STR>(lane);
  const uint32_t do_nk = dos + lane_off_nk<STR>(lane), do_kn = dos + lane_off_kn<STR>(lane);
  const int ntiles = Npad / 16;
  const int full_end = (Npad / 32) * 32;
  for (int kt = warp; kt < ntiles; kt += nwarps) {
    const int k0 = kt * 16;
    uint32_t ka[KS][4], va[KS][4];
    load_a_rows<HD>(ka, rows, ld3, 1, k0, N, cb, sm, lane);
    load_a_rows<HD>(va, rows, ld3, 2, k0, N, cb, sm, lane);
    float dk[DT][4], dv[DT][4];
    zero_acc<DT>(dk);
    zero_acc<DT>(dv);
    int q0 = 0;
    for (; q0 < full_end; q0 += 32)
      bwd_kv_chunk<HD, 4>(ka, va, dk, dv, q_nk, q_kn, do_nk, do_kn, lse_s, del_s, q0, scale_log2, lane);
    if (q0 < Npad) bwd_kv_chunk<HD, 2>(ka, va, dk, dv, q_nk, q_kn, do_nk, do_kn, lse_s, del_s, q0, scale_log2, lane);

    const int r_lo = k0 + (lane >> 2), r_hi = r_lo + 8;
#pragma unroll
    for (int dt = 0; dt < DT; ++dt) {
      const int jp = dt * 4 + (lane & 3);
      const int ck = cb[jp] + sm[jp], cv = cb[jp] + 2 * sm[jp];
      if (r_lo < N) {
        *reinterpret_cast<uint32_t*>(drows + r_lo * ld3 + ck) = pack_bf16(dk[dt][0] * scale, dk[dt][1] * scale);
        *reinterpret_cast<uint32_t*>(drows + r_lo * ld3 + cv) = pack_bf16(dv[dt][0], dv[dt][1]);
      }
      if (r_hi < N) {
        *reinterpret_cast<uint32_t*>(drows + r_hi * ld3 + ck) = pack_bf16(dk[dt][2] * scale, dk[dt][3] * scale);
        *reinterpret_cast<uint32_t*>(drows + r_hi * ld3 + cv) = pack_bf16(dv[dt][2], dv[dt][3]);
      }
    }
  }
}

// ------------------------------------------- backward pass B: dQ -------------------------------------------
// Warps own 16 query rows.  S = Q K^T, P = exp(S*scale - lse), dP = dO V^T, dS = P (dP - delta), dQ = scale * dS K.
template <int HD, int NT, bool MASK>
__device__ __forceinline__ void bwd_q_chunk(const uint32_t (&qa)[HD / 16][4], const uint32_t (&doa)[HD / 16][4],
                                            float (&dq)[HD / 8][4], uint32_t k_nk, uint32_t k_kn, uint32_t v_nk, int kv0,
                                            int N, float lse_lo, float lse_hi, float del_lo, float del_hi,
                                            float scale_log2, int lane) {
  float s[NT][4], dp[NT][4];
  zero_acc<NT>(s);
  zero_acc<NT>(dp);
  mma_a_mt<HD, NT>(s, qa, k_nk, kv0);
  mma_a_mt<HD, NT>(dp, doa, v_nk, kv0);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    float p0 = exp2f(fmaf(s[nt][0], scale_log2, -lse_lo)), p1 = exp2f(fmaf(s[nt][1], scale_log2, -lse_lo));
    float p2 = exp2f(fmaf(s[nt][2], scale_log2, -lse_hi)), p3 = exp2f(fmaf(s[nt][3], scale_log2, -lse_hi));
    if (MASK) {
      const int col = kv0 + nt * 8 + (lane & 3) * 2;
      if (col >= N) { p0 = 0.f; p2 = 0.f; }
      if (col + 1 >= N) { p1 = 0.f; p3 = 0.f; }
    }
    dp[nt][0] = p0 * (dp[nt][0] - del_lo); dp[nt][1] = p1 * (dp[nt][1] - del_lo);
    dp[nt][2] = p2 * (dp[nt][2] - del_hi); dp[nt][3] = p3 * (dp[nt][3] - del_hi);
  }
  mma_p_m<HD, NT>(dq, dp, k_kn, kv0);
}

template <int HD>
__global__ void __launch_bounds__(192, (HD <= 80) ? 2 : 1) attn_bwd_q_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                         const __nv_bfloat16* __restrict__ d_o,
                                                         const float* __restrict__ lse, const float* __restrict__ delta,
                                                         __nv_bfloat16* __restrict__ dqkv, int N, int H, HeadMap m,
                                                         float scale, float scale_log2) {
  constexpr int STR = HD + 8, KS = HD / 16, DT = HD / 8;
  extern __shared__ __align__(16) uint8_t smem[];
  const int Npad = (N + 15) & ~15;
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* Vs = Ks + Npad * STR;
  int* cb = reinterpret_cast<int*>(Vs + Npad * STR);
  int* sm = cb + HD / 2;
  int* ocb = sm + HD / 2;

  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const long ld3 = 3L * m.D;
  const __nv_bfloat16* rows = qkv + static_cast<long>(b) * N * ld3;
  const __nv_bfloat16* dorows = d_o + static_cast<long>(b) * N * m.D;
  __nv_bfloat16* drows = dqkv + static_cast<long>(b) * N * ld3;

  if (threadIdx.x < HD / 2) {
    int base, smul;
    qkv_col(m, h, threadIdx.x * 2, base, smul);
    cb[threadIdx.x] = base; sm[threadIdx.x] = smul;
    ocb[threadIdx.x] = o_col(m, h, threadIdx.x * 2);
  }
  __syncthreads();
  stage_rows<HD>(Ks, rows, ld3, 1, N, Npad, cb, sm);
  stage_rows<HD>(Vs, rows, ld3, 2, N, Npad, cb, sm);
  cp_async_wait_all();
  __syncthreads();

  const uint32_t ks_ = static_cast<uint32_t>(__cvta_generic_to_shared(Ks));
  const uint32_t vs_ = static_cast<uint32_t>(__cvta_generic_to_shared(Vs));
  const uint32_t k_nk = ks_ + lane_off_nk<STR>(lane), k_kn = ks_ + lane_off_kn<STR>(lane);
  const uint32_t v_nk = vs_ + lane_off_nk<STR>(lane);
  const int ntiles = Npad / 16;
  const int full_end = (N / 32) * 32;
  for (int qt = warp; qt < ntiles; qt += nwarps) {
    const int q0 = qt * 16;
    uint32_t qa[KS][4], doa[KS][4];
    load_a_rows<HD>(qa, rows, ld3, 0, q0, N, cb, sm, lane);
    load_a_o_rows<HD>(doa, dorows, m.D, q0, N, ocb, lane);
    const int r_lo = q0 + (lane >> 2), r_hi = r_lo + 8;
    const long soff = (static_cast<long>(b) * H + h) * N;
    const float lse_lo = r_lo < N ? lse[soff + r_lo] * kLog2e : INFINITY;
    const float lse_hi = r_hi < N ? lse[soff + r_hi] * kLog2e : INFINITY;
    const float del_lo = r_lo < N ? delta[soff + r_lo] : 0.f;
    const float del_hi = r_hi < N ? delta[soff + r_hi] : 0.f;
    float dq[DT][4];
    zero_acc<DT>(dq);
    int kv0 = 0;
    for (; kv0 < full_end; kv0 += 32)
      bwd_q_chunk<HD, 4, false>(qa, doa, dq, k_nk, k_kn, v_nk, kv0, N, lse_lo, lse_hi, del_lo, del_hi, scale_log2, lane);
    if (Npad - kv0 >= 32)
      bwd_q_chunk<HD, 4, true>(qa, doa, dq, k_nk, k_kn, v_nk, kv0, N, lse_lo, lse_hi, del_lo, del_hi, scale_log2, lane);
    else if (Npad - kv0 == 16)
      bwd_q_chunk<HD, 2, true>(qa, doa, dq, k_nk, k_kn, v_nk, kv0, N, lse_lo, lse_hi, del_lo, del_hi, scale_log2, lane);
#pragma unroll
    for (int dt = 0; dt < DT; ++dt) {
      const int cq = cb[dt * 4 + (lane & 3)];
      if (r_lo < N) *reinterpret_cast<uint32_t*>(drows + r_lo * ld3 + cq) = pack_bf16(dq[dt][0] * scale, dq[dt][1] * scale);
      if (r_hi < N) *reinterpret_cast<uint32_t*>(drows + r_hi * ld3 + cq) = pack_bf16(dq[dt][2] * scale, dq[dt][3] * scale);
    }
  }
}

// ------------------------------------------------------ host ------------------------------------------------------
static int attn_warps(int N) {
  const int ntiles = (N + 15) / 16;
  const int rounds = (ntiles + 5) / 6;   // at most 6 warps (192 threads) per CTA: 2 CTAs/SM with <= 168 registers
  const int w = (ntiles + rounds - 1) / rounds;
  return w < 2 ? 2 : w;                  // the staging loops need blockDim >= hd/2 threads
}
static int make_head_map(HeadMap* m, int H, int hd, int octic) {
  if (H <= 0 || hd <= 0 || (hd % 16) != 0 || octic < 0 || octic > 2) return OCTIC_ERR_ARG;
  m->octic = octic; m->hd = hd; m->D = H * hd; m->C = m->D / 8; m->ch = hd / 8;
  return OCTIC_OK;
}

template <int HD>
static int launch_fwd(const void* qkv, void* o, float* lse, int B, int N, int H, const HeadMap& m, cudaStream_t s) {
  const int Npad = (N + 15) & ~15;
  const int smem = 2 * Npad * (HD + 8) * 2 + 3 * (HD / 2) * 4;
  static bool done = false;
  if (!done) {
    if (cudaFuncSetAttribute(attn_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return OCTIC_ERR_CUDA;
    done = true;
  }
  if (smem > 227 * 1024) return OCTIC_ERR_ARG;
  const float scale_log2 = kLog2e / sqrtf(static_cast<float>(HD));
  attn_fwd_kernel<HD><<<B * H, attn_warps(N) * 32, smem, s>>>(static_cast<const __nv_bfloat16*>(qkv),
                                                              static_cast<__nv_bfloat16*>(o), lse, N, H, m, scale_log2);
  return cudaGetLastError() == cudaSuccess ? OCTIC_OK : OCTIC_ERR_CUDA;
}

template <int HD>
static int launch_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, float* delta, void* dqkv, int B,
                      int N, int H, const HeadMap& m, cudaStream_t s) {
  const int Npad = (N + 15) & ~15;
  const int smem_kv = 2 * Npad * (HD + 8) * 2 + 2 * Npad * 4 + 3 * (HD / 2) * 4;
  const int smem_q = 2 * Npad * (HD + 8) * 2 + 3 * (HD / 2) * 4;
  static bool done = false;
  if (!done) {
    if (cudaFuncSetAttribute(attn_bwd_kv_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(attn_bwd_q_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return OCTIC_ERR_CUDA;
    done = true;
  }
  if (smem_kv > 227 * 1024) return OCTIC_ERR_ARG;
  const float scale = 1.0f / sqrtf(static_cast<float>(HD));
  const float scale_log2 = kLog2e * scale;
  launch_attn_delta(static_cast<const __nv_bfloat16*>(o), static_cast<const __nv_bfloat16*>(d_o), delta, B, N, H, m, 0, s);
  const int threads = attn_warps(N) * 32;
  attn_bwd_kv_kernel<HD><<<B * H, threads, smem_kv, s>>>(static_cast<const __nv_bfloat16*>(qkv),
                                                         static_cast<const __nv_bfloat16*>(d_o), lse, delta,
                                                         static_cast<__nv_bfloat16*>(dqkv), N, H, m, scale, scale_log2);
  attn_bwd_q_kernel<HD><<<B * H, threads, smem_q, s>>>(static_cast<const __nv_bfloat16*>(qkv),
                                                       static_cast<const __nv_bfloat16*>(d_o), lse, delta,
                                                       static_cast<__nv_bfloat16*>(dqkv), N, H, m, scale, scale_log2);
  return cudaGetLastError() == cudaSuccess ? OCTIC_OK : OCTIC_ERR_CUDA;
}

// The tcgen05 kernels (attention_tc.cu) read head-major q, k, v (layouts OCTIC_ATTN_DENSE and
// OCTIC_ATTN_OCTIC_HEADMAJOR) and serve every shape inside their shared-memory / TMEM envelope; the mma.sync kernels
// above serve the packed-input octic layout and longer dense sequences.  OCTIC_ATTN_LEGACY=1 forces the latter for
// the dense layout (A/B timing, tests).
static bool legacy_forced() {
  static int legacy = -1;
  if (legacy < 0) {
    const char* e = getenv("OCTIC_ATTN_LEGACY");
    legacy = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return legacy != 0;
}
static bool use_tc_path(int N, int hd, int layout, bool backward) {
  if (layout == OCTIC_ATTN_OCTIC_PACKED) return false;
  if (layout == OCTIC_ATTN_DENSE && legacy_forced()) return false;
  return attn_tc_supported(N, hd, backward);
}

}  // namespace octic

using namespace octic;

extern "C" {

int octic_attention_headmajor_supported(int N, int hd, int backward) {
  return attn_tc_supported(N, hd, backward != 0) ? 1 : 0;
}

int octic_attention_fwd(const void* qkv, void* o, float* lse, int B, int N, int H, int hd, int octic_layout,
                        void* stream) {
  if (!qkv || !o || B <= 0 || N <= 0) return OCTIC_ERR_ARG;
  HeadMap m;
  int rc = make_head_map(&m, H, hd, octic_layout);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (use_tc_path(N, hd, octic_layout, false)) return launch_attn_fwd_tc(qkv, o, lse, B, N, H, m, s);
  if (octic_layout == OCTIC_ATTN_OCTIC_HEADMAJOR) return OCTIC_ERR_ARG;   // only the tcgen05 path reads this layout
  switch (hd) {
    case 32: return launch_fwd<32>(qkv, o, lse, B, N, H, m, s);
    case 64: return launch_fwd<64>(qkv, o, lse, B, N, H, m, s);
    case 80: return launch_fwd<80>(qkv, o, lse, B, N, H, m, s);
    case 96: return launch_fwd<96>(qkv, o, lse, B, N, H, m, s);
    case 128: return launch_fwd<128>(qkv, o, lse, B, N, H, m, s);
    default: return OCTIC_ERR_ARG;
  }
}

size_t octic_attention_bwd_workspace_bytes(int N, int hd) { return attn_bwd_tc_workspace_bytes(N, hd); }

int octic_attention_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, float* delta_ws,
                        void* dqkv, int B, int N, int H, int hd, int octic_layout, void* stream) {
  return octic_attention_bwd_ws(qkv, o, d_o, lse, delta_ws, dqkv, B, N, H, hd, octic_layout, nullptr, 0, stream);
}

int octic_attention_bwd_ws(const void* qkv, const void* o, const void* d_o, const float* lse, float* delta_ws,
                           void* dqkv, int B, int N, int H, int hd, int octic_layout, void* ws, size_t ws_bytes,
                           void* stream) {
  if (!qkv || !o || !d_o || !lse || !delta_ws || !dqkv || B <= 0 || N <= 0) return OCTIC_ERR_ARG;
  HeadMap m;
  int rc = make_head_map(&m, H, hd, octic_layout);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (use_tc_path(N, hd, octic_layout, true)) {
    launch_attn_delta(static_cast<const __nv_bfloat16*>(o), static_cast<const __nv_bfloat16*>(d_o), delta_ws, B, N, H, m,
                      octic_layout == OCTIC_ATTN_OCTIC_HEADMAJOR, s);
    if (cudaGetLastError() != cudaSuccess) return OCTIC_ERR_CUDA;
    return launch_attn_bwd_tc(qkv, d_o, lse, delta_ws, dqkv, B, N, H, m, ws, ws_bytes, s);
  }
  if (octic_layout == OCTIC_ATTN_OCTIC_HEADMAJOR) return OCTIC_ERR_ARG;
  switch (hd) {
    case 32: return launch_bwd<32>(qkv, o, d_o, lse, delta_ws, dqkv, B, N, H, m, s);
    case 64: return launch_bwd<64>(qkv, o, d_o, lse, delta_ws, dqkv, B, N, H, m, s);
    case 80: return launch_bwd<80>(qkv, o, d_o, lse, delta_ws, dqkv, B, N, H, m, s);
    case 96: return launch_bwd<96>(qkv, o, d_o, lse, delta_ws, dqkv, B, N, H, m, s);
    case 128: return launch_bwd<128>(qkv, o, d_o, lse, delta_ws, dqkv, B, N, H, m, s);
    default: return OCTIC_ERR_ARG;
  }
}

}  // extern "C"
