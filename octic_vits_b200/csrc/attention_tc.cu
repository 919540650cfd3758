// Octic / dense multi-head attention on tcgen05 (sm_100a), forward and backward, for ViT-length sequences
// (N = 65..~300 tokens: the whole K, V, Q, dO of one (image, head) live in shared memory).
//
// Reference: AttentionD8.forward (octic_vits/d8_layers.py:623-656) = pack 5-tuple -> F.scaled_dot_product_attention ->
// unpack; dense Attention (deit/vit.py:36-50).  As in attention.cu the head vector is gathered from the packed qkv
// row while it is staged into shared memory (cp.async) and the result is scattered straight back into packed rows.
//
// Layout.  Every operand tile is a row-major [rows][hd] bf16 matrix in the SWIZZLE_32B "atom column" format of
// sm100_ptx.cuh: it can be fed to tcgen05.mma both K-major (contraction over hd: Q K^T, dO V^T) and MN-major
// (contraction over tokens: P V, P^T dO, dS^T Q, dS K) without a transposed copy.  hd = 80 (ViT-H/14) costs no padding.
//
// Forward (one CTA per (image, head), 128 threads = 128 TMEM lanes = 128 query rows per tile, two CTAs per SM):
//   pass 1  S chunk = Q_tile K_chunk^T -> TMEM (two ping-pong buffers) -> row max           (exact, no online rescale)
//   pass 2  S chunk again -> p = exp2(s*c - max*c) -> bf16 P written in place over S in TMEM -> O += P V_chunk with
//           the A operand read from TMEM (tcgen05.mma .ts form), accumulators in TMEM.
//   The tensor pipe executes MMAs in issue order, so "O += P_c V_c" followed by "S chunk c+2" into the same columns
//   needs no barrier; the only synchronisation per chunk is one mbarrier wait + one __syncthreads.
// Backward (one CTA per (image, head), 256 threads: two warps per TMEM lane quarter split the columns):
//   phase 1 (lanes = keys)     S^T = K Q^T, dP^T = V dO^T per query chunk -> P^T, dS^T (bf16, in place) ->
//                              dV += P^T dO, dK += dS^T Q
//   phase 2 (lanes = queries)  S = Q K^T, dP = dO V^T per key chunk -> dS in place -> dQ += dS K
//   delta = rowsum(dO * O) comes from attn_delta_kernel (attention.cu).
#include "attention_common.cuh"
#include "sm100_ptx.cuh"

namespace octic {

struct ChunkPlan {
  int n;
  int off[8];
  int w[8];
};

__host__ __device__ constexpr int tc_chunk_width(int hd) { return hd <= 96 ? 80 : 64; }

static bool make_plan(ChunkPlan* p, int N, int hd) {
  const int cw = tc_chunk_width(hd);
  const int Rk = (N + 15) / 16 * 16;
  const int units = Rk / 16;
  const int n = (Rk + cw - 1) / cw;
  if (n > 8) return false;
  int off = 0;
  for (int i = 0; i < n; ++i) {
    const int w = 16 * (units / n + (i < units % n ? 1 : 0));
    p->off[i] = off;
    p->w[i] = w;
    off += w;
  }
  for (int i = n; i < 8; ++i) { p->off[i] = off; p->w[i] = 0; }
  p->n = n;
  return true;
}

__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
  const __nv_bfloat162 b = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&b);
}

template <int GRAN>
__device__ __forceinline__ void cp_async_g(uint32_t dst, const void* src) {
  if (GRAN == 16) cp_async_16(dst, src); else cp_async_4(dst, src);
}
template <int GRAN>
__device__ __forceinline__ void st_zero_g(uint32_t dst) {
  if (GRAN == 16) asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
  else asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst), "r"(0u) : "memory");
}

// Stage rows [0, nrows) of a [R][HD] SWIZZLE_32B tile: rows < nvalid come from global (row r at src + r*ld, element
// unit u at column `col` chosen per thread by the caller's map), the rest are zero-filled.  A "unit" is GRAN bytes.
// Caller: cp_async_commit_wait_all(); fence_proxy_async_smem(); __syncthreads().
template <int HD, int GRAN>
__device__ __forceinline__ void stage_tile(uint32_t dst, int R, const __nv_bfloat16* src, long ld, int nvalid, int nrows,
                                           const int* cb, const int* sm, int s) {
  constexpr int EU = GRAN / 2, NU = HD / EU;
  const int rpi = blockDim.x / NU;
  const int u = threadIdx.x % NU, r0 = threadIdx.x / NU;
  if (r0 >= rpi) return;
  const int c = u * EU;
  const int col = cb[u] + (sm != nullptr ? s * sm[u] : 0);
  const uint32_t cbase = dst + (c >> 4) * (R * 32) + (c & 7) * 2;
  const int half = (c >> 3) & 1;
  const __nv_bfloat16* g = src + static_cast<long>(r0) * ld + col;
  const long gstep = static_cast<long>(rpi) * ld;
  for (int r = r0; r < nrows; r += rpi, g += gstep) {
    const uint32_t a = cbase + r * 32 + ((half ^ ((r >> 2) & 1)) << 4);
    if (r < nvalid) cp_async_g<GRAN>(a, g);
    else st_zero_g<GRAN>(a);
  }
}

// Rows [0, nvalid) of a plain row-major [.][HD] bf16 staging tile -> global rows (scatter through the column map).
template <int HD, int GRAN>
__device__ __forceinline__ void store_tile(const uint8_t* stg, __nv_bfloat16* dst, long ld, int nvalid, const int* cb,
                                           const int* sm, int s) {
  constexpr int EU = GRAN / 2, NU = HD / EU;
  const int rpi = blockDim.x / NU;
  const int u = threadIdx.x % NU, r0 = threadIdx.x / NU;
  if (r0 >= rpi) return;
  const int col = cb[u] + (sm != nullptr ? s * sm[u] : 0);
  for (int r = r0; r < nvalid; r += rpi) {
    const uint8_t* sp = stg + r * (HD * 2) + u * GRAN;
    __nv_bfloat16* gp = dst + static_cast<long>(r) * ld + col;
    if (GRAN == 16) *reinterpret_cast<uint4*>(gp) = *reinterpret_cast<const uint4*>(sp);
    else *reinterpret_cast<uint32_t*>(gp) = *reinterpret_cast<const uint32_t*>(sp);
  }
}

// 16 fp32 accumulator columns of this thread's TMEM lane -> * scale -> bf16 -> 32 bytes of its staging row
__device__ __forceinline__ void acc_piece_to_stg(uint32_t taddr, float scale, uint8_t* dst) {
  uint32_t r[16];
  tmem_ld_32x16(taddr, r);
  tmem_ld_wait();
  uint4 v0, v1;
  v0.x = pack2_bf16(__uint_as_float(r[0]) * scale, __uint_as_float(r[1]) * scale);
  v0.y = pack2_bf16(__uint_as_float(r[2]) * scale, __uint_as_float(r[3]) * scale);
  v0.z = pack2_bf16(__uint_as_float(r[4]) * scale, __uint_as_float(r[5]) * scale);
  v0.w = pack2_bf16(__uint_as_float(r[6]) * scale, __uint_as_float(r[7]) * scale);
  v1.x = pack2_bf16(__uint_as_float(r[8]) * scale, __uint_as_float(r[9]) * scale);
  v1.y = pack2_bf16(__uint_as_float(r[10]) * scale, __uint_as_float(r[11]) * scale);
  v1.z = pack2_bf16(__uint_as_float(r[12]) * scale, __uint_as_float(r[13]) * scale);
  v1.w = pack2_bf16(__uint_as_float(r[14]) * scale, __uint_as_float(r[15]) * scale);
  *reinterpret_cast<uint4*>(dst) = v0;
  *reinterpret_cast<uint4*>(dst + 16) = v1;
}

// =====================================================================================================================
//  forward
// =====================================================================================================================
template <int HD, int GRAN>
__global__ void __launch_bounds__(128) attn_fwd_tc_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ o,
                                                          float* __restrict__ lse, int N, int H, HeadMap m,
                                                          float scale_log2, ChunkPlan cp) {
  constexpr int KS = HD / 16, CW = tc_chunk_width(HD), OCOL = 2 * CW, NU = HD / (GRAN / 2), NPMAX = CW / 16;
  constexpr uint32_t kTmemCols = 256;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int Rk = cp.off[cp.n - 1] + cp.w[cp.n - 1];
  const uint32_t kv_bytes = static_cast<uint32_t>(Rk) * HD * 2;
  uint8_t* Ks = smem;
  uint8_t* Vs = Ks + kv_bytes;
  uint8_t* Qs = Vs + kv_bytes;
  int* cb = reinterpret_cast<int*>(Qs + 128 * HD * 2);
  int* sm = cb + NU;
  int* ocb = sm + NU;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ocb + NU);     // [0], [1]: S buffers; [2]: O complete
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const long ld3 = 3L * m.D;
  const __nv_bfloat16* rows = qkv + static_cast<long>(b) * N * ld3;
  __nv_bfloat16* orows = o + static_cast<long>(b) * N * m.D;

  if (tid < NU) {
    int base, smul;
    qkv_col(m, h, tid * (GRAN / 2), base, smul);
    cb[tid] = base; sm[tid] = smul;
    ocb[tid] = o_col(m, h, tid * (GRAN / 2));
  }
  if (tid == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  __syncthreads();
  const uint32_t q_addr = smem_u32(Qs), k_addr = smem_u32(Ks), v_addr = smem_u32(Vs);
  stage_tile<HD, GRAN>(k_addr, Rk, rows, ld3, N, Rk, cb, sm, 1);
  stage_tile<HD, GRAN>(v_addr, Rk, rows, ld3, N, Rk, cb, sm, 2);
  stage_tile<HD, GRAN>(q_addr, 128, rows, ld3, min(128, N), 128, cb, sm, 0);
  cp_async_commit_wait_all();
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);

  auto issue_l1 = [&](int c, int buf) {      // S chunk c = Q_tile K_c^T  -> buffer buf
    const uint32_t idesc = make_idesc_bf16(128, cp.w[c], 0, 0);
    const uint32_t d = tmem_base + buf * CW;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
      umma_bf16(d, desc_sw32_k(q_addr, 128, 0, ks), desc_sw32_k(k_addr, Rk, cp.off[c], ks), idesc, ks != 0);
    umma_commit(&bars[buf]);
  };
  auto issue_l2 = [&](int c, int buf, bool first) {   // O (+)= P_c V_c, P read from TMEM
    const uint32_t idesc = make_idesc_bf16(128, HD, 0, 1);
    const int nk = cp.w[c] >> 4;
    for (int kk = 0; kk < nk; ++kk)
      umma_bf16_ts(tmem_base + OCOL, tmem_base + buf * CW + kk * 8, desc_sw32_mn(v_addr, Rk, cp.off[c] + 16 * kk), idesc,
                   !(first && kk == 0));
  };

  uint32_t ph[2] = {0u, 0u}, pho = 0u;
  const int nt = (N + 127) >> 7, nc = cp.n, njobs = 2 * nc;
  for (int t = 0; t < nt; ++t) {
    const bool warp_valid = t * 128 + warp * 32 < N;
    if (tid == 0) issue_l1(0, 0);
    float mx = -INFINITY, l = 0.f, moff = 0.f;
    for (int j = 0; j < njobs; ++j) {
      const int c = j < nc ? j : j - nc, buf = j & 1;
      if (tid == 0 && j + 1 < njobs) issue_l1(j + 1 < nc ? j + 1 : j + 1 - nc, (j + 1) & 1);
      mbar_wait(&bars[buf], ph[buf]);
      ph[buf] ^= 1u;
      tc_fence_after();
      if (warp_valid) {
        const int w = cp.w[c], k0 = cp.off[c];
        const uint32_t ta = t_lane + buf * CW;
        const bool need_mask = k0 + w > N;
        uint32_t r[NPMAX][16];
#pragma unroll
        for (int p = 0; p < NPMAX; ++p)
          if (p * 16 < w) tmem_ld_32x16(ta + p * 16, r[p]);
        tmem_ld_wait();
        if (j < nc) {
#pragma unroll
          for (int p = 0; p < NPMAX; ++p)
            if (p * 16 < w) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                float v = __uint_as_float(r[p][i]);
                if (need_mask && k0 + p * 16 + i >= N) v = -INFINITY;
                mx = fmaxf(mx, v);
              }
            }
          if (j == nc - 1) moff = mx * scale_log2;
        } else {
#pragma unroll
          for (int p = 0; p < NPMAX; ++p)
            if (p * 16 < w) {
              uint32_t pk[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float e0 = exp2f(fmaf(__uint_as_float(r[p][2 * i]), scale_log2, -moff));
                float e1 = exp2f(fmaf(__uint_as_float(r[p][2 * i + 1]), scale_log2, -moff));
                if (need_mask) {
                  if (k0 + p * 16 + 2 * i >= N) e0 = 0.f;
                  if (k0 + p * 16 + 2 * i + 1 >= N) e1 = 0.f;
                }
                l += e0 + e1;
                pk[i] = pack2_bf16(e0, e1);
              }
              tmem_st_32x8(ta + p * 8, pk);
            }
          tmem_st_wait();
        }
      }
      tc_fence_before();
      __syncthreads();
      if (j >= nc && tid == 0) {
        tc_fence_after();
        issue_l2(c, buf, j == nc);
        if (j == njobs - 1) umma_commit(&bars[2]);
      }
    }
    mbar_wait(&bars[2], pho);
    pho ^= 1u;
    tc_fence_after();
    const int row = t * 128 + tid;
    if (warp_valid) {
      const float inv = 1.0f / l;
#pragma unroll
      for (int p = 0; p < KS; ++p) acc_piece_to_stg(t_lane + OCOL + p * 16, inv, Qs + tid * (HD * 2) + p * 32);
      if (lse != nullptr && row < N) lse[(static_cast<long>(b) * H + h) * N + row] = (mx * scale_log2 + log2f(l)) * kLn2;
    }
    tc_fence_before();
    __syncthreads();
    store_tile<HD, GRAN>(Qs, orows + static_cast<long>(t) * 128 * m.D, m.D, min(128, N - t * 128), ocb, nullptr, 0);
    __syncthreads();
    if (t + 1 < nt) {
      stage_tile<HD, GRAN>(q_addr, 128, rows + static_cast<long>(t + 1) * 128 * ld3, ld3, min(128, N - (t + 1) * 128), 128,
                           cb, sm, 0);
      cp_async_commit_wait_all();
      fence_proxy_async_smem();
      __syncthreads();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// =====================================================================================================================
//  backward
// =====================================================================================================================
template <int HD, int GRAN>
__global__ void __launch_bounds__(256, 1) attn_bwd_tc_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                             const __nv_bfloat16* __restrict__ d_o,
                                                             const float* __restrict__ lse, const float* __restrict__ delta,
                                                             __nv_bfloat16* __restrict__ dqkv, int N, int H, HeadMap m,
                                                             float scale, float scale_log2, ChunkPlan cp) {
  constexpr int KS = HD / 16, CW = tc_chunk_width(HD), ACC = 4 * CW, NU = HD / (GRAN / 2);
  constexpr int NPW = (CW / 16 + 1) / 2;        // max 16-column pieces per warp in the math step
  constexpr int KS0 = (KS + 1) / 2;             // epilogue: pieces of the accumulator handled by column-half 0
  constexpr uint32_t kTmemCols = 512;
  static_assert(ACC + 2 * HD <= 512, "TMEM budget");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int Rk = cp.off[cp.n - 1] + cp.w[cp.n - 1];
  const uint32_t mat_bytes = static_cast<uint32_t>(Rk) * HD * 2;
  uint8_t* Qs = smem;
  uint8_t* Ks = Qs + mat_bytes;
  uint8_t* Vs = Ks + mat_bytes;
  uint8_t* dOs = Vs + mat_bytes;
  uint8_t* stg = dOs + mat_bytes;                 // [128][HD] bf16 staging; also absorbs the A-tile over-read of dOs
  float* lse_s = reinterpret_cast<float*>(stg + 128 * HD * 2);
  float* del_s = lse_s + Rk;
  int* cb = reinterpret_cast<int*>(del_s + Rk);
  int* sm = cb + NU;
  int* ocb = sm + NU;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ocb + NU);     // [0], [1]: first-level buffers; [2]: accumulators complete
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q4 = warp & 3, hh = warp >> 2;
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const long ld3 = 3L * m.D;
  const __nv_bfloat16* rows = qkv + static_cast<long>(b) * N * ld3;
  const __nv_bfloat16* dorows = d_o + static_cast<long>(b) * N * m.D;
  __nv_bfloat16* drows = dqkv + static_cast<long>(b) * N * ld3;

  if (tid < NU) {
    int base, smul;
    qkv_col(m, h, tid * (GRAN / 2), base, smul);
    cb[tid] = base; sm[tid] = smul;
    ocb[tid] = o_col(m, h, tid * (GRAN / 2));
  }
  if (tid == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  for (int i = tid; i < Rk; i += blockDim.x) {
    const long off = (static_cast<long>(b) * H + h) * N + i;
    lse_s[i] = i < N ? lse[off] * kLog2e : INFINITY;     // +inf -> P = 0 for padded queries
    del_s[i] = i < N ? delta[off] : 0.f;
  }
  for (int i = tid; i < 128 * HD * 2 / 16; i += blockDim.x) reinterpret_cast<uint4*>(stg)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  const uint32_t q_addr = smem_u32(Qs), k_addr = smem_u32(Ks), v_addr = smem_u32(Vs), do_addr = smem_u32(dOs);
  stage_tile<HD, GRAN>(q_addr, Rk, rows, ld3, N, Rk, cb, sm, 0);
  stage_tile<HD, GRAN>(k_addr, Rk, rows, ld3, N, Rk, cb, sm, 1);
  stage_tile<HD, GRAN>(v_addr, Rk, rows, ld3, N, Rk, cb, sm, 2);
  stage_tile<HD, GRAN>(do_addr, Rk, dorows, m.D, N, Rk, ocb, nullptr, 0);
  cp_async_commit_wait_all();
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16);

  // first level: X = A1[tile rows] B1[chunk rows]^T -> buffer, Y = A2[tile rows] B2[chunk rows]^T -> buffer + CW
  auto issue_l1 = [&](uint32_t a1, uint32_t b1, uint32_t a2, uint32_t b2, int row0, int c, int buf) {
    const uint32_t idesc = make_idesc_bf16(128, cp.w[c], 0, 0);
    const uint32_t d = tmem_base + buf * 2 * CW;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
      umma_bf16(d, desc_sw32_k(a1, Rk, row0, ks), desc_sw32_k(b1, Rk, cp.off[c], ks), idesc, ks != 0);
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
      umma_bf16(d + CW, desc_sw32_k(a2, Rk, row0, ks), desc_sw32_k(b2, Rk, cp.off[c], ks), idesc, ks != 0);
    umma_commit(&bars[buf]);
  };
  // second level: acc (+)= (bf16 in TMEM at a_col) * B[chunk rows] (MN-major)
  auto issue_l2 = [&](uint32_t acc_col, uint32_t a_col, uint32_t bmat, int c, bool first) {
    const uint32_t idesc = make_idesc_bf16(128, HD, 0, 1);
    const int nk = cp.w[c] >> 4;
    for (int kk = 0; kk < nk; ++kk)
      umma_bf16_ts(tmem_base + acc_col, tmem_base + a_col + kk * 8, desc_sw32_mn(bmat, Rk, cp.off[c] + 16 * kk), idesc,
                   !(first && kk == 0));
  };
  // accumulator -> * sc -> bf16 staging -> packed global rows (tensor s of the qkv row)
  auto flush_acc = [&](uint32_t acc_col, float sc, int row0, int s, bool warp_valid) {
    if (warp_valid) {
      const int p0 = hh == 0 ? 0 : KS0, p1 = hh == 0 ? KS0 : KS;
      for (int p = p0; p < p1; ++p)
        acc_piece_to_stg(t_lane + acc_col + p * 16, sc, stg + (q4 * 32 + lane) * (HD * 2) + p * 32);
    }
    tc_fence_before();
    __syncthreads();
    store_tile<HD, GRAN>(stg, drows + static_cast<long>(row0) * ld3, ld3, min(128, N - row0), cb, sm, s);
    __syncthreads();
  };

  uint32_t ph[2] = {0u, 0u}, pha = 0u;
  const int nt = (N + 127) >> 7, nc = cp.n;

  for (int phase = 0; phase < 2; ++phase) {
    for (int t = 0; t < nt; ++t) {
      const int row0 = t * 128;
      const int my_row = row0 + q4 * 32 + lane;
      const bool warp_valid = row0 + q4 * 32 < N;
      // phase 0: lanes = keys,    X = S^T = K Q^T,  Y = dP^T = V dO^T
      // phase 1: lanes = queries, X = S   = Q K^T,  Y = dP   = dO V^T
      const uint32_t a1 = phase == 0 ? k_addr : q_addr, b1 = phase == 0 ? q_addr : k_addr;
      const uint32_t a2 = phase == 0 ? v_addr : do_addr, b2 = phase == 0 ? do_addr : v_addr;
      float lse_r = 0.f, del_r = 0.f;
      if (phase == 1) {
        lse_r = my_row < Rk ? lse_s[my_row] : INFINITY;
        del_r = my_row < Rk ? del_s[my_row] : 0.f;
      }
      if (tid == 0) issue_l1(a1, b1, a2, b2, row0, 0, 0);
      for (int c = 0; c < nc; ++c) {
        const int buf = c & 1;
        if (tid == 0 && c + 1 < nc) issue_l1(a1, b1, a2, b2, row0, c + 1, (c + 1) & 1);
        mbar_wait(&bars[buf], ph[buf]);
        ph[buf] ^= 1u;
        tc_fence_after();
        if (warp_valid) {
          const int w = cp.w[c], k0 = cp.off[c];
          const int np = w >> 4, np0 = (np + 1) >> 1;
          const int pb = hh == 0 ? 0 : np0, pe = hh == 0 ? np0 : np;
          const uint32_t tx = t_lane + buf * 2 * CW, ty = tx + CW;
          uint32_t rx[NPW][16], ry[NPW][16];
#pragma unroll
          for (int pp = 0; pp < NPW; ++pp)
            if (pb + pp < pe) {
              tmem_ld_32x16(tx + (pb + pp) * 16, rx[pp]);
              tmem_ld_32x16(ty + (pb + pp) * 16, ry[pp]);
            }
          tmem_ld_wait();
          // both warps of this lane quarter hold their inputs in registers before either overwrites the columns
          tc_fence_before();
          named_bar_sync(1 + q4, 64);
          tc_fence_after();
#pragma unroll
          for (int pp = 0; pp < NPW; ++pp)
            if (pb + pp < pe) {
              const int col0 = k0 + (pb + pp) * 16;
              uint32_t pkp[8], pkd[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float p0, p1, d0, d1;
                if (phase == 0) {
                  const float2 l2 = *reinterpret_cast<const float2*>(lse_s + col0 + 2 * i);
                  const float2 dl = *reinterpret_cast<const float2*>(del_s + col0 + 2 * i);
                  p0 = exp2f(fmaf(__uint_as_float(rx[pp][2 * i]), scale_log2, -l2.x));
                  p1 = exp2f(fmaf(__uint_as_float(rx[pp][2 * i + 1]), scale_log2, -l2.y));
                  d0 = p0 * (__uint_as_float(ry[pp][2 * i]) - dl.x);
                  d1 = p1 * (__uint_as_float(ry[pp][2 * i + 1]) - dl.y);
                } else {
                  p0 = exp2f(fmaf(__uint_as_float(rx[pp][2 * i]), scale_log2, -lse_r));
                  p1 = exp2f(fmaf(__uint_as_float(rx[pp][2 * i + 1]), scale_log2, -lse_r));
                  if (col0 + 2 * i >= N) p0 = 0.f;
                  if (col0 + 2 * i + 1 >= N) p1 = 0.f;
                  d0 = p0 * (__uint_as_float(ry[pp][2 * i]) - del_r);
                  d1 = p1 * (__uint_as_float(ry[pp][2 * i + 1]) - del_r);
                }
                pkp[i] = pack2_bf16(p0, p1);
                pkd[i] = pack2_bf16(d0, d1);
              }
              if (phase == 0) tmem_st_32x8(tx + (pb + pp) * 8, pkp);
              tmem_st_32x8(ty + (pb + pp) * 8, pkd);
            }
          tmem_st_wait();
        }
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
          tc_fence_after();
          const uint32_t xcol = buf * 2 * CW, ycol = xcol + CW;
          if (phase == 0) {
            issue_l2(ACC + HD, xcol, do_addr, c, c == 0);    // dV += P^T dO
            issue_l2(ACC, ycol, q_addr, c, c == 0);          // dK += dS^T Q
          } else {
            issue_l2(ACC, ycol, k_addr, c, c == 0);          // dQ += dS K
          }
          if (c == nc - 1) umma_commit(&bars[2]);
        }
      }
      mbar_wait(&bars[2], pha);
      pha ^= 1u;
      tc_fence_after();
      if (phase == 0) {
        flush_acc(ACC, scale, row0, 1, warp_valid);          // dK
        flush_acc(ACC + HD, 1.0f, row0, 2, warp_valid);      // dV
      } else {
        flush_acc(ACC, scale, row0, 0, warp_valid);          // dQ
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// =====================================================================================================================
//  host
// =====================================================================================================================
constexpr int kMaxSmem = 232448;   // 227 KiB

static size_t fwd_smem(int Rk, int hd, int gran) {
  const int nu = hd / (gran / 2);
  return static_cast<size_t>(2 * Rk + 128) * hd * 2 + 3 * nu * 4 + 3 * 8 + 16 + 1024;
}
static size_t bwd_smem(int Rk, int hd, int gran) {
  const int nu = hd / (gran / 2);
  return static_cast<size_t>(4 * Rk + 128) * hd * 2 + 2 * Rk * 4 + 3 * nu * 4 + 3 * 8 + 16 + 1024;
}

bool attn_tc_supported(int N, int hd, bool backward) {
  if (hd != 64 && hd != 80 && hd != 32 && hd != 96 && hd != 128) return false;
  ChunkPlan p;
  if (N < 1 || !make_plan(&p, N, hd)) return false;
  const int Rk = (N + 15) / 16 * 16;
  return (backward ? bwd_smem(Rk, hd, 4) : fwd_smem(Rk, hd, 4)) <= static_cast<size_t>(kMaxSmem);
}

static int pick_gran(const HeadMap& m) { return (!m.octic || (m.ch % 8) == 0) ? 16 : 4; }

template <int HD, int GRAN>
static int launch_fwd_t(const void* qkv, void* o, float* lse, int B, int N, int H, const HeadMap& m, const ChunkPlan& cp,
                        cudaStream_t s) {
  const int Rk = (N + 15) / 16 * 16;
  const size_t smem = fwd_smem(Rk, HD, GRAN);
  static bool done = false;
  if (!done) {
    if (cudaFuncSetAttribute(attn_fwd_tc_kernel<HD, GRAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess)
      return OCTIC_ERR_CUDA;
    done = true;
  }
  const float scale_log2 = kLog2e / sqrtf(static_cast<float>(HD));
  attn_fwd_tc_kernel<HD, GRAN><<<B * H, 128, smem, s>>>(static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(o),
                                                        lse, N, H, m, scale_log2, cp);
  return cudaGetLastError() == cudaSuccess ? OCTIC_OK : OCTIC_ERR_CUDA;
}
template <int HD, int GRAN>
static int launch_bwd_t(const void* qkv, const void* d_o, const float* lse, const float* delta, void* dqkv, int B, int N,
                        int H, const HeadMap& m, const ChunkPlan& cp, cudaStream_t s) {
  const int Rk = (N + 15) / 16 * 16;
  const size_t smem = bwd_smem(Rk, HD, GRAN);
  static bool done = false;
  if (!done) {
    if (cudaFuncSetAttribute(attn_bwd_tc_kernel<HD, GRAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess)
      return OCTIC_ERR_CUDA;
    done = true;
  }
  const float scale = 1.0f / sqrtf(static_cast<float>(HD));
  attn_bwd_tc_kernel<HD, GRAN><<<B * H, 256, smem, s>>>(static_cast<const __nv_bfloat16*>(qkv),
                                                        static_cast<const __nv_bfloat16*>(d_o), lse, delta,
                                                        static_cast<__nv_bfloat16*>(dqkv), N, H, m, scale, kLog2e * scale, cp);
  return cudaGetLastError() == cudaSuccess ? OCTIC_OK : OCTIC_ERR_CUDA;
}

#define OCTIC_TC_DISPATCH(FN, ...)                                                     \
  switch (m.hd) {                                                                      \
    case 32: return gran == 16 ? FN<32, 16>(__VA_ARGS__) : FN<32, 4>(__VA_ARGS__);    \
    case 64: return gran == 16 ? FN<64, 16>(__VA_ARGS__) : FN<64, 4>(__VA_ARGS__);    \
    case 80: return gran == 16 ? FN<80, 16>(__VA_ARGS__) : FN<80, 4>(__VA_ARGS__);    \
    case 96: return gran == 16 ? FN<96, 16>(__VA_ARGS__) : FN<96, 4>(__VA_ARGS__);    \
    case 128: return gran == 16 ? FN<128, 16>(__VA_ARGS__) : FN<128, 4>(__VA_ARGS__); \
    default: return OCTIC_ERR_ARG;                                                     \
  }

int launch_attn_fwd_tc(const void* qkv, void* o, float* lse, int B, int N, int H, const HeadMap& m, cudaStream_t s) {
  ChunkPlan cp;
  if (!attn_tc_supported(N, m.hd, false) || !make_plan(&cp, N, m.hd)) return OCTIC_ERR_ARG;
  const int gran = pick_gran(m);
  OCTIC_TC_DISPATCH(launch_fwd_t, qkv, o, lse, B, N, H, m, cp, s)
}

int launch_attn_bwd_tc(const void* qkv, const void* d_o, const float* lse, const float* delta, void* dqkv, int B, int N,
                       int H, const HeadMap& m, cudaStream_t s) {
  ChunkPlan cp;
  if (!attn_tc_supported(N, m.hd, true) || !make_plan(&cp, N, m.hd)) return OCTIC_ERR_ARG;
  const int gran = pick_gran(m);
  OCTIC_TC_DISPATCH(launch_bwd_t, qkv, d_o, lse, delta, dqkv, B, N, H, m, cp, s)
}

}  // namespace octic
