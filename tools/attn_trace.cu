// In-kernel timeline of the tcgen05 attention forward kernel (CTA 0): build with
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --use_fast_math -DOCTIC_ATTN_TRACE \
//        -o build/attn_trace tools/attn_trace.cu -lcuda
// and run on a B200:  build/attn_trace [B]   (prints cycle deltas between events of the control and math threads)
#include "../octic_vits_b200/csrc/attention_tc.cu"
#include <cstdio>
#include <vector>

using namespace octic;

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 20, N = 257, H = 16, hd = 80, D = H * hd;
  const bool bwd = argc > 2 && argv[2][0] == 'b';
  const int octic = argc > 3 ? atoi(argv[3]) : 0;
  const int staged = argc > 4 ? atoi(argv[4]) : 1;      // backward: staged dQ (L2 scratch) or the two-pass kernel
  const size_t nq = static_cast<size_t>(B) * N * 3 * D;
  std::vector<__nv_bfloat16> h(nq);
  unsigned s = 12345u;
  for (size_t i = 0; i < nq; ++i) {
    s = s * 1664525u + 1013904223u;
    h[i] = __float2bfloat16((static_cast<float>(s >> 8) / 16777216.0f - 0.5f) * 3.0f);
  }
  __nv_bfloat16 *qkv, *o, *dqkv;
  float *lse, *delta;
  cudaMalloc(&dqkv, nq * 2);
  cudaMalloc(&delta, static_cast<size_t>(B) * H * N * 4);
  cudaMemset(delta, 0, static_cast<size_t>(B) * H * N * 4);
  cudaMalloc(&qkv, nq * 2);
  cudaMalloc(&o, static_cast<size_t>(B) * N * D * 2);
  cudaMalloc(&lse, static_cast<size_t>(B) * H * N * 4);
  cudaMemcpy(qkv, h.data(), nq * 2, cudaMemcpyHostToDevice);
  void* ws = nullptr;
  // argv[5]: number of scratch slots (default: what the library asks for, 2 x SM count); must be >= the SM count
  const int slots = argc > 5 ? atoi(argv[5]) : 0;
  const size_t rk = static_cast<size_t>((N + 15) / 16 * 16);
  const size_t ws_bytes = !staged ? 0 : slots > 0 ? 4096 + slots * rk * rk * 2 : attn_bwd_tc_workspace_bytes(N, hd);
  if (ws_bytes) { cudaMalloc(&ws, ws_bytes); cudaMemset(ws, 0, ws_bytes); }
  HeadMap m;
  m.octic = octic; m.hd = hd; m.D = D; m.C = D / 8; m.ch = hd / 8;
  for (int it = 0; it < 3; ++it) {
#ifdef OCTIC_ATTN_TRACE
    int zero[2] = {0, 0};
    cudaMemcpyToSymbol(g_trace_n, zero, sizeof(zero));
#endif
    if (bwd && it == 0) {
      launch_attn_fwd_tc(qkv, o, lse, B, N, H, m, 0);
      cudaDeviceSynchronize();
#ifdef OCTIC_ATTN_TRACE
      cudaMemcpyToSymbol(g_trace_n, zero, sizeof(zero));
#endif
    }
    int rc = bwd ? launch_attn_bwd_tc(qkv, o /* stands in for dO */, lse, delta, dqkv, B, N, H, m, ws, ws_bytes, 0)
                 : launch_attn_fwd_tc(qkv, o, lse, B, N, H, m, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (rc || e != cudaSuccess) { printf("launch rc=%d cuda=%s\n", rc, cudaGetErrorString(e)); return 1; }
  }
  {
    // wall time of the launch at this B (CUDA events, 5 launches after the warm-up above)
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int it = 0; it < 5; ++it) {
      if (bwd) launch_attn_bwd_tc(qkv, o, lse, delta, dqkv, B, N, H, m, ws, ws_bytes, 0);
      else launch_attn_fwd_tc(qkv, o, lse, B, N, H, m, 0);
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("B=%d %s octic=%d staged=%d: %.1f us per launch\n", B, bwd ? "bwd" : "fwd", octic, staged && ws_bytes, ms * 200.f);
  }
#ifdef OCTIC_ATTN_TRACE
  std::vector<long long> tr(2 * 2048);
  int n[2];
  cudaMemcpyFromSymbol(tr.data(), g_trace, sizeof(long long) * 2 * 2048);
  cudaMemcpyFromSymbol(n, g_trace_n, sizeof(n));
  long long t0 = tr[2048] >> 8;
  for (int slot = 0; slot < 2; ++slot) {
    printf("slot %d: %d events (id:cycles since first math event, delta)\n", slot, n[slot]);
    long long prev = t0;
    for (int i = 0; i < n[slot] && i < 2048; ++i) {
      const long long v = tr[slot * 2048 + i], t = v >> 8;
      printf("  %lld:%lld(+%lld)", v & 255, t - t0, t - prev);
      prev = t;
      if (i % 6 == 5) printf("\n");
    }
    printf("\n");
  }
#endif
  return 0;
}
