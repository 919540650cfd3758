"""Per-kernel microbenchmark at the headline shapes (hybrid ViT-H/14: D=1280, N=257, H=16, hd=80), CUDA-event timed,
with the algorithmic bytes / flops and the roofline fraction of each.  `--profile` wraps one call of every op in a
cudaProfilerStart/Stop range for `ncu --profile-from-start off ...`.

    python tools/microbench_ops.py --batch 64 [--only ln,gelu] [--profile]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from octic_vits_b200 import ops  # noqa: E402
from octic_vits_b200._lib import EPI_BF16, EPI_GELU_BF16, EPI_GELU_BWD, EPI_RESID  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--only", default="")
ap.add_argument("--profile", action="store_true")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--attn-layout", type=int, default=2, help="0 dense, 1 packed octic (mma.sync), 2 head-major octic (tcgen05)")
args = ap.parse_args()
ATTN_LAYOUT = args.attn_layout

peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {
    "hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
dev = "cuda"
B, N, D, H, hd = args.batch, 257, 1280, 16, 80
T, C = B * N, D // 8
g = torch.Generator(device=dev).manual_seed(0)
bf = torch.bfloat16


def rn(*shape, dtype=torch.float32):
    return torch.randn(*shape, device=dev, generator=g).to(dtype)


x32, dres = rn(T, D), rn(T, D)
xb, dyb = rn(T, D, dtype=bf), rn(T, D, dtype=bf)
h4, g4 = rn(T, 4 * D, dtype=bf), rn(T, 4 * D, dtype=bf)
qkv = rn(T, 3 * D, dtype=bf)
alpha, beta, gamma = rn(D) + 1, rn(C), rn(D)
w = {k: rn(*s) * 0.03 for k, s in dict(a=(3 * C, C), e=(6 * C, 2 * C)).items()}
pk_qkv = ops.pack_linear_d8(w["a"], w["a"].clone(), w["a"].clone(), w["a"].clone(), w["e"])
wp = {k: rn(*s) * 0.03 for k, s in dict(a=(C, C), e=(2 * C, 2 * C)).items()}
pk_proj = ops.pack_linear_d8(wp["a"], wp["a"].clone(), wp["a"].clone(), wp["a"].clone(), wp["e"])
w1 = {k: rn(*s) * 0.03 for k, s in dict(a=(4 * C, C), e=(8 * C, 2 * C)).items()}
pk_fc1 = ops.pack_linear_d8(w1["a"], w1["a"].clone(), w1["a"].clone(), w1["a"].clone(), w1["e"])
w2 = {k: rn(*s) * 0.03 for k, s in dict(a=(C, 4 * C), e=(2 * C, 8 * C)).items()}
pk_fc2 = ops.pack_linear_d8(w2["a"], w2["a"].clone(), w2["a"].clone(), w2["a"].clone(), w2["e"])
dq = ops.pack_linear(rn(3 * D, D) * 0.03)
dp = ops.pack_linear(rn(D, D) * 0.03)
d1 = ops.pack_linear(rn(4 * D, D) * 0.03)
d2 = ops.pack_linear(rn(D, 4 * D) * 0.03)
out3 = torch.empty(T, 3 * D, dtype=bf, device=dev)
out4 = torch.empty(T, 4 * D, dtype=bf, device=dev)
out1 = torch.empty(T, D, dtype=bf, device=dev)
res_out = torch.empty(T, D, device=dev)
colsum4 = torch.zeros(4 * D, device=dev)
branch = torch.empty(T, D, dtype=bf, device=dev)
_, stats8 = ops.layernorm_fwd(x32, alpha, beta, 1e-5, True)
_, stats2 = ops.layernorm_fwd(x32, alpha, alpha, 1e-6, False)
o_attn, lse = ops.attention_fwd(qkv, B, N, H, hd, ATTN_LAYOUT)
oct_f = 2.0 * T * D * 3 / 16   # x Dout

CASES = {
    # name: (callable, bytes, flops)
    "ln_d8_fwd": (lambda: ops.layernorm_fwd(x32, alpha, beta, 1e-5, True), T * D * 6, 0),
    "ln_d8_bwd": (lambda: ops.layernorm_bwd(dyb, x32, stats8, alpha, True, dx_in=dres), T * D * (2 + 4 + 4 + 4), 0),
    "ln_fwd": (lambda: ops.layernorm_fwd(x32, alpha, alpha, 1e-6, False), T * D * 6, 0),
    "ln_bwd": (lambda: ops.layernorm_bwd(dyb, x32, stats2, alpha, False, dx_in=dres), T * D * 14, 0),
    "gelu_d8_fwd": (lambda: ops.gelu_d8_fwd(h4), T * 4 * D * 4, 0),
    "gelu_d8_bwd": (lambda: ops.gelu_d8_bwd(g4, h4), T * 4 * D * 6, 0),
    "gelu_bwd": (lambda: ops.gelu_bwd(g4, h4), T * 4 * D * 6, 0),
    "layerscale_bwd": (lambda: ops.layerscale_bwd(dres, dyb, gamma, None, N), T * D * 8, 0),
    "colsum": (lambda: ops.colsum_bf16(h4, 4 * C), T * 4 * C * 2, 0),
    "attn_fwd": (lambda: ops.attention_fwd(qkv, B, N, H, hd, ATTN_LAYOUT), T * D * 8, 4.0 * B * H * N * N * hd),
    "attn_bwd": (lambda: ops.attention_bwd(qkv, o_attn, dyb, lse, B, N, H, hd, ATTN_LAYOUT), T * D * 18, 10.0 * B * H * N * N * hd),
    "d8_qkv": (lambda: ops.linear_d8(xb, pk_qkv, None, EPI_BF16, out=out3), T * D * 8, oct_f * 3 * D),
    "d8_proj_resid": (lambda: ops.linear_d8(xb, pk_proj, None, EPI_RESID, gamma=gamma, resid_in=x32, resid_out=res_out,
                                            branch_out=branch), T * D * (2 + 4 + 4 + 2), oct_f * D),
    "d8_fc1": (lambda: ops.linear_d8(xb, pk_fc1, None, EPI_BF16, out=out4), T * D * 10, oct_f * 4 * D),
    "d8_fc2_resid": (lambda: ops.linear_d8(h4, pk_fc2, None, EPI_RESID, gamma=gamma, resid_in=x32, resid_out=res_out,
                                           branch_out=branch), T * D * (8 + 4 + 4 + 2), oct_f * 4 * D),
    "d8_fc1_dgrad": (lambda: ops.linear_d8_dgrad(g4, pk_fc1), T * D * 10, oct_f * 4 * D),
    "d8_fc1_wgrad": (lambda: ops.linear_d8_wgrad(g4, xb, D, 4 * D), T * D * 10, oct_f * 4 * D),
    "dense_qkv": (lambda: ops.linear_dense(xb, dq.w, 3 * D, D, None, EPI_BF16, out=out3), T * D * 8, 2.0 * T * D * 3 * D),
    "dense_proj_resid": (lambda: ops.linear_dense(xb, dp.w, D, D, None, EPI_RESID, gamma=gamma, resid_in=x32,
                                                  resid_out=res_out, branch_out=branch), T * D * 12, 2.0 * T * D * D),
    "dense_fc1_gelu": (lambda: ops.linear_dense(xb, d1.w, 4 * D, D, None, EPI_GELU_BF16, out=out4, branch_out=g4),
                       T * D * 18, 2.0 * T * D * 4 * D),
    "dense_fc2_resid": (lambda: ops.linear_dense(h4, d2.w, D, 4 * D, None, EPI_RESID, gamma=gamma, resid_in=x32,
                                                 resid_out=res_out, branch_out=branch), T * D * 18, 2.0 * T * D * 4 * D),
    "dense_fc1_wgrad": (lambda: ops.linear_dense_wgrad(g4, xb, 4 * D, D), T * D * 10, 2.0 * T * D * 4 * D),
    "dense_fc2_wgrad": (lambda: ops.linear_dense_wgrad(dyb, h4, D, 4 * D), T * D * 10, 2.0 * T * D * 4 * D),
    "dense_qkv_wgrad": (lambda: ops.linear_dense_wgrad(qkv, xb, 3 * D, D), T * D * 8, 2.0 * T * D * 3 * D),
    "dense_proj_wgrad": (lambda: ops.linear_dense_wgrad(dyb, xb, D, D), T * D * 4, 2.0 * T * D * D),
    "dense_proj_plain": (lambda: ops.linear_dense(xb, dp.w, D, D, None, EPI_BF16, out=out1), T * D * 4, 2.0 * T * D * D),
    "dense_fc2_dgrad_geluBwd": (lambda: ops.linear_dense(dyb, d2.w_t, 4 * D, D, None, EPI_GELU_BWD, out=out4, gelu_pre=h4,
                                                         colsum=colsum4), T * D * 18, 2.0 * T * D * 4 * D),
    "dense_fc1_dgrad": (lambda: ops.linear_dense(g4, d1.w_t, D, 4 * D, None, EPI_BF16, out=out1), T * D * 10, 2.0 * T * D * 4 * D),
    "d8_qkv_headmajor": (lambda: ops.linear_d8(xb, pk_qkv, None, EPI_BF16, out=out3, head=(H, 3)), T * D * 8, oct_f * 3 * D),
    "d8_fc2_dgrad": (lambda: ops.linear_d8_dgrad(dyb, pk_fc2), T * D * 10, oct_f * 4 * D),
    "d8_qkv_wgrad": (lambda: ops.linear_d8_wgrad(qkv, xb, D, 3 * D), T * D * 8, oct_f * 3 * D),
}

only = [s for s in args.only.split(",") if s]
print(f"# batch {B}: T={T} tokens, D={D}; peaks: {peaks['hbm_gbs']:.0f} GB/s, {peaks['bf16_tflops']:.0f} TFLOP/s (burst)")
print(f"{'op':18s} {'us':>9s} {'GB/s':>8s} {'%hbm':>6s} {'TFLOP/s':>9s} {'%tc':>6s}")
for name, (fn, nbytes, flops) in CASES.items():
    if only and not any(name.startswith(o) for o in only):
        continue
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if args.profile:
        torch.cuda.profiler.start()
        fn()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        continue
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / args.iters
    gbs = nbytes / us / 1e3
    tf = flops / us / 1e6
    print(f"{name:18s} {us:9.1f} {gbs:8.0f} {100 * gbs / peaks['hbm_gbs']:6.1f} {tf:9.1f} {100 * tf / peaks['bf16_tflops']:6.1f}")
