"""Times the UNMODIFIED reference (davnords/octic-vits, vendored by tools/vendor_reference.sh into the git-ignored
baseline/_ref/) on the same B200 the product is measured on, next to this repo's kernels.

What it measures (VERDICT r01 "missing 2"):
  1. reference `hybrid_deit_huge_patch14` forward+backward, batch B, 224 px, bf16 autocast, CE loss -- eager and
     `torch.compile` (the DeiT recipe runs compile=True, deit/main.py:341-342); its stock code path = cuBLAS linears,
     SDPA, TritonGeluD8 (octic_vits/d8_gelu.py:103-482).  Protocol of experiments/complexity.py:40-56 with CUDA events.
  2. reference inference forward (BASELINE.json configs[1]) at batch 256, bf16 autocast, eager and compiled.
  3. Triton d8_gelu fwd / bwd vs octic_gelu_d8_fwd / _bwd at (B*N = 32 896 tokens, C = hidden/8 = 640), bf16.
  4. the reference's own bf16 D8 equivariance error of the 16-block octic trunk (experiments/test_equivariance.py
     :145-161 logic, all 8 group elements), the number ours must stay at.
Writes gpurun_out/reference_gpu.json (copied to profiles/r02_reference_gpu.json by hand after the run).

Nothing here is on the product path; the only stand-in is tools/timm_shim (timm==1.0.12 is not installed).
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
REF = ROOT / "baseline" / "_ref"
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools" / "timm_shim"))
sys.path.insert(0, str(REF))


def cuda_time(fn, warmup, iters):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return {"mean_ms": sum(ts) / len(ts), "median_ms": ts[len(ts) // 2], "min_ms": ts[0]}


def train_step_fn(model, img, tgt):
    params = [p for p in model.parameters() if p.requires_grad]

    def step():
        for p in params:
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = torch.nn.functional.cross_entropy(model(img), tgt)
        loss.backward()
        return loss
    return step


def bench_model(out, batch, infer_batch, iters, do_compile, compile_budget_s):
    import octic_vits.deit_models  # noqa: F401  (registers the factories in the shim registry)
    from timm.models import create_model
    torch.manual_seed(0)
    dev = torch.device("cuda", 0)
    model = create_model("hybrid_deit_huge_patch14", num_classes=1000).to(dev)
    n_params = sum(p.numel() for p in model.parameters())
    out["model"] = {"name": "hybrid_deit_huge_patch14", "params_M": n_params / 1e6}
    img = torch.randn(batch, 3, 224, 224, device=dev)
    tgt = torch.randint(0, 1000, (batch,), device=dev)

    # ---- training step, eager
    model.train()
    try:
        t = cuda_time(train_step_fn(model, img, tgt), 3, iters)
        out["train_eager"] = {"batch": batch, **t, "img_per_s": batch / t["median_ms"] * 1e3}
    except torch.OutOfMemoryError as e:
        out["train_eager"] = {"batch": batch, "error": f"OOM: {str(e)[:120]}"}
    print("train eager:", out["train_eager"], flush=True)
    torch.cuda.empty_cache()

    # ---- inference, eager (configs[1])
    model.eval()
    ximg = torch.randn(infer_batch, 3, 224, 224, device=dev)

    def infer():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return model(ximg)
    t = cuda_time(infer, 3, iters)
    out["infer_eager"] = {"batch": infer_batch, **t, "img_per_s": infer_batch / t["median_ms"] * 1e3}
    print("infer eager:", out["infer_eager"], flush=True)

    if do_compile:
        # torch.compile of the whole model, as deit/main.py:341-342 does.  Compilation of 32 blocks of tuple code can be
        # slow; give up (and say so) past the budget.
        t0 = time.time()
        try:
            cmodel = torch.compile(model)
            model.train()
            step = train_step_fn(cmodel, img, tgt)
            step()
            torch.cuda.synchronize()
            out["compile_s_train"] = time.time() - t0
            if out["compile_s_train"] < compile_budget_s:
                t = cuda_time(step, 3, iters)
                out["train_compiled"] = {"batch": batch, **t, "img_per_s": batch / t["median_ms"] * 1e3}
            else:
                out["train_compiled"] = {"error": f"compile took {out['compile_s_train']:.0f} s (> budget); one step ran"}
            print("train compiled:", out.get("train_compiled"), flush=True)
            model.eval()
            t1 = time.time()

            def cinfer():
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                    return cmodel(ximg)
            cinfer()
            torch.cuda.synchronize()
            out["compile_s_infer"] = time.time() - t1
            t = cuda_time(cinfer, 3, iters)
            out["infer_compiled"] = {"batch": infer_batch, **t, "img_per_s": infer_batch / t["median_ms"] * 1e3}
            print("infer compiled:", out["infer_compiled"], flush=True)
        except Exception as e:  # noqa: BLE001  (whatever Dynamo / Inductor / Triton raises is the result)
            out["compiled_error"] = f"{type(e).__name__}: {str(e)[:300]}"
            print("compiled: FAILED", out["compiled_error"], flush=True)
    del model
    torch.cuda.empty_cache()


def bench_equivariance(out):
    """Reference octic trunk (patch embed + pos + cls + 16 octic blocks), bf16 autocast vs fp32, under all of D8."""
    import octic_vits.deit_models  # noqa: F401
    from timm.models import create_model
    from octic_vits.d8_utils import convert_8tuple_to_5tuple, isotypic_dim_interpolation, interpolate_spatial_tuple
    from oracle import octic_oracle as O
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    img = torch.randn(2, 3, 224, 224, device=dev)

    def trunk(model, x):
        B, _, H, W = x.shape
        xs = model.patch_embed(x)
        pos = convert_8tuple_to_5tuple(isotypic_dim_interpolation(model.pos_embed, dim=0))
        pos = interpolate_spatial_tuple(xs, pos, H, W, model.patch_embed.patch_size)
        xs = tuple(a + v.flatten(0, 1) for a, v in zip(xs, pos))
        cls = tuple(model.cls_token[i].expand(B, *model.cls_token[i].shape[1:]) for i in range(5))
        xs = tuple(torch.cat((cls[i], xs[i]), dim=1) for i in range(5))
        for blk in model.blocks[:model.octic_equi_break_layer]:
            xs = blk(xs)
        return tuple(t.float() for t in xs)

    def rel(a, b):
        return float((a - b).norm() / b.norm().clamp_min(1e-12))

    for tag, randomize in (("default_init", False), ("o1_weights", True)):
        model = create_model("hybrid_deit_huge_patch14", num_classes=1000).to(dev).eval()
        if randomize:
            g = torch.Generator().manual_seed(5)
            with torch.no_grad():
                for n, p in model.named_parameters():
                    if not p.requires_grad:
                        continue
                    if "alpha" in n or "gamma" in n:
                        p.copy_((1.0 + 0.1 * torch.randn(p.shape, generator=g)).to(dev))
                    elif p.dim() >= 2 and "pos_embed" not in n and "cls" not in n:
                        p.copy_((torch.randn(p.shape, generator=g) / (p.shape[-1] ** 0.5)).to(dev))
        for mode in ("bf16", "fp32"):
            worst_rel, worst_abs = 0.0, 0.0
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode == "bf16"):
                base = trunk(model, img)
                scale = max(float(t.abs().max()) for t in base)
                for ge in O.GROUP:
                    moved = trunk(model, O.image_action(ge, img).contiguous())
                    want = O.token_action(ge, base, has_cls=True)
                    for a, b in zip(moved, want):
                        worst_rel = max(worst_rel, rel(a, b))
                        worst_abs = max(worst_abs, float((a - b).abs().max()))
            out[f"ref_trunk_equivariance_{tag}_{mode}"] = {"rel_l2_max_over_group": worst_rel, "max_abs": worst_abs,
                                                           "max_abs_of_output": scale}
            print(f"reference equivariance {tag} {mode}:", out[f"ref_trunk_equivariance_{tag}_{mode}"], flush=True)
        del model
        torch.cuda.empty_cache()


def bench_gelu(out, tokens, C, iters):
    """Triton d8_gelu fwd/bwd (reference) vs octic_gelu_d8_fwd/bwd (this repo), bf16, hidden = 8 C."""
    from octic_vits.d8_gelu import d8_gelu_fwd, d8_gelu_bwd
    from octic_vits_b200 import ops
    dev = torch.device("cuda", 0)
    torch.manual_seed(1)
    packed = torch.randn(tokens, 8 * C, device=dev).to(torch.bfloat16)
    gpacked = torch.randn(tokens, 8 * C, device=dev).to(torch.bfloat16)
    # the reference takes five contiguous tensors [1, T, C] x4 and [1, T, 2, 2C]
    five = tuple(packed[:, i * C:(i + 1) * C].contiguous().unsqueeze(0) for i in range(4)) + \
        (packed[:, 4 * C:].reshape(1, tokens, 2, 2 * C).contiguous(),)
    gfive = tuple(gpacked[:, i * C:(i + 1) * C].contiguous().unsqueeze(0) for i in range(4)) + \
        (gpacked[:, 4 * C:].reshape(1, tokens, 2, 2 * C).contiguous(),)
    res = {"tokens": tokens, "C": C, "dtype": "bf16"}
    t = cuda_time(lambda: d8_gelu_fwd(*five), 5, iters)
    res["triton_fwd_us"] = t["median_ms"] * 1e3
    t = cuda_time(lambda: d8_gelu_bwd(*gfive, *five), 5, iters)
    res["triton_bwd_us"] = t["median_ms"] * 1e3
    t = cuda_time(lambda: ops.gelu_d8_fwd(packed), 5, iters)
    res["octic_fwd_us"] = t["median_ms"] * 1e3
    t = cuda_time(lambda: ops.gelu_d8_bwd(gpacked, packed), 5, iters)
    res["octic_bwd_us"] = t["median_ms"] * 1e3
    # values: both against each other (bf16 I/O, fp32 maths)
    y_ref = d8_gelu_fwd(*five)
    y_ours = ops.gelu_d8_fwd(packed)
    y_ref_packed = torch.cat([y.reshape(tokens, -1) for y in y_ref], dim=1)
    res["fwd_max_abs_diff_vs_triton"] = float((y_ref_packed.float() - y_ours.float()).abs().max())
    fwd_bytes, bwd_bytes = tokens * 8 * C * 2 * 2, tokens * 8 * C * 2 * 3
    for k, b in (("triton_fwd", fwd_bytes), ("triton_bwd", bwd_bytes), ("octic_fwd", fwd_bytes), ("octic_bwd", bwd_bytes)):
        res[k + "_GBps"] = b / (res[k + "_us"] * 1e-6) / 1e9
    res["speedup_fwd"] = res["triton_fwd_us"] / res["octic_fwd_us"]
    res["speedup_bwd"] = res["triton_bwd_us"] / res["octic_bwd_us"]
    out["gelu_d8"] = res
    print("gelu:", res, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--infer-batch", type=int, default=256)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--no-compile", action="store_true")
    ap.add_argument("--compile-budget-s", type=float, default=420.0)
    ap.add_argument("--skip", default="", help="comma list of: model, equivariance, gelu")
    ap.add_argument("--out", default="gpurun_out/reference_gpu.json")
    args = ap.parse_args()
    if not REF.is_dir():
        raise SystemExit("baseline/_ref is missing: run tools/vendor_reference.sh in the build container first")
    skip = set(args.skip.split(","))
    out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "autocast": "bfloat16",
           "reference": "davnords/octic-vits, unmodified, baseline/_ref (timm stand-in: tools/timm_shim)"}
    torch.set_float32_matmul_precision("high")      # as experiments/complexity.py:92
    for name, fn in (("gelu", lambda: bench_gelu(out, 32896, 640, 50)),
                     ("equivariance", lambda: bench_equivariance(out)),
                     ("model", lambda: bench_model(out, args.batch, args.infer_batch, args.iters, not args.no_compile,
                                                   args.compile_budget_s))):
        if name in skip:
            continue
        try:
            fn()
        except Exception as e:  # noqa: BLE001
            out[name + "_error"] = f"{type(e).__name__}: {str(e)[:400]}"
            print(name, "FAILED:", out[name + "_error"], flush=True)
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
