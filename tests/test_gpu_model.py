"""GPU parity of the reference-facing modules (octic_vits_b200.layers / .model) against
 (a) the committed golden vectors produced by the reference itself (tests/golden, tools/make_golden.py) and
 (b) the oracle at larger, seeded sizes (BASELINE.json configs[0]: hybrid ViT-S/16, b8),
plus the reference's equivariance / invariance properties (experiments/test_equivariance.py logic).

Tolerance: the kernels compute like the reference under torch.autocast(bfloat16) (bf16 GEMM I/O, fp32 accumulate and
residual) while goldens/oracle are fp32, so comparisons use a relative L2 error bound of 2e-2 per tensor (bf16 has
2^-9 = 2e-3 relative rounding; a handful of chained roundings) and a max-abs bound of 6e-2 of the tensor's max.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import octic_oracle as O

if torch.cuda.is_available():
    from octic_vits_b200 import layers as L
    from octic_vits_b200.deit_models import create_model
    from octic_vits_b200.model import OcticVisionTransformer
    from octic_vits_b200._lib import OcticError

DEV = "cuda"


def rel_err(got, want):
    got, want = got.float(), want.float()
    return float((got - want).norm() / want.norm().clamp_min(1e-12))


def check(got, want, rel=2e-2, mx=6e-2, what=""):
    got, want = got.float().cpu(), want.float().cpu()
    assert got.shape == want.shape, (what, got.shape, want.shape)
    r = rel_err(got, want)
    m = float((got - want).abs().max() / want.abs().max().clamp_min(1e-12))
    assert r < rel and m < mx, f"{what}: rel L2 {r:.3e} (< {rel}), max-abs/max {m:.3e} (< {mx})"


def to_dev(xs):
    return tuple(x.to(DEV) for x in xs)


def run_with_grads(module, xs, gout):
    xs = tuple(x.clone().to(DEV).requires_grad_(True) for x in xs)
    out = module(xs)
    loss = sum((o.float() * g.to(DEV)).sum() for o, g in zip(out, gout))
    loss.backward()
    return out, [x.grad for x in xs], {k: p.grad for k, p in module.named_parameters() if p.grad is not None}


LAYER_CASES = {
    "linear_d8": lambda fx: L.LinearD8(64, 128),
    "layernorm_d8": lambda fx: L.LayerNormD8(64),
    "gelu_d8": lambda fx: L.TritonGeluD8(),
    "mlp_d8": lambda fx: L.MlpD8(64, 256),
    "attention_d8": lambda fx: L.AttentionD8(128, num_heads=2, qkv_bias=True),
    "block_deit_d8": lambda fx: L.Layer_scale_init_BlockD8(128, 2, qkv_bias=True),
    "block_dinov2_d8": lambda fx: L.BlockD8(128, 2, qkv_bias=True, init_values=0.5),
}


@pytest.mark.parametrize("name", list(LAYER_CASES))
def test_layer_matches_reference_golden(golden, name):
    fx = golden(name)
    mod = LAYER_CASES[name](fx).to(DEV)
    if "sd" in fx:
        mod.load_state_dict(fx["sd"], strict=True)
    out, gin, gpar = run_with_grads(mod, fx["in"], fx["gout"])
    for i, (a, b) in enumerate(zip(out, fx["out"])):
        check(a.detach(), b, what=f"{name} out[{i}]")
    for i, (a, b) in enumerate(zip(gin, fx["gin"])):
        check(a, b, rel=3e-2, mx=8e-2, what=f"{name} gin[{i}]")
    for k, g in fx.get("gparams", {}).items():
        check(gpar[k], g, rel=3e-2, mx=8e-2, what=f"{name} grad {k}")


def test_power_spectrum_golden(golden):
    fx = golden("power_spectrum")
    inv = L.PowerSpectrumInvariant(64)
    xs = tuple(x.clone().to(DEV).requires_grad_(True) for x in fx["in"])
    y = inv(xs)
    check(y.detach(), fx["out"], rel=5e-3, mx=1e-2, what="power spectrum")
    (y.float() * fx["gout"].to(DEV)).sum().backward()
    for x, g in zip(xs, fx["gin"]):
        check(x.grad, g, rel=5e-3, mx=1e-2, what="power spectrum grad")


@pytest.mark.parametrize("tag", ["image", "tokens"])
def test_isotypic_to_patch_golden(golden, tag):
    """IsotypicToPatchD8 (d8_layers.py:499-588) vs the reference's own output and autograd gradients."""
    fx = golden("isotypic_to_patch")[tag]
    mod = L.IsotypicToPatchD8(dim=64, **fx["kw"]).to(DEV)
    mod.load_state_dict(fx["sd"], strict=True)
    xs = tuple(x.clone().to(DEV).requires_grad_(True) for x in fx["in"])
    y = mod(xs)
    check(y.detach(), fx["out"], what="patches")
    (y * fx["gout"].to(DEV)).sum().backward()
    for i, (a, b) in enumerate(zip(xs, fx["gin"])):
        check(a.grad, b, rel=3e-2, mx=8e-2, what=f"gin[{i}]")
    for k, g in fx["gparams"].items():
        check(dict(mod.named_parameters())[k].grad, g, rel=3e-2, mx=8e-2, what=f"grad {k}")


def test_dense_block_golden(golden):
    fx = golden("dense_block_deit")
    blk = L.Layer_scale_init_Block(64, 2, qkv_bias=True).to(DEV)
    blk.load_state_dict(fx["sd"], strict=True)
    x = fx["in"].clone().to(DEV).requires_grad_(True)
    y = blk(x)
    check(y.detach(), fx["out"], what="dense block")
    (y * fx["gout"].to(DEV)).sum().backward()
    check(x.grad, fx["gin"], rel=3e-2, mx=8e-2, what="dense block gin")
    for k, g in fx["gparams"].items():
        check(dict(blk.named_parameters())[k].grad, g, rel=3e-2, mx=8e-2, what=f"dense block grad {k}")


def build_from_cfg(cfg, sd, **extra):
    model = OcticVisionTransformer(img_size=cfg["img_size"], patch_size=cfg["patch"], embed_dim=cfg["embed_dim"],
                                   depth=cfg["depth"], num_heads=cfg["num_heads"], num_classes=cfg["num_classes"],
                                   qkv_bias=True, invariant=cfg.get("invariant", False),
                                   standard_block_layers=L.Layer_scale_init_Block,
                                   octic_block_layers=L.Layer_scale_init_BlockD8, **extra).to(DEV)
    model.load_state_dict(sd, strict=True)
    return model


@pytest.mark.parametrize("tag", ["hybrid", "invariant"])
def test_whole_model_golden(golden, tag):
    fx = golden(f"model_{tag}")
    model = build_from_cfg(fx["cfg"], fx["sd"]).eval()
    img = fx["img"].to(DEV)
    with torch.no_grad():
        pe = model.patch_embed(img)
        for a, b in zip(pe, fx["patch_embed"]):
            check(a, b, what="patch embed")
        tok = L.OF.unpack_five(model.embed_tokens_packed(img))
        for a, b in zip(tok, fx["tokens0"]):
            check(a, b, what="tokens")
        trunk = L.OF.unpack_five(model.forward_trunk_packed(img))
        for a, b in zip(trunk, fx["trunk"]):
            check(a, b, what="trunk")
        check(model(img), fx["logits"], rel=3e-2, mx=8e-2, what="logits")
    model.train()
    out = model(img)
    (out * fx["loss_weight"].to(DEV)).sum().backward()
    params = dict(model.named_parameters())
    for k, g in fx["gparams"].items():
        assert params[k].grad is not None, k
        upstream_of_abs = tag == "invariant" and not (k.startswith("blocks.3") or k.startswith("head") or k.startswith("invariant_proj"))
        if upstream_of_abs:
            # d|x|/dx = sign(x): a bf16-level perturbation of the trunk flips the sign of near-zero entries, and in
            # this 272-element fixture one flip moves the gradient by ~12 % (measured: 1 % input noise -> 16 % change
            # of the fp32 oracle's own gradient).  The same kernels are held to 6e-2 by the hybrid model above and
            # by test_power_spectrum_and_bridge; here only gross errors are excluded.
            check(params[k].grad, g, rel=0.8, mx=1.5, what=f"grad {k}")
        else:
            # deep-chain bf16 gradients: the reference's own bf16-autocast run deviates from its fp32 run by up to
            # 4.3e-2 rel. L2 on these tensors; 8-entry vectors of this tiny fixture scatter
            # around that level
            check(params[k].grad, g, rel=6e-2, mx=1.5e-1, what=f"grad {k}")
    # frozen zero cls tokens of the non-A1 irreps get no gradient (reference model.py:99-106)
    for i in range(1, 5):
        assert params.get(f"cls_token.{i}") is None or params[f"cls_token.{i}"].grad is None


def test_timm_default_blocks_golden(golden):
    fx = golden("model_timm_default")
    cfg = fx["cfg"]
    model = OcticVisionTransformer(img_size=cfg["img_size"], patch_size=cfg["patch"], embed_dim=cfg["embed_dim"],
                                   depth=cfg["depth"], num_heads=cfg["num_heads"], num_classes=cfg["num_classes"],
                                   init_scale=0.7).to(DEV).eval()
    model.load_state_dict(fx["sd"], strict=True)
    with torch.no_grad():
        check(model(fx["img"].to(DEV)), fx["logits"], rel=3e-2, mx=8e-2, what="timm-default logits")


def _randomize(model, seed=0, std=0.5):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if not p.requires_grad:
                continue
            if "alpha" in name or "gamma" in name or ("norm" in name and name.endswith("weight")):
                p.copy_((1.0 + 0.2 * torch.randn(p.shape, generator=g)).to(p.device))
            elif p.dim() >= 2:
                fan_in = p[0].numel()
                p.copy_((torch.randn(p.shape, generator=g) * (std / fan_in ** 0.5)).to(p.device))
            else:
                p.copy_((0.1 * torch.randn(p.shape, generator=g)).to(p.device))


def test_config1_vit_s16_against_oracle():
    """BASELINE.json configs[0]: hybrid octic ViT-S/16 (embed 384, depth 12, heads 6), b8, 224 px; layer scale ~ 1 and
    O(1) weights so that no block is ~identity."""
    torch.manual_seed(0)
    model = create_model("hybrid_deit_small_patch16", num_classes=100).to(DEV).eval()
    _randomize(model, seed=1)
    img = torch.randn(8, 3, 224, 224, generator=torch.Generator().manual_seed(2))
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    want_trunk = O.octic_vit_forward(img, sd, patch=16, depth=12, num_heads=6, return_trunk=True)
    want = O.octic_vit_forward(img, sd, patch=16, depth=12, num_heads=6)
    with torch.no_grad():
        got_trunk = L.OF.unpack_five(model.forward_trunk_packed(img.to(DEV)))
        got = model(img.to(DEV))
    for a, b in zip(got_trunk, want_trunk):
        check(a, b, rel=3e-2, mx=1e-1, what="S/16 trunk")
    check(got, want, rel=5e-2, mx=1.5e-1, what="S/16 logits")


def test_trunk_equivariance_and_logit_invariance():
    """experiments/test_equivariance.py:145-161, 302-315 on the GPU path: g . f(x) == f(g . x) for the octic trunk of
    the hybrid model, and logits of the invariant model are the same for all 8 image transforms.  The reference's
    fp32 level is ~1e-5 of |out|; the bf16 path is held to 2e-2 relative L2."""
    torch.manual_seed(0)
    hyb = OcticVisionTransformer(img_size=64, patch_size=8, embed_dim=128, depth=4, num_heads=2, num_classes=10,
                                 qkv_bias=True, standard_block_layers=L.Layer_scale_init_Block,
                                 octic_block_layers=L.Layer_scale_init_BlockD8).to(DEV).eval()
    inv = OcticVisionTransformer(img_size=64, patch_size=8, embed_dim=128, depth=4, num_heads=2, num_classes=10,
                                 qkv_bias=True, invariant=True, standard_block_layers=L.Layer_scale_init_Block,
                                 octic_block_layers=L.Layer_scale_init_BlockD8).to(DEV).eval()
    _randomize(hyb, seed=3)
    _randomize(inv, seed=4)
    img = torch.randn(5, 3, 64, 64, device=DEV)
    with torch.no_grad():
        base = L.OF.unpack_five(hyb.forward_trunk_packed(img))
        base_logits = inv(img)
        flipped_channels = inv(img.flip(1))
        assert rel_err(flipped_channels, base_logits) > 1e-2      # the guard of test_equivariance.py:314-315
        for g in O.GROUP:
            moved = L.OF.unpack_five(hyb.forward_trunk_packed(O.image_action(g, img).contiguous()))
            want = O.token_action(g, tuple(t.float() for t in base), has_cls=True)
            for a, b in zip(moved, want):
                check(a, b, rel=2e-2, mx=8e-2, what=f"equivariance {g}")
            check(inv(O.image_action(g, img).contiguous()), base_logits, rel=2e-2, mx=8e-2, what=f"invariance {g}")


def test_drop_path_matches_oracle_with_injected_mask(golden):
    fx = golden("block_deit_d8")
    blk = L.Layer_scale_init_BlockD8(128, 2, qkv_bias=True, drop_path=0.5).to(DEV).train()
    blk.load_state_dict(fx["sd"], strict=True)
    masks = [torch.tensor([2.0, 0.0], device=DEV), torch.tensor([0.0, 2.0], device=DEV)]
    it = iter(masks)
    blk.drop_path.sample = lambda batch, device: next(it)
    out = blk(to_dev(fx["in"]))
    want = O.block_d8(fx["in"], fx["sd"], "", 2, "deit", masks[0].cpu(), masks[1].cpu())
    for a, b in zip(out, want):
        check(a.detach(), b, what="drop path")


def test_error_behaviour():
    with pytest.raises(ValueError):
        L.LinearD8(60, 128)
    with pytest.raises(ValueError):
        L.AffineD8(12)
    with pytest.raises(NotImplementedError):
        L.AttentionD8(128, num_heads=2, rope=object())
    lin = L.LinearD8(64, 64)
    xs = tuple(torch.randn(1, 3, 8) for _ in range(4)) + (torch.randn(1, 3, 2, 16),)
    with pytest.raises(OcticError):
        lin(xs)                       # CPU tensors: there is no CPU fallback
    with pytest.raises(AssertionError):
        lin.to(DEV)(to_dev(xs)[:4])   # not a 5-tuple
    blk = L.NestedTensorBlockD8(64, 2).to(DEV)
    with pytest.raises(AssertionError):
        blk(torch.zeros(1, device=DEV))
    out = blk([to_dev(xs), to_dev(xs)])
    assert isinstance(out, list) and len(out) == 2 and len(out[0]) == 5


def test_graphed_train_step_matches_eager():
    """parallel.GraphedTrainStep (fwd + loss + bwd replayed from a CUDA graph) leaves the same loss and gradients in the
    flat buffer as the eager step on the same inputs."""
    from octic_vits_b200.parallel import FlatGrads, GraphedTrainStep
    torch.manual_seed(3)
    model = OcticVisionTransformer(img_size=64, patch_size=16, embed_dim=128, depth=4, num_heads=2, num_classes=10,
                                   qkv_bias=True).to(DEV).train()
    fg = FlatGrads(model.parameters())
    img = torch.randn(4, 3, 64, 64, device=DEV)
    tgt = torch.randint(0, 10, (4,), device=DEV)
    step = GraphedTrainStep(model, fg, img.shape, warmup=2)
    assert step.graphed, getattr(step, "capture_error", "")
    loss_g = float(step(img, tgt))
    grads_g = fg.flat.clone()
    eager = GraphedTrainStep(model, fg, img.shape, use_graph=False)
    loss_e = float(eager(img, tgt))
    assert abs(loss_g - loss_e) <= 1e-5 * max(1.0, abs(loss_e))
    # wgrad accumulates with fp32 red.add in a non-deterministic order: equal up to summation order
    assert rel_err(grads_g, fg.flat) < 1e-4
    # a second replay with new inputs changes the result (the static inputs are really read)
    img2 = torch.randn_like(img)
    loss_g2 = float(step(img2, tgt))
    assert loss_g2 != loss_g
    # fused gradient accumulation (wgrad kernels add straight into the flat .grad views; FlatGrads switched it on)
    # against plain autograd accumulation
    from octic_vits_b200 import functional as OF
    assert OF.ACCUMULATE_INTO_GRAD
    OF.ACCUMULATE_INTO_GRAD = False
    try:
        float(eager(img2, tgt))
        plain = fg.flat.clone()
    finally:
        OF.ACCUMULATE_INTO_GRAD = True
    float(eager(img2, tgt))
    assert rel_err(fg.flat, plain) < 1e-4
    # staged input pipeline (pinned host batch -> copy stream -> static inputs) gives the same step
    step.stage(img2.cpu().pin_memory(), tgt.cpu().pin_memory())
    assert abs(float(step.run()) - loss_g2) <= 1e-5 * max(1.0, abs(loss_g2))


def test_graphed_step_sees_parameter_updates_of_any_optimizer():
    """The weight packs live inside the captured graph (repack_weights, default): after a plain torch optimizer changed
    the parameters, a replay computes with the NEW values -- loss and gradients equal the eager step at those values."""
    from octic_vits_b200.parallel import FlatGrads, GraphedTrainStep
    torch.manual_seed(4)
    model = OcticVisionTransformer(img_size=64, patch_size=16, embed_dim=128, depth=4, num_heads=2, num_classes=10,
                                   qkv_bias=True, init_scale=0.5).to(DEV).train()
    fg = FlatGrads(model.parameters())
    img = torch.randn(4, 3, 64, 64, device=DEV)
    tgt = torch.randint(0, 10, (4,), device=DEV)
    step = GraphedTrainStep(model, fg, img.shape, warmup=2)
    eager = GraphedTrainStep(model, fg, img.shape, use_graph=False)
    assert step.graphed, getattr(step, "capture_error", "")
    sgd = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=0.5)
    loss0 = float(step(img, tgt))
    sgd.step()                                        # reads the flat .grad views, bumps the parameters' versions
    loss_e = float(eager(img, tgt))
    grads_e = fg.flat.clone()
    loss_g = float(step(img, tgt))
    assert abs(loss_e - loss0) > 1e-3                 # the update really moved the loss
    assert abs(loss_g - loss_e) <= 1e-4 * max(1.0, abs(loss_e))
    assert rel_err(fg.flat, grads_e) < 1e-3


def test_failed_graph_capture_falls_back_to_a_correct_eager_step():
    """ADVICE r01 (parallel.py): kernels issued during a FAILED capture are only recorded, so the bf16 weight packs
    allocated in that attempt hold uninitialised memory under valid-looking cache keys.  A loss_fn that synchronises
    (.item()) makes the capture fail; the eager fallback must then equal a plain eager step, not run on garbage packs."""
    from octic_vits_b200.parallel import FlatGrads, GraphedTrainStep
    torch.manual_seed(7)
    model = OcticVisionTransformer(img_size=64, patch_size=16, embed_dim=128, depth=4, num_heads=2, num_classes=10,
                                   qkv_bias=True, init_scale=0.5).to(DEV).train()
    fg = FlatGrads(model.parameters())
    img = torch.randn(4, 3, 64, 64, device=DEV)
    tgt = torch.randint(0, 10, (4,), device=DEV)
    ref = GraphedTrainStep(model, fg, img.shape, use_graph=False)
    loss_ref = float(ref(img, tgt))
    grads_ref = fg.flat.clone()

    def syncing_loss(logits, target):
        loss = torch.nn.functional.cross_entropy(logits, target)
        loss.item()                                   # illegal during stream capture
        return loss
    # every parameter gets a new version, so the failed capture has to re-pack (into never-executed buffers)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.0)
    step = GraphedTrainStep(model, fg, img.shape, loss_fn=syncing_loss, warmup=1)
    assert not step.graphed and "capture_error" in step.__dict__
    loss_fb = float(step(img, tgt))
    assert abs(loss_fb - loss_ref) <= 1e-5 * max(1.0, abs(loss_ref))
    assert rel_err(fg.flat, grads_ref) < 1e-4


def test_fused_grad_accumulation_is_scoped_to_the_flat_buffer_owner():
    """ADVICE r01 (parallel.py:47): a FlatGrads must not change the autograd semantics of models it does not manage --
    tensor hooks on their weights still see the gradient and .grad is produced by autograd."""
    from octic_vits_b200.parallel import FlatGrads
    torch.manual_seed(8)
    mk = lambda: OcticVisionTransformer(img_size=32, patch_size=16, embed_dim=64, depth=2, num_heads=2, num_classes=4,
                                        qkv_bias=True, init_scale=0.5).to(DEV).train()
    managed, free = mk(), mk()
    fg = FlatGrads(managed.parameters())
    seen = []
    w = free.blocks[0].attn.qkv.lin_E.weight
    w.register_hook(lambda g: seen.append(g.detach().clone()))
    img = torch.randn(2, 3, 32, 32, device=DEV)
    free(img).sum().backward()
    assert len(seen) == 1 and w.grad is not None and torch.equal(seen[0], w.grad)
    fg.zero()
    managed(img).sum().backward()
    wm = managed.blocks[0].attn.qkv.lin_E.weight
    assert wm.grad.data_ptr() == fg.flat[fg.offsets[[id(p) for p in fg.params].index(id(wm))]:].data_ptr()
    assert float(wm.grad.abs().sum()) > 0
    fg.close()
    assert not any(getattr(p, "_octic_accumulate", False) for p in managed.parameters())


def test_headline_model_equivariance_report(capsys):
    """BASELINE.json metric 'D8 equiv error' on the HEADLINE model: hybrid octic ViT-H/14 (DeiT-III init, layer scale
    1e-4 as shipped, and an O(1)-weights variant), bf16 GPU path, 224 px.  Equivariance of the 16-block octic trunk under
    all 8 elements of D8 (experiments/test_equivariance.py:145-161 logic) and invariance of the logits of the invariant
    ViT-H/14.  The numbers are printed (and written to gpurun_out/equivariance_h14.json when that directory exists) so
    that every GPU test run records them; bounds: 2e-2 relative L2 (bf16 path; the reference's fp32 level is 1.1e-6
    max-abs)."""
    import json
    import os
    torch.manual_seed(0)
    report = {}
    img = torch.randn(2, 3, 224, 224, device=DEV)
    for tag, randomize in (("default_init", False), ("o1_weights", True)):
        hyb = create_model("hybrid_deit_huge_patch14", num_classes=1000).to(DEV).eval()
        if randomize:
            _randomize(hyb, seed=5)
        worst_rel, worst_abs = 0.0, 0.0
        with torch.no_grad():
            base = tuple(t.float() for t in L.OF.unpack_five(hyb.forward_trunk_packed(img)))
            scale = max(float(t.abs().max()) for t in base)
            for g in O.GROUP:
                moved = L.OF.unpack_five(hyb.forward_trunk_packed(O.image_action(g, img).contiguous()))
                want = O.token_action(g, base, has_cls=True)
                for a, b in zip(moved, want):
                    worst_rel = max(worst_rel, rel_err(a, b))
                    worst_abs = max(worst_abs, float((a.float() - b).abs().max()))
        report[f"trunk_equivariance_{tag}"] = {"rel_l2_max_over_group": worst_rel, "max_abs": worst_abs, "max_abs_of_output": scale}
        # the REFERENCE's own level at the same dtype (last row): the oracle restatement of the
        # reference path under bf16 autocast, same weights, same images, on the host
        sd = {k: v.detach().cpu() for k, v in hyb.state_dict().items()}
        del hyb
        icpu = img.cpu()
        ref_rel, ref_abs = 0.0, 0.0
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            rbase = tuple(t.float() for t in O.octic_vit_forward(icpu, sd, patch=14, depth=32, num_heads=16, return_trunk=True))
            for g in O.GROUP:
                moved = O.octic_vit_forward(O.image_action(g, icpu).contiguous(), sd, patch=14, depth=32, num_heads=16,
                                            return_trunk=True)
                want = O.token_action(g, rbase, has_cls=True)
                for a, b in zip(moved, want):
                    ref_rel = max(ref_rel, rel_err(a, b))
                    ref_abs = max(ref_abs, float((a.float() - b).abs().max()))
        report[f"trunk_equivariance_{tag}"].update({"reference_bf16_rel_l2": ref_rel, "reference_bf16_max_abs": ref_abs})
        assert worst_rel < 2e-2, report
        # "must stay at the reference's level": within 2x of what the reference's own bf16 arithmetic gives
        assert worst_rel <= 2.0 * ref_rel + 1e-5, report
    inv = OcticVisionTransformer(img_size=224, patch_size=14, embed_dim=1280, depth=32, num_heads=16, num_classes=1000,
                                 qkv_bias=True, invariant=True, standard_block_layers=L.Layer_scale_init_Block,
                                 octic_block_layers=L.Layer_scale_init_BlockD8).to(DEV).eval()
    _randomize(inv, seed=6)
    with torch.no_grad():
        base_logits = inv(img).float()
        worst = max(rel_err(inv(O.image_action(g, img).contiguous()), base_logits) for g in O.GROUP)
    report["logit_invariance_o1_weights"] = {"rel_l2_max_over_group": worst, "max_abs_of_output": float(base_logits.abs().max())}
    assert worst < 2e-2, report
    with capsys.disabled():
        print("\nD8 equivariance report (hybrid / invariant ViT-H/14, bf16 GPU path):", json.dumps(report))
    if os.path.isdir("gpurun_out"):
        with open("gpurun_out/equivariance_h14.json", "w") as f:
            json.dump(report, f, indent=1)
