"""Pins oracle/octic_oracle.py against vectors produced by the reference itself (tools/make_golden.py)."""
import pytest
import torch

from oracle import octic_oracle as O

TOL = dict(rtol=1e-5, atol=2e-5)


def close(a, b, **kw):
    tol = dict(TOL)
    tol.update(kw)
    torch.testing.assert_close(a, b, **tol)


def test_gelu_known_answer(golden):
    fx = golden("gelu_kat")
    out = O.gelu_d8_eight(fx["in"])
    for a, b in zip(out, fx["out"]):
        close(a, b, atol=1e-6)
    # the survey's printed KAT, fp32
    want = [1.445854, 0.349454, 0.318992, 0.289325, 0.991137, 0.960675, 0.931008, 0.479196]
    got = [float(t) for t in out]
    assert got == pytest.approx(want, abs=2e-6)


def test_transforms_and_tuple_maps(golden):
    fx = golden("transforms")
    for a, b in zip(O.isotypic_to_regular(fx["in"]), fx["i2r"]):
        close(a, b)
    for a, b in zip(O.regular_to_isotypic(fx["in"]), fx["r2i"]):
        close(a, b)
    five = O.eight_to_five(fx["in"])
    for a, b in zip(five, fx["five"]):
        assert torch.equal(a, b)
    for a, b in zip(O.five_to_eight(five), fx["in"]):
        assert torch.equal(a, b)
    # packed rows round trip
    for a, b in zip(O.unpack_rows(O.pack_rows(five)), five):
        assert torch.equal(a, b)


def test_group_actions(golden):
    fx = golden("actions")
    for g in O.GROUP:
        assert torch.equal(O.image_action(g, fx["img"]), fx["image"][g])
        for a, b in zip(O.isotypic_action(g, fx["xs8"]), fx["isotypic"][g]):
            assert torch.equal(a, b)
        got = O.five_to_eight(O.token_action(g, O.eight_to_five(fx["xs8"]), has_cls=False))
        for a, b in zip(got, fx["spatial_isotypic"][g]):
            assert torch.equal(a, b)


def _grads(fn, xs, gout, params):
    xs = tuple(x.clone().requires_grad_(True) for x in xs)
    params = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    out = fn(xs, params)
    loss = sum((o * g).sum() for o, g in zip(out, gout))
    loss.backward()
    return out, [x.grad for x in xs], {k: v.grad for k, v in params.items()}


@pytest.mark.parametrize("name,fn", [
    ("linear_d8", lambda xs, w, fx: O.linear_d8(xs, w, "")),
    ("layernorm_d8", lambda xs, w, fx: O.layernorm_d8(xs, w, "")),
    ("gelu_d8", lambda xs, w, fx: O.gelu_d8(xs)),
    ("mlp_d8", lambda xs, w, fx: O.mlp_d8(xs, w, "")),
    ("attention_d8", lambda xs, w, fx: O.attention_d8(xs, w, "", fx["num_heads"])),
    ("block_deit_d8", lambda xs, w, fx: O.block_d8(xs, w, "", fx["num_heads"], "deit")),
    ("block_dinov2_d8", lambda xs, w, fx: O.block_d8(xs, w, "", fx["num_heads"], "dinov2")),
])
def test_layers_forward_and_backward(golden, name, fn):
    fx = golden(name)
    sd = fx.get("sd", {})
    out, gin, gpar = _grads(lambda xs, w: fn(xs, w, fx), fx["in"], fx["gout"], sd)
    for a, b in zip(out, fx["out"]):
        close(a.detach(), b, rtol=2e-5, atol=5e-5)
    for a, b in zip(gin, fx["gin"]):
        close(a, b, rtol=1e-4, atol=2e-4)
    for k, g in fx.get("gparams", {}).items():
        close(gpar[k], g, rtol=1e-4, atol=5e-4)


def test_power_spectrum(golden):
    fx = golden("power_spectrum")
    xs = tuple(x.clone().requires_grad_(True) for x in fx["in"])
    y = O.power_spectrum(xs)
    close(y.detach(), fx["out"])
    (y * fx["gout"]).sum().backward()
    for x, g in zip(xs, fx["gin"]):
        close(x.grad, g)


def test_dense_block(golden):
    fx = golden("dense_block_deit")
    x = fx["in"].clone().requires_grad_(True)
    w = {k: v.clone().requires_grad_(True) for k, v in fx["sd"].items()}
    y = O.dense_block(x, w, "", fx["num_heads"], eps=1e-5)   # Layer_scale_init_Block default norm_layer=nn.LayerNorm
    close(y.detach(), fx["out"], rtol=2e-5, atol=5e-5)
    (y * fx["gout"]).sum().backward()
    close(x.grad, fx["gin"], rtol=1e-4, atol=2e-4)
    for k, g in fx["gparams"].items():
        close(w[k].grad, g, rtol=1e-4, atol=5e-4)


@pytest.mark.parametrize("tag", ["hybrid", "invariant"])
def test_whole_model(golden, tag):
    fx = golden(f"model_{tag}")
    cfg = fx["cfg"]
    pe = O.patch_embed_d8(fx["img"], fx["sd"], "patch_embed.", cfg["patch"])
    for a, b in zip(pe, fx["patch_embed"]):
        close(a, b, rtol=2e-5, atol=5e-5)
    tok = O.embed_tokens(fx["img"], fx["sd"], cfg["patch"])
    for a, b in zip(tok, fx["tokens0"]):
        close(a, b, rtol=2e-5, atol=5e-5)
    trunk = O.octic_vit_forward(fx["img"], fx["sd"], patch=cfg["patch"], depth=cfg["depth"],
                                num_heads=cfg["num_heads"], invariant=cfg["invariant"], return_trunk=True)
    for a, b in zip(trunk, fx["trunk"]):
        close(a, b, rtol=1e-4, atol=2e-4)
    logits = O.octic_vit_forward(fx["img"], fx["sd"], patch=cfg["patch"], depth=cfg["depth"],
                                 num_heads=cfg["num_heads"], invariant=cfg["invariant"])
    close(logits, fx["logits"], rtol=1e-4, atol=5e-4)
    # parameter gradients of the training-mode loss (drop_path = 0, so train == eval maths)
    w = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in fx["sd"].items()}
    out = O.octic_vit_forward(fx["img"], w, patch=cfg["patch"], depth=cfg["depth"], num_heads=cfg["num_heads"],
                              invariant=cfg["invariant"])
    (out * fx["loss_weight"]).sum().backward()
    for k, g in fx["gparams"].items():
        close(w[k].grad, g, rtol=2e-4, atol=2e-3)


def test_timm_default_blocks(golden):
    fx = golden("model_timm_default")
    cfg = fx["cfg"]
    logits = O.octic_vit_forward(fx["img"], fx["sd"], patch=cfg["patch"], depth=cfg["depth"],
                                 num_heads=cfg["num_heads"], style="dinov2")
    close(logits, fx["logits"], rtol=1e-4, atol=5e-4)


def test_oracle_equivariance_of_the_trunk(golden):
    """experiments/test_equivariance.py logic on the oracle: g . f(x) == f(g . x) for the octic trunk."""
    fx = golden("model_hybrid")
    cfg = fx["cfg"]
    f = lambda im: O.octic_vit_forward(im.double(), {k: v.double() for k, v in fx["sd"].items()}, patch=cfg["patch"],
                                       depth=cfg["depth"], num_heads=cfg["num_heads"], return_trunk=True)
    base = f(fx["img"])
    for g in O.GROUP:
        moved = f(O.image_action(g, fx["img"]))
        want = O.token_action(g, base, has_cls=True)
        for a, b in zip(moved, want):
            torch.testing.assert_close(a, b, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("tag", ["model_dinov2", "model_dinov2_inv"])
def test_dinov2_backbone(golden, tag):
    """OcticDinoVisionTransformer (dinov2_models.py:40-260): token preparation with iBOT masks + registers, the
    feature dict, forward(), get_intermediate_layers and parameter gradients (tools/make_golden_dinov2.py)."""
    fx = golden(tag)
    cfg = fx["cfg"]
    kw = dict(patch=cfg["patch_size"], depth=cfg["depth"], num_heads=cfg["num_heads"], invariant=cfg["invariant"])
    tok = O.dino_prepare_tokens(fx["img"], fx["sd"], cfg["patch_size"], fx["masks"])
    for a, b in zip(tok, fx["tokens0"]):
        close(a, b, rtol=2e-5, atol=5e-5)
    feat = O.octic_dino_forward_features(fx["img"], fx["sd"], masks=fx["masks"], **kw)
    for k, v in fx["feat"].items():
        close(feat[k], v, rtol=1e-4, atol=5e-4)
    assert feat["x_norm_regtokens"].shape[1] == cfg["num_register_tokens"]
    feat2 = O.octic_dino_forward_features(fx["img2"], fx["sd"], **kw)
    for k, v in fx["feat2"].items():
        close(feat2[k], v, rtol=1e-4, atol=5e-4)
    plain, taken = O.octic_dino_forward_features(fx["img"], fx["sd"], take=(cfg["depth"] - 1,), **kw)
    close(plain["x_norm_clstoken"], fx["plain"], rtol=1e-4, atol=5e-4)
    import torch.nn.functional as F
    last = F.layer_norm(taken[0], (taken[0].shape[-1],), fx["sd"]["norm.weight"], fx["sd"]["norm.bias"], 1e-6)
    R = cfg["num_register_tokens"]
    close(last[:, 1 + R:], fx["inter_patch"], rtol=1e-4, atol=5e-4)
    close(last[:, 0], fx["inter_cls"], rtol=1e-4, atol=5e-4)
    w = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in fx["sd"].items()}
    out = O.octic_dino_forward_features(fx["img"], w, masks=fx["masks"], **kw)
    ((out["x_norm_clstoken"] * fx["w_cls"]).sum() + (out["x_norm_patchtokens"] * fx["w_patch"]).sum()).backward()
    assert len(fx["gparams"]) > 40
    for k, g in fx["gparams"].items():
        close(w[k].grad, g, rtol=2e-4, atol=2e-3)


@pytest.mark.parametrize("tag", ["image", "tokens"])
def test_isotypic_to_patch(golden, tag):
    """IsotypicToPatchD8 (d8_layers.py:499-588), forward and gradients (tools/make_golden_extra.py)."""
    fx = golden("isotypic_to_patch")[tag]
    w = {k: v.clone().requires_grad_(True) for k, v in fx["sd"].items()}
    xs = tuple(x.clone().requires_grad_(True) for x in fx["in"])
    kw = fx["kw"]
    y = O.isotypic_to_patch(xs, w, "", kw["patch_side"], kw["out_channels"], kw["reshape_to_image"])
    close(y, fx["out"])
    (y * fx["gout"]).sum().backward()
    for a, b in zip(xs, fx["gin"]):
        close(a.grad, b, rtol=1e-4, atol=1e-4)
    for k, g in fx["gparams"].items():
        close(w[k].grad, g, rtol=1e-4, atol=1e-4)


def test_isotypic_to_patch_is_equivariant(golden):
    """experiments/test_equivariance.py:257-274 on the oracle: acting on the octic tokens and then decoding to an image
    equals decoding and then acting on the image."""
    fx = golden("isotypic_to_patch")["image"]
    w = {k: v.double() for k, v in fx["sd"].items()}
    xs = tuple(x.double() for x in fx["in"])
    f = lambda t: O.isotypic_to_patch(t, w, "", fx["kw"]["patch_side"], fx["kw"]["out_channels"], True)
    base = f(xs)
    for g in O.GROUP:
        moved = f(O.token_action(g, xs, has_cls=False))
        torch.testing.assert_close(moved, O.image_action(g, base), rtol=1e-10, atol=1e-10)
