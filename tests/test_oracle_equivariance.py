"""The reference's own test suite (experiments/test_equivariance.py) restated on the CPU oracle in fp64: every layer of
the hot path commutes with the D8 action, the invariantisation and the invariant model are invariant, the Fourier
transforms invert each other and the isotypic action is a group action.  (The same properties are checked for the CUDA
path in tests/test_gpu_model.py; here they guard the oracle itself, at 1e-9 instead of the reference's 1e-5..1e-6.)"""
import pytest
import torch

from oracle import octic_oracle as O

TOL = dict(rtol=1e-9, atol=1e-9)


def rand5(B, N, C, seed):
    g = torch.Generator().manual_seed(seed)
    off = torch.randn(B, N, 1, generator=g, dtype=torch.float64)       # per-token offset, as test_equivariance.py:147-150
    return tuple(torch.randn(B, N, C, generator=g, dtype=torch.float64) + off for _ in range(4)) + \
        (torch.randn(B, N, 2, 2 * C, generator=g, dtype=torch.float64) + off.unsqueeze(-1),)


def act(g, xs):
    return O.eight_to_five(O.isotypic_action(g, O.five_to_eight(xs)))


def d8_weights(prefix, din, dout, seed, bias=True):
    g = torch.Generator().manual_seed(seed)
    w = {f"{prefix}lin_{n}.weight": torch.randn(dout // 8, din // 8, generator=g, dtype=torch.float64) * 0.2 for n in O.IRREPS}
    w[f"{prefix}lin_E.weight"] = torch.randn(dout // 4, din // 4, generator=g, dtype=torch.float64) * 0.2
    if bias:
        w[f"{prefix}lin_A1.bias"] = torch.randn(dout // 8, generator=g, dtype=torch.float64)
    return w


def affine(prefix, dim, seed, bias):
    g = torch.Generator().manual_seed(seed)
    w = {f"{prefix}alpha_{n}": 1 + 0.3 * torch.randn(dim // 8, generator=g, dtype=torch.float64) for n in O.IRREPS}
    w[f"{prefix}alpha_E"] = 1 + 0.3 * torch.randn(dim // 4, generator=g, dtype=torch.float64)
    if bias:
        w[f"{prefix}beta"] = torch.randn(dim // 8, generator=g, dtype=torch.float64)
    return w


def block_weights(dim, seed, style):
    w = {}
    w.update(affine("norm1.scaling.", dim, seed + 1, True))
    w.update(affine("norm2.scaling.", dim, seed + 2, True))
    w.update(d8_weights("attn.qkv.", dim, 3 * dim, seed + 3))
    w.update(d8_weights("attn.proj.", dim, dim, seed + 4))
    w.update(d8_weights("mlp.fc1.", dim, 4 * dim, seed + 5))
    w.update(d8_weights("mlp.fc2.", 4 * dim, dim, seed + 6))
    ls = ("gamma_1.", "gamma_2.") if style == "deit" else ("ls1.", "ls2.")
    w.update(affine(ls[0], dim, seed + 7, False))
    w.update(affine(ls[1], dim, seed + 8, False))
    return w


LAYERS = {
    "GeluD8": lambda xs: O.gelu_d8(xs),
    "LinearD8": lambda xs: O.linear_d8(xs, d8_weights("", 64, 128, 1), ""),
    "LayerNormD8": lambda xs: O.layernorm_d8(xs, affine("scaling.", 64, 2, True), ""),
    "MlpD8": lambda xs: O.mlp_d8(xs, {**d8_weights("fc1.", 64, 256, 3), **d8_weights("fc2.", 256, 64, 4)}, ""),
    "AttentionD8": lambda xs: O.attention_d8(xs, {**d8_weights("qkv.", 64, 192, 5), **d8_weights("proj.", 64, 64, 6)}, "", 2),
    "Layer_scale_init_BlockD8": lambda xs: O.block_d8(xs, block_weights(64, 10, "deit"), "", 2, "deit"),
    "BlockD8": lambda xs: O.block_d8(xs, block_weights(64, 20, "dinov2"), "", 2, "dinov2"),
}


@pytest.mark.parametrize("name", list(LAYERS))
def test_layer_is_equivariant(name):
    """test_equi_isotypic_to_isotypic (test_equivariance.py:145-161): layer(g . x) == g . layer(x) for all 8 elements."""
    f = LAYERS[name]
    xs = rand5(3, 7, 8, seed=sum(map(ord, name)))
    base = f(xs)
    assert all(float(t.abs().max()) > 1e-3 for t in base)
    for g in O.GROUP:
        for a, b in zip(f(act(g, xs)), act(g, base)):
            torch.testing.assert_close(a, b, **TOL)
    if name != "LayerNormD8":
        # "bad test" guard of the reference (:138): the output is not itself invariant
        assert not torch.allclose(act("r", base)[2], base[2], atol=1e-6)


def test_isotypic_action_is_a_group_action():
    """test_group_action (:51-66): acting with g then h equals acting with the product, checked through the relations
    r^4 = m^2 = e and m r = r^3 m on random features."""
    xs8 = O.five_to_eight(rand5(2, 3, 4, seed=5))
    r = lambda t: O.isotypic_action("r", t)
    m = lambda t: O.isotypic_action("m", t)
    for a, b in zip(r(r(r(r(xs8)))), xs8):
        torch.testing.assert_close(a, b, **TOL)
    for a, b in zip(m(m(xs8)), xs8):
        torch.testing.assert_close(a, b, **TOL)
    for a, b in zip(m(r(xs8)), O.isotypic_action("mr", xs8)):
        torch.testing.assert_close(a, b, **TOL)
    for a, b in zip(r(r(r(m(xs8)))), m(r(xs8))):
        torch.testing.assert_close(a, b, **TOL)
    for g in ("rr", "rrr", "mrr", "mrrr"):
        step = xs8
        for _ in range(g.count("r")):
            step = r(step)
        if g.startswith("m"):
            step = m(step)
        for a, b in zip(step, O.isotypic_action(g, xs8)):
            torch.testing.assert_close(a, b, **TOL)


def test_fourier_transforms_are_inverse_and_intertwine_the_regular_representation():
    """test_fourier_transforms_inverses / test_fourier_transforms (:87-120): R2I(I2R(x)) == x, and in the regular
    (signal-on-the-group) basis the action of g is a permutation of the 8 components."""
    xs8 = O.five_to_eight(rand5(2, 3, 4, seed=6))
    for a, b in zip(O.regular_to_isotypic(O.isotypic_to_regular(xs8)), xs8):
        torch.testing.assert_close(a, b, **TOL)
    for a, b in zip(O.isotypic_to_regular(O.regular_to_isotypic(xs8)), xs8):
        torch.testing.assert_close(a, b, **TOL)
    reg = torch.stack(O.isotypic_to_regular(xs8), dim=-1)
    for g in O.GROUP:
        moved = torch.stack(O.isotypic_to_regular(O.isotypic_action(g, xs8)), dim=-1)
        # every component of the moved signal equals exactly one component of the original: a permutation
        d = (moved.unsqueeze(-1) - reg.unsqueeze(-2)).abs().flatten(0, -3).max(dim=0).values        # [8, 8]
        perm = d < 1e-9
        assert bool((perm.sum(0) == 1).all()) and bool((perm.sum(1) == 1).all()), g


def test_patch_embed_is_equivariant(golden):
    """test_equi_img_to_flattened_isotypic (:197-210) for PatchEmbedD8 with the golden model's lifting filters."""
    fx = golden("model_hybrid")
    w = {k: v.double() for k, v in fx["sd"].items() if k.startswith("patch_embed.")}
    img = fx["img"].double()
    base = O.patch_embed_d8(img, w, "patch_embed.", fx["cfg"]["patch"])
    for g in O.GROUP:
        moved = O.patch_embed_d8(O.image_action(g, img), w, "patch_embed.", fx["cfg"]["patch"])
        for a, b in zip(moved, O.token_action(g, base, has_cls=False)):
            torch.testing.assert_close(a, b, **TOL)


def test_power_spectrum_is_invariant():
    """test_inv_power_spectrum_invariant (:347-354)."""
    xs = rand5(2, 5, 8, seed=7)
    base = O.power_spectrum(xs)
    assert base.shape[-1] == 6 * 8
    for g in O.GROUP:
        torch.testing.assert_close(O.power_spectrum(act(g, xs)), base, **TOL)


def test_invariant_model_logits_are_invariant(golden):
    """test_invariance_img_to_logits (:302-315) on the tiny invariant golden model; the hybrid model is NOT invariant."""
    fx = golden("model_invariant")
    cfg = fx["cfg"]
    w = {k: v.double() for k, v in fx["sd"].items()}
    f = lambda im: O.octic_vit_forward(im, w, patch=cfg["patch"], depth=cfg["depth"], num_heads=cfg["num_heads"], invariant=True)
    img = fx["img"].double()
    base = f(img)
    for g in O.GROUP:
        torch.testing.assert_close(f(O.image_action(g, img)), base, rtol=1e-8, atol=1e-8)
    hy = golden("model_hybrid")
    wh = {k: v.double() for k, v in hy["sd"].items()}
    fh = lambda im: O.octic_vit_forward(im, wh, patch=cfg["patch"], depth=cfg["depth"], num_heads=cfg["num_heads"])
    assert not torch.allclose(fh(O.image_action("r", img)), fh(img), atol=1e-4)
