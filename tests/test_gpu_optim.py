"""GPU parity of the fused parameter update (octic_optim_* of include/octic_b200.h through optim.FusedOptimizer) against
oracle/optim_oracle.py.  fp32 arithmetic on both sides: tolerance 1e-5 relative on parameters after several steps (the
per-tensor norms are summed in a different order than the oracle's; everything else is element-wise)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import optim_oracle as OO

if torch.cuda.is_available():
    from octic_vits_b200 import functional as OF
    from octic_vits_b200.optim import FusedOptimizer
    from octic_vits_b200.parallel import FlatGrads

DEV = "cuda"


class Bag(torch.nn.Module):
    """parameters of awkward sizes: a tail that is not a multiple of 4, a tensor longer than one chunk, a scalar-ish one"""

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(5)
        self.weight = torch.nn.Parameter(torch.randn(129, 67, generator=g))          # 8643 elements: 2 chunks + tail
        self.bias = torch.nn.Parameter(torch.randn(1001, generator=g))
        self.big = torch.nn.Parameter(torch.randn(3, 8192 + 5, generator=g))
        self.tiny = torch.nn.Parameter(torch.randn(3, generator=g))
        self.frozen = torch.nn.Parameter(torch.randn(10, generator=g), requires_grad=False)

    def no_weight_decay(self):
        return {"big"}


def run(kind, steps=4, ema=False, **kw):
    torch.manual_seed(0)
    model = Bag().to(DEV)
    ema_model = Bag().to(DEV) if ema else None
    names = [n for n, p in model.named_parameters() if p.requires_grad]
    ref_p = [p.detach().cpu().clone() for p in model.parameters() if p.requires_grad]
    ref_e = [p.detach().cpu().clone() for n, p in ema_model.named_parameters() if n in names] if ema else None
    m = [torch.zeros_like(p) for p in ref_p]
    v = [torch.zeros_like(p) for p in ref_p]
    fg = FlatGrads(model.parameters())
    opt = FusedOptimizer(model, fg, kind=kind, ema=(ema_model, 0.99) if ema else None, **kw)
    wds = [w for w, _ in opt.seg_hparams]
    assert wds == [kw.get("weight_decay", 0.0), 0.0, 0.0, 0.0]         # weight, bias (1-D), big (no_weight_decay), tiny (1-D)
    gen = torch.Generator().manual_seed(9)
    epoch0 = OF._param_epoch
    for step in range(1, steps + 1):
        scale = 5.0 if step == 2 else 0.01                               # step 2 clips, the others do not
        grads = [scale * torch.randn(p.shape, generator=gen) for p in ref_p]
        for p, g in zip(fg.params, grads):
            p.grad.copy_(g)
        opt.step()
        if kind == "lamb":
            OO.lamb_step(ref_p, grads, m, v, step, opt.lr, opt.betas, opt.eps, wds, None, opt.max_grad_norm)
        else:
            if opt.max_grad_norm > 0:
                gn = torch.sqrt(sum(g.double().pow(2).sum() for g in grads)).item()
                grads = [g / max(gn / opt.max_grad_norm, 1.0) for g in grads]
            OO.adamw_step(ref_p, grads, m, v, step, opt.lr, opt.betas, opt.eps, wds)
        if ema:
            OO.ema_update(ref_e, ref_p, 0.99)
    assert OF._param_epoch == epoch0 + steps
    for n, p, r in zip(names, fg.params, ref_p):
        torch.testing.assert_close(p.detach().cpu(), r, rtol=2e-5, atol=2e-6, msg=lambda s: f"{kind} {n}: {s}")
    if ema:
        for n, r in zip(names, ref_e):
            torch.testing.assert_close(dict(ema_model.named_parameters())[n].detach().cpu(), r, rtol=2e-5, atol=2e-6)
        assert torch.equal(ema_model.frozen, Bag().frozen.to(DEV))
    assert torch.equal(model.frozen.cpu(), Bag().frozen)
    opt.final_params = [p.detach().clone() for p in fg.params]
    return opt


def test_fused_lamb_matches_oracle():
    opt = run("lamb", lr=3e-3, weight_decay=0.05)
    assert float(opt.grad_norm()) > 0


def test_fused_update_is_bit_reproducible():
    """no float atomics in the norm reductions: two runs from the same state give identical bits, which is what keeps
    data-parallel replicas (identical all-reduced gradients) identical"""
    a = run("lamb", lr=3e-3, weight_decay=0.05, steps=6)
    b = run("lamb", lr=3e-3, weight_decay=0.05, steps=6)
    for x, y in zip(a.final_params, b.final_params):
        assert torch.equal(x, y)
    assert torch.equal(a.exp_avg_sq, b.exp_avg_sq) and float(a.grad_norm()) == float(b.grad_norm())


def test_fused_lamb_with_ema_matches_oracle():
    run("lamb", lr=1e-2, weight_decay=0.02, ema=True)


def test_fused_adamw_matches_oracle_and_torch():
    run("adamw", lr=1e-3, weight_decay=0.04, betas=(0.9, 0.999))
    run("adamw", lr=1e-3, weight_decay=0.04, max_grad_norm=3.0, ema=True)
    # and against torch.optim.AdamW on the device directly
    torch.manual_seed(1)
    a, b = Bag().to(DEV), Bag().to(DEV)
    fg = FlatGrads(a.parameters())
    opt = FusedOptimizer(a, fg, kind="adamw", lr=2e-3, weight_decay=0.1, no_decay=())
    ref = torch.optim.AdamW([p for p in b.parameters() if p.requires_grad], lr=2e-3, weight_decay=0.1)
    for _ in range(3):
        for p, q in zip(fg.params, [p for p in b.parameters() if p.requires_grad]):
            g = torch.randn_like(p)
            p.grad.copy_(g)
            q.grad = g.clone()
        opt.step()
        ref.step()
    for p, q in zip(fg.params, [p for p in b.parameters() if p.requires_grad]):
        torch.testing.assert_close(p, q, rtol=2e-5, atol=2e-6)


def test_graphed_train_step_with_optimizer_learns_and_matches_eager():
    """fwd + bwd from the CUDA graph (weight packs captured inside it), then the fused LAMB step: the loss on a fixed
    batch falls, and the parameters follow the eager path (same kernels, atomics aside)."""
    from octic_vits_b200.model import OcticVisionTransformer
    from octic_vits_b200.parallel import GraphedTrainStep

    def make():
        torch.manual_seed(3)
        return OcticVisionTransformer(img_size=64, patch_size=16, embed_dim=128, depth=4, num_heads=2, num_classes=10,
                                      qkv_bias=True, init_scale=0.1).to(DEV).train()
    img = torch.randn(8, 3, 64, 64, device=DEV)
    tgt = torch.randint(0, 10, (8,), device=DEV)
    results = []
    for use_graph in (True, False):
        model = make()
        fg = FlatGrads(model.parameters())
        opt = FusedOptimizer(model, fg, kind="lamb", lr=2e-2, weight_decay=0.05)
        step = GraphedTrainStep(model, fg, img.shape, use_graph=use_graph, optimizer=opt)
        assert step.graphed == use_graph, getattr(step, "capture_error", None)
        losses = [float(step(img, tgt)) for _ in range(8)]
        assert opt.step_count == 8
        results.append((losses, {n: p.detach().clone() for n, p in model.named_parameters()}))
        # eval after training uses freshly packed weights (pack epoch), not the graph's
        model.eval()
        with torch.no_grad():
            assert torch.isfinite(model(img)).all()
    (lg, pg), (le, pe) = results
    assert lg[-1] < lg[0], lg
    for a, b in zip(lg, le):
        assert a == pytest.approx(b, rel=3e-2, abs=3e-2), (lg, le)
    num = sum(float((pg[n] - pe[n]).pow(2).sum()) for n in pg)
    den = sum(float(pe[n].pow(2).sum()) for n in pe)
    assert (num / den) ** 0.5 < 2e-2


def test_schedulable_hyperparameters_and_state_round_trip():
    """ADVICE r01 (optim.py:90): weight decay and per-parameter lr multipliers can be changed after construction (cosine
    weight-decay schedule, last-layer freezing of the DINOv2 recipe), every hyper-parameter survives
    state_dict() -> load_state_dict(), and a re-allocated parameter anywhere in the list is detected."""
    torch.manual_seed(0)
    model = Bag().to(DEV)
    fg = FlatGrads(model.parameters())
    opt = FusedOptimizer(model, fg, kind="adamw", lr=1e-2, weight_decay=0.1, ema=None)
    ref_p = [p.detach().cpu().clone() for p in fg.params]
    m = [torch.zeros_like(p) for p in ref_p]
    v = [torch.zeros_like(p) for p in ref_p]
    gen = torch.Generator().manual_seed(3)
    sched = [(0.1, 1.0), (0.3, 1.0), (0.3, 0.0)]          # (weight decay, lr multiplier of `bias`) per step
    for step, (wd, bias_scale) in enumerate(sched, start=1):
        opt.set_weight_decay(wd)
        opt.set_lr_scales({"bias": bias_scale})
        assert [w for w, _ in opt.seg_hparams] == [wd, 0.0, 0.0, 0.0]
        grads = [0.01 * torch.randn(p.shape, generator=gen) for p in ref_p]
        for p, g in zip(fg.params, grads):
            p.grad.copy_(g)
        opt.step()
        # lr multiplier 0 freezes the tensor (its moments still advance)
        OO.adamw_step(ref_p, grads, m, v, step, opt.lr, opt.betas, opt.eps, [wd, 0.0, 0.0, 0.0], [1.0, bias_scale, 1.0, 1.0])
    for p, r in zip(fg.params, ref_p):
        torch.testing.assert_close(p.detach().cpu(), r, rtol=2e-5, atol=2e-6)
    with pytest.raises(ValueError):
        opt.set_lr_scales({"no_such_parameter": 1.0})
    sd = opt.state_dict()
    opt2 = FusedOptimizer(model, fg, kind="adamw", lr=5.0, betas=(0.5, 0.5), eps=1.0, weight_decay=0.0)
    opt2.set_ema_momentum(0.5)
    opt2.load_state_dict(sd)
    assert (opt2.lr, opt2.betas, opt2.eps, opt2.ema_momentum, opt2.step_count) == (opt.lr, opt.betas, opt.eps, opt.ema_momentum, 3)
    assert opt2.seg_hparams == opt.seg_hparams and torch.equal(opt2.segs, opt.segs)
    assert torch.equal(opt2.exp_avg, opt.exp_avg)
    # a parameter re-allocated after construction (here the LAST one) is caught before any kernel writes through a stale pointer
    model.tiny.data = model.tiny.data.clone()
    with pytest.raises(Exception, match="re-allocated"):
        opt.step()
