"""Fused parameter update for the training step: LAMB (the DeiT-III recipe,
experiments/train_deit.py:42 `fusedlamb` -> apex FusedLAMB through timm.optim.create_optimizer, deit/main.py:365),
AdamW (the DINOv2 recipe, dinov2/train/train.py:67-68) and the EMA / teacher update (deit/engine.py:81-82,
dinov2/train/ssl_meta_arch.py:370-379) -- two or three kernel launches per step for the WHOLE model instead of one
multi-tensor launch group per optimizer phase.

    fg  = FlatGrads(model.parameters())
    opt = FusedOptimizer(model, fg, kind="lamb", lr=3e-3, weight_decay=0.05)        # or kind="adamw"
    loss.backward(); fg.all_reduce(); opt.step()

Gradients are read from the flat buffer of `parallel.FlatGrads` (LAMB overwrites them with the update direction; the
buffer is zeroed at the start of the next step anyway); moments are two more flat fp32 buffers with the same offsets.
Parameters stay where they are and are reached through a chunk table (include/octic_b200.h `octic_optim_chunk`).
The kernels write parameters through raw pointers, so `step()` bumps functional's pack epoch: the bf16 GEMM operands
are re-packed on next use.  No CPU path: CPU parameters raise OcticError.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, Optional, Tuple

import torch

from . import functional as OF
from ._lib import OcticError, OptimChunk, OptimSeg, call
from .parallel import FlatGrads

CHUNK = 8192
SQNORM_PARTIALS = 1184      # OCTIC_OPTIM_SQNORM_PARTIALS


def timm_no_decay(model: torch.nn.Module) -> set:
    """Names without weight decay under timm's `param_groups_weight_decay` (what create_optimizer applies at
    deit/main.py:365): 1-D parameters, biases, and model.no_weight_decay() (model.py:229-234)."""
    skip = set(model.no_weight_decay()) if hasattr(model, "no_weight_decay") else set()
    return {n for n, p in model.named_parameters()
            if p.requires_grad and (p.ndim <= 1 or n.endswith(".bias") or n in skip)}


def build_chunk_table(numels, offsets, ptrs, ema_ptrs, hparams, chunk: int = CHUNK):
    """Host-side tables of the C ABI (include/octic_b200.h): one `octic_optim_seg` per parameter tensor and its
    `octic_optim_chunk`s (<= `chunk` elements each, contiguous per tensor).  `ptrs` / `ema_ptrs` are device addresses
    (0 = no EMA copy), `offsets` the tensors' first elements in the flat gradient buffer, `hparams` (weight_decay,
    lr_scale) per tensor.  Returns (chunks, segs) as lists of tuples in the structs' field order."""
    chunks, segs = [], []
    for seg, (n, off, ptr, eptr, (wd, scale)) in enumerate(zip(numels, offsets, ptrs, ema_ptrs, hparams)):
        segs.append((float(wd), float(scale), len(chunks), (n + chunk - 1) // chunk))
        for s in range(0, n, chunk):
            chunks.append((ptr + 4 * s, eptr + 4 * s if eptr else 0, off + s, min(chunk, n - s), seg))
    return chunks, segs


class FusedOptimizer:
    """kind = "lamb": apex FusedLAMB defaults (bias_correction, grad_averaging, adam_w_mode, max_grad_norm=1.0,
    use_nvlamb=False: the trust ratio only applies to tensors with weight decay);  kind = "adamw": torch.optim.AdamW.
    `lr_scales` maps parameter names to a learning-rate multiplier (layer-wise decay / patch-embed multiplier of the
    DINOv2 recipe); `set_weight_decay` / `set_lr_scales` change them between steps (cosine weight-decay schedule,
    last-layer freezing).  Default eps: 1e-6 for LAMB (apex FusedLAMB's default; the DeiT recipe passes --opt-eps, whose
    default is 1e-8: give eps= explicitly to match a run), 1e-8 for AdamW.  `ema` = (ema_model, momentum): ema_model's parameters with the same names are updated as
    ema = momentum*ema + (1-momentum)*param inside the kernel that writes the parameter."""

    def __init__(self, model: torch.nn.Module, flat_grads: FlatGrads, kind: str = "lamb", lr: float = 1e-3,
                 betas: Tuple[float, float] = (0.9, 0.999), eps: Optional[float] = None, weight_decay: float = 0.0,
                 max_grad_norm: Optional[float] = None, no_decay: Optional[Iterable[str]] = None,
                 lr_scales: Optional[Dict[str, float]] = None, grad_averaging: bool = True, use_nvlamb: bool = False,
                 ema: Optional[Tuple[torch.nn.Module, float]] = None):
        if kind not in ("lamb", "adamw"):
            raise ValueError(f"unknown optimizer kind {kind!r}")
        self.kind, self.lr, self.betas = kind, float(lr), (float(betas[0]), float(betas[1]))
        self.eps = float(eps) if eps is not None else (1e-6 if kind == "lamb" else 1e-8)
        self.max_grad_norm = float(max_grad_norm) if max_grad_norm is not None else (1.0 if kind == "lamb" else 0.0)
        self.grad_averaging, self.use_nvlamb = grad_averaging, use_nvlamb
        self.fg, self.step_count = flat_grads, 0
        names = {id(p): n for n, p in model.named_parameters()}
        no_decay = timm_no_decay(model) if no_decay is None else set(no_decay)
        lr_scales = lr_scales or {}
        dev = flat_grads.flat.device
        ema_params, self.ema_momentum = {}, 0.0
        if ema is not None:
            ema_params, self.ema_momentum = dict(ema[0].named_parameters()), float(ema[1])
        ema_ptrs, hparams = [], []
        for p in flat_grads.params:
            name = names.get(id(p))
            if name is None:
                raise ValueError("flat_grads holds a parameter that is not in model.named_parameters()")
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise OcticError(f"{name}: parameters must be contiguous fp32")
            e = ema_params.get(name)
            if e is not None and (e.shape != p.shape or e.dtype != torch.float32 or not e.is_contiguous() or e.device != p.device):
                raise OcticError(f"{name}: EMA copy must match the parameter (shape, fp32, contiguous, same device)")
            ema_ptrs.append(e.data_ptr() if e is not None else 0)
            hparams.append((0.0 if name in no_decay else float(weight_decay), float(lr_scales.get(name, 1.0))))
        chunks, segs = build_chunk_table([p.numel() for p in flat_grads.params], flat_grads.offsets,
                                         [p.data_ptr() for p in flat_grads.params], ema_ptrs, hparams)
        self.nchunks, self.nseg = len(chunks), len(segs)
        self.seg_hparams = [(wd, sc) for wd, sc, _, _ in segs]
        self._seg_layout = [(c0, nc) for _, _, c0, nc in segs]
        self._names = [names[id(p)] for p in flat_grads.params]
        self._decayed = [name not in no_decay for name in self._names]
        if dev.type != "cuda":
            raise OcticError("FusedOptimizer needs CUDA parameters: there is no CPU path")
        self._ptrs = [p.data_ptr() for p in flat_grads.params]
        ch = (OptimChunk * len(chunks))(*[OptimChunk(*c) for c in chunks])
        sg = (OptimSeg * len(segs))(*[OptimSeg(*s) for s in segs])
        self.chunks = torch.frombuffer(bytearray(bytes(ch)), dtype=torch.uint8).to(dev)
        self.segs = torch.frombuffer(bytearray(bytes(sg)), dtype=torch.uint8).to(dev)
        self.exp_avg = torch.zeros_like(flat_grads.flat)
        self.exp_avg_sq = torch.zeros_like(flat_grads.flat)
        # reduction workspace: [gnorm^2 | pad | per-CTA gnorm partials | (|p|^2, |u|^2) per tensor | ... per chunk]
        self._o_part, self._o_seg = 4, 4 + SQNORM_PARTIALS
        self._o_chunk = self._o_seg + 2 * self.nseg
        self.scratch = torch.zeros(self._o_chunk + 2 * self.nchunks, dtype=torch.float32, device=dev)

    def set_lr(self, lr: float) -> None:
        self.lr = float(lr)

    def _write_segs(self) -> None:
        """Re-upload the per-tensor table (weight decay, lr multiplier) -- a few KB, stream ordered."""
        segs = [(wd, sc, c0, nc) for (wd, sc), (c0, nc) in zip(self.seg_hparams, self._seg_layout)]
        sg = (OptimSeg * len(segs))(*[OptimSeg(*s_) for s_ in segs])
        self.segs.copy_(torch.frombuffer(bytearray(bytes(sg)), dtype=torch.uint8), non_blocking=False)

    def set_weight_decay(self, weight_decay: float) -> None:
        """New weight decay for every decayed tensor (the cosine weight-decay schedule of the DINOv2 recipe,
        dinov2/train/train.py:87-93); tensors in `no_decay` stay at 0."""
        self.seg_hparams = [(float(weight_decay) if d else 0.0, sc) for (_, sc), d in zip(self.seg_hparams, self._decayed)]
        self._write_segs()

    def set_lr_scales(self, lr_scales: Dict[str, float]) -> None:
        """New per-parameter learning-rate multipliers by name (layer-wise decay, last-layer freezing = 0.0); names that
        are not given keep their multiplier."""
        unknown = set(lr_scales) - set(self._names)
        if unknown:
            raise ValueError(f"unknown parameter names: {sorted(unknown)[:4]}")
        self.seg_hparams = [(wd, float(lr_scales.get(n, sc))) for (wd, sc), n in zip(self.seg_hparams, self._names)]
        self._write_segs()

    def set_ema_momentum(self, m: float) -> None:
        self.ema_momentum = float(m)

    def step(self) -> None:
        """Consumes the (already all-reduced) flat gradients.  5 launches for LAMB (grad-norm partials + final sum,
        stage 1, per-tensor norm sums, stage 2), 1 for AdamW (3 with clipping).  No atomics: bit-reproducible."""
        if any(p.data_ptr() != q for p, q in zip(self.fg.params, self._ptrs)):
            raise OcticError("parameters were re-allocated after the optimizer was built (call .to()/.cuda() first)")
        self.step_count += 1
        b1, b2 = self.betas
        bc1, bc2 = 1.0 - b1 ** self.step_count, 1.0 - b2 ** self.step_count
        beta3 = (1.0 - b1) if (self.grad_averaging or self.kind == "adamw") else 1.0
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        g, gnorm = self.fg.flat, None
        lamb = self.kind == "lamb"
        base = self.scratch.data_ptr()
        if self.max_grad_norm > 0.0:
            call("octic_optim_sqnorm", g.data_ptr(), g.numel(), base + 4 * self._o_part, base, st)
            gnorm = base
        chunk_norms = base + 4 * self._o_chunk
        call("octic_optim_stage1", self.chunks.data_ptr(), self.nchunks, self.segs.data_ptr(), g.data_ptr(),
             self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), chunk_norms if lamb else None, gnorm,
             self.max_grad_norm, b1, b2, beta3, self.eps, bc1, bc2, self.lr, 0 if lamb else 1, self.ema_momentum, st)
        if lamb:
            call("octic_optim_lamb_stage2", self.chunks.data_ptr(), self.nchunks, self.segs.data_ptr(), self.nseg,
                 g.data_ptr(), chunk_norms, base + 4 * self._o_seg, self.lr, 1 if self.use_nvlamb else 0,
                 self.ema_momentum, st)
        OF.bump_param_epoch()

    def grad_norm(self) -> torch.Tensor:
        """global gradient norm seen by the last step (device scalar; 0 when clipping is off)"""
        return self.scratch[0].sqrt()

    def state_dict(self) -> dict:
        return {"kind": self.kind, "step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "lr": self.lr, "betas": self.betas, "eps": self.eps, "max_grad_norm": self.max_grad_norm,
                "ema_momentum": self.ema_momentum, "grad_averaging": self.grad_averaging, "use_nvlamb": self.use_nvlamb,
                "seg_hparams": list(self.seg_hparams), "names": list(self._names)}

    def load_state_dict(self, sd: dict) -> None:
        """Restores moments, step count AND every hyper-parameter that state_dict() saved (older dicts without the
        hyper-parameter keys keep the current values)."""
        if sd["kind"] != self.kind or sd["exp_avg"].numel() != self.exp_avg.numel():
            raise ValueError("optimizer state does not match this optimizer")
        if "names" in sd and list(sd["names"]) != self._names:
            raise ValueError("optimizer state was saved for a different parameter list")
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.lr = float(sd.get("lr", self.lr))
        self.betas = tuple(float(b) for b in sd.get("betas", self.betas))
        self.eps = float(sd.get("eps", self.eps))
        self.max_grad_norm = float(sd.get("max_grad_norm", self.max_grad_norm))
        self.ema_momentum = float(sd.get("ema_momentum", self.ema_momentum))
        self.grad_averaging = bool(sd.get("grad_averaging", self.grad_averaging))
        self.use_nvlamb = bool(sd.get("use_nvlamb", self.use_nvlamb))
        if "seg_hparams" in sd:
            self.seg_hparams = [(float(wd), float(sc)) for wd, sc in sd["seg_hparams"]]
            self._write_segs()
