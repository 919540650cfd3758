"""tests/golden/isotypic_to_patch.pt: the reference IsotypicToPatchD8 (octic_vits/d8_layers.py:499-588, the
octic-feature -> image-patch head exercised by experiments/test_equivariance.py:257-274) run on CPU.
Build container only; same stand-ins as tools/make_golden.py."""
import sys
import warnings
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
warnings.filterwarnings("ignore")
from make_golden import ROOT, rand5, randomize  # noqa: E402

from octic_vits import d8_layers  # noqa: E402


def main():
    gen = torch.Generator().manual_seed(777)
    out = {}
    # both cases: LinearD8(64, 128), i.e. 16 output channels per one-dimensional irrep
    for tag, kw in (("image", dict(patch_side=4, out_channels=4, reshape_to_image=True)),
                    ("tokens", dict(patch_side=8, out_channels=1, reshape_to_image=False, bias=False))):
        torch.manual_seed(1)
        mod = d8_layers.IsotypicToPatchD8(dim=64, **kw)
        randomize(mod, gen, std=0.3)
        xs = tuple(x.clone().requires_grad_(True) for x in rand5(2, 9, 8, gen))
        y = mod(xs)
        gy = torch.randn(y.shape, generator=gen)
        (y * gy).sum().backward()
        out[tag] = {"sd": mod.state_dict(), "in": [x.detach() for x in xs], "out": y.detach(), "gout": gy,
                    "gin": [x.grad for x in xs], "gparams": {k: p.grad for k, p in mod.named_parameters()}, "kw": kw}
    path = ROOT / "tests" / "golden" / "isotypic_to_patch.pt"
    torch.save(out, path)
    print(f"isotypic_to_patch {path.stat().st_size / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
