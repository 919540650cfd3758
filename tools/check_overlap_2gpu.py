"""torchrun --nproc-per-node 2 tools/check_overlap_2gpu.py
Two-rank checks of the data-parallel step on real GPUs (NCCL):
 1. gradients of the graphed step WITH the early all-reduce captured inside the CUDA graph == gradients of the eager
    step with one all-reduce after backward (same shards);
 2. after several fused-LAMB steps on different shards the parameters of the two replicas are bit-identical
    (deterministic optimizer reductions + identical all-reduced gradients)."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from octic_vits_b200.model import OcticVisionTransformer  # noqa: E402
from octic_vits_b200.optim import FusedOptimizer  # noqa: E402
from octic_vits_b200.parallel import FlatGrads, GraphedTrainStep, install_early_allreduce  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)


def make():
    torch.manual_seed(3)                       # same replica on both ranks
    return OcticVisionTransformer(img_size=64, patch_size=16, embed_dim=128, depth=4, num_heads=2, num_classes=10,
                                  qkv_bias=True, init_scale=0.1).to(dev).train()


g = torch.Generator().manual_seed(100 + rank)  # different shard per rank
img = torch.randn(8, 3, 64, 64, generator=g).to(dev)
tgt = torch.randint(0, 10, (8,), generator=g).to(dev)

# 1. overlap inside the graph vs plain
ref_model = make()
ref_fg = FlatGrads(ref_model.parameters())
ref_step = GraphedTrainStep(ref_model, ref_fg, img.shape, use_graph=False)
ref_step(img, tgt)
want = ref_fg.flat.clone()
model = make()
fg = FlatGrads(model.parameters())
assert install_early_allreduce(model, fg)
step = GraphedTrainStep(model, fg, img.shape)
for _ in range(3):
    loss = step(img, tgt)
torch.cuda.synchronize()
err = float((fg.flat - want).norm() / want.norm())
print(f"[rank {rank}] graphed={step.graphed} early_in_graph={step.early_in_graph} split={fg.split}/{fg.flat.numel()} "
      f"capture_error={getattr(step, 'capture_error', None)} rel err vs single all-reduce {err:.2e}", flush=True)
assert err < 2e-3, err
other = fg.flat.clone()
dist.broadcast(other, src=0)
assert torch.equal(other, fg.flat), "all-reduced gradients differ between ranks"

# 2. replicas stay bit-identical through fused LAMB steps
opt_model = make()
ofg = FlatGrads(opt_model.parameters())
install_early_allreduce(opt_model, ofg)
opt = FusedOptimizer(opt_model, ofg, kind="lamb", lr=1e-2, weight_decay=0.05)
ostep = GraphedTrainStep(opt_model, ofg, img.shape, optimizer=opt)
losses = [float(ostep(img, tgt).detach()) for _ in range(6)]
flat_p = torch.cat([p.detach().flatten() for p in opt_model.parameters()])
other = flat_p.clone()
dist.broadcast(other, src=0)
same = torch.equal(other, flat_p)
print(f"[rank {rank}] LAMB x6: losses {losses[0]:.4f} -> {losses[-1]:.4f}, replicas bit-identical: {same}", flush=True)
assert same
step.close()          # graphs that captured NCCL kernels must go before the communicator does
ostep.close()
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("overlap + replica consistency ok")
