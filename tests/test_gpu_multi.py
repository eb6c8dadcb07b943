"""GPU, 2 ranks (skipped on a single-GPU box): the data-parallel training step of BASELINE config 4 -- replicas with
their own pair slices, ONE flat NCCL all-reduce of the gradient per step, identical Adam updates."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, ROOT)
import se3_equi_graph_registration_b200 as P
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
model = P.build_model(os.path.join(ROOT, "tests", "golden", "checkpoint-3dmatch.pth"), device=dev, variant="train")
with torch.no_grad():
    model.egnn.embedding_out.weight.mul_(0.005); model.egnn.embedding_out.bias.mul_(0.005)
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
def batch(seed):
    d = {k: v.to(dev) for k, v in P.synthetic.make_batch(seed, 2, n=512).items()}
    es, et = P.knn_graph_batch(d["src_pts"], 16), P.knn_graph_batch(d["tgt_pts"], 16)
    ones = torch.ones(2, 512 * 16, 1, device=dev)
    return (d["src_feat"], d["src_pts"], es, ones, d["tgt_feat"], d["tgt_pts"], et, ones, d["corr"], d["labels"], d["gt_pose"])
print('phase 1', rank, flush=True)
# (1) the all-reduced gradient is the mean of the ranks' local gradients
b = batch(100 + rank)
model.train(); opt.zero_grad(set_to_none=True)
P.train.training_loss(model(*b), b[10]).backward()
params = [p for p in model.parameters()]
live = [p for p in params if p.grad is not None]
assert len(live) == 85
local = torch.cat([p.grad.flatten() for p in live])
gathered = [torch.empty_like(local) for _ in range(world)]
dist.all_gather(gathered, local)
P.train.allreduce_gradients(params)
after = torch.cat([p.grad.flatten() for p in live])
want = torch.stack(gathered).mean(0)
assert torch.allclose(after, want, rtol=1e-5, atol=1e-7 * float(want.abs().max())), float((after - want).abs().max())
assert float((gathered[0] - gathered[1]).abs().max()) > 0          # the ranks really had different data
assert sum(p.grad is not None for p in params) == 85                # dead parameters stay None on every rank
print('phase 2', rank, flush=True)
# (2) replicas stay identical over several eager steps on different data
for s in range(3):
    loss = P.train.train_step(model, opt, batch(200 + 10 * s + rank))
    assert torch.isfinite(loss)
def checksum():
    chk = torch.cat([p.detach().flatten() for p in params]).double().sum().reshape(1)
    both = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(both, chk)
    return both
both = checksum()
assert both[0].item() == both[1].item(), (both[0].item(), both[1].item())
print('phase 3', rank, flush=True)
# (3) the graphed step (one all-reduce of the persistent flat gradient inside the CUDA graph): its 2-rank gradient is the
# mean of the two ranks' 1-rank gradients, and the replicas stay bit-identical
keys = ("src_feat", "src_pts", "tgt_feat", "tgt_pts", "corr", "labels", "gt_pose")
def raw_batch(seed):
    d = P.synthetic.make_batch(seed, 2, n=512)
    return tuple(d[k].to(dev) for k in keys)
opt2 = torch.optim.Adam(model.parameters(), lr=1e-4, capturable=True)
step = P.train.GraphedTrainStep(model, opt2, raw_batch(300 + rank), k=16)
for s in range(3):
    loss = step(raw_batch(310 + 10 * s + rank))
    assert torch.isfinite(loss)
print('phase 3 stepped', rank, flush=True)
gl = [torch.empty_like(step.state.flat_grad) for _ in range(world)]
dist.all_gather(gl, step.state.flat_grad)
assert torch.equal(gl[0], gl[1])                                    # every rank holds the same (all-reduced) gradient
both = checksum()
assert both[0].item() == both[1].item(), (both[0].item(), both[1].item())
if rank == 0:
    print("DDP_OK", float(loss), flush=True)
del step                       # the CUDA graph holds a captured NCCL collective: release it before tearing NCCL down
torch.cuda.synchronize()
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_training_step_two_ranks_nccl(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(f"ROOT = {ROOT!r}\n" + WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and "DDP_OK" in out.stdout, (out.stdout[-1500:], out.stderr[-3000:])
