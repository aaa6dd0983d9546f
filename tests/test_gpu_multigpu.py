"""Data-parallel gradient exchange on real devices.

* the hook-driven early reduction of the word-table gradient through the REAL encoder backward (ops.EncodeDocsFn) with two
  ranks — gloo backend, both ranks on cuda:0, so it runs on a one-GPU box;
* the library's own NVLS kernel (csrc/multimem.cu: multimem.ld_reduce / multimem.st through the NVSwitch) against
  ncclAllReduce on >= 2 GPUs — eager, overlapped, and captured in a CUDA graph (skipped on a one-GPU box; bench.py repeats
  the comparison before timing whenever it runs at N > 1 and prints it as `allreduce_check`).
"""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_COMMON = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["RBR_ROOT"])
import rbr_b200
from rbr_b200 import parallel, synth
U, I, V, E, H, K, L, B = 40, 30, 800, 64, 24, 16, 96, 64
def build(precision="fp32"):
    params = synth.deepconn_params(U, I, V, E, H, K, (3,), seed=3)
    m = rbr_b200.DeepCoNNpp(U, I, V, [3], E, H, K, L, None, 0.0, precision=precision)
    m.load_state_dict(params)
    return m.cuda().train()
def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12))
def step(m, batch, ratings):
    m.zero_grad(set_to_none=True)
    loss = torch.nn.MSELoss()(m(*batch), ratings)
    loss.backward()
    return loss
'''

_GLOO_WORKER = _COMMON + r'''
rank, local, world = parallel.init_from_env("gloo")
torch.cuda.set_device(0)                                   # both ranks share the one GPU (gloo moves CUDA tensors through the host)
model = build()
parallel.broadcast_parameters(model)
batch, ratings = synth.deepconn_batch(B, L, V, U, I, seed=21)
batch, ratings = [t.cuda() for t in batch], ratings.cuda()
# single-process reference: the whole batch on this rank
step(model, batch, ratings)
ref = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
sh = parallel.shard_range(B, rank, world)
mine, mine_r = [t[sh.start:sh.stop] for t in batch], ratings[sh.start:sh.stop]
for mode in ("plain", "overlap"):
    if mode == "overlap":
        parallel.enable_overlap(model)
    step(model, mine, mine_r)
    arena = model.last_arena
    lo = arena.flat.data_ptr(); hi = lo + arena.flat.numel() * 4
    for k, p in model.named_parameters():
        assert lo <= p.grad.data_ptr() < hi, (mode, k, "gradient left the arena (cloned by AccumulateGrad?)")
    n = parallel.allreduce_gradients(model)
    assert n == 1, (mode, n)
    torch.cuda.synchronize()
    for k, p in model.named_parameters():
        assert rel(p.grad, ref[k]) < 3e-5, (mode, k, rel(p.grad, ref[k]))
dist.barrier()
print("rank", rank, "ok")
'''

_NVLS_WORKER = _COMMON + r'''
rank, local, world = parallel.init_from_env("nccl")
torch.cuda.set_device(local)
from rbr_b200.graphs import GraphedTrainStep
results = {}
for overlap, kind in ((False, "multimem"), (True, "multimem"), (False, "p2p"), (True, "p2p")):
    model = build("bf16")
    parallel.broadcast_parameters(model)
    ok = parallel.enable_nvls_allreduce(model, overlap=overlap, kind=kind)
    assert not ok or model.__dict__["_rbr_nvls"].kind == kind
    if not ok:
        print("rank", rank, "SKIP no NVLS multicast")
        dist.barrier(); dist.destroy_process_group(); sys.exit(0)
    batch, ratings = synth.deepconn_batch(B, L, V, U, I, seed=100 + rank)
    batch, ratings = [t.cuda() for t in batch], ratings.cuda()
    # expected: NCCL average of every rank's LOCAL gradients — taken with the early-reduction hook detached: with it, backward
    # already starts averaging the word-table slice on the side stream, and only allreduce_gradients() waits for that
    hook = model.ngram.table_grad_hook
    model.ngram.table_grad_hook = None
    step(model, batch, ratings)
    model.ngram.table_grad_hook = hook
    flat = model.last_arena.flat
    expect = flat.detach().clone()
    dist.all_reduce(expect, op=dist.ReduceOp.AVG)
    for trial in range(3):                                  # the persistent symmetric-memory arena is re-zeroed each step
        step(model, batch, ratings)
        assert model.last_arena.flat.data_ptr() == model.__dict__["_rbr_arena_buffer"].data_ptr()
        n = parallel.allreduce_gradients(model)
        assert n == 1
        torch.cuda.synchronize()
        err = rel(model.last_arena.flat, expect)
        assert err < 2e-6, (kind, overlap, trial, err)
    # the same exchange captured in a CUDA graph with the step
    gs = GraphedTrainStep(model, torch.nn.MSELoss(), batch, ratings, post_backward=lambda: parallel.allreduce_gradients(model))
    for trial in range(3):
        gs.replay()
        torch.cuda.synchronize()
        err = rel(model.last_arena.flat, expect)
        assert err < 2e-6, ("graph", kind, overlap, trial, err)
    # gradient accumulation into the persistent arena is refused, not silently wiped
    model.zero_grad(set_to_none=False)
    try:
        step_ok = True
        loss = torch.nn.MSELoss()(model(*batch), ratings); loss.backward()
    except RuntimeError as e:
        step_ok = "set_to_none=True" in str(e)
        assert step_ok, e
    model.zero_grad(set_to_none=True)
dist.barrier()
print("rank", rank, "ok")
dist.destroy_process_group()
'''


def _torchrun(script, port, tmp_path, nproc=2, timeout=600):
    path = tmp_path / "worker.py"
    path.write_text(script)
    env = dict(os.environ, RBR_ROOT=ROOT, OMP_NUM_THREADS="2")
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                           "--master-addr", "127.0.0.1", "--master-port", str(port), str(path)],
                          env=env, capture_output=True, text=True, timeout=timeout)


def test_overlapped_table_reduction_through_encoder_backward_two_ranks_one_gpu(tmp_path):
    out = _torchrun(_GLOO_WORKER, 29741, tmp_path)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 2


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (NVLS multicast through the NVSwitch)")
def test_nvls_multimem_allreduce_matches_nccl(tmp_path):
    out = _torchrun(_NVLS_WORKER, 29743, tmp_path)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    if "SKIP" in out.stdout:
        pytest.skip("GPUs expose no NVLS multicast")
    assert out.stdout.count("ok") == 2
