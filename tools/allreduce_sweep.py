"""2+ GPUs: the gradient exchange alone on a 66 MB symmetric-memory arena — own multimem kernel, own peer load/store kernel (CTA
sweep), the two barriers, and ncclAllReduce.   torchrun --nproc-per-node N tools/allreduce_sweep.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

from rbr_b200 import parallel
from rbr_b200._lib import lib

rank, local, world = parallel.init_from_env("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
n = 16152896 // 1024 * 1024 + 1024
buf = symm_mem.empty(n, dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
mc = int(hdl.multicast_ptr or 0)
ptrs = [int(x) for x in hdl.buffer_ptrs]
peers = (ctypes.c_uint64 * len(ptrs))(*ptrs)
s = torch.cuda.current_stream().cuda_stream


def timeit(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t) * 1e3


def mm(ctas):
    hdl.barrier(channel=0)
    lib.check(lib.rbr_multimem_allreduce_f32(mc, n, rank, world, 1.0 / world, ctas, s), "mm")
    hdl.barrier(channel=1)


def p2p(ctas):
    hdl.barrier(channel=0)
    lib.check(lib.rbr_p2p_allreduce_f32(ctypes.cast(peers, ctypes.c_void_p), 0, n, rank, world, 1.0 / world, ctas, s), "p2p")
    hdl.barrier(channel=1)


def bars():
    hdl.barrier(channel=0)
    hdl.barrier(channel=1)


# correctness of p2p against NCCL
buf.copy_(torch.randn(n, device=dev) * (rank + 1))
exp = buf.clone(); dist.all_reduce(exp, op=dist.ReduceOp.AVG)
torch.cuda.synchronize(); dist.barrier()
p2p(0); torch.cuda.synchronize()
err = float((buf - exp).abs().max() / exp.abs().max())
out = [f"world {world}, arena {n * 4 / 1e6:.1f} MB; p2p vs ncclAllReduce(AVG): max rel err {err:.2e}"]
buf.zero_()
out.append(f"two barriers alone: {timeit(bars):.1f} us")
if mc:
    for c in (16, 32, 64):
        out.append(f"multimem, {c:3d} CTAs: {timeit(lambda: mm(c)):.1f} us")
for c in (16, 32, 64, 96, 128, 148):
    out.append(f"p2p,      {c:3d} CTAs: {timeit(lambda: p2p(c)):.1f} us")
out.append(f"ncclAllReduce: {timeit(lambda: dist.all_reduce(buf, op=dist.ReduceOp.AVG)):.1f} us")
if rank == 0:
    print("\n".join(out))
dist.barrier()
dist.destroy_process_group()
