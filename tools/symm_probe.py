"""Probe torch symmetric-memory all-reduce variants (NVLS multimem, two-shot) at the gradient-arena size vs NCCL."""
import os, sys, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
w = dist.get_world_size()
n = 16_500_000
group = dist.group.WORLD
t = symm_mem.empty(n, dtype=torch.float32, device=f"cuda:{local}")
hdl = symm_mem.rendezvous(t, group)
if rank == 0:
    print("rendezvous ok; multicast_ptr", getattr(hdl, "multicast_ptr", None), "world", hdl.world_size)
gname = group.group_name
def bench(fn, label):
    try:
        for _ in range(5):
            t.fill_(1.0); fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        ok = bool((t[:1000] == float(w)).all().item()) and bool((t[-1000:] == float(w)).all().item())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 30
        if rank == 0:
            print(f"{label}: {ms*1e3:.1f} us, busbw {2*(w-1)/w*n*4/ms/1e6:.0f} GB/s, correct={ok}")
    except Exception as e:
        if rank == 0:
            print(f"{label}: FAILED {type(e).__name__}: {str(e)[:200]}")
bench(lambda: dist.all_reduce(t), "nccl sum")
bench(lambda: torch.ops.symm_mem.multimem_all_reduce_(t, "sum", gname), "multimem_all_reduce_")
bench(lambda: torch.ops.symm_mem.two_shot_all_reduce_(t, "sum", gname), "two_shot_all_reduce_")
dist.barrier()
dist.destroy_process_group()
