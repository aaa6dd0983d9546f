"""Time one fp32 all-reduce of the gradient-arena size (66 MB) under the NCCL settings given in the environment."""
import os, sys, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 16_500_000
x = torch.ones(n, device="cuda")
op = dist.ReduceOp.SUM if os.environ.get("PROBE_OP", "avg") == "sum" else dist.ReduceOp.AVG
for _ in range(10):
    dist.all_reduce(x, op=op)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    dist.all_reduce(x, op=op)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
w = dist.get_world_size()
if rank == 0:
    print(f"op={os.environ.get('PROBE_OP','avg')} algo={os.environ.get('NCCL_ALGO','default')} proto={os.environ.get('NCCL_PROTO','default')} n={n} ({n*4/1e6:.0f} MB): {ms*1e3:.1f} us, busbw {2*(w-1)/w*n*4/ms/1e6:.0f} GB/s")
dist.destroy_process_group()
