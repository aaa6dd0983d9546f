"""2+ GPU check of the NVLS gradient all-reduce: same averaged gradients as NCCL, and timing of one step either way."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from rbr_b200 import parallel
rank, local, world = parallel.init_from_env("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
loss_fn = torch.nn.MSELoss()
hb = bench.make_batches("deepconn", 2, rank)
db = [([t.to(dev) for t in b], r.to(dev)) for b, r in hb]
res = {}
for kind in ("nccl", "nvls", "nvls_overlap"):
    torch.manual_seed(0)
    model = bench.build("deepconn", dev, "bf16")
    model.eval()            # no dropout: identical local gradients in both runs
    model.train(False)
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    model.train()
    model.fm.dropout.p = 0.0
    parallel.broadcast_parameters(model)
    if kind.startswith("nvls"):
        ok = parallel.enable_nvls_allreduce(model, overlap=(kind == "nvls_overlap"))
        if rank == 0:
            print("enable_nvls_allreduce ->", ok)
    bench.step(model, *db[0], loss_fn, world)
    torch.cuda.synchronize()
    res[kind] = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    for _ in range(5):
        bench.step(model, *db[1], loss_fn, world)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30):
        bench.step(model, *db[i & 1], loss_fn, world)
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"{kind}: {e0.elapsed_time(e1)/30:.3f} ms/step")
# breakdown (model = the NVLS one): the collective alone, and the step without any collective on either arena
def timeit(fn, n=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
model.__dict__["_rbr_nvls"].early = None
model.ngram.table_grad_hook = None
flat = model.last_arena.flat
t_ar = timeit(lambda: parallel._nvls_allreduce(model, flat, True))
hdl = model.__dict__["_rbr_nvls"].hdl
t_bar = timeit(lambda: hdl.barrier(channel=0))
t_nccl = timeit(lambda: dist.all_reduce(flat, op=dist.ReduceOp.AVG))
t_step_symm = timeit(lambda: bench.step(model, *db[0], loss_fn, 1))
del model.__dict__["_rbr_arena_buffer"]
t_step_plain = timeit(lambda: bench.step(model, *db[0], loss_fn, 1))
if rank == 0:
    print(f"nvls allreduce alone {t_ar*1e3:.1f} us (one barrier {t_bar*1e3:.1f} us); nccl alone {t_nccl*1e3:.1f} us; "
          f"step w/o collective: symmetric arena {t_step_symm:.3f} ms, plain arena {t_step_plain:.3f} ms")
worst = 0.0
for k in res["nccl"]:
    for kind in ("nvls", "nvls_overlap"):
        a, b = res[kind][k].double(), res["nccl"][k].double()
        worst = max(worst, float((a - b).abs().max() / b.abs().max().clamp_min(1e-12)))
print(f"rank {rank}: max rel diff nvls vs nccl averaged grads = {worst:.2e}")
dist.barrier(); dist.destroy_process_group()
