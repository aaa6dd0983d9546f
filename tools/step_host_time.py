"""Host-side (Python + launch) time per training step vs device time: is the step CPU-bound?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
model_name = sys.argv[1] if len(sys.argv) > 1 else "deepconn"
dev = torch.device("cuda", 0)
model = bench.build(model_name, dev, "bf16")
hb = bench.make_batches(model_name, 4, 0)
db = [([t.to(dev) for t in b], r.to(dev)) for b, r in hb]
loss_fn = torch.nn.MSELoss()
for i in range(5):
    bench.step(model, *db[i % 4], loss_fn, 1)
torch.cuda.synchronize()
for K in (20, 50):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(K):
        bench.step(model, *db[i % 4], loss_fn, 1)
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{model_name} K={K}: host enqueue {1e3*(t1-t0)/K:.3f} ms/step, device {e0.elapsed_time(e1)/K:.3f} ms/step, wall {1e3*(t2-t0)/K:.3f}")
# with a sync every step (pure GPU time per step + launch latency)
ts = []
for i in range(10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    bench.step(model, *db[i % 4], loss_fn, 1)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    ts.append((t1 - t0, t2 - t0))
print("per-step host %.3f ms, host+drain %.3f ms" % (1e3 * sorted(t[0] for t in ts)[5], 1e3 * sorted(t[1] for t in ts)[5]))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(20):
    bench.step(model, *db[i % 4], loss_fn, 1)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
