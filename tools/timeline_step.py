"""Per-kernel timeline of the GRAPHED training step at real clocks (CUPTI activity records through torch.profiler: no kernel
replay, no clock control — unlike ncu the durations are those of warm, back-to-back launches).  Prints the average duration per
kernel and the step's GPU-busy time against its wall time.
    python tools/timeline_step.py deepconn|narre [replays] [--full]   (--full: trainer step with clip + Adam in the graph)"""
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from rbr_b200.graphs import GraphedTrainStep

name = sys.argv[1] if len(sys.argv) > 1 else "deepconn"
reps = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 20
dev = torch.device("cuda", 0)
model = bench.build(name, dev, "bf16")
batches = [([t.to(dev) for t in b], r.to(dev)) for b, r in bench.make_batches(name, 2, 0)]
loss_fn = torch.nn.MSELoss()
graphs, pool = [], None
for b, r in batches:
    g = GraphedTrainStep(model, loss_fn, b, r, warmup=1, pool=pool)
    pool = g.pool
    graphs.append(g)
for i in range(200):
    graphs[i % 2].replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(reps):
    graphs[i % 2].replay()
e1.record()
torch.cuda.synchronize()
plain_ms = e0.elapsed_time(e1) / reps
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(reps):
        graphs[i % 2].replay()
    torch.cuda.synchronize()
agg = defaultdict(lambda: [0, 0.0])
t_min, t_max = None, None
for ev in prof.events():
    if ev.device_type.name != "CUDA":
        continue
    n = ev.name
    agg[n][0] += 1
    agg[n][1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    tr = ev.time_range
    t_min = tr.start if t_min is None else min(t_min, tr.start)
    t_max = tr.end if t_max is None else max(t_max, tr.end)
busy = sum(v[1] for v in agg.values())
print(f"{name}: {reps} replays, {plain_ms * 1e3:.1f} us per step unprofiled; under CUPTI: span {(t_max - t_min) / reps:.1f} us per step, "
      f"sum of kernel durations {busy / reps:.1f} us per step")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {t / reps:8.1f} us/step  {c / reps:5.1f} x {t / c:8.1f} us  {n[:110]}")
