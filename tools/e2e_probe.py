"""Where does the end-to-end loop lose time against the device-timed replay?  Variants of bench.py's e2e loop on one GPU.
    python tools/e2e_probe.py [deepconn|narre] [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from rbr_b200.graphs import GraphedTrainStep

name = sys.argv[1] if len(sys.argv) > 1 else "deepconn"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 200
dev = torch.device("cuda", 0)
model = bench.build(name, dev, "bf16")
NB = 4
host = bench.make_batches(name, NB, 0)
devb = [([t.to(dev) for t in b], r.to(dev)) for b, r in host]
loss_fn = torch.nn.MSELoss()
HOST_LOSS = os.environ.get("RBR_PROBE_HOST_LOSS", "1") == "1"      # loss read-back as the graph's last node
steps, pool = [], None
for j in range(2):
    g = GraphedTrainStep(model, loss_fn, *devb[j], warmup=1, pool=pool, staged=True, host_loss=HOST_LOSS)
    pool = g.pool
    steps.append(g)
packed = [steps[0].staged.pack(b, r) for b, r in host]
copy_stream = torch.cuda.Stream(device=dev)
loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
B = bench.CFG[name]["B"]


def loop(n, d2h=True, depth=1, h2d=True):
    main = torch.cuda.current_stream()
    free_ev = [None, None]

    def load(i):
        o = steps[i & 1]
        if free_ev[i & 1] is not None:
            copy_stream.wait_event(free_ev[i & 1])
        o.load_packed(packed[i % NB], stream=copy_stream)
        ev = torch.cuda.Event()
        ev.record(copy_stream)
        return ev
    nxt = load(0) if h2d else None
    pend = []
    for i in range(n):
        ev = nxt
        if h2d:
            if i + 1 < n:
                nxt = load(i + 1)
            main.wait_event(ev)
        loss = steps[i & 1].replay()
        if d2h:
            loss_host[i & 1].copy_(loss.detach(), non_blocking=True)
        done = torch.cuda.Event()
        done.record()
        free_ev[i & 1] = done
        pend.append(done)
        if len(pend) > depth:
            pend.pop(0).synchronize()
    for e in pend:
        e.synchronize()


def loop_graph_only(n):
    """bench.py's loop: nothing but graph launches on the compute stream; the host waits for each upload."""
    def load(i):
        steps[i & 1].load_packed(packed[i % NB], stream=copy_stream)
        ev = torch.cuda.Event()
        ev.record(copy_stream)
        return ev
    done = [None, None]
    seen = 0.0
    load(0).synchronize()
    for i in range(n):
        steps[i & 1].replay()
        d = torch.cuda.Event()
        d.record()
        done[i & 1] = d
        if i + 1 < n:
            if done[(i + 1) & 1] is not None:
                done[(i + 1) & 1].synchronize()
                if HOST_LOSS:
                    seen += float(steps[(i + 1) & 1].loss_host[0])
            load(i + 1).synchronize()
    torch.cuda.synchronize()
    return seen


def timed(label, fn=None, **kw):
    loop = fn or globals()["loop"]
    loop(300, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loop(K, **kw)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / K
    print(f"{label:58s} {dt * 1e3:7.3f} ms/step  {B / dt / 1e6:6.3f} M samples/s")


timed("replay only (no H2D, no D2H), host 1 step ahead", d2h=False, h2d=False)
timed("replay + D2H loss", d2h=True, h2d=False)
timed("H2D + replay, no D2H", d2h=False, h2d=True)
timed("bench.py's loop: H2D + replay + D2H, host 1 step ahead", d2h=True, h2d=True)
timed("the same, host 2 steps ahead", d2h=True, h2d=True, depth=2)
timed("the same, host 1 step ahead, again", d2h=True, h2d=True)
timed("graph launches only on the compute stream (bench.py's loop)", fn=loop_graph_only)
timed("replay only, again", d2h=False, h2d=False)
# H2D alone
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(50):
    steps[i & 1].load_packed(packed[i % NB], stream=copy_stream)
copy_stream.synchronize()
dt = (time.perf_counter() - t0) / 50
print(f"H2D alone: {steps[0].staged.h2d_bytes / 1e6:.1f} MB in {dt * 1e3:.3f} ms = {steps[0].staged.h2d_bytes / dt / 1e9:.1f} GB/s")
