"""Profiling driver: a few eager training steps (zero_grad + forward + MSELoss + backward + fused clip/Adam) of one model at its
BASELINE.json config, so that `ncu -k regex:<kernel> --launch-skip N -c 1` can pick any kernel of the step by name.
    python tools/prof_step.py deepconn|narre|dual_att [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from rbr_b200.optim import FusedClipAdam
from rbr_b200 import ops

name = sys.argv[1] if len(sys.argv) > 1 else "deepconn"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda", 0)
model = bench.build(name, dev, "bf16")
batches = [([t.to(dev) for t in b], r.to(dev)) for b, r in bench.make_batches(name, 2, 0)]
opt = FusedClipAdam(model, lr=0.002, max_grad_norm=5.0)
loss_fn = torch.nn.MSELoss()
for i in range(steps):
    b, r = batches[i % 2]
    opt.zero_grad()
    out = model(*b)
    loss = loss_fn(out[0] if isinstance(out, tuple) else out, r)
    loss.backward()
    opt.clip_and_step()
if name != "dual_att":
    # standalone K1 gather (the bit-exact nn.Embedding forward) for its dram__bytes
    table = model.word_embeddings.embedding.weight.detach()
    ids = batches[0][0][0]
    for _ in range(3):
        ops.gather_rows(table, ids)
torch.cuda.synchronize()
print("done", float(loss))
