"""Where do the conv_tc2 roles wait?  Runs the CTA-pair conv on one bench-sized side with RBR_TC2_DEBUG=4 and prints, per role,
the share of its cycles spent waiting on the pipeline barriers (median over CTAs).
    RBR_TC2_DEBUG=4 python tools/prof_conv_roles.py [deepconn|narre]"""
import ctypes
import os
import sys

os.environ.setdefault("RBR_TC2_DEBUG", "4")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from rbr_b200 import ops
from rbr_b200._lib import lib

name = sys.argv[1] if len(sys.argv) > 1 else "deepconn"
c = bench.CFG[name]
dev = torch.device("cuda:0")
model = bench.build(name, dev, "bf16")
NROT = int(os.environ.get("RBR_PROF_ROTATE", "1"))      # > 1: rotate over that many batches (x 2 sides) so ids / masks come from HBM
rot = []
for b, _ in bench.make_batches(name, NROT, 0):
    for s in (0, 1):
        i_, m_ = b[s].to(dev), b[2 + s].to(dev)
        if name == "narre":
            i_, m_ = i_.view(-1, c["T"]), m_.view(-1, c["T"])
        rot.append((i_, m_))
ids, mask = rot[0]
we = model.word_embeddings
conv = model.ngram.conv
w0, b0 = conv.list_of_conv1d[0].weight.detach(), conv.list_of_conv1d[0].bias.detach()
args = dict(act=ops.ACT_RELU, precision="bf16", shadow=we.bf16_shadow(), packed=conv.packed(0))
for _ in range(3):
    ops.conv_act_maxpool(we.embedding.weight.detach(), ids, mask, w0, b0, 1, **args)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.conv_act_maxpool(we.embedding.weight.detach(), ids, mask, w0, b0, 1, **args)
e1.record()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for it in range(20):
        ops.conv_act_maxpool(we.embedding.weight.detach(), *rot[it % len(rot)], w0, b0, 1, **args)
    torch.cuda.synchronize()
durs = [ev.device_time for ev in prof.events() if ev.device_type.name == "CUDA" and "conv_tc2_kernel" in ev.name]
n = 148
out = np.zeros((n, 12), dtype=np.int64)
lib.check(lib.rbr_debug_conv_tc2_prof(out.ctypes.data_as(ctypes.c_void_p), n), "prof")
lead = out[0::2]
print(f"{name}: launch {e0.elapsed_time(e1) * 1e3:.1f} us (with the pre-pass kernels); tiles per pair median {np.median(lead[:, 7]):.0f}")
def share(a, b):
    return f"{np.median(a / np.maximum(b, 1)) * 100:5.1f} %"
print(f"MMA warp (leaders)   total {np.median(lead[:, 0]):9.0f} clk | waits operands {share(lead[:, 1], lead[:, 0])} | waits accumulator {share(lead[:, 2], lead[:, 0])}")
print(f"producer warp 0      total {np.median(out[:, 3]):9.0f} clk | waits ring slot {share(out[:, 4], out[:, 3])} | per-tile row-index prologue {share(out[:, 10], out[:, 3])}")
print(f"epilogue warp 0      total {np.median(out[:, 5]):9.0f} clk | waits accumulator {share(out[:, 6], out[:, 5])}")
if durs:
    med = float(np.median(durs))
    print(f"conv_tc2_kernel alone, 20 back-to-back launches over {len(rot)} input sets (CUPTI): median {med:.1f} us, min {min(durs):.1f} us "
          f"-> {np.median(lead[:, 0]) / med / 1e3:.3f} GHz effective SM clock")
print(f"epilogue warp 0      TMEM loads + column max {share(out[:, 8], out[:, 5])} | finalisation (barrier + stores) {share(out[:, 9], out[:, 5])}")
print(f"clk per tile (MMA warp): {np.median(lead[:, 0] / np.maximum(lead[:, 7], 1)):.0f}")
