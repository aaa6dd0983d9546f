"""Per-role cycle counters of the LAST conv_tc2 launch of one D-ATT forward (RBR_TC2_DEBUG=4): who waits for whom at E = 100.
    RBR_TC2_DEBUG=4 python tools/prof_datt_conv.py"""
import ctypes, os, sys
os.environ.setdefault("RBR_TC2_DEBUG", "4")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from rbr_b200._lib import lib

dev = torch.device("cuda:0")
model = bench.build("dual_att", dev, "bf16")
(b, _), = bench.make_batches("dual_att", 1, 0)
b = [t.to(dev) for t in b]
with torch.no_grad():
    for _ in range(3):
        model(*b)
torch.cuda.synchronize()
out = np.zeros((148, 12), dtype=np.int64)
lib.check(lib.rbr_debug_conv_tc2_prof(out.ctypes.data_as(ctypes.c_void_p), 148), "prof")
lead = out[0::2]
sh = lambda a, c: f"{np.median(a / np.maximum(c, 1)) * 100:5.1f} %"
print(f"tiles per pair {np.median(lead[:, 7]):.0f}; clk per tile {np.median(lead[:, 0] / np.maximum(lead[:, 7], 1)):.0f}")
print(f"MMA warp: waits operands {sh(lead[:, 1], lead[:, 0])}, waits accumulator {sh(lead[:, 2], lead[:, 0])}")
print(f"producer warp 0: waits ring slot {sh(out[:, 4], out[:, 3])}, per-tile prologue {sh(out[:, 10], out[:, 3])}")
print(f"epilogue warp 0: waits accumulator {sh(out[:, 6], out[:, 5])}, TMEM loads + column max {sh(out[:, 8], out[:, 5])}, finalisation {sh(out[:, 9], out[:, 5])}")
