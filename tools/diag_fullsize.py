"""Diagnostic (GPU): where does the fp32 gradient error at B=4096 come from?  Compares ours / the fp32 oracle on the GPU against the fp64 oracle."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rbr_b200
from rbr_b200 import synth
from oracle import rbr_oracle as orc

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
c = dict(B=B, L=500, V=50000, E=300, H=100, K=32, U=20000, I=12000)
params = synth.deepconn_params(c["U"], c["I"], c["V"], c["E"], c["H"], c["K"], (3,), seed=0)
batch, ratings = synth.deepconn_batch(c["B"], c["L"], c["V"], c["U"], c["I"], seed=synth.SEED_BASE)
batch, ratings = [t.cuda() for t in batch], ratings.cuda()
model = rbr_b200.DeepCoNNpp(c["U"], c["I"], c["V"], [3], c["E"], c["H"], c["K"], c["L"], None, 0.0, precision="fp32")
model.load_state_dict(params); model.cuda().train()
with torch.no_grad():
    _, _, ua, ia = model.ngram.encode(model.word_embeddings, batch[:2], batch[2:4], return_argmax=True)
model.zero_grad(set_to_none=True)
loss = torch.nn.MSELoss()(model(*batch), ratings); loss.backward()
g = {k: p.grad.detach().double() for k, p in model.named_parameters()}
p64 = {k: v.cuda().double() for k, v in params.items()}
p32 = {k: v.cuda() for k, v in params.items()}
_, _, r64 = orc.loss_and_grads("deepconn", p64, batch, ratings.double(), argmax_override=(ua, ia))
_, _, r32 = orc.loss_and_grads("deepconn", p32, batch, ratings, argmax_override=(ua, ia))
def rel(a, b): return float((a.double() - b.double()).abs().max() / b.double().abs().max())
for k in r64:
    print(f"{k:50s} ours {rel(g[k], r64[k]):.2e}   fp32-ATen {rel(r32[k], r64[k]):.2e}   max|ref| {float(r64[k].abs().max()):.3e}")
k = "word_embeddings.embedding.weight"
d = (g[k] - r64[k]).abs()
row = int(d.max(dim=1).values.argmax())
cnt = int((batch[0] == row).sum() + (batch[1] == row).sum())
print("worst row", row, "occurrences", cnt, "err", float(d[row].max()), "ref row max", float(r64[k][row].abs().max()), "ours", float(g[k][row].abs().max()))
rowmax = r64[k].abs().max(dim=1).values
print("row with max |grad|:", int(rowmax.argmax()), float(rowmax.max()))
# error relative to each row's own magnitude for the hottest rows
for r in (3, 4, 5, 10, 100, 1000):
    print(" row", r, "rel-to-row", float(d[r].max() / r64[k][r].abs().max().clamp_min(1e-30)), "ATen", float((r32[k][r].double() - r64[k][r]).abs().max() / r64[k][r].abs().max().clamp_min(1e-30)))
