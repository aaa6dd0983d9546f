"""Diagnostic: where does the bf16 path differ from the oracle run on bf16-rounded operands?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import rbr_b200
from rbr_b200 import synth, ops
from oracle import rbr_oracle as orc
from conftest import rel_err

def r16(t): return t.to(torch.bfloat16).float()
B, L, V, U, I, E, H, K = 24, 500, 3000, 50, 40, 300, 100, 32
params = synth.deepconn_params(U, I, V, E, H, K, (3,), seed=1)
batch, ratings = synth.deepconn_batch(B, L, V, U, I, seed=123)
rp = dict(params)
for k in params:
    if k == "word_embeddings.embedding.weight" or (k.startswith("ngram.") and k.endswith(".weight")):
        rp[k] = r16(params[k])
model = rbr_b200.DeepCoNNpp(U, I, V, [3], E, H, K, L, None, 0.0, precision="bf16"); model.load_state_dict(params); model.cuda().train()
pred = model(*[t.cuda() for t in batch]); loss = torch.nn.MSELoss()(pred, ratings.cuda()); loss.backward()
grads = {k: p.grad.cpu() for k, p in model.named_parameters()}
opred, oloss, og = orc.loss_and_grads("deepconn", rp, batch, ratings)
print("pred err", rel_err(pred.detach().cpu(), opred))
for k in og:
    print(k, "max-rel", rel_err(grads[k], og[k], 1e-7), "fro", float((grads[k]-og[k]).norm()/og[k].norm().clamp_min(1e-12)))
# argmax comparison
table = params["word_embeddings.embedding.weight"]; w = params["ngram.feature_layer.0.list_of_conv1d.0.weight"]; b = params["ngram.feature_layer.0.list_of_conv1d.0.bias"]
feat, amax = ops.conv_act_maxpool(table.cuda(), batch[0].cuda(), batch[2].cuda(), w.cuda(), b.cuda(), 1, precision="bf16")
x = orc.mask_rows(orc.embedding_gather(r16(table), batch[0]), batch[2])
y = orc.conv1d_same(x, r16(w), b)
ref, ref_arg = orc.first_argmax_pool(y)
mism = (amax.cpu().long() != ref_arg)
print("argmax mismatches", int(mism.sum()), "of", mism.numel())
got = torch.gather(y, 1, amax.cpu().long().unsqueeze(1)).squeeze(1)
gap = ((ref - got).abs() / ref.abs().clamp_min(1e-6))[mism]
print("rel gaps at mismatches:", gap.sort().values[-10:])
print("active (relu>0) mismatches:", int((mism & (ref > 0)).sum()))
