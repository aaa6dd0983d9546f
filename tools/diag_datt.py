import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import rbr_b200
from conftest import Golden, rel_err
from oracle import rbr_oracle as orc
g = Golden("dual_att_small"); m = g.meta
model = rbr_b200.DualAtt(m["V"], m["L"], m["lw"], m["lo"], m["go"], m["E"], m["h1"], m["h2"], 0.0, None, precision="fp32")
model.load_state_dict(g.params); model.cuda().train()
out = model(*[t.cuda() for t in g.batch])
loss = torch.nn.MSELoss()(out, g.ratings.cuda()); loss.backward()
for k, p in model.named_parameters():
    print(f"{k:40s} {rel_err(p.grad.cpu(), g.grads[k]):.3e}  |ref|={float(g.grads[k].abs().max()):.3e} |got|={float(p.grad.abs().max()):.3e}")
tg = model.word_embeddings.embedding.weight.grad.cpu(); rg = g.grads["word_embeddings.embedding.weight"]
d = (tg - rg).abs().max(dim=1).values
print("rows with err:", [(i, float(d[i]), float(rg[i].abs().max())) for i in torch.nonzero(d > 1e-6).flatten().tolist()][:20])
print("ids u", g.batch[0])
