"""Profiling driver: a few launches of the conv kernel (K2) at BASELINE configs[1] size (one side: 4096 docs x 500 tokens)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rbr_b200
from rbr_b200 import ops, synth

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
n_docs = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
L = int(sys.argv[3]) if len(sys.argv) > 3 else 500
H = int(sys.argv[4]) if len(sys.argv) > 4 else 100
V, E = 50000, int(os.environ.get("RBR_PROF_E", "300"))
k = int(sys.argv[5]) if len(sys.argv) > 5 else 3
p = synth.deepconn_params(10, 10, V, E, H, 32, (k,), seed=0)
table = p["word_embeddings.embedding.weight"].cuda()
w = p["ngram.feature_layer.0.list_of_conv1d.0.weight"].cuda()
b = p["ngram.feature_layer.0.list_of_conv1d.0.bias"].cuda()
ids, mask = synth.doc_batch(n_docs, L, V, seed=1, uniform=(os.environ.get("RBR_UNIFORM", "0") == "1"))
ids, mask = ids.cuda(), mask.cuda()
shadow = ops.table_to_bf16(table)
packed = ops.conv_pack(w)
reps = 12
for _ in range(2):
    ops.conv_act_maxpool(table, ids, mask, w, b, 1, precision=prec, shadow=shadow, packed=packed) if False else ops.conv_act_maxpool(table, ids, mask, w, b, (k - 1) // 2, precision=prec, shadow=shadow, packed=packed)
torch.cuda.synchronize()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
evs[0].record()
for i in range(reps):
    feat, _ = ops.conv_act_maxpool(table, ids, mask, w, b, (k - 1) // 2, precision=prec, shadow=shadow, packed=packed)
    evs[i + 1].record()
torch.cuda.synchronize()
per = [evs[i].elapsed_time(evs[i + 1]) for i in range(reps)]
ms = sorted(per)[reps // 2]   # median: the first launches after warm-up can include allocator hiccups
print("per-launch ms:", " ".join(f"{x:.3f}" for x in per))
fl = 2.0 * n_docs * L * H * E * k
print(f"conv[{prec}] n_docs={n_docs} L={L} H={H}: {ms:.4f} ms/launch, {fl/ms/1e9:.1f} TFLOP/s, checksum {float(feat.sum()):.4f}")
