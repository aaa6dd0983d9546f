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
V, E, k = 50000, 300, 3
p = synth.deepconn_params(10, 10, V, E, H, 32, (k,), seed=0)
table = p["word_embeddings.embedding.weight"].cuda()
w = p["ngram.feature_layer.0.list_of_conv1d.0.weight"].cuda()
b = p["ngram.feature_layer.0.list_of_conv1d.0.bias"].cuda()
ids, mask = synth.doc_batch(n_docs, L, V, seed=1)
ids, mask = ids.cuda(), mask.cuda()
shadow = ops.table_to_bf16(table)
packed = ops.conv_pack(w)
reps = 6
for _ in range(2):
    ops.conv_act_maxpool(table, ids, mask, w, b, 1, precision=prec, shadow=shadow, packed=packed)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    feat, _ = ops.conv_act_maxpool(table, ids, mask, w, b, 1, precision=prec, shadow=shadow, packed=packed)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
fl = 2.0 * n_docs * L * H * E * k
print(f"conv[{prec}] n_docs={n_docs} L={L} H={H}: {ms:.4f} ms/launch, {fl/ms/1e9:.1f} TFLOP/s, checksum {float(feat.sum()):.4f}")
