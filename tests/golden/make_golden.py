"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules on the CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For each case: build the reference nn.Module, overwrite its weights with a seeded parameter set
(rbr_b200.synth.*_params, keyed by the reference's own state_dict names), feed a seeded synthetic
batch, run forward + nn.MSELoss + backward with dropout = 0, and store inputs, parameters, outputs
and every parameter gradient.  The fixtures pin the CPU oracle (tests/test_oracle_golden.py) and,
on the GPU box (where /root/reference does not exist), the CUDA path (tests/test_gpu_*.py).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("RBR_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
for name in ("nltk", "nltk.tokenize", "gensim", "gensim.models", "preprocess", "preprocess.divide_and_create_example_sent",
             "preprocess.divide_and_create_example_word"):   # unused imports of the reference
    m = types.ModuleType(name)
    m.word_tokenize = lambda s: s.split()
    m.KeyedVectors = object
    m.clean_str = lambda s: s
    sys.modules.setdefault(name, m)

import rbr_b200  # noqa: E402
from rbr_b200 import synth  # noqa: E402
from models.deepconn.deepconn import DeepCoNNpp  # noqa: E402
from models.narre.narre import NARRE  # noqa: E402
from models.dual_att.dual_att import DualAtt  # noqa: E402
from models.simple_siamese.simple_siamese import SimpleSiamese  # noqa: E402

torch.set_num_threads(1)
torch.use_deterministic_algorithms(True)


def _save(name, model, params, batch, ratings, outs, extra=None):
    rec = {}
    for k, v in params.items():
        rec["param/" + k] = v.numpy()
    for i, t in enumerate(batch):
        rec[f"batch/{i}"] = t.numpy()
    rec["ratings"] = ratings.numpy()
    for k, v in outs.items():
        rec["out/" + k] = v.detach().numpy()
    for k, prm in model.named_parameters():
        g = prm.grad if prm.grad is not None else torch.zeros_like(prm)
        rec["grad/" + k] = g.numpy()
    for k, v in (extra or {}).items():
        rec["meta/" + k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **rec)
    print(f"{name}: {os.path.getsize(path)/1024:.1f} KiB, loss={float(outs['loss']):.6f}")


def _run(model, params, batch, ratings, narre=False):
    sd = model.state_dict()
    assert set(sd) == set(params), (set(sd) ^ set(params))
    model.load_state_dict(params)
    model.train()
    out = model(*batch)
    outs = {}
    if narre:
        pred, u_sc, i_sc = out
        outs["u_att_scores"], outs["i_att_scores"] = u_sc, i_sc
    else:
        pred = out
    loss = torch.nn.MSELoss()(pred, ratings)
    loss.backward()
    outs["pred"], outs["loss"] = pred, loss
    return outs


def deepconn_case(name, B, L, V, U, I, E, H, K, ks, seed, tweak=None):
    params = synth.deepconn_params(U, I, V, E, H, K, ks, seed=seed)
    batch, ratings = synth.deepconn_batch(B, L, V, U, I, seed=synth.SEED_BASE + seed)
    batch = list(batch)
    if tweak:
        tweak(params, batch)
    model = DeepCoNNpp(U, I, V, list(ks), E, H, K, L, None, 0.0)
    outs = _run(model, params, batch, ratings)
    # per-side pooled features (hook-free: recompute through the module's own layers)
    with torch.no_grad():
        outs["u_rev_feats"] = model.ngram(model.word_embeddings(batch[0]), batch[2]).view(B, H)
        outs["i_rev_feats"] = model.ngram(model.word_embeddings(batch[1]), batch[3]).view(B, H)
    _save(name, model, params, batch, ratings, outs,
          dict(model="deepconn", B=B, L=L, V=V, U=U, I=I, E=E, H=H, K=K, ks=list(ks)))


def deepconn_hier_case(name, B, L, V, U, I, E, H, K, k, seed):
    """DeepCoNNpp(arch="HierPooling") — the reference's alternate encoder (models/deepconn/layers.py:62-98, 110-114)."""
    params = synth.deepconn_hier_params(U, I, V, E, H, K, seed=seed)
    batch, ratings = synth.deepconn_batch(B, L, V, U, I, seed=synth.SEED_BASE + seed)
    batch = list(batch)
    model = DeepCoNNpp(U, I, V, [k], E, H, K, L, None, 0.0, arch="HierPooling")
    outs = _run(model, params, batch, ratings)
    _save(name, model, params, batch, ratings, outs, dict(model="deepconn_hier", B=B, L=L, V=V, U=U, I=I, E=E, H=H, K=K, k=k))


def simple_siamese_case(name, B, R, T, V, U, I, E, K, use_ui_bias, latent_transform, seed):
    """The reference's SimpleSiamese (models/simple_siamese/simple_siamese.py) with every dropout at 0."""
    params = synth.simple_siamese_params(U, I, V, E, K, use_ui_bias, latent_transform, seed=seed)
    batch, ratings = synth.simple_siamese_batch(B, R, T, V, U, I, seed=synth.SEED_BASE + seed)
    model = SimpleSiamese(E, K, V, U, I, None, False, 0.0, 0.0, 0.0, use_ui_bias, latent_transform)
    sd = model.state_dict()
    assert set(sd) == set(params), (set(sd) ^ set(params))
    model.load_state_dict(params)
    model.train()
    pred, _, _ = model(*batch)
    loss = torch.nn.MSELoss()(pred, ratings)
    loss.backward()
    _save(name, model, params, batch, ratings, {"pred": pred, "loss": loss},
          dict(model="simple_siamese", B=B, R=R, T=T, V=V, U=U, I=I, E=E, K=K, ui=int(use_ui_bias), lt=int(latent_transform)))


def narre_case(name, B, R, T, V, U, I, E, H, A, K, seed):
    params = synth.narre_params(U, I, V, E, H, A, K, (3,), seed=seed)
    batch, ratings = synth.narre_batch(B, R, T, V, U, I, seed=synth.SEED_BASE + seed)
    model = NARRE(U, I, V, [3], H, E, A, K, R, T, 0.0, 0, 0, 0, None, "CNN")
    outs = _run(model, params, batch, ratings, narre=True)
    _save(name, model, params, batch, ratings, outs,
          dict(model="narre", B=B, R=R, T=T, V=V, U=U, I=I, E=E, H=H, A=A, K=K))


def dual_att_case(name, B, L, V, E, lw, lo, go, h1, h2, seed):
    params = synth.dual_att_params(V, L, lw, lo, go, E, h1, h2, seed=seed)
    batch, ratings = synth.dual_att_batch(B, L, V, seed=synth.SEED_BASE + seed)
    model = DualAtt(V, L, lw, lo, go, E, h1, h2, 0.0, None)
    outs = _run(model, params, batch, ratings)
    _save(name, model, params, batch, ratings, outs,
          dict(model="dual_att", B=B, L=L, V=V, E=E, lw=lw, lo=lo, go=go, h1=h1, h2=h2))


def edge_tweak(params, batch):
    """Edge cases the reference's behaviour defines:
    doc 0 of the user side fully padded (→ pooled == relu(bias)); doc 1 a constant doc (every
    interior position ties → gradient to the FIRST arg-max); a pad token in the middle of doc 2;
    a mask that disagrees with ids != 0 (forward must honour the mask it is given)."""
    u_revs, i_revs, u_m, i_m = batch[0], batch[1], batch[2], batch[3]
    u_revs[0] = 0
    u_m[0] = False
    u_revs[1] = 5
    u_m[1] = True
    i_revs[2, 3] = 0
    i_m[2, 3] = False
    i_m[1, 0] = False            # real token masked out by the caller
    # positive biases so that relu(bias) can win the max on padded docs
    params["ngram.feature_layer.0.list_of_conv1d.0.bias"] = \
        params["ngram.feature_layer.0.list_of_conv1d.0.bias"].abs() + 0.05


def trainer_case(kind):
    """N Adam steps through the reference's UNMODIFIED trainer loop (train_one_epoch: zero_grad → forward → MSELoss →
    backward → clip_grad_norm_(5.0) → Adam(lr 0.002).step) with the reference's own model, on the stub dataset of
    tests/trainer_harness.py: per-step losses and the final parameters."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import trainer_harness as th
    losses, final, _ = th.run_trainer_epoch(kind, model_module=None)
    rec = {"losses": np.asarray(losses, dtype=np.float64)}
    for k, v in final.items():
        rec["final/" + k] = v.numpy()
    path = os.path.join(HERE, f"trainer_{kind}_{th.N_STEPS}steps.npz")
    np.savez_compressed(path, **rec)
    print(f"trainer_{kind}: {os.path.getsize(path)/1024:.1f} KiB, losses={losses}")


if __name__ == "__main__":
    for kind in ("deepconn", "narre", "dual_att"):
        trainer_case(kind)
    deepconn_case("deepconn_small", B=4, L=20, V=60, U=9, I=7, E=12, H=8, K=6, ks=(3,), seed=1)
    deepconn_case("deepconn_edge", B=4, L=16, V=40, U=9, I=7, E=10, H=6, K=5, ks=(3,), seed=2, tweak=edge_tweak)
    deepconn_case("deepconn_multik", B=3, L=18, V=50, U=6, I=6, E=9, H=12, K=4, ks=(3, 5), seed=3)
    deepconn_case("deepconn_odd", B=5, L=37, V=80, U=11, I=13, E=20, H=10, K=7, ks=(3,), seed=4)
    deepconn_hier_case("deepconn_hier", B=5, L=24, V=70, U=9, I=7, E=12, H=8, K=6, k=3, seed=8)
    deepconn_hier_case("deepconn_hier_noproj", B=4, L=17, V=50, U=6, I=5, E=8, H=8, K=4, k=5, seed=9)
    simple_siamese_case("simple_siamese_small", B=4, R=5, T=11, V=60, U=9, I=7, E=12, K=6, use_ui_bias=True, latent_transform=False, seed=10)
    simple_siamese_case("simple_siamese_lt", B=3, R=4, T=9, V=50, U=6, I=8, E=8, K=5, use_ui_bias=False, latent_transform=True, seed=11)
    narre_case("narre_small", B=3, R=4, T=10, V=60, U=9, I=7, E=12, H=8, A=5, K=6, seed=5)
    narre_case("narre_h150ish", B=2, R=5, T=12, V=70, U=8, I=9, E=16, H=15, A=6, K=4, seed=6)
    dual_att_case("dual_att_small", B=3, L=16, V=50, E=10, lw=5, lo=8, go=6, h1=20, h2=5, seed=7)
