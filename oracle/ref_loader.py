"""Loader of the vendored UNMODIFIED reference (oracle/_ref, produced by oracle/Makefile) — TEST / BASELINE INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs (`cpu_baseline`, `--impl reference`) import this.  It puts
oracle/_ref on sys.path and stubs the imports the reference makes but never uses on this path: nltk (models/dual_att/
dual_att.py:4), gensim (trainer/train_*.py:16, used only with use_pretrain) and preprocess.* (trainer/train_*.py:21,
`clean_str`, used only by load_pretrained_embeddings).  Nothing here edits a reference file.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "models", "deepconn", "deepconn.py"))


def _stub(name: str, **attrs) -> None:
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m


def activate() -> None:
    """Make `models.*`, `utils`, `experiment`, `trainer.*` importable from oracle/_ref."""
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `make -C oracle` where /root/reference exists (build() does)")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    _stub("nltk", word_tokenize=lambda s: s.split())
    _stub("nltk.tokenize", word_tokenize=lambda s: s.split())
    _stub("gensim")
    _stub("gensim.models", KeyedVectors=object)
    _stub("preprocess")
    _stub("preprocess.divide_and_create_example_sent", clean_str=lambda s: s)
    _stub("preprocess.divide_and_create_example_word", clean_str=lambda s: s)
    _stub("preprocess.divide_and_create_example_doc", clean_str=lambda s: s)


def reference_classes():
    """(DeepCoNNpp, NARRE, DualAtt) of the unmodified reference."""
    activate()
    d = importlib.import_module("models.deepconn.deepconn")
    n = importlib.import_module("models.narre.narre")
    a = importlib.import_module("models.dual_att.dual_att")
    return d.DeepCoNNpp, n.NARRE, a.DualAtt


def build_reference(model: str, cfg: dict, params: dict, dropout: float = 0.0):
    """Reference nn.Module for `model` ("deepconn" | "narre" | "dual_att") with `params` (state_dict names) loaded."""
    import contextlib
    D, N, A = reference_classes()
    with contextlib.redirect_stdout(sys.stderr):            # the reference's constructors print; stdout belongs to bench.py's JSON line
        return _build(D, N, A, model, cfg, params, dropout)


def _build(D, N, A, model, cfg, params, dropout):
    if model == "deepconn":
        m = D(cfg["U"], cfg["I"], cfg["V"], list(cfg["ks"]), cfg["E"], cfg["H"], cfg["K"], cfg["L"], None, dropout)
    elif model == "narre":
        m = N(cfg["U"], cfg["I"], cfg["V"], list(cfg["ks"]), cfg["H"], cfg["E"], cfg["A"], cfg["K"], cfg["R"], cfg["T"], dropout,
              0, 0, 0, None, "CNN")
    elif model == "dual_att":
        m = A(cfg["V"], cfg["L"], cfg["lw"], cfg["lo"], cfg["go"], cfg["E"], cfg["h1"], cfg["h2"], dropout, None)
    else:
        raise ValueError(model)
    m.load_state_dict(params)
    return m
