"""GPU parity of the tcgen05 (bf16) conv kernel.  Two references:
 (1) the CPU oracle evaluated on bf16-ROUNDED table and weights — isolates the kernel logic (tiling, tap shifts,
     document packing, masks, arg-max) from the quantisation, so the tolerance can be tight (1e-3);
 (2) the fp32 oracle / reference goldens at BASELINE.json's bf16 tolerance (1e-2 relative)."""
import pytest
import torch

import rbr_b200
from conftest import Golden, grad_floor, rel_err
from oracle import rbr_oracle as orc
from rbr_b200 import ops, synth
from test_gpu_parity import DEEPCONN_CASES, NARRE_CASES, build_model, run_step

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[ops.CONV_TC_SINGLE_CTA, ops.CONV_TC_PAIR_ONLY], ids=["cta1-cpasync", "pair-tma"])
def tc_flags(request):
    """Tests taking this fixture run on BOTH tensor-core kernels: conv_tc.cu (single CTA, cp.async operands) and
    conv_tc2.cu (CTA pair, cta_group::2, TMA gather4 operands), selected per call through the C-ABI's `flags` argument
    (either flag errors out instead of falling back)."""
    return request.param


def _bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


# (n_docs, L, V, E, H, k)  — covers: one K-step / many, one tile / many (long docs), several docs per tile (short docs),
# filter passes (H=150), ragged tail tiles, k=1 (no tap shift) / 3 / 5
SHAPES = [
    (5, 40, 50, 16, 16, 1),
    (5, 40, 50, 16, 16, 3),
    (7, 128, 80, 32, 24, 3),
    (3, 129, 80, 48, 40, 3),
    (9, 500, 300, 300, 100, 3),
    (33, 60, 300, 300, 150, 3),
    (40, 10, 60, 24, 20, 3),
    (6, 300, 200, 100, 200, 5),
    (300, 500, 2000, 300, 100, 3),
    # edge shapes: one document, documents shorter than the kernel's tile and than a warp, one filter, filters split into
    # passes (H=300), tiny / odd embedding widths, k=7 halo, odd number of documents for the CTA pair
    (1, 500, 300, 300, 100, 3),
    (1, 3, 20, 8, 1, 3),
    (3, 1, 20, 8, 5, 1),
    (65, 17, 90, 40, 33, 3),
    (11, 200, 400, 300, 300, 3),
    (9, 260, 300, 72, 48, 7),
    (130, 31, 500, 100, 200, 5),
]


@pytest.mark.parametrize("shape", SHAPES)
def test_conv_tc_vs_bf16_rounded_oracle(shape, tc_flags):
    n, L, V, E, H, k = shape
    gen = torch.Generator().manual_seed(n * 1000 + L)
    table = torch.randn(V, E, generator=gen)
    table[0] = 0
    w = (torch.rand(H, E, k, generator=gen) * 2 - 1) / (E * k) ** 0.5
    b = (torch.rand(H, generator=gen) * 2 - 1) * 0.1
    ids, mask = synth.doc_batch(n, L, V, seed=L + 5)
    mask = mask.clone()
    mask[0, min(3, L - 1)] = False                      # a real token masked out by the caller
    if n > 2:
        ids[2] = 0
        mask[2] = False                                  # fully padded document → relu(bias)
    feat, amax = ops.conv_act_maxpool(table.cuda(), ids.cuda(), mask.cuda(), w.cuda(), b.cuda(), (k - 1) // 2,
                                      precision="bf16", flags=tc_flags)
    x = orc.mask_rows(orc.embedding_gather(_bf16_round(table), ids), mask)
    y = orc.conv1d_same(x, _bf16_round(w), b)
    ref, ref_arg = orc.first_argmax_pool(torch.relu(y))
    assert rel_err(feat.cpu(), ref) < 1e-3
    # arg-max: the position the kernel reports attains the oracle's max up to fp32 summation-order noise
    pre = y.max(dim=1).values
    got = torch.gather(y, 1, amax.cpu().long().unsqueeze(1)).squeeze(1)
    assert float(((pre - got).abs() / pre.abs().clamp_min(1e-3)).max()) < 1e-5
    if n > 2:
        assert torch.allclose(feat[2].cpu(), torch.relu(b), atol=1e-6)


def test_conv_tc_tanh_and_no_mask(tc_flags):
    gen = torch.Generator().manual_seed(4)
    V, E, H, k, n, L = 100, 64, 32, 3, 6, 200
    table = torch.randn(V, E, generator=gen)
    w = (torch.rand(H, E, k, generator=gen) * 2 - 1) / (E * k) ** 0.5
    b = (torch.rand(H, generator=gen) * 2 - 1) * 0.1
    ids = torch.randint(0, V, (n, L), generator=gen)
    feat, _ = ops.conv_act_maxpool(table.cuda(), ids.cuda(), None, w.cuda(), b.cuda(), 1, act=ops.ACT_TANH, precision="bf16",
                                   flags=tc_flags)
    y = orc.conv1d_same(orc.embedding_gather(_bf16_round(table), ids), _bf16_round(w), b)
    assert rel_err(feat.cpu(), torch.tanh(y).max(dim=1).values) < 1e-3


def _rounded_params(params):
    """What the tensor-core kernel actually multiplies: bf16-rounded embedding table and conv weights."""
    out = dict(params)
    for k in params:
        if k == "word_embeddings.embedding.weight" or (k.startswith("ngram.") and k.endswith(".weight")):
            out[k] = _bf16_round(params[k])
    return out


def _fro_err(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-12))


@pytest.mark.parametrize("case", DEEPCONN_CASES + NARRE_CASES)
def test_golden_bf16(case, tc_flags):
    """End-to-end bf16 step.  Outputs against the reference's own numbers at the 1e-2 tolerance.  Gradients:
    (a) within 1e-2 of the oracle run on the bf16-rounded operands (same arg-max routing as the kernel);
    (b) against the fp32 reference on the Frobenius norm — bf16 rounding can move an arg-max to a near-tied
        position, which moves a whole gradient row, so max-abs is not a stable metric there."""
    g = Golden(case)
    model = build_model(g, "bf16")
    model.ngram.conv_flags = tc_flags
    out, loss, grads = run_step(model, g.batch, g.ratings)
    pred = out[0] if isinstance(out, tuple) else out
    assert rel_err(pred.detach().cpu(), g.out["pred"]) < 1e-2
    assert rel_err(loss, g.out["loss"]) < 1e-2
    rp, rl, rg = orc.loss_and_grads(g.model, _rounded_params(g.params), g.batch, g.ratings)
    assert rel_err(pred.detach().cpu(), rp) < 2e-3
    for k, ref in g.grads.items():
        assert rel_err(grads[k], rg[k], max(grad_floor(k), 1e-6)) < 1e-2, k
        if ref.abs().max() > 1e-6:
            assert _fro_err(grads[k], ref) < 1e-1, k


def test_seeded_midsize_bf16_vs_oracle(tc_flags):
    B, L, V, U, I, E, H, K = 24, 500, 3000, 50, 40, 300, 100, 32
    params = synth.deepconn_params(U, I, V, E, H, K, (3,), seed=1)
    batch, ratings = synth.deepconn_batch(B, L, V, U, I, seed=123)
    model = rbr_b200.DeepCoNNpp(U, I, V, [3], E, H, K, L, None, 0.0, precision="bf16")
    model.load_state_dict(params)
    model.cuda()
    model.ngram.conv_flags = tc_flags
    out, loss, grads = run_step(model, batch, ratings)
    rp, rl, rg = orc.loss_and_grads("deepconn", params, batch, ratings)
    assert rel_err(out.detach().cpu(), rp) < 1e-2
    assert rel_err(loss, rl) < 1e-2
    for k in rg:
        assert _fro_err(grads[k], rg[k]) < 1.5e-1, k      # arg-max flips under bf16 rounding move whole rows
    _, _, rgr = orc.loss_and_grads("deepconn", _rounded_params(params), batch, ratings)
    for k in rgr:
        assert rel_err(grads[k], rgr[k], 1e-7) < 1e-2, k


def test_bf16_full_batch_properties(tc_flags):
    """B=4096-sized run of the tensor-core path: deterministic, mask-invariant, permutation-equivariant, and
    equal (within bf16 tolerance) to the fp32 CUDA-core variant on the same inputs."""
    B, L, V, U, I, E, H, K = 1024, 500, 50000, 2000, 1200, 300, 100, 32
    params = synth.deepconn_params(U, I, V, E, H, K, (3,), seed=2)
    batch, _ = synth.deepconn_batch(B, L, V, U, I, seed=77)
    b = [t.cuda() for t in batch]
    preds = {}
    for prec in ("bf16", "fp32"):
        model = rbr_b200.DeepCoNNpp(U, I, V, [3], E, H, K, L, None, 0.5, precision=prec)
        model.load_state_dict(params)
        model.cuda().eval()
        model.ngram.conv_flags = tc_flags if prec == "bf16" else 0
        with torch.no_grad():
            preds[prec] = model(*b)
            if prec == "bf16":
                assert torch.equal(model(*b), preds[prec])
                scr = b[0].clone()
                scr[~b[2]] = 17
                assert torch.equal(model(scr, b[1], b[2], b[3], b[4], b[5]), preds[prec])
                perm = torch.randperm(B, device="cuda")
                assert torch.equal(model(*[t[perm] for t in b]), preds[prec][perm])
    assert rel_err(preds["bf16"].cpu(), preds["fp32"].cpu()) < 1e-2


def test_variants_agree_bitwise_on_values():
    """The two kernels issue the same K-step / tap order into fp32 TMEM accumulators: identical pooled values."""
    gen = torch.Generator().manual_seed(12)
    V, E, H, k, n, L = 4000, 300, 100, 3, 300, 500
    table = torch.randn(V, E, generator=gen)
    w = (torch.rand(H, E, k, generator=gen) * 2 - 1) / (E * k) ** 0.5
    b = (torch.rand(H, generator=gen) * 2 - 1) * 0.1
    ids, mask = synth.doc_batch(n, L, V, seed=99)
    outs = []
    for fl in (ops.CONV_TC_SINGLE_CTA, ops.CONV_TC_PAIR_ONLY):
        outs.append(ops.conv_act_maxpool(table.cuda(), ids.cuda(), mask.cuda(), w.cuda(), b.cuda(), 1, precision="bf16", flags=fl))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])
