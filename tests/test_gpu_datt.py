"""GPU parity of the D-ATT row: the gate kernels (K5) and the gated tanh convolutions, through
rbr_b200.DualAtt, against the reference golden fixture and the CPU oracle.  fp32 variant 1e-5 (3e-5 on batch-summed
gradients), bf16 tensor-core variant 1e-2 (BASELINE.json)."""
import pytest
import torch

import rbr_b200
from conftest import Golden, rel_err
from oracle import rbr_oracle as orc
from rbr_b200 import synth
from rbr_b200._lib import lib
from test_gpu_parity import FP32_GRAD_TOL, FP32_TOL, BF16_TOL, run_step

pytestmark = pytest.mark.gpu


def _model(V, L, lw, lo, go, E, h1, h2, params, precision):
    model = rbr_b200.DualAtt(V, L, lw, lo, go, E, h1, h2, 0.0, None, precision=precision)
    assert set(model.state_dict()) == set(params)
    model.load_state_dict(params)
    return model.cuda()


def test_golden_dual_att_fp32():
    """pred, loss and every parameter gradient against the numbers the unmodified reference produced."""
    g = Golden("dual_att_small")
    m = g.meta
    model = _model(m["V"], m["L"], m["lw"], m["lo"], m["go"], m["E"], m["h1"], m["h2"], g.params, "fp32")
    out, loss, grads = run_step(model, g.batch, g.ratings)
    assert rel_err(out.detach().cpu(), g.out["pred"]) < FP32_TOL
    assert rel_err(loss, g.out["loss"]) < FP32_TOL
    assert set(grads) == set(g.grads)
    for k, ref in g.grads.items():
        assert grads[k].shape == ref.shape, k
        assert rel_err(grads[k], ref) < FP32_GRAD_TOL, k


@pytest.mark.parametrize("cfg", [
    # B, L, V, E, lw, lo, go, h1, h2
    (6, 120, 400, 100, 5, 200, 100, 500, 50),      # the reference's dims at a short doc
    (3, 500, 900, 100, 5, 200, 100, 500, 50),      # full doc length (global gate spans 500 tokens)
    (5, 37, 90, 20, 3, 24, 12, 30, 7),             # odd sizes, window 3
])
def test_seeded_dual_att_vs_oracle_fp32(cfg):
    B, L, V, E, lw, lo, go, h1, h2 = cfg
    params = synth.dual_att_params(V, L, lw, lo, go, E, h1, h2, seed=3)
    batch, ratings = synth.dual_att_batch(B, L, V, seed=B * 10 + L)
    model = _model(V, L, lw, lo, go, E, h1, h2, params, "fp32")
    out, loss, grads = run_step(model, batch, ratings)
    rp, rl, rg = orc.loss_and_grads("dual_att", params, batch, ratings)
    assert rel_err(out.detach().cpu(), rp) < FP32_TOL
    assert rel_err(loss, rl) < FP32_TOL
    for k in rg:
        assert rel_err(grads[k], rg[k]) < FP32_GRAD_TOL, k
    assert lib.rbr_consume_oob_count(None) == 0


def test_dual_att_encoder_features_fp32():
    """The fused encoder output (cat of local + 3 global pooled features) against the oracle's, plus the quirk the
    reference has: pad tokens are NOT masked (row 0 is zero → they contribute tanh(bias) to the max)."""
    B, L, V, E, lw, lo, go, h1, h2 = 4, 64, 300, 100, 5, 200, 100, 500, 50
    params = synth.dual_att_params(V, L, lw, lo, go, E, h1, h2, seed=5)
    batch, _ = synth.dual_att_batch(B, L, V, seed=8)
    batch[0][1] = 0                                                    # an all-padding document
    model = _model(V, L, lw, lo, go, E, h1, h2, params, "fp32")
    with torch.no_grad():
        feat = model.encode([batch[0].cuda()], ["u"])[0].cpu()
    _, aux = orc.dual_att_forward(params, batch[0], batch[1], return_aux=True)
    assert rel_err(feat, aux["u_cat"]) < FP32_TOL
    bias_cat = torch.cat([params["u_local_atten.conv.0.bias"]] + [params[f"u_global_atten.conv{c}.0.bias"] for c in (1, 2, 3)])
    assert torch.allclose(feat[1], torch.tanh(bias_cat), atol=1e-6)


def test_dual_att_bf16_tensor_core():
    """bf16 conv operands on tcgen05 (gates and their backward stay fp32): 1e-2 on predictions, loss and the
    batch-summed dense-parameter gradients."""
    B, L, V, E, lw, lo, go, h1, h2 = 16, 500, 3000, 100, 5, 200, 100, 500, 50
    params = synth.dual_att_params(V, L, lw, lo, go, E, h1, h2, seed=4)
    batch, ratings = synth.dual_att_batch(B, L, V, seed=21)
    model = _model(V, L, lw, lo, go, E, h1, h2, params, "bf16")
    out, loss, grads = run_step(model, batch, ratings)
    rp, rl, rg = orc.loss_and_grads("dual_att", params, batch, ratings)
    assert rel_err(out.detach().cpu(), rp) < BF16_TOL
    assert rel_err(loss, rl) < BF16_TOL
    for k in ("fc.0.weight", "fc.3.weight", "fc.0.bias", "u_local_atten.conv.0.bias", "i_global_atten.conv2.0.bias"):
        assert rel_err(grads[k], rg[k]) < 3 * BF16_TOL, k
    # arg-max routing can move under bf16 rounding (a whole gradient row moves with it): compare in Frobenius norm
    for k in rg:
        num = (grads[k].double() - rg[k].double()).norm()
        den = rg[k].double().norm().clamp_min(1e-12)
        assert float(num / den) < 0.15, (k, float(num / den))
    assert float(grads["word_embeddings.embedding.weight"][0].abs().max()) == 0.0


def test_dual_att_eval_and_wrong_doc_len():
    B, L, V = 4, 32, 100
    params = synth.dual_att_params(V, L, 5, 16, 8, 12, 20, 6, seed=1)
    model = _model(V, L, 5, 16, 8, 12, 20, 6, params, "fp32").eval()
    batch, _ = synth.dual_att_batch(B, L, V, seed=2)
    with torch.no_grad():
        p0 = model(batch[0].cuda(), batch[1].cuda())
    assert rel_err(p0.cpu(), orc.dual_att_forward(params, *batch)) < FP32_TOL
    with pytest.raises(ValueError):
        model(batch[0][:, :20].cuda(), batch[1][:, :20].cuda())
