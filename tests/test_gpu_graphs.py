"""GPU tests of rbr_b200.graphs.GraphedTrainStep: a captured-and-replayed training step gives the same loss and gradients
as the eager nn.Module path on every batch fed to it, and the FM dropout mask still changes from replay to replay."""
import pytest
import torch

import rbr_b200
from conftest import rel_err
from rbr_b200 import synth
from rbr_b200.graphs import GraphedTrainStep

pytestmark = pytest.mark.gpu


def _model(dropout, precision="fp32"):
    U, I, V, E, H, K, L = 40, 30, 600, 64, 24, 16, 96
    params = synth.deepconn_params(U, I, V, E, H, K, (3,), seed=3)
    model = rbr_b200.DeepCoNNpp(U, I, V, [3], E, H, K, L, None, dropout, precision=precision)
    model.load_state_dict(params)
    return model.cuda().train(), (U, I, V, L)


def _eager(model, batch, ratings):
    model.zero_grad(set_to_none=True)
    loss = torch.nn.MSELoss()(model(*batch), ratings)
    loss.backward()
    return float(loss), {k: p.grad.detach().clone() for k, p in model.named_parameters()}


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_graphed_step_matches_eager_on_every_batch(precision):
    model, (U, I, V, L) = _model(0.0, precision)
    batches = []
    for seed in (5, 6, 7):
        b, r = synth.deepconn_batch(32, L, V, U, I, seed=seed)
        batches.append(([t.cuda() for t in b], r.cuda()))
    ref = [_eager(model, b, r) for b, r in batches]
    step = GraphedTrainStep(model, torch.nn.MSELoss(), *batches[0])
    for (b, r), (ref_loss, ref_grads) in zip(batches, ref):
        # inputs from pinned HOST memory for one of them: H2D straight into the graph's static buffers
        src = [t.cpu().pin_memory() for t in b] if ref_loss == ref[1][0] else b
        loss = step(src, r)
        torch.cuda.synchronize()
        assert abs(float(loss) - ref_loss) <= 1e-6 * max(1.0, abs(ref_loss))
        for k, p in model.named_parameters():
            assert rel_err(p.grad.cpu(), ref_grads[k].cpu()) < 1e-5, k        # fp32 atomics land in a different order


def test_graphed_step_dropout_mask_changes_between_replays():
    model, (U, I, V, L) = _model(0.5)
    b, r = synth.deepconn_batch(64, L, V, U, I, seed=9)
    b, r = [t.cuda() for t in b], r.cuda()
    step = GraphedTrainStep(model, torch.nn.MSELoss(), b, r)
    losses = []
    for _ in range(4):
        losses.append(float(step.replay()))
    assert len({round(x, 6) for x in losses}) > 1, losses          # same inputs, different FM dropout masks
    model.eval()
    with torch.no_grad():
        p1, p2 = model(*b), model(*b)
    assert torch.equal(p1, p2)


def _eager_step(model, batch, ratings):
    model.zero_grad(set_to_none=True)
    out = model(*batch)
    pred = out[0] if isinstance(out, tuple) else out
    loss = torch.nn.MSELoss()(pred, ratings)
    loss.backward()
    return float(loss), {k: p.grad.detach().clone() for k, p in model.named_parameters()}


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("fuse_mse", [True, False])
def test_graphed_step_narre_matches_eager(precision, fuse_mse):
    """NARRE: two attention sides on two streams (forward and backward) inside the capture."""
    U, I, V, E, H, A, K, R, T = 40, 30, 600, 64, 48, 16, 16, 6, 20
    params = synth.narre_params(U, I, V, E, H, A, K, (3,), seed=4)
    model = rbr_b200.NARRE(U, I, V, [3], H, E, A, K, R, T, 0.0, 0, 0, 0, None, "CNN", precision=precision)
    model.load_state_dict(params)
    model.cuda().train()
    batches = []
    for seed in (5, 6, 7):
        b, r = synth.narre_batch(24, R, T, V, U, I, seed=seed)
        batches.append(([t.cuda() for t in b], r.cuda()))
    ref = [_eager_step(model, b, r) for b, r in batches]
    step = GraphedTrainStep(model, torch.nn.MSELoss(), *batches[0], fuse_mse=fuse_mse)
    assert step.fused_loss == fuse_mse
    for rep in range(2):                                         # every batch twice: replays do not leak state
        for (b, r), (ref_loss, ref_grads) in zip(batches, ref):
            loss = step(b, r)
            torch.cuda.synchronize()
            assert abs(float(loss) - ref_loss) <= 2e-6 * max(1.0, abs(ref_loss))
            for k, p in model.named_parameters():                # replay vs eager: the fp32 atomics land in a different order
                assert rel_err(p.grad.cpu(), ref_grads[k].cpu(), 1e-9) < 1e-5, k


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_graphed_step_dual_att_matches_eager(precision):
    """D-ATT: the item side's backward runs on a side stream inside the capture; loss through torch's own MSELoss."""
    V, L, E = 500, 64, 32
    params = synth.dual_att_params(V, L, 5, 24, 16, E, 40, 10, seed=5)
    model = rbr_b200.DualAtt(V, L, 5, 24, 16, E, 40, 10, 0.0, None, precision=precision)
    model.load_state_dict(params)
    model.cuda().train()
    batches = []
    for seed in (5, 6, 7):
        b, r = synth.dual_att_batch(16, L, V, seed=seed)
        batches.append(([t.cuda() for t in b], r.cuda()))
    ref = [_eager_step(model, b, r) for b, r in batches]
    step = GraphedTrainStep(model, torch.nn.MSELoss(), *batches[0])
    assert not step.fused_loss                                   # DualAtt has no FM head: nn.MSELoss stays a torch op
    for rep in range(2):
        for (b, r), (ref_loss, ref_grads) in zip(batches, ref):
            loss = step(b, r)
            torch.cuda.synchronize()
            assert abs(float(loss) - ref_loss) <= 2e-6 * max(1.0, abs(ref_loss))
            for k, p in model.named_parameters():
                assert rel_err(p.grad.cpu(), ref_grads[k].cpu(), 1e-9) < 2e-6, k


@pytest.mark.parametrize("wire", ["uint16", "int32"])
@pytest.mark.parametrize("model_name", ["deepconn", "narre", "dual_att"])
def test_staged_inputs_int32_ids_and_derived_masks_match_the_reference_wire_format(model_name, wire):
    """one pinned arena per step (token ids uint16 — the vocabulary fits — or int32, masks derived on the device
    as ids != 0) uploaded with one copy into the graph's static inputs gives the same loss and gradients as int64 ids + bool
    masks, eager and graphed."""
    if model_name == "deepconn":
        model, (U, I, V, L) = _model(0.0, "bf16")
        mk = lambda seed: synth.deepconn_batch(32, L, V, U, I, seed=seed)
    elif model_name == "narre":
        U, I, V, E, H, A, K, R, T = 40, 30, 600, 64, 48, 16, 16, 6, 20
        model = rbr_b200.NARRE(U, I, V, [3], H, E, A, K, R, T, 0.0, 0, 0, 0, None, "CNN", precision="bf16")
        model.load_state_dict(synth.narre_params(U, I, V, E, H, A, K, (3,), seed=4))
        model.cuda().train()
        mk = lambda seed: synth.narre_batch(24, R, T, V, U, I, seed=seed)
    else:
        V, L, E = 500, 64, 32
        model = rbr_b200.DualAtt(V, L, 5, 24, 16, E, 40, 10, 0.0, None, precision="bf16")
        model.load_state_dict(synth.dual_att_params(V, L, 5, 24, 16, E, 40, 10, seed=5))
        model.cuda().train()
        mk = lambda seed: synth.dual_att_batch(16, L, V, seed=seed)
    if wire == "int32":                                                    # as for a vocabulary of more than 65536 rows
        model.staging_spec = dict(model.staging_spec, vocab=1 << 20)
    host = [mk(s) for s in (5, 6, 7)]
    ref = [_eager_step(model, [t.cuda() for t in b], r.cuda()) for b, r in host]
    step = GraphedTrainStep(model, torch.nn.MSELoss(), [t.cuda() for t in host[0][0]], host[0][1].cuda(), staged=True)
    st = step.staged
    ref_bytes = sum(t.numel() * t.element_size() for t in host[0][0]) + host[0][1].numel() * 4
    assert st.token_dtype == (torch.uint16 if wire == "uint16" else torch.int32)
    assert st.h2d_bytes < (0.52 if wire == "int32" else 0.30) * ref_bytes + 4096     # narrow ids, no masks
    assert all(v is None or v.dtype != torch.bool for v in st.batch)
    for (b, r), (ref_loss, ref_grads) in zip(host, ref):
        arena = st.pack(b, r)                                              # pinned host arena in wire format
        assert arena.is_pinned() and arena.numel() == st.h2d_bytes
        step.load_packed(arena)
        loss = step.replay()
        torch.cuda.synchronize()
        assert abs(float(loss) - ref_loss) <= 2e-6 * max(1.0, abs(ref_loss))
        for k, p in model.named_parameters():
            assert rel_err(p.grad.cpu(), ref_grads[k].cpu(), 1e-9) < 1e-5, k
    # eager call on the staged views (the public nn.Module API with int32 ids and masks=None)
    st.upload(st.pack(*host[1]))
    out = model(*st.batch)
    pred = out[0] if isinstance(out, tuple) else out
    ref_out = model(*[t.cuda() for t in host[1][0]])
    ref_pred = ref_out[0] if isinstance(ref_out, tuple) else ref_out
    assert torch.equal(pred, ref_pred)
