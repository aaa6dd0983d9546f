"""the fused global-norm clip + Adam (csrc/optim.cu, rbr_b200.optim.FusedClipAdam) against
nn.utils.clip_grad_norm_ + torch.optim.Adam — the reference loop's trainer/train_deepconn_pp.py:135,167-168 — on the same
gradients, eager and captured in a CUDA graph with the rest of the step."""
import copy

import pytest
import torch

import rbr_b200
from conftest import rel_err
from rbr_b200 import ops, synth
from rbr_b200.graphs import GraphedTrainStep
from rbr_b200.optim import FusedClipAdam

pytestmark = pytest.mark.gpu


def _deepconn(precision):
    U, I, V, E, H, K, L = 40, 30, 600, 64, 24, 16, 96
    model = rbr_b200.DeepCoNNpp(U, I, V, [3], E, H, K, L, None, 0.0, precision=precision)
    model.load_state_dict(synth.deepconn_params(U, I, V, E, H, K, (3,), seed=3))
    batches = []
    for seed in (5, 6, 7, 8):
        b, r = synth.deepconn_batch(48, L, V, U, I, seed=seed)
        batches.append(([t.cuda() for t in b], r.cuda()))
    return model.cuda().train(), batches


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("max_norm", [5.0, 0.05])           # 5.0: the reference's value (rarely clips); 0.05: always clips
def test_fused_clip_adam_matches_torch(precision, max_norm):
    ours, batches = _deepconn(precision)
    theirs = copy.deepcopy(ours)
    opt_o = FusedClipAdam(ours, lr=0.002, max_grad_norm=max_norm)
    opt_t = torch.optim.Adam(theirs.parameters(), lr=0.002)
    loss_fn = torch.nn.MSELoss()
    for b, r in batches:
        opt_o.zero_grad()
        loss_fn(ours(*b), r).backward()
        g_o = opt_o.clip_and_step()
        opt_t.zero_grad()
        loss_fn(theirs(*b), r).backward()
        g_t = torch.nn.utils.clip_grad_norm_(theirs.parameters(), max_norm)
        opt_t.step()
        torch.cuda.synchronize()
        assert abs(float(g_o) - float(g_t)) <= 2e-5 * float(g_t)
        for (k, po), (_, pt) in zip(ours.named_parameters(), theirs.named_parameters()):
            # same gradients up to fp32 atomics order; Adam's first steps are lr * sign-like, so compare on the parameter scale
            assert float((po - pt).abs().max()) <= 2e-5 * max(1.0, float(pt.abs().max())), k
    assert int(opt_o.step_dev.item()) == len(batches)
    # parameters are views of one flat buffer; state_dict is unchanged in names and shapes
    assert set(ours.state_dict()) == set(theirs.state_dict())
    if precision == "bf16":
        # the update kept the bf16 shadow of the word table current (no re-cast kernel in the next forward)
        we = ours.word_embeddings
        key, shadow = we._shadow
        assert key == (we.embedding.weight._version, we.embedding.weight.data_ptr())
        assert torch.equal(shadow, ops.table_to_bf16(we.embedding.weight.detach()))


def test_whole_trainer_step_captured_in_a_graph_matches_eager():
    """zero_grad + forward + MSELoss + backward + clip + Adam as ONE graph replay (device-resident step counter)."""
    ours, batches = _deepconn("bf16")
    ref = copy.deepcopy(ours)
    opt_r = FusedClipAdam(ref, lr=0.002, max_grad_norm=5.0)
    loss_fn = torch.nn.MSELoss()
    ref_losses = []
    for b, r in batches:
        opt_r.zero_grad()
        loss = loss_fn(ref(*b), r)
        loss.backward()
        opt_r.clip_and_step()
        ref_losses.append(float(loss))
    opt = FusedClipAdam(ours, lr=0.002, max_grad_norm=5.0)
    state0 = {k: v.detach().clone() for k, v in ours.state_dict().items()}
    step = GraphedTrainStep(ours, loss_fn, *batches[0], optimizer=opt, max_grad_norm=5.0)
    # construction ran warm-up + capture steps: rewind parameters and optimizer state, then replay the 4 batches
    ours.load_state_dict(state0)
    opt.exp_avg.zero_(); opt.exp_avg_sq.zero_(); opt.step_dev.zero_()
    step.refresh_operands()                                  # the graph reads the shadow buffer it was captured with: re-cast in place
    losses = []
    for b, r in batches:
        losses.append(float(step(b, r)))
    torch.cuda.synchronize()
    for a, b_ in zip(losses, ref_losses):
        assert abs(a - b_) <= 1e-5 * abs(b_), (losses, ref_losses)
    for (k, po), (_, pr) in zip(ours.named_parameters(), ref.named_parameters()):
        assert float((po - pr).abs().max()) <= 2e-5 * max(1.0, float(pr.abs().max())), k
    assert int(opt.step_dev.item()) == len(batches)
