"""GPU tests of the loss glue and of the non-blocking path: `sanitize_loss` against the
reference's `_sanitize_loss` flow (/root/reference/codes/engine.py:12-32, 77, 84) restated on the float64 oracle,
CTC_B200_FLAG_NO_SYNC with the device-side log-space detour, NaN activations, repeated backward."""
import numpy as np
import pytest
import torch

from tests.helpers import synth_problem

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-4
GRAD_ATOL = 1e-5


def _loader_batch(seed, T, B, V, lmin, lmax, tmin):
    """What the reference's loader hands to the trainer: B x T x V model output, flat targets, input_percentages."""
    acts, labels, al, ll = synth_problem(seed, T, B, V, lmin, lmax, tmin=tmin)
    pct = torch.tensor(al.astype(np.float64) / T, dtype=torch.float32)
    al_ref = (pct * T).int().numpy()                                   # engine.py:16 (float32 product, truncation)
    return acts, labels, al_ref, ll, pct


def test_sanitize_loss_matches_reference_flow():
    from aes_lac_2018_b200 import CTCLoss, _lib, sanitize_loss
    from oracle import ctc_f64
    acts, labels, al, ll, pct = _loader_batch(61, 150, 6, 29, 5, 40, 90)
    B = acts.shape[1]
    weight = 0.3
    out = torch.tensor(acts).cuda().transpose(0, 1).contiguous().requires_grad_()      # B x T x V, like model(inputs)
    n0 = _lib.launch_count()
    loss, status = sanitize_loss(CTCLoss(), out * 1.0, torch.tensor(labels), pct, torch.tensor(ll), average=B,
                                 weight=weight, return_status=True)
    assert loss.is_cuda and loss.dim() == 0 and status.is_cuda
    loss.backward()
    assert _lib.launch_count() > n0
    oc, og = ctc_f64.ctc_loss_module(acts, labels, al, ll)
    want = weight * oc / B
    assert abs(loss.item() - want) <= LOSS_RTOL * abs(want)
    got = out.grad.transpose(0, 1).cpu().numpy()
    assert np.abs(got - weight * og / B).max() <= GRAD_ATOL
    assert not status.cpu().numpy().any()
    # a non-unit upstream factor still arrives (the scale kernel runs for real)
    out2 = torch.tensor(acts).cuda().transpose(0, 1).contiguous().requires_grad_()
    l2 = sanitize_loss(CTCLoss(), out2, torch.tensor(labels), pct, torch.tensor(ll), average=B)
    (2.5 * l2).backward()
    assert np.abs(out2.grad.transpose(0, 1).cpu().numpy() - 2.5 * og / B).max() <= 2.5 * GRAD_ATOL
    # the fused gradient is single-use
    out3 = torch.tensor(acts).cuda().transpose(0, 1).contiguous().requires_grad_()
    l3 = sanitize_loss(CTCLoss(), out3, torch.tensor(labels), pct, torch.tensor(ll), average=B)
    l3.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="consumed"):
        l3.backward()


def test_sanitize_loss_inf_guard_on_device():
    """engine.py:27-30: an infinite batch loss is replaced by 0 and the step carries no gradient."""
    from aes_lac_2018_b200 import CTCLoss, sanitize_loss
    acts, labels, al, ll, pct = _loader_batch(62, 60, 3, 15, 4, 10, 60)
    labels[0] = 2
    acts[:, 0, 2] = -1e30                                               # utterance 0 cannot be aligned: cost +inf
    out = torch.tensor(acts).cuda().transpose(0, 1).contiguous().requires_grad_()
    loss, status = sanitize_loss(CTCLoss(), out, torch.tensor(labels), pct, torch.tensor(ll), average=3, return_status=True)
    loss.backward()
    assert loss.item() == 0.0
    assert not out.grad.any()
    assert status[0].item() & 0x2


def test_no_sync_path_and_device_detour():
    """CTC_B200_FLAG_NO_SYNC: costs and status stay on the device; out-of-range utterances are still redone in log
    space (the detour is a kernel behind the fast ones), so the numbers match the blocking call."""
    from aes_lac_2018_b200 import ctc_loss_raw
    from oracle import ctc_f64
    rng = np.random.default_rng(7)
    acts, labels, al, ll = synth_problem(63, 300, 5, 29, 40, 100)
    acts[:, 1] *= 40.0                                                  # utterance 1: far outside the linear-domain range
    wrong = rng.integers(1, 29, 300)
    acts[np.arange(300), 3, wrong] = 65.0                               # utterance 3: confident and wrong
    a = torch.tensor(acts).cuda()
    args = [torch.tensor(x) for x in (labels, al, ll)]
    oc, og = ctc_f64.ctc_batch(acts, labels, al, ll)
    for mode in ("warp32", "warp", "latency", "throughput8"):
        c_d, g_d, s_d = ctc_loss_raw(a, *args, mode=mode, no_sync=True)
        assert c_d.is_cuda and s_d.is_cuda
        c_b, g_b, s_b = ctc_loss_raw(a, *args, mode=mode)
        torch.cuda.synchronize()
        assert torch.equal(c_d.cpu(), c_b) and torch.equal(g_d, g_b) and torch.equal(s_d.cpu(), s_b), mode
        st = s_b.numpy()
        assert st[1] & 0x10 and not (st & 0x8).any(), (mode, st)             # (utterance 3 is inside the warp ladder's range)
        if mode == "warp32":                                                 # ... and is redone by the fp64 tier of the fp32 ladder
            assert st[3] & 0x20 and st[1] & 0x20, (mode, st)
        elif mode != "warp":
            assert st[3] & 0x10, (mode, st)
        rel = np.abs(c_b.numpy() - oc) / np.maximum(1.0, np.abs(oc))
        assert rel.max() <= LOSS_RTOL and np.abs(g_b.cpu().numpy() - og).max() <= GRAD_ATOL, mode


def test_nan_activations_give_nan_not_an_error():
    """Hostile data is not an invalid argument: upstream returns a NaN cost and carries on (the reference's trainer
    only tests for inf, engine.py:27); so does the engine, with the status bit set."""
    from aes_lac_2018_b200 import CTCLoss, ctc_loss_raw
    acts, labels, al, ll = synth_problem(64, 80, 4, 29, 5, 30)
    acts[17, 2, 5] = np.nan
    a = torch.tensor(acts).cuda()
    for mode in ("warp32", "warp", "latency", "throughput8"):
        c, g, st = ctc_loss_raw(a, torch.tensor(labels), torch.tensor(al), torch.tensor(ll), mode=mode)
        assert np.isnan(c[2].item()) and st[2].item() & 0x8, mode
        ok = [0, 1, 3]
        assert torch.isfinite(c[ok]).all() and torch.isfinite(g[:, ok]).all() and not st[ok].any(), mode
    loss = CTCLoss()(a.requires_grad_(), torch.tensor(labels), torch.tensor(al), torch.tensor(ll))
    assert torch.isnan(loss).all()


def test_repeated_backward_does_not_compound():
    """`mul_` in backward compounded grad_output on a second backward through the same node."""
    from aes_lac_2018_b200 import CTCLoss
    from oracle import ctc_f64
    acts, labels, al, ll = synth_problem(65, 70, 3, 29, 5, 25)
    _, og = ctc_f64.ctc_loss_module(acts, labels, al, ll)
    x = torch.tensor(acts).cuda().requires_grad_()
    loss = CTCLoss()(x, torch.tensor(labels), torch.tensor(al), torch.tensor(ll))
    (loss / 3).sum().backward(retain_graph=True)
    g1 = x.grad.clone()
    x.grad = None
    (loss / 3).sum().backward()
    assert torch.equal(x.grad, g1)
    assert np.abs(g1.cpu().numpy() - og / 3).max() <= GRAD_ATOL


def test_workspace_release_and_bidirectional_budget():
    """the small-batch bidirectional path must not ask for gigabytes of column spill on long utterances,
    and cached workspaces can be dropped."""
    import ctypes
    from aes_lac_2018_b200 import _lib, ctc_loss, ctc_loss_raw, release_workspaces
    from oracle import ctc_f64
    lib = _lib.load()
    B, T, L, V = 96, 1500, 400, 29
    ll = np.full(B, L, np.int32)
    al = np.full(B, T, np.int32)
    need = ctypes.c_size_t(0)
    assert lib.ctc_b200_workspace_size(ll.ctypes.data, al.ctypes.data, V, B, T, 1, ctypes.byref(need)) == 0
    assert need.value < (700 << 20), need.value                        # was ~1.2 GB with the unconditional spill
    acts, labels, al2, ll2 = synth_problem(66, 400, 2, 29, 150, 180)     # long transcript, tiny batch: still exact
    c, g, st = ctc_loss_raw(torch.tensor(acts).cuda(), torch.tensor(labels), torch.tensor(al2), torch.tensor(ll2))
    oc, og = ctc_f64.ctc_batch(acts, labels, al2, ll2)
    assert np.abs(g.cpu().numpy() - og).max() <= GRAD_ATOL
    release_workspaces()
    assert not ctc_loss._dev_ws and not ctc_loss._host_ws
