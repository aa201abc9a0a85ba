"""Classifier head: float64 restatement pinned against torch's float64 BatchNorm1d + Linear +
autograd on the CPU; the fused sm_100a kernels against the restatement on the GPU."""
import numpy as np
import pytest
import torch

from oracle import head_f64


def _case(seed, N, H, V, mean_shift=0.0):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((N, H)) * rng.uniform(0.5, 3.0, H) + rng.standard_normal(H) + mean_shift).astype(np.float32)
    W = (rng.standard_normal((V, H)) / np.sqrt(H)).astype(np.float32)
    g = rng.uniform(0.5, 1.5, H).astype(np.float32)
    b = (0.1 * rng.standard_normal(H)).astype(np.float32)
    rm = rng.standard_normal(H).astype(np.float32)
    rv = rng.uniform(0.5, 2.0, H).astype(np.float32)
    dl = rng.standard_normal((N, V)).astype(np.float32)
    return x, W, g, b, rm, rv, dl


@pytest.mark.parametrize("training", [True, False])
def test_restatement_matches_torch_float64(training):
    x, W, g, b, rm, rv, dl = _case(0, 60, 24, 7, mean_shift=4.0)
    bn = torch.nn.BatchNorm1d(24).double()
    lin = torch.nn.Linear(24, 7, bias=False).double()
    with torch.no_grad():
        bn.weight.copy_(torch.tensor(g)); bn.bias.copy_(torch.tensor(b))
        bn.running_mean.copy_(torch.tensor(rm)); bn.running_var.copy_(torch.tensor(rv))
        lin.weight.copy_(torch.tensor(W))
    bn.train(training)
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    out = lin(bn(xt))
    out.backward(torch.tensor(dl, dtype=torch.float64))
    o, cache, nrm, nrv = head_f64.head_forward(x, W, g, b, rm, rv, training)
    dx, dW, dg, db = head_f64.head_backward(dl, cache)
    for got, want in ((o, out), (dx, xt.grad), (dW, lin.weight.grad), (dg, bn.weight.grad), (db, bn.bias.grad),
                      (nrm, bn.running_mean), (nrv, bn.running_var)):
        assert np.abs(got - want.detach().numpy()).max() < 1e-11
    p, _, _, _ = head_f64.head_forward(x, W, g, b, rm, rv, training, softmax=True)
    assert np.abs(p - torch.softmax(out, -1).detach().numpy()).max() < 1e-13


def test_state_dict_names_match_the_reference_layout():
    from aes_lac_2018_b200 import SequenceWiseClassifier
    keys = list(SequenceWiseClassifier(16, 5).state_dict().keys())
    # model.py:205-212: self.fc = Sequential(SequenceWise(Sequential(BatchNorm1d, Linear)))
    assert keys == ["fc.0.module.0.weight", "fc.0.module.0.bias", "fc.0.module.0.running_mean",
                    "fc.0.module.0.running_var", "fc.0.module.0.num_batches_tracked", "fc.0.module.1.weight"]


def _rel(got, want):
    want = np.asarray(want, np.float64)
    return np.abs(np.asarray(got, np.float64) - want).max() / max(1e-30, np.abs(want).max())


SHAPES = [(50, 4, 800, 29, 0.0), (33, 3, 672, 43, 0.0), (20, 2, 64, 64, 0.0), (7, 1, 16, 5, 0.0), (129, 1, 100, 33, 0.0),
          (40, 5, 800, 29, 60.0), (1, 1, 8, 2, 0.0), (300, 8, 800, 29, 3.0)]


@pytest.mark.gpu
@pytest.mark.parametrize("T,B,H,V,shift", SHAPES)
@pytest.mark.parametrize("training", [True, False])
def test_gpu_head_forward_backward_vs_restatement(T, B, H, V, shift, training):
    from aes_lac_2018_b200.head import _HeadFn
    N = T * B
    x, W, g, b, rm, rv, dl = _case(T * 1000 + V, N, H, V, shift)
    if N == 1 and training:
        pytest.skip("BatchNorm training statistics need more than one row (torch raises)")
    o, cache, nrm, nrv = head_f64.head_forward(x, W, g, b, rm, rv, training)
    dx, dW, dg, db = head_f64.head_backward(dl, cache)
    dev = "cuda"
    xt = torch.tensor(x, device=dev).view(T, B, H).requires_grad_(True)
    Wt = torch.tensor(W, device=dev).requires_grad_(True)
    gt = torch.tensor(g, device=dev).requires_grad_(True)
    bt = torch.tensor(b, device=dev).requires_grad_(True)
    rmt, rvt = torch.tensor(rm, device=dev), torch.tensor(rv, device=dev)
    out = _HeadFn.apply(xt, Wt, gt, bt, rmt, rvt, training, 1e-5, 0.1, False)
    assert out.shape == (T, B, V)
    out.backward(torch.tensor(dl, device=dev).view(T, B, V))
    assert _rel(out.detach().cpu().numpy().reshape(N, V), o) < 2e-5
    assert _rel(xt.grad.cpu().numpy().reshape(N, H), dx) < 1e-4
    assert _rel(Wt.grad.cpu().numpy(), dW) < 1e-4
    assert _rel(gt.grad.cpu().numpy(), dg) < 1e-4
    assert _rel(bt.grad.cpu().numpy(), db) < 1e-4
    assert _rel(rmt.cpu().numpy(), nrm) < 1e-5 and _rel(rvt.cpu().numpy(), nrv) < 1e-5
    # eval-mode output of the reference: softmax probabilities
    p, _, _, _ = head_f64.head_forward(x, W, g, b, rm, rv, training, softmax=True)
    with torch.no_grad():
        pt = _HeadFn.apply(xt, Wt, gt, bt, torch.tensor(rm, device=dev), torch.tensor(rv, device=dev), training, 1e-5, 0.1, True)
    # (fp32 logits of magnitude ~10 carry ~1e-6 * 10 of rounding; a probability inherits that absolute error)
    assert np.abs(pt.cpu().numpy().reshape(N, V) - p).max() < 2e-5 * max(1.0, np.abs(o).max() / 4)
    assert np.abs(pt.sum(-1).cpu().numpy() - 1).max() < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("T,B,H,V,training", [(749, 63, 800, 29, True), (300, 61, 672, 43, True), (749, 63, 800, 29, False),
                                              (1500, 40, 1000, 64, True)])
def test_gpu_head_long_pipelines_vs_torch_float64_on_the_device(T, B, H, V, training):
    """Sizes at which every CTA of the tensor-core kernels runs MANY pipeline stages (mbarrier phases wrap, both shared-memory
    stages and both accumulator buffers are reused, row blocks end inside a tile): the small parity shapes above give each
    CTA one stage.  Reference: torch's own BatchNorm1d + Linear in float64 on the same device."""
    from aes_lac_2018_b200.head import _HeadFn
    N = T * B
    gen = torch.Generator(device="cuda").manual_seed(T * 7 + V)
    x = (torch.randn(N, H, device="cuda", generator=gen) * (0.5 + 2.5 * torch.rand(H, device="cuda", generator=gen))
         + torch.randn(H, device="cuda", generator=gen) + 2.0)
    dl = torch.randn(N, V, device="cuda", generator=gen)
    bn = torch.nn.BatchNorm1d(H).cuda().double()
    lin = torch.nn.Linear(H, V, bias=False).cuda().double()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.1)
        bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2.0)
    rm, rv = bn.running_mean.float().clone(), bn.running_var.float().clone()
    bn.train(training)
    xd = x.double().requires_grad_(True)
    want = lin(bn(xd))
    want.backward(dl.double())
    xt = x.view(T, B, H).clone().requires_grad_(True)
    Wt = lin.weight.detach().float().requires_grad_(True)
    gt = bn.weight.detach().float().requires_grad_(True)
    bt = bn.bias.detach().float().requires_grad_(True)
    out = _HeadFn.apply(xt, Wt, gt, bt, rm, rv, training, 1e-5, 0.1, False)
    out.backward(dl.view(T, B, V))

    def rel(got, ref):
        return float((got.double().reshape(ref.shape) - ref).abs().max() / ref.abs().max())
    assert rel(out.detach(), want.detach()) < 2e-5
    assert rel(xt.grad, xd.grad) < 1e-4
    assert rel(Wt.grad, lin.weight.grad) < 1e-4
    assert rel(gt.grad, bn.weight.grad) < 1e-4 and rel(bt.grad, bn.bias.grad) < 1e-4
    if training:
        assert rel(rm, bn.running_mean) < 1e-5 and rel(rv, bn.running_var) < 1e-5


@pytest.mark.gpu
def test_module_is_a_drop_in_for_the_reference_head_and_feeds_the_ctc_engine():
    """Same parameters loaded into torch's own BatchNorm1d + Linear (the reference's head, float64 on the CPU) and into
    the fused module; then head -> CTCLoss -> backward end to end against restatement(head) + oracle(CTC)."""
    from aes_lac_2018_b200 import CTCLoss, SequenceWiseClassifier
    from oracle import ctc_f64
    T, B, H, V = 60, 3, 96, 29
    rng = np.random.default_rng(5)
    head = SequenceWiseClassifier(H, V).cuda()
    ref = torch.nn.Sequential(torch.nn.BatchNorm1d(H), torch.nn.Linear(H, V, bias=False))
    with torch.no_grad():
        ref[0].weight.uniform_(0.5, 1.5); ref[0].bias.normal_(0, 0.1)
        ref[0].running_mean.normal_(); ref[0].running_var.uniform_(0.5, 2.0)
    head.load_state_dict({"fc.0.module." + k: v for k, v in ref.state_dict().items()})
    ref = ref.double()
    x = (rng.standard_normal((T, B, H)) * 1.5 + 0.5).astype(np.float32)

    # eval mode: B x T x V softmax probabilities, running statistics untouched
    head.eval(); ref.eval()
    with torch.no_grad():
        p = head(torch.tensor(x).cuda())
        want = torch.softmax(ref(torch.tensor(x, dtype=torch.float64).view(T * B, H)), -1).view(T, B, V).transpose(0, 1)
    assert p.shape == (B, T, V) and np.abs(p.cpu().numpy() - want.numpy()).max() < 2e-6
    assert int(head.fc[0].module[0].num_batches_tracked) == 0

    # training mode: logits, statistics move, gradients reach every parameter through the CTC loss
    head.train(); ref.train()
    xt = torch.tensor(x).cuda().requires_grad_(True)
    out = head(xt)                                            # B x T x V view, as the reference returns
    assert out.shape == (B, T, V) and int(head.fc[0].module[0].num_batches_tracked) == 1
    label_lens = np.array([10, 0, 25], np.int32)
    labels = rng.integers(1, V, int(label_lens.sum())).astype(np.int32)
    act_lens = np.array([T, T - 7, T], np.int32)
    loss = CTCLoss()(out.transpose(0, 1), torch.tensor(labels), torch.tensor(act_lens), torch.tensor(label_lens))
    (loss / B).sum().backward()

    sd = {k: v.detach().double().numpy() for k, v in ref.state_dict().items()}
    o, cache, nrm, nrv = head_f64.head_forward(x.reshape(T * B, H), sd["1.weight"], sd["0.weight"], sd["0.bias"],
                                               sd["0.running_mean"], sd["0.running_var"], True)
    costs, grads = ctc_f64.ctc_batch(o.reshape(T, B, V).astype(np.float32), labels, act_lens, label_lens)
    dx, dW, dg, db = head_f64.head_backward(grads.reshape(T * B, V) / B, cache)
    assert abs(float(loss.detach()) - costs.sum()) < 1e-4 * costs.sum()
    bn, lin = head.fc[0].module[0], head.fc[0].module[1]
    assert _rel(xt.grad.cpu().numpy().reshape(T * B, H), dx) < 2e-4
    assert _rel(lin.weight.grad.cpu().numpy(), dW) < 2e-4
    assert _rel(bn.weight.grad.cpu().numpy(), dg) < 2e-4 and _rel(bn.bias.grad.cpu().numpy(), db) < 2e-4
    assert _rel(bn.running_mean.cpu().numpy(), nrm) < 1e-5 and _rel(bn.running_var.cpu().numpy(), nrv) < 1e-5


@pytest.mark.gpu
def test_head_rejects_what_it_cannot_do():
    from aes_lac_2018_b200 import SequenceWiseClassifier
    with pytest.raises(RuntimeError):
        SequenceWiseClassifier(8, 3)(torch.zeros(4, 2, 8))             # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        SequenceWiseClassifier(8, 100).cuda()(torch.zeros(4, 2, 8).cuda())   # more than 64 classes
