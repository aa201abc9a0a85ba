"""float64 restatement of the reference's classifier head -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Follows /root/reference/codes/model.py:177-180 (`fc = Sequential(BatchNorm1d(H), Linear(H, V, bias=False))` inside
`SequenceWise`, model.py:10-34: the T x B x H input is viewed as (T*B) x H rows) and model.py:199-207 (training:
logits; eval: softmax over the classes).  BatchNorm1d semantics are PyTorch's (torch.nn.BatchNorm1d, eps = 1e-5,
momentum = 0.1): training normalises with the biased batch variance and moves the running estimates with the
unbiased one; eval normalises with the running estimates.  The backward is the textbook BatchNorm + Linear
gradient.  Pinned in tests/test_head.py against torch's own float64 CPU BatchNorm1d + Linear + autograd (torch
is importable on both boxes), so parity here is pinned by the framework the reference itself calls.
"""
import numpy as np


def head_forward(x, weight, gamma, beta, running_mean, running_var, training=True, eps=1e-5, momentum=0.1,
                 softmax=False):
    """x: [N, H]; weight: [V, H].  Returns (out [N, V], cache, new_running_mean, new_running_var)."""
    x = np.asarray(x, np.float64)
    W = np.asarray(weight, np.float64)
    g = np.asarray(gamma, np.float64)
    b = np.asarray(beta, np.float64)
    N = x.shape[0]
    if training:
        mean = x.mean(axis=0)
        var = x.var(axis=0)                               # biased
        unb = var * (N / (N - 1)) if N > 1 else var
        new_rm = (1 - momentum) * np.asarray(running_mean, np.float64) + momentum * mean
        new_rv = (1 - momentum) * np.asarray(running_var, np.float64) + momentum * unb
    else:
        mean = np.asarray(running_mean, np.float64)
        var = np.asarray(running_var, np.float64)
        new_rm, new_rv = mean.copy(), var.copy()
    invstd = 1.0 / np.sqrt(var + eps)
    xhat = (x - mean) * invstd
    y = xhat * g + b
    out = y @ W.T
    if softmax:
        e = np.exp(out - out.max(axis=1, keepdims=True))
        out = e / e.sum(axis=1, keepdims=True)
    return out, dict(xhat=xhat, invstd=invstd, y=y, W=W, g=g, training=training), new_rm, new_rv


def head_backward(dlogits, cache):
    """Returns (dx [N, H], dW [V, H], dgamma [H], dbeta [H])."""
    dl = np.asarray(dlogits, np.float64)
    xhat, invstd, y, W, g = cache["xhat"], cache["invstd"], cache["y"], cache["W"], cache["g"]
    dW = dl.T @ y
    dy = dl @ W
    dgamma = (dy * xhat).sum(axis=0)
    dbeta = dy.sum(axis=0)
    dxhat = dy * g
    if cache["training"]:
        dx = invstd * (dxhat - dxhat.mean(axis=0) - xhat * (dxhat * xhat).mean(axis=0))
    else:
        dx = invstd * dxhat
    return dx, dW, dgamma, dbeta
