"""float64 CTC oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product path
(``aes_lac_2018_b200``) never does and fails loudly when its CUDA library is absent.

What it restates
----------------
The CTC negative log-likelihood and its gradient w.r.t. *unnormalised* activations,
exactly as the reference obtains them from ``warpctc_pytorch.CTCLoss`` at
``/root/reference/codes/engine.py:22`` and ``/root/reference/codes/metrics.py:51``
(imported at ``train.py:12`` / ``metrics.py:3``, constructed with defaults at
``train.py:179`` / ``metrics.py:43``).  The algorithm itself lives in the third-party
dependency SeanNaren/warp-ctc, cloned at **unpinned HEAD** by
``/root/reference/docker/Dockerfile:52-66`` and absent from ``/root/reference``; what is
restated here is its published algorithm (Graves et al. 2006 forward-backward in log
space, as organised by warp-ctc's ``CpuCTC``: internal softmax, beta that *includes* the
emission at t, ``L + repeats > T`` => cost 0 / gradient untouched), written out from the
maths in ``SURVEY.md`` Appendix C.

PARITY UNPINNED by the reference itself: the reference has no tests and ships no golden
vectors for this path, and warp-ctc cannot be installed offline.  The oracle is instead
pinned (tests/test_oracle.py) against (a) upstream warp-ctc's published known-answer
vectors (``tests/golden/warpctc_known_answers.json``) and (b) an independent
implementation, ``torch.nn.functional.ctc_loss`` in float64 with ``zero_infinity=True``
(``tests/golden/make_golden.py`` generated the committed fixtures).

Everything is float64 and deliberately slow/simple: one Python loop over t per utterance,
numpy vectorised over the 2L+1 blank-extended states.
"""
from __future__ import annotations

import numpy as np

NEG_INF = -np.inf


def log_softmax_rows(x: np.ndarray) -> np.ndarray:
    """Row-wise log of the softmax *probabilities*, float64.

    warp-ctc softmaxes internally (reference README.md:168) and keeps probabilities, taking
    log(p) at each use, so a logit far enough below the row max has p == 0 and log p == -inf
    (fp32: gap > ~103; here in float64: gap > ~745).  Following that form (rather than
    z - log(sum)) keeps the oracle's behaviour on -1e30-style inputs the same as the reference's.
    """
    x = np.asarray(x, dtype=np.float64)
    m = x.max(axis=-1, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0)
    e = np.exp(x - m)
    with np.errstate(divide="ignore"):
        return np.log(e / e.sum(axis=-1, keepdims=True))


def count_repeats(labels: np.ndarray) -> int:
    labels = np.asarray(labels)
    if labels.size < 2:
        return 0
    return int((labels[1:] == labels[:-1]).sum())


def extend_labels(labels: np.ndarray, blank: int) -> np.ndarray:
    """l' = blank, l0, blank, l1, ..., blank  (S = 2L+1)."""
    L = int(len(labels))
    ext = np.full(2 * L + 1, blank, dtype=np.int64)
    ext[1::2] = labels
    return ext


def _shift(v: np.ndarray, k: int) -> np.ndarray:
    """v shifted towards higher s by k (k>0) or lower s (k<0), filled with -inf."""
    out = np.full_like(v, NEG_INF)
    n = len(v)
    if k >= 0:
        if k < n:
            out[k:] = v[:n - k]
    elif -k < n:
        out[:n + k] = v[-k:]
    return out


def ctc_single(acts_tv: np.ndarray, labels: np.ndarray, blank: int = 0,
               warpctc_zero_quirk: bool = False, return_internals: bool = False):
    """Cost and d(cost)/d(acts) for ONE utterance.

    acts_tv : [T, V] unnormalised activations for the valid frames only.
    labels  : [L] int labels (not validated against V, like warp-ctc).
    Returns (cost, grad[T, V]) in float64.  Infeasible (L + repeats > T) => (0.0, zeros),
    the warp-ctc CPU convention.
    """
    acts_tv = np.asarray(acts_tv, dtype=np.float64)
    labels = np.asarray(labels, dtype=np.int64).reshape(-1)
    T, V = acts_tv.shape
    L = len(labels)
    S = 2 * L + 1
    grad = np.zeros((T, V), dtype=np.float64)
    if T == 0 or L + count_repeats(labels) > T:
        return (0.0, grad) if not return_internals else (0.0, grad, None)

    lp = log_softmax_rows(acts_tv)               # [T, V]
    ext = extend_labels(labels, blank)           # [S]
    # skip transition s-2 -> s allowed for non-blank states whose label differs from l'[s-2]
    skip = np.zeros(S, dtype=bool)
    if S > 2:
        skip[2:] = (ext[2:] != blank) & (ext[2:] != ext[:-2])
    emit = lp[:, ext]                            # [T, S] log p_t(l'_s)

    with np.errstate(invalid="ignore"):
        alpha = np.full((T, S), NEG_INF)
        alpha[0, 0] = emit[0, 0]
        if S > 1:
            alpha[0, 1] = emit[0, 1]
        for t in range(1, T):
            prev = alpha[t - 1]
            acc = np.logaddexp(prev, _shift(prev, 1))
            acc = np.where(skip, np.logaddexp(acc, _shift(prev, 2)), acc)
            alpha[t] = acc + emit[t]
        logz = alpha[T - 1, S - 1]
        if S > 1:
            logz = np.logaddexp(logz, alpha[T - 1, S - 2])

        # beta includes the emission at t (warp-ctc convention)
        beta = np.full((T, S), NEG_INF)
        beta[T - 1, S - 1] = emit[T - 1, S - 1]
        if S > 1:
            beta[T - 1, S - 2] = emit[T - 1, S - 2]
        skip_fwd = np.zeros(S, dtype=bool)       # s -> s+2 allowed
        if S > 2:
            skip_fwd[:-2] = skip[2:]
        for t in range(T - 2, -1, -1):
            nxt = beta[t + 1]
            acc = np.logaddexp(nxt, _shift(nxt, -1))
            acc = np.where(skip_fwd, np.logaddexp(acc, _shift(nxt, -2)), acc)
            beta[t] = acc + emit[t]

        cost = -float(logz)
        p = np.exp(lp)
        if not np.isfinite(logz):
            # no path has non-zero probability: cost = +inf; posterior undefined.  We define the
            # gradient as the bare softmax (no NaN), see DESIGN.md "degenerate inputs".
            grad[:] = p
            return (cost, grad) if not return_internals else (cost, grad, (alpha, beta, lp))

        ab = alpha + beta                        # [T, S]
        acc = np.full((T, V), NEG_INF)
        for s in range(S):
            acc[:, ext[s]] = np.logaddexp(acc[:, ext[s]], ab[:, s])
        with np.errstate(over="ignore"):
            post = np.exp(acc - lp - logz)
        use_p = ~np.isfinite(acc) | (p == 0.0)
        if warpctc_zero_quirk:
            use_p |= (acc == 0.0)
        post = np.where(use_p, 0.0, post)
        grad[:] = p - post
    if return_internals:
        return cost, grad, (alpha, beta, lp)
    return cost, grad


def ctc_batch(acts: np.ndarray, flat_labels, act_lens, label_lens, blank: int = 0,
              warpctc_zero_quirk: bool = False):
    """Batch version with the layout the reference call uses.

    acts       : [T_max, B, V] (time-major, as `_sanitize_inputs` produces, engine.py:12-16)
    flat_labels: [sum L_b] concatenated labels (data.py:157)
    act_lens   : [B], label_lens : [B]
    Returns (costs[B] float64, grads[T_max, B, V] float64); frames t >= act_lens[b] get 0.
    """
    acts = np.asarray(acts, dtype=np.float64)
    T_max, B, V = acts.shape
    flat_labels = np.asarray(flat_labels, dtype=np.int64).reshape(-1)
    act_lens = np.asarray(act_lens, dtype=np.int64).reshape(-1)
    label_lens = np.asarray(label_lens, dtype=np.int64).reshape(-1)
    assert len(act_lens) == B and len(label_lens) == B
    costs = np.zeros(B, dtype=np.float64)
    grads = np.zeros_like(acts)
    off = 0
    for b in range(B):
        T, L = int(act_lens[b]), int(label_lens[b])
        lab = flat_labels[off:off + L]
        off += L
        c, g = ctc_single(acts[:T, b, :], lab, blank, warpctc_zero_quirk)
        costs[b] = c
        grads[:T, b, :] = g
    return costs, grads


def ctc_loss_module(acts, flat_labels, act_lens, label_lens, blank=0,
                    size_average=False, length_average=False):
    """What `warpctc_pytorch.CTCLoss.forward` returns: (sum_b cost_b, grads), with the
    averaging flags of the upstream binding: size_average divides cost and
    gradient by B, length_average by sum(act_lens) and supersedes."""
    costs, grads = ctc_batch(acts, flat_labels, act_lens, label_lens, blank)
    total = costs.sum()
    if length_average:
        d = float(np.asarray(act_lens).sum())
        total, grads = total / d, grads / d
    elif size_average:
        d = float(acts.shape[1])
        total, grads = total / d, grads / d
    return total, grads
