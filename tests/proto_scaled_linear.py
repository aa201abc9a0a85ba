"""numpy model of the arithmetic the CUDA kernel uses (csrc/ctc_fused.cuh) -- test infrastructure.

It is NOT an oracle and NOT a product path: it mirrors, state for state, the *scheme* of the sm_100a
kernel so that the exponent bookkeeping can be validated on a CPU box against oracle/ctc_f64.py
before GPU minutes are spent:

  * unnormalised row softmax  p~[t,k] = exp(a[t,k] - max_k a[t,:])  (fp32 exp), row sums kept aside;
  * linear-domain fp64 alpha^/beta^ over the blank-extended states, virtual columns at t=-1 / t=T;
  * exact power-of-two rescale every K steps (chunk boundary), integer exponents Ea_c / Eb_c;
  * alpha checkpoint per chunk + recompute in the backward sweep (identical ops => identical bits);
  * posterior(t,k) = sum_{s in pos(k)} alpha^ beta^ * 2^(Ea_c+Eb_c-Ea_fin) / (Z^ p~[t,k]);
  * optional truncation of the recomputed alpha column to the high 32 bits of the double.
"""
import numpy as np

TA = 256  # target binary exponent of the column max after a rescale
TB = 256


def _hiword_round(x):
    """Keep sign/exponent/20 mantissa bits of a double (TRUNCATED; the kernel removes the mean of the
    truncation error with a factor 1 + 2^-21 on the posterior scale), like the smem alpha column."""
    u = np.asarray(x, dtype=np.float64).view(np.uint64)
    u = u & np.uint64(0xFFFFFFFF00000000)
    return u.view(np.float64) * (1.0 + 0.7213 * 2.0 ** -21)


def _rescale(v, target):
    m = v.max()
    if not (m > 0) or not np.isfinite(m):
        return v, 0
    e = int(np.floor(np.log2(m)))  # kernel: biased exponent field of the max hi-word
    sh = min(target - e, 1023)
    return np.ldexp(v, sh), -sh


def ctc_single(acts_tv, labels, blank=0, K=16, hiword=False):
    acts_tv = np.asarray(acts_tv, dtype=np.float32)
    labels = np.asarray(labels, dtype=np.int64).reshape(-1)
    T, V = acts_tv.shape
    L = len(labels)
    S = 2 * L + 1
    grad = np.zeros((T, V), dtype=np.float32)
    rep = int((labels[1:] == labels[:-1]).sum()) if L > 1 else 0
    if T == 0 or L + rep > T:
        return 0.0, grad
    ext = np.full(S, blank, dtype=np.int64)
    ext[1::2] = labels
    skip = np.zeros(S, dtype=bool)
    skip[2:] = (ext[2:] != blank) & (ext[2:] != ext[:-2])
    skip_f = np.zeros(S, dtype=bool)
    skip_f[:-2] = skip[2:]

    mx = acts_tv.max(axis=1, keepdims=True)
    pt32 = np.exp((acts_tv - mx).astype(np.float32)).astype(np.float32)   # p~, fp32
    rowsum = pt32.sum(axis=1, dtype=np.float32)
    pt = pt32.astype(np.float64)
    emit = pt[:, ext]                                                     # [T, S]

    def a_step(a, t):
        s1 = a.copy()
        s1[1:] += a[:-1]
        s1[2:] += np.where(skip[2:], a[:-2], 0.0)
        return emit[t] * s1

    def b_step(b, t):
        s1 = b.copy()
        s1[:-1] += b[1:]
        s1[:-2] += np.where(skip_f[:-2], b[2:], 0.0)
        return emit[t] * s1

    nC = (T + K - 1) // K
    a = np.zeros(S)
    a[0] = np.ldexp(1.0, TA)
    Ea = -TA
    ckpt, Ea_c = [], []
    logsum = 0.0
    sidx = np.arange(S)
    for c in range(nC):
        a = np.where(sidx < S - 2 * (T - c * K + 1), 0.0, a)   # states that can no longer reach the end
        a, de = _rescale(a, TA)
        Ea += de
        ckpt.append(a.copy())
        Ea_c.append(Ea)
        for t in range(c * K, min(T, (c + 1) * K)):
            a = a_step(a, t)
            logsum += np.log(np.float64(rowsum[t]))
    zhat = a[S - 1] + (a[S - 2] if S > 1 else 0.0)
    Ea_fin = Ea
    if not (zhat > 0) or not np.isfinite(zhat):
        grad[:] = pt32 / rowsum[:, None]
        return np.inf, grad
    cost = -(np.log(zhat) + Ea_fin * np.log(2.0) - logsum)

    b = np.zeros(S)
    b[S - 1] = np.ldexp(1.0, TB)
    Eb = -TB
    for c in range(nC - 1, -1, -1):
        t0, t1 = c * K, min(T, (c + 1) * K)
        a = ckpt[c].copy()
        acol = []
        for t in range(t0, t1):
            a = a_step(a, t)
            acol.append(_hiword_round(a) if hiword else a.copy())
        sc = np.ldexp(1.0 / zhat, Ea_c[c] + Eb - Ea_fin)
        for t in range(t1 - 1, t0 - 1, -1):
            b = b_step(b, t)
            prod = acol[t - t0] * b
            acc = np.zeros(V)
            np.add.at(acc, ext, prod)
            p32 = pt32[t] / rowsum[t]
            with np.errstate(divide="ignore", invalid="ignore"):
                post = np.where(pt32[t] > 0, (acc * sc).astype(np.float32) / pt32[t], 0.0)
            grad[t] = p32 - post.astype(np.float32)
        b = np.where(sidx > 2 * t0 + 1, 0.0, b)               # states the start cannot reach
        b, de = _rescale(b, TB)
        Eb += de
    return float(cost), grad


def ctc_batch(acts, flat_labels, act_lens, label_lens, blank=0, K=16, hiword=False):
    T_max, B, V = acts.shape
    costs = np.zeros(B)
    grads = np.zeros((T_max, B, V), dtype=np.float32)
    off = 0
    for b in range(B):
        T, L = int(act_lens[b]), int(label_lens[b])
        c, g = ctc_single(acts[:T, b], np.asarray(flat_labels)[off:off + L], blank, K, hiword)
        off += L
        costs[b] = c
        grads[:T, b] = g
    return costs, grads
