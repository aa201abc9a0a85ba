"""numpy model of the arithmetic of the one-warp-per-utterance kernel (csrc/ctc_warp.cuh) -- test infrastructure.

NOT an oracle and NOT a product path: it mirrors the *scheme* of the sm_100a kernel so that the exponent
bookkeeping and the error budget can be validated on a CPU box against oracle/ctc_f64.py:

  * RATIO domain: r[t,k] = exp(a[t,k] - a[t,blank]) (fp32 exp on the fraction, kept as the HIGH 32 BITS of the
    double), so every blank state is a plain add: alpha^_t(blank s) = alpha^_{t-1}(s) + alpha^_{t-1}(s-1);
    label states multiply by r.  cost = -(log Z^ + Ea*ln2 - sum_t log s_t), s_t = sum_k r[t,k];
  * linear fp64 alpha^/beta^ with an exact power-of-two rescale per K-step chunk;
  * alpha checkpoint per chunk kept as ROUNDED high words (32 bits), recompute inside the chunk;
    the recomputed alpha of the LABEL states is kept as TRUNCATED high words (mean-corrected);
  * posterior_t(k) = sum_{s in pos(k)} alpha^_t(s) * tb_t(s) * 2^(..) / Z^   where tb is the beta sum BEFORE
    the multiplication by r (no division); blank posterior = 1 - sum of the others;
  * range self-check once per chunk: sum_s pre_s * beta^_{t0}(s) must equal Z^ (pre = alpha sum before * r).
"""
import numpy as np

TE = 192           # target binary exponent of the column max after a rescale
MEANC = 1.0 + 0.7213 * 2.0 ** -21


def hi_trunc(x):
    u = np.asarray(x, dtype=np.float64).view(np.uint64) & np.uint64(0xFFFFFFFF00000000)
    return u.view(np.float64)


def hi_round(x):
    u = np.asarray(x, dtype=np.float64).view(np.uint64)
    u = (u + np.uint64(0x80000000)) & np.uint64(0xFFFFFFFF00000000)
    return u.view(np.float64)


def _rescale(v, target):
    m = v.max()
    if not (m > 0) or not np.isfinite(m):
        return v, 0
    e = int(np.floor(np.log2(m)))
    sh = min(target - e, 1023)
    return np.ldexp(v, sh), -sh


def ratios(acts_tv, blank):
    """r hi-words (float64 holding 21 significant bits), p = softmax (fp32), log s (fp64 sum of fp32 logs)."""
    a = np.asarray(acts_tv, dtype=np.float32)
    d = (a - a[:, blank:blank + 1]).astype(np.float32)
    with np.errstate(over="ignore"):
        r = np.exp(d.astype(np.float64))
    r = (r * (1 + (np.random.default_rng(0).random(r.shape) - 0.5) * 2.0 ** -22))   # ex2.approx-sized noise
    r = hi_trunc(r)
    emax = np.floor(np.log2(r.max(axis=1)))                   # row max exponent
    rs = np.ldexp(r, (-emax).astype(np.int64)[:, None]).astype(np.float32)          # scaled, fp32
    s = rs.sum(axis=1, dtype=np.float32)
    p = (rs / s[:, None]).astype(np.float32)
    logs = np.log(s.astype(np.float64)).astype(np.float32).astype(np.float64) + emax * np.log(2.0)
    return r, p, logs


def ctc_single(acts_tv, labels, blank=0, K=8, return_check=False):
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
    is_lab = np.zeros(S, dtype=bool)
    is_lab[1::2] = True
    skip = np.zeros(S, dtype=bool)
    skip[2:] = is_lab[2:] & (ext[2:] != ext[:-2])
    skip_f = np.zeros(S, dtype=bool)
    skip_f[:-2] = skip[2:]

    r, p32, logs = ratios(acts_tv, blank)
    emit = r[:, ext]
    emit[:, ~is_lab] = 1.0

    def a_pre(a):
        s1 = a.copy()
        s1[1:] += a[:-1]
        s1[2:] += np.where(skip[2:], a[:-2], 0.0)
        return s1

    def b_pre(b):
        s1 = b.copy()
        s1[:-1] += b[1:]
        s1[:-2] += np.where(skip_f[:-2], b[2:], 0.0)
        return s1

    nC = (T + K - 1) // K
    a = np.zeros(S)
    a[0] = np.ldexp(1.0, TE)
    Ea = -TE
    ckpt, Ea_c = [], []
    sidx = np.arange(S)
    for c in range(nC):
        a = np.where(sidx < S - 2 * (T - c * K + 1), 0.0, a)
        a, de = _rescale(a, TE)
        Ea += de
        ckpt.append(hi_round(a))
        Ea_c.append(Ea)
        for t in range(c * K, min(T, (c + 1) * K)):
            a = a_pre(a) * emit[t]
    zhat = a[S - 1] + (a[S - 2] if S > 1 else 0.0)
    Ea_fin = Ea
    if not (zhat > 0) or not np.isfinite(zhat):
        grad[:] = p32
        return np.inf, grad
    cost = -(np.log(zhat) + Ea_fin * np.log(2.0) - logs.sum())

    mz, ez = np.frexp(zhat)
    inv_zm = np.float32(MEANC / mz)
    b = np.zeros(S)
    b[S - 1] = np.ldexp(1.0, TE)
    Eb = -TE
    worst = 0.0
    for c in range(nC - 1, -1, -1):
        t0, t1 = c * K, min(T, (c + 1) * K)
        a = ckpt[c].copy()
        esc = Ea_c[c] + Eb - Ea_fin - ez
        acol = []
        pre0 = None
        for t in range(t0, t1):
            pre = a_pre(a)
            if t == t0:
                pre0 = hi_trunc(pre)
            a = pre * emit[t]
            acol.append(np.ldexp(hi_trunc(a), esc))
        for t in range(t1 - 1, t0 - 1, -1):
            tb = b_pre(b)
            b = tb * emit[t]
            prod = (acol[t - t0] * tb).astype(np.float32)
            acc = np.zeros(V, dtype=np.float32)
            for s_ in range(1, S, 2):
                acc[ext[s_]] += prod[s_]
            post = (acc * inv_zm).astype(np.float32)
            post[blank] = 0.0
            post[blank] = np.float32(1.0) - post.sum(dtype=np.float32)
            grad[t] = p32[t] - post
        q = float((pre0 * b).sum()) * 2.0 ** (Ea_c[c] + Eb - Ea_fin) / zhat * MEANC
        worst = max(worst, abs(q - 1.0))
        b = np.where(sidx > 2 * t0 + 1, 0.0, b)
        b, de = _rescale(b, TE)
        Eb += de
    if return_check:
        return float(cost), grad, worst
    return float(cost), grad


def ctc_batch(acts, flat_labels, act_lens, label_lens, blank=0, K=8):
    T_max, B, V = acts.shape
    costs = np.zeros(B)
    grads = np.zeros((T_max, B, V), dtype=np.float32)
    off = 0
    for b in range(B):
        T, L = int(act_lens[b]), int(label_lens[b])
        c, g = ctc_single(acts[:T, b], np.asarray(flat_labels)[off:off + L], blank, K)
        off += L
        costs[b] = c
        grads[:T, b] = g
    return costs, grads
