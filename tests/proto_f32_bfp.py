"""numpy model of a PER-LANE BLOCK-FLOATING-POINT fp32 recursion -- the prototype the round-1 review asked for
("prototype in tests/proto_* a per-lane block-floating-point fp32 recursion; keep it only if it holds 1e-5").

NOT an oracle and NOT a product path.  It answers one question on the CPU before any GPU minute is spent: can the
T-serial alpha / beta recursion of csrc/ctc_warp.cuh run in fp32 (full-rate pipe, 4-cycle latency, half the registers)
instead of fp64 and still meet the north_star tolerances (1e-5 abs on the gradient, 1e-4 rel on the loss) against
oracle/ctc_f64.py?  SURVEY.md Appendix D scheme F (ONE scale per column, fp32) fails on dynamic range: within a column
alpha^ spans more than e^87.  Here every lane (NS consecutive states of the blank-extended transcript) carries its own
binary exponent:

  * p~ domain: p~[t,k] = exp(a[t,k] - max_k a[t,:])  (fp32, ex2.approx-sized noise), so a column can only shrink
    apart from the <= 3-term sums (growth <= 3^K per chunk of K frames): no overflow by construction;
  * lane l holds float32 x[l][0..NS) and an int e_l: true value = x * 2^e_l.  Once per chunk the lane picks
    e_l from the largest magnitude among itself and the lanes mass can arrive from within the chunk
    (ceil(2K/NS) lanes below for alpha, above for beta), made 64-Lipschitz so that the neighbour factor
    2^(e_{l-1} - e_l) never overflows;
  * neighbour values cross a lane boundary multiplied by that exact power of two;
  * checkpoint = the fp32 column + 32 exponents per chunk; recompute inside the chunk (identical ops);
  * posterior_t(k) = sum_{s in pos(k)} alpha_scaled_t(s) * tb_t(s) / mz with the chunk's checkpoint column scaled
    per lane by 2^(ea_l + eb_l - ez) (clamped to 2^110: a clamped entry meets a tb below 2^-109);
  * loss: log Z^ + exponent of the lane holding the end states - sum_t log s_t (fp64 running product of fp32 sums).

Run `python tests/proto_f32_bfp.py` for the table that DESIGN.md section 4e quotes.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

F = np.float32
TA = 100           # target exponent of a lane group's max, forward sweep
TB = 40            # ... backward sweep (alpha_scaled * tb ~ posterior: the two ranges are reciprocal)
LIP = 64
NLANES = 32


def _exp2i(n):
    """2^n as fp32 for integer arrays n (flushes below 2^-149, saturates above 2^127 -- callers keep n in range)."""
    return np.ldexp(np.float64(1.0), np.clip(n, -200, 127)).astype(F)


def _scale(x, n):
    """x * 2^n in fp32, n integer array broadcast over the lane's states; two exact factors so |n| <= 252 is safe."""
    n = np.clip(n, -252, 252)
    h = n // 2
    return ((x * _exp2i(h)[:, None]).astype(F) * _exp2i(n - h)[:, None]).astype(F)


def _lane_exps(x, e, target, reach, upward):
    """New per-lane exponents: group max over the lanes mass can come from, 64-Lipschitz in the flow direction."""
    m = np.abs(x).max(axis=1)
    with np.errstate(divide="ignore"):
        ex = np.where(m > 0, np.floor(np.log2(np.maximum(m.astype(np.float64), 1e-300))), -10 ** 6).astype(np.int64)
    ab = np.where(m > 0, e + ex, -10 ** 6)
    M = ab.copy()
    for d in range(1, reach + 1):
        sh = np.full_like(ab, -10 ** 6)
        if upward:
            sh[d:] = ab[:-d]
        else:
            sh[:-d] = ab[d:]
        M = np.maximum(M, sh)
    en = M - target
    for d in (1, 2, 4):
        sh = np.full_like(en, -10 ** 7)
        if upward:
            sh[d:] = en[:-d]
        else:
            sh[:-d] = en[d:]
        en = np.maximum(en, sh - LIP * d)
    dead = M <= -10 ** 5                                     # nothing within reach: keep the old exponent
    en = np.where(dead & (en <= -10 ** 5), e, en)
    return en


def ctc_single(acts_tv, labels, blank=0, K=8, NS=None, seed=0, return_check=False, renorm=True):
    acts_tv = np.asarray(acts_tv, dtype=F)
    labels = np.asarray(labels, dtype=np.int64).reshape(-1)
    T, V = acts_tv.shape
    L = len(labels)
    S = 2 * L + 1
    grad = np.zeros((T, V), dtype=F)
    rep = int((labels[1:] == labels[:-1]).sum()) if L > 1 else 0
    if T == 0 or L + rep > T:
        return 0.0, grad
    if NS is None:
        NS = max(2, 2 * ((S + 1 + 63) // 64))
    SP = NLANES * NS
    assert SP >= S + 1
    reach = -(-2 * K // NS)
    ext = np.full(SP, blank, dtype=np.int64)
    ext[1:S:2] = labels
    is_lab = np.zeros(SP, dtype=bool)
    is_lab[1:S:2] = True
    valid = np.arange(SP) < S
    skip = np.zeros(SP, dtype=bool)
    skip[2:] = is_lab[2:] & (ext[2:] != ext[:-2])
    skip_f = np.zeros(SP, dtype=bool)
    skip_f[:-2] = skip[2:]

    mx = acts_tv.max(axis=1, keepdims=True)
    rng = np.random.default_rng(seed)
    pt = np.exp((acts_tv - mx).astype(np.float64)) * (1 + (rng.random(acts_tv.shape) - 0.5) * 2.0 ** -22)
    pt = pt.astype(F)
    ssum = pt.sum(axis=1, dtype=F)
    p32 = (pt * (F(1) / ssum)[:, None]).astype(F)
    logs = np.log(ssum.astype(np.float64))
    emit = pt[:, ext]
    emit[:, ~valid] = 0

    lane_of = np.arange(SP) // NS

    def flat(x):
        return x.reshape(SP)

    def a_step(x, e, em):
        """one alpha step on the [32, NS] block representation"""
        xf = flat(x)
        up = np.zeros((NLANES, 2), dtype=F)                  # x[l-1][NS-1], x[l-1][NS-2] in lane l's frame
        f = _exp2i(np.concatenate(([-(10 ** 3)], e[:-1] - e[1:])))
        up[1:, 0] = x[:-1, NS - 1]
        up[1:, 1] = x[:-1, NS - 2] if NS >= 2 else 0
        up = (up * f[:, None]).astype(F)
        p1 = np.empty(SP, dtype=F)
        p2 = np.empty(SP, dtype=F)
        p1[1:] = xf[:-1]
        p1[0] = 0
        p2[2:] = xf[:-2]
        p2[:2] = 0
        first = np.arange(SP) % NS == 0
        second = np.arange(SP) % NS == 1
        p1[first] = up[:, 0]
        p2[first] = up[:, 1]
        p2[second] = up[:, 0]
        s1 = (xf + p1).astype(F)
        s1 = np.where(skip, (s1.astype(np.float64) + p2.astype(np.float64)).astype(F), s1)   # fma(msk, p2, s1)
        return (s1 * em).astype(F).reshape(NLANES, NS)

    def b_pre(x, e):
        xf = flat(x)
        dn = np.zeros((NLANES, 2), dtype=F)
        f = _exp2i(np.concatenate((e[1:] - e[:-1], [-(10 ** 3)])))
        dn[:-1, 0] = x[1:, 0]
        dn[:-1, 1] = x[1:, 1] if NS >= 2 else 0
        dn = (dn * f[:, None]).astype(F)
        n1 = np.empty(SP, dtype=F)
        n2 = np.empty(SP, dtype=F)
        n1[:-1] = xf[1:]
        n1[-1] = 0
        n2[:-2] = xf[2:]
        n2[-2:] = 0
        last = np.arange(SP) % NS == NS - 1
        last2 = np.arange(SP) % NS == NS - 2
        n1[last] = dn[:, 0]
        n2[last] = dn[:, 1]
        n2[last2] = dn[:, 0]
        s1 = (xf + n1).astype(F)
        s1 = np.where(skip_f, (s1.astype(np.float64) + n2.astype(np.float64)).astype(F), s1)
        return s1.reshape(NLANES, NS)

    sidx = np.arange(SP).reshape(NLANES, NS)
    nC = (T + K - 1) // K
    a = np.zeros((NLANES, NS), dtype=F)
    a[0, 0] = F(2.0 ** TA)
    ea = np.full(NLANES, -TA, dtype=np.int64)
    ck, ck_e = [], []
    for c in range(nC):
        a = np.where(sidx < S - 2 * (T - c * K + 1), F(0), a)
        en = _lane_exps(a, ea, TA, reach, upward=True)
        a = _scale(a, ea - en)
        ea = en
        ck.append(a.copy())
        ck_e.append(ea.copy())
        for t in range(c * K, min(T, (c + 1) * K)):
            a = a_step(a, ea, emit[t])
    e_ref = int(ea[lane_of[S - 1]])
    zloc = float(flat(a)[S - 1])
    if S > 1:
        zloc += float(flat(a)[S - 2]) * 2.0 ** float(np.clip(int(ea[lane_of[S - 2]]) - e_ref, -1000, 1000))
    if not (zloc > 0) or not np.isfinite(zloc):
        grad[:] = p32
        if return_check == "flags":
            return np.inf, grad, np.inf, dict(drift=True, jump=True, last=True)
        return (np.inf, grad, np.inf) if return_check else (np.inf, grad)
    cost = -(np.log(zloc) + e_ref * np.log(2.0) - logs.sum())
    mz, ezl = np.frexp(zloc)
    ez = int(ezl) + e_ref
    inv_mz = F(1.0 / mz)

    b = np.zeros((NLANES, NS), dtype=F)
    flat(b)[S - 1] = F(2.0 ** TB)
    eb = np.full(NLANES, -TB, dtype=np.int64)
    worst = 0.0
    flags = dict(drift=False, jump=False, last=False)
    q_prev = None
    for c in range(nC - 1, -1, -1):
        t0, t1 = c * K, min(T, (c + 1) * K)
        esc = ck_e[c] + eb - ez
        # the scaled column a_sc = alpha^ * 2^(eb_l - ez) lives in the per-lane frame e'_l = ez - eb_l, so the neighbour
        # factor of the recompute is 2^(eb_l - eb_{l-1}) (<= 2^64 by the Lipschitz rule).  No clamp: an overflow ends as
        # inf / NaN in the mass check of the chunk's last frame.
        with np.errstate(over="ignore", invalid="ignore"):
            a = _scale(ck[c], esc)
            eframe = ez - eb
            acol = []
            for t in range(t0, t1):
                a = a_step(a, eframe, emit[t])
                acol.append(a)
            prods = []
            for t in range(t1 - 1, t0 - 1, -1):
                tb = b_pre(b, eb)
                b = (flat(tb) * emit[t]).astype(F).reshape(NLANES, NS)
                prods.append((t, (flat(acol[t - t0]) * flat(tb)).astype(F)))
        # -- mass checks: sum_s alpha_sc(t, s) * tb(t, s) = Z^ at the chunk's FIRST frame (against Z^: the drift of the two
        #    product chains, divided out of the chunk's posteriors) and at its LAST frame (against the first).  All terms
        #    are positive and both recursions are linear, so
        #      * relevant mass lost (flushed, denormal) by the beta recursion at any frame of the chunk never arrives at
        #        the first frame: a deficit there, i.e. a jump against the previous chunk's value;
        #      * relevant mass lost (flushed, overflowed) by the recomputed alpha recursion at any frame of the chunk
        #        never arrives at the last frame: a deficit there;
        #      * mass lost by the forward sweep makes Z^ itself too small: a jump between the chunks around the loss.
        #    The frames in between need no check of their own (and no storage of their blank states). --
        q = float(prods[-1][1][valid].astype(np.float64).sum()) / mz
        q_last = float(prods[0][1][valid].astype(np.float64).sum()) / mz
        worst = max(worst, abs(q - 1.0) if np.isfinite(q) else np.inf)
        if not np.isfinite(q) or abs(q - 1.0) > 1e-4:
            flags["drift"] = True
        ref = q_prev if q_prev is not None else 1.0
        if not np.isfinite(q) or abs(q / ref - 1.0) > 4e-6:
            flags["jump"] = True
        if not np.isfinite(q_last) or abs(q_last / q - 1.0) > 4e-6:
            flags["last"] = True
        q_prev = q
        scale = F(inv_mz / F(q)) if (renorm and np.isfinite(q) and q > 0) else inv_mz
        for t, prod in prods:
            acc = np.zeros(V, dtype=F)
            for s_ in range(1, S, 2):
                acc[ext[s_]] += prod[s_]
            post = (acc * scale).astype(F)
            post[blank] = 0
            post[blank] = F(1) - post.sum(dtype=F)
            grad[t] = p32[t] - post
        b = np.where(sidx > 2 * t0 + 1, F(0), b)
        en = _lane_exps(b, eb, TB, reach, upward=False)
        b = _scale(b, eb - en)
        eb = en
    if return_check == "flags":
        return float(cost), grad, worst, flags
    if return_check:
        return float(cost), grad, worst
    return float(cost), grad


# ---------------------------------------------------------------------------------------------------------------
def _problem(rng, T, L, V, kind, sigma=1.0):
    labels = rng.integers(1, V, size=L)
    if kind == "randn":
        acts = rng.normal(0, sigma, size=(T, V))
    else:
        # trained-model-like: a monotone alignment, blank +m on ~70 % of frames, the aligned label +m otherwise
        m = {"peaky6": 6.0, "peaky12": 12.0, "wrong12": 12.0}[kind]
        acts = rng.normal(0, 1.0, size=(T, V))
        pos = np.sort(rng.choice(T, size=L, replace=False)) if L <= T else np.arange(L)
        tgt = labels if kind != "wrong12" else np.concatenate((labels[: L // 2], rng.integers(1, V, size=L - L // 2)))
        for t in range(T):
            acts[t, 0] += m
        for j, t in enumerate(pos):
            acts[t, 0] -= m
            acts[t, tgt[j]] += m
    return acts.astype(F), labels


def main():
    from oracle.ctc_f64 import ctc_single as oracle_single
    rng = np.random.default_rng(2018)
    rows = []
    shapes = [(200, 50, 29), (750, 60, 29), (750, 120, 29), (750, 200, 29), (750, 200, 43), (1500, 400, 29)]
    if "--long" in sys.argv:
        shapes.append((3000, 600, 29))
    kinds = (("randn", 1.0), ("randn", 3.0), ("peaky6", 1), ("peaky12", 1), ("wrong12", 1))
    if "--peaky" in sys.argv:
        kinds = (("peaky6", 1), ("peaky12", 1))
    for kind, sigma in kinds:
        for T, L, V in shapes:
            acts, labels = _problem(rng, T, L, V, kind, sigma)
            c0, g0 = oracle_single(acts, labels, 0)
            c1, g1, w = ctc_single(acts, labels, 0, K=8, return_check=True, renorm=True)
            _, g2, _ = ctc_single(acts, labels, 0, K=8, return_check=True, renorm=False)
            rows.append((kind, sigma, T, L, V, abs(c1 - c0) / max(1.0, abs(c0)), float(np.abs(g1 - g0).max()),
                         float(np.abs(g2 - g0).max()), w))
            print("%-8s sigma=%g T=%4d L=%3d V=%2d  cost %.3f  rel %.2e  max|dgrad| %.2e (no renorm %.2e)  check %.2e" %
                  (kind, sigma, T, L, V, c0, rows[-1][5], rows[-1][6], rows[-1][7], w), flush=True)
    return rows


if __name__ == "__main__":
    main()
