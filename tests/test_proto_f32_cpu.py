"""CPU check of the numpy model of the fp32 kernel's arithmetic (tests/proto_f32_bfp.py mirrors csrc/ctc_warp32.cuh:
per-lane block exponents, p~ domain, two-ended mass check) against the float64 oracle: the scheme holds the north_star
tolerances on benign inputs, and whatever it cannot hold it flags."""
import numpy as np

from oracle import ctc_f64
from tests import proto_f32_bfp as proto

GRAD_ATOL, LOSS_RTOL = 1e-5, 1e-4


def _run(acts, labels):
    c0, g0 = ctc_f64.ctc_single(acts, labels, 0)
    with np.errstate(all="ignore"):
        c1, g1, _, flags = proto.ctc_single(acts, labels, 0, K=8, return_check="flags")
    err = float(np.abs(g1 - g0).max()) if np.isfinite(g1).all() else np.inf
    rel = abs(c1 - c0) / max(1.0, abs(c0)) if np.isfinite(c1) else np.inf
    return err, rel, any(flags.values())


def test_benign_inputs_hold_the_tolerances_unflagged():
    rng = np.random.default_rng(7)
    for kind, T, L, V in (("randn", 120, 30, 29), ("randn", 90, 10, 43), ("peaky6", 150, 40, 29)):
        acts, labels = proto._problem(rng, T, L, V, kind, 1.0)
        err, rel, flagged = _run(acts, labels)
        assert err <= GRAD_ATOL and rel <= LOSS_RTOL and not flagged, (kind, T, L, V, err, rel, flagged)


def test_what_leaves_the_fp32_range_is_flagged():
    """Wide logits on tight alignments: the band of live states sweeps through lane frames that are fixed for a chunk.
    Every case is either within tolerance or flagged by the two-ended mass check -- never silently wrong."""
    wrong = 0
    for seed in range(12):
        rng = np.random.default_rng(100 + seed)
        L = int(rng.choice([30, 60, 90]))
        T = L + int(rng.integers(0, 12))
        labels = rng.integers(1, 17, size=L)
        for i in range(1, L):
            if labels[i] == labels[i - 1]:
                labels[i] = 1 + (labels[i] % 16)
        acts = rng.normal(0, float(rng.choice([4.0, 6.0])), size=(T, 17)).astype(np.float32)
        err, rel, flagged = _run(acts, labels)
        assert flagged or (err <= GRAD_ATOL and rel <= LOSS_RTOL), (seed, T, L, err, rel)
        wrong += err > GRAD_ATOL
    assert wrong > 0, "expected some of these cases to leave the fp32 range (otherwise the test checks nothing)"
