"""Randomised parity sweep: many small problems of random shape through every kernel ladder against the
float64 oracle (loss 1e-4 rel, gradient 1e-5 abs).  Shapes straddle the variant boundaries (SP = 64, 128, ...),
partial chunks (T mod 4/8/16), ragged lengths, L = 0, infeasible utterances, both alphabets and a far blank."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LOSS_RTOL, GRAD_ATOL = 1e-4, 1e-5


def _random_problem(rng):
    V = int(rng.choice([5, 29, 29, 43, 64]))
    T = int(rng.integers(1, 140))
    B = int(rng.integers(1, 7))
    blank = int(rng.choice([0, 0, V - 1, V // 2]))
    lmax = int(rng.choice([0, 3, 15, 31, 32, 63, 64, 70]))
    act_lens = rng.integers(max(1, T // 2), T + 1, B).astype(np.int32)
    act_lens[rng.integers(0, B)] = T
    label_lens = rng.integers(0, lmax + 1, B).astype(np.int32)
    symbols = np.array([k for k in range(V) if k != blank])
    labels = rng.choice(symbols, int(label_lens.sum())).astype(np.int32)
    if labels.size > 3 and rng.random() < 0.5:                         # sprinkle repeats
        idx = rng.integers(1, labels.size, labels.size // 3)
        labels[idx] = labels[idx - 1]
    sigma = float(rng.choice([0.5, 1.0, 3.0]))
    acts = (rng.standard_normal((T, B, V)) * sigma).astype(np.float32)
    return acts, labels, act_lens, label_lens, blank


@pytest.mark.parametrize("seed", range(24))
def test_random_shapes_all_ladders(seed):
    from aes_lac_2018_b200 import ctc_loss_raw
    from oracle import ctc_f64
    rng = np.random.default_rng(9000 + seed)
    acts, labels, al, ll, blank = _random_problem(rng)
    oc, og = ctc_f64.ctc_batch(acts, labels, al, ll, blank)
    a = torch.tensor(acts).cuda()
    args = [torch.tensor(x) for x in (labels, al, ll)]
    for mode, bidir in (("warp32", True), ("warp", True), ("throughput", True), ("throughput8", True), ("latency", True), ("latency", False)):
        costs, grads, status = ctc_loss_raw(a, *args, blank=blank, mode=mode, bidirectional=bidir)
        c = costs.numpy().astype(np.float64)
        g = grads.cpu().numpy().astype(np.float64)
        tag = f"seed {seed} mode {mode} bidir {bidir} shape {acts.shape} L {ll.tolist()} T {al.tolist()} blank {blank}"
        assert (np.abs(c - oc) <= LOSS_RTOL * np.maximum(1.0, np.abs(oc))).all(), tag
        assert np.abs(g - og).max() <= GRAD_ATOL, tag
        assert not (status.numpy() & 0xC).any(), tag


# Transcript lengths on both sides of every variant's capacity (states per CTA = NS * 32 * W): throughput ladders
# 64, 128, ..., 512, 1024, 2048, 4096 states; latency ladder 64, 128, 256, 512, 1024, 2048, 4096.
_EDGES = [31, 32, 63, 64, 95, 96, 127, 128, 159, 160, 191, 192, 223, 224, 255, 256, 511, 512, 1023, 1024]


@pytest.mark.parametrize("L", _EDGES)
def test_variant_capacity_edges(L):
    """One utterance per call, so the call runs exactly the variant that L selects; T is tight (about 1.25 L, not a
    multiple of any chunk length), a third of the labels are repeats, alphabets alternate between 29 and 64."""
    from aes_lac_2018_b200 import ctc_loss_raw
    from oracle import ctc_f64
    rng = np.random.default_rng(4000 + L)
    V = 29 if (L % 2) else 64
    labels = rng.integers(1, V, L).astype(np.int32)
    idx = rng.integers(1, L, L // 3)
    labels[idx] = labels[idx - 1]
    rep = int((labels[1:] == labels[:-1]).sum())
    T = L + rep + L // 4 + 37
    acts = rng.standard_normal((T, 1, V)).astype(np.float32)
    al, ll = np.array([T], np.int32), np.array([L], np.int32)
    oc, og = ctc_f64.ctc_batch(acts, labels, al, ll)
    a = torch.tensor(acts).cuda()
    args = [torch.tensor(x) for x in (labels, al, ll)]
    for mode, bidir in (("warp32", False), ("warp", False), ("throughput", False), ("throughput8", False), ("latency", True), ("latency", False)):
        costs, grads, status = ctc_loss_raw(a, *args, mode=mode, bidirectional=bidir)
        tag = f"L {L} T {T} V {V} mode {mode} bidir {bidir}"
        assert status.numpy()[0] == 0, tag + f" status {status.numpy()[0]}"      # in particular: no log-space detour
        assert abs(float(costs[0]) - oc[0]) <= LOSS_RTOL * max(1.0, abs(oc[0])), tag
        assert np.abs(grads.cpu().numpy().astype(np.float64) - og).max() <= GRAD_ATOL, tag
