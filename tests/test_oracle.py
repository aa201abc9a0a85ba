"""Pins the CPU oracles (oracle/) against the committed golden vectors.  CPU only."""
import numpy as np
import pytest

from oracle import ctc_f64, warpctc_cpu
from tests.helpers import load_known_answers, load_torch_f64_cases, synth_problem

KNOWN = load_known_answers()
TORCH_CASES = load_torch_f64_cases()


def _check_known(case, costs, grads, extra_tol=0.0):
    tol = case["cost_tol"] + extra_tol
    if "expected_total_cost" in case:
        assert abs(costs.sum() - case["expected_total_cost"]) <= tol * max(1.0, abs(case["expected_total_cost"]))
    if "expected_costs" in case:
        np.testing.assert_allclose(costs, case["expected_costs"], atol=tol * max(1.0, max(case["expected_costs"])))
    if "expected_grads_tbv" in case:
        np.testing.assert_allclose(grads, case["expected_grads_tbv"], atol=case["grad_tol"])
    if "expected_grad_t0_b0" in case:
        np.testing.assert_allclose(grads[0, 0], case["expected_grad_t0_b0"], atol=case["grad_tol"])


@pytest.mark.parametrize("case", KNOWN, ids=[c["name"] for c in KNOWN])
def test_f64_oracle_known_answers(case):
    costs, grads = ctc_f64.ctc_batch(case["acts"], case["labels"], case["act_lens"], case["label_lens"], case["blank"])
    _check_known(case, costs, grads)


@pytest.mark.parametrize("case", KNOWN, ids=[c["name"] for c in KNOWN])
def test_fp32_restatement_known_answers(case):
    costs, grads = warpctc_cpu.ctc_batch(case["acts"], case["labels"], case["act_lens"], case["label_lens"], case["blank"])
    # probabilities are fp32 and log() is taken at use (warp-ctc CPU): exp(-100) is an fp32 denormal
    # with ~5 significant bits, so the x200 case is only good to 2e-4 relative in this arithmetic.
    extra = 5e-4 if "x200" in case["name"] else 2e-6
    _check_known(case, costs.astype(np.float64), grads.astype(np.float64), extra_tol=extra)


@pytest.mark.parametrize("name", sorted(TORCH_CASES))
def test_f64_oracle_matches_torch_float64(name):
    c = TORCH_CASES[name]
    costs, grads = ctc_f64.ctc_batch(c["acts"], c["labels"], c["act_lens"], c["label_lens"], int(c["blank"]))
    np.testing.assert_allclose(costs, c["costs"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(grads, c["grads"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", sorted(TORCH_CASES))
def test_fp32_restatement_matches_torch_float64(name):
    """fp32 warp-ctc arithmetic: loss to 1e-5 rel; gradient only to ~1e-3 (scheme A)."""
    c = TORCH_CASES[name]
    costs, grads = warpctc_cpu.ctc_batch(c["acts"], c["labels"], c["act_lens"], c["label_lens"], int(c["blank"]))
    np.testing.assert_allclose(costs, c["costs"], rtol=1e-5, atol=1e-5)
    assert np.abs(grads - c["grads"]).max() < 2e-3


def test_padded_frames_and_infeasible_rows_are_zero():
    c = TORCH_CASES["tiny"]
    for fn in (ctc_f64.ctc_batch, warpctc_cpu.ctc_batch):
        costs, grads = fn(c["acts"], c["labels"], c["act_lens"], c["label_lens"], 0)
        assert costs[0] == 0.0 and not grads[:, 0].any()          # L=5 > T=4: cost 0, grad 0
        for b, T in enumerate(c["act_lens"]):
            assert not grads[T:, b].any()


def test_gradient_rows_sum_to_zero_and_finite_difference():
    acts, labels, al, ll = synth_problem(7, 12, 2, 6, 2, 4)
    costs, grads = ctc_f64.ctc_batch(acts, labels, al, ll)
    np.testing.assert_allclose(grads.sum(-1), 0.0, atol=1e-12)
    eps = 1e-6
    rng = np.random.default_rng(0)
    a64 = acts.astype(np.float64)
    for _ in range(10):
        t, b, k = rng.integers(0, 12), rng.integers(0, 2), rng.integers(0, 6)
        ap, am = a64.copy(), a64.copy()
        ap[t, b, k] += eps
        am[t, b, k] -= eps
        fd = (ctc_f64.ctc_batch(ap, labels, al, ll)[0].sum() - ctc_f64.ctc_batch(am, labels, al, ll)[0].sum()) / (2 * eps)
        assert abs(fd - grads[t, b, k]) < 1e-6


def test_module_level_averaging_flags():
    acts, labels, al, ll = synth_problem(3, 20, 3, 5, 1, 5, tmin=12)
    tot, g = ctc_f64.ctc_loss_module(acts, labels, al, ll)
    tot_s, g_s = ctc_f64.ctc_loss_module(acts, labels, al, ll, size_average=True)
    tot_l, g_l = ctc_f64.ctc_loss_module(acts, labels, al, ll, size_average=True, length_average=True)
    assert np.isclose(tot_s, tot / 3) and np.allclose(g_s, g / 3)
    assert np.isclose(tot_l, tot / al.sum()) and np.allclose(g_l, g / al.sum())


def test_inf_case_no_nan():
    """upstream inf_test: one needed label column at -1e30 => cost +inf, no NaN in the gradient."""
    rng = np.random.default_rng(5)
    T, V, L = 50, 15, 10
    acts = rng.standard_normal((T, 1, V)).astype(np.float32)
    labels = rng.integers(1, V, L).astype(np.int32)
    labels[0] = 2
    acts[:, 0, 2] = -1e30
    costs, grads = ctc_f64.ctc_batch(acts, labels, [T], [L])
    assert np.isinf(costs[0]) and costs[0] > 0 and not np.isnan(grads).any()
