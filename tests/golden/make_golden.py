"""Generates tests/golden/torch_f64_cases.npz and (with --check) re-verifies the transcribed
warp-ctc known-answer vectors in warpctc_known_answers.json.

The reference's CTC (warpctc_pytorch) cannot be imported offline, so the committed fixtures come
from an INDEPENDENT implementation: torch.nn.functional.ctc_loss in float64 on
log_softmax(acts), zero_infinity=True (= warp-ctc's CPU convention for infeasible utterances:
cost 0, gradient 0 -- SURVEY.md 8c).  Run from the repo root:  python tests/golden/make_golden.py [--check]
Neither this script nor the tests read /root/reference.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))


def torch_ctc(acts, labels, act_lens, label_lens, blank=0):
    a = torch.tensor(np.asarray(acts, dtype=np.float64), requires_grad=True)
    lab = torch.tensor(np.asarray(labels, dtype=np.int64))
    loss = F.ctc_loss(F.log_softmax(a, -1), lab, torch.tensor(np.asarray(act_lens, dtype=np.int64)),
                      torch.tensor(np.asarray(label_lens, dtype=np.int64)), blank=blank,
                      reduction="none", zero_infinity=True)
    loss.sum().backward()
    return loss.detach().numpy(), a.grad.numpy()


def make_case(rng, T, B, V, label_lens, act_lens, sigma=1.0, blank=0, force_repeats=(), peaky=False):
    acts = (rng.standard_normal((T, B, V)) * sigma).astype(np.float32)
    label_lens = np.asarray(label_lens, dtype=np.int32)
    act_lens = np.asarray(act_lens, dtype=np.int32)
    symbols = np.array([k for k in range(V) if k != blank])
    labels = rng.choice(symbols, int(label_lens.sum())).astype(np.int32)
    for i in force_repeats:
        labels[i] = labels[i - 1]
    if peaky:  # mimic a trained model: blank-dominant frames, label spikes
        acts[..., blank] += 6.0 * (rng.random((T, B)) < 0.7)
    costs, grads = torch_ctc(acts, labels, act_lens, label_lens, blank)
    return dict(acts=acts, labels=labels, act_lens=act_lens, label_lens=label_lens,
                blank=np.int32(blank), costs=costs, grads=grads)


def main():
    rng = np.random.default_rng(20181017)
    cases = {
        # BASELINE config 1: the reference's own CPU-runnable call shape
        "c1_b4_t200_v29": make_case(rng, 200, 4, 29, [10, 50, 33, 21], [200] * 4),
        # PT-BR alphabet, ragged lengths, L=0 mixed in, padded frames
        "v43_ragged": make_case(rng, 90, 6, 43, [0, 12, 30, 1, 25, 7], [90, 61, 90, 5, 77, 33], sigma=2.0),
        # repeats: all-same labels at T = 2L-2 (infeasible), 2L-1 (exactly feasible), 2L
        "repeats_boundary": None,
        # T shorter than L; T=1 with L in {0,1}; sum L = 1
        "tiny": None,
        "peaky_v29": make_case(rng, 120, 3, 29, [20, 35, 8], [120, 110, 64], peaky=True),
        "blank_last_v6": make_case(rng, 40, 2, 6, [9, 4], [40, 22], blank=5),
    }
    # all-same labels, L=6: T = 10 (infeasible), 11 (exactly feasible), 12
    T, B, V, L = 12, 3, 5, 6
    acts = rng.standard_normal((T, B, V)).astype(np.float32)
    labels = np.full(3 * L, 3, dtype=np.int32)
    al, ll = np.array([10, 11, 12], np.int32), np.array([L, L, L], np.int32)
    c, g = torch_ctc(acts, labels, al, ll)
    cases["repeats_boundary"] = dict(acts=acts, labels=labels, act_lens=al, label_lens=ll,
                                     blank=np.int32(0), costs=c, grads=g)
    T, B, V = 4, 4, 7
    acts = rng.standard_normal((T, B, V)).astype(np.float32)
    labels = np.array([2, 3, 4, 5, 6, 1], dtype=np.int32)           # utt0 L=5 > T=4; utt1 L=0,T=1; utt2 L=1,T=1
    al, ll = np.array([4, 1, 1, 3], np.int32), np.array([5, 0, 1, 0], np.int32)
    c, g = torch_ctc(acts, labels, al, ll)
    cases["tiny"] = dict(acts=acts, labels=labels, act_lens=al, label_lens=ll, blank=np.int32(0), costs=c, grads=g)

    flat = {}
    for name, d in cases.items():
        for k, v in d.items():
            flat[f"{name}/{k}"] = v
    out = os.path.join(HERE, "torch_f64_cases.npz")
    np.savez_compressed(out, **flat)
    print("wrote", out, os.path.getsize(out), "bytes;", len(cases), "cases")

    if "--check" in sys.argv:
        ka = json.load(open(os.path.join(HERE, "warpctc_known_answers.json")))
        for case in ka["cases"]:
            if case.get("acts_are_log_of_probs"):
                acts = np.log(np.asarray(case["probs_tbv"], dtype=np.float64))
            else:
                acts = np.asarray(case["acts_tbv"], dtype=np.float64) * case.get("scale", 1.0)
            c, g = torch_ctc(acts, case["labels"], case["act_lens"], case["label_lens"], case["blank"])
            msg = [case["name"], "costs", c]
            if "expected_total_cost" in case:
                assert abs(c.sum() - case["expected_total_cost"]) <= case["cost_tol"] * max(1, abs(c.sum())), msg
            if "expected_costs" in case:
                assert np.allclose(c, case["expected_costs"], atol=case["cost_tol"] * max(1, abs(c).max())), msg
            if "expected_grads_tbv" in case:
                assert np.allclose(g, case["expected_grads_tbv"], atol=case["grad_tol"]), msg
            if "expected_grad_t0_b0" in case:
                assert np.allclose(g[0, 0], case["expected_grad_t0_b0"], atol=case["grad_tol"]), msg
            print("verified", *msg)


if __name__ == "__main__":
    main()
