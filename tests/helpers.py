"""Shared test helpers: golden-fixture loading and seeded synthetic CTC problems."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_known_answers():
    with open(os.path.join(GOLDEN, "warpctc_known_answers.json")) as f:
        cases = json.load(f)["cases"]
    for c in cases:
        if c.get("acts_are_log_of_probs"):
            c["acts"] = np.log(np.asarray(c["probs_tbv"], dtype=np.float64))
        else:
            c["acts"] = np.asarray(c["acts_tbv"], dtype=np.float64) * c.get("scale", 1.0)
    return cases


def load_torch_f64_cases():
    z = np.load(os.path.join(GOLDEN, "torch_f64_cases.npz"))
    cases = {}
    for key in z.files:
        name, field = key.split("/")
        cases.setdefault(name, {})[field] = z[key]
    return cases


def synth_problem(seed, T, B, V, lmin, lmax, tmin=None, sigma=1.0, blank=0, peaky=False):
    """Deterministic synthetic DeepSpeech2-shaped CTC problem."""
    rng = np.random.default_rng(seed)
    acts = (rng.standard_normal((T, B, V)) * sigma).astype(np.float32)
    label_lens = rng.integers(lmin, lmax + 1, B).astype(np.int32)
    act_lens = (np.full(B, T) if tmin is None else rng.integers(tmin, T + 1, B)).astype(np.int32)
    if tmin is not None:
        act_lens[rng.integers(0, B)] = T
    symbols = np.array([k for k in range(V) if k != blank])
    labels = rng.choice(symbols, int(label_lens.sum())).astype(np.int32)
    if peaky:
        acts[..., blank] += 6.0 * (rng.random((T, B)) < 0.7)
    return acts, labels, act_lens, label_lens
