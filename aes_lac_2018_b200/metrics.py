"""WER / CER evaluation metrics on the GPU -- mirrors of the reference's `metrics.WER` / `metrics.CER`.

Reference: /root/reference/codes/metrics.py:66-162 (`EditDistance.update`: unflatten targets, `decoder.decode`,
`convert_to_strings`, one `decoder.wer`/`decoder.cer` call per utterance on host strings).  Same constructor
arguments, `reset` / `update` / `compute` / `val` / `den`, same numbers; but `update` runs decode and both
distances as three kernels over the device-resident logits and reads back 2 x B integers.  ignite is not a
dependency: the classes are plain objects with the Metric protocol ignite calls (`reset`, `update`, `compute`).
CUDA only.
"""
from __future__ import annotations

import torch

from .decoder import GreedyDecoder, _drop_blank, edit_distance_raw, greedy_decode_raw

__all__ = ["EditDistance", "WER", "CER"]


class EditDistance:
    """`update(output)` takes `(out[B,T,V] CUDA, targets[sum L], out_sizes[B], target_sizes[B])` like the reference.
    kind: 'wer' or 'cer'.  Non-stateful (the reference's default, train.py:184-185): mean over utterances of
    distance / normaliser (no division when the normaliser is 0, metrics.py:146-148).  Stateful: total distance
    over total normaliser (the reference's stateful branch adds the *normalised* value to its denominator,
    metrics.py:124 -- a latent bug that is not reproduced)."""

    def __init__(self, decoder: GreedyDecoder, kind: str, output_transform=lambda x: x, stateful: bool = False):
        if kind not in ("wer", "cer"):
            raise ValueError("kind must be 'wer' or 'cer'")
        self._decoder = decoder
        self._kind = kind
        self._output_transform = output_transform
        self._stateful = stateful
        self.reset()

    def reset(self):
        self._total_edit_distance = 0
        self._num_examples = 0

    def update(self, output):
        out, targets, out_sizes, target_sizes = self._output_transform(output)
        dec = self._decoder
        tokens, _, counts = greedy_decode_raw(out, out_sizes, dec.blank_index, want_offsets=False)
        refs, lens = _drop_blank(targets, target_sizes, dec.blank_index)
        dist, norm = edit_distance_raw(tokens, counts, refs, lens, dec.space_index, self._kind)
        res = torch.stack([dist, norm]).cpu().to(torch.float64)
        d, n = res[0], res[1]
        if not self._stateful:
            self._total_edit_distance += float(torch.where(n > 0, d / n.clamp(min=1), d).sum())
            self._num_examples += out.shape[0]
        else:
            self._total_edit_distance += float(d.sum())
            self._num_examples += float(n.sum())

    def compute(self):
        if self._num_examples == 0:
            raise RuntimeError("WER must have at least one example before it can be computed")
        return (self._total_edit_distance / self._num_examples) * 100

    @property
    def val(self):
        return self._total_edit_distance

    @property
    def den(self):
        return self._num_examples


class WER(EditDistance):
    def __init__(self, decoder, output_transform=lambda x: x, stateful=False):
        super().__init__(decoder, "wer", output_transform, stateful)


class CER(EditDistance):
    def __init__(self, decoder, output_transform=lambda x: x, stateful=False):
        super().__init__(decoder, "cer", output_transform, stateful)
