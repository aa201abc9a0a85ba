"""GPU greedy CTC decode -- drop-in for the reference's `GreedyDecoder.decode`.

Reference: /root/reference/codes/decoder.py:95-160 (used by codes/metrics.py:111 and test.py:78):
`decode(probs[B,T,V], sizes) -> (strings, offsets)` where `strings[b] == [text]` and `offsets[b] == [IntTensor]`.
There the argmax is one torch kernel but the collapse is a Python loop with one `.item()` sync per frame;
here both happen in one sm_100a kernel (csrc/ctc_decode.cuh) and only the compacted tokens come back.
CUDA only -- no CPU fallback.
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["GreedyDecoder", "greedy_decode_raw"]


def greedy_decode_raw(probs: torch.Tensor, sizes=None, blank: int = 0, want_offsets: bool = True):
    """probs: CUDA float32 B x T x V (any strides with unit last stride).  Returns CUDA int32 tensors
    (tokens[B,T], offsets[B,T] or None, counts[B]); the first counts[b] entries of a row are valid."""
    lib = _lib.load()
    if not probs.is_cuda:
        raise RuntimeError("aes_lac_2018_b200.GreedyDecoder is CUDA-only (B200-native); there is no CPU fallback")
    if probs.dim() != 3 or probs.dtype != torch.float32:
        raise TypeError("probs must be a float32 B x T x V tensor")
    p = probs.detach()
    if p.stride(2) != 1 and p.size(2) > 1:
        p = p.contiguous()
    B, T, V = p.shape
    dev = p.device
    with torch.cuda.device(dev):
        sz = None
        if sizes is not None:
            sz = torch.as_tensor(sizes).to(device=dev, dtype=torch.int32).reshape(-1).contiguous()
            if sz.numel() != B:
                raise ValueError("sizes must have one entry per utterance")
        tokens = torch.empty((B, T), dtype=torch.int32, device=dev)
        offsets = torch.empty((B, T), dtype=torch.int32, device=dev) if want_offsets else None
        counts = torch.empty(B, dtype=torch.int32, device=dev)
        st = lib.ctc_b200_greedy_decode(p.data_ptr(), p.stride(0), p.stride(1), sz.data_ptr() if sz is not None else None,
                                        B, T, V, int(blank), tokens.data_ptr(),
                                        offsets.data_ptr() if want_offsets else None, counts.data_ptr(),
                                        torch.cuda.current_stream(dev).cuda_stream)
        if st != _lib.CTC_STATUS_SUCCESS:
            raise RuntimeError("ctc_b200_greedy_decode: " + _lib.status_string(lib, st))
    return tokens, offsets, counts


class GreedyDecoder:
    """`GreedyDecoder(labels, blank_index=0).decode(probs, sizes)` with the reference's return structure.
    `labels` is the alphabet in index order (a string / list such as data/labels.en.json) or any object with an
    `inverse_transform(list_of_indices)` method (the reference's OrderedLabelEncoder)."""

    def __init__(self, label_encoder, blank_index: int = 0):
        if isinstance(label_encoder, str):
            label_encoder = list(label_encoder)
        self.label_encoder = label_encoder
        self.blank_index = blank_index

    def _to_string(self, ids):
        if not ids:
            return ""
        if hasattr(self.label_encoder, "inverse_transform"):
            return "".join(self.label_encoder.inverse_transform(ids))
        return "".join(self.label_encoder[i] for i in ids)

    def decode(self, probs, sizes=None):
        tokens, offsets, counts = greedy_decode_raw(probs, sizes, self.blank_index, want_offsets=True)
        counts_h = counts.cpu()
        n_max = int(counts_h.max()) if counts_h.numel() else 0
        tok_h = tokens[:, :n_max].cpu()                      # one small D2H instead of T syncs per utterance
        off_h = offsets[:, :n_max].cpu()
        strings, offs = [], []
        for b in range(tok_h.shape[0]):
            n = int(counts_h[b])
            strings.append([self._to_string(tok_h[b, :n].tolist())])
            offs.append([off_h[b, :n].clone().to(torch.int32)])
        return strings, offs
