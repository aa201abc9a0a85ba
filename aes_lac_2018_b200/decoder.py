"""GPU greedy CTC decode and WER/CER scoring -- drop-in for the reference's `GreedyDecoder`.

Reference: /root/reference/codes/decoder.py:95-160 (used by codes/metrics.py:111 and test.py:78):
`decode(probs[B,T,V], sizes) -> (strings, offsets)` where `strings[b] == [text]` and `offsets[b] == [IntTensor]`.
There the argmax is one torch kernel but the collapse is a Python loop with one `.item()` sync per frame;
here both happen in one sm_100a kernel (csrc/ctc_decode.cuh) and only the compacted tokens come back.
Scoring (`wer`, `cer`, decoder.py:49-78; python-Levenshtein on host strings there) is one more kernel
(csrc/ctc_editdist.cuh) over the device-resident token rows; `error_counts` chains decode -> WER -> CER and brings
back four integers per utterance.
CUDA only -- no CPU fallback.
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["GreedyDecoder", "OrderedAlphabet", "greedy_decode_raw", "edit_distance_raw"]


class OrderedAlphabet:
    """Index <-> symbol map in first-appearance order: the `transform` / `inverse_transform` pair of the
    reference's OrderedLabelEncoder (/root/reference/codes/preprocessing.py:10-80), which `Decoder.__init__`
    wraps list / set / str alphabets in (codes/decoder.py:39-46).  Host-side bookkeeping only."""

    def __init__(self, symbols):
        self.classes_ = []
        self.map_classes_ = {}
        for c in symbols:
            if c not in self.map_classes_:
                self.map_classes_[c] = len(self.classes_)
                self.classes_.append(c)

    def transform(self, y):
        try:
            return [self.map_classes_[c] for c in y]
        except KeyError as e:
            raise ValueError(f"y contains previously unseen labels: {e.args[0]!r}") from None

    def inverse_transform(self, y):
        out = []
        for i in y:
            i = int(i)
            if i < 0 or i >= len(self.classes_):
                raise ValueError(f"y contains previously unseen labels: {i}")
            out.append(self.classes_[i])
        return out

    def __len__(self):
        return len(self.classes_)

_EDIT_MODES = {"tokens": 0, "cer": 1, "wer": 2}


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


def edit_distance_raw(hyp_tokens: torch.Tensor, hyp_counts: torch.Tensor, refs, ref_lens, space: int = -1,
                      mode: str = "tokens"):
    """Levenshtein distance of every decoded row against its reference, on the GPU.
    hyp_tokens: CUDA int32 [B, W] (row b holds hyp_counts[b] tokens -- what greedy_decode_raw returns);
    refs: flat int tensor of the concatenated references (CPU or CUDA); ref_lens: [B] (CPU preferred: its maximum
    sizes the kernel).  mode: 'tokens' | 'cer' (drop `space` tokens first) | 'wer' (words between `space` runs).
    Returns CUDA int32 tensors (distances[B], normalisers[B]): normaliser = reference length (tokens, cer) or
    reference word count (wer), the denominators of codes/metrics.py:145-160."""
    lib = _lib.load()
    if not hyp_tokens.is_cuda:
        raise RuntimeError("aes_lac_2018_b200 edit distance is CUDA-only (B200-native); there is no CPU fallback")
    if mode not in _EDIT_MODES:
        raise ValueError("mode must be 'tokens', 'cer' or 'wer'")
    if hyp_tokens.dim() != 2 or hyp_tokens.dtype != torch.int32:
        raise TypeError("hyp_tokens must be an int32 B x W tensor")
    dev = hyp_tokens.device
    hyp = hyp_tokens if hyp_tokens.stride(1) == 1 or hyp_tokens.size(1) <= 1 else hyp_tokens.contiguous()
    B, W = hyp.shape
    lens = torch.as_tensor(ref_lens).reshape(-1)
    if lens.numel() != B or hyp_counts.numel() != B:
        raise ValueError("hyp_counts and ref_lens must have one entry per utterance")
    lens_h = lens.to("cpu", torch.int64)
    if (lens_h < 0).any():
        raise ValueError("negative reference length")
    max_ref = int(lens_h.max()) if B else 0
    refs_t = torch.as_tensor(refs).reshape(-1)
    if int(lens_h.sum()) != refs_t.numel():
        raise ValueError("refs must hold exactly sum(ref_lens) entries")
    with torch.cuda.device(dev):
        offs = (torch.cumsum(lens_h, 0) - lens_h).to(torch.int32)
        meta = torch.stack([offs, lens_h.to(torch.int32)]).to(dev, non_blocking=True)      # one small H2D
        refs_d = refs_t.to(device=dev, dtype=torch.int32)
        if refs_d.numel() == 0:
            refs_d = torch.zeros(1, dtype=torch.int32, device=dev)
        counts = hyp_counts.to(device=dev, dtype=torch.int32).reshape(-1).contiguous()
        out = torch.empty((2, B), dtype=torch.int32, device=dev)
        st = lib.ctc_b200_edit_distance(hyp.data_ptr(), hyp.stride(0), counts.data_ptr(), W, refs_d.data_ptr(),
                                        meta[0].data_ptr(), meta[1].data_ptr(), max_ref, B, int(space),
                                        _EDIT_MODES[mode], out[0].data_ptr(), out[1].data_ptr(),
                                        torch.cuda.current_stream(dev).cuda_stream)
        if st != _lib.CTC_STATUS_SUCCESS:
            raise RuntimeError("ctc_b200_edit_distance: " + _lib.status_string(lib, st))
    return out[0], out[1]


class GreedyDecoder:
    """`GreedyDecoder(labels, blank_index=0).decode(probs, sizes)` with the reference's return structure.
    `labels` is the alphabet in index order (a string / list such as data/labels.en.json; wrapped in an
    `OrderedAlphabet` as the reference wraps it in its OrderedLabelEncoder) or any object with `transform` /
    `inverse_transform` methods (the reference's OrderedLabelEncoder itself)."""

    def __init__(self, label_encoder, blank_index: int = 0):
        if isinstance(label_encoder, str):
            label_encoder = list(label_encoder)
        if isinstance(label_encoder, (set, list, tuple)):    # (decoder.py:42-46 wraps these in an OrderedLabelEncoder)
            label_encoder = OrderedAlphabet(label_encoder)
        self.label_encoder = label_encoder
        self.blank_index = blank_index
        self.space_index = self._find_space()

    def _find_space(self) -> int:
        """Index of ' ' in the alphabet (-1 if it has none: then every transcript is a single word)."""
        try:
            return int(self.label_encoder.transform([" "])[0])
        except (ValueError, KeyError, IndexError):
            return -1

    # -- references to text (decoder.py:100-141; called on the TARGETS at test.py:79 and codes/metrics.py:112) ----
    def convert_to_strings(self, sequences, sizes=None, remove_repetitions=False, return_offsets=False):
        """Given a list of numeric sequences, returns the corresponding strings (`[[text], ...]`, and the frame
        offsets `[[IntTensor], ...]` when asked).  Host-side: the targets are already host tensors."""
        strings = []
        offsets = [] if return_offsets else None
        for i in range(len(sequences)):
            seq_len = int(sizes[i]) if sizes is not None else len(sequences[i])
            string, string_offsets = self.process_string(sequences[i], seq_len, remove_repetitions)
            strings.append([string])                         # one path per utterance
            if return_offsets:
                offsets.append([string_offsets])
        if return_offsets:
            return strings, offsets
        return strings

    def process_string(self, sequence, size, remove_repetitions=False):
        """Drops blanks (and, if asked, symbols equal to the previous FRAME) from `sequence[:size]`; returns the
        text and the kept positions."""
        seq = torch.as_tensor(sequence).reshape(-1)[:size].tolist()      # one transfer instead of one .item() per frame
        kept, offs = [], []
        for i, cur in enumerate(seq):
            if cur == self.blank_index:
                continue
            if remove_repetitions and i != 0 and cur == seq[i - 1]:
                continue
            kept.append(cur)
            offs.append(i)
        return self._to_string(kept), torch.IntTensor(offs)

    # -- scoring (decoder.py:49-78) ------------------------------------------------------------------------
    @staticmethod
    def _score_strings(s1: str, s2: str, mode: str) -> int:
        dev = torch.device("cuda", torch.cuda.current_device())
        a = torch.tensor([[ord(c) for c in s1] or [0]], dtype=torch.int32, device=dev)
        n = torch.tensor([len(s1)], dtype=torch.int32, device=dev)
        b = torch.tensor([ord(c) for c in s2], dtype=torch.int32)
        d, _ = edit_distance_raw(a, n, b, torch.tensor([len(s2)]), space=ord(" "), mode=mode)
        return int(d.item())

    def wer(self, s1: str, s2: str) -> int:
        """Word-level edit distance between two space-separated sentences (decoder.py:49-66).  Words are cut at
        runs of ' ' (the only whitespace the reference's alphabets contain)."""
        return self._score_strings(s1, s2, "wer")

    def cer(self, s1: str, s2: str) -> int:
        """Character-level edit distance after removing spaces (decoder.py:69-78)."""
        return self._score_strings(s1, s2, "cer")

    def error_counts(self, probs, sizes, targets, target_sizes):
        """Decode `probs` (B x T x V, CUDA) and score every utterance against its reference without leaving the
        device: returns a dict of CPU int64 tensors [B] -- 'wer', 'words' (reference word count), 'cer', 'chars'
        (reference length including spaces) -- i.e. what test.py:83-88 accumulates per utterance."""
        tokens, _, counts = greedy_decode_raw(probs, sizes, self.blank_index, want_offsets=False)
        refs, lens = _drop_blank(targets, target_sizes, self.blank_index)
        wd, wn = edit_distance_raw(tokens, counts, refs, lens, self.space_index, "wer")
        cd, cn = edit_distance_raw(tokens, counts, refs, lens, self.space_index, "cer")
        res = torch.stack([wd, wn, cd, cn]).cpu().to(torch.int64)                          # one D2H
        return {"wer": res[0], "words": res[1], "cer": res[2], "chars": res[3]}

    def _to_string(self, ids):
        if not ids:
            return ""
        return "".join(self.label_encoder.inverse_transform(ids))

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


def _drop_blank(targets, target_sizes, blank):
    """convert_to_strings (decoder.py:100-121) skips blank tokens in the references too; they never occur in
    real transcripts, so this is a host-side no-op unless one is present."""
    t = torch.as_tensor(targets).reshape(-1).cpu()
    lens = torch.as_tensor(target_sizes).reshape(-1).cpu().to(torch.int64)
    keep = t != blank
    if bool(keep.all()):
        return t, lens
    owner = torch.repeat_interleave(torch.arange(lens.numel()), lens)
    new_lens = torch.zeros_like(lens).index_add_(0, owner[keep], torch.ones(int(keep.sum()), dtype=torch.int64))
    return t[keep], new_lens
