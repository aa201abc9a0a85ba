"""CPU restatement of the reference's greedy CTC decode -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Follows /root/reference/codes/decoder.py:123-160 (`GreedyDecoder.process_string` / `decode`):
argmax over the last axis of a B x T x V tensor (`torch.max(probs, 2)`: first maximum wins), then per
utterance, for i < size: skip blanks; skip a symbol equal to the previous FRAME's argmax; keep
(symbol, frame index).  PARITY UNPINNED by the reference (it has no tests); pinned here by hand-made cases
in tests/test_decode.py.
"""
import numpy as np


def greedy_decode(probs_btv, sizes=None, blank=0):
    probs = np.asarray(probs_btv)
    B, T, V = probs.shape
    am = probs.argmax(axis=2)
    tokens, offsets = [], []
    for b in range(B):
        n = T if sizes is None else int(min(max(int(sizes[b]), 0), T))
        seq, off = [], []
        for i in range(n):
            c = int(am[b, i])
            if c == blank:
                continue
            if i != 0 and c == int(am[b, i - 1]):
                continue
            seq.append(c)
            off.append(i)
        tokens.append(np.asarray(seq, dtype=np.int32))
        offsets.append(np.asarray(off, dtype=np.int32))
    return tokens, offsets
