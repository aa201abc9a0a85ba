"""Greedy decode: CPU restatement vs hand-made cases, GPU kernel vs the restatement."""
import numpy as np
import pytest
import torch

from oracle.greedy_decode import greedy_decode


def _onehot(seq, V, rng=None):
    x = np.full((len(seq), V), -1.0, np.float32) if rng is None else rng.standard_normal((len(seq), V)).astype(np.float32)
    for t, c in enumerate(seq):
        x[t, c] = 5.0
    return x


def test_oracle_hand_cases():
    V = 5
    #            _  a  a  _  a  b  b  _  _  c
    seq = [0, 1, 1, 0, 1, 2, 2, 0, 0, 3]
    tok, off = greedy_decode(_onehot(seq, V)[None], None, blank=0)
    assert tok[0].tolist() == [1, 1, 2, 3] and off[0].tolist() == [1, 4, 5, 9]
    tok, off = greedy_decode(_onehot(seq, V)[None], [6], blank=0)
    assert tok[0].tolist() == [1, 1, 2] and off[0].tolist() == [1, 4, 5]
    tok, _ = greedy_decode(_onehot([0, 0, 0], V)[None], None)
    assert tok[0].size == 0
    tok, _ = greedy_decode(_onehot([2, 2, 2], V)[None], [0])
    assert tok[0].size == 0
    # ties: the first maximum wins (torch.max / np.argmax on the CPU)
    x = np.zeros((1, 2, V), np.float32)
    tok, _ = greedy_decode(x, None, blank=0)
    assert tok[0].size == 0
    # blank in the middle of the alphabet
    tok, off = greedy_decode(_onehot([3, 3, 2, 3, 1], V)[None], None, blank=3)
    assert tok[0].tolist() == [2, 1] and off[0].tolist() == [2, 4]


@pytest.mark.gpu
@pytest.mark.parametrize("V,T,B", [(29, 750, 17), (43, 333, 9), (29, 31, 3), (5, 64, 4), (100, 70, 3)])
def test_gpu_decode_matches_restatement(V, T, B):
    from aes_lac_2018_b200 import greedy_decode_raw
    rng = np.random.default_rng(V * 1000 + T)
    probs = rng.standard_normal((B, T, V)).astype(np.float32)
    probs[..., 0] += 1.5 * (rng.random((B, T)) < 0.5)                       # plenty of blanks
    rep = rng.random((B, T)) < 0.3                                           # and repeated frames
    for b in range(B):
        for t in range(1, T):
            if rep[b, t]:
                probs[b, t] = probs[b, t - 1]
    sizes = rng.integers(0, T + 1, B).astype(np.int32)
    sizes[0] = T
    for blank in (0, V - 1):
        for sz in (None, sizes):
            tok, off, cnt = greedy_decode_raw(torch.tensor(probs).cuda(), sz, blank=blank)
            want_t, want_o = greedy_decode(probs, sz, blank)
            cnt = cnt.cpu().numpy()
            for b in range(B):
                assert cnt[b] == len(want_t[b])
                assert tok[b, :cnt[b]].cpu().numpy().tolist() == want_t[b].tolist()
                assert off[b, :cnt[b]].cpu().numpy().tolist() == want_o[b].tolist()


@pytest.mark.gpu
def test_gpu_decoder_class_and_strided_view():
    """Same return structure as the reference's GreedyDecoder.decode; also accepts the T x B x V storage viewed as
    B x T x V (what `out.transpose(0, 1)` gives)."""
    from aes_lac_2018_b200 import GreedyDecoder
    labels = "_ 'ABCDEFGHIJKLMNOPQRSTUVWXYZ"                                 # data/labels.en.json order, blank first
    V = len(labels)
    text = "HELLO WORLD"
    seq = []
    for ch in text:
        seq += [labels.index(ch)] * 2 + [0]
    probs = torch.tensor(_onehot(seq, V)[None]).cuda()
    dec = GreedyDecoder(labels, blank_index=0)
    strings, offsets = dec.decode(probs, torch.tensor([len(seq)], dtype=torch.int32))
    assert strings == [[text]]
    assert offsets[0][0].tolist() == [3 * i for i in range(len(text))]
    tbv = probs.transpose(0, 1).contiguous()                                  # T x B x V storage
    strings2, _ = dec.decode(tbv.transpose(0, 1), None)
    assert strings2 == [[text]]
    with pytest.raises(RuntimeError):
        dec.decode(probs.cpu(), None)
