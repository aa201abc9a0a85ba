"""WER / CER scoring (second half): the CPU restatement against classical known answers and
an independent implementation; the GPU kernel (through the C ABI) against the restatement, bit-exact."""
import numpy as np
import pytest
import torch

from oracle import edit_distance as E
from oracle.greedy_decode import greedy_decode

ALPHABET = "_'abcdefghijklmnopqrstuvwxyz "          # data/labels.en.json order: blank first, space last
SPACE = ALPHABET.index(" ")


# ---------------------------------------------------------------- CPU: pin the restatement
@pytest.mark.parametrize("a,b,d", [
    ("kitten", "sitting", 3), ("flaw", "lawn", 2), ("saturday", "sunday", 3), ("", "abc", 3), ("abc", "", 3),
    ("", "", 0), ("abc", "abc", 0), ("intention", "execution", 5), ("gumbo", "gambol", 2), ("a", "b", 1),
    ("rosettacode", "raisethysword", 8), ("sleep", "fleeting", 5),
])
def test_levenshtein_known_answers(a, b, d):
    assert E.levenshtein(a, b) == d
    assert E.levenshtein_plain(a, b) == d
    assert E.levenshtein(b, a) == d


def test_levenshtein_against_independent_implementation_and_axioms():
    rng = np.random.default_rng(11)
    for _ in range(400):
        a = rng.integers(0, 4, rng.integers(0, 40)).tolist()
        b = rng.integers(0, 4, rng.integers(0, 40)).tolist()
        c = rng.integers(0, 4, rng.integers(0, 40)).tolist()
        dab = E.levenshtein(a, b)
        assert dab == E.levenshtein_plain(a, b)
        assert dab == E.levenshtein(b, a)
        assert abs(len(a) - len(b)) <= dab <= max(len(a), len(b))
        assert E.levenshtein(a, c) <= dab + E.levenshtein(b, c)
        assert (dab == 0) == (a == b)


def test_wer_cer_strings():
    # decoder.py:49-78 semantics: split on runs of whitespace; CER ignores spaces entirely
    assert E.wer("the cat sat", "the cat sat") == 0
    assert E.wer("the cat sat", "the cat sat down") == 1
    assert E.wer("  the   cat ", "the cat") == 0
    assert E.wer("", "a b c") == 3
    assert E.wer("cat the", "the cat") == 2
    assert E.wer("a a a", "a") == 2
    assert E.cer("the cat", "thecat") == 0
    assert E.cer("the cat", "the bat") == 1
    assert E.cer("", "a b") == 2


def test_token_scoring_matches_string_scoring():
    rng = np.random.default_rng(5)
    for _ in range(100):
        hyp = rng.choice([SPACE, 2, 3, 4], rng.integers(0, 30)).tolist()
        ref = rng.choice([SPACE, 2, 3, 4], rng.integers(0, 30)).tolist()
        hs, rs = "".join(ALPHABET[i] for i in hyp), "".join(ALPHABET[i] for i in ref)
        assert E.score_tokens(hyp, ref, SPACE, "wer") == (E.wer(hs, rs), len(rs.split()))
        assert E.score_tokens(hyp, ref, SPACE, "cer") == (E.cer(hs, rs), len(rs))


# ---------------------------------------------------------------- GPU: kernel vs restatement
def _random_batch(rng, B, max_h, max_r, vocab, p_space):
    hl = rng.integers(0, max_h + 1, B)
    rl = rng.integers(0, max_r + 1, B)
    hl[0], rl[0] = max_h, max_r
    if B > 2:
        hl[1], rl[2] = 0, 0
    syms = np.arange(1, vocab)

    def seq(n):
        s = rng.choice(syms, n)
        s[rng.random(n) < p_space] = SPACE
        return s.astype(np.int32)
    hyps = [seq(n) for n in hl]
    refs = []
    for b, n in enumerate(rl):                     # references correlated with the hypotheses (like real scoring)
        r = seq(n)
        k = min(n, hl[b])
        keep = rng.random(k) < 0.7
        r[:k][keep] = hyps[b][:k][keep]
        refs.append(r)
    return hyps, refs


def _run_gpu(hyps, refs, mode, space=SPACE):
    from aes_lac_2018_b200 import edit_distance_raw
    B = len(hyps)
    W = max(1, max(len(h) for h in hyps))
    tok = np.full((B, W), 7, np.int32)             # garbage beyond the counts must be ignored
    for b, h in enumerate(hyps):
        tok[b, :len(h)] = h
    cnt = np.array([len(h) for h in hyps], np.int32)
    flat = np.concatenate(refs) if sum(len(r) for r in refs) else np.zeros(0, np.int32)
    d, n = edit_distance_raw(torch.tensor(tok).cuda(), torch.tensor(cnt).cuda(), torch.tensor(flat),
                             torch.tensor([len(r) for r in refs]), space=space, mode=mode)
    return d.cpu().numpy(), n.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("B,max_h,max_r", [(9, 40, 30), (5, 300, 127), (5, 310, 128), (4, 750, 255), (4, 750, 256),
                                           (3, 900, 600), (2, 1500, 1100), (2, 3000, 2047), (6, 1, 1)])
def test_gpu_edit_distance_matches_restatement(B, max_h, max_r):
    rng = np.random.default_rng(B * 100000 + max_h * 10 + max_r)
    for p_space in (0.0, 0.18, 0.6):
        hyps, refs = _random_batch(rng, B, max_h, max_r, 29, p_space)
        for mode in ("tokens", "cer", "wer"):
            d, n = _run_gpu(hyps, refs, mode)
            for b in range(B):
                want = E.score_tokens(hyps[b], refs[b], SPACE, mode)
                assert (int(d[b]), int(n[b])) == want, (mode, p_space, b, len(hyps[b]), len(refs[b]))


@pytest.mark.gpu
def test_gpu_edit_distance_hand_cases():
    enc = lambda s: np.array([ALPHABET.index(c) for c in s], np.int32)
    pairs = [("the cat sat", "the cat sat down"), ("  the   cat ", "the cat"), ("", "a b c"), ("cat the", "the cat"),
             ("a a a", "a"), ("   ", ""), ("", ""), ("kitten", "sitting"), ("hello world", "hello  world"),
             ("abc abd abc", "abd abc abd")]
    hyps, refs = [enc(a) for a, _ in pairs], [enc(b) for _, b in pairs]
    dw, nw = _run_gpu(hyps, refs, "wer")
    dc, nc = _run_gpu(hyps, refs, "cer")
    for i, (a, b) in enumerate(pairs):
        assert dw[i] == E.wer(a, b) and nw[i] == len(b.split()), (a, b)
        assert dc[i] == E.cer(a, b) and nc[i] == len(b), (a, b)
    # an alphabet without a space symbol: every transcript is one word
    d, n = _run_gpu([enc("abc"), enc("abc")], [enc("abc"), enc("abd")], "wer", space=-1)
    assert d.tolist() == [0, 1] and n.tolist() == [1, 1]


@pytest.mark.gpu
def test_decoder_wer_cer_and_metrics_follow_the_reference_flow():
    """GreedyDecoder.wer/cer on strings, error_counts and the WER/CER metric objects against the reference's
    flow restated on the host: decode -> strings -> wer/cer per utterance -> normalise -> mean * 100."""
    from aes_lac_2018_b200 import CER, WER, GreedyDecoder
    dec = GreedyDecoder(ALPHABET, blank_index=0)
    assert dec.space_index == SPACE
    assert dec.wer("the cat sat", "the cat sat down") == 1 and dec.cer("the cat", "the bat") == 1
    assert dec.wer("", "") == 0 and dec.cer("a b", "") == 2

    rng = np.random.default_rng(3)
    B, T, V = 12, 200, len(ALPHABET)
    probs = rng.standard_normal((B, T, V)).astype(np.float32)
    probs[..., 0] += 2.0 * (rng.random((B, T)) < 0.6)
    probs[..., SPACE] += 1.5 * (rng.random((B, T)) < 0.15)
    sizes = rng.integers(T // 2, T + 1, B).astype(np.int32)
    hyp_tok, _ = greedy_decode(probs, sizes, blank=0)
    refs = []
    for b in range(B):                              # references: noisy copies of the hypotheses
        r = hyp_tok[b].copy()
        flip = rng.random(r.size) < 0.2
        r[flip] = rng.integers(1, V, int(flip.sum()))
        refs.append(r[: max(0, r.size - rng.integers(0, 4))].astype(np.int32))
    refs[3] = np.zeros(0, np.int32)                 # empty reference: normaliser 0 => no division
    flat = torch.tensor(np.concatenate(refs))
    lens = torch.tensor([len(r) for r in refs], dtype=torch.int32)
    to_s = lambda ids: "".join(ALPHABET[i] for i in ids)

    want_w = [E.wer(to_s(hyp_tok[b]), to_s(refs[b])) for b in range(B)]
    want_c = [E.cer(to_s(hyp_tok[b]), to_s(refs[b])) for b in range(B)]
    words = [len(to_s(r).split()) for r in refs]
    chars = [len(r) for r in refs]
    out = torch.tensor(probs).cuda()
    got = dec.error_counts(out, torch.tensor(sizes), flat, lens)
    assert got["wer"].tolist() == want_w and got["words"].tolist() == words
    assert got["cer"].tolist() == want_c and got["chars"].tolist() == chars

    for cls, dist, den in ((WER, want_w, words), (CER, want_c, chars)):
        m = cls(dec)
        m.update((out, flat, torch.tensor(sizes), lens))
        m.update((out, flat, torch.tensor(sizes), lens))
        want = sum(d / n if n else d for d, n in zip(dist, den)) * 2 / (2 * B) * 100
        assert abs(m.compute() - want) < 1e-9 * max(1.0, want)
        ms = cls(dec, stateful=True)
        ms.update((out, flat, torch.tensor(sizes), lens))
        assert abs(ms.compute() - sum(dist) / sum(den) * 100) < 1e-9
    with pytest.raises(RuntimeError):
        WER(dec).compute()


# ---------------------------------------------------------------- CPU: host-side helpers of the scoring path
def test_host_helpers_space_index_and_blank_dropping():
    from aes_lac_2018_b200.decoder import GreedyDecoder, _drop_blank
    assert GreedyDecoder(ALPHABET).space_index == SPACE
    assert GreedyDecoder("_abc").space_index == -1                      # alphabet without a space
    assert GreedyDecoder(list("_ ab")).space_index == 1

    class Enc:                                                            # OrderedLabelEncoder-like object
        def transform(self, xs):
            return [{"_": 0, "a": 1, " ": 2}[x] for x in xs]
    assert GreedyDecoder(Enc()).space_index == 2

    t = torch.tensor([1, 2, 3, 4, 5, 6], dtype=torch.int32)
    lens = torch.tensor([2, 0, 4], dtype=torch.int32)
    r, l = _drop_blank(t, lens, blank=0)                                 # nothing to drop: untouched
    assert r.tolist() == t.tolist() and l.tolist() == [2, 0, 4]
    t = torch.tensor([0, 2, 0, 0, 5, 6, 0], dtype=torch.int32)
    lens = torch.tensor([3, 1, 3], dtype=torch.int32)
    r, l = _drop_blank(t, lens, blank=0)                                 # convert_to_strings skips blanks (decoder.py:127)
    assert r.tolist() == [2, 5, 6] and l.tolist() == [1, 0, 2]
