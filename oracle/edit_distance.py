"""CPU restatement of the reference's WER / CER scoring -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Follows /root/reference/codes/decoder.py:49-78 (`Decoder.wer`: split both sentences on whitespace, map words to
integers, Levenshtein distance over those; `Decoder.cer`: remove spaces, Levenshtein distance over characters)
and the normalisers of /root/reference/codes/metrics.py:143-162 (`len(reference.split())` / `len(reference)`,
and "no division when that is 0").  The distance itself lives in a third-party dependency that is absent here:
`python-Levenshtein` (`import Levenshtein as Lev`, decoder.py:19; unpinned in docker/requirements.txt).  Its
`distance(a, b)` is the classical unit-cost Levenshtein distance (insert = delete = substitute = 1), restated
below as the textbook two-row dynamic programme.  PARITY UNPINNED by the reference (no tests); pinned here by
the classical known answers in tests/test_editdist.py (kitten/sitting = 3, flaw/lawn = 2, ...), by the metric
axioms on random inputs, and by an independent recursive implementation on small cases.
"""
import numpy as np


def levenshtein(a, b) -> int:
    """Unit-cost edit distance between two sequences (strings, lists of hashables, integer arrays)."""
    ids = {}
    a = np.asarray([ids.setdefault(x, len(ids)) for x in a], dtype=np.int64)     # exact: equal items <-> equal ids
    b = np.asarray([ids.setdefault(x, len(ids)) for x in b], dtype=np.int64)
    if a.size == 0 or b.size == 0:
        return int(max(a.size, b.size))
    cols = np.arange(b.size + 1)
    prev = cols.copy()
    for i, x in enumerate(a, 1):
        tmp = np.minimum(prev[1:] + 1, prev[:-1] + (b != x))       # vertical / diagonal moves
        # horizontal moves: cur[j] = min_k<=j (tmp[k] + j - k), a running minimum of (value - column)
        prev = np.minimum.accumulate(np.concatenate(([i], tmp)) - cols) + cols
    return int(prev[-1])


def levenshtein_plain(a, b) -> int:
    """The same distance as three nested Python statements -- the independent check of `levenshtein`."""
    a, b = list(a), list(b)
    d = list(range(len(b) + 1))
    for i in range(1, len(a) + 1):
        nd = [i] + [0] * len(b)
        for j in range(1, len(b) + 1):
            nd[j] = min(d[j] + 1, nd[j - 1] + 1, d[j - 1] + (a[i - 1] != b[j - 1]))
        d = nd
    return d[len(b)]


def wer(s1: str, s2: str) -> int:
    """decoder.py:49-66."""
    return levenshtein(s1.split(), s2.split())


def cer(s1: str, s2: str) -> int:
    """decoder.py:69-78."""
    return levenshtein(s1.replace(' ', ''), s2.replace(' ', ''))


def split_words(tokens, space):
    """Token-id analogue of str.split(): maximal runs of non-space tokens."""
    words, cur = [], []
    for t in list(tokens):
        if t == space:
            if cur:
                words.append(tuple(cur))
                cur = []
        else:
            cur.append(int(t))
    if cur:
        words.append(tuple(cur))
    return words


def score_tokens(hyp, ref, space, mode):
    """(distance, normaliser) for one utterance given token-id sequences; mode in {'tokens', 'cer', 'wer'}."""
    hyp, ref = [int(x) for x in hyp], [int(x) for x in ref]
    if mode == "tokens":
        return levenshtein(hyp, ref), len(ref)
    if mode == "cer":
        return levenshtein([t for t in hyp if t != space], [t for t in ref if t != space]), len(ref)
    rw = split_words(ref, space)
    return levenshtein(split_words(hyp, space), rw), len(rw)
