"""Host-side surface of the GreedyDecoder drop-in (no GPU): the methods the reference calls on the TARGETS --
`target_decoder.convert_to_strings(split_targets)` at /root/reference/test.py:79 and
`self._decoder.convert_to_strings(split_targets)` at /root/reference/codes/metrics.py:112 -- and the encoder the
reference's Decoder.__init__ wraps list / str alphabets in (codes/decoder.py:39-46)."""
import torch

from aes_lac_2018_b200.decoder import GreedyDecoder, OrderedAlphabet

LABELS = "_'abcdefghijklmnopqrstuvwxyz "          # data/labels.en.json order: blank first, space last


def _reference_process_string(labels, blank, sequence, size, remove_repetitions):
    """Restatement of the loop at codes/decoder.py:123-141 (one .item() per frame)."""
    out, offs = [], []
    for i in range(size):
        cur = sequence[i].item()
        if cur != blank:
            if remove_repetitions and i != 0 and cur == sequence[i - 1].item():
                continue
            out.append(cur)
            offs.append(i)
    return "".join(labels[c] for c in out), torch.IntTensor(offs)


def test_alphabet_wrapping_and_lookup():
    for alphabet in (LABELS, list(LABELS), tuple(LABELS)):
        d = GreedyDecoder(alphabet, blank_index=0)
        assert isinstance(d.label_encoder, OrderedAlphabet)
        assert d.label_encoder.transform(["a", " ", "_"]) == [2, 28, 0]
        assert d.label_encoder.inverse_transform([2, 28, 0]) == ["a", " ", "_"]
        assert d.space_index == 28
    enc = OrderedAlphabet("abca")                              # first-appearance order, duplicates ignored
    assert enc.classes_ == ["a", "b", "c"] and len(enc) == 3
    try:
        enc.transform(["z"])
    except ValueError:
        pass
    else:
        raise AssertionError("unseen label must raise")
    assert GreedyDecoder("ab_", blank_index=2).space_index == -1


def test_convert_to_strings_matches_the_reference_loop():
    g = torch.Generator().manual_seed(5)
    dec = GreedyDecoder(LABELS, blank_index=0)
    targets = torch.randint(0, len(LABELS), (300,), generator=g, dtype=torch.int32)
    target_sizes = torch.tensor([0, 7, 50, 1, 120, 122], dtype=torch.int32)
    split, off = [], 0
    for size in target_sizes:                                  # the unflatten loop of test.py:60-66
        split.append(targets[off:off + size])
        off += size
    strings = dec.convert_to_strings(split)
    assert len(strings) == len(split)
    for seq, (text,) in zip(split, strings):
        want, _ = _reference_process_string(LABELS, 0, seq, len(seq), False)
        assert text == want
    for rem in (False, True):
        strings, offsets = dec.convert_to_strings(split, sizes=[min(len(s), 40) for s in split],
                                                  remove_repetitions=rem, return_offsets=True)
        for seq, (text,), (offs,) in zip(split, strings, offsets):
            want, want_offs = _reference_process_string(LABELS, 0, seq, min(len(seq), 40), rem)
            assert text == want and offs.dtype == torch.int32 and torch.equal(offs, want_offs)
    text, offs = dec.process_string(torch.zeros(4, dtype=torch.int32), 4)
    assert text == "" and offs.numel() == 0
