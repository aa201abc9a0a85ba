"""numpy model of the operand split the classifier-head kernels feed the tensor cores (csrc/ctc_head_tc.cuh,
csrc/ctc_head_bwd_tc.cuh): a = a_hi + a_lo with a_hi = tf32_rna(a), a_lo = a - a_hi (exact in fp32; the MMA reads its top 19
bits), products a_hi*b_hi + a_hi*b_lo + a_lo*b_hi accumulated in fp32.  Checks the claim the kernels' comments make: the
result is at fp32 fidelity (plain tf32 is ~1e-3), including inputs with a large common offset once they are centred."""
import numpy as np


def tf32_rna(x):
    """cvt.rna.tf32.f32 as the kernels do it: two integer operations on the bit pattern."""
    b = x.astype(np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def tf32_trunc(x):
    """what the tensor core reads of an fp32 operand: sign, exponent, 10 mantissa bits."""
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def split(x):
    hi = tf32_rna(x)
    lo = (x.astype(np.float32) - hi).astype(np.float32)
    return hi, tf32_trunc(lo)


def dot3(a, b):
    """sum_k of the three tf32 products, fp32 accumulation in the order the kernels issue them (K ascending)."""
    ah, al = split(a)
    bh, bl = split(b)
    acc = np.zeros((a.shape[0], b.shape[1]), np.float32)
    for k in range(a.shape[1]):
        acc += np.outer(ah[:, k], bh[k]).astype(np.float32)
        acc += np.outer(ah[:, k], bl[k]).astype(np.float32)
        acc += np.outer(al[:, k], bh[k]).astype(np.float32)
    return acc


def test_split_is_exact_and_small():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(4096) * np.exp(rng.uniform(-20, 20, 4096))).astype(np.float32)
    hi = tf32_rna(x)
    lo = x - hi
    assert np.all(hi.astype(np.float64) + lo.astype(np.float64) == x.astype(np.float64))       # a - a_hi is exact in fp32
    assert np.all(np.abs(lo) <= np.abs(x) * 2.0 ** -11 * (1 + 1e-6))
    assert np.all((hi.view(np.uint32) & np.uint32(0x1FFF)) == 0)


def test_three_products_reach_fp32_fidelity_where_plain_tf32_does_not():
    rng = np.random.default_rng(1)
    a = rng.standard_normal((64, 800)).astype(np.float32)
    b = (rng.standard_normal((800, 29)) / np.sqrt(800)).astype(np.float32)
    want = a.astype(np.float64) @ b.astype(np.float64)
    scale = np.abs(want).max()
    err3 = np.abs(dot3(a, b) - want).max() / scale
    ah, _ = split(a)
    bh, _ = split(b)
    err1 = np.abs(ah.astype(np.float64) @ bh.astype(np.float64) - want).max() / scale
    err_fp32 = np.abs((a @ b).astype(np.float64) - want).max() / scale
    assert err1 > 1e-4                                       # 10 mantissa bits on both operands
    assert err3 < 2e-6 and err3 < 8 * max(err_fp32, 1e-7)    # the split: within a small factor of an fp32 product


def test_centred_inputs_with_a_large_offset():
    """The kernels subtract the batch mean BEFORE the split (folding it into the bias cancels catastrophically)."""
    rng = np.random.default_rng(2)
    x = (rng.standard_normal((128, 256)) + 60.0).astype(np.float32)
    w = (rng.standard_normal((256, 29)) / 16).astype(np.float32)
    mu = x.mean(0, dtype=np.float64).astype(np.float32)
    want = (x.astype(np.float64) - mu.astype(np.float64)) @ w.astype(np.float64)
    got = dot3((x - mu).astype(np.float32), w)
    assert np.abs(got - want).max() / np.abs(want).max() < 2e-6
    folded = dot3(x, w) - (mu.astype(np.float64) @ w.astype(np.float64)).astype(np.float32)
    assert np.abs(folded - want).max() > 5 * np.abs(got - want).max()
