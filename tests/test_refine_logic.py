"""The argument behind knn_refine_kernel's "open rank" shortcut (csrc/prdc.cu), checked on the CPU.

The kernel holds, per row, approximate keys a_c with |a_c - e_c| <= band of the exact keys e_c, sorted by
a.  It (1) never evaluates candidates with a_c > a_(k) + 2 band, and (2) only COUNTS the candidates in front
of the last gap wider than 2 band at or below position k, looking for the (k+1)-th smallest exact key among
the rest.  This restates that selection in numpy and compares it with sorting the exact keys, over random
lists with clustered values, ties and adversarial noise at the band's edge."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st


def select(a_sorted, exact_of, k, band):
    """Index arithmetic of the kernel: returns the (k+1)-th smallest exact key it would report."""
    n = len(a_sorted)
    assert n >= k + 1
    cut = a_sorted[k] + 2 * band
    j0 = 0
    for c in range(1, k + 1):
        if a_sorted[c] - a_sorted[c - 1] > 2 * band:
            j0 = c
    evaluated = [exact_of[c] for c in range(j0, n) if a_sorted[c] <= cut]
    evaluated.sort()
    return evaluated[k - j0]


@settings(max_examples=300, deadline=None)
@given(st.integers(0, 2**31 - 1), st.integers(1, 10), st.sampled_from([1e-4, 1e-3, 1e-2, 0.1]))
def test_open_rank_selection_is_exact(seed, k, band):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(k + 1, 2 * (k + 1) + 3))
    # exact keys: a few clusters so that gaps both wider and narrower than the band occur, plus exact ties
    centres = rng.uniform(0, 1, size=int(rng.integers(1, 5)))
    e = centres[rng.integers(0, len(centres), size=n)] + rng.normal(0, band * rng.choice([0.3, 1.0, 5.0]), size=n)
    if n > 3 and rng.random() < 0.3:
        e[1] = e[0]
    # approximate keys: noise anywhere inside the band, sometimes pushed to its edge
    noise = rng.uniform(-band, band, size=n)
    edge = rng.random(n) < 0.3
    noise[edge] = np.sign(noise[edge]) * band
    a = e + noise
    order = np.argsort(a, kind="stable")
    got = select(a[order], e[order], k, band)
    want = np.sort(e)[k]
    assert got == want


def test_no_gap_and_all_gaps():
    e = np.array([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0])          # identical rows: no gap, everything evaluated
    assert select(e, e, 5, 1e-3) == 0.0
    e = np.arange(8, dtype=float)                               # wide gaps everywhere: one candidate evaluated
    a = e + 0.4 * np.array([1, -1, 1, -1, 1, -1, 1, -1])
    assert select(a, e, 5, 0.4 + 1e-9) == 5.0


def _fma32(a, b, c):
    """fp32 fused multiply-add emulated through fp64 (the product of two fp32 values is exact in fp64;
    the two roundings that follow are both monotone, which is all the argument below uses)."""
    return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(np.float32)


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 2**31 - 1))
def test_group_prefilter_is_a_necessary_condition(seed):
    """epilogues.cuh (TopkEpi / CountEpi): with sc < 0, m = max acc_j over a group and cmin <= |y_j|^2,
    every key fma(acc_j, sc, |y_j|^2) is >= fma(m, sc, cmin) — so `bound < thr` may only fail when no key of
    the group is below thr.  Checked with values chosen to sit on rounding boundaries."""
    rng = np.random.default_rng(seed)
    scale = np.float32(2.0 ** int(rng.integers(-12, 4)))
    sc = np.float32(-2.0) * scale                                   # a power of two, negative
    acc = rng.integers(-2**20, 2**20, size=8).astype(np.float32)    # fp32 accumulators of integer-scaled products
    ny = (rng.uniform(0.5, 2.0, size=8) * rng.choice([1.0, 1e-3, 1e3])).astype(np.float32)
    cmin = np.float32(ny.min() if rng.random() < 0.7 else np.nextafter(ny.min(), np.float32(0)))
    keys = _fma32(acc, sc, ny)
    bound = _fma32(acc.max(keepdims=True), sc, np.array([cmin], dtype=np.float32))[0]
    assert (keys >= bound).all()
    thr = np.float32(rng.choice(list(keys) + [np.nextafter(keys.min(), np.float32(np.inf)), np.float32(np.inf)]))
    if (keys < thr).any():
        assert bound < thr
