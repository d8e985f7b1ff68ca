"""Error budget of the moment tables of the iterative estimators (kgl_gene_b200/csrc/terms_moments.cuh), restated in numpy:
the contract of the path is 1e-6 on F (north_star), the device tests assert 1e-9, this file shows where the expansion itself
stands. No GPU, no oracle: the arithmetic of the header -- bin = exponent + five mantissa bits of r = a/(1-a), six moments of
(r - rc)/w, 1/(f + r) = t sum_j (-(r - rc) t)^j with t = 1/(f + rc) -- against the exact sums."""
import numpy as np

SUB_BITS, J = 5, 6
VALID_MIN = -0.2            # kMomValidMin


def geometry(r):
    m, e = np.frexp(r)                       # r = m 2^e, m in [0.5, 1)  ->  r = (2m) 2^(e-1)
    e = e - 1
    sub = np.floor((2.0 * m - 1.0) * (1 << SUB_BITS)).astype(np.int64)
    rc = np.ldexp(1.0 + (sub + 0.5) / (1 << SUB_BITS), e)
    w = np.ldexp(1.0, e - SUB_BITS - 1)
    return e, sub, rc, w


def table_sums(r, f):
    """S1 = sum 1/(f + r), S2 = sum 1/(f + r)^2 from per-bin moments; bins of the octaves below 4|f| are summed exactly (the
    device takes them from the list of rare homozygous cells)."""
    e, sub, rc, w = geometry(r)
    edge = 0.0
    if f < 0.0:
        m, ex = np.frexp(-4.0 * f)
        if m == 0.5:
            ex -= 1
        edge = np.ldexp(1.0, ex)
    exact = r < edge
    s1 = np.sum(1.0 / (f + r[exact]))
    s2 = np.sum(1.0 / (f + r[exact]) ** 2)
    key = e[~exact] * (1 << SUB_BITS) + sub[~exact]
    v = (r[~exact] - rc[~exact]) / w[~exact]
    assert np.all(np.abs(v) <= 1.0)
    _, first, inv = np.unique(key, return_index=True, return_inverse=True)
    t = 1.0 / (f + rc[~exact][first])
    z = -w[~exact][first] * t
    for j in range(J):
        mom = np.bincount(inv, weights=v ** j, minlength=first.size)
        s1 += np.sum(t * mom * z ** j)
        s2 += np.sum(t * t * (j + 1) * mom * z ** j)
    return s1, s2


def test_bins_are_the_top_bits_of_the_double():
    r = np.array([1.0, 1.03124, 1.03126, 1.999, 2.0, 0.75, 3.0e-4, 4096.5])
    e, sub, rc, w = geometry(r)
    bits = r.view(np.int64)
    assert np.array_equal(e, ((bits >> 52) & 0x7FF) - 1023)
    assert np.array_equal(sub, (bits >> (52 - SUB_BITS)) & ((1 << SUB_BITS) - 1))
    assert np.all(np.abs(r - rc) <= w) and np.all(w / rc <= 1.0 / 64)


def test_expansion_error_over_the_domain():
    rng = np.random.default_rng(7)
    # a site-frequency spectrum like the bench's: a = q for hom-ref cells (close to 1), a = p for rare hom-alt cells
    p = np.clip(rng.beta(0.2, 2.0, size=200_000), 1e-4, 0.9999)
    a = np.concatenate([1.0 - p[:150_000], p[150_000:]])
    r = a / (1.0 - a)
    worst = 0.0
    for f in (VALID_MIN, -0.1, -0.031, -1e-3, 0.0, 1e-6, 0.02, 0.25, 0.7, 1.0):
        s1, s2 = table_sums(r, f)
        keep = (f + r) > 0
        assert keep.all() or f < 0
        e1 = abs(s1 - np.sum(1.0 / (f + r))) / abs(np.sum(1.0 / np.abs(f + r)))
        e2 = abs(s2 - np.sum(1.0 / (f + r) ** 2)) / np.sum(1.0 / (f + r) ** 2)
        worst = max(worst, e1, e2)
    assert worst < 2e-11, worst


def test_worst_case_bin_at_the_left_end_of_the_domain():
    # every cell at the edge of its bin, f at the left end: the bound (1/48)^6 / (1 - 1/48) of the header
    f = VALID_MIN
    rc = 1.0 + 0.5 / 32
    w = 1.0 / 64
    r = np.full(1000, rc + w * (1 - 1e-12))
    s1, _ = table_sums(r, f)
    rel = abs(s1 - np.sum(1.0 / (f + r))) / np.sum(1.0 / (f + r))
    assert rel < (1.0 / 48) ** 6 / (1.0 - 1.0 / 48) * 1.01
    assert rel < 1e-10


def test_root_of_the_likelihood_moves_less_than_the_tolerance_of_the_tests():
    rng = np.random.default_rng(11)
    p = np.clip(rng.beta(0.2, 2.0, size=100_000), 1e-4, 0.9999)
    hom = np.concatenate([1.0 - p[:80_000], p[80_000:90_000]])
    r = hom / (1.0 - hom)
    n_het = 10_000.0

    def g1_exact(f):
        return np.sum(1.0 / (f + r)) - n_het / (1.0 - f)

    def g1_table(f):
        return table_sums(r, f)[0] - n_het / (1.0 - f)

    def root(g):
        lo, hi = -0.9 * r.min(), 0.999
        for _ in range(60):
            mid = 0.5 * (lo + hi)
            lo, hi = (mid, hi) if g(mid) > 0 else (lo, mid)
        return 0.5 * (lo + hi)

    assert abs(root(g1_exact) - root(g1_table)) < 1e-10
