"""Randomised parity: the device path against the CPU oracle on many small populations whose shape, spectrum, phase, missing
cells, missing frequencies, multi-allelic loci, frequency pokes (p = 1, p = 0.5, rare major, tiny) and locus selection (window,
spacing, AF range) are all drawn at random from a fixed seed. Everything the hot path returns is compared: selection bits,
allele counts, class counts (bit-exact), expected sums, the four estimators, logLikelihood on a grid, the CalcFWS bins, the
hetero/homo records, pairwise IBS and the dosage Gram matrix. Sizes keep the oracle at a fraction of a second per case."""
import os

import numpy as np
import pytest

import oracle_py as O
from conftest import results_matrix

pytestmark = pytest.mark.gpu

TIGHT = 1e-11


@pytest.fixture(scope="module")
def gpu():
    from kgl_gene_b200.capi import KglB200
    ctx = KglB200(0)
    yield ctx
    ctx.close()


def rel_err(a, b, floor=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), max(floor, 1e-300))) if a.size else 0.0


def sel_bits(sel):
    return np.bitwise_or.reduce(sel.astype(np.uint8) << np.arange(sel.shape[0], dtype=np.uint8)[:, None], axis=0).astype(np.uint8)


def draw_case(rng):
    from kgl_gene_b200.synth import add_multi_allelic, make_population
    if WIDE:
        n = int(rng.choice([700, 1500, 2504, 2561, 4100]))
        l = int(rng.choice([4097, 30000, 120000]))
    else:
        n = int(rng.choice([1, 2, 31, 64, 65, 100, 129, 200, 449, 513]))
        l = int(rng.choice([1, 5, 63, 64, 257, 1000, 3000, 8000, 20000]))
    kw = dict(n_genomes=n, n_loci=l, seed=int(rng.integers(1, 10**6)), spectrum=str(rng.choice(["sfs", "dense"])),
              grouped=bool(rng.integers(0, 2)), unphased=bool(rng.integers(0, 3) == 0),
              missing_rate=float(rng.choice([0.0, 0.002, 0.03])), missing_af_rate=float(rng.choice([0.0, 0.03])))
    pop, _ = make_population(**kw)
    poke = int(rng.integers(0, 4))
    if poke == 1:
        pop.af[:, ::max(1, l // 9)] = np.float32(rng.choice([1.0, 0.5, 0.996, 0.0004, 1e-7]))
    elif poke == 2:
        pop.af[int(rng.integers(0, 6)), 1::max(2, l // 5)] = np.float32(rng.choice([1.0, 0.9999, 0.0]))
    n_multi = int(rng.choice([0, 0, 1, 7, 40]))
    if n_multi and l >= 64:
        add_multi_allelic(pop, min(n_multi, l // 8), seed=int(rng.integers(1, 10**6)))
    sel_kw = dict(spacing=int(rng.choice([0, 0, 1, 25, 400])), min_af=float(rng.choice([0.0, 0.0, 0.01, 0.2])),
                  max_af=float(rng.choice([1.0, 1.0, 0.5, 0.95])))
    if rng.integers(0, 2) and l > 10:
        lo, hi = sorted(rng.integers(0, l, size=2).tolist())
        sel_kw.update(lower=int(pop.offsets[lo]), upper=int(pop.offsets[hi]) + int(rng.integers(0, 2)))
    return kw, pop, sel_kw


WIDE = os.environ.get("KGL_FUZZ_WIDE") == "1"        # thousands of genomes (the sliced streaming kernel, many CTAs): seconds of oracle per case
CASES_PER_SEED = 4 if WIDE else 25
SEEDS = [int(x) for x in os.environ.get("KGL_FUZZ_SEEDS", "11,23,37,41").split(",")]      # a longer campaign: KGL_FUZZ_SEEDS=1,2,3,...


@pytest.mark.parametrize("seed", SEEDS)
def test_random_populations_match_oracle(gpu, seed):
    from kgl_gene_b200 import fws
    rng = np.random.default_rng(seed)
    for case in range(CASES_PER_SEED):
        kw, pop, sel_kw = draw_case(rng)
        tag = (seed, case, kw, pop.n_multi, sel_kw)
        sel = O.select_all_pops(pop, **sel_kw)
        gpu.upload_population(pop)
        counts = gpu.select_loci(**sel_kw)
        assert np.array_equal(gpu.get_locus_selection(), sel_bits(sel)), tag
        assert np.array_equal(counts[:6], sel.sum(axis=1).astype(np.uint64)), tag

        olc, ogc = O.allele_count(pop)
        lc, res = gpu.count_and_inbreed()
        assert np.array_equal(lc, olc), tag
        start = rng.uniform(0.0, 0.9, size=pop.n_genomes)
        sweeps = int(rng.choice([1, 7, 50]))
        want_simple = O.inbreed(pop, sel, "Simple")
        has_terms = results_matrix(want_simple)[0][:, 4] > 0
        for algo, kwargs, okw in (("Simple", None, {}), ("RitlandLocus", {}, {}),
                                  ("HallME", dict(hall_start=start, hall_sweeps=sweeps), dict(start=start, sweeps=sweeps)),
                                  ("Loglikelihood", {}, {})):
            got = res if kwargs is None else gpu.inbreed(algo, **kwargs)
            want = want_simple if algo == "Simple" else O.inbreed(pop, sel, algo, **okw)
            c_got, f_got = results_matrix(got)
            c_want, f_want = results_matrix(want)
            assert np.array_equal(c_got, c_want), (tag, algo)
            # the expected sums are 64-bit fixed-point sums (quantum 2^-(62 - log2 of the window's rows)): a sum that is
            # itself below 1e-6 is compared on that scale
            assert rel_err(f_got, f_want, floor=1e-6) < 1e-10, (tag, algo)
            a, b = got["inbred_allele_sum"][has_terms], want["inbred_allele_sum"][has_terms]
            assert np.array_equal(np.isnan(a), np.isnan(b)), (tag, algo)
            fin = np.isfinite(b)
            assert np.array_equal(np.isfinite(a), fin), (tag, algo)
            if fin.any():
                bad = np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-3) >= 1e-8
                if algo == "Loglikelihood" and bad.any():
                    # a handful of loci: the objective can be flat where its terms are clamped, or monotone, and then has many
                    # maximisers -- the answer must be as good as the oracle's, not the same point
                    idx = np.flatnonzero(has_terms)[fin][bad]
                    for g in idx:
                        ll = O.loglik_grid(pop, sel, np.array([got["inbred_allele_sum"][g], want["inbred_allele_sum"][g]]))[g]
                        assert ll[0] >= ll[1] - 1e-9 * max(1.0, abs(ll[1])), (tag, algo, int(g), ll)
                    assert int(sel.sum(axis=1).max()) <= 64, (tag, "several maximisers on a selection of this size")
                else:
                    assert not bad.any(), (tag, algo)
        grid = np.array([-0.3, 0.0, 0.11, 0.6])
        g_got, g_want = gpu.loglik_grid(grid)[has_terms], O.loglik_grid(pop, sel, grid)[has_terms]
        fin = np.isfinite(g_want)
        assert np.array_equal(np.isfinite(g_got), fin), tag
        assert rel_err(g_got[fin], g_want[fin]) < 1e-11, tag

        _, gc = gpu.allele_count()
        assert np.array_equal(gc, ogc), tag
        got = fws.calc_fws(gpu, pop=5, n_multi=pop.n_multi)
        want, want_rows = O.fws_bins(pop, 5, fws.FWS_BINS)
        assert np.array_equal(got["bin_variants"], want_rows), tag
        assert np.array_equal(got["genome_bins"][:, :, 1:], want[:, :, 1:3]), tag
        assert np.array_equal(got["genome_bins"][:, :, 0], want[:, :, 0] + want[:, :, 3]), tag
        assert np.array_equal(gpu.hetero_homo(), O.hetero_homo(pop)), tag

        if pop.n_genomes ** 2 * pop.n_loci <= 3e8:
            want_ibs = O.ibs(pop)
            assert np.array_equal(gpu.ibs(), want_ibs), tag
            gpu.set_ibs_tensor_cores(False)
            assert np.array_equal(gpu.ibs(), want_ibs), tag
            gpu.set_ibs_tensor_cores(True)
            assert np.array_equal(gpu.gram(), O.gram(pop)[0]), tag


def test_narrow_window_of_a_long_contig_keeps_small_sums(gpu):
    """Three loci whose major allele has frequency 0.9999 (expected minor-homozygous frequency 1e-8 each) selected out of a
    300,000-locus contig: the fixed-point scale of the sums follows the rows of the selection window, not the contig length
    (its 2^-43 quantum would leave 1e-5 relative on 1e-8; the window's 2^-59 leaves 2e-10)."""
    from kgl_gene_b200.synth import make_population
    pop, _ = make_population(64, 300_000, seed=77, spectrum="sfs", missing_rate=0.001)
    pop.af[:, 150_000:150_003] = np.float32(0.9999)
    sel_kw = dict(lower=int(pop.offsets[150_000]), upper=int(pop.offsets[150_003]))
    sel = O.select_all_pops(pop, **sel_kw)
    assert 1 <= sel.sum(axis=1).max() <= 4
    gpu.upload_population(pop)
    gpu.select_loci(**sel_kw)
    _, res = gpu.count_and_inbreed()
    want = O.inbreed(pop, sel, "Simple")
    c_got, f_got = results_matrix(res)
    c_want, f_want = results_matrix(want)
    assert np.array_equal(c_got, c_want)
    nz = f_want > 0
    assert f_want[nz].min() < 1e-7
    assert rel_err(f_got[nz], f_want[nz]) < 1e-9
    # and a whole-contig selection right after it on the same context (the scale goes back with the window)
    sel = O.select_all_pops(pop)
    gpu.select_loci()
    _, res = gpu.count_and_inbreed()
    c_got, f_got = results_matrix(res)
    c_want, f_want = results_matrix(O.inbreed(pop, sel, "Simple"))
    assert np.array_equal(c_got, c_want) and rel_err(f_got, f_want, floor=1e-6) < TIGHT


def test_misuse_is_an_error_code_and_the_context_survives():
    """Calls out of order or with arguments that do not fit: every one returns an error code with a message (KglError through the
    binding), none crashes or hangs, and the same context then runs a correct sequence to the oracle's results."""
    import ctypes as C
    from kgl_gene_b200.capi import KglB200, KglError
    from kgl_gene_b200.synth import add_multi_allelic, make_population
    ctx = KglB200(0)
    try:
        pop, _ = make_population(70, 900, seed=5, missing_rate=0.01)
        # nothing uploaded yet
        for call in (lambda: ctx.select_loci(), lambda: ctx.allele_count(), lambda: ctx.inbreed("Simple"), lambda: ctx.count_and_inbreed(),
                     lambda: ctx.ibs(0, 1), lambda: ctx.gram(), lambda: ctx.hetero_homo(), lambda: ctx.multi_allele_count(3),
                     lambda: ctx.binned_genome_counts([0.0], [1.0]), lambda: ctx.loglik_grid([0.0]), lambda: ctx.count_loci()):
            with pytest.raises(KglError):
                call()
        # the matrix alone: no frequencies, no super-populations
        ctx.upload_genotypes(pop.packed, pop.n_genomes)
        for call in (lambda: ctx.select_loci(), lambda: ctx.inbreed("Simple"), lambda: ctx.binned_genome_counts([0.0], [1.0])):
            with pytest.raises(KglError):
                call()
        # tables that do not fit the matrix
        with pytest.raises(KglError):
            ctx.upload_loci(pop.af[:, :-1], pop.offsets[:-1]); ctx.set_genome_superpop(pop.superpop); ctx.select_loci(); ctx.inbreed("Simple")
        ctx.upload_genotypes(pop.packed, pop.n_genomes)
        ctx.upload_loci(pop.af, pop.offsets)
        with pytest.raises(KglError):
            ctx.set_genome_superpop(np.full(pop.n_genomes, 6, dtype=np.uint8))                  # index out of range
        with pytest.raises(KglError):
            ctx.set_genome_superpop(pop.superpop[:-1]); ctx.select_loci(); ctx.count_and_inbreed()   # wrong length: a new population without a matrix
        ctx.upload_population(pop)
        with pytest.raises(KglError):
            ctx.set_locus_selection(np.ones(pop.n_loci - 1, dtype=np.uint8))
        with pytest.raises(KglError):
            ctx.set_locus_filter(np.ones(pop.n_loci + 1, dtype=np.uint8))
        with pytest.raises(KglError):
            ctx.upload_multi_allelic(np.array([5, 5], dtype=np.uint32), np.zeros((6, 2, 3), np.float32), np.zeros((2, 70), np.uint8))   # not ascending
        with pytest.raises(KglError):
            ctx.upload_multi_allelic(np.array([pop.n_loci], dtype=np.uint32), np.zeros((6, 1, 3), np.float32), np.zeros((1, 70), np.uint8))   # beyond the table
        scratch = np.zeros((pop.n_genomes, pop.n_genomes, 4), dtype=np.uint32)
        raw_ibs = lambda a, b: ctx._check(ctx.lib.kgl_b200_run_ibs(ctx.h, C.c_uint64(a), C.c_uint64(b), scratch.ctypes.data_as(C.c_void_p)), "run_ibs")
        for call in (lambda: raw_ibs(5, 3), lambda: ctx.ibs(0, pop.n_genomes + 1), lambda: ctx.binned_genome_counts([0.0], [1.0], pop=6),
                     lambda: ctx.count_loci(pop=9), lambda: ctx.multi_allele_count(1)):
            with pytest.raises(KglError):
                call()
        with pytest.raises((KglError, KeyError, ValueError)):
            ctx.inbreed("NoSuchAlgorithm")
        with pytest.raises(KglError):                                                            # an algorithm id the ABI does not know
            ctx._check(ctx.lib.kgl_b200_inbreed_begin(ctx.h, C.c_int(99), None), "inbreed_begin")
        # an empty selection is not an error: no terms, zero counts
        ctx.select_loci(lower=10**8, upper=10**8 + 5)
        _, res = ctx.count_and_inbreed()
        assert int(res["total_allele_count"].sum()) == 0
        for algo in ("RitlandLocus", "HallME", "Loglikelihood"):
            assert int(ctx.inbreed(algo)["total_allele_count"].sum()) == 0
        # ... and the context still computes
        add_multi_allelic(pop, 40, seed=6)
        ctx.upload_population(pop)
        sel = O.select_all_pops(pop, spacing=15)
        ctx.select_loci(spacing=15)
        _, res = ctx.count_and_inbreed()
        want = O.inbreed(pop, sel, "Simple")
        assert np.array_equal(results_matrix(res)[0], results_matrix(want)[0])
        assert rel_err(results_matrix(res)[1], results_matrix(want)[1], floor=1e-6) < 1e-10
        got = ctx.inbreed("Loglikelihood")["inbred_allele_sum"]
        assert np.max(np.abs(got - O.inbreed(pop, sel, "Loglikelihood")["inbred_allele_sum"])) < 1e-9
        assert np.array_equal(ctx.ibs(), O.ibs(pop))
    finally:
        ctx.close()
