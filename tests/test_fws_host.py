"""CPU tests of the host-side Pf7 statistics (kgl_gene_b200/fws.py) and of the oracle's CalcFWS restatement."""
import numpy as np

import oracle_py as O
from kgl_gene_b200 import fws
from kgl_gene_b200.synth import make_population


def test_bins_are_the_references():
    assert len(fws.FWS_BINS) == 11 and fws.FWS_BINS[0] == (0.0, 0.05) and fws.FWS_BINS[-1] == (0.5, 1.0)
    for (a, b), (c, d) in zip(fws.FWS_BINS, fws.FWS_BINS[1:]):
        assert b == c and a < b


def test_oracle_fws_bins_partition_present_variants():
    pop, _ = make_population(60, 2000, seed=3, missing_rate=0.01, missing_af_rate=0.05)
    counts, rows = O.fws_bins(pop, 5, fws.FWS_BINS)
    codes = pop.codes()
    af = pop.af[5]
    present = ((codes == 1) | (codes == 2)).any(axis=1) & ~np.isnan(af) & (af < 1.0)
    assert int(rows.sum()) == int(present.sum())
    assert np.array_equal(counts.sum(axis=(0, 2))[:], np.full(60, present.sum(), dtype=np.uint64))
    # against the whole-matrix counts of the oracle's VariantDBVariant restatement
    _, gc = O.allele_count(pop)
    tot = counts.sum(axis=0)
    assert np.all(tot[:, 1] <= gc[:, 1]) and np.all(tot[:, 2] <= gc[:, 2])


def test_hetero_homo_and_fis():
    gc = np.array([[10, 4, 1, 0], [8, 6, 1, 0], [12, 0, 0, 3], [5, 5, 5, 0]], dtype=np.uint64)
    s = fws.hetero_homo_summary(gc)
    assert s["total_variants"].tolist() == [6, 8, 3, 15]          # a code-3 cell is one entry of another allele
    assert s["heterozygous_reference_minor_alleles"].tolist() == [4, 6, 3, 5]
    s0 = fws.hetero_homo_summary(gc, other_allele_entries=0)       # ... or nothing at all (a missing call)
    assert s0["total_variants"].tolist() == [6, 8, 0, 15] and s0["heterozygous_reference_minor_alleles"].tolist() == [4, 6, 0, 5]
    s = s0
    assert s["homozygous_minor_alleles"].tolist() == [1, 1, 0, 5]
    fis = fws.wrights_fis(s, np.array([0, 0, 0, 1]))
    h_exp = 10 / 14
    assert np.allclose(fis[:2], [(h_exp - 4 / 6) / h_exp, (h_exp - 6 / 8) / h_exp])
    assert fis[2] == 0.0 and fis[3] == 0.0          # no variants -> 0 (:396); a one-genome group has H_obs = H_exp
