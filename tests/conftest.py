import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["sfs_phased", "dense_spacing_af", "unphased_pf", "missing_af_ragged", "rare_major", "multi_allelic", "multi_allelic_unphased"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    """Returns (FlatPopulation, dict of reference arrays, selection kwargs)."""
    from kgl_gene_b200.flatfile import FlatPopulation
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    pop = FlatPopulation(z["in_offsets"], z["in_af"], z["in_superpop"], z["in_packed"], int(z["in_n_genomes"][0]),
                         bool(z["in_unphased"][0]))
    if "in_multi_rows" in z.files:
        pop.multi_rows, pop.multi_af, pop.multi_cells = z["in_multi_rows"], z["in_multi_af"], z["in_multi_cells"]
    ref = {k[4:]: z[k] for k in z.files if k.startswith("ref_")}
    sel_kw = dict(spacing=int(z["arg_spacing"][0]), min_af=float(z["arg_min_af"][0]), max_af=float(z["arg_max_af"][0]),
                  lower=int(z["arg_lower"][0]), upper=int(z["arg_upper"][0]))
    return pop, ref, sel_kw


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return (request.param,) + load_golden(request.param)


def results_matrix(r):
    """LocusResults structured array -> (counts [N,5] majHom,majHet,minHom,minHet,total ; freqs [N,4] same order)."""
    counts = np.stack([r["major_homo_count"], r["major_hetero_count"], r["minor_homo_count"], r["minor_hetero_count"],
                       r["total_allele_count"]], axis=1).astype(np.uint64)
    freqs = np.stack([r["major_homo_freq"], r["major_hetero_freq"], r["minor_homo_freq"], r["minor_hetero_freq"]], axis=1)
    return counts, freqs
