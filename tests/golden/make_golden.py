#!/usr/bin/env python
"""Regenerates tests/golden/*.npz by running the REFERENCE's own code (oracle/_ref/kgl_ref_harness, built from
/root/reference by `make -C oracle ref`) on small seeded populations. Only runs in the build container (the GPU box has
no /root/reference and does not need it: the fixtures are committed).

Each fixture stores the flattened input (offsets, af, superpop, packed, flags) and everything the reference computed:
locus selection per super-population, LocusResults of all four estimators, logLikelihood(f) on a grid, the optimiser's
start/end points, and the VariantDBVariant allele summaries. HallME / Loglikelihood are made reproducible by pinning
std::random_device inside the harness (oracle/ref_stubs.cpp), never by editing reference code.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import oracle_py as O  # noqa: E402
from kgl_gene_b200.synth import add_multi_allelic, make_population  # noqa: E402

CASES = {
    # name: (make_population kwargs, harness kwargs)
    "sfs_phased": (dict(n_genomes=40, n_loci=1500, seed=101, spectrum="sfs"), dict(seed=7)),
    "dense_spacing_af": (dict(n_genomes=33, n_loci=1200, seed=202, spectrum="dense"),
                         dict(seed=11, spacing=25, min_af=0.1, max_af=0.45, lower=1500, upper=11000)),
    "unphased_pf": (dict(n_genomes=24, n_loci=800, seed=303, spectrum="sfs", unphased=True), dict(seed=3)),
    "missing_af_ragged": (dict(n_genomes=70, n_loci=600, seed=404, spectrum="sfs", missing_af_rate=0.05, missing_rate=0.02,
                               grouped=False), dict(seed=5, min_af=0.01)),
    "rare_major": (dict(n_genomes=20, n_loci=500, seed=505, spectrum="sfs"), dict(seed=9)),   # af pushed towards 1 below
    # loci with two or three alternate alleles (add_multi_allelic below): AlleleFreqVector over several alleles, MINOR_HETEROZYGOUS
    # with two different alleles, the normalised class frequencies of kga_analysis_inbreed_freq.cpp:127-217
    "multi_allelic": (dict(n_genomes=60, n_loci=1400, seed=606, spectrum="sfs", missing_af_rate=0.01), dict(seed=13, spacing=15)),
    "multi_allelic_unphased": (dict(n_genomes=30, n_loci=900, seed=707, spectrum="dense", unphased=True), dict(seed=15)),
}
MULTI = {"multi_allelic": (160, 61), "multi_allelic_unphased": (120, 71)}      # name: (loci made multi-allelic, seed)


def main():
    if not O.have_reference_harness():
        raise SystemExit("oracle/_ref/kgl_ref_harness missing: run `make -C oracle ref` (needs /root/reference)")
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, (pop_kw, ref_kw) in CASES.items():
        pop, inbreeding = make_population(**pop_kw)
        if name == "rare_major":
            # exercise the q <= 0.01 drop rule (kga_analysis_inbreed_freq.cpp:532) and p <= 0.001 (calc.cpp:397)
            pop.af[:, ::7] = np.float32(0.995)
            pop.af[:, 3::11] = np.float32(0.0005)
            pop.af[:, 5::13] = np.float32(1.0)
        if name in MULTI:
            add_multi_allelic(pop, *MULTI[name])
        ref = O.run_reference(pop, grid=21, fws=True, **ref_kw)
        stderr = ref.pop("_stderr")
        arrays = {"in_offsets": pop.offsets, "in_af": pop.af, "in_superpop": pop.superpop, "in_packed": pop.packed,
                  "in_n_genomes": np.array([pop.n_genomes]), "in_unphased": np.array([int(pop.unphased)]),
                  "in_true_inbreeding": inbreeding}
        if pop.n_multi:
            arrays.update(in_multi_rows=pop.multi_rows, in_multi_af=pop.multi_af, in_multi_cells=pop.multi_cells)
        for k in ("spacing", "min_af", "max_af", "lower", "upper", "seed"):
            default = {"spacing": 0, "min_af": 0.0, "max_af": 1.0, "lower": 0, "upper": 10**9, "seed": 0}[k]
            arrays["arg_" + k] = np.array([ref_kw.get(k, default)], dtype=np.float64)
        for k, v in ref.items():
            arrays["ref_" + k] = v
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB  {stderr.strip().splitlines()[0]}")


if __name__ == "__main__":
    main()
