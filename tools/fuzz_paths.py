"""Differential fuzz (development tool, GPU): the equivalent implementations of one result against each other on random shapes --
HallME / Loglikelihood from the moment tables (tensor-core builder, CUDA-core builder) against the kernels that evaluate every cell,
pairwise IBS on the tensor cores against the popcount kernel. No oracle involved. Usage: fuzz_paths.py [n_cases] [seed]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kgl_gene_b200.capi import KglB200
from kgl_gene_b200.synth import make_population

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
WIDE = os.environ.get("FUZZ_WIDE") == "1"      # also draws populations of thousands of genomes and hundreds of thousands of loci
ctx = KglB200(0)
worst = {"HallME": 0.0, "Loglikelihood": 0.0}
for case in range(n_cases):
    n = int(rng.choice([1, 2, 33, 64, 65, 127, 130, 256, 300, 511, 700, 1500] + ([2504, 4100] if WIDE else [])))
    l = int(rng.choice([1, 31, 32, 100, 1000, 4097, 20000, 60000] + ([8191, 150000, 400000] if WIDE else [])))
    kw = dict(n_genomes=n, n_loci=l, seed=int(rng.integers(1, 10**6)), spectrum=str(rng.choice(["sfs", "dense"])),
              grouped=bool(rng.integers(0, 2)), unphased=bool(rng.integers(0, 4) == 0),
              missing_rate=float(rng.choice([0.0, 0.001, 0.02])), missing_af_rate=float(rng.choice([0.0, 0.02])))
    pop, _ = make_population(**kw)
    if rng.integers(0, 2):
        pop.af[int(rng.integers(0, 6)), ::max(1, l // 7)] = np.float32(rng.choice([1.0, 0.5, 1e-9, 0.9999, 0.0004]))
    ctx.upload_population(pop)
    sel = dict(spacing=int(rng.choice([0, 0, 7, 50])))
    if rng.integers(0, 2) and l > 10:
        lo, hi = sorted(rng.integers(0, l, size=2).tolist())
        sel.update(lower=int(pop.offsets[lo]), upper=int(pop.offsets[hi]))
    ctx.select_loci(**sel)
    start = rng.uniform(0.0, 1.0, size=n)
    for algo, opts in (("HallME", dict(hall_start=start, hall_sweeps=int(rng.choice([1, 7, 50])))), ("Loglikelihood", {})):
        exact = ctx.inbreed(algo, exact_sweeps=True, **opts)
        ctx.select_loci(**sel)
        fast = ctx.inbreed(algo, **opts)
        path = ctx.used_moment_tables()
        ctx.select_loci(**sel)
        cores = ctx.inbreed(algo, moments_on_cuda_cores=True, **opts)
        a, b, c2 = exact["inbred_allele_sum"], fast["inbred_allele_sum"], cores["inbred_allele_sum"]
        ok = np.isfinite(a)
        assert np.array_equal(np.isnan(a), np.isnan(b)), (case, kw, sel, algo)
        d = float(np.max(np.abs(a[ok] - b[ok]))) if ok.any() else 0.0
        d2 = float(np.max(np.abs(b[ok] - c2[ok]))) if ok.any() else 0.0
        worst[algo] = max(worst[algo], d)
        assert d < 1e-9 and d2 < 1e-12, (case, kw, sel, algo, path, d, d2)
        for f in ("major_homo_count", "major_hetero_count", "minor_homo_count", "minor_hetero_count", "total_allele_count"):
            assert np.array_equal(exact[f], fast[f]), (case, f)
    if n * n * l <= 4e9:
        ctx.set_ibs_tensor_cores(True); t = ctx.ibs(); used = ctx.ibs_used_tensor_cores()
        ctx.set_ibs_tensor_cores(False); p = ctx.ibs()
        ctx.set_ibs_tensor_cores(True)
        assert np.array_equal(t, p), (case, kw, "ibs", used)
    print(case, n, l, kw["spectrum"], "grouped" if kw["grouped"] else "mixed", "unphased" if kw["unphased"] else "phased", sel, "ok", flush=True)
print("all equal; worst fast-vs-exact", worst)
ctx.close()
