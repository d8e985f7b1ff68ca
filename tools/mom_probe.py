"""Development probe: moment-table build phases (run under ncu for per-kernel times) and HallME endpoint starts."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kgl_gene_b200.capi import KglB200
from kgl_gene_b200.synth import make_genomes, make_loci, make_population

if len(sys.argv) > 1 and sys.argv[1] == "endpoints":
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as O
    pop, _ = make_population(64, 3000, seed=41)
    ctx = KglB200(0)
    ctx.upload_population(pop); ctx.select_loci()
    sel = O.select_all_pops(pop)
    start = np.linspace(0.0, 1.0, pop.n_genomes)
    for sweeps in (1, 2, 50):
        want = O.inbreed(pop, sel, "HallME", start=start, sweeps=sweeps)["inbred_allele_sum"]
        fast = ctx.inbreed("HallME", hall_start=start, hall_sweeps=sweeps)["inbred_allele_sum"]
        exact = ctx.inbreed("HallME", hall_start=start, hall_sweeps=sweeps, exact_sweeps=True)["inbred_allele_sum"]
        print(sweeps, "fast-oracle", np.abs(fast - want).max(), "exact-oracle", np.abs(exact - want).max(), "ends", want[[0, -1]], fast[[0, -1]], exact[[0, -1]])
    sys.exit(0)

n, l = 2504, 1_100_000
ctx = KglB200(0)
offsets, af = make_loci(l, 2)
superpop, f = make_genomes(n, 2)
ctx.upload_loci(af, offsets)
ctx.set_genome_superpop(superpop)
ctx.synth_genotypes(2, n, l, f, missing_rate=0.001)
for rep in range(2):
    for algorithm in ("HallME", "Loglikelihood"):
        ctx.select_loci()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ctx.inbreed(algorithm)
        torch.cuda.synchronize(); print(algorithm, "cold ms", (time.perf_counter() - t0) * 1e3, flush=True)
ctx.close()
