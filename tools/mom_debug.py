"""Development probe: which of the equivalent HallME implementations disagrees with the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as O
from kgl_gene_b200.capi import KglB200
from kgl_gene_b200.synth import make_population
for (n, l) in ((30, 2000), (300, 20000)):
    pop, _ = make_population(n, l, seed=41)
    ctx = KglB200(0)
    ctx.upload_population(pop); ctx.select_loci()
    sel = O.select_all_pops(pop)
    want = O.inbreed(pop, sel, "HallME", sweeps=50)["inbred_allele_sum"]
    for name, kw in (("mma+run", {}), ("mma+sweeps", dict(sweep_by_sweep=True)), ("cores+run", dict(moments_on_cuda_cores=True)),
                     ("cores+sweeps", dict(moments_on_cuda_cores=True, sweep_by_sweep=True)), ("exact", dict(exact_sweeps=True))):
        ctx.select_loci()
        got = ctx.inbreed("HallME", hall_sweeps=50, **kw)["inbred_allele_sum"]
        print(n, l, name, ctx.used_moment_tables(), float(np.nanmax(np.abs(got - want))), flush=True)
    want = O.inbreed(pop, sel, "Loglikelihood")["inbred_allele_sum"]
    for name, kw in (("mma+run", {}), ("mma+sweeps", dict(sweep_by_sweep=True)), ("cores+run", dict(moments_on_cuda_cores=True))):
        ctx.select_loci()
        got = ctx.inbreed("Loglikelihood", **kw)["inbred_allele_sum"]
        print(n, l, "LL", name, ctx.used_moment_tables(), float(np.nanmax(np.abs(got - want))), flush=True)
    ctx.close()
