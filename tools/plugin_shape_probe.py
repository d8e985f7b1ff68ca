"""Development probe: the plugin's call pattern (20 windows x 4 algorithms after one upload), wall clock per call."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kgl_gene_b200.capi import KglB200
from kgl_gene_b200.synth import make_genomes, make_loci

n, l = 2504, 1_100_000
ctx = KglB200(0)
offsets, af = make_loci(l, 2)
superpop, f = make_genomes(n, 2)
ctx.upload_loci(af, offsets)
ctx.set_genome_superpop(superpop)
ctx.synth_genotypes(2, n, l, f, missing_rate=0.001)
per = l // 20
for rep in range(2):
    for algorithm in ("Simple", "HallME", "Loglikelihood"):
        times = []
        for w in range(20):
            t0 = time.perf_counter()
            ctx.select_loci(lower=int(offsets[w * per]), upper=int(offsets[min(l, (w + 1) * per) - 1]))
            ctx.inbreed(algorithm)
            times.append((time.perf_counter() - t0) * 1e3)
        print(rep, algorithm, "total %.1f ms" % sum(times), " ".join("%.1f" % t for t in times), "path", ctx.used_moment_tables(), flush=True)
ctx.close()
