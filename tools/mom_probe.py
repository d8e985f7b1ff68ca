"""Development probe: HallME / Loglikelihood on a fresh selection at BASELINE config 2 (run under ncu for per-kernel times)."""
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
for rep in range(2):
    for algorithm in ("HallME", "Loglikelihood"):
        ctx.select_loci()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ctx.inbreed(algorithm)
        torch.cuda.synchronize(); print(algorithm, "cold ms", (time.perf_counter() - t0) * 1e3, flush=True)
ctx.close()
