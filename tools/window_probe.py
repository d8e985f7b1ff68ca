"""Where the time of one plugin window goes (development tool): select_loci and run_inbreed per algorithm on 55,000-locus windows of
the BASELINE config-2 shape, host wall clock per call."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kgl_gene_b200.capi import KglB200
from kgl_gene_b200.synth import make_genomes, make_loci

n, l = 2504, 1_100_000
dev = torch.device("cuda", 0)
ctx = KglB200(0)
offsets, af = make_loci(l, 2)
superpop, f = make_genomes(n, 2)
ctx.upload_loci(af, offsets)
ctx.set_genome_superpop(superpop)
ctx.synth_genotypes(2, n, l, f, missing_rate=0.001)
per = l // 20
def wall(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
w = 7
lo, hi = int(offsets[w * per]), int(offsets[(w + 1) * per - 1])
out = {"select_window_ms": wall(lambda: ctx.select_loci(lower=lo, upper=hi))}
def cold(algorithm, **kw):        # a fresh selection: preparation and moment tables are rebuilt
    ctx.select_loci(**kw)
    ctx.inbreed(algorithm)
for algorithm in ("Simple", "RitlandLocus", "HallME", "Loglikelihood"):
    out[algorithm + "_window_ms"] = wall(lambda: ctx.inbreed(algorithm))
    out[algorithm + "_window_with_select_ms"] = wall(lambda: cold(algorithm, lower=lo, upper=hi))
out["select_all_ms"] = wall(lambda: ctx.select_loci())
for algorithm in ("Simple", "RitlandLocus", "HallME", "Loglikelihood"):
    out[algorithm + "_all_ms"] = wall(lambda: ctx.inbreed(algorithm), 3)
    out[algorithm + "_all_with_select_ms"] = wall(lambda: cold(algorithm), 3)
    if algorithm in ("HallME", "Loglikelihood"):
        out[algorithm + "_all_exact_sweeps_ms"] = wall(lambda: ctx.inbreed(algorithm, exact_sweeps=True), 2)
print(json.dumps(out), flush=True)
ctx.close()
