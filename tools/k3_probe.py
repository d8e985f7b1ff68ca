"""Times the estimator passes that sit behind the streaming pass (development tool): RitlandLocus (one sparse sweep),
HallME (50 EM sweeps), Loglikelihood (bracketed Newton) on the BASELINE config-2 shape, resident matrix.
Prints one JSON line per algorithm: passes, ms per pass, genotype-loci/s per pass and for the whole estimator."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kgl_gene_b200.capi import KglB200
from kgl_gene_b200.synth import make_genomes, make_loci

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2504
l = int(sys.argv[2]) if len(sys.argv) > 2 else 1_100_000
algos = sys.argv[3].split(",") if len(sys.argv) > 3 else ["Simple", "RitlandLocus", "HallME", "Loglikelihood"]
dev = torch.device("cuda", 0)
ctx = KglB200(0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
offsets, af = make_loci(l, 2)
superpop, f = make_genomes(n, 2)
ctx.upload_loci(af, offsets)
ctx.set_genome_superpop(superpop)
ctx.synth_genotypes(2, n, l, f, missing_rate=0.001)
ctx.select_loci()
torch.cuda.synchronize()


def run(algorithm):
    ev = []
    ctx.inbreed_begin(algorithm)
    finished = False
    t_all0 = torch.cuda.Event(enable_timing=True); t_all1 = torch.cuda.Event(enable_timing=True)
    t_all0.record(stream)
    while not finished:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.inbreed_accumulate()
        e1.record(stream)
        finished = ctx.inbreed_update()
        ev.append((e0, e1))
    t_all1.record(stream)
    res = ctx.inbreed_fetch()
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in ev]
    return res, ms, t_all0.elapsed_time(t_all1)


for algorithm in algos:
    run(algorithm)                       # warm-up: sample-major copy, buffers
    res, ms, total = run(algorithm)
    sweeps = ms[1:] if len(ms) > 1 else ms
    per = float(np.median(sweeps))
    err = float(np.max(np.abs(res["inbred_allele_sum"] - f)))
    print(json.dumps({"algorithm": algorithm, "workload": f"{n} x {l}", "passes": len(ms), "first_pass_ms": ms[0], "ms_per_sweep": per,
                      "total_ms": total, "genotype_loci_per_s_per_sweep": n * l / (per * 1e-3),
                      "genotype_loci_per_s_estimator": n * l / (total * 1e-3), "max_abs_dev_from_planted_F": err,
                      "ms_per_pass": [round(m, 3) for m in ms] if algorithm == "Loglikelihood" else None}), flush=True)
ctx.close()
