"""Times kgl_b200_select_loci on the BASELINE config-2 tables (development tool): dense (spacing 0) and spaced windows."""
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
for kw in (dict(spacing=0), dict(spacing=1000), dict(spacing=25), dict(spacing=1000, lower=2_000_000, upper=2_010_000)):
    ctx.select_loci(**kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        counts = ctx.select_loci(**kw)
    torch.cuda.synchronize()
    print(json.dumps({"args": kw, "ms_per_call_incl_count_readback": (time.perf_counter() - t0) / reps * 1e3, "selected_ALL": int(counts[5])}), flush=True)
ctx.close()
