"""Worker of test_run_inbreed_sharded_two_gpus: one rank per GPU, each with a locus shard; shards.run_inbreed_sharded for all
four estimators (on a real torch stream and on the legacy default stream), compared with the oracle on the whole population."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    import oracle_py as O
    from conftest import results_matrix
    from kgl_gene_b200 import shards
    from kgl_gene_b200.capi import KglB200
    from kgl_gene_b200.flatfile import FlatPopulation
    from kgl_gene_b200.synth import make_population

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pop, _ = make_population(300, 60_000, seed=81, missing_rate=0.003)
    sel = O.select_all_pops(pop)
    a, b = shards.locus_shard(pop.n_loci, rank, world)
    sh = FlatPopulation(pop.offsets[a:b], np.ascontiguousarray(pop.af[:, a:b]), pop.superpop, np.ascontiguousarray(pop.packed[a:b]),
                        pop.n_genomes, pop.unphased)
    ctx = KglB200(local)
    ctx.upload_population(sh)
    ctx.select_loci()
    start = np.linspace(0.05, 0.5, pop.n_genomes)
    for use_stream in (True, False):
        stream = torch.cuda.Stream(device=dev) if use_stream else torch.cuda.default_stream(dev)
        with torch.cuda.stream(stream):
            for algo in ("Simple", "RitlandLocus", "HallME", "Loglikelihood"):
                kw = dict(hall_start=start, hall_sweeps=50) if algo == "HallME" else {}
                okw = dict(start=start, sweeps=50) if algo == "HallME" else {}
                got = shards.run_inbreed_sharded(ctx, algo, dev, **kw)
                want = O.inbreed(pop, sel, algo, **okw)
                assert np.array_equal(results_matrix(got)[0], results_matrix(want)[0]), algo
                assert np.max(np.abs(got["inbred_allele_sum"] - want["inbred_allele_sum"])) < 1e-9, (algo, use_stream)
        ctx.set_stream(None)
    ctx.close()
    dist.barrier()
    open(os.path.join(sys.argv[1], f"ok{rank}"), "w").close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
