#!/bin/bash
# ncu evidence of round 2 (run on the GPU box, after the same commands have exited 0 without ncu): launch list of the bench step
# and of the estimators, full captures of the kernels DESIGN.md names. Reports land in gpurun_out/; summaries are cut from them
# into profiles/ (tools/ncu_summary.py).
set -x
python bench.py --steps 3 --warmup 3 --no-kinship --no-cpu-baseline > gpurun_out/r02_profile_bench_plain.json 2> gpurun_out/r02_profile_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_bench_final.csv \
    python bench.py --steps 3 --warmup 3 --no-kinship --no-cpu-baseline > /dev/null 2>&1
for k in k_stream_count_ct k_tail k_locus_prepare k_mom_mma k_mom_unit_fill k_mom_run; do
  ncu --set full --import-source on --clock-control none -k regex:$k -s 2 -c 1 -o gpurun_out/r02_ncu_$k \
      python bench.py --steps 3 --warmup 3 --no-kinship --no-cpu-baseline > /dev/null 2>&1
done
ls -la gpurun_out/*.ncu-rep
