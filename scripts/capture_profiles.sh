#!/bin/bash
# Round profile captures (run on the GPU box through gpurun; outputs under gpurun_out/, summaries copied to profiles/).
#   1. launch list of the default bench command (per-launch gpu__time_duration, cold cache, serialised)
#   2. ncu --set full of the dominant kernel (QP stage) and of the other kernels of a step
R=${1:-r01}
O=gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${R}_bench_plain.json 2> $O/${R}_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${R}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${R}_bench_under_ncu.log 2>&1
for k in qp ls fd; do
  ncu --set full --clock-control none --import-source on -k regex:tg_sqp_${k}_kernel -s 10 -c 1 -o $O/${R}_prof_${k}_c2 \
      python scripts/one_solve.py C2 65536 1 fd > $O/${R}_ncu_${k}.log 2>&1
done
for c in C3 C4; do
  ncu --set full --clock-control none --import-source on -k regex:tg_sqp_qp_kernel -s 10 -c 1 -o $O/${R}_prof_qp_${c} \
      python scripts/one_solve.py $c 65536 1 fd > $O/${R}_ncu_qp_${c}.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:tg_sample_kernel -s 2 -c 1 -o $O/${R}_prof_sample_c2 python scripts/prof_small.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tg_eval_kernel -s 2 -c 1 -o $O/${R}_prof_eval_c2 python scripts/prof_small.py > /dev/null 2>&1
ls -la $O/${R}_*
