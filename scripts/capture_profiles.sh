#!/bin/bash
# Round profile captures (run on the GPU box through gpurun; outputs under gpurun_out/, summaries copied to profiles/).
#   1. launch list of the bench command (per-launch gpu__time_duration, cold cache, serialised)
#   2. ncu --set full of the dominant kernel (QP stage) and of the other kernels of a step, headline config C4
R=${1:-r02}
O=gpurun_out
BENCH="python bench.py --quick --steps 2 --warmup 1 --no-cpu-baseline"
$BENCH > $O/${R}_bench_plain.json 2> $O/${R}_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/${R}_launches_bench.csv \
    $BENCH > $O/${R}_bench_under_ncu.log 2>&1
for c in C4 C2 C3; do
  python scripts/one_solve.py $c 65536 1 fd > $O/${R}_plain_$c.log 2>&1 || exit 1
  ncu --set full --clock-control none --import-source on -k regex:tg_sqp_qp_kernel -s 10 -c 1 -o $O/${R}_prof_qp_${c} \
      python scripts/one_solve.py $c 65536 1 fd > $O/${R}_ncu_qp_${c}.log 2>&1
done
for k in ls fd; do
  ncu --set full --clock-control none --import-source on -k regex:tg_sqp_${k}_kernel -s 10 -c 1 -o $O/${R}_prof_${k}_C4 \
      python scripts/one_solve.py C4 65536 1 fd > $O/${R}_ncu_${k}.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:tg_sqp_${k}_kernel -s 10 -c 1 -o $O/${R}_prof_${k}_C2 \
      python scripts/one_solve.py C2 65536 1 fd > $O/${R}_ncu_${k}_c2.log 2>&1
done
python scripts/prof_small.py C4 65536 > $O/${R}_plain_small.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:tg_sample_kernel -s 2 -c 1 -o $O/${R}_prof_sample_C4 python scripts/prof_small.py C4 65536 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tg_eval_kernel -s 2 -c 1 -o $O/${R}_prof_eval_C4 python scripts/prof_small.py C4 65536 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tg_eval_kernel -s 2 -c 1 -o $O/${R}_prof_eval_C2 python scripts/prof_small.py C2 65536 > /dev/null 2>&1
ls -la $O/${R}_*
