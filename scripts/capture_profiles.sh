#!/bin/bash
# Round profile captures (run on the GPU box through gpurun; outputs under gpurun_out/, summaries copied to profiles/).
#   1. launch list of the bench command (per-launch gpu__time_duration, cold cache, serialised)
#   2. ncu --set full of the dominant kernel (QP stage) and of the other stage kernels of a step, headline config C4
# The raw / source pages are exported on the box and the reports deleted (gpurun brings back at most 64 MiB); the
# report of the headline kernel is kept.  PARTS="bench qp stages small" selects what is captured.
R=${1:-r02}
PARTS=${PARTS:-"bench qp stages small"}
O=gpurun_out
BENCH="python bench.py --quick --steps 2 --warmup 1 --no-cpu-baseline"
export_rep() {          # export_rep <name> [keep]
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1.raw.csv 2>/dev/null
  ncu -i $O/$1.ncu-rep --page source --csv --print-source cuda,sass > $O/$1.src.csv 2>/dev/null
  python scripts/ncu_source_agg.py $O/$1.src.csv 45 > $O/$1.lines.txt 2>/dev/null
  rm -f $O/$1.src.csv
  [ -z "$2" ] && rm -f $O/$1.ncu-rep
}
if [[ $PARTS == *bench* ]]; then
  $BENCH > $O/${R}_bench_plain.json 2> $O/${R}_bench_plain.err || exit 1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/${R}_launches_bench.csv \
      $BENCH > $O/${R}_bench_under_ncu.log 2>&1
fi
if [[ $PARTS == *qp* ]]; then
  for c in C4 C2 C3; do
    python scripts/one_solve.py $c 65536 1 fd > $O/${R}_plain_$c.log 2>&1 || exit 1
    ncu --set full --clock-control none --import-source on -k regex:tg_sqp_qp_kernel -s 10 -c 1 -o $O/${R}_prof_qp_${c} \
        python scripts/one_solve.py $c 65536 1 fd > $O/${R}_ncu_qp_${c}.log 2>&1
    if [ $c == C4 ]; then export_rep ${R}_prof_qp_${c} keep; else export_rep ${R}_prof_qp_${c}; fi
  done
fi
if [[ $PARTS == *stages* ]]; then
  for k in ls fd; do
    for c in C4 C2; do
      ncu --set full --clock-control none --import-source on -k regex:tg_sqp_${k}_kernel -s 10 -c 1 -o $O/${R}_prof_${k}_${c} \
          python scripts/one_solve.py $c 65536 1 fd > $O/${R}_ncu_${k}_${c}.log 2>&1
      export_rep ${R}_prof_${k}_${c}
    done
  done
fi
if [[ $PARTS == *small* ]]; then
  python scripts/prof_small.py C4 65536 > $O/${R}_plain_small.log 2>&1 || exit 1
  ncu --set full --clock-control none --import-source on -k regex:tg_sample_kernel -s 2 -c 1 -o $O/${R}_prof_sample_C4 python scripts/prof_small.py C4 65536 > /dev/null 2>&1
  export_rep ${R}_prof_sample_C4
  for c in C4 C2; do
    ncu --set full --clock-control none --import-source on -k regex:tg_eval_kernel -s 2 -c 1 -o $O/${R}_prof_eval_${c} python scripts/prof_small.py $c 65536 > /dev/null 2>&1
    export_rep ${R}_prof_eval_${c}
  done
fi
ls -la $O/${R}_*
du -sh $O
