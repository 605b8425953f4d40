#!/bin/bash
# One GPU call of ncu captures (they count as one): launch list of the bench command, --set full of the dominant
# kernels.  Every command first runs plain and must exit 0.  usage: tools/gpu_profile.sh <out-dir>
out=gpurun_out/${1:-prof}
mkdir -p $out
B="python bench.py --steps 20 --warmup 3 --no-aux --cpu-seconds 0"
cap() { # name regex skip count cmd...
  name=$1; rx=$2; skip=$3; cnt=$4; shift 4
  timeout 300 "$@" > $out/${name}_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o $out/$name "$@" > $out/${name}_ncu.log 2>&1
  echo "$name rc=$?"
  # gpurun_out/ may carry 64 MiB back: export the two pages the summaries are made from, drop the report
  if [ -f $out/$name.ncu-rep ]; then
    ncu -i $out/$name.ncu-rep --page raw --csv > $out/${name}_raw.csv 2>/dev/null
    ncu -i $out/$name.ncu-rep --page source --csv --print-source sass > $out/${name}_src.csv 2>/dev/null
    [ "$name" = "pure" ] || rm -f $out/$name.ncu-rep
  fi
}
timeout 300 $B > $out/bench_plain.json 2> $out/bench_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $out/launches_bench.csv $B > $out/launches_ncu.log 2>&1
echo "launch list rc=$?"
cap pure q_rounds_pure 40 2 $B
cap round0 'q_round_kernel' 10 2 python tools/exp_round_parts.py
cap mixed q_rounds_mixed 2 1 python tools/exp_mixed.py 0.5
cap rollout4 rollout_lean 4 2 python tools/exp_rollout.py 4 24 10 0
cap reach reach_bfs 1 1 python tools/exp_reach.py 2
ls -la $out
