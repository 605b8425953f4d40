#!/bin/bash
# One GPU call: ncu --set full captures of the scan-mode kernels (4 cars, 2^20 envs).  The command first runs plain
# and must exit 0.  usage: tools/gpu_profile_scan.sh <out-dir>
out=gpurun_out/${1:-profscan}
mkdir -p $out
C="python tools/exp_scan.py 4 20 8 64"
timeout 300 $C > $out/plain.log 2>&1 || { echo "plain run failed"; tail -5 $out/plain.log; exit 1; }
cap() { # name regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o $out/$1 $C > $out/$1_ncu.log 2>&1
  echo "$1 rc=$?"
  if [ -f $out/$1.ncu-rep ]; then
    ncu -i $out/$1.ncu-rep --page raw --csv > $out/$1_raw.csv 2>/dev/null
    ncu -i $out/$1.ncu-rep --page source --csv --print-source sass > $out/$1_src.csv 2>/dev/null
    rm -f $out/$1.ncu-rep
  fi
}
cap purescan q_rounds_pure_scan 20 2
cap pullscan q_pull_scan 100 3
cap envsub 'q_env_kernel' 8 4
ls -la $out
