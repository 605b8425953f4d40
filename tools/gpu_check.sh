#!/bin/bash
# One GPU call: the -m gpu tests, then the A/B timings of this round's kernels.  Every step runs under its own
# `timeout` (a hung kernel must cost a step, not the call's whole limit).  usage: tools/gpu_check.sh <out-dir>
out=gpurun_out/${1:-check}
mkdir -p $out
(time timeout 900 python -m pytest tests -m gpu -x -q -o timeout_method=thread --timeout 300) > $out/pytest.log 2>&1
echo "pytest rc=$?"; tail -12 $out/pytest.log
for e in 0.8 0.5 0.0; do
  timeout 120 python tools/exp_mixed.py $e
done 2>&1 | tee $out/mixed.log
(timeout 120 python tools/exp_rollout.py 4 24 10; MDPP_LEAN_ROLLOUT=0 timeout 120 python tools/exp_rollout.py 4 24 10;
 timeout 120 python tools/exp_rollout.py 2 20 20; MDPP_LEAN_ROLLOUT=0 timeout 120 python tools/exp_rollout.py 2 20 20) 2>&1 | tee $out/rollout.log
timeout 900 python bench.py --steps 300 --warmup 5 > $out/bench.json 2> $out/bench.err
echo "bench rc=$?"; tail -3 $out/bench.err
if [ -n "$2" ]; then   # optional: one ncu capture, "<kernel regex> <command...>" after the plain run above succeeded
  shift
  rx=$1; shift
  timeout 300 "$@" > $out/ncu_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s 4 -c 2 -o $out/prof "$@" > $out/ncu.log 2>&1
  echo "ncu rc=$?"
fi
