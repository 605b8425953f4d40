#!/bin/bash
# Multi-GPU call (gpurun --gpus N): the replica-exchange tests (incl. one process per GPU over CUDA IPC) and the bench
# under torchrun.  usage: tools/gpu_check_multi.sh <out-dir> <N>
out=gpurun_out/${1:-multi}; N=${2:-2}
mkdir -p $out
nvidia-smi -L | head -8
(time timeout 600 python -m pytest tests/test_gpu_sync.py -x -q -m gpu -o timeout_method=thread --timeout 300) > $out/pytest_sync.log 2>&1
echo "pytest rc=$?"; tail -6 $out/pytest_sync.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 40 --warmup 5 > $out/bench_${N}gpu.json 2> $out/bench_${N}gpu.err
echo "bench rc=$?"; tail -5 $out/bench_${N}gpu.err; head -c 3000 $out/bench_${N}gpu.json
