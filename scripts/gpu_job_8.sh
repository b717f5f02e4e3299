#!/bin/bash
# The multi-GPU session of a round (one box, N GPUs; charged N x): bench under torchrun (weak-scaling steps + whole-
# workload time to solution through the ABI and the CLI), then BASELINE configs 4 and 5 at their stated sizes.
#   gpurun --gpus 8 --timeout 1500 -- bash scripts/gpu_job_8.sh 8
set -u
N=${1:-8}
OUT=gpurun_out; mkdir -p $OUT
{ nproc; free -g | head -2; df -h /dev/shm | tail -1; nvidia-smi -L | wc -l; nvidia-smi topo -m | head -12; } > $OUT/r2_probe_${N}gpu.log 2>&1
head -6 $OUT/r2_probe_${N}gpu.log
echo "=== bench, $N ranks ($(date +%T)) ==="
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 5 --warmup 3 \
  > $OUT/r2_bench_${N}gpu.json 2> $OUT/r2_bench_${N}gpu.err; echo "rc=$?"; cat $OUT/r2_bench_${N}gpu.json | cut -c1-6000; tail -3 $OUT/r2_bench_${N}gpu.err
for n in ${EXTRA_N:-}; do
  echo "=== bench, $n ranks ($(date +%T)) ==="
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus $n --steps 3 --warmup 3 --no-e2e \
    > $OUT/r2_bench_${n}gpu.json 2> $OUT/r2_bench_${n}gpu.err; echo "rc=$?"; cat $OUT/r2_bench_${n}gpu.json | cut -c1-6000
done
echo "=== config 4 ($(date +%T)) ==="
KEEP=1 bash scripts/full_configs.sh cfg4 $N > $OUT/r2_cfg4_${N}gpu.log 2>&1; cat $OUT/r2_cfg4_${N}gpu.log
echo "--- config 4 again, every GPU fed from the host instead of GPU to GPU"
NGSLD_CLI_HOST_UPLOAD=1 NO_CHECK=1 bash scripts/full_configs.sh cfg4 $N 2>&1 | grep -E "^\[time|process start" | tee -a $OUT/r2_cfg4_${N}gpu.log
echo "=== config 5 ($(date +%T)) ==="
bash scripts/full_configs.sh cfg5 $N > $OUT/r2_cfg5_${N}gpu.log 2>&1; cat $OUT/r2_cfg5_${N}gpu.log
echo "=== done ($(date +%T)) ==="
