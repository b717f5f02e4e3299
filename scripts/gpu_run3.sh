# GPU run 3: parity tests with the warp-per-pair kernel, then a register-depth sweep
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r1_pytest_gpu3.log; tail -5 gpurun_out/r1_pytest_gpu3.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["roofline"]["launch_ms_avg"], d["roofline"]["fp64"]["frac"], d["config"]["pairs_per_step_per_gpu"])'
for r in 8 7 6 5 4; do echo "WARP R=$r"; NGSLD_WARP_R=$r $B | python -c "$P"; done
echo "OLD tile"; NGSLD_EM_PATH=tile $B | python -c "$P"
echo "N100 default"; $B --n-sites 10000 --n-ind 100 | python -c "$P"
for r in 4 3 2; do echo "N100 warp R=$r"; NGSLD_EM_PATH=warp NGSLD_WARP_R=$r $B --n-sites 10000 --n-ind 100 | python -c "$P"; done
