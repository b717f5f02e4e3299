(timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5)
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["roofline"]["launch_ms_avg"], d["roofline"]["fp64"]["frac"], d["config"]["pairs_per_step_per_gpu"])'
for r in 8 7 6 5; do echo "WARP R=$r U2"; NGSLD_WARP_R=$r $B | python -c "$P"; done
for r in 8 6; do echo "WARP R=$r U1"; NGSLD_WARP_U1=1 NGSLD_WARP_R=$r $B | python -c "$P"; done
