(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5)
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["launch_ms_avg"], d["roofline"]["fp64"]["frac"], d["config"]["pairs_per_step_per_gpu"])'
for r in 6 5 4; do echo "WARP R=$r"; NGSLD_WARP_R=$r $B | python -c "$P"; done
echo "R=6 U1"; NGSLD_WARP_U1=1 $B | python -c "$P"
