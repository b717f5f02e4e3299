(timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5)
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["launch_ms_avg"], d["roofline"]["fp64"]["frac"], d["config"]["pairs_per_step_per_gpu"])'
echo "n500 default"; $B | python -c "$P"
echo "n1000 warp G2"; $B --n-sites 20000 --n-ind 1000 | python -c "$P"
echo "n1000 old"; NGSLD_EM_PATH=tile $B --n-sites 20000 --n-ind 1000 | python -c "$P"
echo "n2000 warp G4"; $B --n-sites 10000 --n-ind 2000 | python -c "$P"
echo "n2000 old"; NGSLD_EM_PATH=tile $B --n-sites 10000 --n-ind 2000 | python -c "$P"
echo "n250 warp"; $B --n-sites 20000 --n-ind 250 | python -c "$P"
echo "n250 old"; NGSLD_EM_PATH=tile $B --n-sites 20000 --n-ind 250 | python -c "$P"
