(timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -5)
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["launch_ms_avg"], d["roofline"]["fp64"]["issue_frac"], d["roofline"]["kernel"], d["config"]["pairs_per_step_per_gpu"])'
echo "n100"; $B --n-sites 10000 --n-ind 100 | python -c "$P"
echo "n100 warp"; NGSLD_EM_PATH=warp $B --n-sites 10000 --n-ind 100 | python -c "$P"
echo "n100 tile"; NGSLD_EM_PATH=tile $B --n-sites 10000 --n-ind 100 | python -c "$P"
echo "n48"; $B --n-sites 10000 --n-ind 48 | python -c "$P"
echo "n128"; $B --n-sites 10000 --n-ind 128 | python -c "$P"
echo "n128 warp"; NGSLD_EM_PATH=warp $B --n-sites 10000 --n-ind 128 | python -c "$P"
echo "n500"; $B | python -c "$P"
