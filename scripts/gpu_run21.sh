(timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -4)
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["fp64"]["issue_frac"], d["roofline"]["kernel"])'
echo n100; $B --n-sites 10000 --n-ind 100 | python -c "$P"
echo n48; $B --n-sites 10000 --n-ind 48 | python -c "$P"
echo n24; $B --n-sites 10000 --n-ind 24 | python -c "$P"
echo n500; $B | python -c "$P"
