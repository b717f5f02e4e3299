(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2)
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["fp64"]["issue_frac"], d["roofline"]["kernel"])'
echo n500; $B | python -c "$P"
echo n100; $B --n-sites 10000 --n-ind 100 | python -c "$P"
echo n1000; $B --n-sites 20000 --n-ind 1000 | python -c "$P"
echo n2000; $B --n-sites 10000 --n-ind 2000 | python -c "$P"
