(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4)
python bench.py --steps 3 --warmup 3 --no-cpu-baseline | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["fp64"]["issue_frac"])'
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --n-sites 10000 --n-ind 100 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("n100", d["value"], d["ms_per_step"])'
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --strict | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("strict", d["value"], d["ms_per_step"])'
