(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5)
python bench.py --steps 3 --warmup 3 --no-cpu-baseline | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["e2e"], d["roofline"]["traffic"], d["roofline"]["algorithmic_bytes_per_launch"])'
