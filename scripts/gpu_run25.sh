(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -2)
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["fp64"]["issue_frac"])'
python scripts/contract_at_scale.py --s1-hi 200
