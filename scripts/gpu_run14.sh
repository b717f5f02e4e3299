(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4)
python bench.py --steps 3 --warmup 3 --no-cpu-baseline | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["fp64"]["issue_frac"])'
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r1_launches_pe2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --batch-pairs 4000000 > /dev/null 2>&1; grep "pearson\|em_warp" gpurun_out/r1_launches_pe2.csv | awk -F'","' '{print $5, $NF}' | tail -4
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --strict | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("strict", d["value"], d["ms_per_step"])'
