(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2)
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["launch_ms_avg"]*d["roofline"]["launches_timed"]/3, d["roofline"]["fp64"]["issue_frac"])'
echo "pearson128"; $B | python -c "$P"
echo "pearson96"; NGSLD_PEARSON_THREADS=96 $B | python -c "$P"
echo "pearson64"; NGSLD_PEARSON_THREADS=64 $B | python -c "$P"
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/r1_launches_pe3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --batch-pairs 4000000 > /dev/null 2>&1; grep "pearson" gpurun_out/r1_launches_pe3.csv | awk -F'","' '{print $5, $NF}' | tail -2
