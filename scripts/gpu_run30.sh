B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["launch_ms_avg"]*d["roofline"]["launches_timed"]/3, d["roofline"]["fp64"]["issue_frac"])'
echo "160 pearson64"; NGSLD_PEARSON_THREADS=64 NGSLD_B200_LIB=$PWD/ngsld_b200/libexp_mr160.so $B | python -c "$P"
echo "160 pearson32"; NGSLD_PEARSON_THREADS=32 NGSLD_B200_LIB=$PWD/ngsld_b200/libexp_mr160.so $B | python -c "$P"
echo "160 pearson64 n1000"; NGSLD_PEARSON_THREADS=64 NGSLD_B200_LIB=$PWD/ngsld_b200/libexp_mr160.so $B --n-sites 20000 --n-ind 1000 | python -c "$P"
echo "152 pearson32"; NGSLD_PEARSON_THREADS=32 $B | python -c "$P"
