B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["launch_ms_avg"]*d["roofline"]["launches_timed"]/3, d["roofline"]["fp64"]["issue_frac"])'
echo "maxnreg 144"; $B | python -c "$P"
echo "maxnreg 152"; NGSLD_B200_LIB=$PWD/ngsld_b200/libexp_mr152.so $B | python -c "$P"
echo "maxnreg 136"; NGSLD_B200_LIB=$PWD/ngsld_b200/libexp_mr136.so $B | python -c "$P"
