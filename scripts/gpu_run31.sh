B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["fp64"]["issue_frac"], d["roofline"]["kernel"])'
for r in 7 5 8; do echo "R=$r"; NGSLD_WARP_R=$r $B | python -c "$P"; done
echo "R=6 U1"; NGSLD_WARP_U1=1 $B | python -c "$P"
