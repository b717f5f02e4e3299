B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --n-sites 30000 --n-ind 320"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["fp64"]["issue_frac"], d["roofline"]["kernel"])'
echo "n320 R4 16 warps"; NGSLD_WARP_R=4 $B | python -c "$P"
echo "n320 R4 12 warps"; NGSLD_WARP_R=4 NGSLD_WARP_PAD_SMEM=32000 $B | python -c "$P"
echo "n320 R4 8 warps"; NGSLD_WARP_R=4 NGSLD_WARP_PAD_SMEM=70000 $B | python -c "$P"
echo "n320 R6 12 warps"; NGSLD_WARP_R=6 $B | python -c "$P"
echo "n320 R6 8 warps"; NGSLD_WARP_R=6 NGSLD_WARP_PAD_SMEM=70000 $B | python -c "$P"
