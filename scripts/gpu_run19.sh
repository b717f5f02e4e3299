B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["fp64"]["issue_frac"], d["roofline"]["kernel"])'
echo "orig"; $B | python -c "$P"
echo "swap (tail first)"; NGSLD_B200_LIB=$PWD/ngsld_b200/libexp_swap.so $B | python -c "$P"
echo "pipelined tail loads"; NGSLD_B200_LIB=$PWD/ngsld_b200/libexp_pipe.so $B | python -c "$P"
echo "orig n1000"; $B --n-sites 20000 --n-ind 1000 | python -c "$P"
echo "pipe n1000"; NGSLD_B200_LIB=$PWD/ngsld_b200/libexp_pipe.so $B --n-sites 20000 --n-ind 1000 | python -c "$P"
