B="python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:em_tile -s 1 -c 1 -f -o gpurun_out/r1_prof_em python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --batch-pairs 2000000 > gpurun_out/r1_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pearson -s 1 -c 1 -f -o gpurun_out/r1_prof_pearson python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --batch-pairs 2000000 > gpurun_out/r1_ncu_full2.log 2>&1
for t in 9 6 4; do echo "TILE=$t"; NGSLD_TILE=$t $B | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['launch_ms_avg'], d['roofline']['fp64']['frac'])"; done
echo LIST; NGSLD_EM_PATH=list $B | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['launch_ms_avg'], d['roofline']['fp64']['frac'])"
echo N100; $B --n-sites 10000 --n-ind 100 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['launch_ms_avg'], d['roofline']['fp64'])"
echo STRICT; $B --strict | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['launch_ms_avg'], d['roofline']['fp64'])"
