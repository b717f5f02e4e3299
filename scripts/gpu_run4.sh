ncu --set full --clock-control none --import-source on -k regex:em_warp -s 1 -c 1 -f -o gpurun_out/r1_prof_warp python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --batch-pairs 2000000 > gpurun_out/r1_ncu_warp.log 2>&1
tail -2 gpurun_out/r1_ncu_warp.log | cut -c1-300
