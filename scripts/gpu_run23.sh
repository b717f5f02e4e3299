ncu --set full --clock-control none --import-source on -k regex:em_warp -s 2 -c 1 -f -o gpurun_out/r1_prof_warp_g2 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --batch-pairs 2000000 --n-sites 20000 --n-ind 1000 > gpurun_out/r1_ncu_warp_g2.log 2>&1
tail -1 gpurun_out/r1_ncu_warp_g2.log | cut -c1-100
