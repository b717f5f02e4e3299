ncu --set full --clock-control none --import-source on -k regex:em_list -s 2 -c 1 -f -o gpurun_out/r1_prof_list_n100 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --batch-pairs 4000000 --n-sites 10000 --n-ind 100 > gpurun_out/r1_ncu_list_n100.log 2>&1
tail -1 gpurun_out/r1_ncu_list_n100.log | cut -c1-200
