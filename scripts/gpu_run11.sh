# round-1 evidence run: launch list of the bench command, one full ncu capture of the dominant kernel, the bench itself
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r1_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --batch-pairs 8000000 > gpurun_out/r1_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:em_warp -s 2 -c 1 -f -o gpurun_out/r1_prof_warp_r6 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --batch-pairs 4000000 > gpurun_out/r1_ncu_warp_r6.log 2>&1
python bench.py > gpurun_out/r1_bench_final.json 2> gpurun_out/r1_bench_final.err; cat gpurun_out/r1_bench_final.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_reference.json 2>/dev/null; cat gpurun_out/r1_bench_reference.json
