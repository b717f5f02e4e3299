(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2)
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r1_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --batch-pairs 8000000 > gpurun_out/r1_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:em_warp -s 2 -c 1 -f -o gpurun_out/r1_prof_warp_final python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --batch-pairs 4000000 > gpurun_out/r1_ncu_warp_final.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_final_reference.json 2>/dev/null
python bench.py > gpurun_out/r1_final_bench.json 2> gpurun_out/r1_final_bench.err; cat gpurun_out/r1_final_bench.json | cut -c1-400
