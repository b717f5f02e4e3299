# round-end style verification: GPU tests, smoke, reference arm, bench
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_final_reference.json 2>/dev/null; cut -c1-260 gpurun_out/r1_final_reference.json
python bench.py > gpurun_out/r1_final_bench.json 2> gpurun_out/r1_final_bench.err; cat gpurun_out/r1_final_bench.json; tail -2 gpurun_out/r1_final_bench.err
