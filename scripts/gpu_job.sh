#!/bin/bash
# One parametrised GPU-box job (replaces the one-off scripts of round 1).  Run under gpurun from the repo root:
#   gpurun --timeout 1500 -- bash scripts/gpu_job.sh probe tests bench sweep ncu
# Every stage writes its log under gpurun_out/ (merged back by gpurun); copy what should be judged into profiles/.
set -u
OUT=gpurun_out
mkdir -p $OUT
TAG=${TAG:-r2}
for stage in "$@"; do
  echo "=== stage $stage ($(date +%T)) ==="
  case $stage in
    probe)
      { nproc; free -g; df -h /dev/shm /tmp; nvidia-smi -L; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv; lscpu | head -20; } > $OUT/${TAG}_probe.log 2>&1
      cat $OUT/${TAG}_probe.log | head -40 ;;
    smoke)
      timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 $OUT/${TAG}_smoke.log ;;
    sanitize)
      timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $OUT/${TAG}_sanitizer_memcheck.log
      timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 $OUT/${TAG}_sanitizer_racecheck.log ;;
    tests)
      timeout 2400 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/${TAG}_pytest_gpu.log ;;
    tests_all)
      timeout 2400 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 $OUT/${TAG}_pytest_gpu.log ;;
    bench)
      timeout 900 python bench.py > $OUT/${TAG}_bench_1gpu.json 2> $OUT/${TAG}_bench_1gpu.err; echo "bench rc=$?"; cat $OUT/${TAG}_bench_1gpu.json; tail -3 $OUT/${TAG}_bench_1gpu.err ;;
    bench_dense)
      timeout 900 python bench.py --em-path warp --no-cpu-baseline > $OUT/${TAG}_bench_1gpu_dense.json 2> $OUT/${TAG}_bench_1gpu_dense.err; echo "bench rc=$?"; cat $OUT/${TAG}_bench_1gpu_dense.json ;;
    bench_ref)
      timeout 900 python bench.py --impl reference > $OUT/${TAG}_bench_reference_arm.json 2>&1; cat $OUT/${TAG}_bench_reference_arm.json ;;
    sweep)
      timeout 900 python scripts/sweep.py ${SWEEP_ARGS:-} > $OUT/${TAG}_sweep.log 2>&1; echo "sweep rc=$?"; cat $OUT/${TAG}_sweep.log ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_bench_cmd.csv \
        python bench.py --steps 2 --warmup 3 --batch-pairs 8000000 --no-cpu-baseline --no-e2e > $OUT/${TAG}_launches_bench.log 2>&1; echo "launches rc=$?"
      grep -v "^==" $OUT/${TAG}_launches_bench_cmd.csv | awk -F'","' 'NR>1{n[$5]++; t[$5]+=$NF+0} END{for(k in n) print n[k], t[k], k}' | sort -k2 -n -r | head -12 ;;
    ncu)
      timeout 1200 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-em_cell_kernel} -s ${NCU_SKIP:-1} -c 1 -f -o $OUT/${TAG}_${NCU_NAME:-em_cell} \
        python scripts/sweep.py --pairs 4000000 --reps 2 ${NCU_ARGS:-} > $OUT/${TAG}_ncu_${NCU_NAME:-em_cell}.log 2>&1; echo "ncu rc=$?"; tail -3 $OUT/${TAG}_ncu_${NCU_NAME:-em_cell}.log ;;
    ab)
      # A/B of alternative builds of the library (make -C ngsld_b200/csrc ALT=_x ALTFLAGS=... lib): parity tests on each
      # alternative build, then the same sweep on the default build and on the alternatives
      for alt in ${AB_PYTEST_LIBS-${AB_LIBS:-_x}}; do
        NGSLD_B200_LIB=$PWD/ngsld_b200/libngsld_b200${alt}.so timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q \
          > $OUT/${TAG}_ab_pytest${alt}.log 2>&1; echo "ab pytest ${alt} rc=$?"; tail -4 $OUT/${TAG}_ab_pytest${alt}.log
      done
      for alt in "" ${AB_LIBS:-_x}; do
        for set in ${AB_SETS:-500:50000}; do  # n_ind:n_sites
          nind=${set%%:*}; nsites=${set##*:}
          NGSLD_B200_LIB=$PWD/ngsld_b200/libngsld_b200${alt}.so timeout 600 python scripts/sweep.py --n-ind $nind --n-sites $nsites ${AB_ARGS:-} \
            > $OUT/${TAG}_ab_sweep${alt}_n${nind}.log 2>&1; echo "ab sweep lib${alt} n_ind=$nind rc=$?"; cat $OUT/${TAG}_ab_sweep${alt}_n${nind}.log | cut -c1-330
        done
      done ;;
    contract)
      # numerics contract at scale: fast kernel vs bit-faithful kernel on ~40 M pairs of the bench workload
      timeout 1200 python scripts/contract_at_scale.py > $OUT/${TAG}_contract_at_scale.log 2>&1; echo "contract rc=$?"; tail -4 $OUT/${TAG}_contract_at_scale.log | cut -c1-600 ;;
    configs)
      # BASELINE configs 2, 4, 5 scaled to one GPU (results left in HBM): pairs/s per configuration
      { timeout 600 python scripts/run_config.py --n-sites 10000 --n-ind 100;
        timeout 900 python scripts/run_config.py --n-sites 100000 --n-ind 1000 --max-kb-dist 500;
        timeout 900 python scripts/run_config.py --n-sites 60000 --n-ind 2000 --rnd-sample 0.01 --seed 1; } > $OUT/${TAG}_scaled_configs.log 2>&1
      echo "configs rc=$?"; grep -v "^$" $OUT/${TAG}_scaled_configs.log | cut -c1-400 | tail -12 ;;
    parity2)
      # BASELINE config 2 at full size: unmodified reference vs the CLI (strict: md5 of the sorted outputs; fast: contract)
      timeout 2000 bash scripts/full_parity_config2.sh > $OUT/${TAG}_full_parity_config2.log 2>&1; echo "parity2 rc=$?"; cat $OUT/${TAG}_full_parity_config2.log | cut -c1-300 ;;
    *) echo "unknown stage $stage" ;;
  esac
done
echo "=== done ($(date +%T)) ==="
