python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
S="compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5"
echo "== memcheck: warp kernel G=1 (n 500), G=2 (n 1000), group kernel (n 100), strict, tsv, sampling"
$S python scripts/run_config.py --n-sites 300 --n-ind 500 --mode tsv 2>&1 | tail -3 | cut -c1-300
$S python scripts/run_config.py --n-sites 200 --n-ind 1001 --mode rows 2>&1 | tail -3 | cut -c1-300
$S python scripts/run_config.py --n-sites 200 --n-ind 2047 --mode rows --rnd-sample 0.3 2>&1 | tail -3 | cut -c1-300
$S python scripts/run_config.py --n-sites 400 --n-ind 100 --mode rows --max-kb-dist 30 2>&1 | tail -3 | cut -c1-300
$S python scripts/run_config.py --n-sites 300 --n-ind 333 --mode rows --strict 2>&1 | tail -3 | cut -c1-300
echo "== racecheck: warp kernel G=2"
compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 5 python scripts/run_config.py --n-sites 120 --n-ind 1001 --mode rows 2>&1 | tail -4 | cut -c1-300
