(timeout 1200 python -m pytest tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -8)
free -g | head -2
echo "config 4 scaled: 100k x 1000 banded 500 kb"; python scripts/run_config.py --n-sites 100000 --n-ind 1000 --max-kb-dist 500
echo "config 5 scaled: 60k x 2000 rnd 0.01"; python scripts/run_config.py --n-sites 60000 --n-ind 2000 --rnd-sample 0.01 --seed 1
echo "config 2: 10k x 100 all pairs rows"; python scripts/run_config.py --n-sites 10000 --n-ind 100 --mode rows
echo "config 3 slice tsv: 50k x 500, first 400 first-sites, tsv"; python scripts/run_config.py --n-sites 50000 --n-ind 500 --s1-hi 400 --mode tsv --data-seed 11
echo "config 3 slice rows"; python scripts/run_config.py --n-sites 50000 --n-ind 500 --s1-hi 400 --mode rows --data-seed 11
