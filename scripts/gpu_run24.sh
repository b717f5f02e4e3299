(timeout 900 python -m pytest tests/test_cli.py -m gpu -x -q 2>&1 | tail -2)
bash scripts/full_config3_cli.sh 2>&1 | tail -2
