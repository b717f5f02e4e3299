(timeout 900 python -m pytest tests/test_cli.py -m gpu -x -q 2>&1 | tail -3)
bash scripts/full_parity_config2.sh 10000 2>&1 | grep -v "^$"
