timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
python scripts/cfg4_probe.py 2>&1 | grep -i "^z "
