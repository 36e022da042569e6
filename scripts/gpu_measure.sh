python -m pytest tests -q -m gpu -x -k "fetch_narrow or symm or assembled" 2>&1 | tail -2
python scripts/e2e_breakdown.py 1000 2>&1 | grep "e2e_step\|fetch rowval" | head -5
