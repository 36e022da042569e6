python -m pytest tests -q -m gpu -x -k "explicit or omega" 2>&1 | tail -2
