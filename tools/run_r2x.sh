timeout 300 python -m pytest tests -m gpu -x -q -k "nan_and_inf" 2>&1 | tail -15
