timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scratch/facade_e2e.py 2>&1 | tail -26
