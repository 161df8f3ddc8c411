set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tests/_sweep.py 16 2>&1 | tail -1
PIVB200_SOA=0 python tests/_sweep.py 16 2>&1 | tail -1
PIVB200_SOA_SYNC=1 python tests/_sweep.py 16 2>&1 | tail -1
PIVB200_NWARPS=16 python tests/_sweep.py 16 2>&1 | tail -1
PIVB200_NWARPS=12 python tests/_sweep.py 16 2>&1 | tail -1
