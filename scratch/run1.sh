set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/_sweep.py 16 2>&1 | tail -1
PIVB200_SOA=0 python tools/_sweep.py 16 2>&1 | tail -1
PIVB200_SOA_SYNC=1 python tools/_sweep.py 16 2>&1 | tail -1
PIVB200_NWARPS=16 python tools/_sweep.py 16 2>&1 | tail -1
PIVB200_NWARPS=12 python tools/_sweep.py 16 2>&1 | tail -1
