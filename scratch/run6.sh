python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/_sweep.py 16 2>&1 | tail -1
PIVB200_SOA_SYNC=0 python tools/_sweep.py 16 2>&1 | tail -1
PIVB200_SOA=0 python tools/_sweep.py 16 2>&1 | tail -1
for m in 1 2 8 16 12 20 36; do PIVB200_SOA_SYNC=$m python tools/_sweep.py 16 2>&1 | tail -1 | sed "s/^/sync=$m: /"; done
