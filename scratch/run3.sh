for cfg in "0 4" "63 0" "21 0" "4 0" "42 0" "1 0" "5 0"; do set -- $cfg; PIVB200_SOA_SYNC=$1 PIVB200_SOA_GROUP=$2 python tools/_sweep.py 16 2>&1 | tail -1 | sed "s/^/sync=$1 group=$2: /"; done
PIVB200_SOA64=1 python -m pytest tests -m gpu -x -q -k "not pass_by_pass and not chained_plan_explained and not seeded_stress" 2>&1 | tail -5
PIVB200_SOA64=1 python tools/_sweep.py 16 2>&1 | tail -1
PIVB200_SOA64=1 PIVB200_SOA_SYNC=4 PIVB200_SOA_GROUP=4 python tools/_sweep.py 16 2>&1 | tail -1
PIVB200_SOA64=1 PIVB200_SOA_SYNC=63 PIVB200_SOA_GROUP=0 python tools/_sweep.py 16 2>&1 | tail -1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "pass_by_pass or chained_plan_explained or seeded_stress" 2>&1 | tail -30
