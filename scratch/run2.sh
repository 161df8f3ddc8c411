set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/_sweep.py 16 2>&1 | tail -1
for cfg in "4 4" "63 4" "63 20" "63 10" "21 4" "4 20" "8 20" "1 20" "36 10"; do set -- $cfg; PIVB200_SOA_SYNC=$1 PIVB200_SOA_GROUP=$2 python tools/_sweep.py 16 2>&1 | tail -1 | sed "s/^/sync=$1 group=$2: /"; done
ncu --set full --clock-control none --import-source on -k regex:"piv_soa" -c 1 -o gpurun_out/r02b_soa32 -f python tools/_prof.py 8 > gpurun_out/ncu_r02b.log 2>&1
ncu -i gpurun_out/r02b_soa32.ncu-rep --page raw --csv > gpurun_out/r02b_raw.csv
ncu -i gpurun_out/r02b_soa32.ncu-rep --page source --csv --print-source sass > gpurun_out/r02b_sass.csv
python profiles/key_metrics.py gpurun_out/r02b_raw.csv
python profiles/sass_mix.py gpurun_out/r02b_sass.csv | head -32
