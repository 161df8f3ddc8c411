set -x
ncu --set full --clock-control none --import-source on -k regex:"piv_soa" -c 1 -o gpurun_out/r02a_soa32 -f python tools/_prof.py 8 > gpurun_out/ncu_r02a.log 2>&1
ncu -i gpurun_out/r02a_soa32.ncu-rep --page raw --csv > gpurun_out/r02a_raw.csv
ncu -i gpurun_out/r02a_soa32.ncu-rep --page source --csv --print-source sass > gpurun_out/r02a_sass.csv
python profiles/key_metrics.py gpurun_out/r02a_raw.csv
python profiles/sass_mix.py gpurun_out/r02a_sass.csv | head -45
