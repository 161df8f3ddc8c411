python bench.py --steps 20 --warmup 5 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; tail -c 400 gpurun_out/r02e_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02e_bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02e_launches.csv python bench.py --steps 1 --warmup 1 --batches-per-step 2 --files-pairs 0 --torch-pairs 0 --cpu-pairs 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"piv_soa|predictor" -c 4 -o gpurun_out/r02e_prof -f python tools/_prof.py 8 > gpurun_out/ncu_r02e.log 2>&1
ncu -i gpurun_out/r02e_prof.ncu-rep --page raw --csv > gpurun_out/r02e_raw.csv
ncu -i gpurun_out/r02e_prof.ncu-rep --page source --csv --print-source sass > gpurun_out/r02e_sass.csv
python -c "import json; d=json.load(open('gpurun_out/r02e_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['pass_first']['frac'], d['roofline']['whole_step_frac'], d['e2e_files'])"
