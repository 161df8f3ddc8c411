python bench.py --steps 3 --warmup 3 --files-pairs 2048 --torch-pairs 0 --cpu-pairs 1 2>gpurun_out/bench_try.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps(d['e2e_files'])); print(d['value'], d['e2e']['value'])"
tail -3 gpurun_out/bench_try.err
nproc; lscpu | grep -E "Model name|Thread|Core|Socket|MHz" | head
python tests/micro/h2d_bw.py 1 2>&1 | tail -8
ncu --set full --clock-control none --import-source on -k regex:"piv_soa" -c 2 -o gpurun_out/r02d_soa -f python tools/_prof.py 8 > gpurun_out/ncu_r02d.log 2>&1
ncu -i gpurun_out/r02d_soa.ncu-rep --page raw --csv > gpurun_out/r02d_raw.csv
ncu -i gpurun_out/r02d_soa.ncu-rep --page source --csv --print-source sass > gpurun_out/r02d_sass.csv
python profiles/key_metrics.py gpurun_out/r02d_raw.csv
