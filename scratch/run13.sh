nproc; nvidia-smi topo -m 2>/dev/null | head -14
python tests/micro/h2d_bw.py 8 2>&1 | tee gpurun_out/r02e_h2d_bw.txt
python bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02e_bench_n8.json 2> gpurun_out/r02e_bench_n8.err; tail -c 600 gpurun_out/r02e_bench_n8.err
python -c "import json; d=json.load(open('gpurun_out/r02e_bench_n8.json')); print(d['value'], d['e2e'], d['e2e_sequential']['value'], d['e2e_files'])"
