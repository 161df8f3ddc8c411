python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "worker_process or offline" 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 --files-pairs 1024 --torch-pairs 0 --cpu-pairs 1 2>gpurun_out/bench_try.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps(d['e2e_files'], indent=1)); print(d['value'], d['e2e']['value'])"
tail -3 gpurun_out/bench_try.err
