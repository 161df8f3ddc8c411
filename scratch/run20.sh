python -m pytest tests -m gpu -x -q -k "offline or stencil or worker or field" 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --files-pairs 4000 --torch-pairs 0 --cpu-pairs 0 2>gpurun_out/bench_try.err | python -c "import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][0]); print(json.dumps(d['e2e_files'])); print(d['value'], d['e2e']['value'])"
