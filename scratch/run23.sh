echo "== default (pair-packed 64 px first pass)"; python tools/_sweep.py 16 2>&1 | tail -1
echo "== PIVB200_SOA64=0 (round-1 kernel, FP32 row transform)"; PIVB200_SOA64=0 python tools/_sweep.py 16 2>&1 | tail -1
echo "== PIVB200_SOA64=0 PIVB200_TC=1 (round-1 kernel, tcgen05 row transform)"; PIVB200_SOA64=0 PIVB200_TC=1 python tools/_sweep.py 16 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; tail -c 300 gpurun_out/r02f_bench.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02f_bench.json').read().splitlines() if l.startswith('{')][0])
print(d['value'], d['e2e']['value'], d['e2e']['h2d_gbs'], d['e2e']['h2d_ceiling_gbs'], d['roofline']['frac'], d['roofline']['pass_first']['frac'], d['roofline']['whole_step_frac'])
print(d['e2e_files']); print(d['cpu_baseline']['value'], d['torch_eager_baseline']['value'], d['torch_eager_baseline']['offline_piv_from_files'])"
