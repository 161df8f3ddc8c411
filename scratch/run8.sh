python bench.py --steps 5 --warmup 3 --files-pairs 512 > gpurun_out/bench_try.json 2> gpurun_out/bench_try.err; tail -c 6000 gpurun_out/bench_try.json; tail -5 gpurun_out/bench_try.err
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -3
