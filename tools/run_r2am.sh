set -x
timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python tools/sweep_e2e.py 2>&1 | tail -1
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2am.json 2> gpurun_out/bench_r2am.err; echo "bench rc=$?"
tail -2 gpurun_out/bench_r2am.err
