set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2x.json 2> gpurun_out/bench_r2x.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r2x.err
ncu --set full --clock-control none --import-source on -k regex:"k_glv|k_p3" -o gpurun_out/prof_a_r2x -f python tools/prof_all.py 20 > gpurun_out/ncu_a_r2x.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/
