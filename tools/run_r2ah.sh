set -x
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_glv|k_p3" -o gpurun_out/prof_a_r2ah -f python tools/prof_all.py 22 > gpurun_out/ncu_a_r2ah.log 2>&1; echo "ncu rc=$?"
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2ah.json 2> gpurun_out/bench_r2ah.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r2ah.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r2ah.json 2> gpurun_out/bench_ref_r2ah.err; echo "ref rc=$?"
ls -la gpurun_out/
