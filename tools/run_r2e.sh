set -x
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r2e.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r2e.json 2> gpurun_out/bench_ref_r2e.err; echo "ref rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_glv|k_p3|k_x25519|k_wmul_fixed|k_ed25519_pubkey" -o gpurun_out/prof_all_r2e -f python tools/prof_all.py 20 > gpurun_out/ncu_all_r2e.log 2>&1; echo "ncu rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2e.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-field > gpurun_out/ncu_launch_r2e.log 2>&1; echo "ncu2 rc=$?"
