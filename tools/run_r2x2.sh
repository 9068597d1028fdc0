set -x
ncu --set full --clock-control none --import-source on -k regex:"k_x25519|k_wmul_fixed|k_ed25519_pubkey" -o gpurun_out/prof_b_r2x -f python tools/prof_all.py 20 > gpurun_out/ncu_b_r2x.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2x.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-field > gpurun_out/ncu_launch_r2x.log 2>&1; echo "ncu2 rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r2x.json 2> gpurun_out/bench_ref_r2x.err; echo "ref rc=$?"
