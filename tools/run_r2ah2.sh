set -x
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_x25519|k_wmul_fixed|k_ed25519_pubkey" -o gpurun_out/prof_b_r2ah -f python tools/prof_all.py 22 > gpurun_out/ncu_b_r2ah.log 2>&1; echo "ncu rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2ah.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-field > gpurun_out/ncu_launch_r2ah.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out/
