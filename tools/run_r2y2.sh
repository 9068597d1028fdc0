set -x
ncu --set full --clock-control none --import-source on -k regex:"k_x25519|k_wmul_fixed|k_ed25519_pubkey" -o gpurun_out/prof_b_r2y -f python tools/prof_all.py 22 > gpurun_out/ncu_b_r2y.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/
