set -x
ncu --set full --clock-control none --import-source on -k regex:"k_x25519|k_wmul_fixed|k_ed25519_pubkey" -o gpurun_out/prof_b_r2ae -f python tools/prof_all.py 22 > gpurun_out/ncu_b_r2ae.log 2>&1; echo "ncu rc=$?"
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2ae.json 2> gpurun_out/bench_r2ae.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r2ae.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r2ae.json 2> gpurun_out/bench_ref_r2ae.err; echo "ref rc=$?"
ls -la gpurun_out/
