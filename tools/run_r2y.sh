set -x
ncu --set full --clock-control none --import-source on -k regex:"k_glv|k_p3" -o gpurun_out/prof_a_r2y -f python tools/prof_all.py 22 > gpurun_out/ncu_a_r2y.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/
