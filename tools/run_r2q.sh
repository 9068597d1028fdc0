set -x
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2q.json 2> gpurun_out/bench_r2q.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r2q.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r2q.json 2> gpurun_out/bench_ref_r2q.err; echo "ref rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_glv|k_p3" -o gpurun_out/prof_a_r2q -f python tools/prof_all.py 20 > gpurun_out/ncu_a_r2q.log 2>&1; echo "ncu rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2q.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-field > gpurun_out/ncu_launch_r2q.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out/
