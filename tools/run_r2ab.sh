set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_r2ab_8gpu.json 2> gpurun_out/bench_r2ab_8gpu.err; echo "rc=$?"
tail -3 gpurun_out/bench_r2ab_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 3 --warmup 3 --no-strong --no-single-process > gpurun_out/bench_r2ab_4gpu.json 2> gpurun_out/bench_r2ab_4gpu.err; echo "rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/bench_r2ab_8gpu.json','gpurun_out/bench_r2ab_4gpu.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value", d['value']/1e6, "e2e", d['e2e']['value']/1e6, "strong", (d.get('strong') or {}).get('value'), "single", (d.get('e2e_single_process') or {}).get('value'), "failures", d.get('failures'))
PY
