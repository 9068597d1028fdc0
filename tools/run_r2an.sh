set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_device or several_devices or devices" 2>&1 | tail -3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r2an_2gpu.json 2> gpurun_out/bench_r2an_2gpu.err; echo "rc=$?"
tail -3 gpurun_out/bench_r2an_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2an_2gpu.json').read())
print("value", d['value']/1e6, "e2e", d['e2e']['value']/1e6, "pageable", d['e2e']['pageable']['value']/1e6, "strong", d['strong'], "single", d['e2e_single_process'], "failures", d['failures'])
PY
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_r2an_2gpu.json 2>/dev/null; echo "ref rc=$?"
