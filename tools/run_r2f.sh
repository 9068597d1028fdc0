set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 2 --warmup 3 --no-others --no-cpu > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2f.json').read())
print("value", d['value']/1e6, "e2e", d['e2e']['value']/1e6, "pageable", d['e2e']['pageable'], "failures", d['failures'])
print("roofline", d['roofline']['frac'], d['roofline']['frac_executed'], d['roofline']['issue_port'])
PY
FORGE_EC_STAGE=0 python bench.py --steps 2 --warmup 3 --no-others --no-cpu --no-strong 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('no staging: pageable', d['e2e']['pageable'])"
FORGE_EC_CT_TABLE=1 python tools/sweep_fixed.py 2>&1 | tail -6
python tools/sweep_fixed.py 2>&1 | tail -6
