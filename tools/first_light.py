"""Quick device timings of every kernel through the device-pointer C ABI (development aid)."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
fec = importlib.import_module("forge-ec_b200")

def timeit(fn, reps=3):
    st = torch.cuda.current_stream()
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3

res = {}
with fec.Engine() as eng:
    s = torch.cuda.current_stream().cuda_stream
    for v in (0, 1, 2):
        res[f"imad_peak_v{v}_instr_per_s"] = eng.imad_peak(v)
    lanes = 1 << 22
    a = torch.empty(lanes * 32, dtype=torch.uint8, device="cuda"); b = torch.empty_like(a); o = torch.empty_like(a)
    for prime in (0, 1, 2):
        eng.synth_dev(0, 2, 0, 0, lanes, a, stream=s); eng.synth_dev(0, 2, 1, 0, lanes, b, stream=s)
        for op, name in ((0, "mul"), (1, "sqr")):
            iters = 256
            t = timeit(lambda: eng.field_bench_dev(prime, op, lanes, iters, a, b, o, stream=s))
            res[f"field_p{prime}_{name}_per_s"] = lanes * iters / t
    n = 1 << 20
    k = torch.empty(n * 32, dtype=torch.uint8, device="cuda"); r = torch.empty_like(k)
    P = torch.empty(n * 64, dtype=torch.uint8, device="cuda"); out = torch.empty(n * 33, dtype=torch.uint8, device="cuda")
    for curve in (0, 1):
        eng.synth_dev(1 + curve, 3, 0, 0, n, k, stream=s); eng.synth_dev(1 + curve, 3, 1, 0, n, r, stream=s)
        t = timeit(lambda: eng.batch_mul_dev(curve, r, None, n, P, out_fmt=2, stream=s))
        res[f"fixed_c{curve}_per_s"] = n / t
        t = timeit(lambda: eng.batch_mul_dev(curve, k, P, n, out, stream=s))
        res[f"var_c{curve}_per_s"] = n / t
    t = timeit(lambda: eng.batch_x25519_dev(k, None, n, r, stream=s))
    res["x25519_per_s"] = n / t
    t = timeit(lambda: eng.batch_ed25519_pubkey_dev(k, True, n, r, stream=s))
    res["ed25519_seed_per_s"] = n / t
print(json.dumps(res, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "first_light.json"), "w"), indent=1)
