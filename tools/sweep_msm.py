"""Development aid: time fec_msm_dev (secp256k1) for several n; FORGE_EC_MSM_IMPL=products selects the non-bucket path."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
fec = importlib.import_module("forge-ec_b200")
eng = fec.Engine()
s = torch.cuda.current_stream()
N = 1 << 22
k = torch.empty(N * 32, dtype=torch.uint8, device="cuda"); r = torch.empty_like(k)
P = torch.empty(N * 64, dtype=torch.uint8, device="cuda"); out = torch.empty(64, dtype=torch.uint8, device="cuda")
for curve in (0, 1):
    eng.synth_dev(1 + curve, 3, 0, 0, N, k, stream=s.cuda_stream); eng.synth_dev(1 + curve, 3, 1, 0, N, r, stream=s.cuda_stream)
    eng.batch_mul_dev(curve, r, None, N, P, out_fmt=2, stream=s.cuda_stream)
    torch.cuda.synchronize()
    for lg in (15, 16, 18, 20, 22):
        n = 1 << lg
        fn = lambda: eng.msm_dev(curve, k, P, n, out, stream=s.cuda_stream)
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(3): fn()
        e1.record(s); torch.cuda.synchronize()
        dt = e0.elapsed_time(e1) / 3 * 1e-3
        print(f"curve {curve} n=2^{lg}: {dt*1e3:8.2f} ms  {n/dt/1e6:8.1f} M points/s", flush=True)
