"""Development aid: P-256 / secp256k1 k*G at 2^22 lanes for each output format."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
fec = importlib.import_module("forge-ec_b200")
eng = fec.Engine(); s = torch.cuda.current_stream()
n = 1 << 22
k = torch.empty(n * 32, dtype=torch.uint8, device="cuda"); out = torch.empty(n * 65, dtype=torch.uint8, device="cuda")
for curve in (0, 1):
    eng.synth_dev(1 + curve, 1, 0, 0, n, k, stream=s.cuda_stream)
    for fmt in (0, 1, 2, 3):
        fn = lambda: eng.batch_mul_dev(curve, k, None, n, out, out_fmt=fmt, stream=s.cuda_stream)
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); fn(); fn(); fn(); e1.record(s); torch.cuda.synchronize()
        print(f"curve {curve} fmt {fmt}: {n/(e0.elapsed_time(e1)/3*1e-3)/1e6:.1f} M/s", flush=True)
