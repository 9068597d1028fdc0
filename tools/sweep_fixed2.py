"""Development aid: time k*G for both curves (2^22 lanes) through the device-pointer ABI."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
fec = importlib.import_module("forge-ec_b200")
eng = fec.Engine(); s = torch.cuda.current_stream()
n = 1 << 22
k = torch.empty(n * 32, dtype=torch.uint8, device="cuda"); out = torch.empty(n * 33, dtype=torch.uint8, device="cuda")
res = []
for curve in (0, 1):
    eng.synth_dev(1 + curve, 1, 0, 0, n, k, stream=s.cuda_stream)
    fn = lambda: eng.batch_mul_dev(curve, k, None, n, out, stream=s.cuda_stream)
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s); fn(); fn(); fn(); e1.record(s); torch.cuda.synchronize()
    res.append(f"{'k1' if curve == 0 else 'p256'} kG {n/(e0.elapsed_time(e1)/3*1e-3)/1e6:.1f} M/s")
print("  ".join(res))
