"""Development aid: time fixed-base k*G (both curves) and Ed25519 keygen through the device-pointer ABI."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
fec = importlib.import_module("forge-ec_b200")
eng = fec.Engine()
s = torch.cuda.current_stream()
n = 1 << int(os.environ.get("SWEEP_LG", "22"))
k = torch.empty(n * 32, dtype=torch.uint8, device="cuda"); k2 = torch.empty_like(k)
out = torch.empty(n * 65, dtype=torch.uint8, device="cuda")
eng.synth_dev(1, 3, 0, 0, n, k, stream=s.cuda_stream); eng.synth_dev(2, 5, 0, 0, n, k2, stream=s.cuda_stream)
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(reps): fn()
    e1.record(s); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
r = []
r.append(("k1_kG", n / t(lambda: eng.batch_mul_dev(0, k, None, n, out, stream=s.cuda_stream))))
r.append(("p256_kG", n / t(lambda: eng.batch_mul_dev(1, k2, None, n, out, stream=s.cuda_stream))))
r.append(("ed_seed", n / t(lambda: eng.batch_ed25519_pubkey_dev(k, True, n, out, stream=s.cuda_stream))))
r.append(("ed_scalar", n / t(lambda: eng.batch_ed25519_pubkey_dev(k, False, n, out, stream=s.cuda_stream))))
r.append(("x25519_base", n / t(lambda: eng.batch_x25519_dev(k, None, n, out, stream=s.cuda_stream))))
r.append(("x25519_var", n / t(lambda: eng.batch_x25519_dev(k, k2, n, out, stream=s.cuda_stream), 2)))
print("  ".join(f"{a}={b/1e6:.1f}M/s" for a, b in r), flush=True)
