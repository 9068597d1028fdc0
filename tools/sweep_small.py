"""Development aid: throughput of every entry point at small and medium batch sizes (device-pointer ABI)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
fec = importlib.import_module("forge-ec_b200")
eng = fec.Engine(); s = torch.cuda.current_stream()
N = 1 << 20
k = torch.empty(N * 32, dtype=torch.uint8, device="cuda"); r = torch.empty_like(k); k2 = torch.empty_like(k)
P = torch.empty(N * 64, dtype=torch.uint8, device="cuda"); P2 = torch.empty_like(P); out = torch.empty(N * 65, dtype=torch.uint8, device="cuda")
eng.synth_dev(1, 3, 0, 0, N, k, stream=s.cuda_stream); eng.synth_dev(1, 3, 1, 0, N, r, stream=s.cuda_stream)
eng.synth_dev(2, 5, 0, 0, N, k2, stream=s.cuda_stream)
eng.batch_mul_dev(0, r, None, N, P, out_fmt=2, stream=s.cuda_stream)
eng.batch_mul_dev(1, k2, None, N, P2, out_fmt=2, stream=s.cuda_stream)
torch.cuda.synchronize()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(reps): fn()
    e1.record(s); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
sp = s.cuda_stream
for lg in (10, 12, 14, 16, 18, 20):
    n = 1 << lg
    runs = {
        "k1_kG": lambda: eng.batch_mul_dev(0, k, None, n, out, stream=sp),
        "p256_kG": lambda: eng.batch_mul_dev(1, k2, None, n, out, stream=sp),
        "k1_kP": lambda: eng.batch_mul_dev(0, k, P, n, out, stream=sp),
        "p256_kP": lambda: eng.batch_mul_dev(1, k2, P2, n, out, stream=sp),
        "x25519": lambda: eng.batch_x25519_dev(k, r, n, out, stream=sp),
        "x25519_base": lambda: eng.batch_x25519_dev(k, None, n, out, stream=sp),
        "ed25519": lambda: eng.batch_ed25519_pubkey_dev(k, True, n, out, stream=sp),
    }
    print(f"n=2^{lg}: " + "  ".join(f"{name} {n/t(fn)/1e6:7.1f}" for name, fn in runs.items()), flush=True)
