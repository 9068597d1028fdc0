"""Development aid: X25519 ladder throughput at 2^22 lanes."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
fec = importlib.import_module("forge-ec_b200")
eng = fec.Engine(); s = torch.cuda.current_stream()
n = 1 << 22
k = torch.empty(n * 32, dtype=torch.uint8, device="cuda"); u = torch.empty_like(k); out = torch.empty_like(k)
eng.synth_dev(0, 4, 0, 0, n, k, stream=s.cuda_stream); eng.synth_dev(0, 4, 1, 0, n, u, stream=s.cuda_stream)
fn = lambda: eng.batch_x25519_dev(k, u, n, out, stream=s.cuda_stream)
fn(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s); fn(); fn(); fn(); e1.record(s); torch.cuda.synchronize()
print(f"x25519 {n/(e0.elapsed_time(e1)/3*1e-3)/1e6:.2f} M/s")
