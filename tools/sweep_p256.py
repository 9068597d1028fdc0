import importlib, os, sys
sys.path.insert(0, "/root/repo" if os.path.exists("/root/repo/bench.py") else os.getcwd())
import torch
fec = importlib.import_module("forge-ec_b200")
eng = fec.Engine(); s = torch.cuda.current_stream()
n = 1 << int(os.environ.get("SWEEP_LG", "22"))
k = torch.empty(n * 32, dtype=torch.uint8, device="cuda"); P = torch.empty(n * 64, dtype=torch.uint8, device="cuda"); out = torch.empty(n * 33, dtype=torch.uint8, device="cuda")
eng.synth_dev(2, 5, 0, 0, n, k, stream=s.cuda_stream)
eng.batch_mul_dev(1, k, None, n, P, out_fmt=2, stream=s.cuda_stream)
fn = lambda: eng.batch_mul_dev(1, k, P, n, out, stream=s.cuda_stream)
fn(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s); fn(); fn(); e1.record(s); torch.cuda.synchronize()
print(f"p256 kP {n/(e0.elapsed_time(e1)/2*1e-3)/1e6:.2f} M/s")
