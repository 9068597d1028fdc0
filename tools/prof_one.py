"""Development aid: ONE call of a chosen entry point at 2^LG lanes, for `ncu -k regex:...` captures.
  python tools/prof_one.py p256_kP|k1_kP|x25519|p256_kG|k1_kG [lg]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
fec = importlib.import_module("forge-ec_b200")
what = sys.argv[1]
n = 1 << int(sys.argv[2] if len(sys.argv) > 2 else 20)
eng = fec.Engine(); s = torch.cuda.current_stream().cuda_stream
k = torch.empty(n * 32, dtype=torch.uint8, device="cuda"); r = torch.empty_like(k)
P = torch.empty(n * 64, dtype=torch.uint8, device="cuda"); out = torch.empty(n * 33, dtype=torch.uint8, device="cuda")
curve = 1 if what.startswith("p256") else 0
eng.synth_dev(1 + curve, 3, 0, 0, n, k, stream=s); eng.synth_dev(1 + curve, 3, 1, 0, n, r, stream=s)
if what.endswith("kP"):
    eng.batch_mul_dev(curve, r, None, n, P, out_fmt=2, stream=s)
    eng.batch_mul_dev(curve, k, P, n, out, stream=s)
elif what.endswith("kG"):
    eng.batch_mul_dev(curve, k, None, n, out, stream=s)
else:
    eng.batch_x25519_dev(k, P[: n * 32], n, r, stream=s)
torch.cuda.synchronize()
print("done", what, n)
