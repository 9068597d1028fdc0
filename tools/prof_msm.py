"""Development aid: ONE fec_msm_dev call at 2^LG points (for an ncu launch list)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
fec = importlib.import_module("forge-ec_b200")
curve = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = 1 << int(sys.argv[2] if len(sys.argv) > 2 else 22)
eng = fec.Engine(); s = torch.cuda.current_stream().cuda_stream
k = torch.empty(n * 32, dtype=torch.uint8, device="cuda"); r = torch.empty_like(k)
P = torch.empty(n * 64, dtype=torch.uint8, device="cuda"); out = torch.empty(64, dtype=torch.uint8, device="cuda")
eng.synth_dev(1 + curve, 3, 0, 0, n, k, stream=s); eng.synth_dev(1 + curve, 3, 1, 0, n, r, stream=s)
eng.batch_mul_dev(curve, r, None, n, P, out_fmt=2, stream=s)
eng.msm_dev(curve, k, P, n, out, stream=s)
torch.cuda.synchronize()
eng.msm_dev(curve, k, P, n, out, stream=s)
torch.cuda.synchronize()
print("done")
