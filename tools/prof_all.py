"""Development aid: ONE call of every hot entry point at 2^LG lanes (for one `ncu --set full` capture of all kernels)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
fec = importlib.import_module("forge-ec_b200")
n = 1 << int(sys.argv[1] if len(sys.argv) > 1 else 20)
eng = fec.Engine(); s = torch.cuda.current_stream().cuda_stream
k = torch.empty(n * 32, dtype=torch.uint8, device="cuda"); r = torch.empty_like(k); k2 = torch.empty_like(k)
P = torch.empty(n * 64, dtype=torch.uint8, device="cuda"); P2 = torch.empty_like(P); out = torch.empty(n * 65, dtype=torch.uint8, device="cuda")
eng.synth_dev(1, 3, 0, 0, n, k, stream=s); eng.synth_dev(1, 3, 1, 0, n, r, stream=s); eng.synth_dev(2, 5, 0, 0, n, k2, stream=s)
eng.batch_mul_dev(0, r, None, n, P, out_fmt=2, stream=s)      # k_wmul_fixed<CurveK1>
eng.batch_mul_dev(1, k2, None, n, P2, out_fmt=2, stream=s)    # k_wmul_fixed_vm<CurveP256>
eng.batch_mul_dev(0, k, P, n, out, stream=s)                  # k_glv_prep / ladder / finish
eng.batch_mul_dev(1, k2, P2, n, out, stream=s)                # k_p3_prep / ladder / finish
eng.batch_x25519_dev(k, r, n, out, stream=s)                  # k_x25519
eng.batch_ed25519_pubkey_dev(k, True, n, out, stream=s)       # k_ed25519_pubkey
torch.cuda.synchronize()
print("done")
