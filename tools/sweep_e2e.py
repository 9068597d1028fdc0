"""Development aid: end-to-end (pinned host buffers) secp256k1 k*P rate of fec_batch_mul at 2^24 lanes."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
fec = importlib.import_module("forge-ec_b200")
eng = fec.Engine(); s = torch.cuda.current_stream().cuda_stream
n = 1 << 24
k = torch.empty(n * 32, dtype=torch.uint8, device="cuda"); r = torch.empty_like(k)
P = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
eng.synth_dev(1, 3, 0, 0, n, k, stream=s); eng.synth_dev(1, 3, 1, 0, n, r, stream=s)
eng.batch_mul_dev(0, r, None, n, P, out_fmt=2, stream=s)
torch.cuda.synchronize()
hk = k.cpu().pin_memory(); hp = P.cpu().pin_memory()
ho = torch.empty(n * 33, dtype=torch.uint8).pin_memory(); hs = torch.empty(n, dtype=torch.uint8).pin_memory()
def call():
    rc = eng._lib.fec_batch_mul(eng._ctx, 0, 0, hk.data_ptr(), hp.data_ptr(), 0, n, 0, ho.data_ptr(), hs.data_ptr())
    assert rc == 0
call()
t0 = time.perf_counter(); call(); call(); dt = (time.perf_counter() - t0) / 2
print(f"chunk_lg={os.environ.get('FORGE_EC_CHUNK_LG','18')} e2e {n/dt/1e6:.2f} M/s")
