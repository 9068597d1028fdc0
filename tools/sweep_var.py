"""Development aid: time k*P (secp256k1) through the device-pointer ABI for several batch sizes."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
fec = importlib.import_module("forge-ec_b200")
eng = fec.Engine()
s = torch.cuda.current_stream()
N = 1 << 24
k = torch.empty(N * 32, dtype=torch.uint8, device="cuda"); r = torch.empty_like(k)
P = torch.empty(N * 64, dtype=torch.uint8, device="cuda"); out = torch.empty(N * 33, dtype=torch.uint8, device="cuda")
eng.synth_dev(1, 3, 0, 0, N, k, stream=s.cuda_stream); eng.synth_dev(1, 3, 1, 0, N, r, stream=s.cuda_stream)
eng.batch_mul_dev(0, r, None, N, P, out_fmt=2, stream=s.cuda_stream)
torch.cuda.synchronize()
def t(fn, reps=2):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(reps): fn()
    e1.record(s); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
LGS = tuple(int(x) for x in os.environ.get("SWEEP_LG", "16,17,18,19,20,22,24").split(","))
for lg in LGS:
    n = 1 << lg
    dt = t(lambda: eng.batch_mul_dev(0, k, P, n, out, stream=s.cuda_stream))
    print(f"n=2^{lg}: {n/dt/1e6:.2f} M/s  ({dt*1e3:.2f} ms)", flush=True)
# chunked on one stream
for lg in ((18, 20) if "SWEEP_LG" not in os.environ else ()):
    c = 1 << lg
    def chunked():
        for off in range(0, N, c):
            eng.batch_mul_dev(0, k[off*32:], P[off*64:], c, out[off*33:], stream=s.cuda_stream)
    dt = t(chunked, 1)
    print(f"2^24 in chunks of 2^{lg} on one stream: {N/dt/1e6:.2f} M/s", flush=True)
if os.environ.get("SWEEP_STAGES"):
    n = 1 << 22
    eng.stage_timing(True)
    eng.batch_mul_dev(0, k, P, n, out, stream=s.cuda_stream)
    torch.cuda.synchronize()
    a, b, c, ch = eng.stage_ms()
    eng.stage_timing(False)
    print(f"stages at 2^22: prepare {a:.2f} ms, ladder {b:.2f} ms ({n/b/1e3:.1f} M/s), finish {c:.2f} ms", flush=True)
