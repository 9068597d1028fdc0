"""Development aid: quick bit-exact check of the k*P paths of the library named by FORGE_EC_CUDA_LIB against the
oracle (pymodel) on a few hundred lanes incl. edge scalars -- run before timing an A/B variant."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
fec = importlib.import_module("forge-ec_b200")
import pymodel as pm
ok = True
with fec.Engine() as eng:
    for curve, n in ((0, 320), (1, 320)):
        k = pm.random_scalars_be(curve, n, stream=11)
        r = pm.random_scalars_be(curve, n, stream=11, sub=1, edge=False)
        pts, _ = pm.batch_mul(curve, r.tobytes(), None, n, pm.OUT_AFFINE_XY_BE)
        exp, est = pm.batch_mul(curve, k.tobytes(), pts, n, 0)
        import numpy as np
        out, st = eng.batch_mul(curve, k, np.frombuffer(pts, dtype=np.uint8).reshape(n, 64))
        good = out.tobytes() == exp and st.tobytes() == est
        ok &= good
        print(f"curve {curve}: {'bit-exact' if good else 'MISMATCH'} on {n} lanes", flush=True)
sys.exit(0 if ok else 1)
