#!/usr/bin/env python
"""bench.py -- throughput of the batched scalar-multiplication hot path on B200.

Workload (BASELINE.json configs[2], the one the metric is quoted on): secp256k1 variable-base
k_i * P_i over 2^24 random (k, P) pairs per GPU with batch affine normalisation and SEC1 compressed
output.  Inputs are synthetic (device Philox): k_i < n, P_i = r_i * G.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--lanes L]

One JSON line on stdout (rank 0).  `value` = lanes/s with inputs resident in HBM (CUDA events on
the launching stream, max over ranks); `e2e` = the same metric through the host-buffer C-ABI call
(fec_batch_mul) with pinned HOST buffers, H2D and D2H inside the timed region; `roofline` = the
k*P kernel against the measured int32-multiply (IMAD.WIDE) issue rate, plus its HBM side;
`cpu_baseline` = the CPU oracle port on the box's host cores over a bounded sample of the same
workload (also used to check those lanes of the GPU output bit for bit).

--impl reference times the reference's CPU implementation of the path (the oracle port: the
reference is Rust and cannot be built in this image) with every host thread, rank 0 only.
"""
import argparse
import ctypes
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "scalar-mults/sec (secp256k1 k*P, 2^24 random (k,P) pairs, batch affine normalisation, SEC1 compressed out)"
UNIT = "scalar-mults/s"
CANON_IMUL32_PER_OP = 172_562   # SURVEY.md 8d, config 3 (fixed 4-bit window, Jacobian, Montgomery-trick inversion)
IO_BYTES_PER_OP = 32 + 64 + 33  # scalar + affine point in, SEC1 compressed out
SYNTH_STREAM = 3
# What the implemented algorithm (GLV split, 2 x 32 signed 4-bit digits, effective-affine table, shared inversion) really
# executes is NOT a constant of this file: it is read from profiles/exec_counts.json, the artefact tools/ncu_opcounts.py writes
# from the round's `ncu --set full --import-source on` capture of the same kernels (per kernel: IMAD.WIDE and total warp
# instructions per lane, DRAM bytes).  Without the artefact the executed-instruction figures are reported as null.
EXEC_COUNTS_PATH = os.path.join(ROOT, "profiles", "exec_counts.json")
ISSUE_CYCLES_IMAD_WIDE = 4  # measured (tools/ubench.cu, profiles/r2a_ubench_issue_model.log): an IMAD.WIDE holds the issue port 4 clk


def load_exec_counts():
    try:
        d = json.load(open(EXEC_COUNTS_PATH))
        k = d["kernels"]
        return {"source": d.get("source"), "note": d.get("note"),
                "prepare": k["k_glv_prep"], "ladder": k["k_glv_ladder"], "finish": k["k_glv_finish"]}
    except Exception:
        return None
# canonical IMUL32 per operation of the secondary configurations
CANON_FIELD = {"secp256k1": {"mul": 74, "sqr": 46}, "p256": {"mul": 64, "sqr": 36}, "25519": {"mul": 73, "sqr": 45}}
CANON_OTHER = {"x25519_base_point": 153_576, "secp256k1_kG": 44_716, "p256_kG": 37_716, "p256_kP": 142_992, "x25519": 153_576, "ed25519_keygen": 32_631}


# ------------------------------------------------------------------------------------------
# CPU oracle (checker / baseline only)
# ------------------------------------------------------------------------------------------
def load_oracle():
    path = os.path.join(ROOT, "oracle", "_build", "liboracle_std.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(path)
    lib.ostd_batch_mul.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_int,
                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    lib.ostd_batch_mul.restype = None
    return lib


def load_oracle_ref():
    path = os.path.join(ROOT, "oracle", "_build", "liboracle_ref.so")
    if not os.path.exists(path):
        return None
    lib = ctypes.CDLL(path)
    lib.oref_secp256k1_batch_mul.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int]
    lib.oref_secp256k1_batch_mul.restype = None
    return lib


def host_inputs(n, first_lane=0):
    """the same Philox inputs as the device generator, restated on the CPU (oracle/pymodel.py)"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pymodel as pm

    k = pm.random_scalars_be(0, n, SYNTH_STREAM, first_lane, edge=True, sub=0)
    r = pm.random_scalars_be(0, n, SYNTH_STREAM, first_lane, edge=False, sub=1)
    return np.ascontiguousarray(k), np.ascontiguousarray(r)


def cpu_points(lib, r, threads):
    n = r.shape[0]
    pts = np.empty((n, 64), dtype=np.uint8)
    lib.ostd_batch_mul(0, r.ctypes.data, None, 0, n, 2, pts.ctypes.data, None, threads)
    return pts


def time_cpu(lib_std, lib_ref, k, pts, threads):
    """returns (lanes/s, out, kind-description).  Prefers the literal restatement of the reference's
    own algorithm (oracle_ref) for timing; outputs for the parity check always come from oracle_std."""
    n = k.shape[0]
    out = np.empty((n, 33), dtype=np.uint8)
    t0 = time.perf_counter()
    lib_std.ostd_batch_mul(0, k.ctypes.data, pts.ctypes.data, 0, n, 0, out.ctypes.data, None, threads)
    t_std = time.perf_counter() - t0
    t_ref = None
    if lib_ref is not None:
        o2 = np.empty((n, 33), dtype=np.uint8)
        t0 = time.perf_counter()
        lib_ref.oref_secp256k1_batch_mul(k.ctypes.data, pts.ctypes.data, n, o2.ctypes.data, threads)
        t_ref = time.perf_counter() - t0
    return n / t_std, (n / t_ref if t_ref else None), out


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.dev, self.proc, self.lines = dev, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.dev)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
def run_reference(args):
    """the reference arm: the path's CPU implementation on the host cores (oracle port), rank 0 only"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    lib_std, lib_ref = load_oracle(), load_oracle_ref()
    # bounded sample per step: about 2 s of work per step
    probe = 64 * threads
    k, r = host_inputs(probe)
    pts = cpu_points(lib_std, r, threads)
    rate_std, rate_ref, _ = time_cpu(lib_std, lib_ref, k, pts, threads)
    rate = rate_ref or rate_std
    sample = max(threads, int(rate * 2.0) // threads * threads)
    k, r = host_inputs(sample)
    pts = cpu_points(lib_std, r, threads)
    times = []
    for i in range(args.warmup + args.steps):
        out = np.empty((sample, 33), dtype=np.uint8)
        t0 = time.perf_counter()
        if lib_ref is not None:
            lib_ref.oref_secp256k1_batch_mul(k.ctypes.data, pts.ctypes.data, sample, out.ctypes.data, threads)
        else:
            lib_std.ostd_batch_mul(0, k.ctypes.data, pts.ctypes.data, 0, sample, 0, out.ctypes.data, None, threads)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = sample * args.steps / total
    kind = "port"
    what = ("oracle/ref: literal C restatement of forge-ec's own limb code and 256-step ladder (secp256k1.rs:2635-2692)"
            if lib_ref is not None else "oracle/std: C port with the reference's algorithm shape (double-and-add, per-point Fermat inversion)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 limbs (u64 on the CPU)", "data": "synthetic (Philox-4x32-10, seed 'forge-ec')",
            "config": {"workload": "secp256k1 k*P, 2^24 (k,P) pairs [BASELINE.json configs[2]]; each step = a bounded sample of it",
                       "lanes_per_step": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": f"{sample} lanes per step; {what}"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lanes", type=int, default=1 << 24, help="lanes per GPU (default 2^24 = the named configuration)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (development)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (development)")
    ap.add_argument("--no-others", action="store_true", help="skip the secondary configurations (development)")
    ap.add_argument("--no-gather", action="store_true", help="skip the optional NCCL gather measurement (N > 1)")
    ap.add_argument("--no-field", action="store_true", help="skip the field microbench (BASELINE.json configs[1])")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg (2^24 lanes in total over the N GPUs)")
    ap.add_argument("--no-single-process", action="store_true", help="skip the one-process multi-device leg (N > 1, rank 0)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # keep stdout to the ONE JSON line: libraries (NCCL's version banner, for one) write to file descriptor 1 directly,
    # so everything that is not the result goes to stderr and the result is written to the saved descriptor at the end
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    cpu_pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # a CPU-side group for the one wait that must not occupy the GPUs: while rank 0 drives all N devices from its own process,
        # an NCCL barrier kernel spinning on the other ranks' GPUs would time-slice with rank 0's work there
        cpu_pg = dist.new_group(backend="gloo")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    failures = []  # anything that must make this run exit non-zero (a number next to a wrong result is worthless)
    fec = importlib.import_module("forge-ec_b200")  # raises loudly if the CUDA library is missing
    eng = fec.Engine()                               # current device; no CPU fallback exists
    n = args.lanes
    first_lane = rank * n                            # contiguous lane range of this rank (weak scaling)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    dev = torch.device("cuda", local)
    d_k = torch.empty(n * 32, dtype=torch.uint8, device=dev)
    d_r = torch.empty(n * 32, dtype=torch.uint8, device=dev)
    d_p = torch.empty(n * 64, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n * 33, dtype=torch.uint8, device=dev)
    d_st = torch.empty(n, dtype=torch.uint8, device=dev)
    # untimed set-up: k_i, r_i from the device Philox generator, P_i = r_i * G with the fixed-base kernel
    eng.synth_dev(1, SYNTH_STREAM, 0, first_lane, n, d_k, edge=True, stream=sp)
    eng.synth_dev(1, SYNTH_STREAM, 1, first_lane, n, d_r, edge=False, stream=sp)
    eng.batch_mul_dev(fec.CURVE_SECP256K1, d_r, None, n, d_p, out_fmt=fec.OUT_AFFINE_XY_BE, stream=sp)
    torch.cuda.synchronize()
    # roofline denominator: measured multiply-pipe issue rates on this GPU (IMUL32/s)
    peak_wide = eng.imad_peak(0)      # IMAD.WIDE.U32, one instruction per IMUL32
    peak_lohi = eng.imad_peak(1) / 2  # IMAD + IMAD.HI, two instructions per IMUL32
    peak_chain = eng.imad_peak(2)     # carry-chained IMAD.WIDE.U32.X
    imad_peak = max(peak_wide, peak_lohi, peak_chain)

    def step():
        eng.batch_mul_dev(fec.CURVE_SECP256K1, d_k, d_p, n, d_out, d_st, stream=sp)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record(stream)
    for i in range(args.steps):
        step()
        evs[i + 1].record(stream)
    barrier()
    launches = eng.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = evs[0].elapsed_time(evs[-1])
    kernel_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    # per-kernel durations of the same step, live: one more (untimed) step with CUDA events between the three kernels
    eng.stage_timing(True)
    step()
    torch.cuda.synchronize()
    ms_prep, ms_lad, ms_fin, chunks = eng.stage_ms()
    eng.stage_timing(False)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * n * args.steps / (total_ms_max * 1e-3)
    kernel_s = statistics.mean(kernel_ms) * 1e-3

    # ---- end to end through the host-buffer C ABI: pinned host buffers, H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        h_k = torch.empty(n * 32, dtype=torch.uint8).pin_memory()
        h_p = torch.empty(n * 64, dtype=torch.uint8).pin_memory()
        h_out = torch.empty(n * 33, dtype=torch.uint8).pin_memory()
        h_st = torch.empty(n, dtype=torch.uint8).pin_memory()
        h_k.copy_(d_k); h_p.copy_(d_p)
        torch.cuda.synchronize()

        def e2e_step():
            eng.batch_mul_ptr(fec.CURVE_SECP256K1, h_k.data_ptr(), h_p.data_ptr(), n, h_out.data_ptr(), h_st.data_ptr())

        e2e_steps = max(1, args.steps)
        e2e_step()  # warm-up (allocates the pipeline buffers)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(h_out, d_out.cpu())) and bool(torch.equal(h_st, d_st.cpu()))
        e2e = {"value": world * n * e2e_steps / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": n * 96,
               "d2h_bytes_per_step": n * 34, "steps": e2e_steps, "timer": "host wall clock around the synchronous fec_batch_mul call",
               "host_buffers": "pinned (torch pin_memory)", "matches_device_resident_run": same}
        if not same:
            failures.append("e2e: the host-buffer call returned different bytes than the device-resident run")
        # the same call on PAGEABLE host memory -- what a Rust Vec<u8> or a numpy array is: cudaMemcpyAsync stages such copies
        # through the driver's own pinned buffer and blocks the calling thread, so the three-stream pipeline overlaps less
        p_k, p_p = h_k.numpy().copy(), h_p.numpy().copy()
        p_out, p_st = np.empty(n * 33, dtype=np.uint8), np.empty(n, dtype=np.uint8)
        eng.batch_mul_ptr(fec.CURVE_SECP256K1, p_k.ctypes.data, p_p.ctypes.data, n, p_out.ctypes.data, p_st.ctypes.data)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            eng.batch_mul_ptr(fec.CURVE_SECP256K1, p_k.ctypes.data, p_p.ctypes.data, n, p_out.ctypes.data, p_st.ctypes.data)
        barrier()
        dtp = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dtp, op=dist.ReduceOp.MAX)
        e2e["pageable"] = {"value": world * n * e2e_steps / float(dtp.item()), "unit": UNIT,
                           "matches_device_resident_run": bool(np.array_equal(p_out, h_out.numpy())) and bool(np.array_equal(p_st, h_st.numpy()))}
        del p_k, p_p, p_out, p_st

    # ---- strong scaling (BASELINE.json configs[2] read as 2^24 pairs IN TOTAL at 1/2/4/8 GPUs): every rank takes 2^24 / N lanes
    strong = None
    if not args.no_strong:
        ns = max(1024, (1 << 24) // world)
        ns = min(ns, n)

        def strong_step():
            eng.batch_mul_dev(fec.CURVE_SECP256K1, d_k, d_p, ns, d_out, d_st, stream=sp)

        strong_step()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(args.steps):
            strong_step()
        s1.record(stream)
        barrier()
        ts = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        strong = {"scaling": "strong", "lanes_total": ns * world, "lanes_per_gpu": ns, "value": world * ns * args.steps / (float(ts.item()) * 1e-3),
                  "unit": UNIT, "ms_per_step": float(ts.item()) / args.steps}
        if not args.no_e2e:
            eng.batch_mul_ptr(fec.CURVE_SECP256K1, h_k.data_ptr(), h_p.data_ptr(), ns, h_out.data_ptr(), h_st.data_ptr())
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                eng.batch_mul_ptr(fec.CURVE_SECP256K1, h_k.data_ptr(), h_p.data_ptr(), ns, h_out.data_ptr(), h_st.data_ptr())
            barrier()
            dts = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dts, op=dist.ReduceOp.MAX)
            strong["e2e"] = {"value": world * ns * args.steps / float(dts.item()), "unit": UNIT, "host_buffers": "pinned"}

    # ---- SURVEY 8e, optional and outside every timed throughput figure: the device-side gather of the shards' results over
    # NVLink / NVSwitch (ncclAllGather of n x 33 B per rank), timed on the device, max over ranks
    gather = None
    if world > 1 and not args.no_gather:
        g_all = torch.empty(world * n * 33, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(g_all, d_out)   # warm-up (communicator set-up)
        torch.cuda.synchronize()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(3):
            dist.all_gather_into_tensor(g_all, d_out)
        g1.record()
        torch.cuda.synchronize()
        gt = torch.tensor([g0.elapsed_time(g1) / 3.0], dtype=torch.float64, device=dev)
        dist.all_reduce(gt, op=dist.ReduceOp.MAX)
        ok = bool(torch.equal(g_all[rank * n * 33:(rank + 1) * n * 33], d_out))
        gather = {"collective": "ncclAllGather (torch.distributed all_gather_into_tensor)", "bytes_per_rank": n * 33, "ms": float(gt.item()),
                  "algorithmic_GBps_per_gpu": (world - 1) * n * 33 / (float(gt.item()) * 1e-3) / 1e9,
                  "share_of_step": float(gt.item()) / (total_ms_max / args.steps), "own_shard_intact": ok,
                  "note": "not part of value / e2e: the path has no exchange step, results normally leave by per-device D2H copies"}
        del g_all

    # ---- the other configurations of BASELINE.json, device-resident, a few repetitions each (reported, not the headline)
    others = None
    if not args.no_others:
        others = {}
        m = min(n, 1 << 22)
        d_o = torch.empty(m * 65, dtype=torch.uint8, device=dev)

        def timed(fn, reps=3):
            fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                fn()
            e1.record(stream)
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps * 1e-3

        runs = {
            "secp256k1_kG": lambda: eng.batch_mul_dev(fec.CURVE_SECP256K1, d_k, None, m, d_o, stream=sp),
            "x25519": lambda: eng.batch_x25519_dev(d_k, d_r, m, d_o, stream=sp),
            "x25519_base_point": lambda: eng.batch_x25519_dev(d_k, None, m, d_o, stream=sp),
            "ed25519_keygen": lambda: eng.batch_ed25519_pubkey_dev(d_k, True, m, d_o, stream=sp),
        }
        for name, fn in runs.items():
            t = timed(fn)
            others[name] = {"value": world * m / t, "unit": "ops/s", "lanes_per_gpu": m,
                            "canonical_imad_frac": CANON_OTHER[name] * m / t / imad_peak}
        # BASELINE.json configs[0] at its own size (2^16 scalars: a batch that does not fill the persistent grids)
        m0 = min(m, 1 << 16)
        t = timed(lambda: eng.batch_mul_dev(fec.CURVE_SECP256K1, d_k, None, m0, d_o, stream=sp), reps=10)
        others["secp256k1_kG_2p16"] = {"value": world * m0 / t, "unit": "ops/s", "lanes_per_gpu": m0,
                                       "canonical_imad_frac": CANON_OTHER["secp256k1_kG"] * m0 / t / imad_peak}
        # BASELINE.json configs[3], as stated: X25519 key exchange over 2^24 random keypairs (per GPU).  One exchange = the four
        # calls A = X25519(a, 9), B = X25519(b, 9), s1 = X25519(a, B), s2 = X25519(b, A); the two shared secrets must be equal.
        mx = min(n, 1 << 24)
        d_A = torch.empty(mx * 32, dtype=torch.uint8, device=dev)
        d_B = torch.empty(mx * 32, dtype=torch.uint8, device=dev)
        d_s1 = torch.empty(mx * 32, dtype=torch.uint8, device=dev)
        d_s2 = torch.empty(mx * 32, dtype=torch.uint8, device=dev)

        def exchange():
            eng.batch_x25519_dev(d_k, None, mx, d_A, stream=sp)
            eng.batch_x25519_dev(d_r, None, mx, d_B, stream=sp)
            eng.batch_x25519_dev(d_k, d_B, mx, d_s1, stream=sp)
            eng.batch_x25519_dev(d_r, d_A, mx, d_s2, stream=sp)

        t = timed(exchange, reps=2)
        agree = bool(torch.equal(d_s1, d_s2)) and bool(d_s1.view(mx, 32).any(dim=1).all())
        others["x25519_key_exchange"] = {"value": world * 4 * mx / t, "unit": "scalar-mults/s (4 per exchange)", "keypairs_per_gpu": mx,
                                         "exchanges_per_s": world * mx / t, "shared_secrets_agree": agree,
                                         "canonical_imad_frac": 4 * mx / t * CANON_OTHER["x25519"] / imad_peak}
        del d_A, d_B, d_s1, d_s2
        if not agree:
            failures.append("x25519_key_exchange: the two parties' shared secrets differ")
        # P-256 needs scalars below its own order and points on its own curve
        d_k2 = torch.empty(m * 32, dtype=torch.uint8, device=dev)
        d_p2 = torch.empty(m * 64, dtype=torch.uint8, device=dev)
        eng.synth_dev(2, 5, 0, first_lane, m, d_k2, edge=True, stream=sp)
        t = timed(lambda: eng.batch_mul_dev(fec.CURVE_P256, d_k2, None, m, d_p2, out_fmt=fec.OUT_AFFINE_XY_BE, stream=sp))
        others["p256_kG"] = {"value": world * m / t, "unit": "ops/s", "lanes_per_gpu": m, "canonical_imad_frac": CANON_OTHER["p256_kG"] * m / t / imad_peak}
        # BASELINE.json configs[4], as stated: P-256 k*G + Ed25519 key generation with COMPRESSED-point encoding, 2^26 scalars over
        # 8 GPUs = 2^23 lanes per GPU (SEC1 compressed 33 B for P-256, the RFC 8032 32-byte encoding for Ed25519)
        m4 = min(n, 1 << 23)
        d_k4 = torch.empty(m4 * 32, dtype=torch.uint8, device=dev)
        d_o4 = torch.empty(m4 * 33, dtype=torch.uint8, device=dev)
        eng.synth_dev(2, 5, 0, rank * m4, m4, d_k4, edge=True, stream=sp)
        t1 = timed(lambda: eng.batch_mul_dev(fec.CURVE_P256, d_k4, None, m4, d_o4, out_fmt=fec.OUT_SEC1_COMPRESSED, stream=sp))
        t2 = timed(lambda: eng.batch_ed25519_pubkey_dev(d_k4, True, m4, d_o4, stream=sp))
        others["p256_kG_2p23_compressed"] = {"value": world * m4 / t1, "unit": "ops/s", "lanes_per_gpu": m4,
                                             "canonical_imad_frac": CANON_OTHER["p256_kG"] * m4 / t1 / imad_peak}
        others["ed25519_keygen_2p23"] = {"value": world * m4 / t2, "unit": "ops/s", "lanes_per_gpu": m4,
                                         "canonical_imad_frac": CANON_OTHER["ed25519_keygen"] * m4 / t2 / imad_peak}
        others["config4_p256_kG_plus_ed25519_keygen"] = {"value": world * 2 * m4 / (t1 + t2), "unit": "ops/s (both key types for every scalar)",
                                                         "lanes_per_gpu": m4, "canonical_imad_frac": None}
        del d_k4, d_o4
        t = timed(lambda: eng.batch_mul_dev(fec.CURVE_P256, d_k2, d_p2, m, d_o, stream=sp))
        others["p256_kP"] = {"value": world * m / t, "unit": "ops/s", "lanes_per_gpu": m, "canonical_imad_frac": CANON_OTHER["p256_kP"] * m / t / imad_peak}
        # SURVEY 8f rows next to the path: ECDH (k*P, x-coordinate, HKDF-SHA-256 on the device) and the multi-scalar
        # multiplication sum k_i*P_i (bucket method); per GPU, no canonical IMUL32 figure exists for them
        d_st2 = torch.empty(m, dtype=torch.uint8, device=dev)
        t = timed(lambda: eng.batch_ecdh_dev(fec.CURVE_SECP256K1, d_k, d_p, m, d_o, d_st2, kdf=fec.KDF_HKDF_SHA256, info=b"bench", stream=sp))
        others["secp256k1_ecdh_hkdf"] = {"value": world * m / t, "unit": "ops/s", "lanes_per_gpu": m, "canonical_imad_frac": None}
        t = timed(lambda: eng.batch_ecdh_dev(fec.CURVE_P256, d_k2, d_p2, m, d_o, d_st2, kdf=fec.KDF_HKDF_SHA256, info=b"bench", stream=sp))
        others["p256_ecdh_hkdf"] = {"value": world * m / t, "unit": "ops/s", "lanes_per_gpu": m, "canonical_imad_frac": None}
        t = timed(lambda: eng.msm_dev(fec.CURVE_SECP256K1, d_k, d_p, m, d_o, stream=sp))
        others["secp256k1_msm"] = {"value": world * m / t, "unit": "points/s (one sum per GPU)", "lanes_per_gpu": m, "canonical_imad_frac": None}
        t = timed(lambda: eng.msm_dev(fec.CURVE_P256, d_k2, d_p2, m, d_o, stream=sp))
        others["p256_msm"] = {"value": world * m / t, "unit": "points/s (one sum per GPU)", "lanes_per_gpu": m, "canonical_imad_frac": None}
        # BASELINE.json configs[1]: field mul / square microbench, 2^26 independent lanes, `iters` dependent operations per lane in
        # registers (one operation per lane would measure HBM, not the multiply pipe); unit = field operations/s
        if not args.no_field:
            fl = 1 << 26
            d_fa = torch.empty(fl * 32, dtype=torch.uint8, device=dev)
            d_fb = torch.empty(fl * 32, dtype=torch.uint8, device=dev)
            d_fo = torch.empty(fl * 32, dtype=torch.uint8, device=dev)
            eng.synth_dev(0, 2, 0, first_lane, fl, d_fa, stream=sp)
            eng.synth_dev(0, 2, 1, first_lane, fl, d_fb, stream=sp)
            d_fa.view(fl, 32)[:, 31] &= 0x3F   # little-endian records below 2^254: canonical residues of all three primes
            d_fb.view(fl, 32)[:, 31] &= 0x3F
            torch.cuda.synchronize()
            iters = 1024
            for pname, prime in (("secp256k1", fec.PRIME_SECP256K1), ("p256", fec.PRIME_P256), ("25519", fec.PRIME_25519)):
                for oname, op in (("mul", fec.FIELD_MUL), ("sqr", fec.FIELD_SQR)):
                    t = timed(lambda: eng.field_bench_dev(prime, op, fl, iters, d_fa, d_fb, d_fo, stream=sp), reps=1)
                    others[f"field_{pname}_{oname}"] = {
                        "value": world * fl * iters / t, "unit": "field operations/s", "lanes_per_gpu": fl, "iters_per_lane": iters,
                        "canonical_imad_frac": CANON_FIELD[pname][oname] * fl * iters / t / imad_peak}
            del d_fa, d_fb, d_fo
        if world > 1:
            for v in others.values():  # max-over-ranks time -> min over ranks of the rate
                if v["canonical_imad_frac"] is None:
                    tv = torch.tensor([v["value"]], dtype=torch.float64, device=dev)
                    dist.all_reduce(tv, op=dist.ReduceOp.MIN)
                    v["value"] = float(tv[0])
                    continue
                tv = torch.tensor([v["value"], v["canonical_imad_frac"]], dtype=torch.float64, device=dev)
                dist.all_reduce(tv, op=dist.ReduceOp.MIN)
                v["value"], v["canonical_imad_frac"] = float(tv[0]), float(tv[1])
        others["note"] = ("canonical_imad_frac = SURVEY 8d's IMUL32 per op x ops/s / measured IMAD.WIDE rate; above 1 where the "
                          "implemented algorithm executes fewer multiplications than the canonical one (16-bit windows, GLV, "
                          "base-point X25519 through the Edwards table)")
        del d_o, d_k2, d_p2

    # ---- the ONE-PROCESS multi-device form (fec_ctx_create with a device list, one fec_batch_mul call on
    # host buffers, the library shards the lanes over its devices with one host thread each).  Rank 0 alone runs it on all N
    # GPUs while the other ranks wait at the barrier; 2^24 lanes in total, pinned and pageable host buffers.
    single = None
    if world > 1 and not args.no_single_process and not args.no_e2e:
        barrier()
        torch.cuda.synchronize()
        if rank == 0:
            tot = min(n, 1 << 24)
            with fec.Engine(devices=list(range(world))) as many:
                res = {}
                for kind in ("pinned", "pageable"):
                    if kind == "pinned":
                        a_k, a_p, a_o, a_s = h_k.data_ptr(), h_p.data_ptr(), h_out.data_ptr(), h_st.data_ptr()
                    else:
                        q_k, q_p = h_k.numpy()[: tot * 32].copy(), h_p.numpy()[: tot * 64].copy()
                        q_o, q_s = np.empty(tot * 33, dtype=np.uint8), np.empty(tot, dtype=np.uint8)
                        a_k, a_p, a_o, a_s = q_k.ctypes.data, q_p.ctypes.data, q_o.ctypes.data, q_s.ctypes.data
                    many.batch_mul_ptr(fec.CURVE_SECP256K1, a_k, a_p, tot, a_o, a_s)  # warm-up: allocates every device's pipeline
                    t0 = time.perf_counter()
                    for _ in range(max(1, args.steps)):
                        many.batch_mul_ptr(fec.CURVE_SECP256K1, a_k, a_p, tot, a_o, a_s)
                    dt1 = time.perf_counter() - t0
                    res[kind] = tot * max(1, args.steps) / dt1
                ok = bool(np.array_equal(q_o, h_out.numpy()[: tot * 33])) and bool(torch.equal(h_out[: tot * 33], d_out[: tot * 33].cpu()))
                single = {"value": res["pinned"], "pageable_value": res["pageable"], "unit": UNIT, "lanes_total": tot, "devices": world,
                          "what": "fec.Engine(devices=range(N)) in rank 0's process: one fec_batch_mul call on host buffers, lanes sharded "
                                  "by the library (one host thread, three streams per device)",
                          "matches_per_rank_results": ok}
                if not ok:
                    failures.append("e2e_single_process: the multi-device context returned different bytes")
        dist.barrier(group=cpu_pg)  # the other ranks wait here on the CPU (gloo), their GPUs idle
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline on the host cores, bounded sample of the same lanes; doubles as a parity check
    cpu = None
    if not args.no_cpu:
        threads = os.cpu_count() or 1
        lib_std, lib_ref = load_oracle(), load_oracle_ref()
        probe = 32 * threads
        k, r = host_inputs(probe)
        pts = cpu_points(lib_std, r, threads)
        rate_std, rate_ref, _ = time_cpu(lib_std, lib_ref, k, pts, threads)
        slow = min(rate_std, rate_ref) if rate_ref else rate_std
        sample = int(min(n, max(probe, int(slow * 8.0) // threads * threads)))
        k, r = host_inputs(sample)
        pts = cpu_points(lib_std, r, threads)
        rate_std, rate_ref, out = time_cpu(lib_std, lib_ref, k, pts, threads)
        gpu_head = d_out[: sample * 33].cpu().numpy().reshape(sample, 33)
        gpu_pts = d_p[: sample * 64].cpu().numpy().reshape(sample, 64)
        parity = bool(np.array_equal(gpu_head, out)) and bool(np.array_equal(gpu_pts, pts))
        cpu = {"value": rate_ref if rate_ref else rate_std, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": (f"first {sample} lanes of the workload; " +
                          ("oracle/ref = literal C restatement of forge-ec's limb code and 256-step ladder"
                           if rate_ref else "oracle/std = C port shaped like the reference (double-and-add, per-point Fermat inversion)")),
               "std_value": rate_std, "gpu_bit_exact_on_sample": parity}
        if not parity:
            print("PARITY FAILURE: GPU output differs from the CPU oracle on the sampled lanes", file=sys.stderr)
            failures.append("gpu_bit_exact_on_sample is false")

    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        hbm_src = "MEASURED_PEAKS.json"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = CANON_IMUL32_PER_OP * n / kernel_s
    stage_total = ms_prep + ms_lad + ms_fin
    lanes_per_launch = n / max(chunks, 1)
    ladder_ms = ms_lad / max(chunks, 1)
    ex = load_exec_counts()
    stages = {
        "kernels": ["k_glv_prep", "k_glv_ladder", "k_glv_finish"], "launches_per_step": 3 * chunks, "lanes_per_launch": lanes_per_launch,
        "ms_per_step": [ms_prep, ms_lad, ms_fin], "share_of_step": [ms_prep / stage_total, ms_lad / stage_total, ms_fin / stage_total],
        "source": "CUDA events between the kernels of one extra step (fec_ctx_stage_timing)",
    }
    frac_executed = issue = ladder_traffic = None
    if ex is not None:
        names = ("prepare", "ladder", "finish")
        wide = {k: ex[k]["imad_wide_per_lane"] for k in names}
        tot = {k: ex[k]["total_per_lane"] for k in names}
        stages["executed_imul32_per_op"] = wide
        stages["executed_instructions_per_op"] = tot
        stages["counts_source"] = f"{os.path.relpath(EXEC_COUNTS_PATH, ROOT)}: {ex['source']}; {ex['note']}"
        # IMAD.WIDE instructions the kernels really issue / measured IMAD.WIDE issue rate: how busy the multiply pipe is
        stages["pipe_utilisation"] = {"ladder": wide["ladder"] * n / (ms_lad * 1e-3) / imad_peak,
                                      "step": sum(wide.values()) * n / kernel_s / imad_peak}
        frac_executed = stages["pipe_utilisation"]["step"]
        # The issue port, not the multiply pipe, is what the kernels saturate: an IMAD.WIDE holds an SM sub-partition's issue port
        # for 4 cycles and every other instruction for 1 (tools/ubench.cu).  cycles needed = (4 W + others) warp instructions per
        # warp-operation; cycles available = SMSPs x clock x time.
        sm_hz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
        smsps = 4 * torch.cuda.get_device_properties(local).multi_processor_count

        def port(k, ms):
            need = (ISSUE_CYCLES_IMAD_WIDE * wide[k] + (tot[k] - wide[k])) * n / 32.0
            return need / (smsps * sm_hz * 1e6 * ms * 1e-3)

        issue = {"model": "issue cycles = 4 x IMAD.WIDE + 1 x every other warp instruction (measured: profiles/r2a_ubench_issue_model.log)",
                 "issue_port_utilisation": {"prepare": port("prepare", ms_prep), "ladder": port("ladder", ms_lad), "finish": port("finish", ms_fin)},
                 "imad_wide_share_of_issue_cycles": {k: ISSUE_CYCLES_IMAD_WIDE * wide[k] / (ISSUE_CYCLES_IMAD_WIDE * wide[k] + tot[k] - wide[k]) for k in names},
                 "sm_mhz_used": sm_hz,
                 "reading": "the ladder fills its issue slots; its IMAD.WIDE share of them is what pipe_utilisation reports, so more "
                            "multiply throughput needs fewer non-multiply instructions per multiplication, not more overlap"}
        dram = ex["ladder"].get("dram_bytes_per_lane")
        ladder_traffic = dram * lanes_per_launch if dram else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 limbs (8 x 32-bit, IMAD.WIDE.U32 carry chains)", "data": "synthetic (device Philox-4x32-10, seed 'forge-ec')",
        "config": {"workload": "secp256k1 variable-base k*P + batch affine normalisation + SEC1 compressed [BASELINE.json configs[2]]",
                   "lanes_per_gpu": n, "sharding": "contiguous lane ranges, one rank per GPU, no collective on the data path (weak scaling: "
                                                   "2^24 lanes per GPU; the `strong` key holds 2^24 lanes in total)",
                   "l2": "inputs 1.5 GiB + outputs 0.55 GiB per step exceed the 126 MB L2; no flush needed"},
        "roofline": {"bound": "imad", "achieved": achieved / 1e9, "peak": imad_peak / 1e9, "unit": "GIMUL32/s", "frac": achieved / imad_peak,
                     "frac_executed": frac_executed,
                     "frac_note": "frac counts SURVEY 8d's canonical 172 562 IMUL32 per operation (the GLV split executes fewer, hence > 1); "
                                  "frac_executed counts the IMAD.WIDE instructions the kernels execute (ncu) and is the multiply pipe's real load",
                     "issue_port": issue,
                     "traffic": ladder_traffic,
                     "traffic_note": "k_glv_ladder, bytes per launch, from one ncu --set full capture (profiles/); its algorithmic "
                                     "traffic is 64 table entries x 64 B + 36 B in + 96 B out = 4.2 KB per lane, most of it served by L2",
                     "kernel": "k_glv_prep + k_glv_ladder + k_glv_finish (the canonical figure covers the whole operation; "
                               "k_glv_ladder dominates, see stages)",
                     "kernel_ms": kernel_s * 1e3, "dominant_kernel_ms_per_launch": ladder_ms,
                     "algorithmic_imul32_per_op": CANON_IMUL32_PER_OP, "stages": stages,
                     "peak_source": "fec_imad_peak measured in this run (IMUL32/s): " +
                                    f"IMAD.WIDE.U32 {peak_wide:.4g}, IMAD+IMAD.HI pairs {peak_lohi:.4g}, chained IMAD.WIDE.U32.X {peak_chain:.4g}",
                     "hbm": {"bound": "hbm", "achieved": IO_BYTES_PER_OP * n / kernel_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": IO_BYTES_PER_OP * n / kernel_s / 1e9 / hbm_peak, "peak_source": hbm_src,
                             "algorithmic_bytes_per_op": IO_BYTES_PER_OP}},
        "cpu_baseline": cpu, "e2e": e2e, "strong": strong, "e2e_single_process": single, "gpu_launches": int(launches) * world,
        "device_gather": gather, "other_configs": others, "clocks": clocks, "failures": failures,
    }
    sys.stdout.flush()
    os.write(result_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    if failures:  # a throughput next to a wrong result is worthless: make the run fail
        print("BENCH FAILED: " + "; ".join(failures), file=sys.stderr)
        sys.exit(1)


if __name__ == "__main__":
    main()
