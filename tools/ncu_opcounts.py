"""Executed-instruction counts per kernel from an Nsight Compute report (read here, no GPU needed), as the JSON artefact
bench.py derives `roofline.frac_executed` and the issue-port model from (instead of constants in bench.py).

  python tools/ncu_opcounts.py gpurun_out/prof.ncu-rep[,second.ncu-rep...] <lanes of the profiled calls> profiles/exec_counts.json [note]

(several reports, comma-separated: a capture of all kernels with sources is larger than one gpurun call brings back)

The report must come from `ncu --set full --import-source on` (the source page carries "Instructions Executed" per SASS
instruction).  Per kernel: warp instructions executed in total and by class, and the same per LANE as thread-level counts
(x 32 / lanes) -- `imad_wide_per_lane` is the number of 32x32->64 multiplications a lane really executes."""
import collections
import csv
import io
import json
import re
import subprocess
import sys


def klass(op):
    if op.startswith("IMAD.WIDE"):
        return "imad_wide"
    if op.startswith(("IMAD", "HFMA2")):
        return "fma_other"
    if op.startswith(("LDS", "STS", "LDL", "STL", "LDG", "STG", "LDGSTS", "LD.", "ST.", "ATOM", "RED")):
        return "lsu"
    if op.startswith(("CALL", "RET", "BRA", "EXIT", "BSSY", "BSYNC", "WARPSYNC", "BAR", "NOP")):
        return "control"
    return "alu"


def kname(full):
    """kernel name without its parameter list; template arguments kept, `(bool)0` / `0` spelled the same on every page"""
    n = full.replace("fec::", "").replace("(bool)0", "0").replace("(bool)1", "1")
    depth = 0
    for i, ch in enumerate(n):
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            return n[:i]
    return n


def main(reps, lanes, dst, note=""):
    rows = []
    elapsed = collections.OrderedDict()
    for rep in reps.split(","):
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows += list(csv.reader(io.StringIO(out)))
        # elapsed SM cycles, duration and DRAM bytes per launch from the raw page (units are chosen per report)
        raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
        if len(raw) <= 2:
            continue
        h = {n: i for i, n in enumerate(raw[0])}
        for vals in raw[2:]:
            kn = kname(vals[h["Kernel Name"]])
            if kn in elapsed:
                continue

            def num(key):
                try:
                    x = float(vals[h[key]].replace(",", ""))
                    return None if x != x else x  # ncu prints -nan for a counter it could not collect
                except (KeyError, ValueError):
                    return None

            cyc = num("sm__cycles_elapsed.max")
            cyc_src = "sm__cycles_elapsed.max"
            if cyc is None and num("gpu__time_duration.sum") is not None and num("device__attribute_clock_rate") is not None:
                # (seen on persistent-grid kernels: every *_cycles_elapsed counter of the launch is -nan) duration x SM clock
                unit = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "second": 1.0}.get(raw[1][h["gpu__time_duration.sum"]].replace("μ", "u"), None)
                if unit is not None:
                    cyc = num("gpu__time_duration.sum") * unit * num("device__attribute_clock_rate") * 1e3
                    cyc_src = "gpu__time_duration.sum x device__attribute_clock_rate"
            elapsed[kn] = {"sm_cycles_elapsed": cyc, "cycles_source": cyc_src, "duration": num("gpu__time_duration.sum"),
                           "duration_unit": raw[1][h["gpu__time_duration.sum"]] if "gpu__time_duration.sum" in h else None,
                           "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                           "dram_read": num("dram__bytes_read.sum"), "dram_read_unit": raw[1][h["dram__bytes_read.sum"]] if "dram__bytes_read.sum" in h else "",
                           "dram_write": num("dram__bytes_write.sum"), "dram_write_unit": raw[1][h["dram__bytes_write.sum"]] if "dram__bytes_write.sum" in h else ""}
    rep = reps
    kernels = collections.OrderedDict()
    name, ix = None, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            name = kname(r[1])
            while name in kernels:  # the same kernel launched again: keep the first launch
                name += "'"
            kernels[name] = collections.Counter()
            continue
        if r and r[0] == "Address":
            ix = {n: i for i, n in enumerate(r)}
            continue
        if name is None or ix is None or len(r) < len(ix):
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
        try:
            ex = int(r[ix["Instructions Executed"]] or 0)
        except ValueError:
            continue
        op = m.group(2) if m else "?"
        kernels[name]["total"] += ex
        kernels[name][klass(op)] += ex
    res = {"source": f"{rep} (ncu --set full --import-source on, source page)", "lanes": lanes, "note": note, "smsps": 148 * 4, "kernels": {}}
    for k, c in kernels.items():
        if k.endswith("'") or c["total"] == 0:
            continue
        per = {f"{n}_per_lane": c[n] * 32.0 / lanes for n in ("total", "imad_wide", "fma_other", "alu", "lsu", "control")}
        res["kernels"][k] = {"warp_instructions": dict(c), **per,
                             # issue-port model: an IMAD.WIDE occupies the issue port for 4 cycles, anything else for 1
                             "issue_cycles_per_warp_op": (4 * c["imad_wide"] + c["total"] - c["imad_wide"]) * 32.0 / lanes,
                             "imad_wide_share_of_issue": 4.0 * c["imad_wide"] / max(1, 4 * c["imad_wide"] + c["total"] - c["imad_wide"])}
        e = elapsed.get(k)
        if e and e["sm_cycles_elapsed"]:
            # cycles the model predicts for this launch on 592 SMSPs against the cycles it took
            pred = (4 * c["imad_wide"] + c["total"] - c["imad_wide"]) / (148 * 4)
            res["kernels"][k].update({"sm_cycles_elapsed": e["sm_cycles_elapsed"], "cycles_source": e["cycles_source"], "issue_model_cycles": pred,
                                      "issue_port_utilisation": pred / e["sm_cycles_elapsed"], "ncu_issue_active_pct": e["issue_active_pct"],
                                      "duration": e["duration"], "duration_unit": e["duration_unit"]})
        if e and e.get("dram_read") is not None and e.get("dram_write") is not None:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
            b = e["dram_read"] * scale.get(e["dram_read_unit"], 1.0) + e["dram_write"] * scale.get(e["dram_write_unit"], 1.0)
            res["kernels"][k]["dram_bytes_per_lane"] = b / lanes
    json.dump(res, open(dst, "w"), indent=1)
    for k, v in res["kernels"].items():
        print(f"{k:36s} IMAD.WIDE/lane {v['imad_wide_per_lane']:9.0f}  instr/lane {v['total_per_lane']:9.0f}  wide share of issue {v['imad_wide_share_of_issue']:.3f}"
              f"  issue-port utilisation {v.get('issue_port_utilisation', float('nan')):.3f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
