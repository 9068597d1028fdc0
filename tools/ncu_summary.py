"""Summarise an Nsight Compute report (read here, no GPU needed) into a small text file for profiles/.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_var_k1.txt

Emits: launch geometry, duration, registers, occupancy, pipe utilisation, issue statistics, DRAM bytes,
warp-stall breakdown and the executed-instruction histogram by SASS opcode (from the source page)."""
import collections
import csv
import io
import re
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, dst):
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    lines = []
    for vals in raw[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        lines.append(f"== kernel: {d.get('Kernel Name', ('?',))[0]}  grid {d.get('Grid Size', ('?',))[0]} block {d.get('Block Size', ('?',))[0]}")
        keep = [
            "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__maximum_warps_per_active_cycle_pct", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
            "smsp__warps_active.avg.per_cycle_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fmaheavy.sum",
            "sm__inst_executed_pipe_fmalite.sum", "sm__inst_executed_pipe_lsu.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        ]
        for k in keep:
            if k in d:
                lines.append(f"{k:72s} {d[k][0]:>18s} {d[k][1]}")
        lines.append("-- warp stall reasons (warps stalled per issue-active cycle)")
        for k in sorted(d):
            m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", k)
            if m and float(d[k][0] or 0) >= 0.005:
                lines.append(f"   {m.group(1):28s} {float(d[k][0]):8.3f}")
    src = page(rep, "source")
    if len(src) > 2:
        h = src[1]
        ix = {n: i for i, n in enumerate(h)}
        tot, n = collections.Counter(), 0
        for r in src[2:]:
            if len(r) < len(h):
                continue
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
            try:
                ex = int(r[ix["Instructions Executed"]] or 0)
            except ValueError:  # repeated header row (one per function of the kernel image)
                continue
            tot[m.group(2) if m else "?"] += ex
            n += ex
        lines.append(f"-- executed warp instructions by opcode (static SASS instructions: {len(src) - 2}; executed: {n})")
        for op, c in tot.most_common(24):
            lines.append(f"   {op:24s} {c:14d} {100 * c / max(n, 1):5.1f}%")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
