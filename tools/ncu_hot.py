"""Development aid: per-function and per-instruction stall summary from an ncu report's source page.
  python tools/ncu_hot.py report.ncu-rep [function-substring]  -> per-function totals, or the listing of one function"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
sel = sys.argv[2] if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fn = None; hdr = None
per = collections.OrderedDict()
for r in rows:
    if r and r[0] == "Kernel Name":
        fn = r[1]; continue
    if r and r[0] == "Address":
        hdr = {n: i for i, n in enumerate(r)}; continue
    if hdr is None or len(r) < len(hdr): continue
    per.setdefault(fn, []).append(r)
cols = ["stall_wait", "stall_math", "stall_dispatch", "stall_not_selected", "stall_selected", "stall_short_sb", "stall_long_sb", "stall_branch_resolving", "stall_no_inst", "stall_mio", "stall_lg"]
def num(r, c):
    try: return int(r[hdr[c]] or 0)
    except ValueError: return 0
if sel is None:
    print(f"{'function':60s} {'instrs':>6s} {'executed':>12s} {'samples':>8s} " + " ".join(f"{c[6:12]:>7s}" for c in cols))
    for f, rs in per.items():
        print(f"{f[:60]:60s} {len(rs):6d} {sum(num(r,'Instructions Executed') for r in rs):12d} {sum(num(r,'# Samples') for r in rs):8d} " + " ".join(f"{sum(num(r,c) for r in rs):7d}" for c in cols))
else:
    for f, rs in per.items():
        if sel not in f: continue
        print("==", f)
        for r in rs:
            print(f"{r[hdr['Address']][-5:]} {r[hdr['Source']][:70]:70s} ex={num(r,'Instructions Executed'):>10d} smp={num(r,'# Samples'):>6d} " + " ".join(f"{c[6:10]}={num(r,c)}" for c in cols if num(r,c)))
