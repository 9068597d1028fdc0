"""Development aid: instruction mix of every loop (backward branch) and every called routine of one kernel in an object file /
shared library, in the terms of the issue-port model: an IMAD.WIDE occupies the issue port for 4 cycles, any other
instruction for 1.   python tools/sass_loops.py <file> <kernel-name-substring>"""
import collections
import re
import subprocess
import sys


def disasm(path, kernel):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    ins, on = [], False
    for l in out.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m:
            on = kernel in m.group(1)
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m and on:
            ins.append((int(m.group(1), 16), m.group(2)))
    return ins


def klass(text):
    text = re.sub(r"^@!?U?P\d+ ", "", text)
    op = text.split()[0]
    if op.startswith("IMAD.WIDE"):
        return "wide"
    if op.startswith(("LDS", "STS", "LDL", "STL", "LDG", "STG", "LDGSTS", "LD.", "ST.")):
        return "lsu"
    if op.startswith(("IMAD", "HFMA2")):
        return "fma_other"
    if op.startswith(("CALL", "RET", "BRA", "EXIT", "BSSY", "BSYNC", "WARPSYNC")):
        return "ctl"
    if op == "NOP":
        return "nop"
    return "alu"


def mix(body):
    c = collections.Counter(klass(t) for _, t in body)
    c.pop("nop", None)
    n = sum(c.values())
    w = c.get("wide", 0)
    return n, w, dict(c), (4.0 * w / (4.0 * w + (n - w)) if n else 0.0)


if __name__ == "__main__":
    ins = disasm(sys.argv[1], sys.argv[2])
    print(f"{sys.argv[2]}: {len(ins)} instructions")
    addr = {a: i for i, (a, _) in enumerate(ins)}
    targets = collections.Counter()
    for a, t in ins:
        m = re.search(r"CALL\S* 0x([0-9a-f]+)", t)
        if m:
            targets[int(m.group(1), 16)] += 1
    rets = [a for a, t in ins if t.startswith("RET")]
    for tgt, cnt in sorted(targets.items()):
        end = min([r for r in rets if r >= tgt], default=None)
        if end is None:
            continue
        n, w, c, share = mix([x for x in ins if tgt <= x[0] <= end])
        print(f"  routine 0x{tgt:x} ({cnt} call sites): {n} instr, {w} IMAD.WIDE, issue cycles {4 * w + n - w}, wide share {share:.3f}  {c}")
    for a, t in ins:
        m = re.search(r"BRA\S* (?:!?U?P\d, )?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            lo = int(m.group(1), 16)
            n, w, c, share = mix([x for x in ins if lo <= x[0] <= a])
            if n >= 100:
                print(f"  loop 0x{lo:x}..0x{a:x}: {n} instr, {w} IMAD.WIDE, issue cycles {4 * w + n - w}, wide share {share:.3f}  {c}")
