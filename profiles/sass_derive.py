"""Instruction counts per DP cell of the packed Gotoh kernel, derived from the SASS of the built library -- the
constant behind bench.py's `roofline.achieved` (GCUPS x ALU-pipe instructions per cell), so that it is measured on
the binary that runs instead of typed in.

    python profiles/sass_derive.py [path/to/libigd_b200.so] [--kernel 16,22,0,0,1] [--json out.json] [--excerpt out.txt]

Method: `cuobjdump -sass`, the function whose name carries the template arguments, the innermost loop that holds the
DPX instructions (VIADDMNMX / VIMNMX3); the hot path through that loop takes a forward conditional branch exactly when
the region it skips holds no DPX instruction (the per-gene epilogue: result store + column re-initialisation), and one
trip of the loop is one column step of C cells per lane.  Pipes (B300_MICROARCH.md): FFMA / FADD / FMUL / IMAD* issue
to the FMA pipe; LOP3, IADD3, VIADD, VIADDMNMX, VIMNMX*, SEL, ISETP, SHF, PRMT, MOV, LEA to the ALU pipe; everything
else (loads, shuffles, conversions, branches, constant loads) is counted as "other".
"""
import argparse
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FMA = ("FFMA", "FADD", "FMUL", "IMAD", "HFMA2", "FMNMX")      # FMNMX is ALU on this part; listed below
ALU = ("LOP3", "IADD3", "IADD", "VIADD", "VIADDMNMX", "VIMNMX", "VIMNMX3", "SEL", "ISETP", "SHF", "PRMT", "MOV", "LEA",
       "FSEL", "FSETP", "FMNMX", "PLOP3", "POPC", "FLO", "BREV", "R2P", "P2R", "CSEL", "ISCADD")
DPX = ("VIADDMNMX", "VIMNMX3")


def functions(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    name, cur = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            cur[name] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and name:
            cur[name].append((int(m.group(1), 16), m.group(2).strip()))
    return cur


def opcode(text):
    t = text.split()
    if t[0].startswith("@"):
        t = t[1:]
    return t[0].split(".")[0]


def hot_loop(ins):
    """(head index, back-edge index) of the innermost loop holding the most DPX instructions"""
    addr = {a: i for i, (a, _) in enumerate(ins)}
    best = None
    for i, (a, text) in enumerate(ins):
        if opcode(text) != "BRA":
            continue
        m = re.search(r"0x([0-9a-f]+)", text.split("BRA")[-1])
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr:
            h = addr[tgt]
            n = sum(opcode(t) in DPX for _, t in ins[h:i + 1])
            size = i - h
            if n and (best is None or n > best[2] or (n == best[2] and size < best[3])):
                best = (h, i, n, size)
    if best is None:
        raise RuntimeError("no loop with DPX instructions found")
    return best[0], best[1]


def hot_path(ins, h, e):
    addr = {a: i for i, (a, _) in enumerate(ins)}
    path, i = [], h
    while i <= e:
        a, text = ins[i]
        path.append((a, text))
        if opcode(text) == "BRA" and text.startswith("@") and i != e:
            m = re.search(r"0x([0-9a-f]+)", text.split("BRA")[-1])
            tgt = int(m.group(1), 16) if m else None
            if tgt is not None and tgt > a and tgt in addr and addr[tgt] <= e + 1:
                skipped = ins[i + 1:addr[tgt]]
                if not any(opcode(t) in DPX for _, t in skipped):
                    i = addr[tgt]
                    continue
        i += 1
    return path


def derive(lib, targs, cells):
    pat = "gotoh_packed_kernelILi%sELi%sELi%sELb%sELb%sE" % tuple(targs)
    fns = functions(lib)
    names = [n for n in fns if pat in n]
    if len(names) != 1:
        raise RuntimeError("%d functions match %s" % (len(names), pat))
    ins = fns[names[0]]
    h, e = hot_loop(ins)
    path = hot_path(ins, h, e)
    counts = {}
    for _, text in path:
        counts[opcode(text)] = counts.get(opcode(text), 0) + 1
    fma = sum(v for k, v in counts.items() if k in ("FFMA", "FADD", "FMUL", "IMAD", "HFMA2"))
    alu = sum(v for k, v in counts.items() if k in ALU)
    total = sum(counts.values())
    return {"library": os.path.relpath(lib, ROOT), "function": names[0], "template_args": list(targs), "cells_per_trip": cells,
            "loop": ["0x%04x" % ins[h][0], "0x%04x" % ins[e][0]], "hot_path_instructions": total,
            "counts": dict(sorted(counts.items(), key=lambda kv: -kv[1])),
            "alu_pipe_per_cell": alu / cells, "fma_pipe_per_cell": fma / cells, "other_per_cell": (total - alu - fma) / cells,
            "all_per_cell": total / cells}, path


def excerpt(lib, out):
    """SASS lines that show what the kernels are built from: TMA bulk copies + mbarrier transactions in the scan
    (cp.async.bulk -> UBLKCP, expect_tx -> SYNCS), DPX max instructions in the aligner."""
    fns = functions(lib)
    with open(out, "w") as f:
        for pat, ops in (("rss_scan_kernelILb1ELb1E", ("UBLKCP", "SYNCS", "LDS", "SHF", "LOP3")),
                         ("gotoh_packed_kernelILi16ELi22ELi0ELb0ELb1E", ("VIADDMNMX", "VIMNMX3", "FFMA", "IMAD", "LOP3"))):
            for name in [n for n in fns if pat in n]:
                ins = fns[name]
                f.write("# %s: %d SASS instructions\n" % (name, len(ins)))
                hist = {}
                for _, t in ins:
                    hist[opcode(t)] = hist.get(opcode(t), 0) + 1
                f.write("# opcode histogram: " + ", ".join("%s %d" % kv for kv in sorted(hist.items(), key=lambda kv: -kv[1])[:24]) + "\n")
                shown = 0
                for a, t in ins:
                    if opcode(t) in ops[:2] and shown < 24:
                        f.write("  /*%04x*/ %s ;\n" % (a, t))
                        shown += 1
                f.write("\n")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("lib", nargs="?", default=os.path.join(ROOT, "igdetective_b200", "libigd_b200.so"))
    ap.add_argument("--kernel", default="16,22,0,0,1", help="template arguments G,C,MIXED,LONG,FPINC")
    ap.add_argument("--json")
    ap.add_argument("--excerpt")
    a = ap.parse_args()
    targs = a.kernel.split(",")
    res, path = derive(a.lib, targs, int(targs[1]))
    print(json.dumps(res, indent=1))
    if a.json:
        json.dump(res, open(a.json, "w"), indent=1)
    if a.excerpt:
        excerpt(a.lib, a.excerpt)
        with open(a.excerpt, "a") as f:
            f.write("# hot path of one column step (%d cells per lane) of %s\n" % (res["cells_per_trip"], res["function"]))
            for adr, t in path:
                f.write("  /*%04x*/ %s ;\n" % (adr, t))
