#!/usr/bin/env python3
"""SASS-level accounting of one profiled launch from an .ncu-rep (read here, no GPU needed).

    python tools/ncu_sass.py gpurun_out/x.ncu-rep [units_per_launch] [--regions] [--top N]

Prints, per SASS opcode, the warp instructions executed and thread instructions per unit (LUT entry /
pixel), the stall samples attributed to it, and -- with --regions -- the same per contiguous address
range between branch targets (basic blocks merged into runs of >= 1 % of the executed instructions), so
that the cost of a loop or a phase can be read without source attribution (inlined code is attributed
to the callee's file in the source view).
"""
import collections
import csv
import subprocess
import sys


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    insts = []
    for r in rows:
        if len(r) > 5 and r[0] == "Address":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) - 2:
            continue
        d = dict(zip(hdr, r))
        try:
            insts.append({"addr": int(d["Address"], 16), "sass": d["Source"].strip(), "exec": int(d["Instructions Executed"]),
                          "thr": int(d["Thread Instructions Executed"]), "samples": int(d["# Samples"]),
                          "stalls": {k: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}})
        except (ValueError, KeyError):
            continue
    return insts


def opcode(s):
    t = s.split()
    if not t:
        return "?"
    if t[0].startswith("@"):
        t = t[1:]
    return t[0].split(".")[0] if t else "?"


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    rep = args[0]
    units = float(args[1]) if len(args) > 1 else None
    top = 40
    if "--top" in sys.argv:
        top = int(sys.argv[sys.argv.index("--top") + 1])
    insts = load(rep)
    tot = sum(i["exec"] for i in insts)
    tot_s = sum(i["samples"] for i in insts) or 1
    print(f"SASS instructions: {len(insts)}, warp instructions executed: {tot}" + (f", thread instr per unit: {sum(i['thr'] for i in insts) / units:.1f}" if units else ""))
    per = collections.Counter()
    per_s = collections.Counter()
    for i in insts:
        op = opcode(i["sass"])
        # keep the packed / double flavours apart
        full = i["sass"].split()
        full = full[1] if full and full[0].startswith("@") else (full[0] if full else "?")
        if op in ("MUFU", "LDS", "STS", "LDG", "STG", "LDL", "STL", "ATOMG", "RED", "SHFL", "REDUX"):
            op = ".".join(full.split(".")[:2])
        per[op] += i["exec"]
        per_s[op] += i["samples"]
    print(f"{'opcode':18s} {'warp instr':>12s} {'%':>6s} {'thr/unit':>9s} {'samples %':>9s}")
    for op, v in per.most_common(top):
        u = f"{v * 32 / units:9.2f}" if units else ""
        print(f"{op:18s} {v:12d} {v / tot * 100:6.2f} {u} {per_s[op] / tot_s * 100:9.2f}")
    stall = collections.Counter()
    for i in insts:
        for k, v in i["stalls"].items():
            stall[k] += v
    st = sum(stall.values()) or 1
    print("\nstall samples:", ", ".join(f"{k[6:]} {v / st * 100:.1f}%" for k, v in stall.most_common(10)))
    if "--regions" in sys.argv:
        # runs of instructions with similar execution counts (a loop body has one count throughout)
        print("\nregions (consecutive instructions, split where the execution count changes by more than 2x):")
        start = 0
        regs = []
        for j in range(1, len(insts) + 1):
            if j == len(insts) or not (0.5 <= (insts[j]["exec"] + 1) / (insts[j - 1]["exec"] + 1) <= 2.0):
                seg = insts[start:j]
                regs.append((start, j, sum(x["exec"] for x in seg), sum(x["samples"] for x in seg)))
                start = j
        for a, b, e, sm in regs:
            if e / tot < 0.004:
                continue
            seg = insts[a:b]
            ops = collections.Counter(opcode(x["sass"]) for x in seg)
            u = f"{e * 32 / units:7.1f}/unit" if units else ""
            print(f"  [{a:5d},{b:5d}) n={b - a:4d} exec/inst={e // (b - a):9d} {e / tot * 100:5.1f}% {u} samples {sm / tot_s * 100:5.1f}%  "
                  + " ".join(f"{k}:{v}" for k, v in ops.most_common(6)))


if __name__ == "__main__":
    main()
