#!/usr/bin/env python
"""Per-source-line instruction accounting of one kernel from an .ncu-rep captured with --import-source on.

usage: python tools/ncu_lines.py REPORT.ncu-rep [--warps W] [--iters I] [--top N] [--sass-order]

Reads `ncu -i REPORT --page source --csv --print-source cuda,sass`, attributes every SASS instruction to its CUDA
line, and prints warp-instructions per line (optionally normalised per warp and per loop iteration) with the average
number of active lanes.  --sass-order lists runs of consecutive SASS instructions that map to the same line in
address order (the control-flow picture).  Needs only ncu on the host: no GPU.
"""
import argparse, csv, subprocess, sys, collections

def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    insts = {}  # address -> (file, line, src, sass, executed, threads)
    fname = None
    hdr = None
    cur = None
    for r in rows:
        if not r:
            continue
        if r[0] in ("File Path", "File Name"):
            fname = r[1].split("/")[-1]; continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}; continue
        if hdr is None or len(r) < 10:
            continue
        if r[0] != "":
            cur = (fname, int(r[0]), r[1].strip()); continue
        if r[2].startswith("0x"):
            ie = int(r[7]); th = float(r[10] or 0)
            insts[int(r[2], 16)] = (cur[0], cur[1], cur[2], r[3].strip(), ie, th)
    return insts

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep"); ap.add_argument("--warps", type=float, default=1.0); ap.add_argument("--iters", type=float, default=1.0)
    ap.add_argument("--top", type=int, default=60); ap.add_argument("--sass-order", action="store_true")
    a = ap.parse_args()
    insts = load(a.rep)
    norm = a.warps * a.iters
    tot = sum(v[4] for v in insts.values())
    print(f"# {len(insts)} SASS instructions, {tot} warp-instructions executed ({tot / norm:.1f} per unit)")
    if a.sass_order:
        run = None
        for ad in sorted(insts):
            f, ln, src, sass, ie, th = insts[ad]
            if run and run[0] == (f, ln):
                run[1] += ie; run[2] += ie * th; run[3] += 1
            else:
                if run: emit(run, norm)
                run = [(f, ln), ie, ie * th, 1, src]
        emit(run, norm)
        return
    per = collections.defaultdict(lambda: [0, 0.0, 0, ""])
    for f, ln, src, sass, ie, th in insts.values():
        p = per[(f, ln)]; p[0] += ie; p[1] += ie * th; p[2] += 1; p[3] = src
    for (f, ln), (ie, thw, n, src) in sorted(per.items(), key=lambda kv: -kv[1][0])[: a.top]:
        print(f"{f:12s} {ln:5d} sass {n:4d} exec {ie / norm:8.1f} ({100 * ie / tot:4.1f}%) lanes {thw / max(ie, 1):5.1f}  {src[:90]}")

def emit(run, norm):
    (f, ln), ie, thw, n, src = run
    if ie == 0: return
    print(f"{f:12s} {ln:5d} sass {n:3d} exec {ie / norm:7.1f} lanes {thw / max(ie, 1):5.1f}  {src[:80]}")

if __name__ == "__main__":
    main()
