#!/usr/bin/env python
"""Where the warps of one kernel WAIT: stall samples per CUDA source line from an .ncu-rep captured with
--set full --import-source on.

usage: python tools/ncu_stalls.py REPORT.ncu-rep [--top N] [--reason long_sb] [--sass]

Per line: share of all stall samples, executed warp-instructions, and the line's top stall reasons.  --reason restricts
the ranking to one reason (long_sb, short_sb, wait, branch_resolving, mio, no_inst, barrier, math, lg, not_selected ...);
--sass lists single SASS instructions instead of lines.  Needs only ncu on the host: no GPU.
"""
import argparse, collections, csv, subprocess

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep"); ap.add_argument("--top", type=int, default=40); ap.add_argument("--reason", default=None)
    ap.add_argument("--sass", action="store_true")
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    hdr = None; cur = None; fname = None
    per = collections.defaultdict(lambda: [0, collections.Counter(), ""])  # key -> [executed, stalls, text]
    reasons = []
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] in ("File Path", "File Name"):
            fname = r[1].split("/")[-1]; continue
        if r[0] == "Line No":
            hdr = r
            reasons = [(i, h[len("stall_"):]) for i, h in enumerate(r) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or len(r) < 10:
            continue
        if r[0] != "":
            cur = (fname, int(r[0]), r[1].strip()); continue
        if r[2].startswith("0x") and cur:
            key = (cur[0], cur[1], r[2]) if a.sass else (cur[0], cur[1])
            e = per[key]
            e[0] += int(r[7] or 0)
            e[2] = (r[3].strip() + "   // " + cur[2][:60]) if a.sass else cur[2]
            for i, name in reasons:
                v = r[i]
                if v and v != "0":
                    e[1][name] += int(float(v))
    total = collections.Counter()
    for e in per.values():
        total.update(e[1])
    allsamples = sum(total.values())
    print(f"# {allsamples} stall samples; by reason: " + ", ".join(f"{k}={v / allsamples:.1%}" for k, v in total.most_common(10)))
    score = (lambda e: e[1][a.reason]) if a.reason else (lambda e: sum(e[1].values()))
    denom = total[a.reason] if a.reason else allsamples
    for key, e in sorted(per.items(), key=lambda kv: -score(kv[1]))[:a.top]:
        top = ", ".join(f"{k}={v}" for k, v in e[1].most_common(3))
        print(f"{score(e) / max(denom, 1):6.1%}  exec {e[0]:>9}  {key[0]}:{key[1]:<5} {top:<50} {e[2][:110]}")

if __name__ == "__main__":
    main()
