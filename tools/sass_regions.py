"""Dynamic instruction mix of one kernel from an ncu report, split at the block barriers.

usage: ncu -i rep.ncu-rep --page source --csv --kernel-name regex:<name> > /tmp/src.csv
       python tools/sass_regions.py /tmp/src.csv <frames-processed>  [--list REGION]
"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    frames = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    show = int(sys.argv[sys.argv.index("--list") + 1]) if "--list" in sys.argv else None
    h = rows[1]
    si, ei = h.index("Source"), h.index("Instructions Executed")
    region, regs = 0, collections.OrderedDict()
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":           # next kernel of a multi-kernel export
            break
        if len(r) <= ei or not r[ei] or not r[ei].isdigit():
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[si])
        op = m.group(2) if m else "?"
        n = int(r[ei])
        regs.setdefault(region, collections.Counter())[op] += n
        if show == region and n and not re.search(r"\b(FADD|FMUL|FFMA)\b", r[si]):
            print(f"{n:8d} {r[si][:110]}")
        if op == "BAR":
            region += 1
    tot = sum(sum(c.values()) for c in regs.values())
    print(f"total warp-instructions {tot} = {tot / frames:.1f} per frame")
    for k, c in regs.items():
        t = sum(c.values())
        fp = c["FADD"] + c["FMUL"] + c["FFMA"]
        mem = c["LDS"] + c["STS"] + c["LDG"] + c["STG"]
        print(f"region {k}: {t / frames:7.1f}/frame ({100 * t / tot:4.1f}%)  fp {fp / frames:6.1f}  mem {mem / frames:6.1f}  "
              f"other {(t - fp - mem) / frames:6.1f}   " + ", ".join(f"{o}:{n / frames:.1f}" for o, n in c.most_common(8)))


if __name__ == "__main__":
    main()
