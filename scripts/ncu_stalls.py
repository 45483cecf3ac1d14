#!/usr/bin/env python
"""Warp-stall samples of one kernel by code region, from the source page of an `.ncu-rep` (run where ncu is installed).

  python scripts/ncu_stalls.py <file.ncu-rep> <lo>:<hi>:<name> [...]     regions are the low 24 bits of the SASS address
  python scripts/ncu_stalls.py <file.ncu-rep> --top 40                   the instructions with the most samples
"""
import collections
import csv
import io
import subprocess
import sys


def load(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    out = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        out.append((int(r[idx["Address"]], 16) & 0xFFFFFF, r[idx["Source"]].strip(), int(r[idx["# Samples"]] or 0),
                    int(r[idx["Instructions Executed"]] or 0), {c[6:]: int(r[idx[c]] or 0) for c in cols}))
    return rows[0][1], out


def main():
    kernel, ins = load(sys.argv[1])
    total = sum(i[2] for i in ins)
    print(f"# {sys.argv[1]}\n# {kernel}\n# {total} warp samples over all warps of the CTA")
    if sys.argv[2] == "--top":
        for a, src, n, ie, st in sorted(ins, key=lambda t: -t[2])[:int(sys.argv[3])]:
            top = [(k, v) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v]
            print(f"{n:7d} {100 * n / total:5.1f}%  executed {ie:9d}  {a:06x}  {src[:64]:64s} {top}")
        return
    for spec in sys.argv[2:]:
        lo, hi, name = spec.split(":")
        lo, hi = int(lo, 16), int(hi, 16)
        n = ie = 0
        c = collections.Counter()
        for a, _, ns, nie, st in ins:
            if lo <= a < hi:
                n += ns
                ie += nie
                c.update(st)
        print(f"{name:22s} {n:6d} samples {100 * n / total:5.1f}%  {ie:10d} warp instructions   " +
              " ".join(f"{k}={v}" for k, v in c.most_common(6) if v))


if __name__ == "__main__":
    main()
