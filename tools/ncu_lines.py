#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source cuda,sass` output: samples per source line and per region."""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    cur = None
    hdr = None
    data = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            si = hdr.index("# Samples")
            ii = hdr.index("Instructions Executed")
            stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            wsx = hdr.index("L1 Wavefronts Shared Excessive") if "L1 Wavefronts Shared Excessive" in hdr else None
            continue
        if hdr is None or r[0] == "" or r[0] in ("Function Name",):
            continue
        try:
            ln = int(r[0])
            s = int(r[si])
            n = int(r[ii])
        except ValueError:
            continue
        stalls = sorted(((int(r[i] or 0), h) for i, h in stall_cols), reverse=True)[:2]
        data.append((s, n, cur, ln, r[1][:100], stalls, int(r[wsx] or 0) if wsx is not None else 0))
    tot = sum(d[0] for d in data) or 1
    toti = sum(d[1] for d in data) or 1
    print("total samples", tot, "instructions", toti)
    for d in sorted(data, reverse=True)[:top]:
        st = " ".join(f"{h[6:]}={c}" for c, h in d[5] if c)
        print(f"{100*d[0]/tot:5.2f}% inst {100*d[1]/toti:5.2f}% bankx={d[6]:9d} {d[2]}:{d[3]:<4d} {d[4]}  [{st}]")
    return data, tot


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
