#!/usr/bin/env python
"""Aggregate an ncu SASS-level source page (ncu -i X.ncu-rep --page source --csv --kernel-name ...)
per CUDA source line, using the line table of the cubin (nvdisasm -g -c <cubin>).

    python tools/ncu_lines.py <sass_csv> <nvdisasm_listing> <mangled-name-substring> [top] [--by-samples]

Prints, for the chosen kernel, the source lines ranked by executed warp instructions (or, with --by-samples, by
stall samples, each with its three leading stall reasons).  Needs -lineinfo at compile time
(torchani_b200/build.py has it).  The listing: cuobjdump -xelf all libani_b200.so; nvdisasm -g -c mlp.sm_100a.cubin."""
import csv
import re
import sys


def line_table(listing, func):
    """address -> (file, line) for the function whose section name contains `func`."""
    table, cur, inside = {}, None, False
    inl = None
    for ln in open(listing, errors="replace"):
        if ln.startswith("//--------------------- .text."):
            inside = func in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            inl = m.group(3)
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    by_samples = "--by-samples" in sys.argv
    argv = [a for a in sys.argv if a != "--by-samples"]
    sass_csv, listing, func = argv[1:4]
    top = int(argv[4]) if len(argv) > 4 else 40
    rows = list(csv.reader(open(sass_csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    table = line_table(listing, func)
    base = None
    agg, total_i, total_s = {}, 0, 0
    for r in rows[hi + 1:]:
        if len(r) <= ii or not r[ia]:
            continue
        try:
            addr = int(r[ia], 16)
        except ValueError:
            continue
        if base is None:
            base = addr
        key, _ = table.get(addr - base, (("?", 0), ""))
        n = int(float(r[ii] or 0))
        s = int(float(r[isamp] or 0))
        a = agg.setdefault(key, [0, 0, 0, {}])
        a[0] += n
        a[1] += s
        a[2] += 1
        for i, h in stall_cols:
            v = int(float(r[i] or 0)) if i < len(r) else 0
            if v:
                a[3][h[6:]] = a[3].get(h[6:], 0) + v
        total_i += n
        total_s += s
    print(f"total warp instructions {total_i}, samples {total_s}, sass lines mapped {len(table)}")
    for key, (n, s, c, st) in sorted(agg.items(), key=lambda kv: -kv[1][1 if by_samples else 0])[:top]:
        why = " ".join(f"{h}={v}" for h, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{key[0]:>20s}:{key[1]:<5d} inst {n:>10d} ({100.0 * n / max(total_i, 1):5.1f}%)  samples {s:>7d} "
              f"({100.0 * s / max(total_s, 1):5.1f}%)  sass {c:<4d} {why}")


if __name__ == "__main__":
    main()
