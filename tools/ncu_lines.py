#!/usr/bin/env python
"""Aggregate an ncu SASS-level source page (ncu -i X.ncu-rep --page source --csv --kernel-name ...)
per CUDA source line, using the line table of the cubin (nvdisasm -g -c <cubin>).

    python tools/ncu_lines.py <sass_csv> <nvdisasm_listing> <mangled-name-substring> [top]

Prints, for the chosen kernel, the source lines ranked by executed warp instructions together
with their stall samples.  Needs -lineinfo at compile time (torchani_b200/build.py has it)."""
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
    sass_csv, listing, func = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(sass_csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
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
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += n
        a[1] += s
        a[2] += 1
        total_i += n
        total_s += s
    print(f"total warp instructions {total_i}, samples {total_s}, sass lines mapped {len(table)}")
    for key, (n, s, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{key[0]:>14s}:{key[1]:<5d} inst {n:>10d} ({100.0 * n / max(total_i, 1):5.1f}%)  samples {s:>7d} "
              f"({100.0 * s / max(total_s, 1):5.1f}%)  sass {c}")


if __name__ == "__main__":
    main()
