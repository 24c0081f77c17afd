#!/usr/bin/env python
"""Stall samples of one kernel by SOURCE LINE, from an ncu report that was captured with --import-source on and a library built
with -lineinfo: `ncu --page source --csv` only exports the SASS view (per-instruction samples), so the instructions are matched,
in order, with `nvdisasm -g` of the same cubin, whose `//## File ..., line N [inlined at ...]` comments give each instruction's line.

    cuobjdump -xelf all raw-physics_b200/librawphys_b200.so ; nvdisasm -g -c rp_batch.sm_100a.cubin > dis.txt
    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:k_gjk > src.csv
    python profiles/sass_lines.py dis.txt src.csv _ZN2rp5k_gjkENS_7DevViewE [top]
"""
import collections
import csv
import re
import sys


def disasm_lines(path, func):
    """[(opcode text, innermost (file, line), outermost line)] of one function"""
    out, on = [], False
    cur, outer = ("?", 0), 0
    for ln in open(path):
        s = ln.strip()
        if s.startswith(".text."):
            on = s == ".text.%s:" % func
            continue
        if not on:
            continue
        m = re.match(r'//## File "([^"]+)", line (\d+)(.*)', s)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            chain = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            outer = int(chain[-1][1]) if chain else cur[1]
            continue
        m = re.match(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", s)
        if m:
            out.append((m.group(2), cur, outer))
    return out


def main():
    dis, src, func = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    ins = disasm_lines(dis, func)
    rows = list(csv.reader(open(src)))
    hdr = rows[1]
    si, so = hdr.index("# Samples"), hdr.index("Source")
    data = [r for r in rows[2:] if len(r) > si]
    if len(data) != len(ins):
        print("warning: %d instructions in the report, %d in the disassembly" % (len(data), len(ins)))
    by_line, by_outer, total = collections.Counter(), collections.Counter(), 0
    for r, (op, cur, outer) in zip(data, ins):
        n = int(r[si] or 0)
        total += n
        by_line[cur] += n
        by_outer[outer] += n
    print("%s: %d samples, %d instructions" % (func, total, len(ins)))
    print("-- by innermost source line")
    for (f, l), n in by_line.most_common(top):
        print("%6.2f%%  %s:%d" % (100.0 * n / max(total, 1), f, l))
    print("-- by line of the kernel body (inlined callees folded in)")
    for l, n in by_outer.most_common(top // 2):
        print("%6.2f%%  kernel line %d" % (100.0 * n / max(total, 1), l))


if __name__ == "__main__":
    main()
