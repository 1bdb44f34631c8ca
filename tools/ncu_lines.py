#!/usr/bin/env python
"""Per-source-line instruction / stall attribution for one kernel of an .ncu-rep.

    python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX CUBIN_NAME [--top N] [--inner]

`ncu --page source --csv` only exports the SASS view with metrics; this joins it with
`nvdisasm -gi` line info of the in-tree library (built with -lineinfo) so the executed
instruction count and the stall samples are summed per CUDA source line.  By default a SASS
instruction is charged to the OUTERMOST line of its inline chain inside the kernel's own file
set (what the kernel body calls); --inner charges the innermost (the callee's line).
"""
import argparse
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "point_geometric_features_b200", "libpgeof_b200.so")


def sass_rows(rep, regex, launch=0):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + regex],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None and "Source" in r:
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] is not None and len(r) >= len(cur["hdr"]) - 2:
            cur["rows"].append(r)
    if not blocks:
        sys.exit("no kernel matched")
    return blocks[launch]


def line_map(cubin_name, mangled_hint):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, stdout=subprocess.DEVNULL, check=True)
    cubin = [f for f in os.listdir(tmp) if f.startswith(cubin_name)][0]
    dis = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, text=True).stdout.splitlines()
    funcs, cur, chain = {}, None, []
    for ln in dis:
        m = re.match(r"^\.text\.(\S+):", ln)
        if m:
            cur = funcs.setdefault(m.group(1), [])
            chain = []
            continue
        if cur is None:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            chain.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
        if m:
            if chain:
                cur.append((int(m.group(1), 16), list(chain), m.group(2).strip()))
                last = list(chain)
            else:
                cur.append((int(m.group(1), 16), last if cur else [], m.group(2).strip()))
            chain = []
    return funcs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel")
    ap.add_argument("cubin")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--inner", action="store_true")
    ap.add_argument("--launch", type=int, default=0)
    ap.add_argument("--sass", action="store_true", help="list the hottest SASS instructions too")
    a = ap.parse_args()
    blk = sass_rows(a.report, a.kernel, a.launch)
    hdr = blk["hdr"]
    ia, isrc, iexec, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    funcs = line_map(a.cubin, blk["name"])
    # match the function by instruction count + opcode sequence
    n = len(blk["rows"])
    cand = [(k, v) for k, v in funcs.items() if len(v) == n]
    if not cand:
        cand = sorted(funcs.items(), key=lambda kv: abs(len(kv[1]) - n))[:1]
        print("warning: no exact-length match (%d SASS rows); using %s with %d" % (n, cand[0][0], len(cand[0][1])), file=sys.stderr)
    if len(cand) > 1:
        op0 = [r[isrc].split()[0] for r in blk["rows"][:200]]
        cand = [c for c in cand if [x[2].split()[0].lstrip("@!P0123456789 ") for x in c[1][:200]] == [o for o in op0]] or cand
    name, ins = cand[0]
    print("kernel:", blk["name"][:120])
    print("matched function:", name[-90:], "(%d instructions)" % len(ins))
    per = collections.OrderedDict()
    tot_e = tot_s = 0
    hot = []
    for r, (off, chain, text) in zip(blk["rows"], ins):
        e, s = int(r[iexec] or 0), int(r[isamp] or 0)
        tot_e += e
        tot_s += s
        key = (chain[0] if a.inner else chain[-1]) if chain else ("?", 0)
        d = per.setdefault(key, [0, 0, collections.Counter()])
        d[0] += e
        d[1] += s
        for i in stall_cols:
            v = int(r[i] or 0)
            if v:
                d[2][hdr[i][6:]] += v
        hot.append((s, e, off, text, key))
    warps = int(blk["rows"][0][iexec] or 1)
    print("total warp-instructions %d (%.1f per warp at entry), samples %d" % (tot_e, tot_e / max(warps, 1), tot_s))
    print("%-28s %10s %7s %7s  %s" % ("file:line", "inst/warp", "inst%", "samp%", "top stalls"))
    for key, (e, s, st) in sorted(per.items(), key=lambda kv: -kv[1][1])[: a.top]:
        print("%-28s %10.1f %6.1f%% %6.1f%%  %s" % ("%s:%d" % key, e / warps, 100.0 * e / max(tot_e, 1), 100.0 * s / max(tot_s, 1),
                                                  " ".join("%s=%d" % kv for kv in st.most_common(3))))
    if a.sass:
        print("\nhottest SASS:")
        for s, e, off, text, key in sorted(hot, reverse=True)[: a.top]:
            print("%6.2f%% %9.1f  /*%04x*/ %-60s %s:%d" % (100.0 * s / max(tot_s, 1), e / warps, off, text[:60], key[0], key[1]))


if __name__ == "__main__":
    main()
