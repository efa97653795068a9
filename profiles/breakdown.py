#!/usr/bin/env python
"""Per-function instruction / stall breakdown of a K1 kernel from an `ncu --import-source on`
report (run in the build container; needs no GPU):

    python profiles/breakdown.py REPORT.ncu-rep KERNEL_MANGLED_PREFIX N_PAIRS

The SASS page of the report (`ncu -i REPORT --page source --csv`) is joined by instruction
offset with `nvdisasm --print-line-info` of the kernel in igm_b200/libigmk.so (same build),
and every instruction is attributed to the function of igmk_actdist.cuh / igmk_actdist_list.cuh whose line range
holds the most recent line of those files seen in address order (inlined helpers of
igmk_device.cuh therefore count for their caller).
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def function_ranges(path):
    """(first_line, last_line, name) of the top-level functions of igmk_actdist.cuh."""
    names = []
    for n, line in enumerate(open(path), 1):
        m = re.match(r"^(?:__device__|__global__|template|struct|static)\b.*", line)
        if not m:
            continue
        nm = re.search(r"(\w+)\s*\(", line) or re.search(r"struct\s+(\w+)", line)
        if line.startswith("template") or nm is None:
            continue
        names.append((n, nm.group(1)))
    out = []
    for k, (n, nm) in enumerate(names):
        out.append((n, names[k + 1][0] - 1 if k + 1 < len(names) else 10 ** 9, nm))
    return out


def main():
    rep, kern, npairs = sys.argv[1], sys.argv[2], int(sys.argv[3])
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "igm_b200", "libigmk.so")],
                   cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "--print-line-info", "-c", cubin], cwd=tmp,
                          capture_output=True, text=True).stdout.split("\n")
    start = [n for n, l in enumerate(sass) if l.startswith(".text." + kern)][0]
    ins, cur = [], ("?", 0)
    for l in sass[start + 1:]:
        if l.startswith(".text.") or l.startswith(".section"):
            break
        m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", l)
        if m:
            ins.append((int(m.group(1), 16), cur[0], cur[1], m.group(2).strip()))
    page = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(page)))
    h = rows[1]
    ix = {k: i for i, k in enumerate(h)}
    data = rows[2:]
    base = int(data[0][0], 16)
    prof = {int(r[0], 16) - base: r for r in data if r and r[0].startswith("0x")}
    if len(prof) != len(ins):
        sys.stderr.write("warning: %d profiled vs %d disassembled instructions (different build?)\n"
                         % (len(prof), len(ins)))
    ranges = {f: function_ranges(os.path.join(ROOT, "igm_b200", "csrc", f))
              for f in ("igmk_actdist.cuh", "igmk_actdist_list.cuh")}

    def func(f, ln):
        for a, b, nm in ranges[f]:
            if a <= ln <= b:
                return nm
        return "other"
    stallcols = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    byf, bys = collections.Counter(), collections.Counter()
    byop, st = collections.defaultdict(collections.Counter), collections.defaultdict(collections.Counter)
    tot = tots = 0
    cur = "kernel"
    for off, f, ln, txt in ins:
        r = prof.get(off)
        if r is None:
            continue
        n = int(r[ix["Instructions Executed"]] or 0)
        s = int(r[ix["# Samples"]] or 0)
        if f in ranges:
            cur = func(f, ln)
        byf[cur] += n
        bys[cur] += s
        tot += n
        tots += s
        op = txt.split()[1] if txt.startswith("@") else txt.split()[0]
        byop[cur][op.split(".")[0]] += n
        for c in stallcols:
            st[cur][c] += int(r[ix[c]] or 0)
    print("kernel `%s`, %d pairs: %.1f warp instructions per pair, %d stall samples\n" % (kern, npairs, tot / npairs, tots))
    print("| function | inst / pair | share | stall samples | top opcodes (per pair) | top stalls (share of all samples) |")
    print("|---|---|---|---|---|---|")
    for k, v in byf.most_common():
        print("| %s | %.1f | %.1f %% | %.1f %% | %s | %s |" % (
            k, v / npairs, 100.0 * v / tot, 100.0 * bys[k] / max(1, tots),
            " ".join("%s:%.0f" % (o, c / npairs) for o, c in byop[k].most_common(7)),
            " ".join("%s:%.1f%%" % (c[6:], 100.0 * x / max(1, tots)) for c, x in st[k].most_common(3))))


if __name__ == "__main__":
    main()
