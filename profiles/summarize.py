#!/usr/bin/env python
"""Turn an Nsight Compute report into the compact text summaries kept under
profiles/ (run in the build container: `ncu -i` needs no GPU).

    python profiles/summarize.py gpurun_out/prof.ncu-rep profiles/r01_actdist_ncu.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__lsu_writeback_active.sum.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    lines = ["# ncu summary of `%s`" % rep.split("/")[-1], ""]
    raw = ncu_csv(rep, "raw")
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        lines += ["## " + name, "", "| metric | unit | value |", "|---|---|---|"]
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                lines.append("| %s | %s | %s |" % (w, units[i], row[i]))
        lines.append("")
    src = ncu_csv(rep, "source")
    h = src[1]
    data = [r for r in src[2:] if len(r) >= len(h)]
    ix = {k: i for i, k in enumerate(h)}
    stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    tot = collections.Counter()
    nsamp = ninst = 0
    ops = collections.Counter()
    for r in data:
        nsamp += int(r[ix["# Samples"]] or 0)
        n = int(r[ix["Instructions Executed"]] or 0)
        ninst += n
        for s in stalls:
            tot[s] += int(r[ix[s]] or 0)
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
        if m:
            ops[m.group(2)] += n
    lines += ["## warp stall sampling (source page, first profiled launch)", "",
              "warp instructions executed: %d, samples: %d" % (ninst, nsamp), "",
              "| stall | samples | share |", "|---|---|---|"]
    for s, v in tot.most_common(10):
        lines.append("| %s | %d | %.1f%% |" % (s, v, 100.0 * v / max(1, nsamp)))
    lines += ["", "## dynamic SASS opcode mix", "", "| opcode | warp instructions | share |", "|---|---|---|"]
    for o, v in ops.most_common(20):
        lines.append("| %s | %d | %.1f%% |" % (o, v, 100.0 * v / max(1, ninst)))
    top = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:12]
    lines += ["", "## most-stalled instructions", "", "| samples | executed | SASS | main stall |", "|---|---|---|---|"]
    for r in top:
        st = {s: int(r[ix[s]] or 0) for s in stalls}
        lines.append("| %s | %s | `%s` | %s |" % (r[ix["# Samples"]], r[ix["Instructions Executed"]],
                                                    r[ix["Source"]].strip(), max(st, key=st.get)))
    open(dst, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
