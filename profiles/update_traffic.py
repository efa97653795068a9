#!/usr/bin/env python
"""Record the DRAM traffic of one launch of the dominant kernel from an
`ncu --set full` report into profiles/traffic.json (read by bench.py's
`roofline.traffic`).  Run in the build container (ncu -i needs no GPU):

    python profiles/update_traffic.py REPORT.ncu-rep NSTRUCT N_PAIRS MODE [KERNEL_REGEX]
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, nstruct, n_pairs, mode = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    rx = re.compile(sys.argv[5] if len(sys.argv) > 5 else "actdist")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ix = {k: i for i, k in enumerate(hdr)}
    for row in rows[2:]:
        if not rx.search(row[ix["Kernel Name"]]):
            continue
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(row[ix[m]]) * UNIT[units[ix[m]]]
        entry = {"nstruct": nstruct, "n_pairs": n_pairs, "mode": mode, "kernel": row[ix["Kernel Name"]],
                 "dram_bytes_per_launch": tot, "kernel_ms_under_ncu": float(row[ix["gpu__time_duration.sum"]]) *
                 {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[units[ix["gpu__time_duration.sum"]]],
                 "report": os.path.basename(rep)}
        k = "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"
        if k in ix:
            entry["l1tex_lsu_data_pipe_pct"] = float(row[ix[k]])
        if "smsp__inst_executed.sum" in ix:
            entry["warp_instructions_per_launch"] = float(row[ix["smsp__inst_executed.sum"]])
        path = os.path.join(HERE, "traffic.json")
        data = json.load(open(path)) if os.path.exists(path) else {"captures": []}
        data["captures"] = [e for e in data["captures"]
                            if (e["nstruct"], e["n_pairs"], e["mode"]) != (nstruct, n_pairs, mode)] + [entry]
        json.dump(data, open(path, "w"), indent=1)
        print(json.dumps(entry))
        return
    raise SystemExit("no kernel matching %r in %s" % (rx.pattern, rep))


if __name__ == "__main__":
    main()
