#!/usr/bin/env python
"""Read an ncu report (ncu -i ... --page raw --csv) here, without a GPU, and write the per-kernel summary the bench and the
docs cite: duration, DRAM bytes read/written, achieved DRAM throughput, issue utilisation, registers, shared memory.

    python scripts/ncu_summary.py gpurun_out/<tag>_all_kernels.ncu-rep profiles/<tag>_ncu_summary.csv [lib_sha16] [workload_bytes] [level]

Also writes profiles/traffic_<kernel>.json (dram_bytes_read / dram_bytes_write per launch + the sha256 prefix of the libzling.so
the capture was taken from) — bench.py refuses a traffic figure whose sha differs from the library it runs."""
import csv
import io
import json
import os
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
sha = sys.argv[3] if len(sys.argv) > 3 else None          # "<lib sha16>[:<source sha16>]"
src_sha = None
if sha and ":" in sha:
    sha, src_sha = sha.split(":", 1)
wbytes = int(sys.argv[4]) if len(sys.argv) > 4 else None
level = int(sys.argv[5]) if len(sys.argv) > 5 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "smsp__inst_executed.sum"]
want = [w for w in want if w in col]


def scale(v, u):
    v = float(v.replace(",", ""))
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "usecond": 1e-3, "msecond": 1, "second": 1e3, "nsecond": 1e-6,
                    "us": 1e-3, "ms": 1, "s": 1e3, "ns": 1e-6}.get(u, 1)


with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "duration_ms", "dram_read_bytes", "dram_write_bytes"] + want[4:])
    seen = set()
    for r in data:
        name = r[col["Kernel Name"]].split("(")[0].replace("zl::", "").replace("void ", "")
        dur = scale(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]])
        rd = scale(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = scale(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        w.writerow([name, round(dur, 4), int(rd), int(wr)] + [r[col[k]] for k in want[4:]])
        if sha and name not in seen:
            seen.add(name)
            short = name.split("<")[0]
            with open(os.path.join(os.path.dirname(out), "traffic_%s.json" % short), "w") as g:
                json.dump({"kernel": name, "lib_sha16": sha, "src_sha16": src_sha, "workload_bytes": wbytes, "level": level, "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
                           "duration_ms": round(dur, 4), "source": "%s (ncu --set full --clock-control none, first launch of the kernel)" % os.path.basename(out)}, g, indent=1)
print("wrote", out)
