#!/usr/bin/env python
"""Extract per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of the captured kernels
into profiles/r2_ncu_traffic.json, which bench.py quotes as roofline.traffic.
    python tools/ncu_traffic.py gpurun_out/prof_european_r2.ncu-rep gpurun_out/prof_trajectory_r2.ncu-rep"""
import csv, io, json, subprocess, sys
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("mcb::", "").split("<")[0]
        rd = float(d["dram__bytes_read.sum"]) * UNIT[u["dram__bytes_read.sum"]]
        wr = float(d["dram__bytes_write.sum"]) * UNIT[u["dram__bytes_write.sum"]]
        out[name] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "traffic": rd + wr,
                     "gpu_time_us": float(d["gpu__time_duration.sum"]) * {"us": 1, "ms": 1e3, "ns": 1e-3}[u["gpu__time_duration.sum"]],
                     "grid": d["Grid Size"], "block": d["Block Size"], "source": rep.split("/")[-1]}
json.dump(out, open("profiles/r2_ncu_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
