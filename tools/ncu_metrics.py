#!/usr/bin/env python
"""Extract the per-point DRAM bytes and shared-memory wavefronts of the projection kernel from an .ncu-rep into
profiles/project_ncu_metrics.json (read by bench.py for roofline.traffic / roofline.smem_lsu).
usage: tools/ncu_metrics.py rep.ncu-rep n_points source_label [build_label]"""
import csv, io, json, os, subprocess, sys
rep, npts, label = sys.argv[1], int(sys.argv[2]), sys.argv[3]
build = sys.argv[4] if len(sys.argv) > 4 else subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
for v in rows[2:]:
    if "project4_kernel" in v[h.index("Kernel Name")]:
        break
else:
    raise SystemExit("no project4_kernel launch in " + rep)
def get(name):
    i = h.index(name)
    x = float(v[i].replace(",", ""))
    unit = u[i].lower()
    return x * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
out = {"kernel": v[h.index("Kernel Name")], "points": npts,
       "dram_bytes_per_point": (get("dram__bytes_read.sum") + get("dram__bytes_write.sum")) / npts,
       "smem_wavefronts_per_point": get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / npts,
       "build": build, "source": label}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "project_ncu_metrics.json")
json.dump(out, open(path, "w"))
print(out)
