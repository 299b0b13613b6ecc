"""Extract per-launch DRAM traffic / duration of the captured kernels from `ncu --set full` reports into
profiles/ncu_traffic.json (read by bench.py for roofline.traffic).
usage: python tools/ncu_traffic.py profiles/ncu_traffic.json name=report.ncu-rep [name=report ...]"""
import csv
import io
import json
import subprocess
import sys

out_path = sys.argv[1]
res = {}
for spec in sys.argv[2:]:
    name, rep = spec.split("=", 1)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    def col(metric):
        return hdr.index(metric)
    launches = []
    for r in rows[2:]:
        def val(metric, scale_unit=True):
            i = col(metric)
            v = float(r[i].replace(",", ""))
            u = units[i]
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6,
                    "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}.get(u, 1)
            return v * mult
        launches.append(dict(kernel=r[col("Kernel Name")][:80], grid=r[col("Grid Size")],
                             dram_read=val("dram__bytes_read.sum"), dram_write=val("dram__bytes_write.sum"),
                             us=val("gpu__time_duration.sum"),
                             tensor_pct=float(r[col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")] or 0)
                             if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in hdr else None))
    n = len(launches)
    res[name] = dict(launches=launches, dram_bytes_per_launch=sum(l["dram_read"] + l["dram_write"] for l in launches) / max(n, 1),
                     avg_us=sum(l["us"] for l in launches) / max(n, 1))
json.dump(res, open(out_path, "w"), indent=1)
print(json.dumps({k: dict(dram_MB=v["dram_bytes_per_launch"] / 1e6, us=v["avg_us"], n=len(v["launches"])) for k, v in res.items()}))
