"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a text file under profiles/
and a traffic JSON that bench.py attaches to its roofline object.
    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_name
"""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keep = ("Kernel Name", "Grid Size", "Block Size", "gpu__time_duration", "dram__bytes", "dram__throughput",
        "gpu__dram_throughput", "dram__cycles_active", "sm__warps_active", "launch__registers", "launch__occupancy",
        "sm__throughput", "lts__t_bytes.sum", "lts__t_sector_hit_rate", "l1tex__t_bytes", "smsp__average_warp",
        "sm__inst_executed.sum", "launch__shared_mem", "launch__waves", "sm__pipe_tensor", "smsp__cycles_active.avg")
traffic = {}
with open(out + "_summary.txt", "w") as f:
    for r in rows[2:]:
        f.write("=" * 100 + "\n")
        rec = dict(zip(hdr, r))
        for i, h in enumerate(hdr):
            if any(h.startswith(k) for k in keep):
                f.write(f"{h} [{units[i]}] = {r[i]}\n")

        def val(name):
            i = hdr.index(name)
            v, u = float(r[i].replace(",", "")), units[i]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)

        name = rec["Kernel Name"].split("<")[0].replace("void ", "").strip()
        traffic[name] = {"dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
                         "duration_us_under_ncu": float(rec["gpu__time_duration.sum"]),
                         "dram_pct_of_peak": float(rec["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]),
                         "kernel": rec["Kernel Name"], "grid": rec.get("Grid Size")}
json.dump(traffic, open(out + "_traffic.json", "w"), indent=1)
print(json.dumps(traffic, indent=1))
