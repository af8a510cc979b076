"""profiles/roofline_traffic.json from a raw ncu page: python tools/roofline_traffic.py RAW.csv SOURCE_NAME
   RAW.csv = `ncu -i X.ncu-rep --page raw --csv` of an `ncu --set full` capture of the closest-hit launches of ONE c3_path step."""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {k: hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot, per = 0.0, []
for r in data:
    b = sum(float(r[col[k]].replace(",", "")) * scale[units[col[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    per.append({"kernel": r[col["Kernel Name"]].split("(")[0].replace("void ", ""), "dram_bytes": b,
                "time": r[col["gpu__time_duration.sum"]] + " " + units[col["gpu__time_duration.sum"]]})
    tot += b
out = {"kernel": f"k_trace_closest_engine + k_trace_mis_engine (the {len(data)} closest-hit launches of one c3_path step, 8 spp)",
       "dram_bytes_per_launch": int(tot / len(data)), "launches": len(data),
       "source": f"profiles/{sys.argv[2]} (ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum averaged over the launches)",
       "note": "the 80 MB scene is L2-resident: DRAM moves the ray / hit / queue records and the first touch of the tree", "per_launch": per}
json.dump(out, open("profiles/roofline_traffic.json", "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "per_launch"}, indent=1))
for p in per:
    print(f"  {p['kernel']:40s} {p['dram_bytes'] / 1e6:9.1f} MB  {p['time']}")
