"""profiles/roofline_traffic.json from the raw pages of the round-2 `ncu --set full` captures (tools/ncu_round3.sh; round-2 mid-point captures: tools/ncu_round2.sh, profiles/r02f_*):

    python tools/roofline_traffic.py

For each kernel class bench.py reports a roofline for: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum), L2 bytes per
launch (lts__t_bytes.sum), the L2 / L1 / DRAM throughput ncu rates against its own peaks, hit rates and lane utilisation — all averaged
over the captured launches and weighted by launch duration where it is a rate."""
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "%": 1.0, "": 1.0, "sector": 1.0}


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, r):
            try:
                d[h] = float(v.replace(",", "")) * SCALE.get(u, 1.0)
            except ValueError:
                d[h] = v
        out.append(d)
    return out


def summarise(launches, source, what, note, bound):
    t = [l["gpu__time_duration.sum"] for l in launches]
    tot = sum(t)
    w = lambda key: sum(l[key] * ti for l, ti in zip(launches, t) if isinstance(l.get(key), float)) / tot
    dram = sum(l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"] for l in launches)
    l2 = sum(l["lts__t_bytes.sum"] for l in launches)
    n = len(launches)
    return {"kernel": what, "launches": n, "limited_by": bound, "dram_bytes_per_launch": dram / n, "l2_bytes_per_launch": l2 / n,
            "ms_per_launch_under_ncu": tot / n,
            "dram_gbs_under_ncu": dram / (tot * 1e-3) / 1e9, "l2_gbs_under_ncu": l2 / (tot * 1e-3) / 1e9,
            "l2_frac": w("lts__throughput.avg.pct_of_peak_sustained_elapsed") / 100.0,
            "l1_frac": w("l1tex__throughput.avg.pct_of_peak_sustained_elapsed") / 100.0,
            "dram_frac": w("dram__throughput.avg.pct_of_peak_sustained_elapsed") / 100.0 if isinstance(launches[0].get("dram__throughput.avg.pct_of_peak_sustained_elapsed"), float)
            else w("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") / 100.0,
            "sm_frac": w("sm__throughput.avg.pct_of_peak_sustained_elapsed") / 100.0,
            "l1_hit_rate": w("l1tex__t_sector_hit_rate.pct") / 100.0, "l2_hit_rate": w("lts__t_sector_hit_rate.pct") / 100.0,
            "lanes_per_instruction": w("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "warps_active_frac": w("sm__warps_active.avg.pct_of_peak_sustained_active") / 100.0,
            "source": source, "note": note}


LIM = "L1 data pipe (l1tex__data_pipe_lsu_wavefronts: scattered per-lane node fetches) and instruction issue; DRAM and L2 are far from their peaks"


def main():
    P = lambda f: os.path.join(ROOT, "profiles", f)
    out = {}
    c5 = load(P("r03k_ncu_full_closest_c5_raw.csv"))
    out["c5_path_closest"] = summarise(c5, "profiles/r03k_ncu_full_closest_c5_raw.csv (ncu --set full --clock-control none; the 96 closest-hit launches of one c5_path step, 32 spp)",
                                       "k_trace_closest_engine + k_trace_mis_engine", "the 5 M-triangle scene (640 MB of collapsed nodes + geometry) does not fit the L2: DRAM serves part of "
                                       "the tree besides the ray / hit records; the kernel is bound by the L1 data pipe (l1_frac) and instruction issue, not by DRAM", LIM)
    c3 = load(P("r03k_ncu_full_closest_c3_raw.csv"))
    out["c3_path_closest"] = summarise(c3, "profiles/r03k_ncu_full_closest_c3_raw.csv (the 12 closest-hit launches of one c3_path step, 8 spp)", "k_trace_closest_engine + k_trace_mis_engine",
                                       "the 1 M-triangle scene (130 MB) is mostly L2-resident: DRAM moves the ray / hit records and the first touch of the tree", LIM)
    c4 = load(P("r03k_ncu_full_c4_raw.csv"))
    closest = [l for l in c4 if "closest" in str(l["Kernel Name"])][-1:]
    anyhit = [l for l in c4 if "anyhit" in str(l["Kernel Name"])][-1:]
    out["c4_closest"] = summarise(closest, "profiles/r03k_ncu_full_c4_raw.csv (second 16 Mi-ray launch of k_closest_batch_engine against 10,014,720 triangles)", "k_closest_batch_engine",
                                  "HBM-resident config: 1.3 GB of nodes + geometry; rays binned by origin cell and direction octant before the launch", LIM)
    out["c4_anyhit"] = summarise(anyhit, "profiles/r03k_ncu_full_c4_raw.csv (second 16 Mi-ray launch of k_anyhit_batch_engine)", "k_anyhit_batch_engine", "as c4_closest", LIM)
    json.dump(out, open(P("roofline_traffic.json"), "w"), indent=1)
    for k, v in out.items():
        print(f"{k:18s} dram {v['dram_bytes_per_launch'] / 1e6:9.1f} MB/launch  l2 {v['l2_bytes_per_launch'] / 1e6:9.1f} MB/launch  l2_frac {v['l2_frac']:.2f} l1_frac {v['l1_frac']:.2f} "
              f"dram_frac {v['dram_frac']:.2f} sm {v['sm_frac']:.2f}  L1 hit {v['l1_hit_rate']:.2f} L2 hit {v['l2_hit_rate']:.2f} lanes {v['lanes_per_instruction']:.1f}")


if __name__ == "__main__":
    main()
