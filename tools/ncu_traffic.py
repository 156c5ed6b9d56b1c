"""Read `ncu -i <file>.ncu-rep --page raw --csv` output(s) and write profiles/r2_ncu_traffic.json + a text summary:
per profiled launch the kernel name, duration, DRAM bytes read / written, tensor-pipe activity and achieved DRAM bandwidth.
    ncu -i gpurun_out/r2_prof_infer.ncu-rep --page raw --csv > gpurun_out/r2_prof_infer_raw.csv
    python tools/ncu_traffic.py gpurun_out/r2_prof_infer_raw.csv [more.csv ...]"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {"gpu__time_duration.sum": "time_us", "dram__bytes_read.sum": "dram_read_MB", "dram__bytes_write.sum": "dram_write_MB",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg": "hmma_cycles_active", "sm__cycles_elapsed.max": "cycles_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak", "launch__registers_per_thread": "regs",
        "launch__grid_size": "grid", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "lts__t_sector_hit_rate.pct": "l2_hit_pct", "smsp__inst_executed.sum": "inst_executed"}
UNIT = {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def parse(path):
    lines = [ln for ln in open(path, newline="") if not ln.startswith("==")]
    rows = list(csv.reader(lines))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        rec = {"kernel": re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("hrnb::", "").replace("void ", "")}
        for i, name in enumerate(hdr):
            if name in WANT and i < len(r) and r[i] != "":
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                rec[WANT[name]] = v * UNIT.get(units[i], 1.0)
        if "time_us" in rec and "dram_read_MB" in rec:
            rec["dram_GBps"] = (rec["dram_read_MB"] + rec.get("dram_write_MB", 0.0)) / rec["time_us"] * 1e3
        if "hmma_cycles_active" in rec and rec.get("cycles_elapsed"):
            rec["hmma_active_frac"] = rec["hmma_cycles_active"] / rec["cycles_elapsed"]
        out.append(rec)
    return out


if __name__ == "__main__":
    allrec = {}
    for p in sys.argv[1:]:
        allrec[os.path.basename(p)] = parse(p)
    # bench.py's `roofline.traffic`: DRAM bytes (read + write) of ONE launch of the dominant tensor-pipe kernel - the first
    # conv_tc record of the inference / training capture = the 3x3 32->32 conv at 64x64 (21 % of the network's MACs)
    for name, recs in list(allrec.items()):
        kind = "conv_tc_kernel_infer" if "infer" in name else ("conv_tc_kernel_train" if "train" in name else None)
        first = next((r for r in recs if r["kernel"].startswith("conv_tc") and "dram_read_MB" in r), None)
        if kind and first:
            allrec[kind] = int((first["dram_read_MB"] + first.get("dram_write_MB", 0.0)) * 1e6)
            allrec[kind + "_note"] = "dram__bytes_read.sum + dram__bytes_write.sum of one %s launch (%.1f us), 3x3 32->32 conv at 64x64, from %s" % (
                first["kernel"], first.get("time_us", 0.0), name)
    with open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json"), "w") as f:
        json.dump(allrec, f, indent=1)
    for name, recs in allrec.items():
        if not isinstance(recs, list):
            continue
        print("#", name)
        for r in recs:
            print("%-46s %8.1f us  read %8.1f MB  write %8.1f MB  %7.0f GB/s  hmma %.3f  regs %s grid %s" % (
                r["kernel"][:46], r.get("time_us", 0), r.get("dram_read_MB", 0), r.get("dram_write_MB", 0), r.get("dram_GBps", 0),
                r.get("hmma_active_frac", 0), int(r.get("regs", 0)), int(r.get("grid", 0))))
