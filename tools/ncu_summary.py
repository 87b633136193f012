#!/usr/bin/env python
"""Summarise .ncu-rep files (read here, no GPU needed) into a small text table for profiles/."""
import csv
import io
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__waves_per_multiprocessor"]


def summarise(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return f"{path}: no kernels captured\n"
    hdr, units = rows[0], rows[1]
    txt = ""
    for vals in rows[2:]:
        txt += f"== {path}\n"
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                txt += f"  {w:<68} {vals[i]} {units[i]}\n"
    return txt




def traffic_json(path, out_json):
    """dram bytes (read + write) per launch of the first captured kernel -> profiles/ncu_traffic.json"""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]

    def get(name):
        i = hdr.index(name)
        v = float(vals[i])
        u = units[i].lower()
        mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        return v * mult
    d = {"kernel": vals[hdr.index("Kernel Name")], "fps_dram_bytes_per_launch": get("dram__bytes_read.sum") + get("dram__bytes_write.sum"),
         "source": path, "note": "one ncu --set full capture of the FPS kernel inside bench.py (cfg2)"}
    json.dump(d, open(out_json, "w"), indent=1)
    print(d)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--traffic":
        traffic_json(sys.argv[2], sys.argv[3])
    else:
        for p in sys.argv[1:]:
            print(summarise(p))
