#!/usr/bin/env python
"""`ncu -i X.ncu-rep --page raw --csv` -> the physical-side facts bench.py attaches to its rooflines (profiles/rNN_traffic.json).
usage: ncu_facts.py raw.csv out.json name:kernel_regex:launch_index:elements ..."""
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name, scale_to=None):
    v = r[col[name]].replace(",", "")
    if v in ("", "no data"):
        return None
    v = float(v)
    u = units[col[name]]
    if scale_to == "byte":
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    if scale_to == "us":
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)   # ncu's "us" column may come as another unit
    return v


out = {"kernels": {}}
for spec in sys.argv[3:]:
    name, rx, idx, elems = spec.split(":")
    sel = [r for r in data if re.search(rx, r[col["Kernel Name"]])]
    r = sel[int(idx)]
    dram = val(r, "dram__bytes_read.sum", "byte") + val(r, "dram__bytes_write.sum", "byte")
    sc = {"issue_slots_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
          "alu_pipe_pct": val(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
          "fma_pipe_pct": val(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
          "xu_pipe_pct": val(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
          "sm_throughput_pct": val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
          "lanes_per_instruction": val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
          "achieved_occupancy_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
          "l1_hit_pct": val(r, "l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": val(r, "lts__t_sector_hit_rate.pct"),
          "l1tex_throughput_pct": val(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
          "l2_throughput_pct": val(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
          "dram_throughput_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
          "warp_instructions_per_element": val(r, "smsp__inst_issued.sum") / float(elems),
          "registers_per_thread": val(r, "launch__registers_per_thread"),
          "ncu_duration_us": val(r, "gpu__time_duration.sum", "us")}
    out["kernels"][name] = {"dram_bytes_per_element": dram / float(elems), "elements": int(elems),
                            "kernel": r[col["Kernel Name"]].split("(")[0],
                            "source": f"profiles/{sys.argv[1].split('/')[-1]}: launch {idx} of /{rx}/ (ncu --set full): dram__bytes_read.sum + "
                                      f"dram__bytes_write.sum = {dram / 1e6:.1f} MB for {elems} elements",
                            "secondary_ceilings": sc}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
