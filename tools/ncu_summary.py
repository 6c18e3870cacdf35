#!/usr/bin/env python
"""ncu -i <rep> --page raw --csv  ->  one JSON entry per captured launch with the counters the roofline argument uses."""
import csv, json, subprocess, sys
KEYS = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "xbar2l1_read",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__inst_executed.sum": "warp_instructions",
}
def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        e = {"kernel": r[hdr.index("Kernel Name")][:90]}
        for k, name in KEYS.items():
            if k in hdr:
                i = hdr.index(k)
                try:
                    e[name] = float(r[i].replace(",", ""))
                except ValueError:
                    e[name] = r[i]
                e[name + "_unit"] = units[i]
        res.append(e)
    json.dump(res, open(out, "w"), indent=1)
    for e in res:
        print({k: v for k, v in e.items() if not k.endswith("_unit")})
if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
