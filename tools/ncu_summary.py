"""ncu report -> compact JSON summary (one object per captured launch) for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_ncu_<name>.json"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max"]
STALL = "smsp__average_warps_issue_stalled_"


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                d[k] = f"{r[hdr.index(k)]} {units[hdr.index(k)]}".strip()
        stalls = {}
        for i, h in enumerate(hdr):
            if h.startswith(STALL) and h.endswith("_per_warp_active.pct") is False and h.endswith(".ratio"):
                name = h[len(STALL):].split("_per_")[0]
                try:
                    stalls[name] = round(float(r[i]), 3)
                except ValueError:
                    pass
        if stalls:
            d["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:9])
        res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    print(f"{out}: {len(res)} launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
