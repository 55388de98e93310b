#!/usr/bin/env python
"""Summarise an `ncu --set full` report: the metrics DESIGN.md / bench.py quote, one JSON object per kernel launch.
usage: tools/ncu_summary.py report.ncu-rep out.json [--traffic profiles/dominant_kernel_traffic.json]"""
import csv, io, json, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "lts__t_sectors_op_red.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units = rows[0], rows[1]
    res, traffic = [], {}
    for r in rows[2:]:
        d = {"kernel": r[h.index("Kernel Name")]}
        for w in WANT:
            if w in h:
                d[w] = f"{r[h.index(w)]} {units[h.index(w)]}".strip()
        try:
            b = sum(float(r[h.index(k)].replace(",", "")) * UNIT[units[h.index(k)]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            d["dram_bytes_total"] = int(b)
            for key in ("raster_forward", "raster_backward"):
                if key in d["kernel"]:
                    traffic[key] = int(b)
        except Exception:
            pass
        res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    if "--traffic" in sys.argv:
        p = sys.argv[sys.argv.index("--traffic") + 1]
        traffic["_source"] = f"ncu --set full, {out} (dram__bytes_read.sum + dram__bytes_write.sum per launch, C4 view)"
        json.dump(traffic, open(p, "w"), indent=1)
    print(json.dumps(res, indent=1)[:3000])


if __name__ == "__main__":
    main()
