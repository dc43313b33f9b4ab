"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into a small JSON: per profiled launch the duration, DRAM traffic,
achieved DRAM GB/s, and the utilisation figures the roofline discussion in DESIGN.md quotes.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.json [--alg name=bytes ...]"""
import csv, io, json, subprocess, sys

WANT = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed": "xu_pipe_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_elapsed": "fp64_pipe_pct",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__cycles_elapsed.max.per_second": "sm_ghz",
}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6,
         "Ghz": 1, "Mhz": 1e-3}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    alg = dict(a.split("=") for a in sys.argv[3:] if "=" in a and not a.startswith("--"))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0]}
        for k, name in WANT.items():
            if k in hdr:
                i = hdr.index(k)
                try:
                    d[name] = float(r[i].replace(",", "")) * SCALE.get(units[i], 1)
                except ValueError:
                    pass
        if "dram_read" in d:
            d["dram_traffic_bytes"] = d.pop("dram_read") + d.pop("dram_write")
            d["dram_GBps"] = d["dram_traffic_bytes"] / d["time_us"] / 1e3
        for k, v in alg.items():
            if k in d["kernel"]:
                d["algorithmic_bytes"] = float(v)
                d["algorithmic_GBps"] = float(v) / d["time_us"] / 1e3
        res.append(d)
    json.dump({"source": rep, "note": "ncu --set full --clock-control none; per-launch, cold-cache, serialised", "launches": res},
              open(out, "w"), indent=1)
    for d in res:
        print(d["kernel"][:28].ljust(28), f'{d.get("time_us", 0):9.1f} us', f'{d.get("dram_GBps", 0):8.0f} GB/s dram',
              f'{d.get("dram_pct_of_peak", 0):5.1f}% dram', f'{d.get("sm_pct_of_peak", 0):5.1f}% sm', f'{d.get("fp64_pipe_pct", 0):5.1f}% fp64',
              f'{d.get("tensor_pipe_pct", 0):5.1f}% tensor', f'{d.get("xu_pipe_pct", 0):5.1f}% xu', f'regs {d.get("registers", 0):.0f}')


main()
