"""Summarise an .ncu-rep (read here, no GPU needed) into profiles/<name>.md + .json.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_scan [--traffic-config N]
--traffic-config N also records the scan kernel's DRAM bytes per launch in profiles/traffic.json under "configN":
bench.py reports that number as roofline.traffic for BASELINE config N."""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v) * m.get(unit, 1)


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    res = []
    for r in data:
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    d[k] = float(r[i].replace(",", ""))
                except ValueError:
                    d[k] = r[i]
                d[k + "@unit"] = units[i]
        if "dram__bytes_read.sum" in d:
            d["dram_bytes_per_launch"] = to_bytes(d["dram__bytes_read.sum"], d["dram__bytes_read.sum@unit"]) + \
                to_bytes(d["dram__bytes_write.sum"], d["dram__bytes_write.sum@unit"])
        res.append(d)
    json.dump(res, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write(f"# ncu summary of `{rep}` (--set full --clock-control none)\n\n")
        for d in res:
            f.write(f"## {d['kernel']}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| `{k}` | {d[k]} | {d[k + '@unit']} |\n")
            if "dram_bytes_per_launch" in d:
                f.write(f"| dram bytes per launch (read+write) | {d['dram_bytes_per_launch']:.4g} | byte |\n")
            f.write("\n")
    print(json.dumps([{k: v for k, v in d.items() if "@" not in k} for d in res], indent=1)[:3000])


def record_traffic(out, config):
    import os
    res = json.load(open(out + ".json"))
    scan = [d for d in res if "k_pass1" in d["kernel"] and "dram_bytes_per_launch" in d]
    if not scan:
        return
    path = os.path.join(os.path.dirname(os.path.abspath(out)), "traffic.json")
    try:
        t = json.load(open(path))
        if "k_pass1_dram_bytes_per_launch" in t:  # round-1 layout
            t = {}
    except Exception:
        t = {}
    t[f"config{config}"] = {"k_pass1_dram_bytes_per_launch": scan[0]["dram_bytes_per_launch"], "kernel": scan[0]["kernel"],
                            "duration_ms_under_ncu": scan[0].get("gpu__time_duration.sum"),
                            "source": os.path.basename(out) + ".json (tools/ncu_summary.py from the .ncu-rep of `ncu --set full "
                                      f"--clock-control none ... python bench.py --config {config} --steps 1 --warmup 3 --no-e2e "
                                      "--no-cpu-baseline --no-hot-spin`)"}
    json.dump(t, open(path, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
    if "--traffic-config" in sys.argv:
        record_traffic(sys.argv[2], int(sys.argv[sys.argv.index("--traffic-config") + 1]))
