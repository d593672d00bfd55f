"""Turn the ncu exports in gpurun_out/ into the small tracked summaries under profiles/."""
import csv, glob, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.environ.get("PROFILES_OUT") or os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)
TAG = sys.argv[1] if len(sys.argv) > 1 else "r1"
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum"]


def to_bytes(v, unit):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit)
    return float(v) * f if f else float(v)


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2:]


traffic = {}
for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"{TAG}_*.ncu-rep"))):
    hdr, units, data = raw(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    name = os.path.basename(rep)[:-8]
    wl = name.rsplit("_", 1)[-1]
    with open(os.path.join(OUT, name + "_summary.csv"), "w", newline="") as f:
        w = csv.writer(f)
        cols = ["Kernel Name"] + [k for k in KEYS if k in idx]
        w.writerow(cols)
        w.writerow([""] + [units[idx[k]] for k in cols[1:]])
        for r in data:
            w.writerow([r[idx["Kernel Name"]][:60]] + [r[idx[k]] for k in cols[1:]])
            kn = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
            if "dram__bytes_read.sum" in idx:
                tb = to_bytes(r[idx["dram__bytes_read.sum"]].replace(",", ""), units[idx["dram__bytes_read.sum"]]) + \
                     to_bytes(r[idx["dram__bytes_write.sum"]].replace(",", ""), units[idx["dram__bytes_write.sum"]])
                traffic.setdefault(wl, {})[kn] = tb
    print("wrote", name + "_summary.csv")
if traffic:
    old = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            old = json.load(f)
    except Exception:  # noqa: BLE001
        pass
    for wl, d in traffic.items():
        old.setdefault(wl, {}).update(d)
    traffic = old
    with open(os.path.join(OUT, "traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1, sort_keys=True)
for lst in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"{TAG}_*launches*.csv"))):
    rows = list(csv.reader(open(lst)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    cols = rows[h]
    ki, vi, ui = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Metric Unit")
    with open(os.path.join(OUT, os.path.basename(lst)), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["launch", "kernel", "gpu__time_duration.sum", "unit"])
        for n, r in enumerate(rows[h + 1:]):
            if len(r) > vi:
                w.writerow([n, r[ki][:70], r[vi], r[ui]])
    print("wrote", os.path.basename(lst))
