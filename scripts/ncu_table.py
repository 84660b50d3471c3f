"""Turn an `ncu --set full` report into the small per-launch table + family totals kept under profiles/.
    python scripts/ncu_table.py <report.ncu-rep> <out.csv> <out.json> "<command that was profiled>"
(runs `ncu -i <report> --page raw --csv`; run it where the report lives, reports are too big to bring back)"""
import csv, io, json, re, subprocess, sys

rep, out_csv, out_json, cmd = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
units = rows[1]
data = rows[2:]


def col(name):
    for i, h in enumerate(hdr):
        if h == name:
            return i
    for i, h in enumerate(hdr):
        if h.startswith(name):
            return i
    return None


def num(r, i, scale_units=True):
    if i is None or r[i] in ("", "n/a"):
        return float("nan")
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    if scale_units:
        if u in ("kbyte", "kb"): v *= 1e3
        elif u in ("mbyte", "mb"): v *= 1e6
        elif u in ("gbyte", "gb"): v *= 1e9
        elif u in ("ns", "nsecond"): v *= 1e-3  # -> us
        elif u in ("ms", "msecond"): v *= 1e3
        elif u in ("s", "second"): v *= 1e6
    return v


c = {k: col(v) for k, v in {
    "name": "Kernel Name", "grid": "Grid Size", "dur": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum",
    "wr": "dram__bytes_write.sum", "l2sm": "lts__t_sectors_srcunit_tex_op_read.sum",
    "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "lts": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "regs": "launch__registers_per_thread"}.items()}
if c["tensor"] is None:
    c["tensor"] = col("sm__pipe_tensor_cycles_active")
if c["dram"] is None:
    c["dram"] = col("gpu__dram_throughput")
tot = {"launches": 0, "dram_bytes_read": 0.0, "dram_bytes_write": 0.0, "duration_us_serialised": 0.0, "tw": 0.0}
with open(out_csv, "w", newline="") as f:
    f.write(f"# {cmd}\n# per launch, in launch order; kernel template args joined with '/'; cold-cache, serialised (ncu replays)\n")
    w = csv.writer(f)
    w.writerow(["kernel", "grid_ctas", "duration_us", "dram_read_MB", "dram_write_MB", "l2_to_sm_read_MB", "tensor_pipe_active_pct",
                "lts_throughput_pct", "dram_throughput_pct", "registers"])
    for r in data:
        name = re.sub(r"\(.*", "", r[c["name"]]).replace("void ", "").replace("ttb::", "").replace(", ", "/")
        grid = r[c["grid"]].strip("()").replace(", ", "x")
        g = 1
        for p in grid.split("x"):
            g *= int(p)
        dur, rd, wr = num(r, c["dur"]), num(r, c["rd"]), num(r, c["wr"])
        l2 = num(r, c["l2sm"], False) * 32 if c["l2sm"] is not None else float("nan")
        tp = num(r, c["tensor"], False)
        w.writerow([name, g, f"{dur:.1f}", f"{rd / 1e6:.1f}", f"{wr / 1e6:.1f}", f"{l2 / 1e6:.1f}", f"{tp:.1f}",
                    f"{num(r, c['lts'], False):.1f}", f"{num(r, c['dram'], False):.1f}", r[c["regs"]] if c["regs"] is not None else ""])
        tot["launches"] += 1
        tot["dram_bytes_read"] += rd
        tot["dram_bytes_write"] += wr
        tot["duration_us_serialised"] += dur
        tot["tw"] += tp * dur
out = {"what": "sum over the captured launches (see the csv beside this file)", "command": cmd, "launches": tot["launches"],
       "dram_bytes_read": tot["dram_bytes_read"], "dram_bytes_write": tot["dram_bytes_write"],
       "dram_bytes_total": tot["dram_bytes_read"] + tot["dram_bytes_write"],
       "duration_us_serialised_cold": tot["duration_us_serialised"],
       "time_weighted_tensor_pipe_active_pct": tot["tw"] / max(tot["duration_us_serialised"], 1e-9)}
json.dump(out, open(out_json, "w"), indent=1)
print(json.dumps(out))
