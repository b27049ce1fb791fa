"""One line per launch of an .ncu-rep (--set full capture): duration, DRAM bytes read/written, achieved DRAM GB/s against
the measured copy peak, L2 bytes, occupancy, registers, and sm__pipe_tensor_subpipe_hmma_cycles_active_realtime / elapsed
cycles -- on this part a WORK counter (executed MACs / 1024 per cycle per SM, DESIGN.md section 3.1: x 0.25 = fraction of
the nominal 4096 MAC/clk/SM dense bf16 rate), not a busy counter.    python tools/ncu_table.py x.ncu-rep [y.ncu-rep ...]"""
import csv, json, os, subprocess, sys

PEAK = 6539.9
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    PEAK = json.load(open(p))["hbm_gbs"]


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return float("nan")


def col(hdr, name):
    for i, h in enumerate(hdr):
        if h == name or h.endswith(name):
            return i
    return None


out = {}
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {k: col(hdr, k) for k in ("Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
                                   "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
                                   "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                                   "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
                                   "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg")}
    if ix["lts__t_bytes.sum"] is None:
        ix["lts__t_bytes.sum"] = col(hdr, "lts__t_bytes.sum.per_second")

    def scaled(r, key):
        i = ix[key]
        if i is None:
            return float("nan")
        v = num(r[i])
        u = units[i].lower()
        for pre, m in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("byte", 1.0), ("msecond", 1e-3), ("usecond", 1e-6),
                       ("nsecond", 1e-9), ("second", 1.0), ("ms", 1e-3), ("us", 1e-6), ("ns", 1e-9), ("s", 1.0)):
            if u.startswith(pre):
                return v * m
        return v
    print(f"# {rep}  (ncu --set full --clock-control none; DRAM GB/s = (read + written) / duration, peak {PEAK} GB/s measured copy)")
    print(f"{'kernel':34s} {'grid':>14s} {'us':>8s} {'rd MB':>8s} {'wr MB':>8s} {'GB/s':>7s} {'of peak':>7s} {'L2 MB':>8s} {'warps%':>6s} {'regs':>4s} {'MAC/1024/cyc/SM':>16s}")
    for r in data:
        import re
        mm = re.search(r"(\w+_kernel\w*)", r[ix["Kernel Name"]])
        name = (mm.group(1) if mm else r[ix["Kernel Name"]])[-34:]
        t = scaled(r, "gpu__time_duration.sum")
        rd, wr = scaled(r, "dram__bytes_read.sum"), scaled(r, "dram__bytes_write.sum")
        l2 = scaled(r, "lts__t_bytes.sum")
        gbs = (rd + wr) / t / 1e9
        print(f"{name:34s} {r[ix['Grid Size']].replace(' ', ''):>14s} {t * 1e6:8.1f} {rd / 1e6:8.1f} {wr / 1e6:8.1f} {gbs:7.0f} "
              f"{gbs / PEAK:7.2f} {l2 / 1e6:8.1f} {num(r[ix['sm__warps_active.avg.pct_of_peak_sustained_active']]):6.1f} "
              f"{r[ix['launch__registers_per_thread']]:>4s} "
              f"{(num(r[ix['sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg']]) / num(r[ix['sm__cycles_elapsed.avg']])) if ix['sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg'] is not None and ix['sm__cycles_elapsed.avg'] is not None else float('nan'):16.3f}")
        key = name
        out.setdefault(key, []).append(dict(grid=r[ix["Grid Size"]], us=t * 1e6, dram_bytes=rd + wr, source=os.path.basename(rep)))
    print()
if os.environ.get("NCU_JSON"):
    json.dump(out, open(os.environ["NCU_JSON"], "w"), indent=1)
