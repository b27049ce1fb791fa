"""Summarise an .ncu-rep (one kernel) into the few numbers DESIGN.md / bench.py quote.
python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.avg",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum"]
print(f"# {rep}")
for k in want:
    for h in d:
        if h == k or h.endswith("." + k):
            print(f"{k:78s} {d[h][1]:>18s} {d[h][0]}")
            break
try:
    el = float(d[[h for h in d if h.endswith("sm__cycles_elapsed.avg")][0]][1].replace(",", ""))
    act = float(d[[h for h in d if h.endswith("hmma_cycles_active_realtime.avg")][0]][1].replace(",", ""))
    print(f"{'tensor pipe (hmma) active / elapsed cycles':78s} {act / el:18.3f}")
except Exception as e:  # noqa
    print("tensor active fraction: n/a", e)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print(f"\n# warp-state samples: {tot}; SASS mnemonics proving the Blackwell path:")
for m in ("UTCHMMA", "LDTM", "UTCBAR", "LDGSTS", "SYNCS.PHASECHK", "UTMALDG"):
    n = sum(1 for r in data if m in r[ix["Source"]])
    ex = sum(int(r[ix["Instructions Executed"]]) for r in data if m in r[ix["Source"]])
    print(f"  {m:16s} static {n:5d}   executed (warp-level) {ex}")
print("\n# top stalled instructions (samples, executed, SASS)")
for i, r in sorted(enumerate(data), key=lambda x: -int(x[1][ix["# Samples"]]))[:12]:
    print(f"  {r[ix['# Samples']]:>6s} {r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:90]}")
