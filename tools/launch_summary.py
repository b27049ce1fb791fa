"""Per-family / per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
python tools/launch_summary.py profiles/x_launches.csv "title line" > profiles/x_launches_summary.txt"""
import collections, csv, re, sys

FAMILIES = [
    ("norm family", r"^norm_"),
    ("TMA + tcgen05 row GEMM (1x1x1 conv / linear / deconv k2s2)", r"^rowgemm_tma"),
    ("tcgen05 deep GEMM conv (TMA halo tiles / cp.async)", r"^conv_gemm_(tma|tc)"),
    ("tcgen05 conv fwd/dgrad", r"^conv3_tc"),
    ("tcgen05 wgrad", r"^wgrad3_tc|^wgrad_gemm_tc"),
    ("wgrad reduce", r"^wgrad_reduce"),
    ("mma.sync igemm/pw/legacy wgrad/pack", r"^igemm|^pw_conv|^wgrad_kernel|^pack_weight"),
    ("DSA + LayerNorm", r"^dsa_|^ln_|^tile_group_sum|^keep_scale"),
    ("loss + outconv", r"^loss_|^outconv|^tv_"),
]
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr, data = None, []
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))


def short(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", n)
    m = re.match(r"([\w:]+)", n)
    return (m.group(1) if m else n).split("(")[0]


per = collections.OrderedDict()
for d in data:
    k = short(d["Kernel Name"])
    us = float(d["Metric Value"].replace(",", "")) / 1000.0
    c = per.setdefault(k, [0, 0.0])
    c[0] += 1
    c[1] += us
total = sum(v[1] for v in per.values())
print(f"# {sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]}")
print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES, not absolutes)")
print(f"# {sum(v[0] for v in per.values())} launches, {total / 1000:.2f} ms of kernel time\n\n## by family")
fam = collections.OrderedDict((f, [0, 0.0]) for f, _ in FAMILIES)
fam["other (pool, upsample, copies, layout, dropout masks, optimizer, sw, post-process)"] = [0, 0.0]
for k, (n, us) in per.items():
    for f, pat in FAMILIES:
        if re.search(pat, k.split("::")[-1]):
            fam[f][0] += n; fam[f][1] += us
            break
    else:
        f = list(fam)[-1]
        fam[f][0] += n; fam[f][1] += us
for f, (n, us) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print(f"{f:82s} {n:5d} launches {us:10.1f} us   {us / total:.3f}")
print("\n## by kernel\nkernel                                               launches         us   share")
for k, (n, us) in sorted(per.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[-50:]:50s} {n:10d} {us:10.1f}   {us / total:.3f}")
