"""Per-kernel table from an `ncu --csv --metrics ...` log of scripts/prof_all_kernels.py (second round only).
usage: python scripts/summarize_kernels.py kernels.csv [hbm_peak_gbs]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 8]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6544.0
h = rows[0]
ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
launches = collections.OrderedDict()
for r in rows[1:]:
    launches.setdefault(int(r[ii]), {"name": r[ki].split("(")[0]})[r[mi]] = float(r[vi].replace(",", ""))
ids = sorted(launches)
ids = ids[len(ids) // 2:]  # the second round
agg = collections.OrderedDict()
for i in ids:
    L = launches[i]
    a = agg.setdefault(L["name"], collections.defaultdict(float))
    t = L["gpu__time_duration.sum"]
    a["n"] += 1
    a["t"] += t
    a["bytes"] += L["dram__bytes_read.sum"] + L["dram__bytes_write.sum"]
    for k, key in (("fma", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                   ("issue", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   ("tensor", "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"),
                   ("fp64", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                   ("hmma", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                   ("warps", "sm__warps_active.avg.pct_of_peak_sustained_active")):
        a[k] += L.get(key, 0.0) * t
print(f"| kernel | launches | time us | dram GB/s (read+write) | % of {peak:.0f} GB/s | FP32-FMA pipe active % | issue active % | fp16 tensor ops (tcgen05) % of peak | tensor pipe active % (incl. HMMA) | FP64 pipe % | warps active % |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
    t = a["t"]
    gbs = a["bytes"] / t  # bytes per ns = GB/s
    print(f"| {name[:60]} | {int(a['n'])} | {t / 1e3:.1f} | {gbs:.0f} | {100 * gbs / peak:.0f} | {a['fma'] / t:.1f} | {a['issue'] / t:.1f} | "
          f"{a['tensor'] / t:.1f} | {a['hmma'] / t:.1f} | {a['fp64'] / t:.1f} | {a['warps'] / t:.1f} |")
