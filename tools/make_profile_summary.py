"""Turn the ncu outputs of one gpurun call into the tracked summaries under profiles/.

  python tools/make_profile_summary.py <tag> <launches.csv> <raw.csv>

  <launches.csv>  ncu --metrics gpu__time_duration.sum --clock-control none ... --csv --log-file (bench.py under ncu)
  <raw.csv>       ncu -i prof.ncu-rep --page raw --csv                                     (--set full capture)
Writes profiles/<tag>_launches.md (per-kernel launch count, device time, share of the step),
profiles/<tag>_ncu_full.md (per profiled launch: duration, registers, occupancy, DRAM bytes, L2/L1 hit rates, pipes,
top stall reasons) and refreshes profiles/ncu_traffic.json (dram read+write bytes per launch, read by bench.py).
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches_csv, raw_csv = sys.argv[1:4]
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)


def short(name):
    n = name.replace("void ", "").replace("sicp::", "")
    return n.split("(")[0]


# ---------------------------------------------------------------- launch list
rows = list(csv.reader(l for l in open(launches_csv) if not l.startswith("==")))
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(r[iu], 1.0)
    a = agg.setdefault(short(r[ik]), [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
with open(os.path.join(ROOT, "profiles", f"{tag}_launches.md"), "w") as f:
    cmd = os.environ.get("PROFILE_CMD", "python bench.py --steps 2 --warmup 1 --no-cpu-baseline")
    f.write(f"# {tag}: launch list of `{cmd}` under ncu\n\n")
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch device times are cold-cache and serialised;\n"
            "compare SHARES, not absolutes.  Includes warm-up, the device-resident and the host (e2e) legs and the profiled single registration.\n"
            "The batch legs launch each LM solve on 37 CTAs (sized for 8 registrations in flight); ncu serialises the launches, so an LM launch\n"
            "runs alone on a quarter of the GPU and its share here (≈ 64 %) is above the live one (bench.py's per-stage events of a lone\n"
            "registration on 148 CTAs: LM ≈ 54 % of the four stages; marginal cost inside a batch, tools/sweep.py: ≈ 40 %).\n\n")
    f.write("| kernel | launches | total device time (us) | share |\n|---|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |\n")
    f.write(f"| total | {sum(v[0] for v in agg.values())} | {tot:.1f} | 100% |\n")

# ---------------------------------------------------------------- full capture
rows = list(csv.reader(open(raw_csv)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
stall = [h for h in hdr if "average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
ikn = hdr.index("Kernel Name")


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


traffic = collections.defaultdict(list)
with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full.md"), "w") as f:
    f.write(f"# {tag}: `ncu --set full --clock-control none --import-source on` on one KITTI-shaped EM-ICP registration (tools/probe_one.py)\n\n")
    f.write("One block per profiled launch, in launch order.  Times are under the profiler (cold caches, serialised) and are never bench values.\n")
    for r in rows[2:]:
        name = short(r[ikn])
        f.write(f"\n## `{name}`\n\n| metric | value |\n|---|---|\n")
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                f.write(f"| {w} | {r[i]} {units[i]} |\n")
        st = []
        for h in stall:
            v = num(r[hdr.index(h)])
            if v is not None and v >= 0.3:
                st.append((v, h.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        if st:
            f.write("| top stall reasons (warps stalled per issue-active cycle) | " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:6]) + " |\n")
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        rd, wr = num(r[ir]), num(r[iw])
        if rd is not None and wr is not None:
            traffic[name].append(to_bytes(rd, units[ir]) + to_bytes(wr, units[iw]))

stage = {"lm_kernel": "lm", "self_knn_pca_kernel": "cov", "cross_knn_kernel": "knn", "estep_kernel": "estep"}
out = {"_source": f"profiles/{tag}_ncu_full.md (dram__bytes_read.sum + dram__bytes_write.sum, mean per launch)"}
for name, vals in traffic.items():
    for pref, key in stage.items():
        if name.startswith(pref):
            out[key] = sum(vals) / len(vals)
with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote profiles/%s_launches.md, profiles/%s_ncu_full.md, profiles/ncu_traffic.json" % (tag, tag))
