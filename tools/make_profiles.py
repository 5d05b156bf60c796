"""Turn the scratch outputs of tools/gpu_final.sh (gpurun_out/) into the tracked evidence under profiles/:
launch lists, per-kernel ncu metric summaries, bench lines and the measured DRAM traffic bench.py reports."""
import collections, csv, json, pathlib, shutil, subprocess, sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
G, P = ROOT / "gpurun_out", ROOT / "profiles"
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
csv.field_size_limit(10**9)

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__icc_request_hit_rate.pct", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum"]


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def stall_page(rep, pat):
    """Warp-stall samples per kernel from the SASS view of the source page (one block per captured launch, headed by a
    "Kernel Name" row; the stall_* columns are per instruction)."""
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "source", "--print-source", "sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    func, hdr, res = None, None, {}
    for r in rows:
        if len(r) >= 2 and r[0] == "Kernel Name": func = r[1]; continue
        if len(r) > 5 and r[0] == "Address": hdr = r; continue
        if hdr is None or func is None or len(r) < len(hdr) or pat not in func: continue
        agg = res.setdefault(func, collections.Counter())
        for h, v in zip(hdr, r):
            if h.startswith("stall_") and "Not" not in h and v.isdigit(): agg[h[6:]] += int(v)
    return res


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[start + 1:]:
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum": continue
        v = float(d["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[d["Metric Unit"]]
        tot[d["Kernel Name"]] += v; cnt[d["Kernel Name"]] += 1
    shutil.copy(src, dst)
    T = sum(tot.values())
    return [(k, cnt[k], v, 100 * v / T) for k, v in sorted(tot.items(), key=lambda x: -x[1])]


summary, traffic = [], {}
SETS = {"r01": (("config 2 (1-D pipelined kernels)", "prof_pipe_final.ncu-rep", "k1d_pipe<", "launches_c2.csv", "c2"),
                ("config 4 (cooperative PCG)", "prof_pcg_final.ncu-rep", "k_pcg", "launches_c4.csv", "c4")),
        "r02": (("config 2 (1-D pipelined kernels; default bench line)", "prof_pipe_r2.ncu-rep", "k1d_pipe<", "launches_c2.csv", "c2"),
                ("config 5a (sweep: forward + fused misfit adjoint)", "prof_c5a_r2.ncu-rep", "k1d_pipe<", "launches_c5a.csv", "c5a"),
                ("config 4 (multigrid-preconditioned CG, structured assembly)", "prof_c4_r2.ncu-rep", "k_", "launches_c4.csv", "c4"),
                ("config 5b (banded batch: block TRSM on the FP64 tensor cores, stencil load / gradient kernels)", "prof_c5b_r2.ncu-rep", "k_band", "launches_c5b.csv", "c5b"),
                ("config 2 variant c2e (per-sample per-element kappa, two-pass kernels)", "prof_c2e_r2.ncu-rep", "k1d_pe", "launches_c2e.csv", "c2e"))}
for name, rep, pat, src, key in SETS.get(TAG, SETS["r02"]):
    summary.append(f"## {name}\n")
    if (G / src).exists():
        summary.append(f"Launch list `{TAG}_{key}_launches.csv` (ncu --metrics gpu__time_duration.sum, serialised, cold cache):\n")
        for k, n, v, pct in launches(G / src, P / f"{TAG}_{key}_launches.csv")[:8]:
            summary.append(f"    {pct:5.1f} %  n={n:3d}  avg {v / n:10.1f} us  {k[:90]}")
        summary.append("")
    if (G / rep).exists():
        recs, units = raw_page(G / rep)
        stalls = stall_page(G / rep, pat)
        for d in recs:
            kn = d["Kernel Name"]
            summary.append(f"`{kn[:100]}` (ncu --set full --clock-control none):\n")
            for k in WANT:
                if k in d and d[k] not in ("", "n/a"): summary.append(f"    {k:90s} {d[k]} {units.get(k, '')}")
            norm = lambda x: x.replace("(bool)", "").replace("(int)", "").replace(" ", "").replace("void", "").split("(")[0]
            for f, agg in stalls.items():
                if norm(f)[:70] == norm(kn)[:70]:
                    t = sum(agg.values())
                    summary.append("    warp stall samples: " + ", ".join(f"{a} {100 * b / t:.1f}%" for a, b in agg.most_common(8)))
            summary.append("")
            rd, wr = float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            traffic[f"{key}:{kn[:110]}"] = rd * scale[units["dram__bytes_read.sum"]] + wr * scale[units["dram__bytes_write.sum"]]
(P / f"{TAG}_ncu_summary.md").write_text("# ncu summaries generated by tools/make_profiles.py from the round's final gpurun call\n\n" + "\n".join(summary) + "\n")
(P / f"{TAG}_traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")
lines = []
for f in ("bench", "bench_ref", "bench_c5a", "bench_c3", "bench_c4", "bench_c4_jacobi", "bench_c5b", "bench_c2e", "bench_n2", "bench_n4", "bench_n8",
          "bench_ref_n2", "bench_ref_n4", "bench_ref_n8", "bench_c5b_n2", "bench_c5b_n4", "bench_c5b_n8"):
    fp = G / f"{f}.json"
    if fp.exists():
        for ln in fp.read_text().splitlines():
            if ln.startswith("{"): lines.append(json.dumps({"file": f, **json.loads(ln)}))
(P / f"{TAG}_bench_lines.jsonl").write_text("\n".join(lines) + "\n")
for f in ("ubench.txt", "ubench2.txt"):
    if (G / f).exists(): shutil.copy(G / f, P / f"{TAG}_{f}")
print("\n".join(summary)[:6000])
