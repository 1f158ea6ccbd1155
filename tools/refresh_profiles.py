"""Copy the artefacts of tools/gpu_r02_final.sh + tools/gpu_launchlist.sh from gpurun_out/ into profiles/ and write the
summaries (launch list by kernel, ncu raw metrics, ncu source pages).  Run here, after the GPU call."""
import collections, csv, json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
G, P = "gpurun_out", "profiles"
for a, b in (("bench_r02.json", "bench_r02_final_n1.json"), ("bench_r02_ref.json", "bench_r02_reference_arm.json"),
             ("ked_traffic_r02.json", "ked_traffic_r02.json"), ("launches_r02.csv", "launches_r02.csv")):
    shutil.copy(os.path.join(G, a), os.path.join(P, b))
d = json.load(open(os.path.join(G, "bench_r02.json")))
print("value %.4g e2e %.4g ms/step %.1f launches %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"]))
print("ked", d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"].get("fp64_cublas_dgemm_tflops"), d["roofline"]["peak"])
print(d["rooflines"]["stage_ms"], d["rooflines"]["gwr_kernel"]["frac"], d["rooflines"]["ked_kernel"]["share_of_step"])
print(json.dumps(d["secondary"]))
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "ref", json.load(open(os.path.join(G, "bench_r02_ref.json")))["value"])
rows = [r for r in csv.reader(open(os.path.join(P, "launches_r02.csv"))) if len(r) > 10 and r[0].isdigit()]
nm = lambda s: re.sub(r"[<(].*", "", s.replace("void ", "").replace("twxi::", ""))
names = [nm(r[4]) for r in rows]
dur = [float(r[14].replace(",", "")) * ({"ns": 1e-3, "us": 1, "ms": 1e3}.get(r[13], 1e-3)) for r in rows]
i0 = next(i for i, n in enumerate(names) if n.startswith("unpack_chunk_kernel"))
i1 = next(i for i in range(i0, len(names)) if names[i].startswith("fixer_kernel"))
agg, cnt = {}, collections.Counter()
for n, t in zip(names[i0:i1 + 1], dur[i0:i1 + 1]):
    agg[n] = agg.get(n, 0) + t
    cnt[n] += 1
tot = sum(agg.values())
st = d["rooflines"]["stage_ms"]
live_tot, ked_live = sum(st.values()), d["rooflines"]["ked_kernel"]["ms"] / 4
out = ["# ncu launch list of `TWX_BENCH_TILES=8 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary` (profiles/launches_r02.csv,",
       "# first 500 launches; --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised).  One tile (unpack to fixer),",
       "# C5 workload (10 000 stations/var), first (full-land) tile of the list, both variables:"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    out.append("%-28s %3d launches %10.1f us  %5.1f%%" % (k, cnt[k], v, 100 * v / tot))
out.append("tile total %.1f us; ked_kernel share %.1f %% (bench.py live with CUDA events, profiles/bench_r02_final_n1.json: ked_kernel %.1f of %.1f ms = %.0f %%)"
           % (tot, 100 * agg["ked_kernel"] / tot, ked_live, live_tot, 100 * ked_live / live_tot))
open(os.path.join(P, "launches_r02_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[3:]))
KEYS = ("Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "registers_per_thread", "occupancy_limit", "warps_active.avg.pct",
        "subpipe_dmma.avg.pct", "pipe_fp64_cycles_active.avg.pct", "pipe_shared_cycles_active.avg.pct", "smsp__inst_executed.sum",
        "lsu_wavefronts_mem_shared.sum.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "issue_stalled",
        "sm__throughput.avg.pct", "smsp__issue_active.avg.pct", "sm__inst_executed.avg.per_cycle_active", "l1tex__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct", "l1tex__throughput.avg.pct")
for rep, outp in (("ked_r02.ncu-rep", "ncu_ked_r02_raw_summary.csv"), ("stages_r02.ncu-rep", "ncu_stages_r02_raw_summary.csv")):
    raw = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units, data = rr[0], rr[1], rr[2:]
    keys = [h for h in hdr if any(k in h for k in KEYS) and "per_second" not in h]
    with open(os.path.join(P, outp), "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch %d" % i for i in range(len(data))])
        for k in keys:
            i = hdr.index(k)
            w.writerow([k, units[i]] + [x[i] for x in data])
for rep, k in (("ked_r02.ncu-rep", "ked_kernel"), ("stages_r02.ncu-rep", "gwr_kernel"), ("stages_r02.ncu-rep", "knn_kernel"), ("stages_r02.ncu-rep", "nngh_params")):
    txt = subprocess.run([sys.executable, "tools/ncu_lines.py", os.path.join(G, rep), k], capture_output=True, text=True).stdout
    open(os.path.join(P, "ncu_%s_r02_source_summary.txt" % k), "w").write(txt)
