"""Copies one final_round.sh output set (gpurun_out/<tag>_*) into profiles/r02_*: bench lines, launch list + summary, the raw
page of the --set full capture + a summary, traffic.json.  Usage: python scripts/collect_profiles.py <tag>"""
import collections, csv, json, os, shutil, subprocess, sys
tag = sys.argv[1]
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(R, "gpurun_out"), os.path.join(R, "profiles")
for c in ("l2s", "l4", "long", "u32", "reference_arm"):
    src = os.path.join(G, f"{tag}_bench_{c}.json")
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, f"r02_bench_{c}.json"))
for a, b in ((f"{tag}_launches.csv", "r02_launches_l2s.csv"), (f"{tag}_gputests.log", "r02_gputests.log")):
    if os.path.exists(os.path.join(G, a)):
        shutil.copy(os.path.join(G, a), os.path.join(P, b))
rep = os.path.join(G, f"{tag}_prof_l2s.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(os.path.join(P, "r02_ncu_full_l2s_10Mreads_raw.csv"), "w").write(raw)
    rows = list(csv.reader(raw.splitlines()))
    h = rows[0]
    keep = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
            "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_op_read.sum", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__cycles_active.avg", "sm__cycles_elapsed.avg"]
    keep += [c for c in h if c.startswith("smsp__average_warps_issue_stalled") and c.endswith("per_issue_active.ratio")]
    keep = [k for k in keep if k in h]
    csv.writer(open(os.path.join(P, "r02_ncu_full_l2s_10Mreads_summary.csv"), "w")).writerows(
        [keep, [rows[1][h.index(k)] for k in keep]] + [[r[h.index(k)] for k in keep] for r in rows[2:]])
    tr = {"_doc": "ncu dram__bytes_read.sum + dram__bytes_write.sum PER LAUNCH of each hot kernel at the bench shape (one resident batch of "
                  "10,000,000 x 150 bp reads, RC), one --set full capture: profiles/r02_ncu_full_l2s_10Mreads_raw.csv (scripts/final_round.sh). "
                  "bench.py copies the figure of its roofline kernel into roofline.traffic when the step has exactly these reads.",
          "l2s": {"reads": 10000000}}
    for r in rows[2:]:
        key = r[h.index("Kernel Name")].replace("void ", "").split("<")[0].split("(")[0]
        b = (float(r[h.index("dram__bytes_read.sum")]) + float(r[h.index("dram__bytes_write.sum")])) * 1e9
        tr["l2s"][key] = round(b)
        print(key, round(b / 1e9, 2), "GB", r[h.index("gpu__time_duration.sum")], "ms", r[h.index("smsp__inst_executed.sum")], "warp instr")
    tr["l2s"]["vote"] = tr["l2s"].get("vote_thread_kernel")
    json.dump(tr, open(os.path.join(P, "traffic.json"), "w"), indent=1)
lc = os.path.join(P, "r02_launches_l2s.csv")
if os.path.exists(lc):
    rows = [r for r in csv.reader(l for l in open(lc) if not l.startswith("=="))]
    h = rows[0]; ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        n = r[ki].replace("void ", "").split("(")[0]; t = float(r[vi].replace(",", ""))
        t = t / 1e6 if r[ui] == "ns" else t / 1e3 if r[ui] == "us" else t
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += t
    tot = sum(v[1] for v in agg.values())
    step = {k: v for k, v in agg.items() if k.split("<")[0] in ("pack_kernel", "sieve_kernel", "queue_lookup_kernel", "vote_thread_kernel", "vote_warp_kernel",
                                                                "vote_block_kernel", "vote_big_count_kernel", "vote_big_finish_kernel")}
    st = sum(v[1] for v in step.values())
    d = json.load(open(os.path.join(P, "r02_bench_l2s.json")))
    with open(os.path.join(P, "r02_launches_l2s_summary.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 600, `python bench.py --steps 2 --warmup 1 --no-cpu --no-extra` (L2S, 10 M reads):\n"
                "the first 600 launches of the process -- the table / sieve build of the upload (ktab_build, sieve_build, verify), the input synthesiser\n"
                "(reads_kernel, tools/), then the resident passes of one 10 M-read batch and the e2e passes in 128 MiB batches.  Per-launch times are\n"
                "cold-cache and serialised.\n\n")
        f.write(f"{'kernel':60s} {'launches':>8s} {'total ms':>10s} {'share':>7s}\n")
        for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{n:60s} {c:8d} {t:10.3f} {100 * t / tot:6.1f}%\n")
        f.write(f"{'total':60s} {sum(v[0] for v in agg.values()):8d} {tot:10.3f}\n\n")
        sv = [v[1] for k, v in step.items() if k.startswith("sieve_kernel")][0]
        f.write(f"Among the kernels of the search step (pack + sieve + queue_lookup + vote_*): sieve_kernel {sv:.3f} / {st:.3f} ms = {100 * sv / st:.1f} %.\n"
                f"bench.py's live CUDA-event figure for the same kernel (profiles/r02_bench_l2s.json): {d['roofline']['kernel_ms']} / {d['ms_per_step']} ms = "
                f"{100 * d['roofline']['kernel_share_of_step']:.1f} %.\n")
    print(open(os.path.join(P, "r02_launches_l2s_summary.txt")).read()[-400:])
