#!/usr/bin/env python
"""bench.py -- reads/s and k-mer lookups/s of the SEARCH_GG hot path on B200.

Workload (BASELINE.json configs[1], "L2S"): a synthetic L2-scale CTR (5,000
prokaryote-sized genomes, complevel 2, ~1.1 G records, ~8 GB) against 10 M
synthetic 150 bp reads per GPU with reverse complement.  A "step" is one pass
of the hot path (pack -> lookup -> vote) over the 10 M reads.

  value   : reads/s with the FASTA bytes already resident in HBM, device time
            from CUDA events on the launching stream (utb_batch_rerun_device)
  e2e     : reads/s through the C ABI with HOST buffers (utb_search_mem: H2D of
            the raw FASTA, framing + search + output formatting on the device,
            D2H of the text, copy into the caller's buffer), wall clock
  roofline: the LONGEST kernel of the resident step (every stage kernel is
            listed in roofline.kernels): algorithmic bytes per launch / its
            CUDA-event time, against MEASURED_PEAKS.json hbm_gbs for streaming
            kernels or the live random 32-byte-sector gather bandwidth
            (utb_measure_rand32) for gather kernels; reference_layout_equiv is
            SURVEY 8d's figure (32 B x the sectors the reference's own probe
            sequence touches, counted by the oracle on a sample) over the
            lookup stage's time
  cpu_baseline / --impl reference: the UNMODIFIED reference binary
            (oracle/_ref/utree-search_gg) on the host cores, same CTR + reads

Multi-GPU: one process per GPU (torchrun), reads sharded, CTR replicated, no
collective on the data path; weak scaling (10 M reads per GPU).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from oracle import oracle_api  # noqa: E402  (CPU checker: used only by the cpu_baseline / reference legs)

CONFIGS = {
    # name: universe (phyla, genera, species, strains, genome_len), complevel, ix_bytes, reads/GPU, read_len
    # one strain per species, like a RefSeq-representative set: little k-mer sharing, so the
    # 20 Gb of genome sampled 1/16 (complevel 2) keeps ~1.1 G records (~8 GB), README.md:2
    "l2s": dict(universe=(50, 20, 5, 1, 4_000_000), complevel=2, ix_bytes=2, reads=10_000_000, read_len=150,
                desc="L2-scale synthetic CTR (5000 genomes x 4 Mb, complevel 2) vs 10M x 150bp reads, RC"),
    "l4": dict(universe=(50, 20, 5, 1, 4_000_000), complevel=4, ix_bytes=2, reads=10_000_000, read_len=150,
               desc="L4 synthetic CTR (5000 genomes x 4 Mb, complevel 4) vs 10M x 150bp reads, RC"),
    # IXTYPE=uint32_t (SZ=9), > 65,536 labels, complevel 0 (dense sampling), 250 bp reads (BASELINE.json configs[4])
    "u32": dict(universe=(60, 11, 10, 10, 6000), complevel=0, ix_bytes=4, reads=10_000_000, read_len=250,
                desc="uint32-label synthetic CTR (66000 genomes x 6 kb, 73500 labels, complevel 0) vs 10M x 250bp reads, RC"),
    "small": dict(universe=(4, 3, 3, 3, 400_000), complevel=2, ix_bytes=2, reads=400_000, read_len=150,
                  desc="small synthetic CTR (108 genomes x 0.4 Mb, complevel 2) vs 400k x 150bp reads, RC"),
}
SEED = 20260101


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def work_dir():
    d = os.environ.get("UTB_BENCH_DIR")
    if not d:
        d = "/dev/shm/utb_bench" if os.path.isdir("/dev/shm") else "/tmp/utb_bench"
    os.makedirs(d, exist_ok=True)
    return d


def cfg_key(name, cfg):
    return hashlib.sha1(json.dumps([name, cfg["universe"], cfg["complevel"], cfg["ix_bytes"], SEED]).encode()).hexdigest()[:10]


def ensure_ctr(name, cfg, device):
    """Synthesises the CTR once per box (cached so both arms see the same file)."""
    from utree_b200 import build
    from tools import synthgpu
    build.build_synth()
    path = os.path.join(work_dir(), f"{name}_{cfg_key(name, cfg)}.ctr")
    meta = path + ".json"
    if os.path.exists(path) and os.path.exists(meta):
        return path, json.load(open(meta))
    uni = synthgpu.Universe(SEED, *cfg["universe"])
    t = time.time()
    tmp = path + f".tmp{os.getpid()}"
    n, nl = uni.build_ctr(tmp, complevel=cfg["complevel"], ix_bytes=cfg["ix_bytes"], device=device)
    os.replace(tmp, path)
    m = {"records": n, "labels": nl, "bytes": os.path.getsize(path), "synth_s": round(time.time() - t, 1)}
    json.dump(m, open(meta, "w"))
    log(f"synthesised {path}: {n} records, {nl} labels, {m['bytes'] / 1e9:.2f} GB in {m['synth_s']} s")
    return path, m


def make_reads(cfg, first, n_reads, device, out=None):
    from tools import synthgpu
    uni = synthgpu.Universe(SEED, *cfg["universe"])
    return uni.make_reads(n_reads, read_len=cfg["read_len"], read_seed=SEED + 1, first=first, device=device, out=out)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu_index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        sm = sorted(int(r[1]) for r in self.rows if r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if r[2].isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ---------------------------------------------------------------------------
# reference arm helpers (CPU): the unmodified reference binary
# ---------------------------------------------------------------------------
def ref_binary(cfg):
    exe = os.path.join(ROOT, "oracle", "_ref", "utree-search_gg" + ("_u32" if cfg["ix_bytes"] == 4 else ""))
    return exe if os.path.exists(exe) else None


def run_reference(exe, ctr, fasta, threads):
    out = os.path.join(work_dir(), f"ref_{os.getpid()}.out")
    t = time.time()
    p = subprocess.run([exe, ctr, fasta, out, str(threads), "RC"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    dt = time.time() - t
    if p.returncode:
        raise RuntimeError(f"reference exited {p.returncode}: {p.stderr[-300:]}")
    return dt, out


def cpu_reference_rate(cfg, ctr_path, reads_bytes, rec_bytes, n_avail, budget_s, threads, state=None):
    """reads/s of the reference binary = sample / (wall - wall of a 1-read run)."""
    exe = ref_binary(cfg)
    wd = work_dir()
    one = os.path.join(wd, f"one_{os.getpid()}.fa")
    reads_bytes[:rec_bytes].tofile(one)
    if state is None:
        state = {}
    if "t_load" not in state:
        state["t_load"] = min(run_reference(exe, ctr_path, one, threads)[0] for _ in range(2))
        # calibrate the sample on 20k reads
        cal = os.path.join(wd, f"cal_{os.getpid()}.fa")
        n_cal = min(20000, n_avail)
        reads_bytes[:n_cal * rec_bytes].tofile(cal)
        dt, _ = run_reference(exe, ctr_path, cal, threads)
        rate = n_cal / max(dt - state["t_load"], 1e-3)
        state["sample"] = int(max(20000, min(n_avail, rate * budget_s)))
        os.remove(cal)
    n = state["sample"]
    fa = os.path.join(wd, f"sample_{os.getpid()}.fa")
    if state.get("fa_n") != n:
        reads_bytes[:n * rec_bytes].tofile(fa)
        state["fa_n"] = n
    dt, out = run_reference(exe, ctr_path, fa, threads)
    search = max(dt - state["t_load"], 1e-6)
    state["last_out"] = out
    return n / search, n, search, state


def oracle_sector_stats(ctr_path, reads_bytes, rec_bytes, n_sample, threads):
    """Algorithmic bytes per lookup of SURVEY 8d, counted by the CPU checker on a sample."""
    from utree_b200 import capi
    fa = os.path.join(work_dir(), f"sect_{os.getpid()}.fa")
    reads_bytes[:n_sample * rec_bytes].tofile(fa)
    orc = oracle_api.OracleDb(ctr_path)
    t = time.time()
    rc, st, err = orc.search_file(fa, fa + ".out", do_rc=True, threads=threads)
    dt = time.time() - t
    orc.free()
    assert rc == 0, err
    bpl = 32.0 * (st["sect_idx"] + st["sect_bkt"]) / max(st["lookups"], 1)
    return {"bytes_per_lookup": bpl, "lookups_per_read": st["lookups"] / n_sample, "sample_reads": n_sample,
            "probes_per_lookup": st["probes"] / max(st["lookups"], 1), "hit_rate": st["hits"] / max(st["lookups"], 1),
            "port_reads_per_s": n_sample / dt, "out_file": fa + ".out", "fasta": fa}


_REAL_STDOUT = None


def emit_json(line):
    """The ONE JSON line goes to the process's original stdout; everything else (NCCL banners, library
    chatter) was redirected to stderr at start-up."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=os.environ.get("UTB_BENCH_CONFIG", "l2s"), choices=sorted(CONFIGS))
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU (default: the config's)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.reads:
        cfg["reads"] = args.reads
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ncpu = os.cpu_count() or 1
    rec_bytes = 12 + cfg["read_len"] + 1
    workload = {"workload": cfg["desc"], "config": args.config, "reads_per_gpu": cfg["reads"], "read_len": cfg["read_len"],
                "rc": True, "sharding": f"reads x{world}, CTR replicated", "l2_policy": "inputs larger than L2 (no flush needed)"}

    import torch
    if args.impl == "reference":
        return reference_arm(args, cfg, rank, world, local, ncpu, rec_bytes, workload)

    import torch.distributed as dist
    from utree_b200 import build, capi
    build.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- inputs ------------------------------------------------------------
    if rank == 0:
        ctr_path, ctr_meta = ensure_ctr(args.config, cfg, local)
    barrier()
    if rank != 0:
        ctr_path, ctr_meta = ensure_ctr(args.config, cfg, local)
    n_reads = cfg["reads"]
    pinned = torch.empty(n_reads * rec_bytes, dtype=torch.uint8, pin_memory=True)
    reads_np = pinned.numpy()
    t = time.time()
    make_reads(cfg, rank * n_reads, n_reads, local, out=reads_np)
    log(f"rank {rank}: {n_reads} reads ({reads_np.size / 1e9:.2f} GB FASTA) in {time.time() - t:.1f} s")

    # ---- database resident in HBM ---------------------------------------------
    t = time.time()
    ctr = capi.Ctr(ctr_path)
    db = capi.Db(ctr, local)
    db_load_s = time.time() - t
    log(f"rank {rank}: CTR resident in HBM in {db_load_s:.1f} s ({ctr.num_nodes} records, {ctr.max_ix} labels)")

    # ---- value: device-resident pass ---------------------------------------------
    batch = capi.Batch(db, reads_np.size, n_reads)
    batch.bytes[:reads_np.size] = reads_np
    r = np.arange(n_reads, dtype=np.uint64)
    batch.seq_off[:n_reads] = r * rec_bytes + 12
    batch.seq_len[:n_reads] = cfg["read_len"]
    batch.submit(reads_np.size, n_reads, True)
    res0 = batch.wait()
    lookups, hits = batch.counts()
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        batch.rerun_device(1)
    barrier()
    ms_sum = np.zeros(4)
    det_ms, det_sect = np.zeros(2), [0, 0]
    launches = 0
    for _ in range(args.steps):
        ms, l = batch.rerun_device(1)
        ms_sum += np.array(ms)
        launches += l
        dm, det_sect = batch.lookup_detail()
        det_ms += np.array(dm)
    barrier()
    dev_s = ms_sum[3] / 1e3
    lookup_mode, hbm_bytes = db.lookup_mode(), int(db.hbm_bytes())
    batch.destroy(); db.free()          # the e2e searcher below uploads its own copy: never both resident at once
    searcher = capi.Searcher(ctr, devices=(local,), host_threads=max(2, ncpu // max(world, 1)))

    # ---- e2e: host buffers through the C ABI --------------------------------------
    import ctypes
    ptr = reads_np.ctypes.data
    out_text = None
    for _ in range(max(1, args.warmup)):
        rc_, ex, out_text, st = searcher.search_mem(None, do_rc=True, ptr=ptr, n=reads_np.size)
        assert rc_ == 0, capi.lib().utb_last_error()
    barrier()
    t0 = time.time()
    e2e_stats = []
    for _ in range(args.steps):
        rc_, ex, nbytes, st = searcher.search_mem(None, do_rc=True, ptr=ptr, n=reads_np.size, copy=False)
        assert rc_ == 0 and nbytes == len(out_text)
        e2e_stats.append(st)
    barrier()
    e2e_s = time.time() - t0
    clocks = sampler.summary()

    # ---- reduce over ranks: max time, sum of units -----------------------------------
    tt = torch.tensor([dev_s, e2e_s], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([float(lookups), float(hits), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dev_s_max, e2e_s_max = tt.tolist()
    lookups_all, hits_all, launches_all = cnt.tolist()

    line = None
    if rank == 0:
        total_reads = n_reads * world * args.steps
        value = total_reads / dev_s_max
        lookups_per_s = lookups_all * args.steps / dev_s_max
        e2e_value = total_reads / e2e_s_max
        # roofline of the dominant kernel (lookup), rank 0's GPU
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        stream_peak = peaks.get("hbm_gbs", 6650.0)
        rand32 = capi.measure_rand32(local, ws_bytes=8 << 30, loads=1 << 28, iters=3)
        try:        # the same gather over an L2-resident working set: the ceiling of scattered L2 HITS (probe_kernel's regime)
            l2_scatter = capi.measure_rand32(local, ws_bytes=32 << 20, loads=1 << 28, iters=3)
        except Exception:
            l2_scatter = None
        n_sect = min(n_reads, 100_000)
        sect = oracle_sector_stats(ctr_path, reads_np, rec_bytes, n_sect, ncpu)
        # parity of this very run against the CPU checker on the sample
        want = open(sect["out_file"], "rb").read()
        assert out_text.startswith(want), "bench output differs from the oracle"
        lookup_ms = ms_sum[1] / args.steps
        two_phase = det_ms[0] > 0
        # Reference-layout figure of SURVEY 8d: bytes the reference's own probe sequence would touch.
        ref_bytes = sect["bytes_per_lookup"] * lookups
        ref_equiv = ref_bytes / (lookup_ms * 1e-3) / 1e9
        sieve_lines = ctr.num_nodes // 32 + 1024                    # 128-byte lines of the sieve (SV_RPL records per line)
        n_pos = 32.0 * ((cfg["read_len"] + 1 + 31) // 32) * n_reads  # positions of the packed stream (padded reads)
        probes = det_sect[0] / 2.0                                  # ONE sieve fetch per position serves both strands
        surv_ms = det_ms[1] / args.steps
        vote_ms = ms_sum[2] / args.steps
        # candidates: (kernel, ms per launch, algorithmic bytes per launch, peak it is held against)
        cand = []
        if two_phase:
            cand.append(("sieve_kernel<2> (minimizer-keyed blocked Bloom: one 16 B block per position serves both strands, "
                         "one DRAM line per minimizer run)", det_ms[0] / args.steps,
                         16.0 * probes + 0.375 * n_pos + 12.0 * det_sect[1], "stream",
                         "16 B sieve block per valid position + packed bases in (0.375 B/position) + one 12 B queue entry per survivor"))
            stage_bytes = 16.0 * probes + 0.375 * n_pos + 12.0 * det_sect[1] + 32.0 * det_sect[1]
        elif lookup_mode:
            cand.append(("lookup_kernel<2,true> (sector hash table, sieve off)", lookup_ms, 32.0 * det_sect[1], "rand32",
                         "32 B x the table sectors touched, counted on the device"))
            stage_bytes = 32.0 * det_sect[1]
        else:
            cand.append(("lookup_kernel<2,false> (reference probe sequence)", lookup_ms, ref_bytes, "rand32", "SURVEY 8d reference-layout bytes"))
            stage_bytes = ref_bytes
        if two_phase:
            cand.append(("queue_lookup_kernel (exact sector-hash-table lookup of the sieve survivors)", surv_ms, 32.0 * det_sect[1], "rand32",
                         "32 B x the table sectors the exact lookup touches, counted on the device"))
        cand.append(("vote_thread_kernel (+ vote_warp_kernel / vote_block_kernel for label-rich and long reads; label multiset and aufbau walk)", vote_ms,
                     det_sect[0] / 8.0 + 4.0 * hits + 32.0 * n_reads, "stream",
                     "1 bit per lookup slot of the hit map + 4 B per hit + one 32 B result per read"))
        k_name, k_ms, k_bytes, k_peak, k_alg = max(cand, key=lambda c: c[1])
        achieved = k_bytes / (k_ms * 1e-3) / 1e9
        stream_bound = k_peak == "stream"
        # ncu dram__bytes_read+write per launch of the same kernel/workload shape, when a capture is committed
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            per_lookup = tr.get(args.config, {}).get(k_name.split("<")[0].split(" ")[0] + "_dram_bytes_per_lookup")
            traffic = round(per_lookup * lookups) if per_lookup else None
        except Exception:
            pass
        peak = stream_peak if stream_bound else rand32
        roofline = {"bound": "hbm", "kernel": k_name, "achieved": round(achieved, 1), "peak": round(peak, 1),
                    "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                    "algorithmic_bytes": k_alg,
                    "peak_kind": ("MEASURED_PEAKS.json hbm_gbs (streaming copy)" if stream_bound else
                                  "measured random 32B-sector gather, 8 GiB working set (utb_measure_rand32); the L2 fills whole "
                                  "128 B lines, so this equals ~6 TB/s of DRAM reads (profiles/r01_membench_ncu.txt)"),
                    "rand32_peak": round(rand32, 1),
                    "kernels": [{"kernel": c[0].split(" ")[0], "ms": round(float(c[1]), 3), "alg_gbs": round(c[2] / (c[1] * 1e-3) / 1e9, 1) if c[1] > 0 else None,
                                 "frac": round(c[2] / (c[1] * 1e-3) / 1e9 / (stream_peak if c[3] == "stream" else rand32), 4) if c[1] > 0 else None,
                                 "peak": c[3]} for c in cand],
                    "sieve_probes_per_s": round(probes / (det_ms[0] / args.steps * 1e-3), 1) if two_phase else None,
                    "sieve_lines": int(sieve_lines),
                    "l2_scatter": ({"sectors_per_s": round(l2_scatter * 1e9 / 32.0, 1), "gbs": round(l2_scatter, 1),
                                    "how": "utb_measure_rand32 over 32 MiB"} if l2_scatter else None),
                    "phase_a_ms": {"sieve_kernel": round(float(det_ms[0] / args.steps), 3)} if two_phase else None,
                    "stream_peak": stream_peak, "stream_peak_kind": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback",
                    "kernel_ms": round(float(k_ms), 3), "kernel_share_of_step": round(float(k_ms * args.steps / ms_sum[3]), 4),
                    "lookups_per_launch": lookups,
                    "lookup_stage": {"ms": round(float(lookup_ms), 3), "bytes": stage_bytes,
                                     "gbs": round(stage_bytes / (lookup_ms * 1e-3) / 1e9, 1),
                                     "survivor_kernel_ms": round(float(surv_ms), 3),
                                     "sectors_per_lookup": round((probes + det_sect[1]) / max(lookups, 1), 3) if two_phase else None},
                    "reference_layout_equiv": {"bytes_per_lookup": round(sect["bytes_per_lookup"], 2), "gbs": round(ref_equiv, 1),
                                               "ratio_to_peak": round(ref_equiv / rand32, 3),
                                               "note": "SURVEY 8d figure: what the reference's bisection would touch per lookup; "
                                                       "above 1 because the path no longer performs those probes"},
                    "stage_ms": {"pack": round(float(ms_sum[0] / args.steps), 3), "lookup": round(float(lookup_ms), 3),
                                 "vote": round(float(vote_ms), 3)}}
        cpu = None
        if world == 1 and not args.no_cpu:
            exe = ref_binary(cfg)
            if exe:
                rate, n_s, secs, _ = cpu_reference_rate(cfg, ctr_path, reads_np, rec_bytes, n_reads, 12.0, ncpu)
                cpu = {"value": round(rate, 1), "unit": "reads/s", "cores": ncpu, "kind": "reference",
                       "sample": f"first {n_s} of the same reads, reference binary threads={ncpu}, wall minus 1-read run, {secs:.1f} s",
                       "lookups_per_s": round(rate * sect["lookups_per_read"], 1)}
                try:                                                  # and single-threaded (the parity configuration)
                    r1, n1, s1, _ = cpu_reference_rate(cfg, ctr_path, reads_np, rec_bytes, min(n_reads, 400_000), 4.0, 1)
                    cpu["value_1thread"] = round(r1, 1)
                    cpu["sample_1thread"] = f"first {n1} reads, threads=1, {s1:.1f} s"
                except Exception as e:                                # never fail the bench line over the extra figure
                    cpu["value_1thread"] = None
                    log(f"1-thread reference run failed: {e}")
            else:
                cpu = {"value": round(sect["port_reads_per_s"], 1), "unit": "reads/s", "cores": ncpu, "kind": "port",
                       "sample": f"first {n_sect} of the same reads, oracle port with {ncpu} OpenMP threads (with sector accounting)",
                       "lookups_per_s": round(sect["port_reads_per_s"] * sect["lookups_per_read"], 1)}
        st = e2e_stats[-1]
        line = {
            "metric": "reads/sec (150bp, L2-scale CTR, RC); k-mer lookups/sec in lookups_per_s",
            "value": round(value, 1), "unit": "reads/s", "lookups_per_s": round(lookups_per_s, 1),
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(dev_s_max * 1e3 / args.steps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": workload,
            "e2e": {"value": round(e2e_value, 1), "unit": "reads/s", "h2d_bytes_per_step": int(st["h2d_bytes"]),
                    "d2h_bytes_per_step": int(st["d2h_bytes"]), "ms_per_step": round(e2e_s_max * 1e3 / args.steps, 1),
                    "api": "utb_search_mem (host FASTA buffer -> host text)", "out_bytes_per_step": int(st["out_bytes"]),
                    "host_phase_s": {k: round(st[k], 3) for k in ("rd_wait_slot", "rd_fill", "rd_frame", "rd_submit",
                                                                 "fm_wait_gpu", "fm_format", "fm_emit", "seconds_device")}},
            "gpu_launches": int(launches_all), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "db": {"records": int(ctr.num_nodes), "labels": int(ctr.max_ix), "file_bytes": ctr_meta["bytes"],
                   "hbm_bytes": hbm_bytes, "load_s": round(db_load_s, 1)},
            "hit_rate": round(hits / max(lookups, 1), 4), "lookups_per_read": round(lookups / n_reads, 2),
        }
    searcher.destroy(); ctr.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line:
        emit_json(line)


def reference_arm(args, cfg, rank, world, local, ncpu, rec_bytes, workload):
    """`--impl reference`: the reference's own CPU implementation of the path, all host threads,
    bounded sample per step.  Rank 0 alone works."""
    if rank != 0:
        return
    import torch
    have_gpu = torch.cuda.is_available()
    ctr_path = os.path.join(work_dir(), f"{args.config}_{cfg_key(args.config, cfg)}.ctr")
    if not os.path.exists(ctr_path):
        if not have_gpu:
            emit_json({"impl": "reference", "unavailable": "synthetic CTR needs the GPU synthesiser and no GPU is visible"})
            return
        ctr_path, _ = ensure_ctr(args.config, cfg, local)
    n_avail = min(cfg["reads"], 2_000_000)
    reads_np = make_reads(cfg, 0, n_avail, local)
    exe = ref_binary(cfg)
    steps_total = args.steps + args.warmup
    budget = max(2.0, min(12.0, 150.0 / steps_total))
    times, state = [], None
    kind = "reference" if exe else "port"
    if exe:
        for i in range(steps_total):
            rate, n_s, secs, state = cpu_reference_rate(cfg, ctr_path, reads_np, rec_bytes, n_avail, budget, ncpu, state)
            if i >= args.warmup:
                times.append((n_s, secs))
    else:
        from utree_b200 import capi
        orc = oracle_api.OracleDb(ctr_path)
        fa = os.path.join(work_dir(), f"refport_{os.getpid()}.fa")
        n_s = min(n_avail, 200_000)
        reads_np[:n_s * rec_bytes].tofile(fa)
        for i in range(steps_total):
            t = time.time()
            orc.search_file(fa, fa + ".out", do_rc=True, threads=ncpu)
            if i >= args.warmup:
                times.append((n_s, time.time() - t))
        orc.free()
    tot_reads = sum(n for n, _ in times)
    tot_s = sum(s for _, s in times)
    value = tot_reads / tot_s
    line = {"impl": "reference", "metric": "reads/sec (150bp, L2-scale CTR, RC); k-mer lookups/sec in lookups_per_s",
            "value": round(value, 1), "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(tot_s * 1e3 / args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": workload,
            "cpu_baseline": {"value": round(value, 1), "unit": "reads/s", "cores": ncpu, "kind": kind,
                             "sample": f"{times[0][0]} reads per step of the same synthetic reads; search time = wall minus a 1-read run"},
            "e2e": {"value": round(value, 1), "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit_json(line)


if __name__ == "__main__":
    main()
