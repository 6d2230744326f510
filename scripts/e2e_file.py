"""GPU box: utb_search_file only (files on /dev/shm), N repetitions; prints wall time and host phase timers.
UTB_TIMELINE=1 adds the per-batch host timestamps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from utree_b200 import capi
name = os.environ.get("E2E_CONFIG", "l2s")
cfg = dict(bench.CONFIGS[name]); n = int(os.environ.get("E2E_READS", str(cfg["reads"]))); reps = int(os.environ.get("E2E_REPS", "3"))
ctr_path, _ = bench.ensure_ctr(name, cfg, 0)
reads, _ = bench.make_reads(cfg, 0, n, 0)
wd = bench.work_dir()
fa, out = os.path.join(wd, "e2e_file.fa"), os.path.join(wd, "e2e_file.out")
reads.tofile(fa)
ctr = capi.Ctr(ctr_path)
s = capi.Searcher(ctr, devices=(0,), host_threads=int(os.environ.get("E2E_THREADS", os.cpu_count())))
for i in range(reps):
    t = time.time()
    rc, ex, st = s.search_file(fa, out, do_rc=True)
    dt = time.time() - t
    print(f"rep {i}: {dt*1e3:.1f} ms  {n/dt/1e6:.1f} M reads/s  in {reads.size/1e6:.0f} MB out {os.path.getsize(out)/1e6:.0f} MB batches {st['batches']} "
          + " ".join(f"{k}={st[k]:.3f}" for k in ("rd_wait_slot", "rd_fill", "rd_frame", "rd_submit", "fm_wait_gpu", "fm_emit", "seconds_device")), flush=True)
os.remove(fa); os.remove(out)
