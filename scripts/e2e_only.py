"""GPU box: utb_search_mem only (no resident pass), N repetitions; prints wall time and host phase timers."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from utree_b200 import capi
cfg = dict(bench.CONFIGS["l2s"]); n = int(os.environ.get("E2E_READS", "10000000")); reps = int(os.environ.get("E2E_REPS", "3"))
ctr_path, _ = bench.ensure_ctr("l2s", cfg, 0)
reads, _ = bench.make_reads(cfg, 0, n, 0, pin=True)
ctr = capi.Ctr(ctr_path)
s = capi.Searcher(ctr, devices=(0,), host_threads=int(os.environ.get("E2E_THREADS", os.cpu_count())))
for i in range(reps):
    t = time.time()
    rc, ex, nbytes, st = s.search_mem(None, do_rc=True, ptr=reads.ctypes.data, n=reads.size, copy=False)
    dt = time.time() - t
    print(f"rep {i}: {dt*1e3:.1f} ms  {n/dt/1e6:.1f} M reads/s  batches {st['batches']} launches {st['kernel_launches']} "
          + " ".join(f"{k}={st[k]:.3f}" for k in ("rd_wait_slot", "rd_fill", "rd_frame", "rd_submit", "fm_wait_gpu", "fm_emit", "seconds_device")), flush=True)
