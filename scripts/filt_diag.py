"""GPU box: times the filter kernel variants on a resident 2M-read batch (diagnostic only)."""
import os, sys, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from utree_b200 import capi
cfg = dict(bench.CONFIGS["l2s"]); n = 2_000_000
ctr_path, _ = bench.ensure_ctr("l2s", cfg, 0)
reads = bench.make_reads(cfg, 0, n, 0)
rec = 12 + cfg["read_len"] + 1
ctr = capi.Ctr(ctr_path); db = capi.Db(ctr, 0)
b = capi.Batch(db, reads.size, n)
b.bytes[:reads.size] = reads
b.seq_off[:n] = np.arange(n, dtype=np.uint64) * rec + 12; b.seq_len[:n] = cfg["read_len"]
b.submit(reads.size, n, True); b.wait()
lk, _ = b.counts()
for diag in ("0", "1", "2", "0"):
    os.environ["UTB_FILT_DIAG"] = diag
    b.rerun_device(2)
    ms, _ = b.rerun_device(5)
    dm, sec = b.lookup_detail()
    print(f"diag={diag}: filter {dm[0]:.2f} ms  survivors {dm[1]:.2f} ms  lookup stage {ms[1]/5:.2f} ms  probes {sec[0]}  -> {sec[0]/dm[0]/1e6:.1f} G probes/s")
