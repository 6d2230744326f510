"""GPU box: device-resident stage times with the match-ranked vs the atomic-ranked partitioner."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from utree_b200 import capi
cfg = dict(bench.CONFIGS["l2s"]); n = int(os.environ.get("AB_READS", "10000000"))
ctr_path, _ = bench.ensure_ctr("l2s", cfg, 0)
reads = bench.make_reads(cfg, 0, n, 0)
rec = 12 + cfg["read_len"] + 1
ctr = capi.Ctr(ctr_path); db = capi.Db(ctr, 0)
b = capi.Batch(db, reads.size, n)
b.bytes[:reads.size] = reads
b.seq_off[:n] = np.arange(n, dtype=np.uint64) * rec + 12; b.seq_len[:n] = cfg["read_len"]
b.submit(reads.size, n, True); b.wait()
for atoms in ("0", "1", "0", "1"):
    os.environ["UTB_PART_ATOMS"] = atoms
    b.rerun_device(2)
    ms, _ = b.rerun_device(3)
    pm = b.partition_detail(); dm, _ = b.lookup_detail()
    print(f"atoms={atoms}: partition {pm[0]:.2f} ms probe {pm[1]:.2f} ms survivors {dm[1]:.2f} ms | pack {ms[0]/3:.2f} lookup {ms[1]/3:.2f} vote {ms[2]/3:.2f} total {ms[3]/3:.2f}")
