"""N>1 path on CPU: two gloo ranks shard the reads, run the host stages of the
product (framer + formatter) around the oracle's lookup/vote, and rank 0's
ordered merge must equal the reference output.  The device stages are covered
on the GPU box by the virtual-device test."""
import os
import subprocess
import sys
import textwrap

import pytest

from conftest import GOLD, ROOT


def test_shard_ranges_cover_everything():
    from utree_b200.shard import shard_range
    for n in (0, 1, 7, 100, 1201):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np
    import torch.distributed as dist
    from utree_b200 import capi, shard
    from oracle import oracle_api
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
    rank, world = dist.get_rank(), dist.get_world_size()
    data = open({fasta!r}, "rb").read()
    lo, hi = shard.split_fasta(data, rank, world)
    mine = data[lo:hi]
    ctr, orc = capi.Ctr({ctr!r}), oracle_api.OracleDb({ctr!r})
    rc, ex, used, recs, (name_off, name_len) = capi.frame_records(mine, threads=2)
    assert rc == 0 and used == len(mine)
    res = np.zeros(len(recs), dtype=capi.RESULT_DTYPE)
    for i, (_, seq) in enumerate(recs):
        hits, _ = orc.slide(seq, do_rc=True)
        v = orc.vote(hits)
        res[i] = (v.kind, v.label, v.cut, v.found, v.uix, v.sl, v.ol, 0)
    text = capi.format_results(ctr, mine, name_off, name_len, res)
    parts = [None, None]
    dist.all_gather_object(parts, (len(recs), text))
    dist.barrier()
    if rank == 0:
        merged = shard.merge_outputs([p[1] for p in parts])
        assert sum(p[0] for p in parts) == 1200
        open({out!r}, "wb").write(merged)
    dist.destroy_process_group()
""")


def test_two_gloo_ranks_merge_to_reference_output(built, ctrs, tmp_path):
    out = str(tmp_path / "merged.out")
    script = str(tmp_path / "worker.py")
    port = 29500 + os.getpid() % 2000
    open(script, "w").write(WORKER.format(root=ROOT, port=port, fasta=os.path.join(GOLD, "toyA_reads.fa"),
                                          ctr=ctrs["toyA"], out=out))
    procs = [subprocess.Popen([sys.executable, script, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    for p in procs:
        o, _ = p.communicate(timeout=300)
        assert p.returncode == 0, o.decode()[-2000:]
    assert open(out, "rb").read() == open(os.path.join(GOLD, "toyA_rc.out"), "rb").read()
