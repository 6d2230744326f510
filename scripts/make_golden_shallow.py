#!/usr/bin/env python
"""Generates the tests/golden/*_shallow_* fixtures by RUNNING THE REFERENCE's non-GG search binary
(oracle/_ref/utree-search[_u32], built from /root/reference/itree.c with -D SEARCH by oracle/Makefile)
on the committed trees and reads.  Run in the build container only; the GPU box uses the committed files.
The binary is single-threaded in this mode and its vote depends on the ORDER of the reads (itree.c:982),
so the outputs are exact as they are.

shallow_reads.fa is added to the committed inputs: exact copies of genome windows of very different lengths
in an order that makes the cross-read dependency bite (long matches first, then shorter ones)."""
import json, os, subprocess, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
GOLD = os.path.join(ROOT, "tests", "golden")

CASES = [("toyA", "toyA_reads.fa", 1), ("toyA", "toyA_reads.fa", 0), ("toyB_u32", "toyB_reads.fa", 1),
         ("dense", "dense_reads.fa", 1), ("dense", "dense_reads.fa", 0), ("quirk", "quirk_reads.fa", 1),
         ("toyA", "long_reads.fa", 1), ("toyA", "edge_reads.fa", 1), ("toyA", "shallow_reads.fa", 1), ("toyA", "shallow_reads.fa", 0)]


def out_name(db, reads, rc):
    return f"shallow_{db}_{reads[:-len('_reads.fa')]}_{'rc' if rc else 'norc'}.out"


def main():
    tmp = tempfile.mkdtemp()
    # extra reads: windows of the toyA genomes (complevel 0: every 32-mer is in the tree), lengths falling then rising
    gA = synth.make_genomes(seed=11, n_phyla=2, n_genera=2, n_species=2, n_strains=2, length=4000)
    asc = [synth.codes_to_ascii(g["codes"]).tobytes() for g in gA]
    rng = np.random.default_rng(41)
    recs = []
    for k, L in enumerate([3000, 1500, 900, 400, 250, 150, 120, 90, 64, 48, 40, 33, 32, 40, 64, 150, 400, 1500] * 6):
        g = int(rng.integers(len(asc)))
        st = int(rng.integers(0, len(asc[g]) - L))
        seq = bytearray(asc[g][st:st + L])
        if k % 5 == 0 and L > 64:
            seq[L // 2] = ord("N")
        if k % 7 == 3 and L > 200:                  # chimera of two genomes: two competing labels
            g2 = (g + 5) % len(asc)
            seq[L // 2:] = asc[g2][st + L // 2:st + L]
        recs.append((f"s{k}_L{L}_g{g}".encode(), bytes(seq)))
    synth.write_reads(recs, os.path.join(GOLD, "shallow_reads.fa"))
    meta = {}
    for db, reads, rc in CASES:
        u32 = db.endswith("u32")
        ctr = os.path.join(tmp, db + ".ctr")
        if not os.path.exists(ctr):
            subprocess.run([os.path.join(REF, "utree-compress" + ("_u32" if u32 else "")), os.path.join(GOLD, db + ".ubt"), ctr],
                           check=True, stdout=subprocess.DEVNULL)
        out = os.path.join(GOLD, out_name(db, reads, rc))
        args = [os.path.join(REF, "utree-search" + ("_u32" if u32 else "")), ctr, os.path.join(GOLD, reads), out, "1"] + (["RC"] if rc else [])
        p = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        meta[os.path.basename(out)] = {"db": db, "reads": reads, "rc": rc, "exit": p.returncode,
                                       "stdout_tail": p.stdout.decode().splitlines()[-2:]}
        print(os.path.basename(out), os.path.getsize(out), p.returncode, meta[os.path.basename(out)]["stdout_tail"])
    json.dump(meta, open(os.path.join(GOLD, "meta_shallow.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
