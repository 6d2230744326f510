#!/usr/bin/env python
"""Generates tests/golden/ by RUNNING THE REFERENCE (oracle/_ref, built from
/root/reference/itree.c by oracle/Makefile).  Run in the build container only;
the GPU box uses the committed files.

For every case: inputs (.ubt or hand-made .ubt, reads .fa) and the output of
the reference search binary with threads=1 are written, plus
sha256 of the reference-made .ctr so that synth.compress can be checked.
"""
import hashlib, json, os, subprocess, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
GOLD = os.path.join(ROOT, "tests", "golden")


def run(*a):
    subprocess.run(list(a), check=True, stdout=subprocess.DEVNULL)


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for b in iter(lambda: f.read(1 << 22), b""):
            h.update(b)
    return h.hexdigest()


def ref_search(ctr, fa, out, rc, u32=False):
    exe = os.path.join(REF, "utree-search_gg" + ("_u32" if u32 else ""))
    args = [exe, ctr, fa, out, "1"] + (["RC"] if rc else [])
    p = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return p.returncode, p.stdout.decode(), p.stderr.decode()


def main():
    os.makedirs(GOLD, exist_ok=True)
    meta = {}
    tmp = tempfile.mkdtemp()

    # ---- case A: u16, complevel 0, regular taxonomy -----------------------
    gA = synth.make_genomes(seed=11, n_phyla=2, n_genera=2, n_species=2, n_strains=2, length=4000)
    fa, mp = os.path.join(tmp, "A.fa"), os.path.join(tmp, "A.map")
    synth.write_fasta_and_map(gA, fa, mp)
    ubtA = os.path.join(GOLD, "toyA.ubt")
    run(os.path.join(REF, "utree-build_gg"), fa, mp, ubtA, "1", "0")
    os.remove(ubtA + ".gg.log")
    ctrA = os.path.join(tmp, "A.ctr")
    run(os.path.join(REF, "utree-compress"), ubtA, ctrA)
    meta["toyA"] = {"ctr_sha256": sha(ctrA), "ix_bytes": 2}
    reads = synth.make_reads(gA, 1200, seed=12, n_frac=0.05, random_frac=0.03, lower_frac=0.05)
    rA = os.path.join(GOLD, "toyA_reads.fa")
    synth.write_reads(reads, rA)
    for rc in (0, 1):
        out = os.path.join(GOLD, f"toyA_{'rc' if rc else 'norc'}.out")
        code, so, _ = ref_search(ctrA, rA, out, rc)
        assert code == 0
        meta[f"toyA_{'rc' if rc else 'norc'}"] = {"stdout_tail": so.splitlines()[-2:]}

    # ---- case B: u32, complevel 1, quirky taxonomy, ragged reads ----------
    gB = synth.make_genomes(seed=21, n_phyla=3, n_genera=2, n_species=2, n_strains=2, length=6000,
                            quirky_tax=True)
    fa, mp = os.path.join(tmp, "B.fa"), os.path.join(tmp, "B.map")
    synth.write_fasta_and_map(gB, fa, mp)
    ubtB = os.path.join(GOLD, "toyB_u32.ubt")
    run(os.path.join(REF, "utree-build_gg_u32"), fa, mp, ubtB, "1", "1")
    os.remove(ubtB + ".gg.log")
    ctrB = os.path.join(tmp, "B.ctr")
    run(os.path.join(REF, "utree-compress_u32"), ubtB, ctrB)
    meta["toyB_u32"] = {"ctr_sha256": sha(ctrB), "ix_bytes": 4}
    reads = synth.make_reads(gB, 400, seed=22, min_len=40, max_len=1500, n_frac=0.08,
                             random_frac=0.03, lower_frac=0.1, chimera_frac=0.3)
    rB = os.path.join(GOLD, "toyB_reads.fa")
    synth.write_reads(reads, rB)
    out = os.path.join(GOLD, "toyB_u32_rc.out")
    code, so, _ = ref_search(ctrB, rB, out, 1, u32=True)
    assert code == 0

    # ---- case Q: first-bin quirk, hand-made dense buckets ----
    words, ixs, tail, _ = synth.ubt_read(ubtA)
    labels = synth.labels_from_tail(tail)
    rng = np.random.default_rng(31)

    def word_to_ascii(w):
        return bytes(b"ACGT"[(int(w) >> (62 - 2 * i)) & 3] for i in range(32))

    def rand_suffixes(n):
        s = np.unique(rng.integers(1, (1 << 40) - 1, n * 2, dtype=np.uint64))
        return np.sort(rng.permutation(s)[:n])

    p1, p2 = 0x1B2C3D, 0x1B2C40                 # p1 holds ONE record -> folded into p2
    suf2 = rand_suffixes(9)
    mid = np.uint64((int(suf2[4]) + int(suf2[5])) // 2)
    qw = [np.uint64((p1 << 40) | int(mid))] + [np.uint64((p2 << 40) | int(x)) for x in suf2]
    p3 = 0x2FFFF0
    qw += [np.uint64((p3 << 40) | int(x)) for x in rand_suffixes(40)]
    qw = np.array(qw, dtype=np.uint64)
    qix = rng.integers(0, len(labels), qw.size)
    ubtQ = os.path.join(GOLD, "quirk.ubt")
    synth.ubt_write(ubtQ, qw, qix, labels, 2)
    ctrQ = os.path.join(tmp, "Q.ctr")
    run(os.path.join(REF, "utree-compress"), ubtQ, ctrQ)
    meta["quirk"] = {"ctr_sha256": sha(ctrQ), "ix_bytes": 2}
    qreads = [(b"alone_in_bin1", word_to_ascii(qw[0])),
              (b"foreign_suffix_in_bin2", word_to_ascii((p2 << 40) | int(mid))),
              (b"below_all_in_bin2", word_to_ascii((p2 << 40) | 0)),
              (b"above_all_in_bin2", word_to_ascii((p2 << 40) | 0xFFFFFFFFFF))]
    for i, w in enumerate(qw[1:]):
        qreads.append((f"member{i}".encode(), word_to_ascii(w)))
    for i in range(len(suf2) - 1):               # between consecutive members
        qreads.append((f"between{i}".encode(), word_to_ascii((p2 << 40) | (int(suf2[i]) + 1))))
    for i, w in enumerate(qw):                   # flanked: several valid windows per read
        qreads.append((f"flank{i}".encode(), b"ACGTTGCA" + word_to_ascii(w) + b"GGATCCAT"))
    qreads.append((b"all_members", b"N".join(word_to_ascii(w) for w in qw)))
    rQ = os.path.join(GOLD, "quirk_reads.fa")
    synth.write_reads(qreads, rQ)
    for rc in (0, 1):
        code, so, _ = ref_search(ctrQ, rQ, os.path.join(GOLD, f"quirk_{'rc' if rc else 'norc'}.out"), rc)
        assert code == 0

    # ---- case D: dense buckets of many sizes (exercises xtSuffixBS) --------
    sizes = [1, 2, 3, 4, 5, 6, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 127, 128, 129,
             255, 256, 257, 300, 511, 512, 513]
    dw, dreads = [], []
    for bi, n in enumerate(sizes):
        pre = 0x100000 + bi * 3 + (0xE00000 if bi % 2 else 0)     # both halves of the table
        suf = rand_suffixes(n)
        ws = [(pre << 40) | int(x) for x in suf]
        dw += ws
        pick = sorted(set([0, n - 1, n // 2] + list(rng.integers(0, n, 6))))
        for j in pick:
            dreads.append((f"b{n}_m{j}".encode(), word_to_ascii(ws[j])))
            dreads.append((f"b{n}_m{j}_plus1".encode(), word_to_ascii(ws[j] + 1)))
        dreads.append((f"b{n}_lo".encode(), word_to_ascii(pre << 40)))
        dreads.append((f"b{n}_hi".encode(), word_to_ascii((pre << 40) | 0xFFFFFFFFFF)))
        dreads.append((f"b{n}_all".encode(), b"N".join(word_to_ascii(w) for w in ws[:200])))
        dreads.append((f"empty_bin_after{n}".encode(), word_to_ascii((pre + 1) << 40 | int(suf[0]))))
    dw = np.array(dw, dtype=np.uint64)
    dix = rng.integers(0, len(labels), dw.size)
    ubtD = os.path.join(GOLD, "dense.ubt")
    synth.ubt_write(ubtD, dw, dix, labels, 2)
    ctrD = os.path.join(tmp, "D.ctr")
    run(os.path.join(REF, "utree-compress"), ubtD, ctrD)
    meta["dense"] = {"ctr_sha256": sha(ctrD), "ix_bytes": 2}
    rD = os.path.join(GOLD, "dense_reads.fa")
    synth.write_reads(dreads, rD)
    for rc in (0, 1):
        code, so, _ = ref_search(ctrD, rD, os.path.join(GOLD, f"dense_{'rc' if rc else 'norc'}.out"), rc)
        assert code == 0

    # ---- case L: long queries (many labels per read) on the toyA tree ------
    asc = [synth.codes_to_ascii(g["codes"]).tobytes() for g in gA]
    lreads = [(b"all_genomes", b"".join(asc)),
              (b"all_genomes_N", b"N".join(asc)),
              (b"phylum0", b"".join(asc[:8])),
              (b"rc_of_half", synth.revcomp_ascii(np.frombuffer(b"".join(asc[4:12]), dtype=np.uint8)).tobytes()),
              (b"one_genome_x3", asc[3] * 3),
              (b"short_after_long", asc[9][100:250])]
    rL = os.path.join(GOLD, "long_reads.fa")
    synth.write_reads(lreads, rL)
    code, so, _ = ref_search(ctrA, rL, os.path.join(GOLD, "long_rc.out"), 1)
    assert code == 0

    # ---- case E: reader edge cases (App. B.1) on the toyA tree -------------
    g0 = synth.codes_to_ascii(gA[0]["codes"]).tobytes()
    g5 = synth.codes_to_ascii(gA[5]["codes"]).tobytes()
    edge = (b">crlf one\r\n" + g0[100:260] + b"\r\n"
            b">tab\tin header\n" + g0[300:460] + b"\n"
            b">nospace\n" + g5[10:200] + b"\n"
            b">short\n" + g0[0:31] + b"\n"
            b">exact32\n" + g0[0:32] + b"\n"
            b">emptyline\n\n"
            b">lower\n" + g5[500:700].lower() + b"\n"
            b">manyN\n" + g0[600:640] + b"NNN" + g0[643:700] + b"n" + g0[701:800] + b"\n"
            b">other chars\n" + g0[900:950] + b"RYKM-" + g0[955:1100] + b"\n"
            b">cr_in_name\rX\n" + g5[1000:1150] + b"\n"
            b">nofinalnewline\n" + g5[2000:2200])
    rE = os.path.join(GOLD, "edge_reads.fa")
    open(rE, "wb").write(edge)
    code, so, _ = ref_search(ctrA, rE, os.path.join(GOLD, "edge_rc.out"), 1)
    assert code == 0
    # malformed inputs: expected exit codes (App. C)
    bad = {"bad_noheader.fa": b"ACGT\nACGT\n",
           "bad_seq_is_header.fa": b">a\n" + g0[:100] + b"\n>b\n>c\n",
           "bad_truncated.fa": b">a\n" + g0[:100] + b"\n>b\n"}
    for name, data in bad.items():
        pth = os.path.join(GOLD, name)
        open(pth, "wb").write(data)
        code, so, se = ref_search(ctrA, pth, os.path.join(GOLD, name + ".out"), 1)
        meta[name] = {"exit": code}

    json.dump(meta, open(os.path.join(GOLD, "meta.json"), "w"), indent=1, sort_keys=True)
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
