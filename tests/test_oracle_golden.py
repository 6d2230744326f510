"""The oracle (oracle/oracle.c) against the golden outputs produced by the
reference binary, and its unit-level pieces.  CPU only."""
import json
import os

import numpy as np
import pytest

import conftest  # noqa: F401  (puts the repo root on sys.path)
from oracle import oracle_api
from conftest import BAD_CASES, CASES, GOLD, gold, read_fasta


@pytest.fixture(scope="module")
def orcs(built, ctrs):
    from utree_b200 import capi
    d = {k: oracle_api.OracleDb(v) for k, v in ctrs.items()}
    yield d
    for o in d.values():
        o.free()


@pytest.mark.parametrize("db_name,reads,out,rc", CASES)
@pytest.mark.parametrize("threads", [1, 3])
def test_oracle_reproduces_reference_output(orcs, tmp_path, db_name, reads, out, rc, threads):
    o = str(tmp_path / "o.txt")
    code, st, err = orcs[db_name].search_file(gold(reads), o, do_rc=bool(rc), threads=threads)
    assert code == 0, err
    assert open(o, "rb").read() == open(gold(out), "rb").read()
    assert st["good_finds"] == sum(1 for _ in open(gold(out), "rb"))


@pytest.mark.parametrize("bad", BAD_CASES)
def test_oracle_malformed_inputs(orcs, tmp_path, meta, bad):
    o = str(tmp_path / "o.txt")
    code, st, err = orcs["toyA"].search_file(gold(bad), o, do_rc=True)
    assert code == meta[bad]["exit"] == 2
    assert open(o, "rb").read() == open(gold(bad + ".out"), "rb").read()


def test_good_finds_match_reference_stdout(orcs, tmp_path, meta):
    for key, rc in (("toyA_rc", 1), ("toyA_norc", 0)):
        _, st, _ = orcs["toyA"].search_file(gold("toyA_reads.fa"), str(tmp_path / "o"), do_rc=bool(rc))
        assert meta[key]["stdout_tail"] == [f"Good finds: {st['good_finds']}", f"Searched {st['reads']} queries"]


def test_revcomp_word_is_rc_of_text(built):
    """rc(kmer) equals the k-mer of the reverse-complemented text."""
    from utree_b200 import capi
    from tools import synth
    rng = np.random.default_rng(3)
    codes = rng.integers(0, 4, 200, dtype=np.uint8)
    w = synth.kmer_words(codes)
    rc_text = (3 - codes)[::-1]
    w_rc = synth.kmer_words(rc_text)[::-1]
    O = oracle_api.oracle()
    assert [O.orc_revcomp_word(int(x)) for x in w] == [int(x) for x in w_rc]
    assert np.array_equal(synth.revcomp_words(w), w_rc)


def test_slide_skips_ambiguous_windows(orcs):
    """Every all-ACGT window exactly once, none containing another byte (App. B.2)."""
    from tools import synth
    orc = orcs["toyA"]
    seq = read_fasta(gold("edge_reads.fa"))[7][1]          # manyN
    assert b"N" in seq and b"n" in seq
    _, words = orc.slide(seq, do_rc=False, want_words=True)
    exp = []
    for i in range(len(seq) - 31):
        win = seq[i:i + 32].upper()
        if all(c in b"ACGT" for c in win):
            exp.append(int(synth.kmer_words(np.array([b"ACGT".index(c) for c in win], dtype=np.uint8))[0]))
    assert [int(x) for x in words] == exp


def test_lookup_on_dense_buckets(orcs):
    """xtSuffixBS on buckets of size 1..513: members hit with their id, neighbours miss."""
    from tools import synth
    words, ixs, _, _ = synth.ubt_read(gold("dense.ubt"))
    orc = orcs["dense"]
    got = orc.lookup_many(words)
    # the first bin holds one record, so the compressor's quirk loses it
    assert got[0] == 0xFFFFFFFF
    assert np.array_equal(got[1:], ixs[1:].astype(np.uint32))
    member = set(int(w) for w in words)
    near = [int(w) + 1 for w in words[::7] if int(w) + 1 not in member]
    assert all(orc.lookup(w) == 0xFFFFFFFF for w in near)


def test_quirk_bucket_semantics(orcs):
    """the lone record of the first bin is lost; the folded bucket
    is searched with the reference's probe sequence."""
    from tools import synth
    words, ixs, _, _ = synth.ubt_read(gold("quirk.ubt"))
    orc = orcs["quirk"]
    got = orc.lookup_many(words)
    assert got[0] == 0xFFFFFFFF                  # alone in its bin -> unreachable
    assert np.array_equal(got[1:], ixs[1:].astype(np.uint32))
    p2 = int(words[1]) >> 40
    foreign = (p2 << 40) | (int(words[0]) & 0xFFFFFFFFFF)
    assert orc.lookup(foreign) == 0xFFFFFFFF      # its suffix is not below the bucket's minimum


def test_vote_branches(orcs):
    """Hand-made multisets covering the branches of App. B.4."""
    orc = orcs["toyA"]
    labs = {orc.label(i): i for i in range(orc.max_ix)}
    strain = [i for l, i in labs.items() if b";t__" in l]
    v = orc.vote(np.array([], dtype=np.uint32)); assert v.kind == 0
    v = orc.vote(np.array([strain[0]], dtype=np.uint32)); assert (v.kind, v.found, v.uix) == (1, 1, 1)
    v = orc.vote(np.array([strain[0]] * 9, dtype=np.uint32)); assert (v.kind, v.found, v.uix) == (1, 9, 1)
    # 3/4 majority for one strain -> full label
    v = orc.vote(np.array([strain[0]] * 9 + [strain[1]] * 1, dtype=np.uint32))
    assert v.kind == 2 and v.label == strain[0] and v.cut == 0xFFFFFFFE
    # even split between two phyla -> stops at the shared prefix
    a = [i for l, i in labs.items() if b"p__Ba;" in l and b";t__" in l][0]
    b = [i for l, i in labs.items() if b"p__Bact;" in l and b";t__" in l][0]
    v = orc.vote(np.array([a] * 5 + [b] * 5, dtype=np.uint32))
    assert v.kind == 2 and v.uix == 2 and (v.sl, v.ol) == (5, 10)      # the failing level is reported (App. D #8)
    assert orc.label(v.label)[:v.cut] == b"k__Bacteria"


# ---- the non-GG binary (-D SEARCH): SPARSITY-skip slide + shallow top-2 vote ------------------------
SHALLOW = json.load(open(os.path.join(GOLD, "meta_shallow.json")))


@pytest.mark.parametrize("name", sorted(SHALLOW))
def test_oracle_shallow_search_reproduces_the_reference(ctrs, tmp_path, name):
    """oracle/_ref/utree-search (itree.c with -D SEARCH) produced these files (scripts/make_golden_shallow.py); the
    restatement must give the same bytes and the same "Good finds" / "Searched" counts.  The vote of a read depends on
    the reads before it (itree.c:982 reads one entry past the hit list), so this also pins the order dependence."""
    m = SHALLOW[name]
    orc = oracle_api.OracleDb(ctrs[m["db"]])
    out = str(tmp_path / "o.out")
    rc, st, err = orc.search_file_shallow(os.path.join(GOLD, m["reads"]), out, do_rc=bool(m["rc"]))
    orc.free()
    assert rc == m["exit"], err
    assert open(out, "rb").read() == open(os.path.join(GOLD, name), "rb").read()
    assert [f"Good finds: {st['good_finds']}", f"Searched {st['reads']} queries"] == m["stdout_tail"]
