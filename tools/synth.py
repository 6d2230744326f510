"""Seeded synthetic data + file-format tools for tests and bench (NOT on the
search path).

* genomes / taxonomy map / reads in the shapes SURVEY.md App. E describes;
* ``.ubt`` reader/writer (itree.c:1317-1343) and a restatement of
  ``utree-compress`` (itree.c:1234-1315, including the first-bin quirk of
  :1282-1289) so fixtures committed as ``.ubt`` can be turned into the ``.ctr``
  the reference would have produced -- byte-identical, which
  tests/test_formats.py checks against oracle/_ref/utree-compress;
* a direct CTR writer for databases too large for the reference builder.
"""
from __future__ import annotations

import numpy as np

NUMBINS = (1 << 24) + 1
_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.full(256, ord("N"), dtype=np.uint8)
for _a, _b in zip(b"ACGTacgt", b"TGCATGCA"):
    _COMP[_a] = _b


# --------------------------------------------------------------------------
# genomes, taxonomy, reads
# --------------------------------------------------------------------------
def _mutate(rng, seq, rate):
    out = seq.copy()
    m = rng.random(seq.size) < rate
    out[m] = (out[m] + rng.integers(1, 4, int(m.sum()), dtype=np.uint8)) & 3
    return out


def make_genomes(seed=1, n_phyla=2, n_genera=2, n_species=2, n_strains=2,
                 length=50_000, rates=(0.10, 0.03, 0.005), quirky_tax=False):
    """Mutation tree phylum -> genus -> species -> strain (App. E).  Returns a
    list of dicts {name, tax, codes(uint8 0..3)}.  ``quirky_tax`` produces the
    taxonomy shapes of SURVEY 4.3 #4: empty ranks, phyla whose names diverge
    mid-token, labels that stop early."""
    rng = np.random.default_rng(seed)
    phy_names = ["Ba", "Bact", "Bacz", "Cy", "Prot", "Firm", "Act", "Spi"]
    out = []
    gid = 0
    for p in range(n_phyla):
        anc = rng.integers(0, 4, length, dtype=np.uint8)
        pn = phy_names[p % len(phy_names)] + (str(p // len(phy_names)) if p >= len(phy_names) else "")
        for g in range(n_genera):
            gen = _mutate(rng, anc, rates[0])
            for s in range(n_species):
                spe = _mutate(rng, gen, rates[1])
                for t in range(n_strains):
                    stn = _mutate(rng, spe, rates[2])
                    order = f"o__O{p}" if not (quirky_tax and g % 2) else "o__"
                    tax = (f"k__Bacteria;p__{pn};c__C{p};{order};f__F{p};"
                           f"g__G{p}x{g};s__S{p}x{g}x{s};t__T{p}x{g}x{s}x{t}")
                    if quirky_tax and (s + t) % 3 == 2:
                        tax = (f"k__Bacteria;p__{pn};c__C{p};{order};f__F{p};"
                               f"g__G{p}x{g};s__;t__")
                    if quirky_tax and gid % 7 == 5:
                        tax = f"k__Bacteria;p__{pn};c__C{p};{order};f__F{p};g__G{p}x{g}"
                    out.append({"name": f"genome{gid}", "tax": tax, "codes": stn})
                    gid += 1
    return out


def codes_to_ascii(codes):
    return _BASES[codes]


def write_fasta_and_map(genomes, fasta_path, map_path):
    with open(fasta_path, "wb") as f, open(map_path, "wb") as m:
        for g in genomes:
            f.write(b">" + g["name"].encode() + b"\n")
            f.write(codes_to_ascii(g["codes"]).tobytes() + b"\n")
            m.write(g["name"].encode() + b"\t" + g["tax"].encode() + b"\n")


def revcomp_ascii(a):
    return _COMP[a[::-1]]


def make_reads(genomes, n, seed=2, min_len=150, max_len=150, sub_rate=0.01,
               rc_frac=0.5, n_frac=0.02, random_frac=0.02, lower_frac=0.0,
               chimera_frac=0.0, name_prefix="r"):
    """Returns a list of (header_bytes_without_gt, seq_bytes).  Reads are
    sampled from the genomes, mutated, half reverse-complemented, a few get
    an N, a few are pure random (no hits expected)."""
    rng = np.random.default_rng(seed)
    recs = []
    for i in range(n):
        L = int(rng.integers(min_len, max_len + 1))
        if rng.random() < random_frac:
            a = _BASES[rng.integers(0, 4, L)]
        else:
            g = genomes[int(rng.integers(len(genomes)))]["codes"]
            L = min(L, g.size)
            st = int(rng.integers(0, g.size - L + 1))
            a = _BASES[_mutate(rng, g[st:st + L], sub_rate)]
            if chimera_frac and rng.random() < chimera_frac:
                g2 = genomes[int(rng.integers(len(genomes)))]["codes"]
                h = L // 2
                s2 = int(rng.integers(0, g2.size - h + 1))
                a = np.concatenate([a[:L - h], _BASES[g2[s2:s2 + h]]])
        a = a.copy()
        if rng.random() < rc_frac:
            a = revcomp_ascii(a).copy()
        if rng.random() < n_frac and a.size:
            k = int(rng.integers(1, 3))
            p = int(rng.integers(0, max(1, a.size - k)))
            a[p:p + k] = ord("N")
        if lower_frac and rng.random() < lower_frac:
            a = np.frombuffer(a.tobytes().lower(), dtype=np.uint8).copy()
        hdr = f"{name_prefix}{i} len={L}".encode()
        recs.append((hdr, a.tobytes()))
    return recs


def write_reads(recs, path, newline=b"\n", final_newline=True):
    with open(path, "wb") as f:
        for i, (h, s) in enumerate(recs):
            last = i == len(recs) - 1
            f.write(b">" + h + newline + s)
            if not last or final_newline:
                f.write(newline)


# --------------------------------------------------------------------------
# 2-bit words
# --------------------------------------------------------------------------
def kmer_words(codes, k=32):
    """All k-mer words of a clean uint8 code array (first base most
    significant, itree.c:924).  Returns uint64[len-k+1]."""
    n = codes.size - k + 1
    if n <= 0:
        return np.zeros(0, dtype=np.uint64)
    w = np.zeros(n, dtype=np.uint64)
    c = codes.astype(np.uint64)
    for j in range(k):
        w = (w << np.uint64(2)) | c[j:j + n]
    return w


def revcomp_words(w):
    x = ~w
    r = np.zeros_like(x)
    for _ in range(32):
        r = (r << np.uint64(2)) | (x & np.uint64(3))
        x = x >> np.uint64(2)
    return r


# --------------------------------------------------------------------------
# .ubt / .ctr files
# --------------------------------------------------------------------------
def _label_tail(labels, counts):
    return b"".join(l + b"\t" + str(int(c)).encode() + b"\n" for l, c in zip(labels, counts))


def ubt_write(path, words, ixs, labels, ix_bytes=2):
    """itree.c:1317-1343: header {8,0,ix_bytes,numNodes}, then (u64 word, ix)
    records ascending by word, then the label tail."""
    words = np.asarray(words, dtype=np.uint64)
    ixs = np.asarray(ixs)
    order = np.argsort(words, kind="stable")
    words, ixs = words[order], ixs[order]
    assert np.all(words[1:] > words[:-1]), "words must be unique"
    dt = np.dtype([("w", "<u8"), ("i", "<u2" if ix_bytes == 2 else "<u4")])
    rec = np.empty(words.size, dtype=dt)
    rec["w"], rec["i"] = words, ixs
    counts = np.bincount(ixs.astype(np.int64), minlength=len(labels))
    with open(path, "wb") as f:
        f.write(np.array([8, 0, ix_bytes, words.size], dtype="<u8").tobytes())
        f.write(rec.tobytes())
        f.write(_label_tail(labels, counts))


def ubt_read(path):
    with open(path, "rb") as f:
        md = np.frombuffer(f.read(32), dtype="<u8")
        assert md[0] == 8 and md[1] == 0 and md[2] in (2, 4)
        ix_bytes, n = int(md[2]), int(md[3])
        dt = np.dtype([("w", "<u8"), ("i", "<u2" if ix_bytes == 2 else "<u4")])
        rec = np.frombuffer(f.read(n * dt.itemsize), dtype=dt)
        tail = f.read()
    return rec["w"].copy(), rec["i"].copy(), tail, ix_bytes


def binix_like_reference(words):
    """Restates itree.c:1281-1289 exactly, quirk included: index 0 doubles as
    'unset', so a first bin holding exactly one record is lost and folded into
    the next non-empty bin."""
    n = words.size
    pre = (words >> np.uint64(40)).astype(np.int64)
    binix = np.zeros(NUMBINS, dtype=np.uint64)
    # first occurrence of each prefix, but "if(!BinIx[v]) BinIx[v]=i" never
    # records i == 0 and keeps overwriting while the stored value is 0.
    idx = np.arange(n, dtype=np.uint64)
    nz = idx > 0
    first = np.full(NUMBINS, np.iinfo(np.uint64).max, dtype=np.uint64)
    np.minimum.at(first, pre[nz], idx[nz])
    has = first != np.iinfo(np.uint64).max
    binix[has] = first[has]
    binix[NUMBINS - 1] = n
    u = int(np.flatnonzero(binix)[0])
    binix[u] = 0
    # backward fill of empty bins above u (i from NUMBINS-2 down to u+1)
    seg = binix[u + 1:NUMBINS].copy()
    if seg.size == 0:          # single-record tree: everything reads as empty
        return binix
    zero = seg == 0
    zero[-1] = False
    pos = np.where(~zero, np.arange(seg.size), seg.size - 1)
    nxt = np.minimum.accumulate(pos[::-1])[::-1]
    seg = seg[nxt]
    binix[u + 1:NUMBINS] = seg
    return binix


def ctr_write(path, words, ixs, label_tail, ix_bytes=2, binix=None):
    """Appendix A writer.  ``binix`` defaults to the reference compressor's
    (quirk and all) so the file equals what utree-compress would emit."""
    words = np.asarray(words, dtype=np.uint64)
    n = words.size
    if binix is None:
        binix = binix_like_reference(words)
    e = "<u4" if n < 0xFFFFFFFF else "<u8"
    sz = 5 + ix_bytes
    rec = np.zeros((n, sz), dtype=np.uint8)
    wb = words.astype("<u8").view(np.uint8).reshape(n, 8)
    rec[:, :5] = wb[:, :5]
    ib = np.asarray(ixs).astype("<u2" if ix_bytes == 2 else "<u4").view(np.uint8).reshape(n, ix_bytes)
    rec[:, 5:] = ib
    with open(path, "wb") as f:
        f.write(np.array([8, 0, ix_bytes, n], dtype="<u8").tobytes())
        f.write(binix.astype(e).tobytes())
        f.write(rec.tobytes())
        f.write(label_tail)


def compress(ubt_path, ctr_path):
    """utree-compress restated (itree.c:1234-1315).  The reference re-emits the
    tail through its label reader, which drops duplicate label lines; the
    fixtures here never contain duplicates, so the tail is copied."""
    words, ixs, tail, ix_bytes = ubt_read(ubt_path)
    ctr_write(ctr_path, words, ixs, tail, ix_bytes)


def labels_from_tail(tail):
    """Distinct labels in order of first appearance (id assignment of
    itree.c:1154-1223 via addSampleUdX :191-220)."""
    seen, out = set(), []
    for line in tail.split(b"\n"):
        if not line:
            continue
        lab = line.split(b"\t", 1)[0]
        if lab not in seen:
            seen.add(lab)
            out.append(lab)
    return out
