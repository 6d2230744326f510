"""ctypes wrapper of libutb_synth.so (tools/synth.cu): GPU-side
generator of bench / large-test INPUTS (a CTR file and FASTA reads of a seeded
synthetic universe).  Not on the search path."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libutb_synth.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.uts_last_error.restype = C.c_char_p
        L.uts_build_ctr.argtypes = [C.c_int, C.c_uint64] + [C.c_uint32] * 7 + [C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
        L.uts_reads_bytes.restype = C.c_uint64
        L.uts_reads_bytes.argtypes = [C.c_uint64, C.c_uint32]
        L.uts_make_reads.argtypes = [C.c_int, C.c_uint64] + [C.c_uint32] * 5 + [C.c_uint64, C.c_uint64, C.c_uint64] + [C.c_uint32] * 4 + [C.c_void_p]
        L.uts_make_long_reads.argtypes = [C.c_int, C.c_uint64] + [C.c_uint32] * 5 + [C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
        L.uts_genome_ascii.argtypes = [C.c_uint64] + [C.c_uint32] * 6 + [C.c_void_p]
        L.uts_genome_ascii.restype = None
        L.uts_genome_tax.argtypes = [C.c_uint64] + [C.c_uint32] * 6 + [C.c_char_p, C.c_size_t]
        L.uts_genome_tax.restype = None
        _lib = L
    return _lib


M64 = (1 << 64) - 1


def _mix64(x):
    """tools/synth.cu mix64 on numpy uint64 arrays."""
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def build_big_ctr(path, n_records, n_labels=1000, device=0):
    """A CTR with closed-form content (word(i) = i * S + mix64(i) % S, label(i) = mix64(i ^ 0x5555) % n_labels): for trees of
    >= 2^32 - 1 records (8-byte prefix index) that no second copy in RAM has to vouch for."""
    lib().uts_build_big_ctr.argtypes = [C.c_int, C.c_uint64, C.c_uint32, C.c_char_p]
    rc = lib().uts_build_big_ctr(device, n_records, n_labels, os.fsencode(path))
    if rc:
        raise RuntimeError("uts_build_big_ctr: " + lib().uts_last_error().decode())


def big_ctr_word_label(i, n_records, n_labels=1000):
    """word(i), label(i) of build_big_ctr for an array of record indices."""
    i = np.ascontiguousarray(i, dtype=np.uint64)
    S = np.uint64(M64 // n_records)
    with np.errstate(over="ignore"):
        return i * S + _mix64(i) % S, (_mix64(i ^ np.uint64(0x5555)) % np.uint64(n_labels)).astype(np.uint32)


class Universe:
    """n_phyla x n_genera x n_species x n_strains genomes of genome_len bases."""

    def __init__(self, seed, n_phyla, n_genera, n_species, n_strains, genome_len):
        self.seed, self.shape, self.genome_len = seed, (n_phyla, n_genera, n_species, n_strains), genome_len
        self.n_genomes = n_phyla * n_genera * n_species * n_strains

    def _args(self):
        return [self.seed, *self.shape, self.genome_len]

    def build_ctr(self, path, complevel=2, ix_bytes=2, device=0):
        n, nl = C.c_uint64(), C.c_uint32()
        rc = lib().uts_build_ctr(device, *self._args(), complevel, ix_bytes, os.fsencode(path), C.byref(n), C.byref(nl))
        if rc:
            raise RuntimeError("uts_build_ctr: " + lib().uts_last_error().decode())
        return n.value, nl.value

    def record_bytes(self, read_len):
        return 12 + read_len + 1

    def make_reads(self, n_reads, read_len=150, read_seed=7, first=0, sub_permille=10, n_permille=20,
                   random_permille=20, device=0, out=None):
        """Returns a uint8 array holding fixed-width FASTA records
        '>r%09u\\n' + bases + '\\n' for reads [first, first + n_reads)."""
        nb = lib().uts_reads_bytes(n_reads, read_len)
        if out is None:
            out = np.empty(nb, dtype=np.uint8)
        assert out.size >= nb
        rc = lib().uts_make_reads(device, *self._args(), read_seed, first, n_reads, read_len, sub_permille,
                                  n_permille, random_permille, out.ctypes.data)
        if rc:
            raise RuntimeError("uts_make_reads: " + lib().uts_last_error().decode())
        return out[:nb]

    def make_long_reads(self, lengths, read_seed=7, first=0, sub_permille=10, device=0, out=None):
        """Variable-length records '>r%09u\\n' + bases + '\\n' (long reads, whole-genome queries that run on through
        the following genomes).  Returns (uint8 array, byte offsets of the records)."""
        lengths = np.ascontiguousarray(lengths, dtype=np.uint32)
        off = np.zeros(lengths.size + 1, dtype=np.uint64)
        off[1:] = np.cumsum(lengths.astype(np.uint64) + 13)
        nb = int(off[-1])
        if out is None:
            out = np.empty(nb, dtype=np.uint8)
        assert out.size >= nb
        rc = lib().uts_make_long_reads(device, *self._args(), read_seed, first, lengths.size, lengths.ctypes.data, sub_permille,
                                       out.ctypes.data)
        if rc:
            raise RuntimeError("uts_make_long_reads: " + lib().uts_last_error().decode())
        return out[:nb], off

    def genome_ascii(self, g):
        buf = np.empty(self.genome_len, dtype=np.uint8)
        lib().uts_genome_ascii(*self._args(), g, buf.ctypes.data)
        return buf.tobytes()

    def genome_tax(self, g):
        buf = C.create_string_buffer(512)
        lib().uts_genome_tax(*self._args(), g, buf, 512)
        return buf.value
