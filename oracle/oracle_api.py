"""ctypes mirror of oracle/oracle.h -- the CPU checker (TEST INFRASTRUCTURE).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this; the
product package utree_b200/ never does."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_PATH = os.path.join(HERE, "liboracle.so")


class OrcStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("reads", "good_finds", "lookups", "hits", "probes",
                                          "sect_idx", "sect_bkt", "out_bytes")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class OrcVote(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("kind", "label", "cut", "found", "uix", "sl", "ol")]


_orc = None


def oracle():
    global _orc
    if _orc is None:
        O = C.CDLL(ORACLE_PATH)
        O.orc_db_load.restype = C.c_void_p
        O.orc_db_load.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
        O.orc_db_free.argtypes = [C.c_void_p]
        O.orc_db_num_nodes.restype = C.c_uint64
        for f in ("orc_db_num_nodes", "orc_db_max_ix", "orc_db_ix_bytes", "orc_db_binix_bytes"):
            getattr(O, f).argtypes = [C.c_void_p]
        O.orc_db_label.restype = C.c_char_p
        O.orc_db_label.argtypes = [C.c_void_p, C.c_uint32]
        O.orc_lookup.restype = C.c_uint32
        O.orc_lookup.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(OrcStats)]
        O.orc_slide.restype = C.c_uint64
        O.orc_slide.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_int, C.c_void_p, C.c_uint64,
                                C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(OrcStats)]
        O.orc_vote.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(OrcVote)]
        O.orc_search_file.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_uint64,
                                      C.POINTER(OrcStats), C.c_char_p, C.c_size_t]
        O.orc_search_file_shallow.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(OrcStats), C.c_char_p, C.c_size_t]
        O.orc_revcomp_word.restype = C.c_uint64
        O.orc_revcomp_word.argtypes = [C.c_uint64]
        _orc = O
    return _orc


class OracleDb:
    def __init__(self, path):
        err = C.create_string_buffer(256)
        self.h = oracle().orc_db_load(os.fsencode(path), err, 256)
        if not self.h:
            raise RuntimeError("oracle: " + err.value.decode())
        self.max_ix = oracle().orc_db_max_ix(self.h)

    def lookup(self, word):
        return oracle().orc_lookup(self.h, int(word), None)

    def lookup_many(self, words):
        return np.array([oracle().orc_lookup(self.h, int(w), None) for w in words], dtype=np.uint32)

    def label(self, ix):
        return oracle().orc_db_label(self.h, ix)

    def slide(self, seq: bytes, do_rc=True, want_words=False):
        cap = 2 * len(seq) + 2
        hits = np.empty(cap, dtype=np.uint32)
        words = np.empty(cap if want_words else 1, dtype=np.uint64)
        nw = C.c_uint64()
        nf = oracle().orc_slide(self.h, seq, len(seq), int(do_rc), hits.ctypes.data, cap,
                                words.ctypes.data if want_words else None, cap if want_words else 0,
                                C.byref(nw), None)
        return hits[:nf].copy(), (words[:nw.value].copy() if want_words else None)

    def vote(self, hits):
        hits = np.ascontiguousarray(hits, dtype=np.uint32)
        v = OrcVote()
        oracle().orc_vote(self.h, hits.ctypes.data, hits.size, C.byref(v))
        return v

    def search_file(self, fasta, out, do_rc=True, threads=1, max_reads=0):
        st = OrcStats()
        err = C.create_string_buffer(256)
        rc = oracle().orc_search_file(self.h, os.fsencode(fasta), os.fsencode(out), int(do_rc), threads,
                                      max_reads, C.byref(st), err, 256)
        return rc, st.as_dict(), err.value.decode()

    def search_file_shallow(self, fasta, out, do_rc=True):
        """The non-GG binary (-D SEARCH): SPARSITY-skip slide + shallow top-2 vote, sequential."""
        st = OrcStats()
        err = C.create_string_buffer(256)
        rc = oracle().orc_search_file_shallow(self.h, os.fsencode(fasta), os.fsencode(out), int(do_rc), C.byref(st), err, 256)
        return rc, st.as_dict(), err.value.decode()

    def free(self):
        if self.h:
            oracle().orc_db_free(self.h)
            self.h = None
