"""ctypes binding of oracle/libstrk_oracle.so -- the CPU checker.  TEST INFRASTRUCTURE ONLY.

Nothing under strkit_b200/ imports this module; tests, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs do.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libstrk_oracle.so")

MODE_SG = 15
MODE_SG_QE = 2
GAP = 5

_i32p = C.POINTER(C.c_int32)


def build(force: bool = False) -> str:
    src = os.path.join(ORACLE_DIR, "strk_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return LIB_PATH


class Oracle:
    def __init__(self, lib: C.CDLL):
        self.lib = lib
        lib.strk_oracle_dna_matrix.argtypes = [C.c_void_p]
        lib.strk_oracle_symbol.argtypes = [C.c_ubyte]
        lib.strk_oracle_sg_align.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                             C.c_int, _i32p, _i32p, _i32p]
        lib.strk_oracle_sg_align_simd.argtypes = lib.strk_oracle_sg_align.argtypes
        lib.strk_oracle_set_simd.argtypes = [C.c_int]
        lib.strk_oracle_score_candidate.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int,
                                                    C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, _i32p]
        lib.strk_oracle_get_repeat_count.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_char_p,
                                                     C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                     C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        lib.strk_oracle_score_ref_boundaries.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int,
                                                         C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                                         C.c_void_p]
        lib.strk_oracle_get_ref_repeat_count.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_char_p,
                                                         C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                                         C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                                         C.c_int, C.c_void_p]
        lib.strk_oracle_count_loci.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                               C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_double)]
        lib.strk_oracle_realign.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                            _i32p, _i32p, C.c_void_p, C.c_int, _i32p]
        self.matrix = np.zeros(17 * 17, dtype=np.int8)
        lib.strk_oracle_dna_matrix(self.matrix.ctypes.data)

    def symbol(self, ch: str) -> int:
        return self.lib.strk_oracle_symbol(ord(ch))

    def set_simd(self, on: bool) -> bool:
        """Route the search / batch entry points through the AVX2 scan kernel (the timed CPU baseline); returns the
        previous setting.  The checker default is the scalar restatement."""
        return bool(self.lib.strk_oracle_set_simd(int(on)))

    def have_simd(self) -> bool:
        return bool(self.lib.strk_oracle_have_simd())

    def sg_align(self, s1: str, s2: str, flags: int, gap_open: int = GAP, gap_extend: int = GAP, simd: bool = False):
        sc, eq, er = C.c_int32(), C.c_int32(), C.c_int32()
        b1, b2 = s1.encode(), s2.encode()
        fn = self.lib.strk_oracle_sg_align_simd if simd else self.lib.strk_oracle_sg_align
        rc = fn(b1, len(b1), b2, len(b2), gap_open, gap_extend, self.matrix.ctypes.data,
                                           flags, C.byref(sc), C.byref(eq), C.byref(er))
        if rc:
            raise ValueError(f"oracle sg_align failed: {rc}")
        return sc.value, eq.value, er.value

    def realign(self, ref_seq: str, query_seq: str, gap_open: int = 7, gap_extend: int = 0, trace_flags: int = 0):
        """parasail.sg_dx_trace_scan_16(ref_seq, query_seq, open, extend) restated: (score, end_ref, cigar uint32[])."""
        b1, b2 = ref_seq.encode(), query_seq.encode()
        cap = 2 * len(b1) + 4
        cigar = np.zeros(cap, dtype=np.uint32)
        sc, er, ln = C.c_int32(), C.c_int32(), C.c_int32()
        rc = self.lib.strk_oracle_realign(b1, len(b1), b2, len(b2), gap_open, gap_extend, self.matrix.ctypes.data,
                                          trace_flags, C.byref(sc), C.byref(er), cigar.ctypes.data, cap, C.byref(ln))
        if rc:
            raise ValueError(f"oracle realign failed: {rc}")
        return sc.value, er.value, cigar[:ln.value].copy()

    def score_candidate(self, tr: str, fl: str, fr: str, motif: str, n: int, flags: int = MODE_SG) -> int:
        db = (fl + tr + fr).encode()
        sc = C.c_int32()
        rc = self.lib.strk_oracle_score_candidate(db, len(db), fl.encode(), len(fl), fr.encode(), len(fr),
                                                  motif.encode(), len(motif), n, GAP, self.matrix.ctypes.data, flags,
                                                  C.byref(sc))
        if rc:
            raise ValueError(f"oracle score_candidate failed: {rc}")
        return sc.value

    def get_repeat_count(self, start_count, tr, fl, fr, motif, max_iters, local_search_range, step_size,
                         flags: int = MODE_SG, tie_flags: int = 0):
        out = np.zeros(4, dtype=np.int32)
        rc = self.lib.strk_oracle_get_repeat_count(start_count, tr.encode(), len(tr), fl.encode(), len(fl),
                                                   fr.encode(), len(fr), motif.encode(), len(motif), max_iters,
                                                   local_search_range, step_size, GAP, self.matrix.ctypes.data, flags,
                                                   tie_flags, out.ctypes.data)
        if rc:
            raise ValueError(f"oracle get_repeat_count failed: {rc}")
        return (int(out[0]), int(out[1])), int(out[2]), int(out[3])

    def score_ref_boundaries(self, tr, fl, fr, motif, n, ref_size):
        db = (fl + tr + fr).encode()
        out = np.zeros(4, dtype=np.int32)
        rc = self.lib.strk_oracle_score_ref_boundaries(db, len(db), fl.encode(), len(fl), fr.encode(), len(fr),
                                                       motif.encode(), len(motif), n, ref_size, GAP,
                                                       self.matrix.ctypes.data, out.ctypes.data)
        if rc:
            raise ValueError(f"oracle score_ref_boundaries failed: {rc}")
        return (int(out[0]), int(out[1])), (int(out[2]), int(out[3]))

    def get_ref_repeat_count(self, start_count, tr, fl, fr, motif, ref_size, vcf_anchor_size, max_iters,
                             local_search_range, step_size, respect_coords=False, flags: int = MODE_SG,
                             tie_flags: int = 0):
        out = np.zeros(8, dtype=np.int32)
        rc = self.lib.strk_oracle_get_ref_repeat_count(start_count, tr.encode(), len(tr), fl.encode(), len(fl),
                                                       fr.encode(), len(fr), motif.encode(), len(motif), ref_size,
                                                       vcf_anchor_size, max_iters, local_search_range, step_size,
                                                       int(respect_coords), GAP, self.matrix.ctypes.data, flags,
                                                       tie_flags, out.ctypes.data)
        if rc:
            raise ValueError(f"oracle get_ref_repeat_count failed: {rc}")
        db = fl + tr + fr
        nfl, nfr = int(out[6]), int(out[7])
        fl2, tr2, fr2 = db[:nfl], db[nfl:len(db) - nfr], db[len(db) - nfr:]
        return ((int(out[0]), int(out[1])), int(out[2]), int(out[3]), (int(out[4]), int(out[5])), (fl2, tr2, fr2))

    def count_loci(self, arena, seq_off, lens, est_cn, read_begin, motif_off, motif_len, max_iters=50,
                   local_search_range=3, step_size=1, flags=MODE_SG, tie_flags=0, n_threads=1):
        """Batch read loop (call_locus.py:1129-1161).  Returns (out[n_reads,4], reference-equivalent cells)."""
        arena = np.ascontiguousarray(arena, dtype=np.uint8)
        seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        est_cn = np.ascontiguousarray(est_cn, dtype=np.int32)
        read_begin = np.ascontiguousarray(read_begin, dtype=np.int64)
        motif_off = np.ascontiguousarray(motif_off, dtype=np.uint64)
        motif_len = np.ascontiguousarray(motif_len, dtype=np.int32)
        n_reads = int(est_cn.shape[0])
        out = np.zeros((n_reads, 4), dtype=np.int32)
        cells = C.c_double(0.0)
        rc = self.lib.strk_oracle_count_loci(arena.ctypes.data, seq_off.ctypes.data, lens.ctypes.data,
                                             est_cn.ctypes.data, read_begin.ctypes.data, motif_off.ctypes.data,
                                             motif_len.ctypes.data, len(motif_len), max_iters, local_search_range,
                                             step_size, GAP, self.matrix.ctypes.data, flags, tie_flags, n_threads,
                                             out.ctypes.data, C.byref(cells))
        if rc:
            raise ValueError(f"oracle count_loci failed: {rc}")
        return out, cells.value


def load() -> Oracle:
    return Oracle(C.CDLL(build()))
