"""Reads into memory (first half): the library's FASTA / FASTQ (.gz) reader, which follows the reference's
ReadFiles / kseq record grammar (ReadFiles.hpp:155-204) and returns the fixed-stride NUL-padded arrays Genotyper.Genotype takes."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def load_reads(path1, path2=None):
    """-> (reads1 uint8 [n, stride], reads2 or None)"""
    r = L.Reads()
    L.check(L.lib().t1k_reads_load(str(path1).encode(), None if path2 is None else str(path2).encode(), C.byref(r)))
    try:
        shape = (r.n_frag, r.stride)
        a = np.ctypeslib.as_array(C.cast(r.reads1, C.POINTER(C.c_uint8)), shape=shape).copy() if r.n_frag else np.zeros((0, r.stride), np.uint8)
        b = None
        if path2 is not None:
            b = np.ctypeslib.as_array(C.cast(r.reads2, C.POINTER(C.c_uint8)), shape=shape).copy() if r.n_frag else np.zeros((0, r.stride), np.uint8)
    finally:
        L.lib().t1k_reads_free(C.byref(r))
    return a, b
