"""ctypes binding of the C ABI declared in include/t1k_b200.h (t1k_b200/libt1k_b200.so).

There is no CPU fallback: if the CUDA library has not been built, loading fails loudly
(build it with `python -c "import __graft_entry__ as g; g.build()"` or `make -C t1k_b200/csrc`).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libt1k_b200.so")

OVERLAP_DT = np.dtype([(n, "<i4") for n in ("seqIdx", "readStart", "readEnd", "seqStart", "seqEnd", "strand",
                                             "matchCnt", "relaxedMatchCnt", "leftClip", "rightClip")])
ASSIGN_DT = np.dtype([("alleleIdx", "<i4"), ("start", "<i4"), ("end", "<i4"),
                      ("weight", "<f4"), ("qual", "<f4"), ("adjustWeight", "<f4")])

T1K_OK, T1K_ERR_NO_DEVICE, T1K_ERR_CUDA, T1K_ERR_ARG, T1K_ERR_UNSUPPORTED, T1K_ERR_NCCL = range(6)

# every symbol include/t1k_b200.h declares
EXPORTS = ["t1k_last_error", "t1k_device_count", "t1k_ref_create", "t1k_ref_destroy", "t1k_ref_n_alleles",
           "t1k_assign_batch", "t1k_assignment_destroy", "t1k_assignment_fetch", "t1k_assignment_stats",
           "t1k_coverage_fetch", "t1k_coverage_reset", "t1k_missing_coverage", "t1k_pair_batch", "t1k_free",
           "t1k_em_run", "t1k_genotype", "t1k_comm_unique_id", "t1k_comm_create", "t1k_comm_destroy",
           "t1k_coverage_allreduce", "t1k_groups_create", "t1k_groups_destroy", "t1k_groups_add_fragments",
           "t1k_groups_serialize", "t1k_groups_merge", "t1k_groups_fetch", "t1k_em_partition",
           "t1k_filter_create", "t1k_filter_destroy", "t1k_filter_batch", "t1k_align_info_batch", "t1k_dpx_peak", "t1k_groups_ec_filter",
           "t1k_assign_batch_async", "t1k_assign_wait", "t1k_pinned_alloc", "t1k_pinned_free", "t1k_reads_load", "t1k_reads_free"]

UNIQUE_ID_BYTES = 128


class RefDesc(C.Structure):
    _fields_ = [("n_alleles", C.c_int32), ("bases", C.c_char_p), ("offset", C.c_void_p), ("exon_ptr", C.c_void_p),
                ("exon_se", C.c_void_p), ("similarity", C.c_double), ("relax_intron", C.c_int32), ("device", C.c_int32)]


class EmProblem(C.Structure):
    _fields_ = [("n_groups", C.c_int32), ("n_ec", C.c_int32), ("row_ptr", C.c_void_p), ("col", C.c_void_p),
                ("count", C.c_void_p), ("ec_len", C.c_void_p), ("x0", C.c_void_p), ("min_squarem_alpha", C.c_double),
                ("filter_frac", C.c_double), ("n_alleles", C.c_int32), ("n_major", C.c_int32), ("n_gene", C.c_int32),
                ("ec_allele_ptr", C.c_void_p), ("ec_alleles", C.c_void_p), ("allele_major", C.c_void_p),
                ("allele_gene", C.c_void_p), ("fast_sums", C.c_int32), ("comm", C.c_void_p)]


class EmResult(C.Structure):
    _fields_ = [("x", C.c_void_p), ("ec_read_count", C.c_void_p), ("iterations", C.c_int32), ("ms_kernel", C.c_float),
                ("n_launches", C.c_uint64)]


class GenotypeParams(C.Structure):
    _fields_ = [("max_assign", C.c_int32), ("min_squarem_alpha", C.c_double), ("filter_frac", C.c_double),
                ("seq_weight", C.c_void_p), ("effective_len", C.c_void_p), ("allele_major", C.c_void_p),
                ("allele_gene", C.c_void_p), ("n_major", C.c_int32), ("n_gene", C.c_int32), ("em_fast_sums", C.c_int32),
                ("comm", C.c_void_p), ("groups_out", C.c_void_p)]


class GenotypeResult(C.Structure):
    _fields_ = [("n_alleles", C.c_int32), ("abundance", C.c_void_p), ("ec_abundance", C.c_void_p),
                ("equivalent_class", C.c_void_p), ("missing_coverage", C.c_void_p), ("fragment_assigned", C.c_void_p),
                ("em_iterations", C.c_int32), ("n_groups", C.c_int32), ("n_ec", C.c_int32),
                ("assigned_fragments", C.c_int32), ("n_unique_ends", C.c_uint64), ("n_overlaps", C.c_uint64),
                ("n_assignments", C.c_uint64), ("avg_alleles_per_read", C.c_double), ("ms_dedup", C.c_float),
                ("ms_align", C.c_float), ("ms_pair", C.c_float), ("ms_coalesce", C.c_float), ("ms_em", C.c_float),
                ("ms_align_kernel", C.c_float), ("ms_pair_kernel", C.c_float), ("ms_em_kernel", C.c_float),
                ("n_postings", C.c_uint64), ("n_candidates", C.c_uint64), ("n_launches", C.c_uint64),
                ("ms_prep_wait", C.c_float), ("ms_exchange", C.c_float), ("n_pair_records", C.c_uint64), ("em_nnz", C.c_uint64),
                ("em_updates", C.c_int32), ("ec_read_count", C.c_void_p),
                ("ec_allele_ptr", C.c_void_p), ("ec_alleles", C.c_void_p), ("allele_kept", C.c_void_p), ("allele_span", C.c_void_p)]


class FilterDesc(C.Structure):
    _fields_ = [("n_seqs", C.c_int32), ("bases", C.c_char_p), ("offset", C.c_void_p), ("kmer_length", C.c_int32),
                ("hit_len_required", C.c_int32), ("similarity", C.c_double), ("device", C.c_int32)]


class FilterStats(C.Structure):
    _fields_ = [("windows", C.c_uint64), ("entries", C.c_uint64), ("chained", C.c_uint64), ("ms_kernel", C.c_float),
                ("kmer_length", C.c_int32)]


class AlignInfoStats(C.Structure):
    _fields_ = [("n_diagonal", C.c_uint64), ("n_dp", C.c_uint64), ("dp_cells", C.c_uint64), ("ms_kernel", C.c_float)]


class Reads(C.Structure):
    _fields_ = [("reads1", C.c_void_p), ("reads2", C.c_void_p), ("stride", C.c_uint32), ("n_frag", C.c_uint32), ("max_len", C.c_uint32),
                ("pinned", C.c_int32)]


class AssignStats(C.Structure):
    _fields_ = [("postings", C.c_uint64), ("candidates", C.c_uint64), ("tiles", C.c_uint64), ("records", C.c_uint64),
                ("ms_kernel", C.c_float), ("grid_blocks", C.c_int32), ("hit_cap", C.c_int32), ("n_sm", C.c_int32)]


class T1KError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("t1k_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError("t1k_b200: %s is missing — the CUDA library must be built (there is no CPU path)" % SO_PATH)
        L = C.CDLL(SO_PATH)
        L.t1k_last_error.restype = C.c_char_p
        L.t1k_device_count.argtypes = [C.POINTER(C.c_int)]
        L.t1k_ref_create.argtypes = [C.POINTER(RefDesc), C.POINTER(C.c_void_p)]
        L.t1k_ref_destroy.argtypes = [C.c_void_p]
        L.t1k_ref_destroy.restype = None
        L.t1k_ref_n_alleles.argtypes = [C.c_void_p]
        L.t1k_assign_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                       C.POINTER(C.c_void_p)]
        L.t1k_assign_batch_async.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]
        L.t1k_assign_wait.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.t1k_pinned_alloc.argtypes = [C.c_uint64, C.POINTER(C.c_void_p)]
        L.t1k_pinned_free.argtypes = [C.c_void_p]
        L.t1k_pinned_free.restype = None
        L.t1k_assignment_destroy.argtypes = [C.c_void_p]
        L.t1k_assignment_destroy.restype = None
        L.t1k_assignment_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        L.t1k_assignment_stats.argtypes = [C.c_void_p, C.POINTER(AssignStats)]
        L.t1k_coverage_fetch.argtypes = [C.c_void_p, C.c_void_p]
        L.t1k_coverage_reset.argtypes = [C.c_void_p]
        L.t1k_missing_coverage.argtypes = [C.c_void_p, C.c_void_p]
        L.t1k_pair_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32,
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p]
        L.t1k_free.argtypes = [C.c_void_p]
        L.t1k_free.restype = None
        L.t1k_em_run.argtypes = [C.POINTER(EmProblem), C.POINTER(EmResult), C.c_int32]
        L.t1k_genotype.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                   C.POINTER(GenotypeParams), C.POINTER(GenotypeResult)]
        L.t1k_comm_unique_id.argtypes = [C.c_void_p]
        L.t1k_comm_create.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
        L.t1k_comm_destroy.argtypes = [C.c_void_p]
        L.t1k_comm_destroy.restype = None
        L.t1k_coverage_allreduce.argtypes = [C.c_void_p, C.c_void_p]
        L.t1k_groups_create.argtypes = [C.POINTER(C.c_void_p)]
        L.t1k_groups_destroy.argtypes = [C.c_void_p]
        L.t1k_groups_destroy.restype = None
        L.t1k_groups_add_fragments.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.t1k_groups_serialize.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.t1k_groups_merge.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.t1k_groups_fetch.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                       C.c_void_p, C.c_void_p]
        L.t1k_groups_ec_filter.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.t1k_em_partition.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.t1k_filter_create.argtypes = [C.POINTER(FilterDesc), C.POINTER(C.c_void_p)]
        L.t1k_filter_destroy.argtypes = [C.c_void_p]
        L.t1k_filter_destroy.restype = None
        L.t1k_filter_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(FilterStats)]
        L.t1k_align_info_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                                           C.c_int32, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(AlignInfoStats)]
        L.t1k_dpx_peak.argtypes = [C.c_int32, C.POINTER(C.c_double)]
        L.t1k_reads_load.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(Reads)]
        L.t1k_reads_free.argtypes = [C.POINTER(Reads)]
        L.t1k_reads_free.restype = None
        _lib = L
    return _lib


def check(rc):
    if rc != T1K_OK:
        raise T1KError(rc, lib().t1k_last_error().decode(errors="replace"))


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None
