"""Read-sharded multi-GPU run of the hot path: one process per GPU, `torch.distributed` only as the
launcher-side plumbing that carries the 128-byte NCCL unique id from rank 0 to the other ranks; everything on the data
path (base-coverage all-reduce, read-group table all-gather, one fp64 all-reduce of per-EC read counts per EMupdate)
is NCCL over NVLink inside the C library.

    comm = Comm.from_torch_distributed(device)          # after dist.init_process_group(...)
    out = genotype_sharded(gt, reads1_shard, reads2_shard, comm)

Host-side helpers (`ReadGroups`, `em_partition`) wrap the library's model steps; they need no device and are what the
world_size-2 gloo tests exercise on the CPU box.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class Comm:
    """One NCCL rank (T1KComm)."""

    def __init__(self, unique_id: bytes, rank: int, world: int, device: int = -1):
        assert len(unique_id) == L.UNIQUE_ID_BYTES
        self.rank, self.world = rank, world
        buf = np.frombuffer(unique_id, dtype=np.uint8).copy()
        h = C.c_void_p()
        L.check(L.lib().t1k_comm_create(L.ptr(buf), rank, world, device, C.byref(h)))
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            L.lib().t1k_comm_destroy(self.h)
            self.h = None

    @staticmethod
    def unique_id() -> bytes:
        buf = np.zeros(L.UNIQUE_ID_BYTES, dtype=np.uint8)
        L.check(L.lib().t1k_comm_unique_id(L.ptr(buf)))
        return buf.tobytes()

    @classmethod
    def from_torch_distributed(cls, device: int = -1):
        """rank 0 makes the id, torch.distributed (any backend) broadcasts it."""
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(box[0], rank, world, device)


def shard_bounds(n: int, world: int):
    """contiguous fragment ranges, sizes differing by at most one"""
    return [n * r // world for r in range(world + 1)]


def genotype_sharded(gt, reads1, reads2, comm: Comm):
    """reads1/reads2: THIS rank's fragments.  Per-allele results are the whole job's, identical on every rank."""
    return gt.Genotype(reads1, reads2, comm=comm)


def em_partition(row_ptr, world: int):
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    out = np.zeros(world + 1, dtype=np.int32)
    L.check(L.lib().t1k_em_partition(L.ptr(row_ptr), len(row_ptr) - 1, world, L.ptr(out)))
    return out


class ReadGroups:
    """Genotyper::CoalesceReadAssignments state (host): add fragments in order, serialize, merge peers in rank order."""

    def __init__(self):
        h = C.c_void_p()
        L.check(L.lib().t1k_groups_create(C.byref(h)))
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            L.lib().t1k_groups_destroy(self.h)
            self.h = None

    def add_fragments(self, row_ptr, entries):
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        entries = np.ascontiguousarray(entries, dtype=L.ASSIGN_DT)
        L.check(L.lib().t1k_groups_add_fragments(self.h, L.ptr(row_ptr), L.ptr(entries), len(row_ptr) - 1))

    def serialize(self) -> bytes:
        p, n = C.c_void_p(), C.c_uint64(0)
        L.check(L.lib().t1k_groups_serialize(self.h, C.byref(p), C.byref(n)))
        try:
            return C.string_at(p, n.value)
        finally:
            L.lib().t1k_free(p)

    def merge(self, blob: bytes):
        buf = np.frombuffer(blob, dtype=np.uint8)
        L.check(L.lib().t1k_groups_merge(self.h, L.ptr(buf), len(buf)))

    def ec_filter(self, allele_len, ec_abundance, ec_allele_ptr, ec_alleles):
        """Genotyper::RemoveLowLikelihoodAlleleInEquivalentClass (Genotyper.hpp:1371-1460) over this table ->
        (allele_kept uint8[n_alleles], allele_span int32[2 * n_alleles])"""
        ln = np.ascontiguousarray(allele_len, dtype=np.int32)
        ab = np.ascontiguousarray(ec_abundance, dtype=np.float64)
        ep = np.ascontiguousarray(ec_allele_ptr, dtype=np.int32)
        ea = np.ascontiguousarray(ec_alleles, dtype=np.int32)
        kept = np.zeros(len(ln), dtype=np.uint8)
        span = np.zeros(2 * len(ln), dtype=np.int32)
        L.check(L.lib().t1k_groups_ec_filter(self.h, len(ln), L.ptr(ln), L.ptr(ab), L.ptr(ep), L.ptr(ea), len(ep) - 1, L.ptr(kept), L.ptr(span)))
        return kept, span

    def fetch(self):
        """-> (ptr[n_groups+1], entries, assigned_fragments)"""
        ng, ne, na = C.c_int32(0), C.c_uint64(0), C.c_uint64(0)
        L.check(L.lib().t1k_groups_fetch(self.h, C.byref(ng), C.byref(ne), C.byref(na), None, None))
        ptr = np.zeros(ng.value + 1, dtype=np.int64)
        ent = np.zeros(ne.value, dtype=L.ASSIGN_DT)
        L.check(L.lib().t1k_groups_fetch(self.h, C.byref(ng), C.byref(ne), C.byref(na), L.ptr(ptr), L.ptr(ent)))
        return ptr, ent, int(na.value)
