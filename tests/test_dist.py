"""The read-sharded (N>1) path.

CPU part (world_size 2, gloo): the host-side protocol of the sharded run — contiguous fragment shards, per-rank
read-group tables (t1k_groups_*), the rank-order merge every rank performs after the all-gather, the row partition
of the EM and the one all-reduce per EMupdate — checked against the single-process oracle.  The device kernels cannot
run here; the per-fragment rows come from the oracle (the checker) and the partial E-step is restated in numpy.

GPU part (needs >= 2 devices; run with `gpurun --gpus 2`): two NCCL ranks through the C ABI against the one-GPU run.
"""
import os
import pickle
import socket
import sys
import tempfile

import numpy as np
import pytest

import oracle_py as O
import workloads as W
from t1k_b200 import dist_em
from t1k_b200._lib import ASSIGN_DT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rows_to_csr(frags):
    ptr = np.zeros(len(frags) + 1, dtype=np.uint64)
    ptr[1:] = np.cumsum([len(f) for f in frags])
    ent = np.concatenate(frags) if len(frags) and ptr[-1] else np.zeros(0, dtype=ASSIGN_DT)
    return ptr, ent.astype(ASSIGN_DT)


def _estep_partial(rowptr, col, count, x, g0, g1, n_ec):
    """rows [g0, g1) of Genotyper::EMupdate's E-step (Genotyper.hpp:379-404)"""
    rc = np.zeros(n_ec)
    for g in range(g0, g1):
        c = col[rowptr[g]:rowptr[g + 1]]
        ps = x[c].sum()
        if ps == 0:
            ps = 1.0
        np.add.at(rc, c, count[g] * x[c] / ps)
    return rc


def _cpu_worker(rank, world, port, path):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        with open(path, "rb") as f:
            D = pickle.load(f)
        frags = D["frags"]
        b = dist_em.shard_bounds(len(frags), world)
        mine = frags[b[rank]:b[rank + 1]]
        local = dist_em.ReadGroups()
        local.add_fragments(*_rows_to_csr(mine))
        blobs = [None] * world
        dist.all_gather_object(blobs, local.serialize())          # the library does this with ncclAllGather
        merged = dist_em.ReadGroups()
        for r in range(world):
            merged.merge(blobs[r])
        ptr, ent, assigned = merged.fetch()
        # sharded EM: same problem on every rank, E-step over the rank's rows, one all-reduce per EMupdate
        P = D["problem"]
        rowptr, col, count, eclen, x0 = P["rowptr"], P["col"], P["count"], P["eclen"], P["x0"]
        bounds = dist_em.em_partition(rowptr, world)
        E = len(eclen)

        def em_update(x):
            import torch
            rc = torch.from_numpy(_estep_partial(rowptr, col, count, x, int(bounds[rank]), int(bounds[rank + 1]), E))
            dist.all_reduce(rc)
            rc = rc.numpy()
            y = rc / eclen
            return y / y.sum(), rc

        x = x0.copy()
        iters = 0
        t = 0
        while t < 1000:                                              # Genotyper.hpp:1234-1290 without the mask
            iters += 1
            x1, _ = em_update(x)
            x2, _ = em_update(x1)
            r, v = x1 - x, x2 - 2 * x1 + x
            sv = np.sqrt((v * v).sum())
            alpha = -1.0 if sv == 0 else -np.sqrt((r * r).sum()) / sv
            x3 = x - 2 * alpha * r + alpha * alpha * v
            xn, rc = em_update(x3)
            diff = np.abs(xn - x).sum()
            x = xn
            if diff < 1e-5 and t < 998:
                t = 998
            t += 1
        out = dict(ptr=ptr, ent=ent, assigned=assigned, bounds=bounds, iters=iters, x=x, rc=rc)
        with open("%s.out%d" % (path, rank), "wb") as f:
            pickle.dump(out, f)
    finally:
        dist.destroy_process_group()


@pytest.fixture(scope="module")
def sharded_cpu_run():
    import torch.multiprocessing as mp
    recs = W.small_dna_ref(seed=51)
    kept, w = O.collapse_reference(recs)
    sw = O.seq_weights(kept, w)
    orc = O.Oracle(kept, 0.9, True, sw)
    r1, r2 = W.reads_for(kept, 240, read_len=100, seed=52)
    R = O.genotype_pipeline(orc, r1, r2, [k[0] for k in kept], sw)
    it, x, rc = O.em(R["problem"]["rowptr"], R["problem"]["col"], R["problem"]["count"], R["problem"]["eclen"], R["problem"]["x0"])
    td = tempfile.mkdtemp(prefix="t1kdist_")
    path = os.path.join(td, "in.pkl")
    with open(path, "wb") as f:
        pickle.dump(dict(frags=R["frags"], problem=R["problem"]), f)
    world = 2
    mp.spawn(_cpu_worker, args=(world, _free_port(), path), nprocs=world, join=True)
    outs = [pickle.load(open("%s.out%d" % (path, r), "rb")) for r in range(world)]
    return R, (it, x, rc), outs


def test_rank_order_merge_equals_single_process_coalescing(sharded_cpu_run):
    R, _, outs = sharded_cpu_run
    groups = R["groups"]
    want_ptr = np.zeros(len(groups) + 1, dtype=np.int64)
    want_ptr[1:] = np.cumsum([len(g) for g in groups])
    want = np.concatenate(groups)
    for o in outs:                                                  # every rank ends with the same table
        assert o["assigned"] == R["assigned"]
        assert np.array_equal(o["ptr"], want_ptr)                   # same groups in the same (first-appearance) order
        assert np.array_equal(o["ent"]["alleleIdx"], want["alleleIdx"])
        assert np.array_equal(o["ent"]["qual"], want["qual"])
        # float32 sums: (rank 0's partial) + (rank 1's partial) instead of fragment by fragment
        np.testing.assert_allclose(o["ent"]["weight"], want["weight"], rtol=1e-6)
        np.testing.assert_allclose(o["ent"]["adjustWeight"], want["adjustWeight"], rtol=1e-6)
    assert np.array_equal(outs[0]["ent"], outs[1]["ent"])


def test_single_process_groups_match_the_oracle(sharded_cpu_run):
    """t1k_groups_add_fragments over all fragments in order == Genotyper::CoalesceReadAssignments, bit for bit."""
    R, _, _ = sharded_cpu_run
    g = dist_em.ReadGroups()
    g.add_fragments(*_rows_to_csr(R["frags"]))
    ptr, ent, assigned = g.fetch()
    want = np.concatenate(R["groups"])
    assert assigned == R["assigned"] and len(ptr) - 1 == len(R["groups"])
    for name in ASSIGN_DT.names:
        assert np.array_equal(ent[name], want[name]), name


def test_row_sharded_em_matches_the_oracle(sharded_cpu_run):
    R, (it, x, rc), outs = sharded_cpu_run
    rowptr = R["problem"]["rowptr"]
    for o in outs:
        b = o["bounds"]
        assert b[0] == 0 and b[-1] == len(rowptr) - 1 and (np.diff(b) >= 0).all()
        assert o["iters"] == it
        np.testing.assert_allclose(o["rc"], rc, rtol=1e-9, atol=1e-9 * rc.max())
        np.testing.assert_allclose(o["x"], x, rtol=1e-9, atol=1e-9 * x.max())    # EM dust moves with the summation order
    assert np.array_equal(outs[0]["x"], outs[1]["x"])               # replicated M-step: identical on every rank
    # the partition balances non-zeros
    nnz = [rowptr[outs[0]["bounds"][r + 1]] - rowptr[outs[0]["bounds"][r]] for r in range(2)]
    assert abs(nnz[0] - nnz[1]) <= max(np.diff(rowptr))


def test_em_partition_edges():
    assert dist_em.em_partition([0], 4).tolist() == [0, 0, 0, 0, 0]
    assert dist_em.em_partition([0, 5], 2).tolist() == [0, 1, 1] or dist_em.em_partition([0, 5], 2).tolist() == [0, 0, 1]
    b = dist_em.em_partition(np.arange(0, 101, 10), 4)
    assert b[0] == 0 and b[-1] == 10 and (np.diff(b) > 0).all()
    assert dist_em.shard_bounds(10, 4) == [0, 2, 5, 7, 10]


# ---------------------------------------------------------------------------------------------------------
def _gpu_worker(rank, world, port, path):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        from t1k_b200.genotyper import Genotyper, QuantifyAlleleEquivalentClass
        from t1k_b200.refset import RefSet
        with open(path, "rb") as f:
            D = pickle.load(f)
        comm = dist_em.Comm.from_torch_distributed(rank)
        ref = RefSet(D["recs"])
        gt = Genotyper(ref, D["sim"], D["relax"], device=rank)
        r1, r2 = D["r1"], D["r2"]
        b = dist_em.shard_bounds(len(r1), world)
        out = dist_em.genotype_sharded(gt, r1[b[rank]:b[rank + 1]], r2[b[rank]:b[rank + 1]], comm)
        P = D["problem"]
        it, x, rc, _ = QuantifyAlleleEquivalentClass(P["rowptr"], P["col"], P["count"], P["eclen"], P["x0"], device=rank, comm=comm)
        res = {k: out[k] for k in ("abundance", "ec_abundance", "equivalent_class", "missing_coverage", "fragment_assigned",
                                   "em_iterations", "n_groups", "n_ec", "assigned_fragments", "avg_alleles_per_read")}
        res.update(em=(it, x, rc))
        with open("%s.out%d" % (path, rank), "wb") as f:
            pickle.dump(res, f)
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_two_rank_nccl_run_matches_one_gpu():
    import ctypes
    import torch.multiprocessing as mp
    from t1k_b200 import _lib as L
    from t1k_b200.genotyper import Genotyper
    from t1k_b200.refset import RefSet
    n = ctypes.c_int(0)
    L.lib().t1k_device_count(ctypes.byref(n))
    if n.value < 2:
        pytest.skip("needs two CUDA devices")
    recs = W.small_dna_ref(seed=61)
    kept, w = O.collapse_reference(recs)
    sw = O.seq_weights(kept, w)
    r1, r2 = W.reads_for(kept, 600, read_len=125, seed=62, insert=(200, 420))
    ref = RefSet(recs)
    one = Genotyper(ref, 0.9, True, device=0).Genotype(r1, r2)
    R = O.genotype_pipeline(O.Oracle(kept, 0.9, True, sw), r1[:200], r2[:200], ref.names, sw)
    oit, ox, orc_ = O.em(R["problem"]["rowptr"], R["problem"]["col"], R["problem"]["count"], R["problem"]["eclen"], R["problem"]["x0"])
    td = tempfile.mkdtemp(prefix="t1kdist_")
    path = os.path.join(td, "in.pkl")
    with open(path, "wb") as f:
        pickle.dump(dict(recs=recs, sim=0.9, relax=True, r1=r1, r2=r2, problem=R["problem"]), f)
    world = 2
    mp.spawn(_gpu_worker, args=(world, _free_port(), path), nprocs=world, join=True)
    outs = [pickle.load(open("%s.out%d" % (path, r), "rb")) for r in range(world)]
    b = dist_em.shard_bounds(len(r1), world)
    for r, o in enumerate(outs):
        # integer outputs: exactly the one-GPU run
        for k in ("equivalent_class", "missing_coverage"):
            assert np.array_equal(o[k], one[k]), k
        for k in ("n_groups", "n_ec", "assigned_fragments"):
            assert o[k] == one[k], k
        assert np.array_equal(o["fragment_assigned"], one["fragment_assigned"][b[r]:b[r + 1]])
        assert abs(o["avg_alleles_per_read"] - one["avg_alleles_per_read"]) < 1e-9
        # abundances: north-star tolerance (sums run in a different order once rows are sharded)
        scale = float(np.abs(one["abundance"]).max())
        np.testing.assert_allclose(o["abundance"], one["abundance"], rtol=1e-5, atol=1e-5 * scale)
        it, x, rc = o["em"]
        assert it == oit
        np.testing.assert_allclose(rc, orc_, rtol=1e-9, atol=1e-9 * orc_.max())
    assert np.array_equal(outs[0]["abundance"], outs[1]["abundance"])       # all-reduced: identical on every rank
