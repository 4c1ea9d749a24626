"""CPU tests of the N>1 path: pattern shards are contiguous, disjoint and cover every
pattern; per-shard partial log-likelihoods summed over ranks (gloo all-reduce,
world_size 2, two real processes) equal the unsharded value.  The per-shard
arithmetic here is the oracle port -- the GPU engine applies the same rule on the
device and all-reduces with NCCL (tests/test_gpu_parity.py covers it on hardware)."""
import os
import socket

import numpy as np
import pytest


def test_shard_ranges_partition_patterns(pkg):
    pf = pkg.pf
    for n in (0, 1, 7, 31, 32, 33, 1000, 1000000, 999983):
        for world in (1, 2, 3, 4, 8):
            edges = [pf.shardRangeFor(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for (a, b), (c, d) in zip(edges[:-1], edges[1:]):
                assert b == c and a <= b
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, root, out):
    import sys
    for p in (root, os.path.join(root, "oracle"), os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import p4_phylogenetics_b200 as P
    import pf_port
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    tree = P.synth.build_config(None, 1, nTax=9, nPatterns=400)
    # this rank's contiguous share of the alignment columns, by the engine's rule
    aln = tree.data.alignments[0]
    comp = pf_port.compress(aln.sequences, aln.symbols, aln.equates)
    nPat = comp["nPatterns"]
    lo, hi = P.pf.shardRangeFor(nPat, rank, world)
    keep = np.isin(comp["sequencePositionPatternIndex"], np.arange(lo, hi))
    sub = [bytes(np.frombuffer(s, dtype=np.uint8)[keep]) for s in aln.sequences]
    tree.data = P.host.Data(None, [P.host.Alignment(None, sub, aln.symbols, aln.equates)])
    part = torch.tensor([pf_port.tree_loglike(tree)], dtype=torch.float64)
    dist.all_reduce(part, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(out, np.array([part.item()]))
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_equals_unsharded(pkg, tmp_path):
    import torch.multiprocessing as mp
    import pf_port
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "lnl.npy")
    mp.spawn(_worker, args=(2, _free_port(), root, out), nprocs=2, join=True)
    got = float(np.load(out)[0])
    whole = pkg.synth.build_config(None, 1, nTax=9, nPatterns=400)
    want = pf_port.tree_loglike(whole)
    assert abs(got - want) <= 1e-11 * abs(want)


def _newt_worker(rank, world, port, root, out):
    import sys
    for p in (root, os.path.join(root, "oracle"), os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import p4_phylogenetics_b200 as P
    import pf_port
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    tree = P.synth.build_config(None, 1, nTax=8, nPatterns=300)
    aln = tree.data.alignments[0]
    comp = pf_port.compress(aln.sequences, aln.symbols, aln.equates)
    lo, hi = P.pf.shardRangeFor(comp["nPatterns"], rank, world)
    keep = np.isin(comp["sequencePositionPatternIndex"], np.arange(lo, hi))
    sub = [bytes(np.frombuffer(s, dtype=np.uint8)[keep]) for s in aln.sequences]
    tree.data = P.host.Data(None, [P.host.Alignment(None, sub, aln.symbols, aln.equates)])
    # the engine's collective for one derivative evaluation: 3 doubles per part (sum n log l, sum n l'/l, sum n (l''l - l'^2)/l^2)
    nodes = [n for n in tree.iterNodesNoRoot()]
    mine = torch.tensor([v for n in nodes for v in pf_port.branch_derivs_of(tree, n)], dtype=torch.float64)
    dist.all_reduce(mine, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(out, mine.numpy())
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_of_the_newton_sums_equals_unsharded(pkg, tmp_path):
    """The Newton-Raphson step's collective: the three sums of a derivative evaluation are sums over
    patterns, so per-shard sums all-reduced over the ranks are the unsharded sums -- every rank then takes the same Newton step."""
    import torch.multiprocessing as mp
    import pf_port
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "newt.npy")
    mp.spawn(_newt_worker, args=(2, _free_port(), root, out), nprocs=2, join=True)
    got = np.load(out)
    whole = pkg.synth.build_config(None, 1, nTax=8, nPatterns=300)
    want = np.array([v for n in whole.iterNodesNoRoot() for v in pf_port.branch_derivs_of(whole, n)])
    assert got.shape == want.shape
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0)) <= 1e-10
