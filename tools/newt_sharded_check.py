"""tools/newt_sharded_check.py -- the Newton-Raphson branch-length step on pattern shards: every rank holds its shard
of the patterns (CLs and cl2 arrays of ALL nodes for that range); the three sums of a derivative evaluation are
all-reduced inside the engine (one ncclAllReduce of 3*nParts doubles per evaluation), so every rank takes the same
Newton steps.  Rank 0 also runs the reference's own p4_newtAround on the whole alignment and compares.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/newt_sharded_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import p4_phylogenetics_b200 as P  # noqa: E402


def build(engine, cfg, nTax, nPatterns):
    tree = P.synth.build_config(engine, cfg, nTax=nTax, nPatterns=nPatterns)
    rng = np.random.default_rng(5)
    for n in tree.iterNodesNoRoot():
        n.br.len = float(min(max(n.br.len * np.exp(rng.normal(0.0, 0.6)), 1e-4), 2.0))
    return tree


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    pf = P.pf
    torch.cuda.set_device(local)
    pf.setDevice(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = [pf.commGetUniqueId() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    pf.commInitRank(uid[0], rank, world)
    worst = 0.0
    for cfg, nTax, nPat in ((1, 14, 4001), (3, 9, 1203)):       # DNA with pInvar; protein; ragged shard sizes
        tree = build(pf, cfg, nTax, nPat)
        start = tree.calcLogLike()
        pf.p4_newtSetup(tree.cTree)
        got = pf.newtAround(tree.cTree, 1.0e-5, 1.0e-7)
        lens = pf.p4_getBrLens(tree.cTree)
        lo, hi = pf.treeShardRange(tree.cTree, 0)
        mine = torch.tensor([got] + [lens[n.nodeNum] for n in tree.iterNodesNoRoot()], dtype=torch.float64, device="cuda")
        ref0 = mine.clone()
        dist.broadcast(ref0, src=0)
        ok = torch.tensor([1 if torch.equal(mine, ref0) else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if rank == 0:
            import ref_loader
            import ref_peek
            rpf = ref_loader.load_ref_pf()
            twin = build(rpf, cfg, nTax, nPat)
            twin.calcLogLike()
            rpf.p4_newtSetup(twin.cTree)
            ref_peek.newt_lib().p4_newtAround(twin.cTree, 1.0e-5, 1.0e-7)
            want = rpf.p4_treeLogLike(twin.cTree, 0)
            dl = abs(got - want) / abs(want)
            db = max(abs(lens[a.nodeNum] - ref_peek.node_brlen(b.cNode)) / max(ref_peek.node_brlen(b.cNode), 1e-3)
                     for a, b in zip(tree.iterNodesNoRoot(), twin.iterNodesNoRoot()))
            worst = max(worst, dl)
            print("sharded Newton-Raphson cfg %d: %d ranks, shard of rank 0 = [%d, %d) of %d; lnL %.6f -> %.6f; identical on all ranks: %s; "
                  "vs the reference engine on the whole alignment: lnL rel. diff %.2e, worst branch-length rel. diff %.2e"
                  % (cfg, world, lo, hi, nPat, start, got, bool(ok.item()), dl, db))
            assert ok.item() == 1 and dl <= 1e-9 and db <= 1e-6
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
