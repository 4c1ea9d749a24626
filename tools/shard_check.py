"""tools/shard_check.py -- multi-GPU correctness of the sharded evaluation (run under torchrun, one rank per GPU).

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/shard_check.py [--patterns 200000]

Every rank holds its pattern shard; the shard sums are combined inside the kernel through NVLink peer mailboxes
(include/p4b200.h p4b_peerReduceState) or, with P4B_PEER_REDUCE=0, by an NCCL all-reduce.  Checked on every rank:
  * p4_treeLogLike (4-state whole-tree kernel, fused fold + exchange), repeated -- identical bits on every rank and call;
  * a queued dirty path + p4_partLogLike; three trees in one batched launch (pf.treesPartLogLike);
  * a 20-state part (like_kernel + like_final_kernel exchange) and a 6-state part (per-node kernels);
and on rank 0, after the communicator is gone, the same trees evaluated unsharded: the sharded values must agree to 1e-12.
Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import p4_phylogenetics_b200 as P  # noqa: E402


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--patterns", type=int, default=200000)
    ap.add_argument("--taxa", type=int, default=60)
    a = ap.parse_args()
    pf = P.pf
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    pf.setDevice(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = [pf.commGetUniqueId() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    pf.commInitRank(uid[0], rank, world)
    pf.setMemoize(0)
    out = {"world": world}
    dna = P.synth.build_config(pf, 2, nTax=a.taxa, nPatterns=a.patterns)
    vals = [dna.calcLogLike() for _ in range(4)]
    out["peer_reduce_state"] = pf.peerReduceState()
    assert vals[0] == vals[1] == vals[2] == vals[3], vals
    t0 = time.perf_counter()
    for _ in range(50):
        pf.p4_treeLogLike(dna.cTree, 0)
    out["ms_per_eval_wall"] = (time.perf_counter() - t0) * 1e3 / 50
    n = list(dna.iterNodesNoRoot())[7]
    n.br.len *= 1.3
    n.br.lenChanged = True
    dirty = dna.recalcAfterBranchChange()
    full = dna.calcLogLike()
    assert rel(dirty, full) <= 1e-13, (dirty, full)
    twins = [dna] + [P.host.clone_tree(dna, pf, data=dna.data) for _ in range(2)]
    for k, t in enumerate(twins[1:]):
        for x in t.iterNodesNoRoot():
            x.br.len *= 1.0 + 0.1 * (k + 1)
    single = [t.calcLogLike() for t in twins]
    for t in twins:
        for x in t.iterInternalsPostOrder():
            pf.p4_setConditionalLikelihoodsOfInternalNodePart(x.cNode, 0)
    batched = pf.treesPartLogLike([t.cTree for t in twins], 0)
    for s, b in zip(single, batched):
        assert rel(b, s) <= 1e-13, (s, b)
    prot = P.synth.build_config(pf, 3, nTax=20, nPatterns=5000)
    lnProt = prot.calcLogLike()
    day = None
    sharded = {"dna": full, "dna_batched": batched, "protein": lnProt}
    # every rank must hold the same bits
    t = torch.tensor([full, lnProt] + list(batched), dtype=torch.float64, device="cuda")
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    for g in gathered[1:]:
        assert torch.equal(g, gathered[0]), "ranks disagree"
    dist.barrier()
    pf.commDestroy()          # back to one process = the whole alignment (the trees lay themselves out again)
    if rank == 0:
        one = {"dna": dna.calcLogLike(), "dna_batched": [t.calcLogLike() for t in twins], "protein": prot.calcLogLike()}
        out["rel_dna"] = rel(sharded["dna"], one["dna"])
        out["rel_protein"] = rel(sharded["protein"], one["protein"])
        out["rel_batched"] = max(rel(x, y) for x, y in zip(sharded["dna_batched"], one["dna_batched"]))
        out["lnL_dna"] = sharded["dna"]
        out["ok"] = bool(max(out["rel_dna"], out["rel_protein"], out["rel_batched"]) <= 1e-12)
        print("SHARDCHECK" + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
