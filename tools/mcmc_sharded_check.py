"""tools/mcmc_sharded_check.py -- the MCMC protocol on pattern shards: every rank runs the same chains (same seed, same
proposals) on its shard of the patterns; the part log-likelihoods are all-reduced inside the engine, so every rank
takes the same accept / reject / swap decisions.  Rank 0 also runs the reference's own Pf engine on the whole
alignment and compares the two traces.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/mcmc_sharded_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import p4_phylogenetics_b200 as P  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    pf = P.pf
    torch.cuda.set_device(local)
    pf.setDevice(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = [pf.commGetUniqueId() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    pf.commInitRank(uid[0], rank, world)

    def run(engine, gens, chains):
        tree = P.synth.build_config(engine, 5, nTax=16, nPatterns=3000)
        m = P.mcmc.Mcmc(tree, nChains=chains, seed=17)
        return m.run(gens), m

    trace, m = run(pf, 80, 4)
    lo, hi = pf.treeShardRange(m.chains[0].curTree.cTree, 0)
    mine = torch.tensor([x for _, row in trace for x in row], dtype=torch.float64, device="cuda")
    ref0 = mine.clone()
    dist.broadcast(ref0, src=0)
    same_everywhere = bool(torch.equal(mine, ref0))
    ok = torch.tensor([1 if same_everywhere else 0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        import ref_loader
        want, _ = run(ref_loader.load_ref_pf(), 80, 4)
        worst = max(abs(a - b) / abs(b) for (_, ra), (_, rb) in zip(trace, want) for a, b in zip(ra, rb))
        print("sharded MCMC: %d ranks, shard of rank 0 = [%d, %d); traces identical on all ranks: %s; worst rel. diff vs the "
              "reference engine over %d generations x 4 chains: %.3e" % (world, lo, hi, bool(ok.item()), len(trace), worst))
        assert ok.item() == 1 and worst <= 1e-9
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
