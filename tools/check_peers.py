#!/usr/bin/env python
"""2+ GPU check (torchrun): the in-kernel peer-store gather must equal the NCCL
all-gather of the single-GPU results, byte for byte, on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    from igm_b200 import synthetic
    from igm_b200.dist import PeerGather, peer_gather_available
    from igm_b200.engine import ActdistEngine
    pop = synthetic.make_population(2_000_000, 300, seed=5, genome_scale=0.05)
    rng = np.random.default_rng(100 + rank)
    n = 20000
    ii = rng.integers(0, pop.n_hap - 1, n).astype(np.int32)
    jj = (ii + 1 + rng.integers(0, 40, n)).clip(max=pop.n_hap - 1).astype(np.int32)
    nc, ch = pop.copy_index.ncopies(), pop.chrom_hap()
    bad = ((ch[ii] == ch[jj]) & (nc[ii] != nc[jj])) | (ii == jj)
    jj[bad] = ii[bad]                      # i == j -> empty result, still a valid slot
    pw = rng.uniform(0.01, 1, n)
    d_i, d_j = torch.from_numpy(ii).to(dev), torch.from_numpy(jj).to(dev)
    d_pw, d_pl = torch.from_numpy(pw).to(dev), torch.zeros(n, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    os.environ["IGMK_PEER_GATHER"] = "1"
    assert peer_gather_available(world)
    with ActdistEngine(pop, lr) as eng:
        mine = torch.zeros((n, 32), dtype=torch.uint8, device=dev)
        eng.actdist_device(d_i, d_j, d_pw, d_pl, mine, n, 2.0, 1, "LB", stream=stream)
        ref = torch.zeros((world, n, 32), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(ref.view(world * n, 32), mine)
        pg = PeerGather(n, rank, world, dev)
        ok = True
        for rep in range(5):
            got = pg.step(eng, d_i, d_j, d_pw, d_pl, n, 2.0, 1, "LB", stream)
            torch.cuda.synchronize()
            ok = ok and bool(torch.equal(got, ref))
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("peer gather == nccl all-gather:", bool(t.item()))
    dist.barrier()
    dist.destroy_process_group()
    if not t.item():
        sys.exit(1)


if __name__ == "__main__":
    main()
