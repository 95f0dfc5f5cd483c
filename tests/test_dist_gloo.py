"""world_size-2 gloo test (CPU) of the N>1 host logic: shard by segment range, score shards independently (no
data-path collective), gather scores; all-reduce of the shard gradients reproduces the full-batch gradient; path sharding
with the pooled logits combined as log-sum-exp partials (the host mirror of pool_allreduce_lse) reproduces the full pool."""
import os
import socket
import sys

import numpy as np
import torch.distributed as tdist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["OMP_NUM_THREADS"] = "2"
    from kprn_b200 import dist
    from tests.util import make_case
    tdist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    o, params, ids, seg, labels = make_case(N=400, seed=31)
    S = len(seg) - 1
    l_ids, l_seg, l_lab, rng = dist.shard(ids, seg, labels, rank, world)
    fwd = o.forward(params, l_ids, l_seg, cache=True)
    scores = dist.gather_scores(fwd["scores"], rng, S)
    # BCE is a mean over ALL S segments: the local gradient uses dL/dscore scaled by 1/S, i.e. S_local/S of the local mean
    _, dsc = o.bce(fwd["scores"], l_lab.astype(np.float64))
    dsc *= len(l_lab) / S
    g = dist.allreduce_host(o.backward(params, l_ids, l_seg, fwd, dsc))
    # path sharding: the paths of every segment spread over the ranks, pooled logits combined as log-sum-exp partials
    p_ids, p_seg = dist.shard_paths(ids, seg, rank, world)
    assert len(p_seg) == len(seg) and p_seg[-1] == len(p_ids)
    pf = o.forward(params, p_ids, p_seg)
    pooled = dist.allreduce_lse(pf["seg_logits"], p_seg, o.cfg.inv_gamma if o.cfg.inv_gamma else 1.0)
    if rank == 0:
        full = o.forward(params, ids, seg, cache=True)
        out.put(("lse", float(np.abs(pooled - full["seg_logits"]).max())))
        _, dfull = o.bce(full["scores"], labels.astype(np.float64))
        gfull = o.backward(params, ids, seg, full, dfull)
        out.put((float(np.abs(scores - full["scores"]).max()), float(np.abs(g - gfull).max() / np.abs(gfull).max())))
    tdist.barrier()
    tdist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = [out.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lse = [g for g in got if g[0] == "lse"][0]
    res = [g for g in got if g[0] != "lse"][0]
    assert lse[1] < 1e-10      # path-sharded pooling == pooling of the whole batch
    assert res[0] < 1e-6       # fp32 transport of the gathered scores
    assert res[1] < 1e-12
