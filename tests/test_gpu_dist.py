"""2-GPU tests (skipped with fewer devices): one process per GPU, the NCCL communicator owned by the ctx (kprn_comm_*, the
same entry points a LuaJIT host binds) -- segment-range sharding of one CSR batch (kprn_b200/dist.py), no data-path
collective for scoring; path sharding with the pooled logits combined as log-sum-exp partials ("path_sharded"); training: all-reduce of the small tensors' gradients + all-gather of the entity-row gradients
(kprn_train_step), and the reference-style dense flow (kprn_allreduce_grads)."""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, id_q, out):
    sys.path.insert(0, ROOT)
    from kprn_b200 import dist, engine, synth
    from tests.util import engine_for, make_case
    torch.cuda.set_device(rank)
    if rank == 0:
        uid, uid2 = engine.Engine.comm_unique_id(), engine.Engine.comm_unique_id()
        for _ in range(world - 1):
            id_q.put((uid, uid2))
    else:
        uid, uid2 = id_q.get(timeout=120)
    V_e = 3000
    o, params, _, _, _ = make_case(N=64, seed=51, d_t=8, d_e=64, d_r=8, H=64, V_e=V_e)
    c = o.cfg
    pads = (c.V_t, c.V_e, c.V_r)
    batches = []
    for b in range(2):
        ids = synth.make_paths(2000, T=c.T, V_e=V_e, seed=510 + b, length_mode="stress")
        seg = synth.make_segments(2000, seed=510 + b)
        batches.append((ids, seg, synth.make_labels(len(seg) - 1, 510 + b)))
    e = engine_for(o, precision=1, device=rank)
    e.comm_init(rank, world, uid)
    if rank == 0:
        e.set_params(params)
    e.broadcast_params(0)
    res = {}
    # ---- sharded scoring: no collective on the data path ----
    ids, seg, labels = batches[0]
    S = len(seg) - 1
    l_ids, l_seg, l_lab, rng = dist.shard(ids, seg, labels, rank, world)
    scores_local = e.forward(l_ids, l_seg)
    full = np.zeros(S)
    full[rng[0]:rng[1]] = scores_local
    # (the test gathers them through the library's own small host all-reduce, 64 values at a time)
    for a in range(0, S, 64):
        full[a:a + 64] = e.allreduce_host(full[a:a + 64])
    ref = o.forward(params, ids, seg)
    res["score_err"] = float(np.max(np.abs(full - ref["scores"]) / ref["scores"]))
    # ---- reference-style dense flow: backward on the shard, one all-reduce of the flat gradient ----
    e.zero_grad()
    e.forward(l_ids, l_seg, train=True)
    e.loss_bce(l_lab)
    e.backward(scale=len(l_lab) / S)
    e.allreduce_grads()
    g = e.get_grads()
    refc = o.forward(params, ids, seg, cache=True)
    _, dsc = o.bce(refc["scores"], labels.astype(np.float64))
    g_ref = o.backward(params, ids, seg, refc, dsc)
    rms = np.sqrt(np.mean(g_ref ** 2))
    res["grad_err"] = float(np.max(np.abs(g - g_ref) / np.maximum(np.abs(g_ref), rms)))
    # ---- sparse training steps with the row exchange: 3 steps over 2 batches ----
    for t in range(3):
        ids, seg, labels = batches[t % 2]
        S = len(seg) - 1
        l_ids, l_seg, l_lab, rng = dist.shard(ids, seg, labels, rank, world)
        e.train_step(l_ids, l_seg, l_lab, pad_rows=pads, grad_scale=len(l_lab) / S)
    res["info"] = e.train_info()
    # ---- path-sharded pooling: the paths of every segment spread over the ranks, pooled logits combined as LSE partials ----
    ids, seg, labels = batches[0]
    f = engine_for(o, precision=1, device=rank)
    f.comm_init(rank, world, uid2)
    f.set_params(params)
    f.set_option("path_sharded", 1)
    p_ids, p_seg = dist.shard_paths(ids, seg, rank, world)
    sc = f.forward(p_ids, p_seg)                           # scores-only path (class column)
    res["ps_score_err"] = float(np.max(np.abs(sc - ref["scores"]) / ref["scores"]))
    sc2, lg2 = f.forward(p_ids, p_seg, want_logits=True)    # all 46 pooled columns
    res["ps_logit_err"] = float(np.max(np.abs(lg2 - ref["seg_logits"]) / np.maximum(np.abs(ref["seg_logits"]), np.sqrt(np.mean(ref["seg_logits"] ** 2)))))
    # training-mode forward + BCE + backward on the shard: the gradient of the WHOLE batch after the all-reduce
    f.zero_grad()
    f.forward(p_ids, p_seg, train=True)
    f.loss_bce(labels)
    f.backward()
    f.allreduce_grads()
    g2 = f.get_grads()
    res["ps_grad_err"] = float(np.max(np.abs(g2 - g_ref) / np.maximum(np.abs(g_ref), rms)))
    f.allreduce_host(np.zeros(1))
    f.comm_destroy()
    f.close()
    res["comm"] = e.comm_info()
    p = e.get_params()
    res["hash"] = hashlib.sha1(p.tobytes()).hexdigest()
    if rank == 0:
        # single-GPU run of the same steps on the whole batches (dense flow) as the reference point
        d = engine_for(o, precision=1, device=rank)
        d.set_params(params)
        for t in range(3):
            ids, seg, labels = batches[t % 2]
            d.train_step(ids, seg, labels, pad_rows=pads, dense=True)
        q = d.get_params()
        diff = np.abs(p.astype(np.float64) - q.astype(np.float64))
        res["param_q999"] = float(np.quantile(diff, 0.999)); res["param_max"] = float(diff.max())
        d.close()
    out.put((rank, res))
    e.allreduce_host(np.zeros(1))        # barrier before tearing the communicator down
    e.comm_destroy()
    e.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_scoring_dense_allreduce_and_sparse_row_exchange():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out, id_q = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, id_q, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=400) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    r0, r1 = got[0], got[1]
    assert r0["score_err"] < 1e-4 and r0["grad_err"] < 1e-4, r0
    assert r0["hash"] == r1["hash"], "replicas diverged"
    assert r0["param_q999"] <= 2e-6 and r0["param_max"] <= 2e-4, r0
    assert r0["info"]["rows_union"] >= r0["info"]["rows_local"] > 0 and r0["comm"]["bytes"] > 0, r0
    for r in (r0, r1):
        assert r["ps_score_err"] < 1e-4 and r["ps_logit_err"] < 1e-4 and r["ps_grad_err"] < 1e-4, r
