"""2-GPU tests (skipped with fewer devices): one process per GPU over NCCL, segment-range sharding of one CSR batch,
no data-path collective for scoring, one all-reduce of the flat gradient buffer for training (kprn_b200/dist.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as tdist
    from kprn_b200 import dist
    from tests.util import engine_for, make_case
    torch.cuda.set_device(rank)
    tdist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                             device_id=torch.device("cuda", rank))
    o, params, ids, seg, labels = make_case(N=2000, seed=51, d_t=8, d_e=64, d_r=8, H=64, V_e=900)
    S = len(seg) - 1
    l_ids, l_seg, l_lab, rng = dist.shard(ids, seg, labels, rank, world)
    e = engine_for(o, precision=1, device=rank)
    e.set_params(params)
    e.zero_grad()
    scores_local = e.forward(l_ids, l_seg, train=True)
    scores = dist.gather_scores(scores_local, rng, S)
    # BCE is a mean over all S segments: scale the local dL/dscore by S_local / S via kprn_backward's scale
    e.loss_bce(l_lab)
    e.backward(scale=len(l_lab) / S)
    dist.allreduce_grads(e)
    g = e.get_grads()
    if rank == 0:
        ref = o.forward(params, ids, seg, cache=True)
        _, dsc = o.bce(ref["scores"], labels.astype(np.float64))
        g_ref = o.backward(params, ids, seg, ref, dsc)
        rms = np.sqrt(np.mean(g_ref ** 2))
        out.put((float(np.max(np.abs(scores - ref["scores"]) / ref["scores"])),
                 float(np.max(np.abs(g - g_ref) / np.maximum(np.abs(g_ref), rms)))))
    tdist.barrier()
    e.close()
    tdist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_scoring_and_gradient_allreduce():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res[0] < 1e-4 and res[1] < 1e-4, res
