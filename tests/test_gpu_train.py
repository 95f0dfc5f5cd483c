"""GPU tests of the training step (kprn_train_step: MyOptimizer:trainBatch, SP/model/optimizer/MyOptimizer.lua:177-221):
the row-sparse evaluation (compact entity gradient + lazily-replayed dense optim.adam) against the reference-literal dense
flow and against the fp64 oracle's trainBatch, BCE and BPR, with the entity rows of successive batches overlapping only
partly (so rows sit out several steps and are replayed)."""
import numpy as np
import pytest

from kprn_b200 import synth
from tests.util import assert_rel, engine_for, make_case

pytestmark = pytest.mark.gpu


def _batches(o, n_batches, N, seed, V_e):
    c = o.cfg
    out = []
    for b in range(n_batches):
        ids = synth.make_paths(N, T=c.T, V_e=V_e, seed=seed + 13 * b, length_mode="stress", V_t=c.V_t, V_r=c.V_r)
        seg = synth.make_segments(N, seed=seed + 13 * b)
        labels = synth.make_labels(len(seg) - 1, seed + 13 * b)
        out.append((ids, seg, labels))
    return out


def _oracle_train(o, params, batches, pads, steps, loss="bce", lr=1e-3):
    """MyOptimizer:trainBatch on the oracle: zero pads -> fwd -> loss -> bwd -> adam -> zero pads."""
    x = params.copy()
    m = np.zeros_like(x); v = np.zeros_like(x)
    losses = []
    for t in range(1, steps + 1):
        ids, seg, labels = batches[(t - 1) % len(batches)]
        for tbl, row in enumerate(pads):
            o.zero_row(x, tbl, row)
        fwd = o.forward(x, ids, seg, cache=True)
        if loss == "bce":
            lv, dsc = o.bce(fwd["scores"], labels.astype(np.float64))
        else:
            S = len(seg) - 1
            pos = np.arange(0, S - 1, 2); neg = pos + 1
            lv, dp, dn = o.bpr(fwd["scores"][pos], fwd["scores"][neg])
            dsc = np.zeros(S); np.add.at(dsc, pos, dp); np.add.at(dsc, neg, dn)
        losses.append(lv)
        g = o.backward(x, ids, seg, fwd, dsc)
        o.adam(x, g, m, v, t, lr=lr)
        for tbl, row in enumerate(pads):
            o.zero_row(x, tbl, row)
    return x, losses


@pytest.mark.parametrize("precision,dims", [(0, (64, 64)), (3, (64, 64)), (3, (128, 128))])
def test_sparse_step_equals_dense_flow_and_oracle(precision, dims):
    """5 steps over 3 rotating batches drawn from a 4000-row table (most rows sit out some steps): parameters after the
    last step -- read once, so every postponed update is replayed by the flush -- match the dense flow and the oracle."""
    V_e = 4000
    o, params, _, _, _ = make_case(N=64, seed=70, d_t=8, d_e=dims[0], d_r=8, H=dims[1], V_e=V_e)      # (128, 128): the fused tensor-core training kernels
    c = o.cfg
    pads = (c.V_t, c.V_e, c.V_r)
    batches = _batches(o, 3, 600, 700, V_e)
    steps = 5
    ref, ref_losses = _oracle_train(o, params, batches, pads, steps)
    outs = {}
    for mode in ("sparse", "dense"):
        e = engine_for(o, precision=precision)
        e.set_params(params)
        losses = []
        for t in range(steps):
            ids, seg, labels = batches[t % 3]
            losses.append(e.train_step(ids, seg, labels, pad_rows=pads, dense=(mode == "dense")))
        outs[mode] = e.get_params().astype(np.float64)
        for a, b in zip(losses, ref_losses):
            assert abs(a - b) <= 2e-4 * abs(b), (mode, losses, ref_losses)
        e.close()
    # sparse vs dense: same arithmetic per element; gradients differ only by float-atomic summation order
    d = np.abs(outs["sparse"] - outs["dense"])
    assert np.quantile(d, 0.999) <= 2e-6 and d.max() <= 2e-4, (np.quantile(d, 0.999), d.max())
    # against the oracle: Adam divides by sqrt(v) ~ |g|, so the bar is on the update (tests/test_gpu_parity.py)
    upd = ref - params
    scale = np.sqrt(np.mean(upd[upd != 0] ** 2))
    for mode in outs:
        err = np.abs(outs[mode] - ref)
        assert np.quantile(err, 0.999) <= 2e-3 * scale * 10 and np.mean(err) <= 2e-3 * scale, (mode, err.max(), scale)
    # untouched rows did not move
    touched = np.zeros(V_e, dtype=bool)
    for ids, _, _ in batches:
        touched[ids[:, :, 1].reshape(-1) - 1] = True
    E0 = o.view(params, "E_e"); E1 = o.view(outs["sparse"], "E_e")
    assert np.array_equal(E1[~touched], E0[~touched].astype(np.float32).astype(np.float64))


def test_lazy_rows_survive_scoring_between_steps_and_param_io():
    """A scoring forward / get_params between training steps flushes the postponed updates; training then continues."""
    V_e = 3000
    o, params, _, _, _ = make_case(N=64, seed=71, d_t=8, d_e=64, d_r=8, H=64, V_e=V_e)
    c = o.cfg
    pads = (c.V_t, c.V_e, c.V_r)
    batches = _batches(o, 3, 500, 710, V_e)
    ref, _ = _oracle_train(o, params, batches, pads, 4)
    e = engine_for(o, precision=1)          # tensor-core scoring forward between the steps: packed bf16 rows must follow
    d = engine_for(o, precision=1)
    e.set_params(params); d.set_params(params)
    for t in range(4):
        ids, seg, labels = batches[t % 3]
        e.train_step(ids, seg, labels, pad_rows=pads)
        d.train_step(ids, seg, labels, pad_rows=pads, dense=True)
        if t == 1:
            ids2, seg2, _ = batches[2]
            s_e, s_d = e.forward(ids2, seg2), d.forward(ids2, seg2)
            assert_rel(s_e, s_d, rtol=2e-5, what="scores between steps: sparse vs dense engine")
    ids2, seg2, _ = batches[0]
    assert_rel(e.forward(ids2, seg2), o.forward(ref, ids2, seg2)["scores"], rtol=5e-4, what="scores after training vs oracle-trained")
    assert_rel(e.forward(ids2, seg2), d.forward(ids2, seg2), rtol=2e-5, what="scores after training: sparse vs dense")
    e.close(); d.close()


@pytest.mark.parametrize("on_logits", [False, True])
def test_bpr_train_step_matches_oracle(on_logits):
    V_e = 2000
    o, params, _, _, _ = make_case(N=64, seed=72, d_t=8, d_e=64, d_r=8, H=64, V_e=V_e)
    c = o.cfg
    pads = (c.V_t, c.V_e, c.V_r)
    (ids, seg, labels), = _batches(o, 1, 700, 720, V_e)
    S = len(seg) - 1
    pos = np.arange(0, S - 1, 2, dtype=np.int32); neg = pos + 1
    x = params.copy()
    for tbl, row in enumerate(pads):
        o.zero_row(x, tbl, row)
    fwd = o.forward(x, ids, seg, cache=True)
    if on_logits:
        a, b = fwd["seg_logits"][pos, 0], fwd["seg_logits"][neg, 0]
        lv, dp, dn = o.bpr(a, b)
        dsl = np.zeros((S, c.L)); np.add.at(dsl[:, 0], pos, dp); np.add.at(dsl[:, 0], neg, dn)
        g_ref = o.backward_logits(x, ids, seg, fwd, dsl)
    else:
        lv, dp, dn = o.bpr(fwd["scores"][pos], fwd["scores"][neg])
        dsc = np.zeros(S); np.add.at(dsc, pos, dp); np.add.at(dsc, neg, dn)
        g_ref = o.backward(x, ids, seg, fwd, dsc)
    e = engine_for(o, precision=0)
    e.set_params(params)
    loss = e.train_step(ids, seg, None, pos, neg, pad_rows=pads, bpr_on_logits=on_logits)
    assert abs(loss - lv) <= 1e-4 * abs(lv)
    g = e.get_grads()                      # dense view of the sparse step's gradient
    for name in o.shapes():
        r = o.view(g_ref, name)
        if np.sqrt(np.mean(r * r)) < 1e-12:       # BPR on logits: every pair adds +s and -s to the class bias -> exactly zero
            assert np.abs(o.view(g, name)).max() < 1e-5, name
            continue
        assert_rel(o.view(g, name), r, what="grad " + name)
    # kprn_train_batch_bpr is the same step
    e2 = engine_for(o, precision=0)
    e2.set_params(params)
    loss2 = e2.train_batch_bpr(ids, seg, pos, neg, on_logits=on_logits, pad_rows=pads)
    assert abs(loss2 - loss) <= 1e-6 * abs(loss)
    e.close(); e2.close()


def test_clip_in_sparse_step_matches_dense_flow():
    V_e = 2000
    o, params, _, _, _ = make_case(N=64, seed=73, d_t=8, d_e=64, d_r=8, H=64, V_e=V_e)
    c = o.cfg
    pads = (c.V_t, c.V_e, c.V_r)
    batches = _batches(o, 2, 400, 730, V_e)
    outs = []
    for dense in (False, True):
        e = engine_for(o, precision=0)
        e.set_params(params)
        for t in range(3):
            ids, seg, labels = batches[t % 2]
            e.train_step(ids, seg, labels, pad_rows=pads, regularize=True, use_clip=True, clip_norm=1e-3, dense=dense)   # clip active
        outs.append(e.get_params().astype(np.float64))
        e.close()
    d = np.abs(outs[0] - outs[1])
    assert np.quantile(d, 0.999) <= 2e-6 and d.max() <= 2e-4, (np.quantile(d, 0.999), d.max())


def test_training_is_bit_reproducible_and_the_next_batch_hint_changes_no_result():
    """Two identical runs of 6 training steps leave bit-identical parameters (no float atomics on the tensor-core training path:
    sorted segment reduce with slots numbered by first occurrence, fixed-order split-K and column sums), and
    kprn_train_prefetch_device is a pure scheduling hint: with the right hint or a wrong one the bits are the same again (the
    hinted rows are replayed at the tail of the running step by the same kernel that would replay them at the head of the next)."""
    import torch
    V_e = 4000
    o, params, _, _, _ = make_case(N=64, seed=90, d_t=8, d_e=128, d_r=8, H=128, V_e=V_e)      # the fused tensor-core training kernels
    c = o.cfg
    pads = (c.V_t, c.V_e, c.V_r)
    batches = _batches(o, 3, 700, 900, V_e)
    dev = [(torch.from_numpy(i).cuda(), torch.from_numpy(s).cuda(), torch.from_numpy(l.astype(np.float32)).cuda(), len(s) - 1) for i, s, l in batches]
    out = []
    for mode in ("none", "none", "right", "wrong"):
        e = engine_for(o, precision=3)
        e.set_params(params)
        for t in range(6):
            ids_t, seg_t, lab_t, S = dev[t % 3]
            if mode == "right":
                e.train_prefetch_device(dev[(t + 1) % 3][0].data_ptr(), 700)
            elif mode == "wrong":
                e.train_prefetch_device(dev[(t + 2) % 3][0].data_ptr(), 700)
            e.train_step_device(ids_t.data_ptr(), seg_t.data_ptr(), 700, S, d_labels=lab_t.data_ptr(), pad_rows=pads)
        out.append(e.get_params().copy())
        e.close()
    assert np.array_equal(out[0], out[1]), "two identical runs differ"
    assert np.array_equal(out[0], out[2]) and np.array_equal(out[0], out[3]), "a hint changed the result"
