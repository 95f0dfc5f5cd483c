"""GPU parity tests proper: the CUDA path, called through the C ABI, against the fp64 oracle on the same inputs.
Tolerance: 1e-4 relative (tests/util.py:assert_rel), the figure BASELINE.json's north_star states for fp32."""
import numpy as np
import pytest

from tests.util import assert_rel, engine_for, make_case

pytestmark = pytest.mark.gpu

FWD_CASES = [
    dict(N=200, seed=1),
    dict(N=333, seed=2, d_t=5, d_e=10, d_r=5, H=25),                     # reference ratio, odd sizes
    dict(N=1000, seed=3, T=3, d_t=32, d_e=32, d_r=32, H=32, V_e=5000),   # BASELINE configs[0]
    dict(N=64, seed=4, seg_mode="dense", fixed_p=4),
    dict(N=257, seed=5, cell=1),
    dict(N=257, seed=6, pool=1),
    dict(N=257, seed=7, pool=2, topk=2),
    dict(N=257, seed=8, head_layers=2, L=1),
    dict(N=257, seed=9, inv_gamma=0.5),
    dict(N=1, seed=10),
]


@pytest.mark.parametrize("kw", FWD_CASES)
def test_forward_matches_oracle_fp32(kw):
    o, params, ids, seg, labels = make_case(**kw)
    ref = o.forward(params, ids, seg)
    e = engine_for(o, precision=0)
    e.set_params(params)
    scores, logits = e.forward(ids, seg, want_logits=True)
    assert_rel(e.path_logits(ids.shape[0]), ref["path_logits"], what="path_logits")
    assert_rel(logits, ref["seg_logits"], what="seg_logits")
    assert_rel(scores, ref["scores"], what="scores")
    e.close()


@pytest.mark.parametrize("kw", FWD_CASES[:9])
def test_bce_backward_matches_oracle(kw):
    o, params, ids, seg, labels = make_case(**kw)
    ref = o.forward(params, ids, seg, cache=True)
    loss_ref, dsc = o.bce(ref["scores"], labels)
    g_ref = o.backward(params, ids, seg, ref, dsc)
    e = engine_for(o, precision=0)
    e.set_params(params)
    e.zero_grad()
    e.forward(ids, seg, train=True)
    loss = e.loss_bce(labels)
    assert abs(loss - loss_ref) <= 1e-4 * abs(loss_ref)
    e.backward()
    g = e.get_grads()
    for name, shp in o.shapes().items():
        assert_rel(o.view(g, name), o.view(g_ref, name), what="grad " + name)
    # accGradParameters accumulates: a second backward with scale doubles + scales
    e.backward(scale=0.5)
    assert_rel(e.get_grads(), 1.5 * g_ref, what="accumulated grads")
    e.close()


def test_bpr_loss_and_grad_match_oracle():
    o, params, ids, seg, labels = make_case(N=300, seed=11)
    ref = o.forward(params, ids, seg, cache=True)
    S = len(seg) - 1
    rng = np.random.default_rng(0)
    pos = rng.integers(0, S, 64).astype(np.int32)
    neg = rng.integers(0, S, 64).astype(np.int32)
    loss_ref, dp, dn = o.bpr(ref["scores"][pos], ref["scores"][neg])
    dsc = np.zeros(S)
    np.add.at(dsc, pos, dp)
    np.add.at(dsc, neg, dn)
    g_ref = o.backward(params, ids, seg, ref, dsc)
    e = engine_for(o)
    e.set_params(params)
    e.zero_grad()
    e.forward(ids, seg, train=True)
    loss = e.loss_bpr(pos, neg)
    assert abs(loss - loss_ref) <= 1e-4 * abs(loss_ref)
    e.backward()
    assert_rel(e.get_grads(), g_ref, what="bpr grads")
    e.close()


def test_external_grad_scores_and_f64_ids_layout():
    o, params, ids, seg, labels = make_case(N=64, seed=12, seg_mode="dense", fixed_p=4)
    ref = o.forward(params, ids, seg, cache=True)
    e = engine_for(o)
    e.set_params(params)
    data = ids.reshape(16, 4, o.cfg.T, 3).astype(np.float64)     # Batcher tensor: [B,P,T,3] DoubleTensor
    scores = e.forward_f64ids(data, train=True)
    assert_rel(scores, ref["scores"], what="scores via f64 ids")
    rng = np.random.default_rng(1)
    dsc = rng.normal(size=16)
    g_ref = o.backward(params, ids, seg, ref, dsc)
    e.zero_grad()
    e.backward(grad_scores=dsc)
    assert_rel(e.get_grads(), g_ref, what="grads from external dL/dscore")
    e.close()


def test_train_batch_matches_oracle_trainBatch():
    """MyOptimizer:trainBatch (MyOptimizer.lua:177-221): pad zero -> fwd -> BCE -> bwd -> clip+l2 -> Adam -> pad zero."""
    o, params, ids, seg, labels = make_case(N=200, seed=13)
    c = o.cfg
    pads = (c.V_t, c.V_e, c.V_r)
    e = engine_for(o)
    p0 = params.copy()
    p0[:] += 0.0
    e.set_params(p0)
    P = p0.copy()
    m = np.zeros_like(P); v = np.zeros_like(P)
    for step in (1, 2, 3):
        loss = e.train_batch_bce(ids, seg, labels, pad_rows=pads, regularize=True, use_clip=True, clip_norm=0.05, l2=1e-3)
        for tbl, row in enumerate(pads):
            o.zero_row(P, tbl, row)
        fwd = o.forward(P, ids, seg, cache=True)
        loss_ref, dsc = o.bce(fwd["scores"], labels.astype(np.float64))
        g = o.backward(P, ids, seg, fwd, dsc)
        o.clip_l2(g, P, True, 0.05, 1e-3)
        o.adam(P, g, m, v, step)
        for tbl, row in enumerate(pads):
            o.zero_row(P, tbl, row)
        assert abs(loss - loss_ref) <= 2e-4 * abs(loss_ref)
    # Adam normalises by sqrt(v): early steps move every touched parameter by ~lr, so compare the update itself
    upd, upd_ref = e.get_params() - p0, P - p0
    assert_rel(upd, upd_ref, rtol=2e-3, what="parameter update after 3 Adam steps")
    e.close()


@pytest.mark.parametrize("precision", [0, 1])
def test_errors_are_loud(precision):
    from kprn_b200.engine import KprnError
    o, params, ids, seg, labels = make_case(N=32, seed=14, d_t=8, d_e=64, d_r=8, H=64)
    e = engine_for(o, precision=precision)
    e.set_params(params)
    for col, v in ((1, o.cfg.V_e + 1), (1, 0), (0, o.cfg.V_t + 1), (2, -3)):     # nn.LookupTable raises on any of these
        bad = ids.copy(); bad[3, 2, col] = v
        with pytest.raises(KprnError):
            e.forward(bad, seg)
    with pytest.raises(KprnError):
        e.backward()                      # no training forward yet
    badseg = seg.copy(); badseg[-1] -= 1
    with pytest.raises(KprnError):
        e.forward(ids, badseg)
    e.forward(ids, seg)                   # still usable afterwards
    e.close()


@pytest.mark.parametrize("precision", [0, 1])
def test_scores_only_fast_path_equals_full(precision):
    """forward without logits materialises only column classId (nn.Select(2,classId)); scores must be identical."""
    from kprn_b200.engine import KprnError
    o, params, ids, seg, labels = make_case(N=700, seed=21, d_t=8, d_e=64, d_r=8, H=64, V_e=500)
    ref = o.forward(params, ids, seg)
    e = engine_for(o, precision=precision)
    e.set_params(params)
    s_fast = e.forward(ids, seg)
    with pytest.raises(KprnError):
        e.path_logits(ids.shape[0])
    s_full, _ = e.forward(ids, seg, want_logits=True)
    assert_rel(s_fast, ref["scores"], what="scores-only")
    assert_rel(s_fast, s_full, rtol=1e-6, what="scores-only vs full")
    e.close()


def test_pinned_host_buffers_give_identical_results():
    from kprn_b200 import engine
    o, params, ids, seg, labels = make_case(N=300, seed=22)
    e = engine_for(o)
    e.set_params(params)
    s_pageable = e.forward(ids, seg)
    pi = engine.pinned_empty(ids.shape, np.int32); pi[...] = ids
    ps = engine.pinned_empty(seg.shape, np.int32); ps[...] = seg
    s_pinned = e.forward(pi, ps)
    np.testing.assert_array_equal(s_pageable, s_pinned)
    e.close()


@pytest.mark.parametrize("precision", [0, 1])
@pytest.mark.parametrize("lens", [[1, 31, 32, 33, 1000, 7], [1, 2, 3, 1, 4, 2, 5, 1, 1, 6, 2, 3, 9, 1, 2] * 9, [0, 3, 0, 0, 5, 1, 0]])
def test_long_short_and_empty_segments_warp_shuffle_pool(precision, lens):
    """Scores-only LSE pool = shuffle butterfly over 32 lanes (long segments) or 4 lanes (batches of short segments) per
    (user,item) pair; with precision 1 the per-path head value comes from the four partial dot products the fused
    tensor-core forward leaves.  Empty segments (nn.MapReduce never produces one, the CSR form allows it) score sigmoid(0)."""
    n = int(np.sum(lens))
    o, params, ids, _, _ = make_case(N=n, seed=23, d_t=8, d_e=64, d_r=8, H=64, V_e=400)
    seg = np.cumsum([0] + list(lens)).astype(np.int32)
    ref = o.forward(params, ids, seg)
    e = engine_for(o, precision=precision)
    e.set_params(params)
    got = e.forward(ids, seg)
    nz = np.array(lens) > 0
    assert_rel(got[nz], ref["scores"][nz], what="scores-only pool")
    assert np.all(got[~nz] == 0.5)
    s_full, lg = e.forward(ids, seg, want_logits=True)
    assert_rel(lg[nz], ref["seg_logits"][nz], what="full pool")
    assert_rel(e.forward(ids, seg)[nz], ref["scores"][nz], what="scores-only again after a full forward")
    e.close()


def test_fallback_kernels_still_match():
    """KPRN_TC_RECUR_PAIRS=0 (1-CTA recurrence) and KPRN_TC_DEDUP=0 (projection of every step instead of every distinct step
    input) are kept as fallbacks: same results."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = "\n".join([
        "import sys; sys.path.insert(0, %r)" % root,
        "from tests.util import make_case, engine_for, assert_rel",
        "o, params, ids, seg, labels = make_case(N=700, seed=24, d_t=8, d_e=128, d_r=8, H=128, V_e=900)",
        "ref = o.forward(params, ids, seg)",
        "e = engine_for(o, precision=1); e.set_params(params)",
        "assert_rel(e.forward(ids, seg), ref['scores'], what='scores'); print('ok')",
    ])
    compile(code, "<fallback>", "exec")
    for env in ({"KPRN_TC_RECUR_PAIRS": "0"}, {"KPRN_TC_DEDUP": "0"}, {"KPRN_TC_DEDUP": "0", "KPRN_TC_RECUR_PAIRS": "0"}):
        out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and "ok" in out.stdout, (env, out.stderr[-800:])


def test_dedup_is_bit_identical_to_projecting_every_step(monkeypatch):
    """Projecting each distinct (type, entity, relation) step input once changes where a row of XW lives, not its value."""
    o, params, ids, seg, labels = make_case(N=3000, seed=31, d_t=16, d_e=128, d_r=16, H=128, V_e=300, length_mode="realistic")
    e1 = engine_for(o, precision=1)
    e1.set_params(params)
    a = e1.forward(ids, seg)
    e1.close()
    monkeypatch.setenv("KPRN_TC_DEDUP", "0")
    e2 = engine_for(o, precision=1)
    e2.set_params(params)
    b = e2.forward(ids, seg)
    e2.close()
    np.testing.assert_array_equal(a, b)
    assert_rel(a, o.forward(params, ids, seg)["scores"], what="scores")
