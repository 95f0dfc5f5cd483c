"""GPU parity tests for the tcgen05 tensor-core path (KPRN_PREC_TC_SPLIT = bf16x3 split, fp32 accumulate in TMEM)
against the fp64 oracle, 1e-4 relative (tests/util.py:assert_rel)."""
import numpy as np
import pytest

from tests.util import assert_rel, engine_for, make_case

pytestmark = pytest.mark.gpu

TC_CASES = [
    dict(N=128, seed=1, d_t=8, d_e=64, d_r=8, H=64, V_e=500),
    dict(N=300, seed=2, d_t=16, d_e=128, d_r=16, H=128, V_e=2000),                 # ragged last tile
    dict(N=1000, seed=3, d_t=32, d_e=256, d_r=32, H=256, V_e=5000),
    dict(N=333, seed=4, d_t=50, d_e=100, d_r=50, H=250, V_e=3000),                 # reference default dims (padded to 128/256)
    dict(N=1, seed=5, d_t=8, d_e=64, d_r=8, H=64, V_e=500),
    dict(N=1000, seed=6, T=3, d_t=32, d_e=32, d_r=32, H=32, V_e=5000),            # BASELINE configs[0] dims
    dict(N=40000, seed=7, d_t=8, d_e=64, d_r=8, H=64, V_e=50000),                  # > 148 tiles: persistent loop
    dict(N=500, seed=8, d_t=8, d_e=64, d_r=8, H=64, V_e=500, pool=1),
    dict(N=500, seed=9, d_t=8, d_e=64, d_r=8, H=64, V_e=500, head_layers=2, L=1),
    dict(N=260, seed=11, T=1, d_t=8, d_e=64, d_r=8, H=64, V_e=500),                 # single step: no recurrent MMA at all
    dict(N=260, seed=12, T=2, d_t=8, d_e=128, d_r=8, H=64, V_e=500),                # d_e > H
    dict(N=260, seed=13, d_t=8, d_e=64, d_r=8, H=192, V_e=500),                     # H padded 192 -> 256? no: 192 = 3*64, odd chunk pairs
    dict(N=700, seed=14, d_t=6, d_e=72, d_r=10, H=96, V_e=900, V_t=9, V_r=12),      # 108 (type,rel) combos -> two one-hot slices, ragged dims
]


# shapes outside the fused kernels (H or d_e > 256, Elman cell): the generic engine with its products on tcgen05 (tc_gemm)
GENERIC_TC_CASES = [
    dict(N=300, seed=15, d_t=8, d_e=512, d_r=8, H=512, V_e=800),                    # BASELINE configs[4], d = 512
    dict(N=300, seed=16, d_t=8, d_e=320, d_r=8, H=128, V_e=800),                    # only the entity width is too large
    dict(N=400, seed=17, d_t=16, d_e=128, d_r=16, H=128, V_e=800, cell=1),          # Elman ReLU + MaskZero
]


@pytest.mark.parametrize("kw", GENERIC_TC_CASES)
def test_generic_engine_on_tensor_cores_matches_oracle(kw):
    o, params, ids, seg, labels = make_case(**kw)
    ref = o.forward(params, ids, seg)
    e = engine_for(o, precision=1)
    e.set_params(params)
    scores, logits = e.forward(ids, seg, want_logits=True)
    assert_rel(e.path_logits(ids.shape[0]), ref["path_logits"], what="path_logits")
    assert_rel(scores, ref["scores"], what="scores")
    assert_rel(e.forward(ids, seg), ref["scores"], what="scores-only path")
    e.close()


@pytest.mark.parametrize("kw", TC_CASES)
def test_tc_forward_matches_oracle(kw):
    o, params, ids, seg, labels = make_case(**kw)
    ref = o.forward(params, ids, seg)
    e = engine_for(o, precision=1)
    e.set_params(params)
    scores, logits = e.forward(ids, seg, want_logits=True)
    assert_rel(e.path_logits(ids.shape[0]), ref["path_logits"], what="path_logits")
    assert_rel(logits, ref["seg_logits"], what="seg_logits")
    assert_rel(scores, ref["scores"], what="scores")
    # a second call with changed parameters must repack the weights
    p2 = params.copy()
    o.view(p2, "Wh")[:] *= 0.5
    e.set_params(p2)
    ref2 = o.forward(p2, ids, seg)
    assert_rel(e.forward(ids, seg), ref2["scores"], what="scores after parameter update")
    e.close()


def test_tc_big_dims_d256_full_batch_properties():
    """BASELINE size (d=256, 65 536 paths): size-independent properties instead of the (slow) oracle:
    (1) segment-permutation invariance, (2) agreement with the fp32 CUDA-core engine on the same inputs."""
    from kprn_b200 import synth
    from kprn_b200.engine import Engine
    N, V_e, d = 65536, 200000, 256
    ids = synth.make_paths(N, V_e=V_e, seed=1)
    seg = synth.make_segments(N, mode="dense", fixed_p=4)
    outs = []
    for prec in (1, 0):
        e = Engine(d_t=d, d_e=d, d_r=d, H=d, V_e=V_e, precision=prec)
        e.init_uniform(0.1, 7)
        e.zero_row(1, V_e)
        s1 = e.forward(ids, seg)
        perm = np.arange(N).reshape(-1, 4)[:, ::-1].reshape(-1)
        s2 = e.forward(ids[perm], seg)
        assert_rel(s2, s1, rtol=1e-5, what="segment permutation invariance")
        outs.append(s1)
        e.close()
    assert_rel(outs[0], outs[1], what="tensor-core vs fp32 engine at d=256, N=65536")


def test_tc_train_forward_and_backward_run():
    """precision=1 with train=1 keeps the fp32 training caches; gradients still match the oracle."""
    o, params, ids, seg, labels = make_case(N=256, seed=10, d_t=8, d_e=64, d_r=8, H=64, V_e=500)
    ref = o.forward(params, ids, seg, cache=True)
    loss_ref, dsc = o.bce(ref["scores"], labels)
    g_ref = o.backward(params, ids, seg, ref, dsc)
    e = engine_for(o, precision=1)
    e.set_params(params)
    e.zero_grad()
    e.forward(ids, seg, train=True)
    e.loss_bce(labels)
    e.backward()
    assert_rel(e.get_grads(), g_ref, what="grads")
    e.close()


TRAIN_CASES = [
    dict(N=700, seed=31, d_t=8, d_e=64, d_r=8, H=64, V_e=500),
    dict(N=1500, seed=32, d_t=16, d_e=128, d_r=16, H=128, V_e=2000),
    dict(N=333, seed=33, d_t=50, d_e=100, d_r=50, H=250, V_e=3000),          # reference default dims: unaligned strides
    dict(N=900, seed=34, d_t=32, d_e=256, d_r=32, H=256, V_e=5000),
]


@pytest.mark.parametrize("kw", TRAIN_CASES)
def test_tc_training_forward_backward_match_oracle(kw):
    """precision=3 (KPRN_PREC_TC_ALL), train=1: the forward with caches and every backward product run through the
    generic tcgen05 GEMM (tc_gemm.cu, bf16x3 split).  Scores and loss hold the 1e-4 bar; gradients are sums with heavy
    cancellation of products carrying ~2^-16 relative error each, so their bar in this mode is 5e-4 of the tensor scale
    (precision 0/1 keep the training products in fp32 and hold 1e-4, tests/test_gpu_parity.py)."""
    o, params, ids, seg, labels = make_case(**kw)
    ref = o.forward(params, ids, seg, cache=True)
    loss_ref, dsc = o.bce(ref["scores"], labels)
    g_ref = o.backward(params, ids, seg, ref, dsc)
    e = engine_for(o, precision=3)
    e.set_params(params)
    e.zero_grad()
    scores = e.forward(ids, seg, train=True)
    assert_rel(scores, ref["scores"], what="scores (train forward)")
    loss = e.loss_bce(labels)
    assert abs(loss - loss_ref) <= 1e-4 * abs(loss_ref)
    e.backward()
    g = e.get_grads()
    for name in o.shapes():
        assert_rel(o.view(g, name), o.view(g_ref, name), rtol=5e-4, what="grad " + name)
    e.close()


@pytest.mark.parametrize("kw", [TRAIN_CASES[1], TRAIN_CASES[3], dict(N=128 * 5 + 17, seed=35, d_t=16, d_e=128, d_r=16, H=200, V_e=900)])
def test_fused_bptt_kernel_equals_the_per_step_kernels(kw, monkeypatch):
    """bptt_kernel (one persistent launch for the whole time loop) runs the same arithmetic in the same order as the
    cell_bwd_kernel + bgemm_kernel launches per step (odd tile counts included); the two compilations may contract the cell
    arithmetic into FMAs differently, so the gradients agree to fp32 rounding (1e-5 of each tensor's scale), not bitwise."""
    o, params, ids, seg, labels = make_case(**kw)
    got = []
    for fused in ("1", "0"):
        monkeypatch.setenv("KPRN_TC_FUSED", fused)
        e = engine_for(o, precision=3)
        e.set_params(params)
        e.zero_grad()
        e.forward(ids, seg, train=True)
        e.loss_bce(labels)
        e.backward()
        got.append(e.get_grads().copy())
        e.close()
    for name in o.shapes():
        assert_rel(o.view(got[0], name), o.view(got[1], name), rtol=1e-5, what="grad " + name)
