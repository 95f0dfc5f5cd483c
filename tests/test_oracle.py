"""CPU tests: the C oracle against an independent PyTorch fp64 autograd model and committed goldens."""
import numpy as np
import pytest
import torch

from oracle.kprn_oracle import Oracle
from kprn_b200 import synth
from tests import torch_ref


def _case(o, N, seed, seg_mode="realistic"):
    c = o.cfg
    ids = synth.make_paths(N, T=c.T, V_e=c.V_e, seed=seed, length_mode="stress", V_t=c.V_t, V_r=c.V_r)
    seg = synth.make_segments(N, seed=seed, mode=seg_mode)
    params = o.init_params(seed)
    for tbl, row in ((0, c.V_t), (1, c.V_e), (2, c.V_r)):
        o.zero_row(params, tbl, row)
    labels = synth.make_labels(len(seg) - 1, seed).astype(np.float64)
    return params, ids, seg, labels


CONFIGS = [
    dict(T=6, d_t=4, d_e=8, d_r=4, H=8, L=46, V_e=40),
    dict(T=3, d_t=32, d_e=32, d_r=32, H=32, L=46, V_e=500),          # BASELINE configs[0] dims
    dict(T=6, d_t=5, d_e=10, d_r=5, H=25, L=46, V_e=60),              # reference ratio 1:2:1, H odd-ish
    dict(T=6, d_t=4, d_e=8, d_r=4, H=8, L=1, V_e=40, head_layers=2),
    dict(T=6, d_t=4, d_e=8, d_r=4, H=8, L=46, V_e=40, cell=1),
    dict(T=6, d_t=4, d_e=8, d_r=4, H=8, L=46, V_e=40, pool=1),
    dict(T=6, d_t=4, d_e=8, d_r=4, H=8, L=46, V_e=40, pool=2, topk=2),
    dict(T=6, d_t=4, d_e=8, d_r=4, H=8, L=46, V_e=40, inv_gamma=0.5),
]


@pytest.mark.parametrize("kw", CONFIGS)
def test_forward_and_grads_match_autograd(kw):
    o = Oracle(**kw)
    params, ids, seg, labels = _case(o, 37, seed=7)
    fwd = o.forward(params, ids, seg, class_id=1, cache=True)
    pt = torch.tensor(params, dtype=torch.float64, requires_grad=True)
    s, lse, p = torch_ref.forward(o, pt, ids, seg)
    np.testing.assert_allclose(fwd["path_logits"], s.detach().numpy(), rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(fwd["seg_logits"], lse.detach().numpy(), rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(fwd["scores"], p.detach().numpy(), rtol=1e-11, atol=1e-13)
    loss_t = torch_ref.bce(p, torch.tensor(labels))
    loss_t.backward()
    loss_o, dsc = o.bce(fwd["scores"], labels)
    assert abs(loss_o - float(loss_t)) < 1e-12
    g = o.backward(params, ids, seg, fwd, dsc)
    np.testing.assert_allclose(g, pt.grad.numpy(), rtol=1e-9, atol=1e-13)


def test_bpr_matches_reference_formula():
    rng = np.random.default_rng(0)
    pos, neg = rng.normal(size=11), rng.normal(size=11)
    loss, dp, dn = Oracle.bpr(pos, neg)
    out = 1 / (1 + np.exp(-(pos - neg))) + 1e-4
    assert abs(loss + np.log(out).mean()) < 1e-14
    np.testing.assert_allclose(dp, -(1 - out), rtol=0, atol=1e-15)   # BPRLoss.lua:28-31: not divided by n
    np.testing.assert_allclose(dn, (1 - out), rtol=0, atol=1e-15)


def test_adam_matches_torch_optim_formula():
    rng = np.random.default_rng(1)
    x = rng.normal(size=50); g = rng.normal(size=50)
    m = np.zeros(50); v = np.zeros(50)
    x0 = x.copy()
    for t in (1, 2, 3):
        Oracle.adam(x, g, m, v, t)
    # closed form for a constant gradient
    mm = np.zeros(50); vv = np.zeros(50); xx = x0.copy()
    for t in (1, 2, 3):
        mm = 0.9 * mm + 0.1 * g; vv = 0.999 * vv + 0.001 * g * g
        xx -= 1e-3 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t) * mm / (np.sqrt(vv) + 1e-8)
    np.testing.assert_allclose(x, xx, rtol=1e-13)


def test_dense_bucket_equals_csr_and_segment_permutation_invariance():
    o = Oracle(**CONFIGS[0])
    params, ids, _, _ = _case(o, 40, seed=3)
    seg = synth.make_segments(40, mode="dense", fixed_p=4)
    a = o.forward(params, ids, seg)
    # permuting paths inside a segment leaves the pooled score unchanged
    perm = np.arange(40).reshape(10, 4)[:, ::-1].reshape(-1)
    b = o.forward(params, ids[perm], seg)
    np.testing.assert_allclose(a["scores"], b["scores"], rtol=1e-13)


def test_pad_rows_by_index_not_semantics():
    """D5: pad handling = zeroing row vocabSize of each table; steps are never masked in the LSTM."""
    o = Oracle(**CONFIGS[0])
    params, ids, seg, _ = _case(o, 20, seed=5)
    a = o.forward(params, ids, seg)
    p2 = params.copy()
    o.view(p2, "E_e")[o.cfg.V_e - 1] = 0.3    # un-zero the entity pad row -> scores of padded paths must move
    b = o.forward(p2, ids, seg)
    assert np.abs(a["path_logits"] - b["path_logits"]).max() > 1e-6


def test_bad_id_raises():
    o = Oracle(**CONFIGS[0])
    params, ids, seg, _ = _case(o, 8, seed=5)
    ids[0, 0, 1] = o.cfg.V_e + 1
    with pytest.raises(ValueError):
        o.forward(params, ids, seg)
