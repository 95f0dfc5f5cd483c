"""GPU tests of the reference-facing operator surface (kprn_b200/nn.py = the Python twin of the LuaJIT shim): they read
like the reference's drivers (MyOptimizer.lua:177-221, test_from_checkpoint.lua:98-124)."""
import numpy as np
import pytest

from kprn_b200 import io, nn
from tests.util import assert_rel, make_case

pytestmark = pytest.mark.gpu


def _model(o, precision):
    c = o.cfg
    return nn.KPRNPathScorer(c.V_t, c.d_t, c.V_e, c.d_e, c.V_r, c.d_r, c.H, rnnType="lstm" if c.cell == 0 else "rnn",
                             topK=2, labelDimension=c.L, classId=1, maxPathLength=c.T, precision=precision)


@pytest.mark.parametrize("precision", [0, 1])
def test_forward_backward_like_MyOptimizer_fEval(precision):
    o, params, ids, seg, labels = make_case(N=120, seed=41, seg_mode="dense", fixed_p=4, d_t=8, d_e=64, d_r=8, H=64, V_e=300)
    inputs = ids.reshape(30, 4, o.cfg.T, 3).astype(np.float64)            # Batcher tensor (DoubleTensor of 1-based ids)
    ref = o.forward(params, ids, seg, cache=True)
    loss_ref, dsc = o.bce(ref["scores"], labels.astype(np.float64))
    g_ref = o.backward(params, ids, seg, ref, dsc)

    model = _model(o, precision).training()
    model.setParameters(params)
    criterion = nn.BCECriterion(model)
    model.zeroGradParameters()
    output = model.forward(inputs)
    err = criterion.forward(output, labels)
    df_do = criterion.backward(output, labels)
    gi = model.backward(inputs, df_do)
    assert gi.size == 0                                                   # MapReduce.lua:83: no grad w.r.t. ids
    assert_rel(output, ref["scores"], what="output")
    assert abs(err - loss_ref) <= 1e-4 * loss_ref
    _, g = model.getParameters()
    assert_rel(g, g_ref, what="gradParameters")
    # updateGradInput and accGradParameters invoked separately (MapReduce.lua:58-71), explicit df_do, idempotent
    model.zeroGradParameters()
    model.forward(inputs)
    model.updateGradInput(inputs, dsc)
    model.accGradParameters(inputs)
    model.accGradParameters(inputs)
    assert_rel(model.getParameters()[1], g_ref, what="split backward")


def test_scoring_loop_like_test_from_checkpoint(tmp_path):
    from kprn_b200 import evalmetrics
    o, params, ids, seg, labels = make_case(N=96, seed=42, seg_mode="dense", fixed_p=3, d_t=8, d_e=64, d_r=8, H=64, V_e=300)
    data = ids.reshape(32, 3, o.cfg.T, 3)
    io.write_int_file(str(tmp_path / "test.txt.3.int"), labels, data)
    (tmp_path / "test.list").write_text("test.txt.3.int\n")
    batcher = io.BatcherFileList(str(tmp_path), 10, shuffle=False, list_file="test.list")
    model = _model(o, 1).evaluate()
    model.setParameters(params)
    preds, labs = [], []
    while True:
        b = batcher.getBatch()
        if b is None:
            break
        preds.extend(model.forward(b[1]).tolist())
        labs.extend(b[0].tolist())
    ref = o.forward(params, ids, seg)
    assert_rel(np.array(preds), ref["scores"], what="scores")
    out = tmp_path / "test.res"
    evalmetrics.write_scores(str(out), preds, labs)
    lines = out.read_text().splitlines()
    assert len(lines) == 32
    assert [ln.split("\t")[1] for ln in lines] == ["%.5f" % s for s in ref["scores"]]     # identical 5-decimal strings


def test_MyOptimizer_train_epochs_decrease_loss(tmp_path):
    o, params, ids, seg, labels = make_case(N=256, seed=43, seg_mode="dense", fixed_p=2, d_t=8, d_e=16, d_r=8, H=16, V_e=80)
    data = ids.reshape(128, 2, o.cfg.T, 3)
    io.write_int_file(str(tmp_path / "train.txt.2.int"), labels, data)
    (tmp_path / "train.list").write_text("train.txt.2.int\n")
    batcher = io.BatcherFileList(str(tmp_path), 32, shuffle=True, list_file="train.list")
    c = o.cfg
    model = _model(o, 0)
    model.setParameters(params)
    opt = nn.MyOptimizer(model, learningRate=5e-3, entityTypePadToken=c.V_t, entityPadToken=c.V_e, relationPadToken=c.V_r, numEpochs=6)
    hist = opt.train(batcher)
    assert hist[-1] < hist[0]
    p = model.getParameters()[0]
    assert np.all(o.view(p, "E_e")[c.V_e - 1] == 0)           # zeroPadTokens after the step (MyOptimizer.lua:219)
