"""Generates the committed fixtures under tests/golden/ (run here, in the build container).

1. oracle_*.npz  -- inputs + fp64 oracle outputs (path/segment logits, scores, BCE loss, flat gradient) for small
   configurations.  PARITY UNPINNED: the reference ships no numeric goldens for this path and cannot run here, so these
   pin the restatement (oracle/kprn_oracle.c) against regressions; its correctness is argued by the independent
   autograd model in tests/test_oracle.py.
2. eval_score_ref.json -- outputs of the REFERENCE's own eval_one_rating (eval/eval_score.py:34-46), imported from
   /root/reference, on seeded inputs incl. ties, all-zero rows and 5-decimal score strings: golden vectors for the
   metric port in kprn_b200/evalmetrics.py.
3. resort_ref.json -- a slice of the reference's shipped combine / sorted fixtures (eval/config1/*) to pin resort().
"""
import importlib.util
import zlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.kprn_oracle import Oracle  # noqa: E402
from kprn_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference/release/songPathRnn"

CASES = {
    "lstm_small": dict(N=96, seed=101, T=6, d_t=8, d_e=16, d_r=8, H=16, L=46, V_e=200),
    "lstm_tc64": dict(N=200, seed=102, T=6, d_t=8, d_e=64, d_r=8, H=64, L=46, V_e=400),
    "lstm_refratio": dict(N=120, seed=103, T=6, d_t=5, d_e=10, d_r=5, H=25, L=46, V_e=150),
    "rnn_small": dict(N=96, seed=104, T=6, d_t=8, d_e=16, d_r=8, H=16, L=46, V_e=200, cell=1),
    "kkbox_plumbing": dict(N=1000, seed=105, T=3, d_t=32, d_e=32, d_r=32, H=32, L=46, V_e=2000),   # BASELINE configs[0]
}


def make_oracle_case(name, N, seed, **kw):
    o = Oracle(**kw)
    c = o.cfg
    ids = synth.make_paths(N, T=c.T, V_e=c.V_e, seed=seed, length_mode="stress", V_t=c.V_t, V_r=c.V_r)
    seg = synth.make_segments(N, seed=seed)
    params = o.init_params(seed)
    for tbl, row in ((0, c.V_t), (1, c.V_e), (2, c.V_r)):
        o.zero_row(params, tbl, row)
    labels = synth.make_labels(len(seg) - 1, seed).astype(np.float64)
    fwd = o.forward(params, ids, seg, cache=True)
    loss, dsc = o.bce(fwd["scores"], labels)
    g = o.backward(params, ids, seg, fwd, dsc)
    np.savez_compressed(os.path.join(OUT, "oracle_%s.npz" % name), cfg=json.dumps(dict(N=N, seed=seed, **kw)),
                        params=params.astype(np.float64), ids=ids, seg=seg, labels=labels,
                        path_logits=fwd["path_logits"], seg_logits=fwd["seg_logits"], scores=fwd["scores"],
                        loss=np.float64(loss), grads=g)


def make_eval_goldens():
    spec = importlib.util.spec_from_file_location("ref_eval_score", os.path.join(REF, "eval", "eval_score.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.Generator(np.random.PCG64(12345))
    cases = []
    for i in range(200):
        n_neg = 100 if i % 3 else int(rng.integers(1, 20))
        mode = i % 5
        if mode == 0:
            pre = [float("%.5f" % v) for v in rng.random(n_neg + 1)]
        elif mode == 1:                       # saturated sigmoid scores: many exact ties at 1.00000
            pre = [float("%.5f" % min(1.0, v)) for v in 0.99995 + 1e-4 * rng.random(n_neg + 1)]
        elif mode == 2:
            pre = [0.0] * (n_neg + 1)         # sum == 0 -> (0, 0)
        elif mode == 3:                       # positive tied with several negatives
            pre = [0.5] + [0.5 if rng.random() < 0.3 else float("%.5f" % rng.random()) for _ in range(n_neg)]
        else:
            pre = [float("%.5f" % v) for v in rng.normal(0.5, 0.2, n_neg + 1).clip(0, 1)]
        gnd = [1.0] + [0.0] * n_neg
        res = [list(mod.eval_one_rating(i_gnd=gnd, i_pre=pre, K=k)) for k in range(1, 16)]
        cases.append({"pre": pre, "res": res})
    json.dump(cases, open(os.path.join(OUT, "eval_score_ref.json"), "w"))
    # resort fixture: the first 400 lines of the shipped combine sample and the reference-sorted order of the same lines
    comb = [ln.rstrip("\n").split("\t") for ln in open(os.path.join(REF, "eval", "config1", "test_combine_sample.txt"))][:400]
    users = sorted({r[0] for r in comb})
    items = [(int(r[0]), r[1], r[2], float(r[3])) for r in comb]
    ref_sorted = sorted(items, key=lambda x: (x[0], -x[-1]))      # eval/resort.py:38
    json.dump({"combine": comb, "users": users, "sorted": [[str(a), b, c, str(d)] for a, b, c, d in ref_sorted]},
              open(os.path.join(OUT, "resort_ref.json"), "w"))


def make_formatter_goldens(n_lines=250):
    """Runs the REFERENCE formatter (data/movie_data_format.py, Python-2 source: `xrange` and integer `/` are patched to
    their Python-3 spellings, nothing else) on the first lines of the shipped samples in a scratch directory and stores
    inputs + the `.int` files it writes."""
    import shutil
    import subprocess
    import tempfile
    tmp = tempfile.mkdtemp(prefix="kprn_fmt_")
    os.makedirs(os.path.join(tmp, "input")); os.makedirs(os.path.join(tmp, "vocab"))
    inputs = {}
    for name in ("positive", "negative", "test"):
        src = os.path.join(REF, "data", "input", "%s_matrix_sample.tsv.translated" % name)
        lines = [ln.rstrip("\n") for ln in open(src)][:n_lines]
        inputs[name] = lines
        open(os.path.join(tmp, "input", "%s_matrix.tsv.translated" % name), "w").write("\n".join(lines) + "\n")
    # vocabularies: the shipped ones; the entity vocab is a 10 000-row sample without the special rows, so they are appended
    ent = [ln.rstrip("\n") for ln in open(os.path.join(REF, "data", "vocab", "all_entity_id_sample.txt"))]
    seen = {ln.split("\t")[0] for ln in ent}
    extra = []
    for lines in inputs.values():          # give every third entity of the inputs an id, the rest stay #UNK_ENTITY
        for ln in lines:
            for tok in ln.split("\t")[2].replace("###", "-").split("-"):
                if tok and tok[0] in "sup" and tok not in seen and (zlib.crc32(tok.encode()) % 3 == 0 or len(extra) < 5):
                    seen.add(tok); extra.append(tok)
    extra = sorted(extra)
    ent += ["%s\t%d" % (t, 10000 + i) for i, t in enumerate(extra)]
    ent += ["#UNK_ENTITY\t%d" % (10000 + len(extra)), "#PAD_TOKEN\t%d" % (10001 + len(extra))]
    open(os.path.join(tmp, "vocab", "all_entity_id.txt"), "w").write("\n".join(ent) + "\n")
    e2t = [ln.rstrip("\n") for ln in open(os.path.join(REF, "data", "vocab", "entity_to_type_sample.txt"))]
    e2t += ["%s\t%s" % (t, {"s": "song", "u": "user", "p": "person"}[t[0]]) for t in extra[::2]]
    open(os.path.join(tmp, "vocab", "entity_to_type.txt"), "w").write("\n".join(e2t) + "\n")
    for f in ("entity_type_id.txt", "all_relation_id.txt", "domain-label"):
        shutil.copy(os.path.join(REF, "data", "vocab", f), os.path.join(tmp, "vocab", f))
    src = open(os.path.join(REF, "data", "movie_data_format.py")).read()
    src = src.replace("xrange", "range").replace("path_len / 2 + 2", "path_len // 2 + 2")
    open(os.path.join(tmp, "ref_formatter_py3.py"), "w").write(src)
    subprocess.check_call([sys.executable, "ref_formatter_py3.py", "-i", "input", "-d", "output", "-o", "0", "-g", "0", "-e", "0",
                           "-m", "6", "-t", "1"], cwd=tmp, stdout=subprocess.DEVNULL)
    outputs = {}
    for sub in ("train", "test"):
        for f in sorted(os.listdir(os.path.join(tmp, "output", sub))):
            outputs[sub + "/" + f] = [ln.rstrip("\n") for ln in open(os.path.join(tmp, "output", sub, f))]
    vocab = {f: open(os.path.join(tmp, "vocab", f)).read() for f in ("all_entity_id.txt", "entity_to_type.txt", "entity_type_id.txt",
                                                                       "all_relation_id.txt", "domain-label")}
    json.dump({"inputs": inputs, "vocab": vocab, "outputs": outputs}, open(os.path.join(OUT, "formatter_ref.json"), "w"))
    shutil.rmtree(tmp)
    print("formatter goldens:", {k: len(v) for k, v in outputs.items()})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    make_formatter_goldens()
    for name, kw in CASES.items():
        make_oracle_case(name, **kw)
    make_eval_goldens()
    print(sorted(os.listdir(OUT)))
