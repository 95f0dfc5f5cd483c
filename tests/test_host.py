"""CPU tests of the host-side logic around the path: .int reader, Batcher mirrors, CSR conversion, partitioning."""
import numpy as np

from kprn_b200 import dist, io, synth


def _bucket(B, P, T=6, seed=0):
    ids = synth.make_paths(B * P, T=T, V_e=100, seed=seed, length_mode="stress").reshape(B, P, T, 3)
    labels = synth.make_labels(B, seed)
    return labels, ids


def test_int_file_round_trip_and_add_one(tmp_path):
    labels, data = _bucket(7, 3)
    p = tmp_path / "train.txt.3.int"
    io.write_int_file(str(p), labels, data)
    first = p.read_text().splitlines()[0]
    lab, paths = first.split("\t")
    assert len(paths.split(";")) == 3 and len(paths.split(";")[0].split(" ")) == 6      # P paths, T steps
    assert [int(v) + 1 for v in paths.split(";")[0].split(" ")[0].split(",")] == data[0, 0, 0].tolist()   # file is 0-based
    l2, d2 = io.read_int_file(str(p))
    np.testing.assert_array_equal(d2, data)
    np.testing.assert_array_equal(l2, labels)
    assert d2.min() >= 1                                      # int2torch.lua -addOne 1


def test_batcher_matches_reference_semantics(tmp_path):
    """model/batcher/test/testBatcher.lua: getBatch non-nil, classId, same number of batches before/after reset."""
    labels, data = _bucket(22, 2)
    b = io.Batcher(labels, data, batch_size=4, shuffle=False, class_id=1)
    n = 0
    while True:
        got = b.getBatch()
        if got is None:
            break
        lab, dat, count, cls = got
        assert cls == 1 and dat.shape[0] == count == len(lab)
        n += 1
    assert n == 6
    b.reset()
    n2 = sum(1 for _ in iter(b.getBatch, None))
    assert n2 == 6
    # shuffle keeps (label, data) rows together
    bs = io.Batcher(labels, data, batch_size=22, shuffle=True)
    lab, dat, _, _ = bs.getBatch()
    key = {tuple(d.reshape(-1)): l for l, d in zip(labels, data)}
    assert all(key[tuple(d.reshape(-1))] == l for l, d in zip(lab, dat))


def test_batcher_file_list_and_csr(tmp_path):
    names = []
    for P in (1, 2, 5):
        labels, data = _bucket(9, P, seed=P)
        name = "train.txt.%d.int" % P
        io.write_int_file(str(tmp_path / name), labels, data)
        names.append(name)
    (tmp_path / "train.list").write_text("\n".join(names) + "\n")
    fl = io.BatcherFileList(str(tmp_path), batch_size=4, shuffle=False, list_file="train.list")
    batches = list(iter(fl.getBatch, None))
    assert len(batches) == 9                                  # 3 files x ceil(9/4)
    fl.reset()
    assert len(list(iter(fl.getBatch, None))) == 9            # testBatcherFilelist.lua: stable across reset
    ids, seg, labels = io.to_csr([(b[0], b[1]) for b in batches])
    assert ids.shape[0] == seg[-1] == 9 * (1 + 2 + 5) and len(labels) == len(seg) - 1 == 27
    assert (np.diff(seg) > 0).all()


def test_partition_never_splits_a_segment_and_balances_paths():
    seg = synth.make_segments(10000, seed=4)
    for world in (1, 2, 3, 8):
        parts = dist.partition_segments(seg, world)
        assert parts[0][0] == 0 and parts[-1][1] == len(seg) - 1
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        counts = [int(seg[b] - seg[a]) for a, b in parts]
        assert sum(counts) == 10000 and max(counts) - min(counts) <= 28 * 2
    ids = synth.make_paths(10000, V_e=50, seed=4)
    l_ids, l_seg, _, (a, b) = dist.shard(ids, seg, None, 1, 4)
    assert l_seg[0] == 0 and l_seg[-1] == len(l_ids) and len(l_seg) == b - a + 1
    np.testing.assert_array_equal(l_ids, ids[seg[a]:seg[b]])


def test_formatter_port_reproduces_reference_script_outputs(tmp_path):
    """tests/golden/formatter_ref.json holds `.int` files written by the reference's own data/movie_data_format.py
    (run by tools/make_golden.py on the first lines of the shipped samples); the port must reproduce them byte for byte."""
    import json
    import os
    z = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "formatter_ref.json")))
    for name, text in z["vocab"].items():
        (tmp_path / name).write_text(text)
    v = io.Vocab.load(str(tmp_path))
    all_lines = z["inputs"]["positive"] + z["inputs"]["negative"] + z["inputs"]["test"]
    T = io.max_path_length(all_lines, 6)
    assert T == 6
    train = {}
    for name, label in (("positive", "1"), ("negative", "-1")):          # both append to train.txt.<P>.int, in this order
        for P, rows in io.format_translated(z["inputs"][name], v, T, fixed_label=label).items():
            train.setdefault(P, []).extend(rows)
    test = io.format_translated(z["inputs"]["test"], v, T)
    got = {"train/train.txt.%d.int" % P: rows for P, rows in train.items()}
    got.update({"test/test.txt.%d.int" % P: rows for P, rows in test.items()})
    assert sorted(got) == sorted(z["outputs"])
    for k in got:
        assert got[k] == z["outputs"][k], k
    # and the reader turns a bucket into the tensors int2torch.lua would: [B,P,T,3], ids shifted by +1
    p = tmp_path / "train.txt.2.int"
    p.write_text("\n".join(got["train/train.txt.2.int"]) + "\n")
    labels, data = io.read_int_file(str(p))
    assert data.shape[1:] == (2, 6, 3) and data.min() >= 1 and set(np.unique(labels)) <= {0.0, 1.0}


def test_bench_prefix_bookkeeping_matches_the_sort_rule():
    """bench.prefix_recur_products replays tc_prefix.cu on the host: paths sorted by leading-token count, a pair of tiles
    (256 rows) starts at the count of its first row."""
    import bench
    from kprn_b200 import synth
    ids = synth.make_paths(1000, T=6, V_e=500, seed=5, length_mode="realistic")          # 70 % two pads, 30 % none
    got = bench.prefix_recur_products(ids)
    dense = len(ids) * 5
    assert 0 < got <= dense
    n0 = int((ids[:, 0, 1] != ids[:, 1, 1]).sum())                                          # paths without a repeated leading token
    # rows [0, n0) have no shared prefix; the tile pair that straddles the boundary runs all 5 products
    expect = sum((5 - (0 if r0 < n0 else 2)) * min(256, 1000 - r0) for r0 in range(0, 1000, 256))
    assert got == expect
    assert bench.prefix_recur_products(synth.make_paths(300, T=6, V_e=500, seed=6, length_mode=6)) == 300 * 5
