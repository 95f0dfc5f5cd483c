"""Input side of the path: the reference's `.int` text format -> int32 id tensors, and mirrors of
model/batcher/Batcher.lua / BatcherFileList.lua so the shipped bucket files can drive the engine without Lua.

`.int` line format (data/movie_data_format.py:298-314):  label \\t path;path;...   with path = step step ... (T steps,
space separated) and step = type,entity,relation (0-based vocabulary ids, left-padded with the PAD triple).  One file per
number of paths P (`<name>.<P>.int`).  data/int2torch.lua -addOne 1 shifts the ids (not the labels) by +1.
Paths relative to /root/reference/release/songPathRnn/.
"""
import os

import numpy as np


def read_int_file(path, add_one=True):
    """Returns (labels[B] float32, data[B,P,T,3] int32 1-based) -- the tensors int2torch.lua writes (there as Double)."""
    labels, rows = [], []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if not line:
                continue
            lab, inp = line.split("\t")
            paths = []
            for p in inp.split(";"):
                steps = [[int(x) for x in tok.split(",")] for tok in p.strip().split(" ")]
                paths.append(steps)
            labels.append(float(lab))
            rows.append(paths)
    P, T = len(rows[0]), len(rows[0][0])
    for r in rows:       # Util:table2tensor asserts a rectangular table (util/Util.lua:121-201)
        if len(r) != P or any(len(p) != T or any(len(s) != 3 for s in p) for p in r):
            raise ValueError("%s: ragged bucket (every pair of a bucket file must have the same P and T)" % path)
    data = np.asarray(rows, dtype=np.int32)
    if add_one:
        data = data + 1
    return np.asarray(labels, dtype=np.float32), data


def write_int_file(path, labels, data_1based):
    """Inverse of read_int_file (ids written 0-based like the formatter does)."""
    with open(path, "w") as f:
        for lab, pair in zip(labels, np.asarray(data_1based) - 1):
            f.write("%d\t%s\n" % (int(lab), ";".join(" ".join(",".join(str(int(v)) for v in s) for s in p) for p in pair)))


class Batcher:
    """model/batcher/Batcher.lua: one bucket file, optional shuffle, `narrow` minibatches."""

    def __init__(self, labels, data, batch_size, shuffle=False, class_id=1, seed=12345):
        self.labels, self.data = np.asarray(labels, dtype=np.float32), np.asarray(data)
        self.batch_size, self.do_shuffle, self.class_id = batch_size, shuffle, class_id
        self.rng = np.random.Generator(np.random.PCG64(seed))
        self.cur = 0
        if shuffle:
            self._shuffle()

    @classmethod
    def from_int_file(cls, path, batch_size, **kw):
        labels, data = read_int_file(path)
        return cls(labels, data, batch_size, **kw)

    def _shuffle(self):
        perm = self.rng.permutation(len(self.labels))
        self.labels, self.data = self.labels[perm], self.data[perm]

    def getBatch(self):
        """(labels, data, count, classId) or None at the end (Batcher.lua:43-54)."""
        if self.cur >= len(self.labels):
            return None
        a, b = self.cur, min(self.cur + self.batch_size, len(self.labels))
        self.cur = b
        return self.labels[a:b], self.data[a:b], b - a, self.class_id

    def reset(self):
        self.cur = 0
        if self.do_shuffle:
            self._shuffle()


class BatcherFileList:
    """model/batcher/BatcherFileList.lua: iterate the bucket files named in a list file, one Batcher per file."""

    def __init__(self, data_dir, batch_size, shuffle=False, list_file="train.list", seed=12345):
        names = [ln.strip() for ln in open(os.path.join(data_dir, list_file)) if ln.strip()]
        self.batchers = []
        for i, n in enumerate(names):
            p = n if os.path.isabs(n) else os.path.join(data_dir, n)
            self.batchers.append(Batcher.from_int_file(p, batch_size, shuffle=shuffle, seed=seed + i))
        self.idx = 0

    def getBatch(self):
        while self.idx < len(self.batchers):
            b = self.batchers[self.idx].getBatch()
            if b is not None:
                return b
            self.idx += 1
        return None

    def reset(self):
        self.idx = 0
        for b in self.batchers:
            b.reset()


def to_csr(batches):
    """Concatenate dense buckets [(labels, data[B,P,T,3]), ...] into one CSR batch (ids[N,T,3], seg[S+1], labels[S])."""
    ids, seg, labels, off = [], [0], [], 0
    for lab, data in batches:
        B, P = data.shape[0], data.shape[1]
        ids.append(np.asarray(data, dtype=np.int32).reshape(B * P, data.shape[2], 3))
        for _ in range(B):
            off += P
            seg.append(off)
        labels.append(np.asarray(lab, dtype=np.float32))
    return np.concatenate(ids), np.asarray(seg, dtype=np.int32), np.concatenate(labels)
