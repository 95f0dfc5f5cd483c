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


# ---------------------------------------------------------------------------------------------------------------------
# The formatter's rules (data/movie_data_format.py, run as `-o 0 -g 0 -m 6 -t 1` by movie_data_format.sh:2): KG path
# text -> `.int` lines.  Ported so the shipped `.translated` samples can drive the engine without Python 2 / Lua; the
# port is pinned against outputs of the reference script itself (tests/golden/formatter_ref.json).


class Vocab:
    """The four vocabularies the formatter loads (movie_data_format.py:42-77); ids are kept as strings like there."""

    def __init__(self, entity, entity_type, relation, entity_to_type, label2int):
        self.entity, self.entity_type, self.relation = entity, entity_type, relation
        self.entity_to_type, self.label2int = entity_to_type, label2int

    @staticmethod
    def _read(path):
        d = {}
        with open(path) as f:
            for line in f:
                a = line.strip().split("\t")
                if len(a) >= 2:
                    d[a[0]] = a[1]
        return d

    @classmethod
    def load(cls, vocab_dir):
        import json
        return cls(cls._read(os.path.join(vocab_dir, "all_entity_id.txt")), cls._read(os.path.join(vocab_dir, "entity_type_id.txt")),
                   cls._read(os.path.join(vocab_dir, "all_relation_id.txt")),
                   {k: [v] for k, v in cls._read(os.path.join(vocab_dir, "entity_to_type.txt")).items()},
                   json.load(open(os.path.join(vocab_dir, "domain-label"))))


def _types_in_order(v, entity_types, length):
    ids = [v.entity_type.get(t, v.entity_type["#UNK_ENTITY_TYPE"]) for t in entity_types]
    ids = sorted(ids)[:length][::-1]              # sorted as strings, sliced, reversed (movie_data_format.py:103-113)
    return ",".join(str(i) for i in ids)


def _feature(v, prev_entity, relation, slots):
    """movie_data_format.py:129-158: `types...,entity,relation` of one step."""
    out = ""
    if prev_entity in v.entity_to_type:
        types = v.entity_to_type[prev_entity]
        length = min(slots, len(types))
        for _ in range(slots - len(types)):
            out += str(v.entity_type["#PAD_TOKEN"]) + ","
        out += _types_in_order(v, types, length) + ","
    else:
        for _ in range(slots):
            out += str(v.entity_type["#UNK_ENTITY_TYPE"]) + ","
    out += str(v.entity.get(prev_entity, v.entity["#UNK_ENTITY"])) + ","
    out += str(v.relation.get(relation, v.relation["#UNK_RELATION"]))
    return out


def path_steps(path):
    return len(path.split("-")) // 2 + 2            # py2 integer division, movie_data_format.py:245


def format_translated(lines, vocab, max_length, fixed_label=None, slots=1):
    """lines of `e1 \t e2 \t path###path... [\t label]` -> {P: ["label_int\tpath;path", ...]} (one bucket per number of
    kept paths P, like the `<name>.<P>.int` files).  Left padding with the PAD triple; the last step is (e2, #END_RELATION);
    paths longer than max_length are dropped; pairs with no path left are dropped (movie_data_format.py:225-314)."""
    pad = ",".join([str(vocab.entity_type["#PAD_TOKEN"])] * slots + [str(vocab.entity["#PAD_TOKEN"]), str(vocab.relation["#PAD_TOKEN"])])
    buckets = {}
    for line in lines:
        split = line.split("\t")
        label = fixed_label if fixed_label is not None else split[3].strip()
        e1, e2 = split[0].strip(), split[1].strip()
        out_line, flag = "", 0
        for each_path in split[2].split("###"):
            each_path = each_path.strip()
            tokens = each_path.split("-")
            n = len(tokens) // 2 + 2
            if n > max_length:
                continue
            vec = " ".join([pad] * (max_length - n))
            prev = e1
            for i, tok in enumerate(tokens):
                if i % 2 == 0:
                    fv = _feature(vocab, prev, tok, slots)
                    vec = fv if (i == 0 and vec == "") else vec + " " + fv
                else:
                    prev = tok
            vec = vec + " " + _feature(vocab, e2, "#END_RELATION", slots)
            if len(vec.split(" ")) != max_length:
                continue
            if flag == 0:
                flag = 1
                out_line += vec
            else:
                out_line += ";" + vec
        out_line = out_line.strip()
        if not out_line:
            continue
        P = len(out_line.split(";"))
        buckets.setdefault(P, []).append(str(vocab.label2int["domain"][label.strip()]) + "\t" + out_line)
    return buckets


def max_path_length(all_lines, cap):
    """movie_data_format.py:80-99: min(cap, longest path over every input file)."""
    m = -1
    for line in all_lines:
        for path in line.split("\t")[2].strip().split("###"):
            m = max(m, path_steps(path))
    return min(cap, m)
