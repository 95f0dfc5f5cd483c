"""Deterministic synthetic KG-path batches.

Layout follows the reference's formatter: ids[N,T,3] int32, 1-based, columns
(type, entity, relation) (release/songPathRnn/data/movie_data_format.py:142-157,
int2torch.lua:58-63 "+1"); paths are LEFT-padded to T with the PAD triple
(movie_data_format.py:161-186,250-254); the last real step carries #END_RELATION
(movie_data_format.py:282). Shipped vocab (+1): type PAD=5 (V_t=6), relation PAD=8,
END=9 (V_r=9); the entity table's last row V_e is PAD.

RNG: numpy Generator(PCG64(seed)), seed 12345 = the constant in model/OneModel.lua:115.
"""
import numpy as np

V_T, V_R = 6, 9
PAD_TYPE, PAD_REL, END_REL = 5, 8, 9


def make_paths(n_paths, T=6, V_e=2851220, seed=12345, length_mode="realistic",
               entity_mode="zipf", V_t=V_T, V_r=V_R):
    """Returns ids[N,T,3] int32 (1-based)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    N = int(n_paths)
    if length_mode == "realistic":      # shipped sample: 70 % 4-step, 30 % 6-step paths
        ln = np.where(rng.random(N) < 0.7, 4, 6)
    elif length_mode == "stress":
        ln = rng.integers(2, 7, N)
    else:
        ln = np.full(N, int(length_mode))
    ln = np.minimum(ln, T).astype(np.int64)
    pad_t = min(PAD_TYPE, V_t)
    pad_r = min(PAD_REL, V_r)
    end_r = min(END_REL, V_r)
    n_types = max(1, min(4, V_t - 2))
    n_rels = max(1, min(6, V_r - 3))
    ids = np.empty((N, T, 3), dtype=np.int32)
    ids[:, :, 0] = rng.integers(1, n_types + 1, (N, T))
    if entity_mode == "zipf":           # heavy-tailed like users/songs
        z = rng.zipf(1.05, (N, T)).astype(np.float64)
        ent = np.mod(z - 1.0, max(1, V_e - 2)).astype(np.int64) + 1
    else:
        ent = rng.integers(1, max(2, V_e - 1), (N, T))
    ids[:, :, 1] = ent
    ids[:, :, 2] = rng.integers(1, n_rels + 1, (N, T))
    ids[:, T - 1, 2] = end_r
    step = np.arange(T)[None, :]
    pad = step < (T - ln)[:, None]
    ids[:, :, 0][pad] = pad_t
    ids[:, :, 1][pad] = V_e
    ids[:, :, 2][pad] = pad_r
    return ids


def make_segments(n_paths, seed=12345, mode="realistic", fixed_p=4):
    """CSR offsets seg[S+1] int32 over consecutive paths.

    realistic: paths/pair ~ 1+Geometric(p=0.57) clipped to 28 (mean ~1.75, like the shipped sample);
    dense: every pair has fixed_p paths (a reference bucket file, Batcher.lua:26-28)."""
    N = int(n_paths)
    if mode == "dense":
        assert N % fixed_p == 0
        return np.arange(0, N + 1, fixed_p, dtype=np.int32)
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    sizes = np.minimum(rng.geometric(0.57, N), 28)
    cs = np.cumsum(sizes)
    k = int(np.searchsorted(cs, N, side="left"))
    offs = np.concatenate([[0], cs[:k], [N]]).astype(np.int64)
    offs = np.unique(offs)
    return offs.astype(np.int32)


def make_labels(n_seg, seed=12345):
    rng = np.random.Generator(np.random.PCG64(seed + 2))
    return (rng.random(int(n_seg)) < 0.5).astype(np.float32)


def device_uniform(indices, scale=0.1, seed=12345):
    """Values kprn_init_uniform (csrc/api.cu init_uniform_kernel: splitmix64 of seed*K + i, top 24 bits) writes at the flat
    parameter indices `indices` -- lets a checker rebuild any slice of a 2.85M-row table without copying it off the device."""
    i = np.asarray(indices, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = np.uint64(seed) * np.uint64(0xD1342543DE82EF95) + i
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    u = ((x >> np.uint64(40)).astype(np.float64) * (1.0 / 16777216.0)).astype(np.float32)
    return ((np.float32(2.0) * u - np.float32(1.0)) * np.float32(scale)).astype(np.float32)


def compact_entities(ids):
    """Renumbers the entity column of ids[N,T,3] to 1..K over the rows actually used (order of first id value).
    Returns (ids_compact, rows_1based[K]): the checker can then hold a K-row table instead of the full vocabulary."""
    ids = np.array(ids, dtype=np.int32, copy=True)
    uniq, inv = np.unique(ids[:, :, 1], return_inverse=True)
    ids[:, :, 1] = inv.reshape(ids.shape[0], ids.shape[1]).astype(np.int32) + 1
    return ids, uniq.astype(np.int64)


def compact_params(full_offsets, compact_offsets, n_compact, rows_1based, d_e, sizes, scale=0.1, seed=12345, zero_rows=()):
    """Flat fp64 parameter vector of the COMPACT model (entity table = rows_1based of the full table, everything else
    identical) holding exactly the values kprn_init_uniform(scale, seed) put on the device.
    full_offsets / compact_offsets: {name: flat offset} of the two layouts (engine.offsets / the checker's); sizes:
    {name: element count} of every tensor but E_e; zero_rows: [(name, row_1based_in_FULL_numbering, width)] rows zeroed
    after the init (MyOptimizer:zeroPadTokens)."""
    out = np.zeros(int(n_compact), dtype=np.float64)
    for name, n in sizes.items():
        fo, co = full_offsets[name], compact_offsets[name]
        if fo < 0 or co < 0 or name == "E_e":
            continue
        out[co:co + n] = device_uniform(fo + np.arange(n, dtype=np.int64), scale, seed)
    rows = np.asarray(rows_1based, dtype=np.int64) - 1
    idx = full_offsets["E_e"] + rows[:, None] * d_e + np.arange(d_e, dtype=np.int64)[None, :]
    co = compact_offsets["E_e"]
    out[co:co + idx.size] = device_uniform(idx.reshape(-1), scale, seed)
    for name, row, width in zero_rows:
        if name == "E_e":
            k = np.nonzero(np.asarray(rows_1based) == row)[0]
            for kk in k:
                out[co + kk * d_e: co + (kk + 1) * d_e] = 0.0
        else:
            c0 = compact_offsets[name] + (row - 1) * width
            out[c0:c0 + width] = 0.0
    return out


def full_offsets(d_t, d_e, d_r, H, L=46, V_t=V_T, V_e=2851220, V_r=V_R):
    """Flat offsets of the LSTM / 1-layer-head parameter layout (E_t,E_e,E_r,Wx,bx,Wh,Wo,bo) -- the same as Engine.offsets,
    computable without a device (csrc/api.cu make_layout)."""
    D, G = d_t + d_e + d_r, 4 * H
    o, out = 0, {}
    for name, n in (("E_t", V_t * d_t), ("E_e", V_e * d_e), ("E_r", V_r * d_r), ("Wx", G * D), ("bx", G), ("Wh", G * H)):
        out[name] = o; o += n
    out["bh"] = out["W1"] = out["b1"] = -1
    out["Wo"] = o; o += L * H
    out["bo"] = o; o += L
    out["total"] = o
    return out
