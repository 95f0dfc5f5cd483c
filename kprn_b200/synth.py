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
