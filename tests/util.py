"""Shared helpers for the parity tests (inputs are seeded numpy -> identical for CUDA and oracle)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def random_cloud(rng, shape, num_per_batch, channels, dtype=np.float32):
    """Unique uniform-random coordinates per sample (spconv/test_utils.py:142-195 semantics)."""
    total = int(np.prod(shape))
    inds = []
    for b, n in enumerate(num_per_batch):
        flat = rng.permutation(total)[:n]
        coords = np.stack(np.unravel_index(flat, shape), axis=-1).astype(np.int32)
        inds.append(np.concatenate([np.full((n, 1), b, np.int32), coords], axis=1))
    indices = np.concatenate(inds, 0)
    feats = rng.uniform(-1, 1, size=(indices.shape[0], channels)).astype(dtype)
    return feats, indices


from bench_utils import surface_cloud  # noqa: E402,F401  (clustered LiDAR-like clouds)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def describe_mismatch(got, ref, name="", max_rows=5):
    """Human-readable summary of where two matrices differ (used in assertion messages so one
    GPU run tells as much as possible)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    err = np.abs(got - ref)
    bad = err > (1e-2 + 1e-2 * np.abs(ref))
    rows = np.unique(np.nonzero(bad)[0])
    cols = np.unique(np.nonzero(bad)[1]) if bad.ndim > 1 else []
    msg = (f"{name}: shape {got.shape} max_abs_err {err.max():.4g} rel_l2 {rel_l2(got, ref):.4g} "
           f"bad {bad.sum()}/{bad.size} bad_rows {len(rows)} (first {rows[:max_rows].tolist()}) "
           f"bad_cols {len(cols)} (first {list(cols[:16])}) nan {np.isnan(got).sum()}")
    if len(rows):
        r = rows[0]
        msg += f"\n  row {r} got {np.round(got[r][:8], 3).tolist()} ref {np.round(ref[r][:8], 3).tolist()}"
    return msg
