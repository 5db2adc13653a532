"""Shared helpers for the parity tests (inputs are seeded numpy -> identical for CUDA and oracle)."""
import numpy as np


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


def surface_cloud(rng, shape, n_target, batch=1):
    """Spatially clustered voxels (random planar patches) ~6 neighbours/voxel like LiDAR data
    (SURVEY 8d).  Returns unique int32 coords [N, 4], N <= n_target * batch."""
    out = []
    for b in range(batch):
        pts = set()
        D, H, W = shape
        while len(pts) < n_target:
            # a tilted planar patch
            cz, cy, cx = rng.integers(0, D), rng.integers(0, H), rng.integers(0, W)
            ext = int(rng.integers(8, 40))
            sy, sx = rng.uniform(-0.3, 0.3, size=2)
            ys = np.arange(max(0, cy - ext), min(H, cy + ext))
            xs = np.arange(max(0, cx - ext), min(W, cx + ext))
            yy, xx = np.meshgrid(ys, xs, indexing="ij")
            zz = np.clip(np.round(cz + sy * (yy - cy) + sx * (xx - cx)).astype(np.int64), 0, D - 1)
            keep = rng.random(yy.shape) < 0.85
            for z, y, x in zip(zz[keep], yy[keep], xx[keep]):
                pts.add((int(z), int(y), int(x)))
                if len(pts) >= n_target:
                    break
        arr = np.array(sorted(pts), dtype=np.int32)
        arr = arr[rng.permutation(arr.shape[0])]
        out.append(np.concatenate([np.full((arr.shape[0], 1), b, np.int32), arr], axis=1))
    return np.concatenate(out, 0)


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
