"""Regenerates tests/golden/*.  Run in the build container only (needs /root/reference).

* ``fixture_coords.npz``  -- voxel coordinates of the reference's own LiDAR fixture
  (test/data/test_spconv.pkl: 125 562 voxels, shape [80, 1600, 1600]); data, not source.
* ``fixture_facts.json``  -- rulebook facts of that fixture computed with an implementation that
  shares nothing with oracle/ (sorted linear keys + np.searchsorted): per-offset SubM pair counts,
  pair / output counts of SparseConv3d(k3, s2, p1).  BASELINE.md section 2 quotes the same totals.
* ``dense_conv_case.npz`` -- one seeded dense-equivalence case of test/test_conv.py:247-357
  (inputs + torch.nn.functional.conv3d outputs / input-grad / weight-grad), so the GPU box can
  check the kernels against torch's dense conv without regenerating anything.
"""
import json
import os
import pickle

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def linear(c, shape):
    k = c[:, 0].astype(np.int64)
    for a, d in enumerate(shape):
        k = k * d + c[:, a + 1]
    return k


def main():
    voxels, coors, shape = pickle.load(open("/root/reference/test/data/test_spconv.pkl", "rb"))
    coors = np.ascontiguousarray(coors.astype(np.int32))
    np.savez_compressed(os.path.join(HERE, "fixture_coords.npz"), coors=coors,
                        shape=np.array(shape, np.int32))
    keys = linear(coors, shape)
    order = np.argsort(keys)
    skeys = keys[order]
    counts = []
    for kz in range(3):
        for ky in range(3):
            for kx in range(3):
                nb = coors.astype(np.int64).copy()
                nb[:, 1] += 1 - kz
                nb[:, 2] += 1 - ky
                nb[:, 3] += 1 - kx
                ok = np.all((nb[:, 1:] >= 0) & (nb[:, 1:] < np.array(shape)), axis=1)
                nk = linear(nb[ok], shape)
                pos = np.searchsorted(skeys, nk)
                pos[pos >= len(skeys)] = 0
                counts.append(int((skeys[pos] == nk).sum()))
    # strided conv k3 s2 p1
    oshape = [(s + 2 - 2 - 1) // 2 + 1 for s in shape]
    pairs = 0
    outs = set()
    for kz in range(3):
        for ky in range(3):
            for kx in range(3):
                h = coors[:, 1:].astype(np.int64) + 1 - np.array([kz, ky, kx])
                ok = np.all((h % 2 == 0) & (h >= 0) & (h // 2 < np.array(oshape)), axis=1)
                o = h[ok] // 2
                pairs += int(ok.sum())
                ok_keys = (o[:, 0] * oshape[1] + o[:, 1]) * oshape[2] + o[:, 2]
                outs.update(np.unique(ok_keys).tolist())
    facts = {"num_voxels": int(coors.shape[0]), "shape": [int(s) for s in shape],
             "subm_k3_pairs_per_offset": counts, "subm_k3_pairs_total": int(sum(counts)),
             "conv_k3s2p1_pairs": pairs, "conv_k3s2p1_outputs": len(outs),
             "conv_k3s2p1_out_shape": oshape}
    json.dump(facts, open(os.path.join(HERE, "fixture_facts.json"), "w"), indent=1)
    print(facts["subm_k3_pairs_total"], pairs, len(outs))

    # dense-equivalence golden (test/test_conv.py:247-357 with its seeds and a small grid)
    np.random.seed(484)
    torch.manual_seed(48848)
    shape3, bs, npts, C, K = [19, 18, 17], 2, 1500, 16, 16
    total = int(np.prod(shape3))
    inds = []
    for b in range(bs):
        flat = np.random.permutation(total)[:npts]
        cc = np.stack(np.unravel_index(flat, shape3), -1).astype(np.int32)
        inds.append(np.concatenate([np.full((npts, 1), b, np.int32), cc], 1))
    inds = np.concatenate(inds, 0)
    feats = np.random.uniform(-1, 1, size=(inds.shape[0], C)).astype(np.float32)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import oracle as orc
    out = {}
    for tag, (k, s, p, d) in {"k3s2p1d1": (3, 2, 1, 1), "k3s1p1d1": (3, 1, 1, 1),
                              "k2s2p0d1": (2, 2, 0, 1)}.items():
        w = np.random.uniform(-1, 1, size=(K, k, k, k, C)).astype(np.float32)
        dense = torch.zeros((bs, C, *shape3))
        dense[inds[:, 0], :, inds[:, 1], inds[:, 2], inds[:, 3]] = torch.from_numpy(feats)
        dense.requires_grad_(True)
        wt = torch.from_numpy(w).permute(0, 4, 1, 2, 3).contiguous().requires_grad_(True)
        y = torch.nn.functional.conv3d(dense, wt, stride=s, padding=p, dilation=d)
        dy = torch.from_numpy(np.random.uniform(-0.2, 0.2, size=tuple(y.shape)).astype(np.float32))
        # the sparse op only defines gradients through its ACTIVE outputs: mask dy to them
        oi, _, _ = orc.get_indice_pairs(inds, bs, shape3, [k] * 3, [s] * 3, [p] * 3, [d] * 3, [0] * 3, False)
        act = torch.zeros_like(y)
        act[oi[:, 0], :, oi[:, 1], oi[:, 2], oi[:, 3]] = 1
        assert float((y.detach() * (1 - act)).abs().max()) == 0.0      # inactive outputs are exactly 0
        y.backward(dy * act)
        out[f"{tag}_w"] = w
        out[f"{tag}_y"] = y.detach().numpy()
        out[f"{tag}_dy"] = dy.numpy()
        out[f"{tag}_dw"] = wt.grad.permute(0, 2, 3, 4, 1).contiguous().numpy()      # back to KRSC
        out[f"{tag}_dx"] = dense.grad[inds[:, 0], :, inds[:, 1], inds[:, 2], inds[:, 3]].numpy()
    np.savez_compressed(os.path.join(HERE, "dense_conv_case.npz"), inds=inds, feats=feats,
                        shape=np.array(shape3), **out)


if __name__ == "__main__":
    main()
