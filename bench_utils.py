"""Shared helpers of bench.py and the tests: synthetic LiDAR-like clouds, roofline arithmetic,
clock sampling.  Not part of the product package."""
from __future__ import annotations

import json
import os
import subprocess
import threading
import time
from typing import Dict, List, Optional

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def surface_cloud(rng: np.random.Generator, shape, n_target: int, batch: int = 1,
                  keep: float = 0.68) -> np.ndarray:
    """Unique voxel coordinates on random tilted planar patches, ``[N, 4]`` int32 (b, z, y, x),
    shuffled.  ``keep=0.68`` gives ~6.3 rulebook pairs per voxel for SubM 3^3 -- the value of
    the reference's own LiDAR fixture; uniform sampling of a KITTI-sized
    grid would give ~1.03 and benchmark nothing."""
    D, H, W = [int(s) for s in shape]
    out = []
    for b in range(batch):
        seen = np.empty((0,), np.int64)
        while seen.shape[0] < n_target:
            cz, cy, cx = rng.integers(0, D), rng.integers(0, H), rng.integers(0, W)
            ext = int(rng.integers(8, 40))
            sy, sx = rng.uniform(-0.3, 0.3, size=2)
            ys = np.arange(max(0, cy - ext), min(H, cy + ext))
            xs = np.arange(max(0, cx - ext), min(W, cx + ext))
            yy, xx = np.meshgrid(ys, xs, indexing="ij")
            zz = np.clip(np.round(cz + sy * (yy - cy) + sx * (xx - cx)).astype(np.int64), 0, D - 1)
            sel = rng.random(yy.shape) < keep
            patch = np.unique((zz[sel] * H + yy[sel]) * W + xx[sel])
            patch = patch[~np.isin(patch, seen)]
            room = n_target - seen.shape[0]
            if patch.shape[0] > room:                 # truncate the last patch, keep whole rows of it
                patch = patch[:room]
            seen = np.concatenate([seen, patch])
        keys = seen[rng.permutation(seen.shape[0])]
        z, rem = np.divmod(keys, H * W)
        y, x = np.divmod(rem, W)
        out.append(np.stack([np.full_like(z, b), z, y, x], axis=1).astype(np.int32))
    return np.concatenate(out, 0)


# ---------------------------------------------------------------------------- roofline
def load_peaks() -> Dict[str, float]:
    """Measured peaks written by the driver; else the B200_PROFILING.md fallback."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                    "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    "source": "measured"}
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback"}


def algorithmic_bytes(kind: str, n_in: int, n_out: int, c_in: int, c_out: int, kv: int,
                      elem: int) -> int:
    """Compulsory HBM bytes per launch (SURVEY 8d / BASELINE.md section 4)."""
    rulebook_read = kv * n_out * 4 + n_out * 4 + n_out * 4        # dense pair table + mask + argsort
    if kind == "fwd":
        return n_in * c_in * elem + n_out * c_out * elem + kv * c_in * c_out * elem + rulebook_read
    if kind == "dgrad":
        rb = kv * n_in * 4 + n_in * 8
        return n_out * c_out * elem + n_in * c_in * elem + kv * c_in * c_out * elem + rb
    if kind == "wgrad":
        return n_in * c_in * elem + n_out * c_out * elem + rulebook_read + kv * c_in * c_out * 4
    if kind == "rulebook_subm":
        # read coords, build + probe a 2N-slot 8-byte table, write both tables + mask
        return n_in * 16 + 2 * n_in * 8 + 2 * kv * n_in * 4 + n_in * 4
    raise ValueError(kind)


def conv_flops(pairs_total: int, c_in: int, c_out: int) -> int:
    return 2 * pairs_total * c_in * c_out


# ---------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0, period_ms: int = 100):
        self.gpu_index = gpu_index
        self.period_ms = period_ms
        self.proc: Optional[subprocess.Popen] = None
        self.lines: List[str] = []
        self._thr: Optional[threading.Thread] = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", str(self.period_ms)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def pump():
            assert self.proc is not None and self.proc.stdout is not None
            for line in self.proc.stdout:
                self.lines.append(line.strip())
        self._thr = threading.Thread(target=pump, daemon=True)
        self._thr.start()

    def stop(self) -> Dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(self.period_ms / 1000.0 * 1.5)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------- BASELINE.json configs[2]
# SECOND-style 6-layer sparse encoder (closest reference pattern
# /root/reference/test/fake_train.py:52-99): (kind, C_in, C_out, indice_key)
ENCODER6_LAYERS = [
    ("subm", 16, 16, "subm1"),
    ("subm", 16, 16, "subm1"),       # shares the rulebook of the layer above (indice_key reuse)
    ("conv", 16, 32, "down1"),       # 3x3x3, stride 2, padding 1
    ("subm", 32, 32, "subm2"),
    ("conv", 32, 64, "down2"),
    ("conv", 64, 128, "down3"),
]


def make_encoder6(spconv, algo=None, bias: bool = False, relu: bool = False):
    """The six conv layers as a list of modules (the caller chains them, so tests can look at every
    intermediate tensor); ``relu=True`` interleaves ``torch.nn.ReLU`` like SECOND does."""
    import torch
    layers = []
    for kind, c_in, c_out, key in ENCODER6_LAYERS:
        if kind == "subm":
            layers.append(spconv.SubMConv3d(c_in, c_out, 3, padding=1, bias=bias, indice_key=key, algo=algo))
        else:
            layers.append(spconv.SparseConv3d(c_in, c_out, 3, stride=2, padding=1, bias=bias, indice_key=key,
                                              algo=algo))
        if relu:
            layers.append(torch.nn.ReLU())
    return layers
