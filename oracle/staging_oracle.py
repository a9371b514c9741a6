"""TEST INFRASTRUCTURE ONLY — CPU restatement of the input staging and action post-processing that sit either
side of `infer_action`. Only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline may import this module; the product path never does.

Restated from:
  * frame resize  — `cv2.resize(image, image_size, interpolation=cv2.INTER_LANCZOS4)` at
    Vlaser_VLA/Simpler/src/agent/env_adapter/simpler.py:80-84. OpenCV is a third-party dependency that is not under
    /root/reference (requirements.txt:154-155 pin opencv-python 4.8.0.74 / headless 4.11.0.86); its published 8-bit
    Lanczos4 algorithm (imgproc/resize.cpp: `resizeGeneric_<HResizeLanczos4<uchar,int,short>, VResizeLanczos4<uchar,int,
    short,FixedPtCast<int,uchar,22>>>`, coefficients from `interpolateLanczos4`, scaled by 2^11 and rounded to int16,
    replicate border, result `(sum + 2^21) >> 22` saturated to u8) is restated here in numpy integer arithmetic and
    pinned against the cv2 build of this container (4.13.0) in tests/test_staging.py and by tests/golden/staging_*.npz.
  * normalisation — `process_images` / `rescale` / `normalize`, Vlaser_VLA/Simpler/src/model/vla/processing.py:27-63,
    mean/std at :303-304.
  * proprio / action (de)normalisation — Vlaser_VLA/Simpler/src/agent/env_adapter/base.py:8-49.
  * postprocess — simpler.py:136-177 (`SimplerAdapter.postprocess`), gripper rules :217-221 (Bridge) and :259-286
    (EDR sticky gripper), euler2axangle — Vlaser_VLA/Simpler/src/utils/geometry.py:261-434 (sxyz).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS
IMAGENET_MEAN = (0.485, 0.456, 0.406)  # processing.py:303
IMAGENET_STD = (0.229, 0.224, 0.225)   # processing.py:304


# ------------------------------------------------------------------------------------------------ resize
def lanczos4_coeffs(x: np.float32) -> np.ndarray:
    """OpenCV `interpolateLanczos4(float x, float* coeffs)`: the eight taps for fractional offset x, float32."""
    s45 = 0.70710678118654752440084436210485
    cs = ((1, 0), (-s45, -s45), (0, 1), (s45, -s45), (-1, 0), (s45, s45), (0, -1), (-s45, s45))
    x = np.float32(x)
    coeffs = np.zeros(8, np.float32)
    y0 = -float(np.float32(x + np.float32(3))) * math.pi * 0.25   # (x+3) is a float expression in OpenCV
    s0, c0 = math.sin(y0), math.cos(y0)
    total = np.float32(0)
    for i in range(8):
        y0_ = np.float32(np.float32(x + np.float32(3)) - np.float32(i))
        if abs(float(y0_)) >= 1e-6:
            y = -float(y0_) * math.pi * 0.25
            coeffs[i] = np.float32((cs[i][0] * s0 + cs[i][1] * c0) / (y * y))
        else:
            coeffs[i] = np.float32(1e30)
        total = np.float32(total + coeffs[i])
    inv = np.float32(np.float32(1.0) / total)
    return (coeffs * inv).astype(np.float32)


def _axis_table(src: int, dst: int) -> Tuple[np.ndarray, np.ndarray]:
    """Per destination index: first source tap index (may be negative / past the end: taps clamp) and the eight
    int16 weights `saturate_cast<short>(coeff * 2048)` (round half to even)."""
    scale = float(src) / float(dst)
    ofs = np.zeros(dst, np.int64)
    w = np.zeros((dst, 8), np.int64)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(math.floor(float(f)))
        f = np.float32(f - np.float32(s))
        c = lanczos4_coeffs(f)
        ofs[d] = s - 3
        w[d] = np.clip(np.rint((c * np.float32(COEF_SCALE)).astype(np.float32)), -32768, 32767).astype(np.int64)
    return ofs, w


def resize_lanczos4_u8(img: np.ndarray, out_hw: Tuple[int, int]) -> np.ndarray:
    """img u8 [H, W, C] -> u8 [out_h, out_w, C]; bit-exact restatement of cv2.resize(..., INTER_LANCZOS4)."""
    assert img.dtype == np.uint8 and img.ndim == 3
    H, W, _ = img.shape
    oh, ow = out_hw
    if (oh, ow) == (H, W):
        return img.copy()
    xo, xw = _axis_table(W, ow)
    yo, yw = _axis_table(H, oh)
    src = img.astype(np.int64)
    # horizontal pass on every source row (int accumulators, no rounding)
    xi = np.clip(xo[:, None] + np.arange(8)[None, :], 0, W - 1)          # [ow, 8]
    hor = np.einsum("hwkc,wk->hwc", src[:, xi, :], xw)                    # [H, ow, C]
    yi = np.clip(yo[:, None] + np.arange(8)[None, :], 0, H - 1)          # [oh, 8]
    ver = np.einsum("hkwc,hk->hwc", hor[yi, :, :], yw)                    # [oh, ow, C]
    out = (ver + (1 << (2 * COEF_BITS - 1))) >> (2 * COEF_BITS)
    return np.clip(out, 0, 255).astype(np.uint8)


def center_crop(img: np.ndarray) -> np.ndarray:
    """simpler.py:72-79 (the CENTER_CROP branch)."""
    h, w = img.shape[:2]
    c = min(h, w)
    sx, sy = (w - c) // 2, (h - c) // 2
    return img[sy:sy + c, sx:sx + c]


# ------------------------------------------------------------------------------------------------ normalise
def normalize_frames(frames_u8_nchw: np.ndarray) -> np.ndarray:
    """processing.py:27-63: `image * (1/255)` then `(image - mean) / std`, all float32. [n,3,H,W] u8 -> f32."""
    x = frames_u8_nchw.astype(np.float32) * np.float32(1 / 255.0)
    mean = np.asarray(IMAGENET_MEAN, np.float32)[None, :, None, None]
    std = np.asarray(IMAGENET_STD, np.float32)[None, :, None, None]
    return ((x - mean) / std).astype(np.float32)


def stage_frames(frames_hwc_u8: np.ndarray, px: int, crop: bool = False) -> np.ndarray:
    """[n, H, W, 3] u8 camera frames -> [n, 3, px, px] float32 pixel_values (simpler.py:70-91 + processing.py:51-63)."""
    out = []
    for f in frames_hwc_u8:
        if crop:
            f = np.ascontiguousarray(center_crop(f))
        out.append(resize_lanczos4_u8(f, (px, px)).transpose(2, 0, 1))
    return normalize_frames(np.stack(out))


def normalize_bound(data, lo, hi, clip_min=-1.0, clip_max=1.0, eps=1e-8):
    """base.py:8-19."""
    return np.clip(2 * (data - lo) / (hi - lo + eps) - 1, clip_min, clip_max)


def normalize_gaussian(data, mean, std, eps=1e-8):
    """base.py:33-40."""
    return (data - mean) / (std + eps)


# ------------------------------------------------------------------------------------------------ postprocess
def euler2axangle_sxyz(ai: float, aj: float, ak: float) -> Tuple[np.ndarray, float]:
    """geometry.py:261-434 for axes 'sxyz' (i=1, j=2, k=3, no parity / repetition / frame swap)."""
    ai, aj, ak = ai / 2.0, aj / 2.0, ak / 2.0
    ci, si, cj, sj, ck, sk = math.cos(ai), math.sin(ai), math.cos(aj), math.sin(aj), math.cos(ak), math.sin(ak)
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    q = np.array([cj * cc + sj * ss, cj * sc - sj * cs, cj * ss + sj * cc, cj * cs - sj * sc])
    eps = np.finfo(np.float64).eps
    Nq = float(np.sum(q ** 2))
    if not np.isfinite(Nq):
        return np.array([1.0, 0, 0]), float("nan")
    if Nq < eps ** 2:
        return np.array([1.0, 0, 0]), 0.0
    if Nq != 1:
        q = q / math.sqrt(Nq)
    xyz = q[1:]
    len2 = float(np.sum(xyz ** 2))
    if len2 < (eps * 3) ** 2:
        return np.array([1.0, 0, 0]), 0.0
    theta = 2 * math.acos(max(min(q[0], 1), -1))
    return xyz / math.sqrt(len2), theta


class StickyGripper:
    """EDRSimplerAdapter gripper state, simpler.py:228-286 (one instance per environment)."""

    def __init__(self, num_repeat: int = 15):
        self.num_repeat = num_repeat
        self.reset()

    def reset(self):
        self.on, self.repeat, self.value = False, 0, 0.0

    def step(self, action: float) -> float:
        action = (action * 2) - 1
        rel = -action
        if abs(rel) > 0.5 and self.on is False:
            self.on, self.value = True, rel
        if self.on:
            self.repeat += 1
            rel = self.value
        if self.repeat == self.num_repeat:
            self.on, self.repeat, self.value = False, 0, 0.0
        return rel


def postprocess(actions: np.ndarray, stats: Dict[str, np.ndarray], norm: str = "bound", gripper: str = "bridge",
                sticky: Optional[StickyGripper] = None) -> np.ndarray:
    """simpler.py:136-177 for ONE environment: actions [T, 7] normalised -> [T, 7] (xyz, axis*angle, gripper)."""
    actions = np.asarray(actions)
    if norm == "bound":
        lo, hi = np.asarray(stats["p01"])[:-1], np.asarray(stats["p99"])[:-1]
        raw = (actions[:, :-1] - (-1)) / 2 * (hi - lo) + lo                   # base.py:21-31
    else:
        raw = actions[:, :-1] * (np.asarray(stats["std"])[:-1] + 1e-8) + np.asarray(stats["mean"])[:-1]  # base.py:42-49
    raw = np.concatenate([raw, actions[:, -1:]], axis=1)
    out = np.zeros((len(raw), 7))
    for i, r in enumerate(raw):
        ax, ang = euler2axangle_sxyz(float(r[3]), float(r[4]), float(r[5]))
        if gripper == "bridge":
            g = 2.0 * (r[-1] > 0.5) - 1.0                                      # simpler.py:217-221
        else:
            g = sticky.step(float(r[-1]))                                      # simpler.py:259-286
        out[i] = np.concatenate([r[:3], ax * ang, [g]])
    return out
