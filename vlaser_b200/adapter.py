"""Device-side mirrors of the two evaluator steps either side of `infer_action`, batched
over environments. Both are thin ctypes bindings of CUDA kernels in csrc/staging.cu; there is no CPU fallback.

  FrameStager          — `SimplerAdapter.preprocess` image branch: optional centre crop, `cv2.resize(.., INTER_LANCZOS4)`
                         (Vlaser_VLA/Simpler/src/agent/env_adapter/simpler.py:70-91) + `process_images`
                         (src/model/vla/processing.py:51-63). uint8 camera frames in, `pixel_values` out, so the
                         host->device copy of a step carries u8 frames instead of bf16 tensors and the dense attention
                         mask is never built (valid lengths are derived from `input_ids` on the device).
  HistoryStacker       — the camera-history branch of `SimplerAdapter.preprocess` (simpler.py:92-104, `num_images > 1`):
                         every environment's previous frame is kept on the device and stacked in front of the
                         current one ([n_env, num_images, ...]); the first call after a reset repeats the current frame.
  Normalizer           — `BaseEnvAdapter.normalize_* / denormalize_*` (base.py:8-49) for a batch on the device: the proprio
                         vector of `SimplerAdapter.preprocess` (simpler.py:107-121), or any chunk of joint targets.
  ActionPostprocessor  — `SimplerAdapter.postprocess` (simpler.py:136-177) with the Bridge / EDR gripper rules
                         (:217-221, :259-286) for every environment of the batch at once (the reference handles only
                         `actions[0]`, eval.py:139); the sticky-gripper state lives on the device between calls.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import torch

from . import _lib


class FrameStager:
    def __init__(self, image_size: int, center_crop: bool = False, dtype: torch.dtype = torch.bfloat16,
                 layout: str = "hwc", mean: Optional[Sequence[float]] = None, std: Optional[Sequence[float]] = None):
        if layout not in ("hwc", "chw"):
            raise ValueError("layout must be 'hwc' or 'chw'")
        if dtype not in (torch.bfloat16, torch.float32):
            raise ValueError("pixel_values dtype must be bfloat16 or float32")
        self._lib = _lib.load()
        self.image_size, self.center_crop, self.dtype, self.layout = int(image_size), bool(center_crop), dtype, layout
        self._mean = (C.c_float * 3)(*mean) if mean is not None else None
        self._std = (C.c_float * 3)(*std) if std is not None else None

    def __call__(self, frames: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """frames: uint8 CUDA tensor [..., H, W, 3] ('hwc') or [..., 3, H, W] ('chw'); leading dims are flattened
        (B, or B x num_images as processing.py:60 `flatten(0, 1)`). Returns [n, 3, px, px] on the same device."""
        if not frames.is_cuda or frames.dtype != torch.uint8:
            raise RuntimeError("FrameStager needs a uint8 CUDA tensor (no CPU fallback)")
        frames = frames.contiguous()
        if self.layout == "hwc":
            H, W, ch = frames.shape[-3:]
        else:
            ch, H, W = frames.shape[-3:]
        if ch != 3:
            raise ValueError("expected 3 colour channels")
        n = frames.numel() // (3 * H * W)
        px = self.image_size
        if out is None:
            out = torch.empty((n, 3, px, px), dtype=self.dtype, device=frames.device)
        with torch.cuda.device(frames.device):
            rc = self._lib.vlaser_op_stage_frames(
                C.c_void_p(frames.data_ptr()), n, H, W, 0 if self.layout == "hwc" else 1, int(self.center_crop),
                C.c_void_p(out.data_ptr()), _lib.BF16 if out.dtype == torch.bfloat16 else _lib.F32, px,
                self._mean, self._std, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "stage_frames")
        return out


class HistoryStacker:
    """`history_image` of SimplerAdapter (simpler.py:60-104) for n_env environments, on the device. frames: uint8 CUDA
    tensor [n_env, ...] of the current step (any trailing layout). Returns [n_env, num_images, ...]: the num_images - 1
    previous frames then the current one; after `reset` the current frame fills the whole window, as the reference does
    for its two-frame window."""

    def __init__(self, num_images: int):
        if num_images < 1:
            raise ValueError("num_images >= 1")
        self.num_images = int(num_images)
        self._hist: Optional[torch.Tensor] = None   # [n_env, num_images - 1, ...]
        self._fresh: Optional[torch.Tensor] = None  # [n_env] bool: no history yet

    def reset(self, env_ids: Optional[Sequence[int]] = None) -> None:
        if env_ids is None or self._fresh is None:
            self._hist = self._fresh = None
        else:
            self._fresh[torch.as_tensor(list(env_ids), device=self._fresh.device)] = True

    def __call__(self, frames: torch.Tensor) -> torch.Tensor:
        if not frames.is_cuda or frames.dtype != torch.uint8:
            raise RuntimeError("HistoryStacker needs a uint8 CUDA tensor (no CPU fallback)")
        cur = frames.unsqueeze(1)
        if self.num_images == 1:
            return cur
        n = frames.shape[0]
        if self._hist is None or self._hist.shape[0] != n or self._hist.shape[2:] != frames.shape[1:]:
            self._hist = cur.expand(n, self.num_images - 1, *frames.shape[1:]).clone()
            self._fresh = torch.zeros(n, dtype=torch.bool, device=frames.device)
        elif bool(self._fresh.any()):
            idx = self._fresh.nonzero().flatten()
            self._hist[idx] = cur[idx].expand(-1, self.num_images - 1, *frames.shape[1:])
            self._fresh.zero_()
        out = torch.cat([self._hist, cur], dim=1)
        self._hist = out[:, 1:].clone()
        return out


class Normalizer:
    """stats: one entry of the reference's dataset statistics json ("proprio" or "action": keys p01 / p99 / mean / std).
    `normalize(x)`: [..., dim] CUDA tensor (f32 / bf16 / f64) -> normalised, dtype `out_dtype`; `denormalize` the inverse."""

    _CODES = {torch.float32: 0, torch.bfloat16: 1, torch.float64: 2}

    def __init__(self, stats: Dict[str, Sequence[float]], normalization: str = "bound", device=None):
        if normalization not in ("bound", "gaussian"):
            raise ValueError("normalization must be 'bound' or 'gaussian'")
        if not torch.cuda.is_available():
            raise RuntimeError("vlaser_b200 needs a CUDA device (no CPU fallback)")
        self._lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        a, b = ("p01", "p99") if normalization == "bound" else ("mean", "std")
        self._a = torch.tensor(list(stats[a]), dtype=torch.float64, device=self.device)
        self._b = torch.tensor(list(stats[b]), dtype=torch.float64, device=self.device)
        self.dim = int(self._a.numel())
        self.norm = 0 if normalization == "bound" else 1

    def _run(self, x: torch.Tensor, out_dtype: torch.dtype, inverse: int) -> torch.Tensor:
        if not x.is_cuda or x.dtype not in self._CODES or out_dtype not in self._CODES:
            raise RuntimeError("Normalizer needs an f32 / bf16 / f64 CUDA tensor (no CPU fallback)")
        if x.shape[-1] != self.dim:
            raise ValueError(f"last dimension must be {self.dim}")
        x = x.contiguous()
        out = torch.empty(x.shape, dtype=out_dtype, device=x.device)
        with torch.cuda.device(x.device):
            rc = self._lib.vlaser_op_normalize(
                C.c_void_p(x.data_ptr()), self._CODES[x.dtype], C.c_void_p(out.data_ptr()), self._CODES[out_dtype],
                x.numel() // self.dim, self.dim, self.norm, inverse, C.c_void_p(self._a.data_ptr()),
                C.c_void_p(self._b.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "normalize")
        return out

    def normalize(self, x: torch.Tensor, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
        return self._run(x, out_dtype, 0)

    def denormalize(self, x: torch.Tensor, out_dtype: torch.dtype = torch.float64) -> torch.Tensor:
        return self._run(x, out_dtype, 1)


class ActionPostprocessor:
    def __init__(self, action_stats: Dict[str, Sequence[float]], normalization: str = "bound", gripper: str = "bridge",
                 n_env: int = 1, device=None, sticky_gripper_num_repeat: int = 15):
        """action_stats: the `"action"` entry of the reference's dataset statistics json (keys p01/p99/mean/std);
        normalization: `action_normalization_type` (simpler.py:34-37); gripper: 'bridge' (BridgeSimplerAdapter) or
        'sticky' (EDRSimplerAdapter)."""
        if normalization not in ("bound", "gaussian"):
            raise ValueError("normalization must be 'bound' or 'gaussian'")
        if gripper not in ("bridge", "sticky"):
            raise ValueError("gripper must be 'bridge' or 'sticky'")
        if not torch.cuda.is_available():
            raise RuntimeError("vlaser_b200 needs a CUDA device (no CPU fallback)")
        self._lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        a, b = ("p01", "p99") if normalization == "bound" else ("mean", "std")
        self._a = torch.tensor(list(action_stats[a]), dtype=torch.float64, device=self.device)
        self._b = torch.tensor(list(action_stats[b]), dtype=torch.float64, device=self.device)
        self.norm = 0 if normalization == "bound" else 1
        self.gripper = 0 if gripper == "bridge" else 1
        self.repeat = int(sticky_gripper_num_repeat)
        self.n_env = int(n_env)
        self._state = torch.zeros((self.n_env, 3), dtype=torch.float64, device=self.device)

    def reset(self, env_ids: Optional[Sequence[int]] = None) -> None:
        """EDRSimplerAdapter.reset (simpler.py:235-240) for all / the given environments."""
        if env_ids is None:
            self._state.zero_()
        else:
            self._state[torch.as_tensor(list(env_ids), device=self.device)] = 0

    def __call__(self, actions: torch.Tensor) -> torch.Tensor:
        """actions: [n_env, T, 7] normalised (bf16 / f32, CUDA) -> float64 [n_env, T, 7] environment actions."""
        if not actions.is_cuda or actions.dtype not in (torch.bfloat16, torch.float32):
            raise RuntimeError("ActionPostprocessor needs a bf16 / f32 CUDA tensor (no CPU fallback)")
        n_env, T, A = actions.shape
        if n_env != self.n_env:
            raise ValueError(f"expected {self.n_env} environments, got {n_env}")
        actions = actions.contiguous()
        out = torch.empty((n_env, T, 7), dtype=torch.float64, device=actions.device)
        with torch.cuda.device(actions.device):
            rc = self._lib.vlaser_op_postprocess_actions(
                C.c_void_p(actions.data_ptr()), _lib.BF16 if actions.dtype == torch.bfloat16 else _lib.F32, n_env, T, A,
                self.norm, C.c_void_p(self._a.data_ptr()), C.c_void_p(self._b.data_ptr()), self.gripper, self.repeat,
                C.c_void_p(self._state.data_ptr()), C.c_void_p(out.data_ptr()),
                C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "postprocess_actions")
        return out
