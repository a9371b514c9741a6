"""Data-parallel sharding of the policy call across the GPUs of one box.

Observations are independent (no cross-sample op anywhere in `infer_action`), so rank r takes rows
[r*B/G, (r+1)*B/G) of every input, runs the full path on its own weight replica / KV cache / CUDA
graph, and the only data-path collective is ONE all-gather of the [B/G, horizon, A] action chunk
after the last Euler step (NCCL over NVLink on the box; gloo in the CPU tests). Noise is drawn for
the GLOBAL batch from a shared seed and sliced, so results do not depend on G.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced row range of `rank` (first `batch % world` ranks get one extra row)."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_inputs(inputs: Dict[str, torch.Tensor], num_images: int, world: int, rank: int) -> Dict[str, torch.Tensor]:
    """Row-shards every per-observation tensor; `pixel_values` has num_images rows per observation."""
    B = inputs["input_ids"].shape[0]
    lo, hi = shard_bounds(B, world, rank)
    out = {}
    for k, v in inputs.items():
        if not torch.is_tensor(v):
            out[k] = v
        elif k in ("pixel_values",):
            out[k] = v[lo * num_images: hi * num_images]
        elif k == "images_u8" or v.shape[0] == B:
            out[k] = v[lo:hi]
        else:
            out[k] = v
    return out


def global_noise(batch: int, H: int, A: int, seed: int, dtype=torch.float32) -> torch.Tensor:
    """The `torch.randn` draw of pizero_internvl.py:879-881 for the global batch (CPU generator so that
    every rank, and every G, sees the same rows)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randn(batch, H, A, generator=g).to(dtype)


def gather_actions(local_actions: torch.Tensor, batch: int, group=None) -> torch.Tensor:
    """All-gathers the per-rank action chunks into [B, horizon, A] (the single collective of the path)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_actions
    world = dist.get_world_size(group)
    sizes = [shard_bounds(batch, world, r) for r in range(world)]
    if len({hi - lo for lo, hi in sizes}) == 1:
        out = torch.empty((batch,) + tuple(local_actions.shape[1:]), dtype=local_actions.dtype,
                          device=local_actions.device)
        dist.all_gather_into_tensor(out, local_actions.contiguous(), group=group)
        return out
    # ragged split: still ONE fixed-size collective -- pad every shard to the largest one, drop the pad rows after
    rows = max(hi - lo for lo, hi in sizes)
    tail = tuple(local_actions.shape[1:])
    padded = local_actions.new_zeros((rows,) + tail)
    padded[: local_actions.shape[0]] = local_actions
    out = local_actions.new_empty((world * rows,) + tail)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * rows: r * rows + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)


def infer_action_dp(policy_fn: Callable[..., torch.Tensor], inputs: Dict[str, torch.Tensor], num_images: int,
                    group=None, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Runs `policy_fn(**local_inputs)` on this rank's shard of a GLOBAL batch and gathers the actions.
    `inputs` holds the global tensors (input_ids, pixel_values, proprios, ...); `noise` is the global draw."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = inputs["input_ids"].shape[0]
    if noise is not None:
        inputs = dict(inputs, noise=noise)
    local = shard_inputs(inputs, num_images, world, rank)
    out = policy_fn(**local)
    return gather_actions(out, B, group)
