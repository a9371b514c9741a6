"""Synthetic observations of the shape the reference's env adapter produces
(`Vlaser_VLA/Simpler/src/agent/env_adapter/simpler.py:65-134`, `src/model/vla/processing.py:27-63,
330-366`): uint8 frames -> `(x/255-mean)/std`, prompt template = 10 ids, n_img*T `<IMG_CONTEXT>` ids,
7+len(prompt) ids, right-padded with `pad_token_id` to `max_image_text_tokens`.

Used by bench.py, the golden-vector generator and the tests (SURVEY.md §8d "Synthetic inputs").
There is no tokenizer on the GPU box, so prompt token COUNTS are taken from the reference's
standard task strings (tokenised once with the vendored Qwen2 tokenizer) and the ids are random.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .config import VlaserConfig

IMAGE_MEAN = (0.485, 0.456, 0.406)   # processing.py:303-304
IMAGE_STD = (0.229, 0.224, 0.225)
TEMPLATE_PREFIX = 10                 # "<|im_start|>system\nNone<|im_end|>\n<|im_start|>user\n<img>"
TEMPLATE_SUFFIX = 7                  # "</img>\n" ... "<|im_end|>\n<|im_start|>assistant\n"
# token counts of the SimplerEnv task strings
BRIDGE_PROMPT_TOKENS = (4, 5, 6, 8)  # put carrot on plate / put eggplant in basket / put the spoon on the towel / stack ...
FRACTAL_PROMPT_TOKENS = (3, 2, 2, 2, 5)  # pick coke can / move near / open drawer / close drawer / place apple into top drawer


def process_images(images_u8: torch.Tensor) -> torch.Tensor:
    """processing.py:51-63: [B,n_img,3,px,px] uint8 -> [B*n_img,3,px,px] fp32 normalised."""
    mean = torch.tensor(IMAGE_MEAN)[None, None, :, None, None]
    std = torch.tensor(IMAGE_STD)[None, None, :, None, None]
    x = images_u8.to(torch.float32) * (1.0 / 255.0)
    x = (x - mean) / std
    return x.flatten(0, 1)


def make_inputs(cfg: VlaserConfig, batch: int, seed: int = 1,
                prompt_tokens: Optional[List[int]] = None,
                tiles_per_sample: Optional[List[int]] = None) -> Dict[str, torch.Tensor]:
    """Returns images_u8, pixel_values (fp32), input_ids (i64 [B,S]), attention_mask, proprios
    (fp32 [B,1,Dp]), noise (fp32 [B,H,A]), valid_len (i32 [B]).

    `tiles_per_sample` (dynamic tiles): sample b holds tiles_per_sample[b] image tiles (that many
    times T `<IMG_CONTEXT>` ids); the tiles of the batch are concatenated in order, and their total must be
    batch * cfg.num_images -- the layout the reference's masked scatter consumes (pizero_internvl.py:757-791)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    S, T, n_img = cfg.max_image_text_tokens, cfg.image_tokens, cfg.num_images
    if prompt_tokens is None:
        prompt_tokens = list(BRIDGE_PROMPT_TOKENS)
    images = torch.randint(0, 256, (batch, n_img, 3, cfg.image_size, cfg.image_size), dtype=torch.uint8, generator=g)
    ids = torch.full((batch, S), cfg.pad_token_id, dtype=torch.int64)
    lo_special = min(cfg.image_token_index, cfg.pad_token_id)
    valid = []
    if tiles_per_sample is not None:
        assert len(tiles_per_sample) == batch and sum(tiles_per_sample) == batch * n_img, "tile counts must sum to batch * num_images"
    for b in range(batch):
        n_txt = prompt_tokens[b % len(prompt_tokens)]
        n_pre, n_post = TEMPLATE_PREFIX, TEMPLATE_SUFFIX + n_txt
        n_it = (tiles_per_sample[b] if tiles_per_sample is not None else n_img) * T
        total = n_pre + n_it + n_post
        assert total <= S, f"prefix of {total} tokens does not fit max_image_text_tokens={S}"
        txt = torch.randint(0, lo_special, (n_pre + n_post,), generator=g)
        ids[b, :n_pre] = txt[:n_pre]
        ids[b, n_pre:n_pre + n_it] = cfg.image_token_index
        ids[b, n_pre + n_it:total] = txt[n_pre:]
        valid.append(total)
    # Values are rounded to the bf16 grid (and kept as fp32): the bf16 device path and the fp32 oracle
    # then see IDENTICAL inputs, so parity measures compute error only (SURVEY.md §7 hard parts).
    r16 = lambda t: t.to(torch.bfloat16).to(torch.float32)
    proprios = r16(torch.rand(batch, 1, cfg.proprio_dim, generator=g) * 2 - 1)
    noise = r16(torch.randn(batch, cfg.num_action_tokens, cfg.action_dim, generator=g))
    return {
        "images_u8": images,
        "pixel_values": r16(process_images(images)),
        "input_ids": ids,
        "attention_mask": (ids != cfg.pad_token_id).to(torch.int64),
        "proprios": proprios,
        "noise": noise,
        "valid_len": torch.tensor(valid, dtype=torch.int32),
    }


def weights_checksum(sd: Dict[str, torch.Tensor]) -> float:
    """Order-dependent fp64 checksum to detect RNG drift between containers."""
    acc = 0.0
    for i, (k, v) in enumerate(sd.items()):
        acc += float(v.double().sum()) * (1.0 + (i % 7)) + float(v.double().abs().sum())
    return acc


def pixel_values_from_u8(images_u8: torch.Tensor) -> torch.Tensor:
    """uint8 frames -> normalised pixel values on the bf16 grid (fp32 tensor), as `make_inputs` does."""
    return process_images(images_u8).to(torch.bfloat16).to(torch.float32)
