"""Text side of the reference's `InternVLAProcessor` (Vlaser_VLA/Simpler/src/model/vla/processing.py:296-360): the chat
query the policy is prompted with and its tokenisation to the fixed `max_image_text_tokens` window. Host-side string /
tokenizer work only (the tokenizer itself is the caller's: a HuggingFace tokenizer object of the checkpoint)."""
from __future__ import annotations

import torch

IMG_CONTEXT = "<IMG_CONTEXT>"


def build_query(instruction: str, num_image_tokens: int) -> str:
    """processing.py:349-357 (non-debug branch): `num_image_tokens` = 64 (224 px) or 256 (448 px) per image x num_images."""
    return ("<|im_start|>system\nNone<|im_end|>\n<|im_start|>user\n<img>{}</img>\n{}<|im_end|>\n<|im_start|>assistant\n"
            .format(IMG_CONTEXT * num_image_tokens, instruction))


def encode_prompt(tokenizer, instruction: str, num_image_tokens: int, max_seq_len: int) -> torch.Tensor:
    """processing.py:359-362: tokenizer(query, max_length=max_seq_len, padding="max_length", truncation=True).
    Returns input_ids int64 [max_seq_len] (right-padded with the tokenizer's pad id)."""
    enc = tokenizer([build_query(instruction, num_image_tokens)], return_tensors="pt", max_length=max_seq_len,
                    padding="max_length", truncation=True)
    return enc["input_ids"][0].to(torch.int64)
