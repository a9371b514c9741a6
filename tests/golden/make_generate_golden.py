"""Golden vectors of the chat-VLM greedy decode (InternVLChatModel.generate -> HF Qwen2 generate, do_sample=False) from the
reference's modules and the UNMODIFIED HF Qwen2ForCausalLM (oracle/ref_harness.py::run_reference_generate), build
container only:

    python tests/golden/make_generate_golden.py

Weights: random_state_dict(cfg, seed) + text_head_state_dict(cfg, seed, tied); inputs are stored.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from vlaser_b200.config import tiny  # noqa: E402
from vlaser_b200.weights import random_state_dict, text_head_state_dict  # noqa: E402
from vlaser_b200.synth import make_inputs, weights_checksum  # noqa: E402

CASES = {
    # name: (cfg, batch, weight seed, input seed, text lengths, tiles per sample, tied head, new tokens)
    "tiny_generate": (tiny(), 3, 0, 11, [5, 9, 14], None, False, 8),
    "tiny_generate_tied_tiles": (tiny().with_(cond_steps=2, max_image_text_tokens=224), 4, 3, 11, None, [1, 3, 2, 2], True, 6),
}


def main():
    from oracle import ref_harness as R
    assert R.available(), "needs /root/reference"
    for name, (cfg, B, wseed, iseed, tlens, tiles, tied, n_new) in CASES.items():
        sd = random_state_dict(cfg, wseed)
        sd.update(text_head_state_dict(cfg, wseed, tied))
        inp = make_inputs(cfg, B, iseed, tlens, tiles_per_sample=tiles)
        tokens, first = R.run_reference_generate(cfg, sd, inp["input_ids"], inp["pixel_values"], n_new)
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(
            path, cfg_json=np.frombuffer(repr(cfg.to_dict()).encode(), dtype=np.uint8), weight_seed=np.int64(wseed),
            weights_checksum=np.float64(weights_checksum(sd)), tied=np.int64(tied), images_u8=inp["images_u8"].numpy(),
            input_ids=inp["input_ids"].numpy(), tokens=tokens.numpy(), first_logits=first.numpy())
        print(name, "->", path, os.path.getsize(path) // 1024, "KiB", tokens.tolist())


if __name__ == "__main__":
    main()
