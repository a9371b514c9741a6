"""Golden vectors of the flow-matching training forward (`PiZero.forward`, pizero_internvl.py:1064-1197) from the
UNMODIFIED reference (oracle/ref_harness.py), build container only:

    python tests/golden/make_flow_golden.py

Same conventions as make_golden.py: weights are re-generated from the seed, inputs are stored.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from vlaser_b200.config import tiny  # noqa: E402
from vlaser_b200.weights import random_state_dict  # noqa: E402
from vlaser_b200.synth import make_inputs, weights_checksum  # noqa: E402

CASES = {
    # name: (cfg, batch, weight seed, input seed, text lengths, flow_sig_min)
    "tiny_flow_bridge": (tiny(), 3, 0, 11, [5, 9, 14], 0.001),
    "tiny_flow_robotwin": (tiny().with_(cond_steps=3, action_dim=14, proprio_dim=14, max_image_text_tokens=224,
                                        time_max_period=100.0), 2, 2, 13, [6, 12], 0.01),
}


def flow_targets(cfg, B, seed):
    """Target chunks x1 in [-1, 1] and flow times in (0, 1) for a case (CPU generator, deterministic)."""
    g = torch.Generator().manual_seed(seed)
    actions = torch.rand((B, cfg.num_action_tokens, cfg.action_dim), generator=g) * 2 - 1
    t = torch.rand((B,), generator=g)
    return actions, t


def main():
    from oracle import ref_harness as R
    assert R.available(), "needs /root/reference"
    for name, (cfg, B, wseed, iseed, tlens, sig) in CASES.items():
        sd = random_state_dict(cfg, wseed)
        inp = make_inputs(cfg, B, iseed, tlens)
        actions, t = flow_targets(cfg, B, 100 + iseed)
        loss, v = R.run_reference_flow_loss(cfg, sd, inp["input_ids"], inp["pixel_values"], inp["proprios"], actions,
                                            inp["noise"], t, sig)
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(
            path, cfg_json=np.frombuffer(repr(cfg.to_dict()).encode(), dtype=np.uint8), weight_seed=np.int64(wseed),
            weights_checksum=np.float64(weights_checksum(sd)), images_u8=inp["images_u8"].numpy(),
            input_ids=inp["input_ids"].numpy(), proprios=inp["proprios"].numpy(), noise=inp["noise"].numpy(),
            actions=actions.numpy(), t=t.numpy(), flow_sig_min=np.float64(sig), loss=loss.numpy(), v_psi=v.numpy())
        print(name, "->", path, os.path.getsize(path) // 1024, "KiB", "loss=", float(loss))


if __name__ == "__main__":
    main()
