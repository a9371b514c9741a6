"""Loader for the committed golden vectors (tests/golden/*.npz, produced by make_golden.py from the
unmodified reference)."""
from __future__ import annotations

import ast
import os

import numpy as np
import torch

from vlaser_b200.config import VlaserConfig
from vlaser_b200.synth import pixel_values_from_u8, weights_checksum
from vlaser_b200.weights import random_state_dict, text_head_state_dict

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["tiny_bridge", "tiny_fractal", "tiny_robotwin"]
DYNAMIC_TILE_CASES = ["tiny_dynamic_tiles"]  # samples own different numbers of image tiles
FLOW_CASES = ["tiny_flow_bridge", "tiny_flow_robotwin"]  # training forward (make_flow_golden.py): loss, v_psi
GENERATE_CASES = ["tiny_generate", "tiny_generate_tied_tiles"]  # chat-VLM greedy decode (make_generate_golden.py)


def load_case(name: str):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    cfg = VlaserConfig(**ast.literal_eval(bytes(z["cfg_json"]).decode()))
    sd = random_state_dict(cfg, int(z["weight_seed"]))
    if "tokens" in z.files:  # chat-VLM decode: + the LLM's final norm and (untied) output head
        sd.update(text_head_state_dict(cfg, int(z["weight_seed"]), bool(int(z["tied"]))))
    assert abs(weights_checksum(sd) - float(z["weights_checksum"])) < 1e-6 * abs(float(z["weights_checksum"])), \
        "weight generator drifted from the one that produced the golden vectors"
    images = torch.from_numpy(z["images_u8"])
    if "tokens" in z.files:
        inp = {"input_ids": torch.from_numpy(z["input_ids"]), "pixel_values": pixel_values_from_u8(images)}
        return cfg, sd, inp, {"tokens": torch.from_numpy(z["tokens"]), "first_logits": torch.from_numpy(z["first_logits"])}
    inp = {
        "input_ids": torch.from_numpy(z["input_ids"]),
        "pixel_values": pixel_values_from_u8(images),
        "proprios": torch.from_numpy(z["proprios"]),
        "noise": torch.from_numpy(z["noise"]),
    }
    if "v_psi" in z.files:  # flow-matching training forward: `actions` is an INPUT (the target chunk x1)
        inp["actions"] = torch.from_numpy(z["actions"])
        inp["t"] = torch.from_numpy(z["t"])
        inp["flow_sig_min"] = float(z["flow_sig_min"])
        return cfg, sd, inp, {"loss": torch.from_numpy(z["loss"]), "v_psi": torch.from_numpy(z["v_psi"])}
    golden = {k: torch.from_numpy(z[k]) for k in z.files if k == "actions" or k.startswith("trace_")}
    return cfg, sd, inp, golden
