"""Loader for the committed golden vectors (tests/golden/*.npz, produced by make_golden.py from the
unmodified reference)."""
from __future__ import annotations

import ast
import os

import numpy as np
import torch

from vlaser_b200.config import VlaserConfig
from vlaser_b200.synth import pixel_values_from_u8, weights_checksum
from vlaser_b200.weights import random_state_dict

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["tiny_bridge", "tiny_fractal", "tiny_robotwin"]
DYNAMIC_TILE_CASES = ["tiny_dynamic_tiles"]  # samples own different numbers of image tiles


def load_case(name: str):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    cfg = VlaserConfig(**ast.literal_eval(bytes(z["cfg_json"]).decode()))
    sd = random_state_dict(cfg, int(z["weight_seed"]))
    assert abs(weights_checksum(sd) - float(z["weights_checksum"])) < 1e-6 * abs(float(z["weights_checksum"])), \
        "weight generator drifted from the one that produced the golden vectors"
    images = torch.from_numpy(z["images_u8"])
    inp = {
        "input_ids": torch.from_numpy(z["input_ids"]),
        "pixel_values": pixel_values_from_u8(images),
        "proprios": torch.from_numpy(z["proprios"]),
        "noise": torch.from_numpy(z["noise"]),
    }
    golden = {k: torch.from_numpy(z[k]) for k in z.files if k == "actions" or k.startswith("trace_")}
    return cfg, sd, inp, golden
