"""Generates tests/golden/staging_*.npz by running the UNMODIFIED reference adapter code (and the cv2 build of this
container for the resize the reference delegates to OpenCV) in the build container.

    python tests/golden/make_staging_golden.py

What is executed from /root/reference (import stubs only for tensorflow / simpler_env, which are absent here):
  * `SimplerAdapter.postprocess`, `BridgeSimplerAdapter.postprocess_gripper`, `EDRSimplerAdapter.postprocess_gripper`
    (+ `reset`) — Vlaser_VLA/Simpler/src/agent/env_adapter/simpler.py:136-177, 217-221, 228-286
  * `BaseEnvAdapter.normalize_bound / normalize_gaussian` — env_adapter/base.py:8-40
  * `process_images` — src/model/vla/processing.py:51-63
  * `cv2.resize(..., INTER_LANCZOS4)` as called at simpler.py:80-84
The dataset statistics used are the reference's own `config/{bridge,fractal}_statistics.json` (values stored in the
fixture so the tests do not need the reference).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))


def load_adapter_module():
    from oracle import ref_harness as R
    assert R.available(), "needs /root/reference"
    R.load_reference()  # sys.path + the stubs the model code needs
    R._stub("tensorflow")
    R._stub("simpler_env")
    R._stub("simpler_env.utils")
    R._stub("simpler_env.utils.env")
    R._stub("simpler_env.utils.env.observation_utils", get_image_from_maniskill2_obs_dict=lambda env, obs: obs["image"])
    import src.agent.env_adapter.simpler as sim
    import src.model.vla.processing as proc
    return sim, proc, R.SIMPLER


def smooth_frame(h, w, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([128 + 100 * np.sin(xx / 17.0 + yy / 23.0 + seed), xx * 255.0 / (w - 1), yy * 255.0 / (h - 1)], -1)
    return np.clip(base + rng.normal(0, 25, base.shape), 0, 255).astype(np.uint8)


def main():
    import cv2
    sim, proc, simpler_root = load_adapter_module()
    rng = np.random.default_rng(7)

    # ---- frames: resize + normalise
    frames = {
        "a": np.stack([smooth_frame(60, 80, 1), rng.integers(0, 256, (60, 80, 3), dtype=np.uint8)]),       # upscale
        "b": np.stack([smooth_frame(180, 240, 2)]),                                                          # downscale
        "c": np.stack([rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)]),                                  # identity size
    }
    targets = {"a": 112, "b": 128, "c": 64}
    out = {"cv2_version": np.array(cv2.__version__)}
    mean = torch.tensor([0.485, 0.456, 0.406])
    std = torch.tensor([0.229, 0.224, 0.225])
    for k, fr in frames.items():
        px = targets[k]
        res = np.stack([cv2.resize(f, (px, px), interpolation=cv2.INTER_LANCZOS4) for f in fr])  # simpler.py:80-84
        images = torch.as_tensor(res, dtype=torch.uint8).permute(0, 3, 1, 2)[:, None]            # [n,1,3,H,W] (:88-90)
        pv = proc.process_images(images, 1 / 255.0, mean, std)                                   # processing.py:51-63
        out[f"frames_{k}"] = fr
        out[f"resized_{k}"] = res
        out[f"px_{k}"] = np.array(px)
        if k != "a":
            out[f"pixel_values_{k}"] = pv.numpy().astype(np.float32)
        else:
            out[f"pixel_values_{k}_bf16bits"] = pv.to(torch.bfloat16).view(torch.int16).numpy()
    # centre crop branch (simpler.py:72-79) on frame b
    f = frames["b"][0]
    h, w = f.shape[:2]
    c = min(h, w)
    crop = f[(h - c) // 2:(h - c) // 2 + c, (w - c) // 2:(w - c) // 2 + c]
    out["resized_b_crop"] = cv2.resize(crop, (128, 128), interpolation=cv2.INTER_LANCZOS4)[None]
    np.savez_compressed(os.path.join(OUT, "staging_frames.npz"), **out)

    # ---- proprio normalisation + action post-processing
    post = {}
    for name, cls, norm in (("bridge", sim.BridgeSimplerAdapter, "bound"), ("fractal", sim.EDRSimplerAdapter, "gaussian"),
                            ("fractal_bound", sim.EDRSimplerAdapter, "bound")):
        stats_file = "bridge_statistics.json" if name == "bridge" else "fractal_statistics.json"
        with open(os.path.join(simpler_root, "config", stats_file)) as fh:
            stats = json.load(fh)
        ad = cls.__new__(cls)
        ad.dataset_statistics = stats
        ad.action_normalization_type = norm
        ad.proprio_normalization_type = norm
        ad.history_image = None
        if cls is sim.EDRSimplerAdapter:
            ad.sticky_gripper_num_repeat = 15
        if cls is sim.BridgeSimplerAdapter:
            ad.default_rot = np.array([[0, 0, 1.0], [0, 1.0, 0], [-1.0, 0, 0]])
        ad.reset()
        n_env, T = 3, 24                     # 6 chunks of 4 actions per environment, state carried across chunks
        acts = rng.uniform(-1.2, 1.2, (n_env, T, 7)).astype(np.float32)
        acts[:, :, -1] = rng.uniform(-0.2, 1.2, (n_env, T))  # gripper channel is not normalised (simpler.py:139)
        acts[0, 5, 3:6] = 0.0                # identity rotation branch of quat2axangle
        outs = []
        for e in range(n_env):
            ad.reset()
            chunks = [ad.postprocess(acts[e, t:t + 4]) for t in range(0, T, 4)]   # simpler.py:136-177, unmodified
            outs.append(np.concatenate(chunks, 0))
        post[f"{name}_actions"] = acts
        post[f"{name}_post"] = np.stack(outs)
        for grp in ("action", "proprio"):
            for key in ("mean", "std", "p01", "p99"):
                post[f"{name}_{grp}_{key}"] = np.asarray(stats[grp][key], np.float64)
        raw_prop = rng.normal(0, 0.5, (5, len(stats["proprio"]["mean"])))
        if norm == "bound":
            post[f"{name}_proprio_norm"] = ad.normalize_bound(raw_prop, np.array(stats["proprio"]["p01"]),
                                                              np.array(stats["proprio"]["p99"]), clip_min=-1, clip_max=1)
        else:
            post[f"{name}_proprio_norm"] = ad.normalize_gaussian(raw_prop, np.array(stats["proprio"]["mean"]),
                                                                 np.array(stats["proprio"]["std"]))
        post[f"{name}_proprio_raw"] = raw_prop
    np.savez_compressed(os.path.join(OUT, "staging_post.npz"), **post)
    for fn in ("staging_frames.npz", "staging_post.npz"):
        print(fn, os.path.getsize(os.path.join(OUT, fn)) // 1024, "KiB")


if __name__ == "__main__":
    main()
