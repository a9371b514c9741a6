"""RoboTwin 2.0 evaluator plugin for the vlaser_b200 policy (BASELINE configs[3]: Vlaser-2B-VLA, 3 cameras, 14-DoF
bimanual joint chunks). Copy (or symlink) this directory to `policy/vlaser_b200/` of a RoboTwin checkout and run
`script/eval_policy.py` with `policy_name: vlaser_b200`: the evaluator imports `policy.<name>.deploy_policy` and calls

    get_model(usr_args)                      Vlaser_VLA/RoboTwin/script/eval_policy.py:38-46, 165
    eval(TASK_ENV, model, observation)       eval_policy.py:209, 292-295 (once per policy step)
    reset_model(model)                       eval_policy.py:210 (once per episode)

exactly as for the stock plugins (Vlaser_VLA/RoboTwin/policy/pi0/deploy_policy.py:14-54). `encode_obs` picks the three
camera frames and the 14-vector joint state the same way pi0's does (:14-22); actions go to `TASK_ENV.take_action`
(Vlaser_VLA/RoboTwin/envs/_base_task.py:1722, action_type 'qpos').

usr_args (the `deploy_policy.yml` of the plugin + command-line overrides):
    checkpoint            path of a torch checkpoint whose "model" (or top-level) state dict carries the reference keys,
                          or a state dict object; "random" builds random-init weights (smoke tests)
    dataset_statistics    path of the dataset-statistics json (or a dict) with "proprio" and "action" entries
                          (p01/p99 or mean/std, 14 values each)
    tokenizer             a HuggingFace tokenizer object, or a path for AutoTokenizer.from_pretrained, or a callable
                          instruction -> int64 ids [max_image_text_tokens]
    action_normalization_type / proprio_normalization_type   "bound" (default) | "gaussian"
    act_steps             actions of the predicted chunk that are executed before the next policy call (default: all)
    preset                vlaser_b200.config preset name (default "vlaser_2b_robotwin")
    device                CUDA device index (default 0); seed: noise seed (default 0)
"""
from __future__ import annotations

import json
import os
import sys
from typing import Callable, Optional

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(os.path.dirname(_HERE))
if _REPO not in sys.path:  # running from a copy inside a RoboTwin checkout: VLASER_B200_ROOT points at the repo
    sys.path.insert(0, os.environ.get("VLASER_B200_ROOT", _REPO))

from vlaser_b200.adapter import FrameStager, Normalizer  # noqa: E402
from vlaser_b200.config import PRESETS  # noqa: E402
from vlaser_b200.policy import PiZeroInference  # noqa: E402
from vlaser_b200.prompt import encode_prompt  # noqa: E402

CAMERAS = ("head_camera", "right_camera", "left_camera")  # order of pi0/deploy_policy.py:15-19


def encode_obs(observation):
    """pi0/deploy_policy.py:14-22: three RGB frames [H, W, 3] uint8 and the joint vector."""
    input_rgb_arr = [observation["observation"][cam]["rgb"] for cam in CAMERAS]
    input_state = observation["joint_action"]["vector"]
    return input_rgb_arr, input_state


class VlaserRoboTwinPolicy:
    def __init__(self, usr_args: dict):
        self.cfg = PRESETS[usr_args.get("preset", "vlaser_2b_robotwin")]()
        dev = int(usr_args.get("device", 0))
        torch.cuda.set_device(dev)
        self.device = torch.device("cuda", dev)
        self.policy = PiZeroInference(self.cfg, max_batch=1, device=dev)
        ck = usr_args.get("checkpoint", "random")
        if isinstance(ck, str) and ck == "random":
            from vlaser_b200.weights import random_state_dict
            sd = random_state_dict(self.cfg, int(usr_args.get("seed", 0)))
        elif isinstance(ck, str):
            blob = torch.load(ck, map_location="cpu", weights_only=True)
            sd = blob.get("model", blob) if isinstance(blob, dict) else blob
        else:
            sd = ck
        res = self.policy.load_state_dict(sd, strict=False)
        if res.missing_keys:
            raise RuntimeError(f"checkpoint misses {len(res.missing_keys)} tensors of the path, e.g. {res.missing_keys[:3]}")
        stats = usr_args["dataset_statistics"]
        if isinstance(stats, str):
            with open(stats) as f:
                stats = json.load(f)
        self.proprio_norm = Normalizer(stats["proprio"], usr_args.get("proprio_normalization_type", "bound"), self.device)
        self.action_norm = Normalizer(stats["action"], usr_args.get("action_normalization_type", "bound"), self.device)
        if self.proprio_norm.dim != self.cfg.proprio_dim or self.action_norm.dim != self.cfg.action_dim:
            raise ValueError("dataset statistics do not match the proprio / action dimensions of the preset")
        self.stager = FrameStager(self.cfg.image_size, center_crop=bool(usr_args.get("center_crop", False)))
        tok = usr_args.get("tokenizer")
        if callable(tok) and not hasattr(tok, "convert_tokens_to_ids"):
            self._tokenize: Callable[[str], torch.Tensor] = tok
        else:
            if isinstance(tok, str):
                from transformers import AutoTokenizer
                tok = AutoTokenizer.from_pretrained(tok, trust_remote_code=True, use_fast=False)
            if tok is None:
                raise ValueError("usr_args['tokenizer'] is required (tokenizer object, path or callable)")
            n_img_tok = self.cfg.image_tokens * self.cfg.num_images
            self._tokenize = lambda text: encode_prompt(tok, text, n_img_tok, self.cfg.max_image_text_tokens)
        self.act_steps = int(usr_args.get("act_steps", self.cfg.horizon_steps))
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(int(usr_args.get("seed", 0)))
        self.instruction: Optional[str] = None
        self.input_ids: Optional[torch.Tensor] = None
        self.observation_window = None

    # ---- the pi0 plugin's model surface (pi_model.py) -------------------------------------------------------------
    def set_language(self, instruction: str) -> None:
        self.instruction = instruction
        ids = torch.as_tensor(self._tokenize(instruction), dtype=torch.int64).reshape(1, -1)
        if ids.shape[1] != self.cfg.max_image_text_tokens:
            raise ValueError(f"tokenizer returned {ids.shape[1]} ids, expected {self.cfg.max_image_text_tokens}")
        self.input_ids = ids.to(self.device)

    def update_observation_window(self, img_arr, state) -> None:
        frames = torch.from_numpy(np.stack([np.ascontiguousarray(im) for im in img_arr])).to(self.device)  # [3, H, W, 3] u8
        self.observation_window = {"frames": frames, "state": torch.as_tensor(np.asarray(state), dtype=torch.float64)}

    def get_action(self) -> np.ndarray:
        """One policy call: [horizon_steps, 14] de-normalised joint targets (float64)."""
        assert self.observation_window is not None and self.input_ids is not None, "set the instruction and an observation first"
        w = self.observation_window
        pix = self.stager(w["frames"])                                                    # [3, 3, px, px] bf16
        prop = self.proprio_norm.normalize(w["state"].to(self.device).reshape(1, 1, -1), torch.bfloat16)
        noise = torch.randn((1, self.cfg.num_action_tokens, self.cfg.action_dim), generator=self.gen, device=self.device,
                            dtype=torch.float32).to(torch.bfloat16)
        act = self.policy.infer_action(self.input_ids, pix, proprios=prop, noise=noise)   # [1, horizon, 14] normalised
        return self.action_norm.denormalize(act[0].float()).cpu().numpy()

    def reset_obsrvationwindows(self) -> None:  # (sic: the evaluator-side name used by pi_model.py)
        self.instruction = None
        self.input_ids = None
        self.observation_window = None


def get_model(usr_args):
    return VlaserRoboTwinPolicy(usr_args)


def eval(TASK_ENV, model, observation):  # noqa: A001  (the evaluator looks the name up)
    """pi0/deploy_policy.py:30-50: instruction once per episode, one policy call, `act_steps` environment steps."""
    if model.observation_window is None:
        model.set_language(TASK_ENV.get_instruction())
    input_rgb_arr, input_state = encode_obs(observation)
    model.update_observation_window(input_rgb_arr, input_state)
    actions = model.get_action()[: model.act_steps]
    for action in actions:
        TASK_ENV.take_action(action)
        observation = TASK_ENV.get_obs()
        input_rgb_arr, input_state = encode_obs(observation)
        model.update_observation_window(input_rgb_arr, input_state)


def reset_model(model):
    model.reset_obsrvationwindows()
