"""GPU: the staging / post-processing kernels (csrc/staging.cu, through the C ABI) against the golden vectors of the
unmodified reference adapter + cv2 and against the CPU oracle at evaluator sizes.

Bars: resize + normalisation bit-exact (integer work, then float32 ops with one rounding each, then the bf16 cast);
post-processing float64 within 1e-12 (libm vs CUDA sin/cos/acos differ in the last ulp)."""
from __future__ import annotations

import os

import numpy as np
import pytest
import torch

from oracle import staging_oracle as S  # checker only

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def frames():
    return np.load(os.path.join(GOLD, "staging_frames.npz"))


@pytest.fixture(scope="module")
def post():
    return np.load(os.path.join(GOLD, "staging_post.npz"))


def _bits(t: torch.Tensor) -> np.ndarray:
    return t.view(torch.int16 if t.dtype == torch.bfloat16 else torch.int32).cpu().numpy()


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_stage_frames_vs_golden(frames, case):
    from vlaser_b200.adapter import FrameStager
    px = int(frames[f"px_{case}"])
    fr = torch.from_numpy(frames[f"frames_{case}"]).cuda()
    want_u8 = frames[f"resized_{case}"]                      # cv2.resize output, [n, px, px, 3]
    want = S.normalize_frames(want_u8.transpose(0, 3, 1, 2))  # processing.py arithmetic (pinned in test_staging.py)
    got = FrameStager(px, dtype=torch.float32)(fr)
    assert np.array_equal(_bits(got), want.view(np.int32))
    got16 = FrameStager(px, dtype=torch.bfloat16)(fr)
    assert np.array_equal(_bits(got16), _bits(torch.from_numpy(want).to(torch.bfloat16)))
    if f"pixel_values_{case}" in frames:                     # the reference's own process_images output
        assert np.array_equal(_bits(got), frames[f"pixel_values_{case}"].view(np.int32))
    if f"pixel_values_{case}_bf16bits" in frames:
        assert np.array_equal(_bits(got16), frames[f"pixel_values_{case}_bf16bits"])
    # CHW input layout gives the same result
    got_chw = FrameStager(px, dtype=torch.float32, layout="chw")(fr.permute(0, 3, 1, 2).contiguous())
    assert torch.equal(got_chw, got)


def test_stage_frames_center_crop(frames):
    from vlaser_b200.adapter import FrameStager
    fr = torch.from_numpy(frames["frames_b"]).cuda()
    got = FrameStager(128, center_crop=True, dtype=torch.float32)(fr)
    want = S.normalize_frames(frames["resized_b_crop"].transpose(0, 3, 1, 2))
    assert np.array_equal(_bits(got), want.view(np.int32))


@pytest.mark.parametrize("h,w,px,n", [(480, 640, 224, 3), (512, 640, 224, 2), (480, 640, 448, 1), (37, 53, 64, 2),
                                      (224, 224, 224, 2), (1080, 1920, 224, 1)])
def test_stage_frames_evaluator_sizes_vs_oracle(h, w, px, n):
    """SimplerEnv camera sizes (640x480 Bridge, 640x512 Google Robot) and edge cases (up-scaling, identity, large
    down-scaling window) against the numpy restatement of cv2's 8-bit Lanczos4."""
    from vlaser_b200.adapter import FrameStager
    rng = np.random.default_rng(h * 7 + w)
    fr = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    want = S.stage_frames(fr, px)
    got = FrameStager(px, dtype=torch.float32)(torch.from_numpy(fr).cuda())
    assert np.array_equal(_bits(got), want.view(np.int32))


def test_stage_frames_full_bench_batch_properties():
    """BASELINE configs[1] batch (64 frames 480x640 -> 224): constant frames stay constant through the resize (the
    taps sum to 2^11 only approximately, so check against the oracle on one constant frame), batch rows are
    independent, and the output feeds infer_action's layout [B, 3, 224, 224] bf16."""
    from vlaser_b200.adapter import FrameStager
    g = torch.Generator().manual_seed(0)
    fr = torch.randint(0, 256, (64, 480, 640, 3), generator=g, dtype=torch.uint8)
    fr[5] = 200
    st = FrameStager(224)
    out = st(fr.cuda())
    assert out.shape == (64, 3, 224, 224) and out.dtype == torch.bfloat16 and torch.isfinite(out.float()).all()
    one = st(fr[17:18].cuda())
    assert torch.equal(one[0], out[17])
    want5 = torch.from_numpy(S.stage_frames(fr[5:6].numpy(), 224)).to(torch.bfloat16)
    assert torch.equal(out[5].cpu(), want5[0])


@pytest.mark.parametrize("name,norm,gripper", [("bridge", "bound", "bridge"), ("fractal", "gaussian", "sticky"),
                                               ("fractal_bound", "bound", "sticky")])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_postprocess_vs_reference_adapter(post, name, norm, gripper, dtype):
    from vlaser_b200.adapter import ActionPostprocessor
    acts = torch.from_numpy(post[f"{name}_actions"])          # [n_env, 24, 7], six chunks of four per environment
    stats = {k: post[f"{name}_action_{k}"] for k in ("mean", "std", "p01", "p99")}
    n_env, T, _ = acts.shape
    pp = ActionPostprocessor(stats, norm, gripper, n_env=n_env)
    pp.reset()
    chunks = [pp(acts[:, t:t + 4].to(dtype).cuda()).cpu().numpy() for t in range(0, T, 4)]  # state carried across calls
    got = np.concatenate(chunks, axis=1)
    if dtype == torch.float32:
        want = post[f"{name}_post"]                           # unmodified reference adapter output
    else:                                                     # bf16 actions: same rule on the rounded inputs (oracle)
        a16 = acts.to(torch.bfloat16).float().numpy()
        want = np.stack([np.concatenate([S.postprocess(a16[e, t:t + 4], stats, norm, gripper, sg) for t in range(0, T, 4)])
                         for e, sg in ((e, S.StickyGripper(15)) for e in range(n_env))])
    assert np.abs(got - want).max() <= 1e-12
    # reset() of one environment restarts only that environment's sticky state
    if gripper == "sticky":
        pp.reset([1])
        again = pp(acts[:, :4].cuda()).cpu().numpy()
        assert np.abs(again[1] - post[f"{name}_post"][1, :4]).max() <= 1e-12


def test_normalizer_matches_base_env_adapter():
    """Device proprio normalisation / joint de-normalisation against the BaseEnvAdapter formulas (base.py:8-49) in float64."""
    from vlaser_b200.adapter import Normalizer
    rng = np.random.default_rng(3)
    for dim in (7, 14):
        lo, hi = rng.normal(size=dim) - 2.0, rng.normal(size=dim) + 2.0
        stats = {"p01": lo.tolist(), "p99": hi.tolist(), "mean": rng.normal(size=dim).tolist(),
                 "std": (rng.random(dim) + 0.1).tolist()}
        x = rng.normal(size=(5, 1, dim)) * 3.0   # some values outside [p01, p99]: the bound rule clips
        xd = torch.from_numpy(x).cuda()
        nb = Normalizer(stats, "bound")
        got = nb.normalize(xd, torch.float64).cpu().numpy()
        assert np.allclose(got, S.normalize_bound(x, lo, hi), rtol=0, atol=1e-15)
        assert got.min() >= -1.0 and got.max() <= 1.0
        y = rng.uniform(-1, 1, size=(5, 4, dim))
        back = nb.denormalize(torch.from_numpy(y).cuda()).cpu().numpy()
        assert np.allclose(back, (y + 1.0) / 2.0 * (hi - lo) + lo, rtol=0, atol=1e-14)
        ng = Normalizer(stats, "gaussian")
        mean, std = np.array(stats["mean"]), np.array(stats["std"])
        assert np.allclose(ng.normalize(xd, torch.float64).cpu().numpy(), S.normalize_gaussian(x, mean, std), rtol=0, atol=1e-14)
        assert np.allclose(ng.denormalize(torch.from_numpy(y).cuda()).cpu().numpy(), y * (std + 1e-8) + mean, rtol=0, atol=1e-14)
        # float32 / bf16 outputs are the float64 result rounded once
        assert torch.equal(nb.normalize(xd, torch.float32).cpu(), torch.from_numpy(S.normalize_bound(x, lo, hi)).float())
        assert torch.equal(nb.normalize(xd, torch.bfloat16).cpu(), torch.from_numpy(S.normalize_bound(x, lo, hi)).float().bfloat16())


def test_history_stacker_matches_reference_window():
    """simpler.py:92-104 with num_images = 2: first call repeats the frame, later calls give [previous, current]; a reset
    of one environment restarts only that environment's window."""
    from vlaser_b200.adapter import HistoryStacker
    g = torch.Generator().manual_seed(0)
    frames = [torch.randint(0, 256, (3, 8, 8, 3), generator=g, dtype=torch.uint8).cuda() for _ in range(4)]
    hs = HistoryStacker(2)
    w0 = hs(frames[0])
    assert w0.shape == (3, 2, 8, 8, 3) and torch.equal(w0[:, 0], frames[0]) and torch.equal(w0[:, 1], frames[0])
    w1 = hs(frames[1])
    assert torch.equal(w1[:, 0], frames[0]) and torch.equal(w1[:, 1], frames[1])
    hs.reset([1])
    w2 = hs(frames[2])
    assert torch.equal(w2[0, 0], frames[1][0]) and torch.equal(w2[1, 0], frames[2][1]) and torch.equal(w2[:, 1], frames[2])
    hs.reset()
    w3 = hs(frames[3])
    assert torch.equal(w3[:, 0], frames[3])
    assert torch.equal(HistoryStacker(1)(frames[0])[:, 0], frames[0])


def test_robotwin_plugin_drives_a_task_env():
    """integrations/robotwin/deploy_policy.py through the evaluator's calls (Vlaser_VLA/RoboTwin/script/eval_policy.py:38-46,
    209-210): get_model / eval / reset_model against a fake TASK_ENV; the joint targets handed to take_action are the
    de-normalised actions of the policy (checked against the fp32 oracle on the same staged inputs)."""
    from integrations.robotwin import deploy_policy as P
    from oracle import vlaser_oracle as O
    from vlaser_b200.config import PRESETS
    from vlaser_b200.weights import random_state_dict
    cfg = PRESETS["tiny_robotwin"]()
    rng = np.random.default_rng(5)
    stats = {k: {"p01": (-1.0 - rng.random(14)).tolist(), "p99": (1.0 + rng.random(14)).tolist()} for k in ("proprio", "action")}
    S_len = cfg.max_image_text_tokens
    n_img = cfg.image_tokens * cfg.num_images

    def tokenize(text):
        ids = torch.full((S_len,), cfg.pad_token_id, dtype=torch.int64)
        ids[:4] = torch.tensor([1, 2, 3, 4])
        ids[4:4 + n_img] = cfg.image_token_index
        ids[4 + n_img:4 + n_img + 5] = torch.tensor([7, 8, 9, 10, 11]) + len(text) % 3
        return ids

    class FakeEnv:
        def __init__(self):
            self.taken, self.t = [], 0

        def get_instruction(self):
            return "stack the blocks"

        def _obs(self):
            r = np.random.default_rng(100 + self.t)
            return {"observation": {cam: {"rgb": r.integers(0, 256, (240, 320, 3), dtype=np.uint8)} for cam in P.CAMERAS},
                    "joint_action": {"vector": r.normal(size=14)}}

        def get_obs(self):
            self.t += 1
            return self._obs()

        def take_action(self, action, action_type="qpos"):
            self.taken.append(np.asarray(action).copy())

    sd = random_state_dict(cfg, 2)
    model = P.get_model({"preset": "tiny_robotwin", "checkpoint": sd, "dataset_statistics": stats, "tokenizer": tokenize,
                         "act_steps": 3, "seed": 11})
    env = FakeEnv()
    obs0 = env._obs()
    P.eval(env, model, obs0)
    assert len(env.taken) == 3 and all(a.shape == (14,) and a.dtype == np.float64 for a in env.taken)
    assert model.observation_window is not None and env.t == 3
    # the same call by hand: staged frames + normalised proprio + the same noise draw -> oracle -> de-normalise
    rgb, state = P.encode_obs(obs0)
    pix = model.stager(torch.from_numpy(np.stack(rgb)).cuda()).float().cpu()
    prop = torch.from_numpy(S.normalize_bound(np.asarray(state), np.array(stats["proprio"]["p01"]),
                                              np.array(stats["proprio"]["p99"]))).float().bfloat16().float().reshape(1, 1, -1)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(11)
    noise = torch.randn((1, cfg.num_action_tokens, 14), generator=gen, device="cuda").bfloat16().float().cpu()
    ref = O.infer_action(cfg, sd, tokenize("stack the blocks")[None], pix, prop, noise)[0].double().numpy()
    lo, hi = np.array(stats["action"]["p01"]), np.array(stats["action"]["p99"])
    want = (ref + 1.0) / 2.0 * (hi - lo) + lo
    assert np.abs(np.stack(env.taken) - want[:3]).max() <= 1e-2 * (hi - lo).max()
    P.reset_model(model)
    assert model.observation_window is None and model.instruction is None
