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
