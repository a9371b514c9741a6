"""CPU: the staging / post-processing oracle (oracle/staging_oracle.py) against the golden
vectors produced by the unmodified reference adapter + cv2 (tests/golden/make_staging_golden.py), and against the
cv2 build of this container when it is importable."""
from __future__ import annotations

import os

import numpy as np
import pytest

from oracle import staging_oracle as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def frames():
    return np.load(os.path.join(GOLD, "staging_frames.npz"))


@pytest.fixture(scope="module")
def post():
    return np.load(os.path.join(GOLD, "staging_post.npz"))


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_resize_bit_exact_vs_golden(frames, case):
    px = int(frames[f"px_{case}"])
    for f, want in zip(frames[f"frames_{case}"], frames[f"resized_{case}"]):
        got = S.resize_lanczos4_u8(f, (px, px))
        assert got.dtype == np.uint8 and np.array_equal(got, want)


def test_center_crop_branch(frames):
    f = frames["frames_b"][0]
    got = S.resize_lanczos4_u8(np.ascontiguousarray(S.center_crop(f)), (128, 128))
    assert np.array_equal(got, frames["resized_b_crop"][0])


def test_normalise_bit_exact_vs_reference_process_images(frames):
    for case in ("b", "c"):
        px = int(frames[f"px_{case}"])
        got = S.stage_frames(frames[f"frames_{case}"], px)
        want = frames[f"pixel_values_{case}"]
        assert got.dtype == np.float32 and np.array_equal(got.view(np.int32), want.view(np.int32))
    # bf16 cast of the model dtype (eval.py:43-45 `.to(dtype)`): compare the bit patterns
    import torch
    got = torch.from_numpy(S.stage_frames(frames["frames_a"], int(frames["px_a"]))).to(torch.bfloat16).view(torch.int16)
    assert np.array_equal(got.numpy(), frames["pixel_values_a_bf16bits"])


def test_resize_matches_live_cv2_at_evaluator_sizes():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for (h, w, o) in [(480, 640, 224), (512, 640, 224), (480, 640, 448), (37, 53, 64), (224, 224, 224)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(S.resize_lanczos4_u8(img, (o, o)), cv2.resize(img, (o, o), interpolation=cv2.INTER_LANCZOS4))


def test_lanczos_taps_are_an_identity_at_zero_offset():
    c = S.lanczos4_coeffs(np.float32(0.0))
    assert c[3] == 1.0 and np.abs(np.delete(c, 3)).max() < 1e-20


@pytest.mark.parametrize("name,norm,gripper", [("bridge", "bound", "bridge"), ("fractal", "gaussian", "sticky"),
                                               ("fractal_bound", "bound", "sticky")])
def test_postprocess_matches_reference_adapter(post, name, norm, gripper):
    acts, want = post[f"{name}_actions"], post[f"{name}_post"]
    stats = {k: post[f"{name}_action_{k}"] for k in ("mean", "std", "p01", "p99")}
    for e in range(acts.shape[0]):
        sticky = S.StickyGripper(15) if gripper == "sticky" else None
        got = np.concatenate([S.postprocess(acts[e, t:t + 4], stats, norm, gripper, sticky) for t in range(0, acts.shape[1], 4)])
        # float64 throughout, same libm: tolerance only for the order of two additions
        assert np.abs(got - want[e]).max() <= 1e-12, (name, e)
    if gripper == "sticky":
        assert len(np.unique(want[..., -1])) > 2  # the sticky branch was exercised (held values differ per env)


def test_proprio_normalisation_matches_reference(post):
    for name, norm in (("bridge", "bound"), ("fractal", "gaussian")):
        raw = post[f"{name}_proprio_raw"]
        if norm == "bound":
            got = S.normalize_bound(raw, post[f"{name}_proprio_p01"], post[f"{name}_proprio_p99"])
        else:
            got = S.normalize_gaussian(raw, post[f"{name}_proprio_mean"], post[f"{name}_proprio_std"])
        assert np.array_equal(got, post[f"{name}_proprio_norm"])
