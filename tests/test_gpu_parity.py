"""-m gpu: end-to-end parity of the CUDA path (through the C ABI, via the PiZeroInference binding)
against (a) the golden vectors produced by the unmodified reference and (b) the fp32 oracle, with
per-stage hidden-state relative errors (SURVEY.md §8d "Correctness gates").

Tolerances (stated by BASELINE.json north_star): normalised actions max-abs error <= 1e-2 vs the fp32
reference on identical bf16-rounded weights/inputs and injected noise; per-layer hidden states are
checked by relative L2 error; indexing work (token layout, pixel_shuffle, patch order) is bit-exact.
"""
from __future__ import annotations

import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import vlaser_oracle as O  # noqa: E402  (checker only)
from tests.golden_util import CASES, DYNAMIC_TILE_CASES, FLOW_CASES, GENERATE_CASES, load_case  # noqa: E402
from tests.ops import max_err, rel_err  # noqa: E402
from vlaser_b200.config import tiny, vlaser_2b_bridge  # noqa: E402
from vlaser_b200.policy import PiZeroInference  # noqa: E402
from vlaser_b200.synth import make_inputs  # noqa: E402
from vlaser_b200.weights import random_state_dict, text_head_state_dict  # noqa: E402

ACTION_TOL = 1e-2


def _policy(cfg, sd, B, **opts):
    pol = PiZeroInference(cfg, max_batch=B)
    pol.load_state_dict(sd, strict=True)
    for k, v in opts.items():
        pol.set_option(k, v)
    return pol


def _run(pol, inp, dtype=torch.float32, masks=False):
    ids = inp["input_ids"].cuda()
    pix = inp["pixel_values"].to(dtype).cuda()
    prop = inp["proprios"].to(dtype).cuda()
    noise = inp["noise"].to(dtype).cuda()
    if masks:
        am = (inp["input_ids"] != pol.pad_token_id).to(torch.int64)
        m, vp, pp, ap = pol.build_causal_mask_and_position_ids(am, dtype)
        itp, amask = pol.split_full_mask_into_submasks(m)
        return pol(ids, pix, itp.cuda(), amask.cuda(), vp.cuda(), pp.cuda(), ap.cuda(), prop, noise=noise).float().cpu()
    return pol.infer_action(ids, pix, proprios=prop, noise=noise).float().cpu()


def _stage_report(cfg, sd, inp, pol, B, log):
    """Per-stage relative errors of the CUDA path vs the fp32 oracle. Returns dict of errors."""
    tr = {}
    O.infer_action(cfg.with_(final_action_clip_value=None), sd, inp["input_ids"], inp["pixel_values"], inp["proprios"],
                   inp["noise"], trace=tr)
    errs = {}
    S, H, A = cfg.max_image_text_tokens, cfg.num_action_tokens, cfg.action_dim
    Ni = B * cfg.num_images
    valid = (inp["input_ids"] != cfg.pad_token_id).sum(-1).tolist()
    # ---- ViT
    pol.set_option("stop_after", 1)
    _run(pol, inp)
    vit = pol.debug_buffer("vit_x", torch.float32)[: Ni * cfg.vit_tokens * cfg.vit_hidden].view(Ni, cfg.vit_tokens, -1)
    errs["vit_embed"] = None
    errs["vit_out"] = rel_err(vit, tr["vit_out"])
    # ---- prefix embedding (packed rows)
    pol.set_option("stop_after", 2)
    _run(pol, inp)
    assert pol.debug_buffer("valid_len", torch.int32)[:B].tolist() == valid
    cu = pol.debug_buffer("cu", torch.int32)[: B + 1].tolist()
    px = pol.debug_buffer("prefix_x", torch.float32)[: cu[-1] * cfg.hidden].view(cu[-1], cfg.hidden)
    ref = torch.cat([tr["prefix_embeds"][b, : valid[b]] for b in range(B)])
    errs["prefix_embeds"] = rel_err(px, ref)
    # text rows are a pure gather of bf16-exact embeddings: bit exact
    is_text = torch.cat([(inp["input_ids"][b, : valid[b]] != cfg.image_token_index) for b in range(B)])
    assert torch.equal(px[is_text], ref[is_text]), "text-token rows of the prefix must be bit-exact"
    xp = pol.debug_buffer("proprio_x", torch.float32)[: B * cfg.hidden_e].view(B, 1, -1)
    errs["proprio_embeds"] = rel_err(xp, tr["proprio_embeds"])
    # ---- prefill: KV caches
    pol.set_option("stop_after", 3)
    _run(pol, inp)
    kv_rows = int(pol.debug_buffer("kv_rows", torch.int64)[0])
    kvoff = pol.debug_buffer("kvoff", torch.int32)[:B].tolist()
    kc = pol.debug_buffer("k_cache", torch.bfloat16).view(cfg.layers, cfg.n_kv_heads, kv_rows, 128).float()
    vc = pol.debug_buffer("v_cache", torch.bfloat16).view(cfg.layers, cfg.n_kv_heads, kv_rows, 128).float()
    for li in (0, cfg.layers - 1):
        for nm, cache in (("k", kc), ("v", vc)):
            got = torch.cat([cache[li, :, kvoff[b]: kvoff[b] + valid[b]] for b in range(B)], dim=1)
            want = torch.cat([tr[f"{nm}_vlm_layer{li}"][b, :, : valid[b]] for b in range(B)], dim=1)
            errs[f"{nm}_vlm_layer{li}"] = rel_err(got, want)
            gotp = torch.stack([cache[li, :, kvoff[b] + valid[b]] for b in range(B)])
            errs[f"{nm}_proprio_layer{li}"] = rel_err(gotp, tr[f"{nm}_proprio_layer{li}"][:, :, 0])
    # ---- first denoise step
    pol.set_option("stop_after", 4)
    _run(pol, inp)
    act = pol.debug_buffer("action", torch.float32)[: B * H * A].view(B, H, A)
    errs["action_step0_maxabs"] = max_err(act, tr["action_step0"])
    pol.set_option("stop_after", 0)
    for k, v in errs.items():
        if v is not None:
            log(f"    {k:28s} {v:.3e}")
    return errs, tr


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("simt", [0, 1])
def test_golden_parity(name, simt, capsys):
    """CUDA path vs the reference's golden vectors (tiny configs: bridge / fractal / 3-camera RoboTwin)."""
    cfg, sd, inp, gold = load_case(name)
    B = inp["input_ids"].shape[0]
    pol = _policy(cfg, sd, B, simt_gemm=simt)
    act = _run(pol, inp)
    e_post = max_err(act, gold["actions"])
    pol_pre = _policy(cfg.with_(final_action_clip_value=None), sd, B, simt_gemm=simt)
    pre = _run(pol_pre, inp)
    e_pre = max_err(pre, gold["trace_actions_preclamp_sliced"])
    with capsys.disabled():
        print(f"\n[{name} simt={simt}] actions max-abs vs reference golden: post-clamp {e_post:.3e}  pre-clamp {e_pre:.3e}")
        errs, _ = _stage_report(cfg, sd, inp, pol, B, print)
    assert e_post <= ACTION_TOL and e_pre <= ACTION_TOL
    assert errs["vit_out"] < 1e-2 and errs["prefix_embeds"] < 1e-2
    assert errs["k_vlm_layer0"] < 1e-2 and errs[f"k_vlm_layer{cfg.layers - 1}"] < 3e-2
    assert errs["action_step0_maxabs"] < ACTION_TOL


def test_graph_equals_direct_and_mask_api():
    cfg, sd, inp, gold = load_case("tiny_bridge")
    B = inp["input_ids"].shape[0]
    pol = _policy(cfg, sd, B, use_graph=1)
    a1 = _run(pol, inp)
    a2 = _run(pol, inp)            # graph replay
    a3 = _run(pol, inp, masks=True)  # reference-style call with dense masks / position ids
    pol2 = _policy(cfg, sd, B, use_graph=0)
    a4 = _run(pol2, inp)
    a5 = _run(pol2, inp, dtype=torch.bfloat16)
    # split-K partial sums are added in split order (option det_splitk): replay, direct launches and the mask API
    # agree bit for bit
    assert torch.equal(a1, a2) and torch.equal(a1, a3) and torch.equal(a1, a4)
    assert max_err(a5, gold["actions"]) < 2e-2  # bf16 I/O rounds the returned actions
    assert pol.last_launch_count() > 50


@pytest.mark.parametrize("name", CASES)
def test_expert_stream_kernel_matches_per_kernel_schedule(name):
    """The persistent expert-stream kernel (one launch per proprio / denoise pass) against the per-kernel schedule
    (7 launches per layer) on the same inputs: same math, different launch structure and split-K summation order."""
    cfg, sd, inp, gold = load_case(name)
    B = inp["input_ids"].shape[0]
    pol1 = _policy(cfg, sd, B, expert_mega=1)
    pol0 = _policy(cfg, sd, B, expert_mega=0)
    a1, a0 = _run(pol1, inp), _run(pol0, inp)
    assert pol1.last_launch_count() < pol0.last_launch_count()
    assert max_err(a1, a0) < 2e-3
    assert max_err(a1, gold["actions"]) <= ACTION_TOL
    # K/V rows written by the proprio pass and by the last denoise step
    for nm in ("k_cache", "v_cache"):
        c1, c0 = pol1.debug_buffer(nm, torch.bfloat16).float(), pol0.debug_buffer(nm, torch.bfloat16).float()
        assert rel_err(c1, c0) < 1e-2


def test_batch_row_independence_and_ragged():
    cfg = tiny()
    sd = random_state_dict(cfg, 3)
    inp = make_inputs(cfg, 5, 21, [3, 14, 7, 1, 9])
    pol = _policy(cfg, sd, 5)
    full = _run(pol, inp)
    for b in (0, 3):
        one = {k: v[b:b + 1] for k, v in inp.items() if k in ("input_ids", "pixel_values", "proprios", "noise")}
        # split-K partial sums land with fp32 atomics, and the split count of the ViT / LLM residual projections follows
        # the tile count of the batch: the summation order (not the math) depends on the batch size
        assert max_err(_run(pol, one), full[b:b + 1]) < 3e-3
    ref = O.infer_action(cfg, sd, inp["input_ids"], inp["pixel_values"], inp["proprios"], inp["noise"])
    assert max_err(full, ref) <= ACTION_TOL


def test_host_entry_point():
    cfg, sd, inp, gold = load_case("tiny_bridge")
    B = inp["input_ids"].shape[0]
    pol = _policy(cfg, sd, B)
    out = pol.infer_action_host(inp["input_ids"], inp["pixel_values"], inp["proprios"], inp["noise"])
    assert max_err(out, gold["actions"]) <= ACTION_TOL


@pytest.mark.skipif(os.environ.get("VLASER_SKIP_FULL", "0") == "1", reason="full-size parity skipped by env")
def test_full_size_2b_parity(capsys):
    """BASELINE config 1 shapes (Vlaser-2B-VLA, 1x224 px, S=384, 10 steps), B=2, vs the fp32 oracle."""
    cfg = vlaser_2b_bridge()
    sd = random_state_dict(cfg, 0)
    B = 2
    inp = make_inputs(cfg, B, 1)
    pol = _policy(cfg.with_(final_action_clip_value=None), sd, B)
    pre = _run(pol, inp)
    with capsys.disabled():
        print("\n[vlaser_2b_bridge B=2] per-stage relative errors vs fp32 oracle:")
        errs, tr = _stage_report(cfg, sd, inp, pol, B, print)
        ref_pre = tr["actions_preclamp"][:, -cfg.horizon_steps:]
        e_pre = max_err(pre, ref_pre)
        e_post = max_err(pre.clamp(-1, 1), ref_pre.clamp(-1, 1))
        print(f"    actions max-abs: pre-clamp {e_pre:.3e} post-clamp {e_post:.3e} (mean-abs pre {float((pre - ref_pre).abs().mean()):.3e})")
    assert errs["vit_out"] < 2e-2 and errs["prefix_embeds"] < 2e-2
    assert e_post <= ACTION_TOL and e_pre <= ACTION_TOL
    assert torch.equal(_run(pol, inp), pre), "repeated calls must agree bit for bit"


def _rows(inp, idx):
    return {k: inp[k][idx] for k in ("input_ids", "pixel_values", "proprios", "noise")}


@pytest.mark.skipif(os.environ.get("VLASER_SKIP_FULL", "0") == "1", reason="full-size parity skipped by env")
def test_ln_overlap_changes_no_bit():
    """Option ln_overlap (default on): the LayerNorm launches of the ViT run under the residual GEMM in front of them,
    released row block by row block through completion counters instead of waiting for the GEMM to finish. Same kernels,
    same arithmetic: the actions equal the serial schedule bit for bit (full-size 2B, B = 16: the ViT GEMMs take the
    CTA-pair kernel that reports the blocks), from call to call as well."""
    from vlaser_b200.config import PRESETS
    cfg = PRESETS["vlaser_2b_bridge"]()
    sd = random_state_dict(cfg, 0)
    B = 16
    inp = make_inputs(cfg, B, 11)
    on = _policy(cfg, sd, B)
    a1 = _run(on, inp, dtype=torch.bfloat16)
    a2 = _run(on, inp, dtype=torch.bfloat16)
    off = _policy(cfg, sd, B, ln_overlap=0)
    a0 = _run(off, inp, dtype=torch.bfloat16)
    assert torch.equal(a1, a2) and torch.equal(a1, a0)
    assert torch.isfinite(a1).all()


@pytest.mark.skipif(os.environ.get("VLASER_SKIP_FULL", "0") == "1", reason="full-size parity skipped by env")
@pytest.mark.parametrize("opt", ["fuse_norm_llm", "half_sm"])
def test_full_size_optional_schedules_vs_oracle(opt, capsys):
    """Schedules that are measured but off by default, at full size and a batch large enough to take them (B = 16: the
    LLM prefill on the CTA-pair kernel, 64 action rows... the o / gate-up / down launches of a denoise step at one tile per
    CTA pair): `fuse_norm_llm` (RMSNorm of the LLM prefill fused across the GEMM chain, transposed read-modify-write
    epilogue + per-row partial sums of squares) and `half_sm` (half-SM variant of the pair kernel). Same bar as the
    default schedule: <= 1e-2 against the fp32 oracle pre- and post-clamp, bit-identical from call to call."""
    from vlaser_b200.config import PRESETS
    cfg = PRESETS["vlaser_2b_bridge"]().with_(final_action_clip_value=None)
    sd = random_state_dict(cfg, 0)
    B = 64 if opt == "half_sm" else 16
    inp = make_inputs(cfg, B, 7)
    for k in ("pixel_values", "proprios", "noise"):
        inp[k] = inp[k].bfloat16().float()
    base = _run(_policy(cfg, sd, B), inp)
    pol = _policy(cfg, sd, B, **{opt: 1})
    got = _run(pol, inp)
    assert torch.equal(_run(pol, inp), got), "repeated calls must agree bit for bit"
    rows = [0, B // 2 - 1, B - 1]
    ref = O.infer_action(cfg, sd, **{k: v for k, v in _rows(inp, rows).items()})
    worst = 0.0
    for i, b in enumerate(rows):
        worst = max(worst, max_err(got[b], ref[i]), max_err(got[b].clamp(-1, 1), ref[i].clamp(-1, 1)))
    with capsys.disabled():
        print(f"\n[{opt}=1, B={B}] max-abs vs fp32 oracle {worst:.3e}; vs the default schedule {max_err(got, base):.3e}")
    assert worst <= ACTION_TOL
    if opt == "half_sm":
        assert torch.equal(got, base), "same splits, same summation order: the half-SM kernels change no bit"
    else:
        assert not torch.equal(got, base), "the fused chain rounds the operand before the row scale: the option did not take"


@pytest.mark.skipif(os.environ.get("VLASER_SKIP_FULL", "0") == "1", reason="full-size parity skipped by env")
@pytest.mark.parametrize("preset,B", [("vlaser_2b_bridge", 64), ("vlaser_2b_fractal", 64), ("vlaser_2b_bridge", 128)])
def test_full_size_b64_default_schedule_vs_oracle(preset, B, capsys):
    """The benched configuration itself (BASELINE configs[1] / [2] shapes: full-size 2B, B = 64, every option at its
    default) against the fp32 oracle on rows spread over the batch: first / last row and one row of each 128-row tile
    of the action stream (rows 4b..4b+3). Inputs are rounded to bf16 first, so the fp32-I/O call, the bf16-I/O call
    (what bench.py times) and the oracle all see the same numbers; pre- and post-clamp, per row."""
    from vlaser_b200.config import PRESETS
    cfg = PRESETS[preset]().with_(final_action_clip_value=None)
    sd = random_state_dict(cfg, 0)
    # (B = 128: the per-GPU shard of BASELINE configs[2] at 4 GPUs -- 512 action rows take the multi-tile GEMM paths)
    inp = make_inputs(cfg, B, 5)
    for k in ("pixel_values", "proprios", "noise"):
        inp[k] = inp[k].bfloat16().float()
    pol = _policy(cfg, sd, B)
    got = _run(pol, inp)
    got16 = _run(pol, inp, dtype=torch.bfloat16)
    assert torch.equal(got16, got.bfloat16().float()), "bf16 I/O must only round the returned actions"
    assert torch.equal(_run(pol, inp), got), "repeated calls must agree bit for bit"
    rows = [0, 13, 31, 32, 47, 63] if B == 64 else [0, 31, 64, 95, 127]
    ref = O.infer_action(cfg, sd, **{k: v for k, v in _rows(inp, rows).items()})
    worst_pre = worst_post = 0.0
    for i, b in enumerate(rows):
        e_pre = max_err(got[b], ref[i])
        e_post = max_err(got[b].clamp(-1, 1), ref[i].clamp(-1, 1))
        worst_pre, worst_post = max(worst_pre, e_pre), max(worst_post, e_post)
        assert e_pre <= ACTION_TOL and e_post <= ACTION_TOL, (preset, b, e_pre, e_post)
    with capsys.disabled():
        print(f"\n[{preset} B={B} default schedule] rows {rows}: max-abs vs fp32 oracle pre-clamp {worst_pre:.3e} "
              f"post-clamp {worst_post:.3e}")


@pytest.mark.skipif(os.environ.get("VLASER_SKIP_FULL", "0") == "1", reason="full-size parity skipped by env")
def test_full_size_2b_bench_batch_expert_stream():
    """BASELINE configs[1] size (B=64: 256 action rows = two 128-row sub-tiles per expert tile, 64 proprio rows):
    the persistent expert-stream kernel against the per-kernel schedule on the same handle, and row independence
    of the batch (sample b of the B=64 call == the same sample run alone)."""
    cfg = vlaser_2b_bridge()
    sd = random_state_dict(cfg, 0)
    B = 64
    inp = make_inputs(cfg, B, 5)
    pol = _policy(cfg, sd, B)
    pol.set_option("expert_mega", 1)
    a1 = _run(pol, inp, dtype=torch.bfloat16)
    pol.set_option("expert_mega", 0)
    a0 = _run(pol, inp, dtype=torch.bfloat16)
    pol.set_option("expert_mega", 1)
    assert torch.isfinite(a1).all()
    assert max_err(a1, a0) <= ACTION_TOL
    # independent sub-grids of the expert-stream kernel (1 = whole grid, 2 = 2 x 128 rows, 4 = 4 x 64 rows)
    for streams in (1, 2, 4):
        pol.set_option("expert_streams", streams)
        a_s = _run(pol, inp, dtype=torch.bfloat16)
        assert torch.isfinite(a_s).all()
        assert max_err(a_s, a0) <= ACTION_TOL, streams
    pol.set_option("expert_streams", 2)
    for b in (7, 40, 63):   # one sample of each sub-grid (and the last row of the batch) run alone
        one = {k: v[b:b + 1] for k, v in inp.items() if k in ("input_ids", "pixel_values", "proprios", "noise")}
        assert max_err(_run(pol, one, dtype=torch.bfloat16), a1[b:b + 1]) <= ACTION_TOL


@pytest.mark.skipif(os.environ.get("VLASER_SKIP_FULL", "0") == "1", reason="full-size parity skipped by env")
@pytest.mark.parametrize("preset,B", [("vlaser_2b_fractal", 2), ("vlaser_2b_robotwin", 1)])
def test_full_size_other_configs_parity(preset, B):
    """BASELINE configs 3 and 4 at full model size: Google Robot / fractal (proprio_dim 8, time_max_period 100) and the
    RoboTwin shapes (3 cameras, 14-DoF chunk, S=896, H=6 action tokens) against the fp32 oracle, pre-clamp."""
    from vlaser_b200.config import PRESETS
    cfg = PRESETS[preset]().with_(final_action_clip_value=None)
    sd = random_state_dict(cfg, 0)
    inp = make_inputs(cfg, B, 3)
    pol = _policy(cfg, sd, B)
    got = _run(pol, inp)
    ref = O.infer_action(cfg, sd, inp["input_ids"], inp["pixel_values"], inp["proprios"], inp["noise"])
    assert got.shape == (B, cfg.horizon_steps, cfg.action_dim)
    assert max_err(got.clamp(-1, 1), ref.clamp(-1, 1)) <= ACTION_TOL
    pol.set_option("expert_mega", 1)
    assert max_err(_run(pol, inp).clamp(-1, 1), ref.clamp(-1, 1)) <= ACTION_TOL


def test_fp32_checkpoint_weight_rounding(capsys):
    """The shipped evaluator configs run the reference in fp32 (`use_bf16: False`); this library rounds every projection
    weight to bf16 at load. Deviation on a checkpoint whose weights are NOT bf16-representable (plain fp32 random init):
    library (fp32 weights in, `.to(torch.float32)` requested -> one RuntimeWarning) vs the fp32 oracle on the un-rounded
    weights, next to the deviation on bf16-representable weights."""
    import warnings
    cfg = tiny().with_(final_action_clip_value=None)
    inp = make_inputs(cfg, 4, 9, [5, 11, 3, 8])
    res = {}
    for rounded in (True, False):
        sd = random_state_dict(cfg, 6, bf16_round=rounded)
        pol = PiZeroInference(cfg, max_batch=4)
        pol.load_state_dict(sd, strict=True)
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            pol.to(torch.float32)
            pol.to(torch.float32)
        assert sum(issubclass(x.category, RuntimeWarning) for x in w) == 1, "fp32 request must warn exactly once"
        got = _run(pol, inp)
        ref = O.infer_action(cfg, sd, inp["input_ids"], inp["pixel_values"], inp["proprios"], inp["noise"])
        res[rounded] = max_err(got, ref)
    with capsys.disabled():
        print(f"\n[tiny, pre-clamp] max-abs vs fp32 oracle: bf16-representable weights {res[True]:.3e}, "
              f"plain fp32 weights (rounded to bf16 at load) {res[False]:.3e}")
    assert res[True] <= ACTION_TOL
    assert res[False] <= 3 * ACTION_TOL, "weight rounding of an fp32 checkpoint moved the actions by more than 3e-2"


def test_checkpoint_key_aliases_load():
    """eval.py:196-212: the released checkpoints carry `_orig_mod.` prefixes (torch.compile), the expert weights under
    `mixtures.proprio.*` as well as `mixtures.action.*` (shared modules), duplicates under `internvl_model.*` and
    tensors the path never reads (the expert's lm_head); `load_state_dict(strict=False)` must take all of that and give the same
    actions as the canonical state dict."""
    cfg, sd, inp, gold = load_case("tiny_bridge")
    B = inp["input_ids"].shape[0]
    ck = {}
    for k, v in sd.items():
        ck["_orig_mod." + k] = v
        if k.startswith("joint_model.mixtures.action."):
            ck["_orig_mod." + k.replace("mixtures.action.", "mixtures.proprio.")] = v
        if k.startswith("vision_tower.vision_model."):
            ck["_orig_mod." + k.replace("vision_tower.vision_model.", "internvl_model.vision_model.")] = v
        if k.startswith("multi_modal_projector."):
            ck["_orig_mod." + k.replace("multi_modal_projector.", "internvl_model.mlp1.")] = v
    # the output head of the LLM is taken (the chat-VLM text path reads it); a tensor no path reads is reported back
    ck["_orig_mod.internvl_model.language_model.lm_head.weight"] = torch.zeros(cfg.embed_rows, cfg.hidden)
    ck["_orig_mod.internvl_model.action_expert.lm_head.weight"] = torch.zeros(8, 8)
    ck["step"] = 123  # non-tensor entries are skipped
    pol = PiZeroInference(cfg, max_batch=B)
    res = pol.load_state_dict(ck, strict=False)
    assert any("action_expert.lm_head" in k for k in res.unexpected_keys)
    assert not any("language_model.lm_head" in k for k in res.unexpected_keys)
    assert max_err(_run(pol, inp), gold["actions"]) <= ACTION_TOL
    # a missing tensor must fail loudly at finalize, not produce numbers
    short = dict(list(sd.items())[:-3])
    pol2 = PiZeroInference(cfg, max_batch=B)
    pol2.load_state_dict(short, strict=False)
    with pytest.raises(RuntimeError):
        _run(pol2, inp)


@pytest.mark.skipif(os.environ.get("VLASER_SKIP_FULL", "0") == "1", reason="full-size parity skipped by env")
def test_8b_widths_448px_reduced_depth_parity(capsys):
    """BASELINE config 5 shapes: the Vlaser-8B widths (LLM 3584 / 28:4 heads / MLP 18944, expert MLP 5600 which is
    padded to 5632 inside the library), a 448 px tile (1025 ViT tokens -> 256 image tokens, generic var-len ViT
    attention path), at reduced depth (2 LLM / expert layers, 2 ViT layers, 1024-row embedding) so that the fp32
    oracle runs in seconds. Prefix encode + prefill are checked stage by stage, then the whole call."""
    from vlaser_b200.config import vlaser_8b_vla
    cfg = vlaser_8b_vla().with_(layers=2, vit_layers=2, embed_rows=1024, image_token_index=1000, pad_token_id=999,
                                final_action_clip_value=None)
    sd = random_state_dict(cfg, 4)
    B = 2
    inp = make_inputs(cfg, B, 9)
    pol = _policy(cfg, sd, B)
    got = _run(pol, inp)
    with capsys.disabled():
        print("\n[vlaser_8b_vla widths, 448 px, 2 layers, B=2] per-stage relative errors vs fp32 oracle:")
        errs, tr = _stage_report(cfg, sd, inp, pol, B, print)
    ref = tr["actions_preclamp"][:, -cfg.horizon_steps:]
    assert errs["vit_out"] < 2e-2 and errs["prefix_embeds"] < 2e-2
    for k, v in errs.items():
        if k.startswith(("k_", "v_")):
            assert v < 3e-2, (k, v)
    assert max_err(got, ref) <= ACTION_TOL
    pol.set_option("expert_mega", 1)
    assert max_err(_run(pol, inp), ref) <= ACTION_TOL


@pytest.mark.parametrize("name", CASES)
def test_single_observation_gemv_kernel(name):
    """B = 1 routes the expert passes through the weight-streaming kernel (expert_gemv.cu; R = 1 proprio row, R = 4 or 8
    action rows): same actions as the per-kernel tensor-core schedule and as the reference, K/V rows of the proprio
    pass and of the last denoise step included."""
    cfg, sd, inp, gold = load_case(name)
    one = {k: v[:1] for k, v in inp.items() if k in ("input_ids", "proprios", "noise")}
    one["pixel_values"] = inp["pixel_values"][: cfg.num_images]
    pol1 = _policy(cfg, sd, 1, expert_gemv=1)
    pol0 = _policy(cfg, sd, 1, expert_gemv=0)
    a1, a0 = _run(pol1, one), _run(pol0, one)
    assert pol1.last_launch_count() < pol0.last_launch_count()
    assert max_err(a1, a0) < 2e-3
    assert max_err(a1, gold["actions"][:1]) <= ACTION_TOL
    assert torch.equal(a1, _run(pol1, one)) and torch.equal(a0, _run(pol0, one))
    for nm in ("k_cache", "v_cache"):
        c1, c0 = pol1.debug_buffer(nm, torch.bfloat16).float(), pol0.debug_buffer(nm, torch.bfloat16).float()
        assert rel_err(c1, c0) < 1e-2


@pytest.mark.skipif(os.environ.get("VLASER_SKIP_FULL", "0") == "1", reason="full-size parity skipped by env")
def test_full_size_2b_single_observation():
    """BASELINE configs[0] shape on the GPU: Vlaser-2B-VLA, one 224 px observation, 10 flow steps (the p50 B=1 latency
    path) against the fp32 oracle, with the graph replay equal to the direct launch."""
    cfg = vlaser_2b_bridge().with_(final_action_clip_value=None)
    sd = random_state_dict(cfg, 0)
    inp = make_inputs(cfg, 1, 17)
    pol = _policy(cfg, sd, 1)
    got = _run(pol, inp)
    ref = O.infer_action(cfg, sd, inp["input_ids"], inp["pixel_values"], inp["proprios"], inp["noise"])
    assert max_err(got.clamp(-1, 1), ref.clamp(-1, 1)) <= ACTION_TOL
    # graph replay == direct launch up to the order of the split-K red.adds into the fp32 residual streams (ViT, LLM and
    # expert projections all split K at B = 1); the un-clamped actions of this random-init model reach |a| ~ 3
    pol.set_option("use_graph", 0)
    # (round 1 saw up to 6e-3 call to call from the fp32 atomics of split-K, profiles/r01aa_b1_parity_spread.txt; the
    # partial sums are now combined in a fixed order -- workspace partials / integer accumulators -- so repeated calls
    # and the graph / direct schedules agree bit for bit)
    again = _run(pol, inp)
    assert torch.equal(again, got)
    pol.set_option("use_graph", 1)
    assert torch.equal(_run(pol, inp), got)
    pol.set_option("expert_gemv", 0)
    other = _run(pol, inp)
    assert max_err(other.clamp(-1, 1), ref.clamp(-1, 1)) <= ACTION_TOL
    assert max_err(other, got) < 5e-3
    assert torch.equal(_run(pol, inp), other)


def test_dynamic_tiles_parity():
    """Dynamic tiles: the samples of one batch hold 1 / 3 / 2 / 2 image tiles; projector
    rows are handed out in batch order as the reference's masked scatter does (oracle pinned to the live reference for
    this layout in tests/test_oracle.py). Whole call against the fp32 oracle; the token merge is bit-exact."""
    cfg, sd, inp, gold = load_case(DYNAMIC_TILE_CASES[0])
    counts = (inp["input_ids"] == cfg.image_token_index).sum(1) // cfg.image_tokens
    assert counts.tolist() == [1, 3, 2, 2]
    pol = _policy(cfg, sd, 4)
    got = _run(pol, inp)
    assert max_err(got, gold["actions"]) <= ACTION_TOL            # the reference's own output
    ref = O.infer_action(cfg, sd, inp["input_ids"], inp["pixel_values"], inp["proprios"], inp["noise"])
    assert max_err(got, ref) <= ACTION_TOL
    # token merge: image rows of the packed prefix land where the reference's masked scatter puts them
    # (bit-exact index check: projector row k -> packed prefix row of the k-th image token in batch order)
    ids = inp["input_ids"]
    valid = (ids != cfg.pad_token_id).sum(1)
    cu = torch.cumsum(valid, 0) - valid
    expected = [int(cu[b]) + j for b in range(ids.shape[0]) for j in range(ids.shape[1]) if int(ids[b, j]) == cfg.image_token_index]
    assert pol.debug_buffer("img_row_map", torch.int32)[: len(expected)].tolist() == expected


def test_prefix_fills_max_tokens_and_batch_bounds():
    """Edge sizes: a prefix that uses every one of the max_image_text_tokens positions (valid_len == S, no pad) beside a
    minimal one in the same batch, the full max_batch, and the error behaviour outside it (negative code -> RuntimeError,
    like the reference's asserts: joint_model.py:427-431)."""
    cfg = tiny()
    sd = random_state_dict(cfg, 4)
    S, T = cfg.max_image_text_tokens, cfg.image_tokens
    longest = S - (10 + T + 7)                      # prompt tokens that make the prefix exactly S long
    inp = make_inputs(cfg, 3, 31, [longest, 0, longest])
    assert inp["valid_len"].tolist() == [S, 10 + T + 7, S]
    pol = _policy(cfg, sd, 3)
    got = _run(pol, inp)
    ref = O.infer_action(cfg, sd, inp["input_ids"], inp["pixel_values"], inp["proprios"], inp["noise"])
    assert max_err(got, ref) <= ACTION_TOL
    # more observations than the handle was created for: loud error, no partial result
    big = make_inputs(cfg, 4, 32)
    with pytest.raises((RuntimeError, AssertionError)):
        _run(pol, big)
    # the handle is still usable afterwards
    assert max_err(_run(pol, inp), ref) <= ACTION_TOL


# ---- flow-matching training forward (pizero_internvl.py:1064-1197) ------------------------------
def _flow(pol, inp, dtype=torch.float32, masks=False):
    ids = inp["input_ids"].cuda()
    kw = {}
    if masks:
        am = (inp["input_ids"] != pol.pad_token_id).to(torch.int64)
        m, vp, pp, ap = pol.build_causal_mask_and_position_ids(am, dtype)
        kw = dict(causal_mask=m.cuda(), vlm_position_ids=vp, proprio_position_ids=pp, action_position_ids=ap)
    loss, v = pol.flow_matching_loss(ids, inp["pixel_values"].to(dtype).cuda(), proprios=inp["proprios"].to(dtype).cuda(),
                                     actions=inp["actions"].to(dtype).cuda(), t=inp["t"].to(dtype).cuda(),
                                     noise=inp["noise"].to(dtype).cuda(), return_velocity=True, **kw)
    return float(loss), v.cpu()


@pytest.mark.parametrize("name", FLOW_CASES)
def test_flow_matching_loss_golden(name, capsys):
    """Loss and velocity of the training forward vs the UNMODIFIED reference (tests/golden/make_flow_golden.py)."""
    cfg, sd, inp, gold = load_case(name)
    B = inp["input_ids"].shape[0]
    pol = _policy(cfg, sd, B)
    pol.cfg = {"flow_sig_min": inp["flow_sig_min"]}  # the binding reads flow_sig_min from its cfg, like PiZero.__init__
    loss, v = _flow(pol, inp)
    loss2, v2 = _flow(pol, inp, masks=True)  # graph replay, mask API
    assert loss == loss2 and torch.equal(v, v2), "the training forward must be bit-identical from call to call"
    e_v = max_err(v, gold["v_psi"])
    with capsys.disabled():
        print(f"\n[{name}] flow loss {loss:.6f} vs reference {float(gold['loss']):.6f}; v_psi max-abs {e_v:.3e}")
    assert e_v <= ACTION_TOL
    assert abs(loss - float(gold["loss"])) <= 5e-3 * float(gold["loss"])
    # the inference entry still gives the inference result afterwards (separate graphs per mode)
    cfg_i, sd_i, inp_i, gold_i = load_case("tiny_bridge" if "bridge" in name else "tiny_robotwin")
    if cfg_i.to_dict() == cfg.to_dict():
        assert max_err(_run(pol, inp_i), gold_i["actions"]) <= ACTION_TOL


def test_flow_matching_loss_full_size_vs_oracle(capsys):
    """Full-size 2B widths, B = 2, bf16 I/O, against the fp32 oracle on the same bf16-rounded weights and inputs."""
    cfg = vlaser_2b_bridge()
    sd = random_state_dict(cfg, 0)
    B = 2
    inp = make_inputs(cfg, B, 21)
    g = torch.Generator().manual_seed(7)
    inp["actions"] = (torch.rand((B, cfg.num_action_tokens, cfg.action_dim), generator=g) * 2 - 1).bfloat16().float()
    inp["t"] = torch.tensor([0.25, 0.8125])
    pol = _policy(cfg, sd, B)
    loss, v = _flow(pol, inp, torch.bfloat16)
    ref_loss, ref_v = O.flow_matching_loss(cfg, sd, inp["input_ids"], inp["pixel_values"].bfloat16().float(),
                                           inp["proprios"].bfloat16().float(), inp["actions"], inp["noise"].bfloat16().float(),
                                           inp["t"], 0.001)
    e_v = max_err(v, ref_v)
    with capsys.disabled():
        print(f"\n[2B flow loss] {loss:.5f} vs oracle {float(ref_loss):.5f}; v_psi max-abs {e_v:.3e}")
    assert e_v <= ACTION_TOL
    assert abs(loss - float(ref_loss)) <= 1e-2 * float(ref_loss)


# ---- chat VLM: causal prefill + greedy decode (modeling_internvl_chat.py:401-439) ----------------
def _real_tokens(cfg, inp, b, n_new):
    """Tokens sample b can emit: its K/V window of S + 1 + H rows holds the prompt and every token that is fed back; the
    last emitted token is not fed back."""
    valid = int((inp["input_ids"][b] != cfg.pad_token_id).sum())
    return min(n_new, cfg.max_image_text_tokens + 1 + cfg.num_action_tokens - valid + 1)


def _check_generation(cfg, sd, inp, toks, first_logits, log, tag):
    """bf16 tensor-core logits against the fp32 oracle, robust to near ties: the oracle is teacher-forced with the tokens
    the GPU emitted, and every emitted token must be within `tol` of the oracle's best logit at that step; the logits
    of the last prompt position are compared directly."""
    n_new = toks.shape[1]
    _, ref = O.generate(cfg, sd, inp["input_ids"], inp["pixel_values"], n_new, forced=toks)
    ref_first = torch.stack([l[0] for l in ref])
    scale = float(ref_first.std())
    e_first = max_err(first_logits, ref_first)
    worst = 0.0
    for b in range(toks.shape[0]):
        for k in range(_real_tokens(cfg, inp, b, n_new)):
            worst = max(worst, float(ref[b][k].max() - ref[b][k][toks[b, k]]))
        assert bool((toks[b, _real_tokens(cfg, inp, b, n_new):] == cfg.pad_token_id).all()), "fill token behind a full window"
    e_rms = float((first_logits - ref_first).pow(2).mean().sqrt())
    log(f"\n[{tag}] first-step logits vs fp32 oracle: max-abs {e_first:.3e} rms {e_rms:.3e} (logit std {scale:.2f}); "
        f"worst oracle-logit gap of an emitted token {worst:.3e}")
    # bf16 operands: well under 1 % of the logit spread rms; the maximum over B x vocab logits sits ~5 sigma out
    assert e_rms <= 1e-2 * scale and e_first <= 5e-2 * scale
    assert worst <= 2e-2 * scale
    return ref


@pytest.mark.parametrize("decode_chunk", [256, 64, 0])
@pytest.mark.parametrize("name", GENERATE_CASES)
def test_generate_golden(name, decode_chunk, capsys):
    """Greedy decode through the C ABI vs the HF / reference golden (tiny configs; untied head, and tied head with dynamic
    tiles 1/3/2/2). decode_chunk: the split-KV decode attention with one chunk per window (256), with two or three (64),
    and the generic attention kernel (0)."""
    cfg, sd, inp, gold = load_case(name)
    B, n_new = gold["tokens"].shape
    pol = PiZeroInference(cfg, max_batch=B)
    pol.load_state_dict(sd, strict=False)
    pol.set_option("decode_chunk", decode_chunk)
    toks, logits = pol.generate(inp["input_ids"].cuda(), inp["pixel_values"].cuda(), max_new_tokens=n_new, return_first_logits=True)
    toks2 = pol.generate(inp["input_ids"].cuda(), inp["pixel_values"].cuda(), max_new_tokens=n_new)  # graph replays
    assert torch.equal(toks, toks2), "generation must be bit-identical from call to call"
    toks, logits = toks.cpu(), logits.cpu()
    scale = float(gold["first_logits"].std())
    assert max_err(logits, gold["first_logits"]) <= 5e-2 * scale
    with capsys.disabled():
        ref = _check_generation(cfg, sd, inp, toks, logits, print, name)
    # where the reference's own margin is clear, the tokens are the reference's
    for b in range(B):
        for k in range(_real_tokens(cfg, inp, b, n_new)):
            top2 = ref[b][k].topk(2).values
            if not torch.equal(toks[b, : k + 1], gold["tokens"][b, : k + 1]):
                assert float(top2[0] - top2[1]) <= 4e-2 * scale, f"sample {b} step {k}: token differs at a clear margin"
                break
    # eos stops a sample: its first emitted token as eos -> everything after it is the fill token
    eos = int(toks[0, 0])
    t3 = pol.generate(inp["input_ids"].cuda(), inp["pixel_values"].cuda(), max_new_tokens=n_new, eos_token_id=eos,
                      fill_token_id=7).cpu()
    assert int(t3[0, 0]) == eos and bool((t3[0, 1:] == 7).all())
    for b in range(1, B):
        if eos not in toks[b].tolist():
            n = _real_tokens(cfg, inp, b, n_new)
            assert torch.equal(t3[b, :n], toks[b, :n]) and bool((t3[b, n:] == 7).all())
    # the policy call still works on the same handle (its own graphs and K/V layout)
    pinp = make_inputs(cfg, B, 5)
    if name == "tiny_generate":
        a = _run(pol, pinp)
        ref_a = O.infer_action(cfg, sd, pinp["input_ids"], pinp["pixel_values"], pinp["proprios"], pinp["noise"])
        assert max_err(a, ref_a) <= ACTION_TOL


def test_generate_text_only_and_window():
    """A text-only prompt (no tile goes through the ViT) and a sample that fills its K/V window."""
    cfg = tiny()
    sd = random_state_dict(cfg, 0)
    sd.update(text_head_state_dict(cfg, 0))
    B, S = 2, cfg.max_image_text_tokens
    g = torch.Generator().manual_seed(3)
    ids = torch.full((B, S), cfg.pad_token_id, dtype=torch.int64)
    ids[0, :7] = torch.randint(0, 400, (7,), generator=g)
    ids[1, :S] = torch.randint(0, 400, (S,), generator=g)  # full prompt: window = S + 1 + H leaves 1 + H new tokens
    pol = PiZeroInference(cfg, max_batch=B)
    pol.load_state_dict(sd, strict=False)
    n_new = 2 + cfg.num_action_tokens + 1
    toks = pol.generate(ids.cuda(), None, max_new_tokens=n_new, fill_token_id=9).cpu()
    _, logits = O.generate(cfg, sd, ids, None, n_new, forced=toks)
    scale = float(logits[0][0].std())
    for k in range(n_new):
        assert float(logits[0][k].max() - logits[0][k][toks[0, k]]) <= 2e-2 * scale
    # sample 1: its window (S + 1 + H rows) takes the prompt and 1 + H fed-back tokens, so 2 + H tokens are emitted (the
    # last one is never fed back); the rest is the fill token
    for k in range(2 + cfg.num_action_tokens):
        assert float(logits[1][k].max() - logits[1][k][toks[1, k]]) <= 2e-2 * scale
    assert bool((toks[1, 2 + cfg.num_action_tokens:] == 9).all())


def test_generate_full_width_reduced_depth(capsys):
    """2B widths (1536 / 12:2 heads / MLP 8960, vocab 151 930) at depth 2, B = 3, bf16 tiles, 12 new tokens."""
    cfg = vlaser_2b_bridge().with_(layers=2, vit_layers=2)
    sd = random_state_dict(cfg, 0)
    sd.update(text_head_state_dict(cfg, 0, tied=True))
    B = 3
    inp = make_inputs(cfg, B, 9)
    pol = PiZeroInference(cfg, max_batch=B)
    pol.load_state_dict(sd, strict=False)
    toks, logits = pol.generate(inp["input_ids"].cuda(), inp["pixel_values"].bfloat16().cuda(), max_new_tokens=12,
                                return_first_logits=True)
    with capsys.disabled():
        _check_generation(cfg, sd, inp, toks.cpu(), logits.cpu(), print, "2B widths depth 2")
