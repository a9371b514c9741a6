"""CPU: pins the oracle restatement (oracle/vlaser_oracle.py) against the golden vectors produced by
the unmodified reference, checks the known-answer invariants of SURVEY.md §8c, and (when
/root/reference is present) compares it with the live reference on fresh inputs."""
from __future__ import annotations

import pytest
import torch

from oracle import ref_harness
from oracle import vlaser_oracle as O
from tests.golden_util import CASES, DYNAMIC_TILE_CASES, FLOW_CASES, GENERATE_CASES, load_case
from vlaser_b200.config import tiny
from vlaser_b200.synth import make_inputs
from vlaser_b200.weights import random_state_dict


@pytest.mark.parametrize("name", CASES + DYNAMIC_TILE_CASES)
def test_oracle_matches_golden(name):
    cfg, sd, inp, gold = load_case(name)
    tr = {}
    act = O.infer_action(cfg, sd, inp["input_ids"], inp["pixel_values"], inp["proprios"], inp["noise"], trace=tr)
    assert act.shape == gold["actions"].shape
    assert (act - gold["actions"]).abs().max() < 2e-5
    pre = tr["actions_preclamp"][:, -cfg.horizon_steps:]
    assert (pre - gold["trace_actions_preclamp_sliced"]).abs().max() < 2e-5
    for k in ("vit_out", "image_features", "expert_out_step0", "k_vlm_layer0", "v_vlm_layer0",
              f"k_vlm_layer{cfg.layers - 1}", f"k_proprio_layer{cfg.layers - 1}", f"v_proprio_layer{cfg.layers - 1}"):
        assert (tr[k] - gold["trace_" + k]).abs().max() < 2e-5, k


@pytest.mark.parametrize("name", FLOW_CASES)
def test_oracle_flow_loss_matches_golden(name):
    """Training forward (PiZero.forward, pizero_internvl.py:1064-1197): the prefill + one-expert-pass restatement gives the
    loss and velocity of the reference's joint forward without KV cache."""
    cfg, sd, inp, gold = load_case(name)
    loss, v = O.flow_matching_loss(cfg, sd, inp["input_ids"], inp["pixel_values"], inp["proprios"], inp["actions"],
                                   inp["noise"], inp["t"], inp["flow_sig_min"])
    assert (v - gold["v_psi"]).abs().max() < 2e-5
    assert abs(float(loss) - float(gold["loss"])) < 1e-5 * max(1.0, float(gold["loss"]))


@pytest.mark.parametrize("name", GENERATE_CASES)
def test_oracle_generate_matches_golden(name):
    """Chat-VLM greedy decode: the causal-LM restatement reproduces the tokens and first-step logits of the unmodified HF
    Qwen2ForCausalLM.generate fed by the reference's own prompt embedding (make_generate_golden.py)."""
    cfg, sd, inp, gold = load_case(name)
    n_new = gold["tokens"].shape[1]
    toks, logits = O.generate(cfg, sd, inp["input_ids"], inp["pixel_values"], n_new)
    assert torch.equal(toks, gold["tokens"])
    first = torch.stack([l[0] for l in logits])
    assert (first - gold["first_logits"]).abs().max() < 1e-4 * max(1.0, float(gold["first_logits"].abs().max()))
    # teacher forcing with the same tokens reproduces the free-running logits
    _, forced = O.generate(cfg, sd, inp["input_ids"], inp["pixel_values"], n_new, forced=toks)
    assert all(torch.equal(a, b) for a, b in zip(logits, forced))


def _tiny_case(B=2, seed=5, prompt_tokens=(5, 11)):
    cfg = tiny()
    sd = random_state_dict(cfg, 3)
    inp = make_inputs(cfg, B, seed, list(prompt_tokens))
    return cfg, sd, inp


def test_batch_row_independence():
    """KAT (3): row b of a batched call == the B=1 call with the same noise row."""
    cfg, sd, inp = _tiny_case()
    full = O.infer_action(cfg, sd, inp["input_ids"], inp["pixel_values"], inp["proprios"], inp["noise"])
    one = O.infer_action(cfg, sd, inp["input_ids"][1:2], inp["pixel_values"][1:2], inp["proprios"][1:2], inp["noise"][1:2])
    assert (full[1:2] - one).abs().max() < 1e-5


def test_pad_length_independence():
    """KAT (2): the output does not depend on how far the prefix is padded."""
    cfg, sd, inp = _tiny_case()
    a = O.infer_action(cfg, sd, inp["input_ids"], inp["pixel_values"], inp["proprios"], inp["noise"])
    cfg2 = cfg.with_(max_image_text_tokens=cfg.max_image_text_tokens + 32)
    ids2 = torch.full((2, cfg2.max_image_text_tokens), cfg.pad_token_id, dtype=torch.int64)
    ids2[:, : cfg.max_image_text_tokens] = inp["input_ids"]
    b = O.infer_action(cfg2, sd, ids2, inp["pixel_values"], inp["proprios"], inp["noise"])
    assert (a - b).abs().max() < 1e-5


def test_pixel_shuffle_closed_form():
    """KAT (4): pixel_shuffle == the closed-form index map, bit exact, g = 16 and 32."""
    for g in (16, 32):
        C = 8
        x = torch.arange(2 * g * g * C, dtype=torch.float32).view(2, g, g, C)
        out = O.pixel_shuffle(x).reshape(2, (g // 2) ** 2, 4, C)
        idx = O.pixel_shuffle_index_map(g, C)
        ref = x.view(2, g * g, C)[:, idx]  # [2, T, 4, C]
        assert torch.equal(out, ref)


def test_prefix_merge_rules():
    """KAT (5): image rows equal projector output in row-major order, pad rows are exactly zero."""
    cfg, sd, inp = _tiny_case()
    tr = {}
    emb = O.prefix_embedding(cfg, sd, inp["input_ids"], inp["pixel_values"], tr)
    ids = inp["input_ids"]
    assert torch.equal(emb[ids == cfg.pad_token_id], torch.zeros_like(emb[ids == cfg.pad_token_id]))
    assert torch.equal(emb[ids == cfg.image_token_index], tr["image_features"].flatten(0, 1))


def test_mask_builder_matches_block_rule():
    cfg = tiny()
    am = torch.zeros(2, cfg.max_image_text_tokens, dtype=torch.int64)
    am[0, :70] = 1
    am[1, :90] = 1
    m, vp, pp, ap = O.build_causal_mask_and_position_ids(am, cfg.max_image_text_tokens, cfg.num_action_tokens)
    S = cfg.max_image_text_tokens
    assert m.shape == (2, 1, S + 5, S + 5)
    assert (m[0, 0, 0, :70] == 0).all() and (m[0, 0, 0, 70:] < 0).all()
    assert m[0, 0, S, S] == 0 and (m[0, 0, S, S + 1:] < 0).all()          # proprio sees itself, not actions
    assert (m[0, 0, S + 1:, S:] == 0).all() and (m[0, 0, 5, S:] < 0).all()  # actions see proprio+actions
    assert vp[0, 0] == 1 and pp[0, 0] == 1 and ap[0].tolist() == [2, 3, 4, 5]


@pytest.mark.skipif(not ref_harness.available(), reason="/root/reference not present (GPU box)")
def test_oracle_matches_live_reference():
    cfg, sd, inp = _tiny_case(B=2, seed=9, prompt_tokens=(4, 13))
    ref_act, ref_tr = ref_harness.run_reference(cfg, sd, inp["input_ids"], inp["pixel_values"], inp["proprios"], inp["noise"])
    tr = {}
    act = O.infer_action(cfg, sd, inp["input_ids"], inp["pixel_values"], inp["proprios"], inp["noise"], trace=tr)
    assert (act - ref_act).abs().max() < 2e-5
    assert (tr["vit_out"] - ref_tr["vit_out"]).abs().max() < 2e-5
    assert (tr["expert_out_step0"] - ref_tr["expert_out_step0"]).abs().max() < 2e-5


@pytest.mark.skipif(not ref_harness.available(), reason="/root/reference not present (GPU box)")
def test_oracle_matches_live_reference_dynamic_tiles():
    """Dynamic tiles: samples hold 1 / 3 / 2 / 2 image tiles (8 = B * num_images in total); the
    reference's masked scatter hands the projector rows out in batch order, and so does the oracle."""
    from vlaser_b200.config import tiny
    from vlaser_b200.synth import make_inputs
    from vlaser_b200.weights import random_state_dict
    from oracle import vlaser_oracle as O
    cfg = tiny().with_(cond_steps=2, max_image_text_tokens=224)
    sd = random_state_dict(cfg, 3)
    inp = make_inputs(cfg, 4, 11, tiles_per_sample=[1, 3, 2, 2])
    ref_act, _ = ref_harness.run_reference(cfg, sd, inp["input_ids"], inp["pixel_values"], inp["proprios"], inp["noise"])
    got = O.infer_action(cfg, sd, inp["input_ids"], inp["pixel_values"], inp["proprios"], inp["noise"])
    assert float((got - ref_act).abs().max()) <= 1e-5
