"""CPU: host-side logic — config mapping, state-dict layout, the vectorised mask builder, synthetic
inputs, the FLOP model."""
from __future__ import annotations

import types

import torch

from oracle import vlaser_oracle as O
from vlaser_b200.config import PRESETS, tiny, vlaser_2b_bridge
from vlaser_b200.synth import make_inputs
from vlaser_b200.weights import canonical_key, random_state_dict, state_dict_spec


def test_presets_dims():
    c = vlaser_2b_bridge()
    assert (c.grid, c.patches, c.vit_tokens, c.image_tokens, c.num_action_tokens) == (16, 256, 257, 64, 4)
    assert c.q_dim == 1536 and c.kv_dim == 256
    r = PRESETS["vlaser_2b_robotwin"]()
    assert r.num_images == 3 and r.num_action_tokens == 6 and r.max_image_text_tokens == 896


def test_state_dict_spec_param_count():
    spec = state_dict_spec(vlaser_2b_bridge())
    n = sum(int(torch.tensor(s).prod()) for s, _ in spec.values())
    # ViT 0.303 B + LLM blocks 1.310 B + expert blocks 0.655 B + embeddings + heads
    assert 2.4e9 < n < 2.6e9
    assert canonical_key("_orig_mod.joint_model.mixtures.proprio.layers.3.mlp.up_proj.weight") == \
        "joint_model.mixtures.action.layers.3.mlp.up_proj.weight"
    assert canonical_key("internvl_model.mlp1.1.weight") == "multi_modal_projector.1.weight"


def test_random_state_dict_is_deterministic_and_bf16_exact():
    a = random_state_dict(tiny(), 5)
    b = random_state_dict(tiny(), 5)
    for k in a:
        assert torch.equal(a[k], b[k])
        assert torch.equal(a[k], a[k].to(torch.bfloat16).float())


def test_synthetic_inputs_layout():
    cfg = vlaser_2b_bridge()
    inp = make_inputs(cfg, 4, 1)
    assert inp["valid_len"].tolist() == [85, 86, 87, 89]           # 17 + 64 + prompt tokens
    assert inp["pixel_values"].shape == (4, 3, 224, 224)
    ids = inp["input_ids"]
    assert (ids[0, 10:74] == cfg.image_token_index).all() and (ids[0, 85:] == cfg.pad_token_id).all()


def test_policy_mask_builder_matches_reference_semantics():
    """The vectorised builder in policy.py produces the same mask/ids as the reference's loop
    (restated in the oracle)."""
    from vlaser_b200.policy import PiZeroInference
    cfg = tiny()
    shim = types.SimpleNamespace(num_action_tokens=cfg.num_action_tokens,
                                 max_image_text_tokens=cfg.max_image_text_tokens, num_proprio_tokens=1)
    am = torch.zeros(3, cfg.max_image_text_tokens, dtype=torch.int64)
    for b, n in enumerate((70, 96, 1)):
        am[b, :n] = 1
    m, vp, pp, ap = PiZeroInference.build_causal_mask_and_position_ids(shim, am, torch.float32)
    m2, vp2, pp2, ap2 = O.build_causal_mask_and_position_ids(am, cfg.max_image_text_tokens, cfg.num_action_tokens)
    assert torch.equal(m, m2) and torch.equal(vp, vp2) and torch.equal(pp, pp2) and torch.equal(ap, ap2)
    a, b = PiZeroInference.split_full_mask_into_submasks(shim, m)
    a2, b2 = O.split_full_mask_into_submasks(m2, cfg.max_image_text_tokens, cfg.num_action_tokens)
    assert torch.equal(a, a2) and torch.equal(b, b2)
    assert ((a[:, 0, 0, : cfg.max_image_text_tokens] == 0).sum(-1) == am.sum(-1)).all()
