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


def _reference_eval_cfg():
    """Exactly the keys `PiZero.__init__` is handed by Vlaser_VLA/Simpler/config/eval/bridge_internvl_448.yaml:33-85
    (after hydra interpolation), as plain nested dicts."""
    mix = dict(hidden_size=768, intermediate_size=4096, use_final_norm=True, cache=True, use_quantize=False, use_lora=False,
               adaptive_mode=None, rope_theta=10000.0)
    return dict(
        flow_sampling="beta", num_inference_steps=10, final_action_clip_value=1.0, use_torch_compile=True, use_bf16=False,
        cond_steps=1, horizon_steps=4, act_steps=4, action_dim=7, proprio_dim=7,
        mixture=dict(vlm=dict(hidden_size=1536, intermediate_size=16384, use_final_norm=False, cache=True, use_quantize=False,
                              use_lora=False, adaptive_mode=None, rope_theta=10000.0),
                     proprio=dict(mix), action=dict(mix, cache=False)),
        action_expert_adaptive_mode=None, time_hidden_size=256, time_max_period=10000.0, action_expert_rope_theta=10000.0,
        quantize=False, lora=False, lora_r=32, lora_dropout=0.0, max_image_text_tokens=384,
        image_token_index=151667, vocab_size=151667, pad_token_id=151643,
    )


def test_config_from_reference_eval_yaml_keys():
    """The drop-in constructor takes the reference's eval cfg as it is: every key that shapes `infer_action` is mapped,
    and settings the library does not implement raise instead of returning the actions of another model."""
    import pytest
    from vlaser_b200.policy import check_supported_cfg, config_from_reference_cfg
    cfg = _reference_eval_cfg()
    v = config_from_reference_cfg(cfg)
    assert (v.max_image_text_tokens, v.cond_steps, v.horizon_steps, v.action_dim, v.proprio_dim) == (384, 1, 4, 7, 7)
    assert v.num_inference_steps == 10 and v.final_action_clip_value == 1.0 and v.time_max_period == 10000.0
    assert (v.image_token_index, v.pad_token_id, v.hidden_e) == (151667, 151643, 768)
    # the yaml's vocab_size (151667) is NOT the embedding row count: the table follows the checkpoint tensor
    assert v.embed_rows > v.image_token_index
    # the InternVL path takes the LLM widths from model_size, not from the PaliGemma-era mixture.vlm keys
    assert (v.hidden, v.mlp) == (1536, 8960)
    assert config_from_reference_cfg(dict(cfg, model_size="8B")).hidden == 3584
    # attribute-style cfg objects (omegaconf) work too
    ns = types.SimpleNamespace(**{k: (types.SimpleNamespace(**{kk: types.SimpleNamespace(**vv) for kk, vv in val.items()})
                                      if k == "mixture" else val) for k, val in cfg.items()})
    assert config_from_reference_cfg(ns) == v
    for bad in (dict(integration_method="rk4"), dict(action_expert_adaptive_mode="adaLN"), dict(lora=True), dict(quantize=True)):
        with pytest.raises(NotImplementedError):
            check_supported_cfg(dict(cfg, **bad))
    broken = _reference_eval_cfg()
    broken["mixture"]["proprio"]["hidden_size"] = 1024
    with pytest.raises(ValueError):
        check_supported_cfg(broken)
    check_supported_cfg(dict(cfg, integration_method="euler"))
