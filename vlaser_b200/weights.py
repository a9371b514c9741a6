"""State-dict key layout of the reference policy and a deterministic random initialiser.

The key names are the ones `PiZero` (reference `pizero_internvl.py:253-334`) exposes after
grafting the InternVL3 modules into the joint model, i.e. what `EvalAgent.load_checkpoint`
(`Vlaser_VLA/Simpler/src/agent/eval.py:196-212`) feeds to `load_state_dict`:

  vision_tower.vision_model.embeddings.{class_embedding,position_embedding,patch_embedding.*}
  vision_tower.vision_model.encoder.layers.N.{attn.qkv,attn.proj,mlp.fc1,mlp.fc2,norm1,norm2}.*, ls1, ls2
  multi_modal_projector.{0,1,3}.*
  embed_tokens.weight
  joint_model.mixtures.vlm.layers.N.{self_attn.{q,k,v,o}_proj,mlp.{gate,up,down}_proj,*layernorm}.*
  joint_model.mixtures.action.layers.N....    (proprio shares the same modules: `:255-256`)
  joint_model.mixtures.action.norm.weight
  action_encoder.linear_{1,2,3}.*, proprio_encoder.*, action_decoder.*

Checkpoints also carry aliases of the same tensors (`joint_model.mixtures.proprio.*`,
`internvl_model.*`, an `_orig_mod.` prefix from torch.compile); `canonical_key` folds them.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Tuple

import torch

from .config import VlaserConfig


def state_dict_spec(cfg: VlaserConfig) -> "OrderedDict[str, Tuple[Tuple[int, ...], str]]":
    """key -> (shape, init kind). Init kinds: n02 (N(0,.02)), randn, one, zero, ls, ku:<fan_in>."""
    s: "OrderedDict[str, Tuple[Tuple[int, ...], str]]" = OrderedDict()
    d, f = cfg.vit_hidden, cfg.vit_mlp
    v = "vision_tower.vision_model."
    s[v + "embeddings.class_embedding"] = ((1, 1, d), "randn")
    s[v + "embeddings.position_embedding"] = ((1, cfg.vit_tokens, d), "randn")
    s[v + "embeddings.patch_embedding.weight"] = ((d, 3, cfg.patch_size, cfg.patch_size), f"ku:{3 * cfg.patch_size ** 2}")
    s[v + "embeddings.patch_embedding.bias"] = ((d,), f"ku:{3 * cfg.patch_size ** 2}")
    for i in range(cfg.vit_layers):
        p = f"{v}encoder.layers.{i}."
        s[p + "norm1.weight"] = ((d,), "one")
        s[p + "norm1.bias"] = ((d,), "n02")
        s[p + "attn.qkv.weight"] = ((3 * d, d), "n02")
        s[p + "attn.qkv.bias"] = ((3 * d,), "n02")
        s[p + "attn.proj.weight"] = ((d, d), "n02")
        s[p + "attn.proj.bias"] = ((d,), "n02")
        s[p + "ls1"] = ((d,), "ls")
        s[p + "norm2.weight"] = ((d,), "one")
        s[p + "norm2.bias"] = ((d,), "n02")
        s[p + "mlp.fc1.weight"] = ((f, d), "n02")
        s[p + "mlp.fc1.bias"] = ((f,), "n02")
        s[p + "mlp.fc2.weight"] = ((d, f), "n02")
        s[p + "mlp.fc2.bias"] = ((d,), "n02")
        s[p + "ls2"] = ((d,), "ls")
    h = cfg.hidden
    s["multi_modal_projector.0.weight"] = ((4 * d,), "one")
    s["multi_modal_projector.0.bias"] = ((4 * d,), "n02")
    s["multi_modal_projector.1.weight"] = ((h, 4 * d), f"ku:{4 * d}")
    s["multi_modal_projector.1.bias"] = ((h,), f"ku:{4 * d}")
    s["multi_modal_projector.3.weight"] = ((h, h), f"ku:{h}")
    s["multi_modal_projector.3.bias"] = ((h,), f"ku:{h}")
    s["embed_tokens.weight"] = ((cfg.embed_rows, h), "n02")

    def llm_block(prefix: str, hid: int, mlp: int):
        for i in range(cfg.layers):
            p = f"{prefix}layers.{i}."
            s[p + "input_layernorm.weight"] = ((hid,), "one")
            s[p + "self_attn.q_proj.weight"] = ((cfg.q_dim, hid), "n02")
            s[p + "self_attn.q_proj.bias"] = ((cfg.q_dim,), "n02")
            s[p + "self_attn.k_proj.weight"] = ((cfg.kv_dim, hid), "n02")
            s[p + "self_attn.k_proj.bias"] = ((cfg.kv_dim,), "n02")
            s[p + "self_attn.v_proj.weight"] = ((cfg.kv_dim, hid), "n02")
            s[p + "self_attn.v_proj.bias"] = ((cfg.kv_dim,), "n02")
            s[p + "self_attn.o_proj.weight"] = ((hid, cfg.q_dim), "n02")
            s[p + "post_attention_layernorm.weight"] = ((hid,), "one")
            s[p + "mlp.gate_proj.weight"] = ((mlp, hid), "n02")
            s[p + "mlp.up_proj.weight"] = ((mlp, hid), "n02")
            s[p + "mlp.down_proj.weight"] = ((hid, mlp), "n02")

    llm_block("joint_model.mixtures.vlm.", cfg.hidden, cfg.mlp)
    llm_block("joint_model.mixtures.action.", cfg.hidden_e, cfg.mlp_e)
    he = cfg.hidden_e
    s["joint_model.mixtures.action.norm.weight"] = ((he,), "one")
    s["action_encoder.linear_1.weight"] = ((he, cfg.action_dim), f"ku:{cfg.action_dim}")
    s["action_encoder.linear_1.bias"] = ((he,), f"ku:{cfg.action_dim}")
    s["action_encoder.linear_2.weight"] = ((he, 2 * he), f"ku:{2 * he}")
    s["action_encoder.linear_2.bias"] = ((he,), f"ku:{2 * he}")
    s["action_encoder.linear_3.weight"] = ((he, he), f"ku:{he}")
    s["action_encoder.linear_3.bias"] = ((he,), f"ku:{he}")
    s["proprio_encoder.weight"] = ((he, cfg.proprio_dim), f"ku:{cfg.proprio_dim}")
    s["proprio_encoder.bias"] = ((he,), f"ku:{cfg.proprio_dim}")
    s["action_decoder.weight"] = ((cfg.action_dim, he), f"ku:{he}")
    s["action_decoder.bias"] = ((cfg.action_dim,), f"ku:{he}")
    return s


def random_state_dict(cfg: VlaserConfig, seed: int = 0, bf16_round: bool = True) -> Dict[str, torch.Tensor]:
    """Deterministic random-init weights (fp32 tensors on CPU).

    Module-default style init: Qwen2 linears/embeddings N(0,0.02); InternViT
    CLS/pos randn, layer-scale 0.1; `nn.Linear` default U(+-1/sqrt(fan_in)) for the small heads.
    Norm gains are 1 + N(0,0.02) and biases are non-zero so that a dropped bias/gain shows up in
    parity tests. With `bf16_round` the values are rounded to bf16 (and returned as fp32) so that
    the fp32 oracle and the bf16 device path see *identical* weights (SURVEY.md §7 hard parts).
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for key, (shape, kind) in state_dict_spec(cfg).items():
        if kind == "n02":
            t = torch.randn(shape, generator=g) * 0.02
        elif kind == "randn":
            t = torch.randn(shape, generator=g)
        elif kind == "one":
            t = 1.0 + torch.randn(shape, generator=g) * 0.02
        elif kind == "ls":
            t = 0.1 + torch.randn(shape, generator=g) * 0.002
        elif kind.startswith("ku:"):
            bound = 1.0 / float(kind[3:]) ** 0.5
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        else:  # pragma: no cover
            raise ValueError(kind)
        if bf16_round:
            t = t.to(torch.bfloat16).to(torch.float32)
        out[key] = t
    return out


def text_head_state_dict(cfg: VlaserConfig, seed: int = 0, tied: bool = False, bf16_round: bool = True) -> Dict[str, torch.Tensor]:
    """The two tensors the chat-VLM text path reads on top of `random_state_dict` (they are optional for the policy
    call, which never reaches the LLM's final norm or head: pizero_internvl.py:192-193): the final RMSNorm gain of the
    `vlm` mixture and, unless `tied` (Qwen2.5-1.5B ties the head to the embedding table), `lm_head.weight`. The head is
    drawn wide (N(0, 0.5)) so that the greedy argmax of a random-init model has a clear margin."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000 + seed)
    out = {"joint_model.mixtures.vlm.norm.weight": 1.0 + torch.randn((cfg.hidden,), generator=g) * 0.1}
    if not tied:
        out["lm_head.weight"] = torch.randn((cfg.embed_rows, cfg.hidden), generator=g) * 0.5
    if bf16_round:
        out = {k: v.to(torch.bfloat16).to(torch.float32) for k, v in out.items()}
    return out


_ALIASES = (
    ("_orig_mod.", ""),
    ("joint_model.mixtures.proprio.", "joint_model.mixtures.action."),
    ("internvl_model.vision_model.", "vision_tower.vision_model."),
    ("internvl_model.mlp1.", "multi_modal_projector."),
    ("internvl_model.language_model.model.embed_tokens.", "embed_tokens."),
    ("internvl_model.language_model.model.layers.", "joint_model.mixtures.vlm.layers."),
    ("internvl_model.action_expert.model.layers.", "joint_model.mixtures.action.layers."),
    ("internvl_model.action_expert.model.norm.", "joint_model.mixtures.action.norm."),
    ("internvl_model.language_model.model.norm.", "joint_model.mixtures.vlm.norm."),
    ("internvl_model.language_model.lm_head.", "lm_head."),
    ("language_model.lm_head.", "lm_head."),
)


def canonical_key(key: str) -> str:
    """Map checkpoint aliases of a tensor to the canonical key used by `state_dict_spec`."""
    for a, b in _ALIASES:
        if key.startswith(a):
            key = b + key[len(a):]
    return key
