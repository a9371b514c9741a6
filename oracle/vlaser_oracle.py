"""ORACLE — test infrastructure, not product code.

CPU fp32 restatement (plain torch ops, no HF / hydra / timm imports) of the reference's
`PiZero.infer_action` hot path. Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import this module; the product path
(`vlaser_b200/`) never does and fails loudly without its CUDA library.

Parity pin: the reference holds no golden vectors for this path, so this
restatement is pinned against outputs of the reference's own code executed in the build
container: `tests/golden/make_golden.py` imports `/root/reference` (unmodified
`PiZero.infer_action`, `JointModel.forward`, `InternVisionModel.forward`, HF Qwen2 primitives),
runs it on seeded weights/inputs and commits the vectors under `tests/golden/*.npz`;
`tests/test_oracle.py` checks this file against them (and, when `/root/reference` is present,
against the live reference on fresh inputs).

Every function cites the reference lines it restates. `S/` = Vlaser_VLA/Simpler/,
`S/vla/` = S/src/model/vla/, `S/ivl/` = S/src/model/internvl_chat/internvl/model/internvl_chat/.
Third-party arithmetic (HF transformers 4.48 Qwen2RMSNorm / Qwen2MLP / Qwen2RotaryEmbedding /
eager_attention_forward / repeat_kv) is restated from its published definition; call sites:
`S/vla/joint_model.py:166,209,219,450-452,574-578,631-656,694`.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def build_causal_mask_and_position_ids(attention_mask: torch.Tensor, S: int, H: int, dtype=torch.float32):
    """S/vla/pizero_internvl.py:517-587 (non-debug branch). attention_mask: [B,S] 0/1."""
    bsz = attention_mask.size(0)
    dev = attention_mask.device
    P0, P1 = S, S + 1
    cnts = attention_mask.sum(dim=1)
    L = S + H + 1
    m = torch.full((bsz, L, L), torch.finfo(dtype).min, dtype=dtype, device=dev)
    for i, c in enumerate(cnts.tolist()):
        m[i, :c, :c] = 0
        m[i, P0:, :c] = 0
    m[:, P0:P1, P0:P1] = 0
    m[:, P1:, P0:] = 0
    m = m.unsqueeze(1)
    vlm_pos = torch.arange(1, S + 1, device=dev).repeat(bsz, 1)
    proprio_pos = torch.arange(1, 2, device=dev).repeat(bsz, 1)
    action_pos = torch.arange(2, H + 2, device=dev).repeat(bsz, 1)
    return m, vlm_pos, proprio_pos, action_pos


def split_full_mask_into_submasks(mask: torch.Tensor, S: int, H: int):
    """S/vla/pizero_internvl.py:589-603 (4-D branch)."""
    return mask[..., : S + 1, : S + 1], mask[..., -H:, :]


def pixel_shuffle(x: torch.Tensor, scale_factor: float = 0.5) -> torch.Tensor:
    """S/vla/pizero_internvl.py:706-716 (== S/ivl/modeling_internvl_chat.py:256-270, v2)."""
    n, w, h, c = x.size()
    x = x.view(n, w, int(h * scale_factor), int(c / scale_factor))
    x = x.permute(0, 2, 1, 3).contiguous()
    x = x.view(n, int(h * scale_factor), int(w * scale_factor), int(c / (scale_factor * scale_factor)))
    x = x.permute(0, 2, 1, 3).contiguous()
    return x


def pixel_shuffle_index_map(g: int, C: int) -> torch.Tensor:
    """Closed form of `pixel_shuffle`: returns int64 [g/2*g/2, 4] patch
    indices; token r2*(g/2)+c2 = concat of patches (2r2,2c2),(2r2,2c2+1),(2r2+1,2c2),(2r2+1,2c2+1)."""
    idx = torch.empty((g // 2) * (g // 2), 4, dtype=torch.int64)
    for r2 in range(g // 2):
        for c2 in range(g // 2):
            t = r2 * (g // 2) + c2
            for b in range(2):
                for a in range(2):
                    idx[t, b * 2 + a] = (2 * r2 + b) * g + (2 * c2 + a)
    return idx


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """HF Qwen2RMSNorm: fp32 inside, cast to input dtype before the gain."""
    xf = x.to(torch.float32)
    var = xf.pow(2).mean(-1, keepdim=True)
    return w * (xf * torch.rsqrt(var + eps)).to(x.dtype)


def rope_cos_sin(pos: torch.Tensor, head_dim: int, theta: float):
    """HF Qwen2RotaryEmbedding (default rope): inv_freq = theta^(-2i/d); emb = cat(freqs, freqs)."""
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64, device=pos.device).to(torch.float32) / head_dim))
    freqs = pos.to(torch.float32)[..., None] * inv  # [B, L, d/2]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    x1 = x[..., : x.shape[-1] // 2]
    x2 = x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """S/vla/joint_model.py:583,597: x [B,heads,L,d], cos/sin [B,L,d] (unsqueeze_dim=1)."""
    cos, sin = cos.to(x.dtype), sin.to(x.dtype)  # HF casts the fp32 table to the activation dtype
    return x * cos.unsqueeze(1) + rotate_half(x) * sin.unsqueeze(1)


def eager_attention(q, k, v, mask, scaling: float, groups: int):
    """HF eager_attention_forward + repeat_kv: softmax_fp32(q k^T * scale + mask) v."""
    k = k.repeat_interleave(groups, dim=1)
    v = v.repeat_interleave(groups, dim=1)
    w = torch.matmul(q, k.transpose(2, 3)) * scaling
    if mask is not None:
        w = w + mask[:, :, :, : k.shape[-2]]
    w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
    o = torch.matmul(w, v)
    return o.transpose(1, 2).contiguous()  # [B, L, heads, d]


# --------------------------------------------------------------------------------------
# model pieces
# --------------------------------------------------------------------------------------
def vit_forward(cfg, sd: Dict[str, torch.Tensor], pixel_values: torch.Tensor, trace: Optional[dict] = None):
    """S/ivl/modeling_intern_vit.py:162-174 (embeddings), :291-293 (layer), :210-227 (naive attn),
    :259-263 (MLP). Returns last_hidden_state [N,1+P,d]."""
    v = "vision_tower.vision_model."
    w = sd[v + "embeddings.patch_embedding.weight"]
    b = sd[v + "embeddings.patch_embedding.bias"]
    x = F.conv2d(pixel_values, w, b, stride=cfg.patch_size)          # [N,d,g,g]
    n = x.shape[0]
    x = x.flatten(2).transpose(1, 2)                                 # [N,P,d]
    cls = sd[v + "embeddings.class_embedding"].expand(n, 1, -1)
    x = torch.cat([cls, x], dim=1)
    # _get_pos_embed is the identity when the grid matches (bicubic resize to same size)
    x = x + sd[v + "embeddings.position_embedding"]
    if trace is not None:
        trace["vit_embed"] = x.clone()
    nh = cfg.vit_heads
    hd = cfg.vit_hidden // nh
    for i in range(cfg.vit_layers):
        p = f"{v}encoder.layers.{i}."
        h = F.layer_norm(x, (cfg.vit_hidden,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], cfg.vit_eps)
        B_, N_, C_ = h.shape
        qkv = F.linear(h, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])
        qkv = qkv.reshape(B_, N_, 3, nh, hd).permute(2, 0, 3, 1, 4)
        q, k, vv = qkv.unbind(0)
        attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
        attn = attn.softmax(dim=-1)
        a = (attn @ vv).transpose(1, 2).reshape(B_, N_, C_)
        a = F.linear(a, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
        x = x + a * sd[p + "ls1"]
        h = F.layer_norm(x, (cfg.vit_hidden,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], cfg.vit_eps)
        h = F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
        h = F.gelu(h)
        h = F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
        x = x + h * sd[p + "ls2"]
        if trace is not None:
            trace[f"vit_layer{i}"] = x.clone()
    return x


def projector(cfg, sd, shuffled: torch.Tensor) -> torch.Tensor:
    """mlp1: S/ivl/modeling_internvl_chat.py:89-94 — LN(4d) -> Linear -> GELU -> Linear."""
    d4 = 4 * cfg.vit_hidden
    h = F.layer_norm(shuffled, (d4,), sd["multi_modal_projector.0.weight"], sd["multi_modal_projector.0.bias"], cfg.proj_eps)
    h = F.linear(h, sd["multi_modal_projector.1.weight"], sd["multi_modal_projector.1.bias"])
    h = F.gelu(h)
    return F.linear(h, sd["multi_modal_projector.3.weight"], sd["multi_modal_projector.3.bias"])


def prefix_embedding(cfg, sd, input_ids: torch.Tensor, pixel_values: torch.Tensor, trace: Optional[dict] = None):
    """S/vla/pizero_internvl.py:718-796 (`_forward_siglip_and_text_embedding`)."""
    emb = F.embedding(input_ids, sd["embed_tokens.weight"])
    vit = vit_forward(cfg, sd, pixel_values, trace)
    feats = vit[:, 1:]
    g = cfg.grid
    feats = feats.reshape(feats.shape[0], g, g, feats.shape[-1])
    sh = pixel_shuffle(feats)
    img = projector(cfg, sd, sh)
    img = img.reshape(img.shape[0], img.shape[1] * img.shape[2], -1)
    if trace is not None:
        trace["vit_out"] = vit.clone()
        trace["image_features"] = img.clone()
    bsz, seq = input_ids.shape
    final = torch.zeros(bsz, seq, emb.shape[-1], dtype=emb.dtype, device=emb.device)
    text_mask = (input_ids != cfg.image_token_index) & (input_ids != cfg.pad_token_id)
    image_mask = input_ids == cfg.image_token_index
    final[text_mask] = emb[text_mask]
    final[image_mask] = img.flatten(0, 1)
    return final


def _attn_block(cfg, sd, prefix: str, li: int, x: torch.Tensor):
    p = f"{prefix}layers.{li}."
    h = rms_norm(x, sd[p + "input_layernorm.weight"], cfg.rms_eps)
    B_, L_, _ = h.shape
    q = F.linear(h, sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.q_proj.bias"]).view(B_, L_, -1, cfg.head_dim).transpose(1, 2)
    k = F.linear(h, sd[p + "self_attn.k_proj.weight"], sd[p + "self_attn.k_proj.bias"]).view(B_, L_, -1, cfg.head_dim).transpose(1, 2)
    v = F.linear(h, sd[p + "self_attn.v_proj.weight"], sd[p + "self_attn.v_proj.bias"]).view(B_, L_, -1, cfg.head_dim).transpose(1, 2)
    return q, k, v


def _post_attn(cfg, sd, prefix: str, li: int, x: torch.Tensor, attn_out: torch.Tensor):
    """S/vla/joint_model.py:196-231: o_proj, residual, post-norm, SwiGLU MLP, residual."""
    p = f"{prefix}layers.{li}."
    x = x + F.linear(attn_out, sd[p + "self_attn.o_proj.weight"])
    h = rms_norm(x, sd[p + "post_attention_layernorm.weight"], cfg.rms_eps)
    gate = F.linear(h, sd[p + "mlp.gate_proj.weight"])
    up = F.linear(h, sd[p + "mlp.up_proj.weight"])
    return x + F.linear(F.silu(gate) * up, sd[p + "mlp.down_proj.weight"])


VLM = "joint_model.mixtures.vlm."
EXP = "joint_model.mixtures.action."


def prefill(cfg, sd, vlm_embeds, proprio_embeds, mask, vlm_pos, proprio_pos, trace: Optional[dict] = None):
    """S/vla/joint_model.py:740-814 with embeds {vlm, proprio}, return_caches=True;
    layer body :140-232, attention :410-696. Returns per-layer (K,V) caches for both mixtures."""
    groups = cfg.n_heads // cfg.n_kv_heads
    scaling = cfg.head_dim ** -0.5
    cos_v, sin_v = rope_cos_sin(vlm_pos, cfg.head_dim, cfg.rope_theta)
    cos_p, sin_p = rope_cos_sin(proprio_pos, cfg.head_dim, cfg.rope_theta)
    xv, xp = vlm_embeds, proprio_embeds
    S = xv.shape[1]
    caches = {"vlm": [], "proprio": []}
    for li in range(cfg.layers):
        last = li == cfg.layers - 1
        qv, kv, vv = _attn_block(cfg, sd, VLM, li, xv)
        qp, kp, vp = _attn_block(cfg, sd, EXP, li, xp)
        kv = apply_rope(kv, cos_v, sin_v)
        qv = apply_rope(qv, cos_v, sin_v)
        kp = apply_rope(kp, cos_p, sin_p)
        qp = apply_rope(qp, cos_p, sin_p)
        caches["vlm"].append((kv, vv))
        caches["proprio"].append((kp, vp))
        if last:
            break  # post_attn_skip_names = (vlm, proprio): outputs discarded (:192-193, 690-691)
        q = torch.cat((qv, qp), dim=-2)
        k = torch.cat((kv, kp), dim=-2)
        v = torch.cat((vv, vp), dim=2)
        o = eager_attention(q, k, v, mask, scaling, groups)
        o = o.reshape(o.shape[0], o.shape[1], -1)
        ov, op = o[:, :S], o[:, S:]
        xv = _post_attn(cfg, sd, VLM, li, xv, ov)
        xp = _post_attn(cfg, sd, EXP, li, xp, op)
        if trace is not None:
            trace[f"prefill_vlm_layer{li}"] = xv.clone()
            trace[f"prefill_proprio_layer{li}"] = xp.clone()
    return caches


def time_embedding(cfg, t: torch.Tensor) -> torch.Tensor:
    """S/vla/modules.py:9-22 SinusoidalPosEmb(hidden_e, time_max_period)."""
    half = cfg.hidden_e // 2
    e = math.log(cfg.time_max_period) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=t.dtype, device=t.device) * -e)
    e = t[:, None] * e[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


def action_encoder(sd, action: torch.Tensor, time_emb: torch.Tensor) -> torch.Tensor:
    """S/vla/modules.py:25-53 ActionEncoder(time_cond=True)."""
    e = F.linear(action, sd["action_encoder.linear_1.weight"], sd["action_encoder.linear_1.bias"])
    tf = time_emb.unsqueeze(1).expand(-1, action.size(1), -1)
    e = torch.cat([tf, e], dim=-1)
    e = F.silu(F.linear(e, sd["action_encoder.linear_2.weight"], sd["action_encoder.linear_2.bias"]))
    return F.linear(e, sd["action_encoder.linear_3.weight"], sd["action_encoder.linear_3.bias"])


def denoise_step(cfg, sd, x, caches, action_mask, action_pos):
    """One pass of the action expert over cached prefix K/V (cache_mode="append_non_active"):
    S/vla/joint_model.py:461-464, 622-624, 646-656; final norm :804-808."""
    groups = cfg.n_heads // cfg.n_kv_heads
    scaling = cfg.head_dim ** -0.5
    cos, sin = rope_cos_sin(action_pos, cfg.head_dim, cfg.rope_theta)
    for li in range(cfg.layers):
        q, k, v = _attn_block(cfg, sd, EXP, li, x)
        k = apply_rope(k, cos, sin)
        q = apply_rope(q, cos, sin)
        kc = torch.cat((caches["vlm"][li][0], caches["proprio"][li][0], k), dim=-2)
        vc = torch.cat((caches["vlm"][li][1], caches["proprio"][li][1], v), dim=2)
        o = eager_attention(q, kc, vc, action_mask, scaling, groups)
        o = o.reshape(o.shape[0], o.shape[1], -1)
        x = _post_attn(cfg, sd, EXP, li, x, o)
    return rms_norm(x, sd["joint_model.mixtures.action.norm.weight"], cfg.rms_eps)


def infer_action(cfg, sd: Dict[str, torch.Tensor], input_ids: torch.Tensor, pixel_values: torch.Tensor,
                 proprios: torch.Tensor, noise: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                 trace: Optional[dict] = None) -> torch.Tensor:
    """S/vla/pizero_internvl.py:798-936. `noise` replaces the `torch.randn` draw at :879-881.
    attention_mask defaults to `input_ids != pad_token_id` (what the tokenizer emits).
    Returns normalised actions [B, horizon_steps, A]; trace (if given) receives intermediates
    incl. `actions_preclamp`."""
    S = cfg.max_image_text_tokens
    H = cfg.num_action_tokens
    assert input_ids.shape[1] == S
    if attention_mask is None:
        attention_mask = (input_ids != cfg.pad_token_id).to(torch.int64)
    wdt = sd["proprio_encoder.weight"].dtype  # fp32 on the CPU (parity oracle); bf16 for bench.py's GPU eager baseline leg
    mask, vlm_pos, proprio_pos, action_pos = build_causal_mask_and_position_ids(attention_mask, S, H, dtype=wdt)
    itp_mask, action_mask = split_full_mask_into_submasks(mask, S, H)

    embeds = prefix_embedding(cfg, sd, input_ids, pixel_values, trace)
    proprio_embeds = F.linear(proprios, sd["proprio_encoder.weight"], sd["proprio_encoder.bias"])
    if trace is not None:
        trace["prefix_embeds"] = embeds.clone()
        trace["proprio_embeds"] = proprio_embeds.clone()
    caches = prefill(cfg, sd, embeds, proprio_embeds, itp_mask, vlm_pos, proprio_pos, trace)
    if trace is not None:
        for li in (0, cfg.layers - 1):
            trace[f"k_vlm_layer{li}"] = caches["vlm"][li][0].clone()
            trace[f"v_vlm_layer{li}"] = caches["vlm"][li][1].clone()
            trace[f"k_proprio_layer{li}"] = caches["proprio"][li][0].clone()
            trace[f"v_proprio_layer{li}"] = caches["proprio"][li][1].clone()

    action = noise.clone().to(wdt)
    bsz = action.shape[0]
    dt = 1.0 / cfg.num_inference_steps
    t = torch.zeros(bsz, dtype=wdt, device=action.device)
    for step in range(cfg.num_inference_steps):
        temb = time_embedding(cfg, t)
        x = action_encoder(sd, action, temb)
        y = denoise_step(cfg, sd, x, caches, action_mask, action_pos)
        vel = F.linear(y, sd["action_decoder.weight"], sd["action_decoder.bias"])
        action = action + dt * vel
        t = t + dt
        if trace is not None:
            trace[f"action_step{step}"] = action.clone()
            if step == 0:
                trace["expert_out_step0"] = y.clone()
    if trace is not None:
        trace["actions_preclamp"] = action.clone()
    if cfg.final_action_clip_value is not None:
        action = torch.clamp(action, -cfg.final_action_clip_value, cfg.final_action_clip_value)
    return action[:, -cfg.horizon_steps:]


def flow_matching_loss(cfg, sd: Dict[str, torch.Tensor], input_ids: torch.Tensor, pixel_values: torch.Tensor,
                       proprios: torch.Tensor, actions: torch.Tensor, noise: torch.Tensor, t: torch.Tensor,
                       flow_sig_min: float = 0.001, attention_mask: Optional[torch.Tensor] = None):
    """S/vla/pizero_internvl.py:1064-1197 (`PiZero.forward`, is_vlm=False, imgfeat off), forward only. `noise` replaces the
    `torch.randn_like` draw at :1094. The reference runs the three mixtures jointly without KV cache under the block mask
    of build_causal_mask_and_position_ids; no block attends to a later one, so the prefix / proprio K/V are those of
    `prefill` and the action block is one `denoise_step` over them (psi_t :1053-1062, loss :1190-1193).
    Returns (loss, v_psi)."""
    S = cfg.max_image_text_tokens
    H = cfg.num_action_tokens
    if attention_mask is None:
        attention_mask = (input_ids != cfg.pad_token_id).to(torch.int64)
    wdt = sd["proprio_encoder.weight"].dtype
    mask, vlm_pos, proprio_pos, action_pos = build_causal_mask_and_position_ids(attention_mask, S, H, dtype=wdt)
    itp_mask, action_mask = split_full_mask_into_submasks(mask, S, H)
    x0, x1 = noise.to(wdt), actions.to(wdt)
    tt = t.to(wdt)[:, None, None]
    psi = (1 - (1 - flow_sig_min) * tt) * x0 + tt * x1
    embeds = prefix_embedding(cfg, sd, input_ids, pixel_values)
    proprio_embeds = F.linear(proprios, sd["proprio_encoder.weight"], sd["proprio_encoder.bias"])
    caches = prefill(cfg, sd, embeds, proprio_embeds, itp_mask, vlm_pos, proprio_pos)
    x = action_encoder(sd, psi, time_embedding(cfg, t.to(wdt)))
    y = denoise_step(cfg, sd, x, caches, action_mask, action_pos)
    v_psi = F.linear(y, sd["action_decoder.weight"], sd["action_decoder.bias"])
    d_psi = x1 - (1 - flow_sig_min) * x0
    return torch.mean((v_psi - d_psi) ** 2), v_psi


# --------------------------------------------------------------------------------------
# chat VLM: causal LLM over the merged prompt embeddings, greedy decode
# --------------------------------------------------------------------------------------
def causal_lm_logits(cfg, sd, embeds: torch.Tensor) -> torch.Tensor:
    """HF Qwen2ForCausalLM.forward(inputs_embeds=...) restated (transformers Qwen2Model / Qwen2DecoderLayer /
    eager_attention_forward with the causal mask; call sites: V/internvl/model/internvl_chat/modeling_internvl_chat.py:
    430-437 `language_model.generate(inputs_embeds=...)`, S/vla/pizero_internvl.py:1027-1040 `infer_text`). embeds
    [B, L, hidden] without padding; positions 0..L-1. Returns logits [B, L, vocab]. Weights: the `vlm` mixture layers,
    `joint_model.mixtures.vlm.norm.weight`, and `lm_head.weight` (or the tied `embed_tokens.weight`)."""
    groups = cfg.n_heads // cfg.n_kv_heads
    scaling = cfg.head_dim ** -0.5
    B_, L_, _ = embeds.shape
    pos = torch.arange(L_, device=embeds.device).unsqueeze(0).expand(B_, -1)
    cos, sin = rope_cos_sin(pos, cfg.head_dim, cfg.rope_theta)
    mask = torch.full((L_, L_), torch.finfo(embeds.dtype).min, dtype=embeds.dtype, device=embeds.device).triu(1)[None, None]
    x = embeds
    for li in range(cfg.layers):
        q, k, v = _attn_block(cfg, sd, VLM, li, x)
        q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
        o = eager_attention(q, k, v, mask, scaling, groups)
        x = _post_attn(cfg, sd, VLM, li, x, o.reshape(B_, L_, -1))
    x = rms_norm(x, sd["joint_model.mixtures.vlm.norm.weight"], cfg.rms_eps)
    return F.linear(x, sd.get("lm_head.weight", sd["embed_tokens.weight"]))


def generate(cfg, sd, input_ids: torch.Tensor, pixel_values: Optional[torch.Tensor], max_new_tokens: int,
             eos_token_id: int = -1, fill_token_id: int = 0, forced: Optional[torch.Tensor] = None):
    """Greedy decode of the chat VLM: V/.../modeling_internvl_chat.py:401-439 (image rows scattered into the token
    embeddings, then HF generate with do_sample=False). No KV cache here: every step re-runs the causal forward over the
    sample's whole sequence (an oracle; the sizes are tiny). `forced` [B, max_new_tokens] (optional): teacher forcing --
    the tokens fed back are the given ones, so that the per-step logits can be compared with a run whose argmax went
    another way on a near tie. Returns (tokens [B, max_new_tokens], step_logits: list over samples of [steps, vocab])."""
    if pixel_values is not None:
        embeds = prefix_embedding(cfg, sd, input_ids, pixel_values)
    else:
        embeds = F.embedding(input_ids.clamp(max=cfg.embed_rows - 1), sd["embed_tokens.weight"])
    valid = (input_ids != cfg.pad_token_id).sum(-1).tolist()
    B_ = input_ids.shape[0]
    tokens = torch.full((B_, max_new_tokens), fill_token_id, dtype=torch.int64)
    all_logits = []
    for b in range(B_):
        seq = embeds[b:b + 1, : valid[b]]
        rows = []
        for k in range(max_new_tokens):
            logits = causal_lm_logits(cfg, sd, seq)[0, -1]
            rows.append(logits)
            tok = int(torch.argmax(logits))
            tokens[b, k] = tok
            fed = int(forced[b, k]) if forced is not None else tok
            if forced is None and tok == eos_token_id:
                break
            seq = torch.cat([seq, sd["embed_tokens.weight"][fed].view(1, 1, -1).to(seq.dtype)], dim=1)
        all_logits.append(torch.stack(rows))
    return tokens, all_logits
