"""ORACLE support — test infrastructure, build-container only.

Imports the UNMODIFIED reference (`/root/reference`, read-only) with import stubs for packages
that are absent here (hydra, omegaconf, timm, bitsandbytes), bypasses the two constructors that
need a HF-hub download (`PiZero.__init__`, `JointModel.__init__`) by setting exactly the
attributes they would (`pizero_internvl.py:207-336`,
`joint_model.py:707-734`), and runs the reference's own `infer_action`. Used by
`tests/golden/make_golden.py` to generate the committed golden vectors and by
`tests/test_oracle.py` (skipped when `/root/reference` is absent, e.g. on the GPU box).
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys
import types
from typing import Dict

import torch
from torch import nn

REF_ROOT = "/root/reference"
SIMPLER = os.path.join(REF_ROOT, "Vlaser_VLA/Simpler")
IVL = os.path.join(SIMPLER, "src/model/internvl_chat/internvl/model/internvl_chat")


def available() -> bool:
    return os.path.isdir(SIMPLER)


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []  # behave as a package
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_loaded = {}


def load_reference():
    """Returns namespace with the reference modules (pizero_internvl, joint_model, intern_vit)."""
    if _loaded:
        return _loaded
    import transformers  # noqa: F401  (must be imported before the stubs)

    class _Identity(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

        def forward(self, x):
            return x

    class _OmegaConf:
        @staticmethod
        def merge(*a, **k):
            raise RuntimeError("stub")

    _stub("omegaconf", OmegaConf=_OmegaConf, open_dict=lambda *a, **k: None)
    _stub("hydra")
    _stub("hydra.utils")
    bnb = _stub("bitsandbytes")
    bnb.nn = _stub("bitsandbytes.nn", Params4bit=type("Params4bit", (), {}), Linear4bit=type("Linear4bit", (nn.Module,), {}))
    _stub("timm")
    _stub("timm.models")
    _stub("timm.models.layers", DropPath=_Identity)
    os.environ["INTERNVL"] = "1"
    if SIMPLER not in sys.path:
        sys.path.insert(0, SIMPLER)
    import src.model.vla.pizero_internvl as pz
    import src.model.vla.joint_model as jm
    import src.model.vla.mixture as mx

    pkg = _stub("_ref_ivl")
    pkg.__path__ = [IVL]

    def _load(modname, fname):
        spec = importlib.util.spec_from_file_location(f"_ref_ivl.{modname}", os.path.join(IVL, fname))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"_ref_ivl.{modname}"] = mod
        spec.loader.exec_module(mod)
        return mod

    cfg_mod = _load("configuration_intern_vit", "configuration_intern_vit.py")
    vit_mod = _load("modeling_intern_vit", "modeling_intern_vit.py")
    _loaded.update(pz=pz, jm=jm, mx=mx, cfg_mod=cfg_mod, vit_mod=vit_mod)
    return _loaded


class _VisionTower(nn.Module):
    """Same forward as the SiglipVisionModel shell (`S/src/model/paligemma/siglip.py:303-320`)."""

    def __init__(self, vision_model):
        super().__init__()
        self.vision_model = vision_model

    def forward(self, pixel_values):
        return self.vision_model(pixel_values=pixel_values)


def build_reference_policy(cfg, sd: Dict[str, torch.Tensor], parts: dict = None):
    """Reference `PiZero` instance (fp32, CPU, eval) holding the weights in `sd`. `parts` (optional dict) receives the HF
    `Qwen2ForCausalLM` objects whose decoder layers the joint model holds ("llm", "expert")."""
    from transformers import Qwen2Config, Qwen2ForCausalLM

    ref = load_reference()
    pz, jm, mx = ref["pz"], ref["jm"], ref["mx"]
    IVConfig = ref["cfg_mod"].InternVisionConfig
    IVModel = ref["vit_mod"].InternVisionModel

    def qcfg(hidden, inter):
        return Qwen2Config(hidden_size=hidden, intermediate_size=inter, num_hidden_layers=cfg.layers,
                           num_attention_heads=cfg.n_heads, num_key_value_heads=cfg.n_kv_heads,
                           head_dim=cfg.head_dim, rope_theta=cfg.rope_theta, rms_norm_eps=cfg.rms_eps,
                           max_position_embeddings=32768, vocab_size=cfg.embed_rows,
                           attn_implementation="eager")

    with torch.device("meta"):
        llm = Qwen2ForCausalLM(qcfg(cfg.hidden, cfg.mlp))
        expert = Qwen2ForCausalLM(qcfg(cfg.hidden_e, cfg.mlp_e))
    # rotary embeddings carry only a non-persistent buffer; rebuild them for real on CPU
    llm = llm.to_empty(device="cpu")
    expert = expert.to_empty(device="cpu")
    from transformers.models.qwen2.modeling_qwen2 import Qwen2RotaryEmbedding
    llm.model.rotary_emb = Qwen2RotaryEmbedding(config=llm.config)
    expert.model.rotary_emb = Qwen2RotaryEmbedding(config=expert.config)

    vcfg = IVConfig(hidden_size=cfg.vit_hidden, num_attention_heads=cfg.vit_heads,
                    intermediate_size=cfg.vit_mlp, num_hidden_layers=cfg.vit_layers,
                    image_size=cfg.image_size, patch_size=cfg.patch_size, qkv_bias=True,
                    qk_normalization=False, norm_type="layer_norm", use_flash_attn=False,
                    initializer_factor=0.1, layer_norm_eps=cfg.vit_eps, drop_path_rate=0.0)
    vit = IVModel(vcfg)

    def shell(layers, norm):
        m = mx.Mixture.__new__(mx.Mixture)
        nn.Module.__init__(m)
        m.layers = layers
        if norm is not None:
            m.norm = norm
        m.adaptive_mode = None
        return m

    joint = jm.JointModel.__new__(jm.JointModel)
    nn.Module.__init__(joint)
    joint.num_hidden_layers = cfg.layers
    joint.cache_names = ["vlm", "proprio"]
    joint.backbone_type = "INTERNVL"
    joint.use_internvl = True
    joint.use_flash_attention = False
    joint.stop_grad_to_vlm = False
    act = shell(expert.model.layers, expert.model.norm)
    joint.mixtures = nn.ModuleDict(dict(vlm=shell(llm.model.layers, llm.model.norm), proprio=act, action=act))
    joint.mixture_names = ["vlm", "proprio", "action"]

    pol = pz.PiZero.__new__(pz.PiZero)
    nn.Module.__init__(pol)
    H = cfg.num_action_tokens
    pol.cfg = types.SimpleNamespace(horizon_steps=cfg.horizon_steps)
    pol.imgfeat = False
    pol.image_448 = False
    pol.no_img = False
    pol.debug_causal = False
    pol.NO_CAUSAL_IMG = False
    pol.num_images = cfg.num_images
    pol.max_image_text_tokens = cfg.max_image_text_tokens
    pol.num_proprio_tokens = 1
    pol.num_action_tokens = H
    pol.horizon_steps = H
    pol.action_dim = cfg.action_dim
    pol.num_inference_steps = cfg.num_inference_steps
    pol.final_action_clip_value = cfg.final_action_clip_value
    pol.integration_method = "euler"
    pol.action_expert_adaptive_mode = None
    pol.image_token_index = cfg.image_token_index
    pol.pad_token_id = cfg.pad_token_id
    pol.joint_model = joint
    pol.vision_tower = _VisionTower(vit)
    d4 = 4 * cfg.vit_hidden
    pol.multi_modal_projector = nn.Sequential(nn.LayerNorm(d4), nn.Linear(d4, cfg.hidden), nn.GELU(),
                                              nn.Linear(cfg.hidden, cfg.hidden))
    pol.embed_tokens = nn.Embedding(cfg.embed_rows, cfg.hidden, cfg.pad_token_id)
    ns = types.SimpleNamespace(
        language_model=types.SimpleNamespace(model=types.SimpleNamespace(rotary_emb=llm.model.rotary_emb)),
        action_expert=types.SimpleNamespace(model=types.SimpleNamespace(rotary_emb=expert.model.rotary_emb)),
    )
    object.__setattr__(pol, "internvl_model", ns)
    pol.action_encoder = pz.ActionEncoder(cfg.action_dim, cfg.hidden_e, time_cond=True)
    pol.time_embedding = pz.SinusoidalPosEmb(cfg.hidden_e, cfg.time_max_period)
    pol.proprio_encoder = nn.Linear(cfg.proprio_dim, cfg.hidden_e)
    pol.action_decoder = nn.Linear(cfg.hidden_e, cfg.action_dim)

    # the 224-px branch hard-codes a 16x16 grid (pizero_internvl.py:739-743); the tiny config
    # uses a different grid, so only there the reshape is re-pointed at cfg.grid via image_448.
    if cfg.grid not in (16, 32):
        raise ValueError("reference supports only 16x16 / 32x32 patch grids; use patched_grid()")
    pol.image_448 = cfg.grid == 32

    # (the harness instance carries no `lm_head` module -- pizero_internvl.py:325-334 -- run_reference_generate loads it into
    # the HF model itself)
    missing, unexpected = pol.load_state_dict({k: v for k, v in sd.items() if k != "lm_head.weight"}, strict=False)
    assert not unexpected, f"unexpected keys: {unexpected[:5]}"
    allowed_missing = ("joint_model.mixtures.vlm.norm.weight",)
    real_missing = [k for k in missing if not k.endswith(allowed_missing) and "proprio" not in k]
    assert not real_missing, f"missing keys: {real_missing[:8]}"
    # vlm final norm is never used on the policy path (skip rule); give it a value when the weights have none
    if "joint_model.mixtures.vlm.norm.weight" not in sd:
        with torch.no_grad():
            llm.model.norm.weight.fill_(1.0)
    if parts is not None:
        parts.update(llm=llm, expert=expert)
    pol.eval()
    return pol


def run_reference(cfg, sd, input_ids, pixel_values, proprios, noise, capture: bool = True):
    """Runs the reference `infer_action` (fp32 CPU) with `noise` injected in place of its
    `torch.randn` draw. Returns (actions, trace)."""
    pol = build_reference_policy(cfg, sd)
    return run_reference_policy(pol, cfg, input_ids, pixel_values, proprios, noise, capture)


def run_reference_policy(pol, cfg, input_ids, pixel_values, proprios, noise, capture: bool = True):
    trace = {}
    attention_mask = (input_ids != cfg.pad_token_id).to(torch.int64)
    mask, vp, pp, ap = pol.build_causal_mask_and_position_ids(attention_mask, torch.float32)
    itp, am = pol.split_full_mask_into_submasks(mask)
    handles = []
    if capture:
        def hook_vit(mod, args, out):
            trace["vit_out"] = out["last_hidden_state"].detach().clone()
        handles.append(pol.vision_tower.register_forward_hook(hook_vit))

        def hook_proj(mod, args, out):
            trace["image_features"] = out.detach().reshape(out.shape[0], -1, out.shape[-1]).clone()
        handles.append(pol.multi_modal_projector.register_forward_hook(hook_proj))

        def hook_dec(mod, args, out):
            if "expert_out_step0" not in trace:
                trace["expert_out_step0"] = args[0].detach().clone()
        handles.append(pol.action_decoder.register_forward_hook(hook_dec))

    real_randn = torch.randn
    clip = pol.final_action_clip_value

    def fake_randn(*a, **k):
        return noise.clone().to(k.get("dtype", torch.float32))

    jm = load_reference()["jm"]
    caches_seen = {}
    real_build = pol.joint_model.build_mixture_caches

    def build_and_keep():
        c = real_build()
        caches_seen.update(c)
        return c

    pol.joint_model.build_mixture_caches = build_and_keep
    try:
        torch.randn = fake_randn
        with torch.inference_mode():
            pol.final_action_clip_value = None
            pre = pol.infer_action(input_ids, pixel_values, itp, am, vp, pp, ap, proprios).clone()
            pol.final_action_clip_value = clip
    finally:
        torch.randn = real_randn
        pol.final_action_clip_value = clip
        del pol.joint_model.build_mixture_caches
        for h in handles:
            h.remove()
    if clip is not None:
        actions = torch.clamp(pre, -clip, clip)
    else:
        actions = pre
    # NOTE: pre-clamp tensor is already sliced to the last horizon_steps tokens.
    trace["actions_preclamp_sliced"] = pre
    if capture and caches_seen:
        for li in (0, cfg.layers - 1):
            for name in ("vlm", "proprio"):
                k, v = caches_seen[name].get(li)
                trace[f"k_{name}_layer{li}"] = k.detach().clone()
                trace[f"v_{name}_layer{li}"] = v.detach().clone()
    return actions, trace


def run_reference_flow_loss(cfg, sd, input_ids, pixel_values, proprios, actions, noise, t, flow_sig_min: float = 0.001):
    """Runs the reference training `PiZero.forward(..., is_vlm=False)` (pizero_internvl.py:1064-1197, fp32 CPU, no
    grad) with `noise` injected in place of its `torch.randn_like` draw. Returns (loss, v_psi)."""
    pol = build_reference_policy(cfg, sd)
    pol.flow_sig_min = flow_sig_min
    attention_mask = (input_ids != cfg.pad_token_id).to(torch.int64)
    mask, vp, pp, ap = pol.build_causal_mask_and_position_ids(attention_mask, torch.float32)
    seen = {}
    h = pol.action_decoder.register_forward_hook(lambda mod, args, out: seen.__setitem__("v_psi", out.detach().clone()))
    real = torch.randn_like
    try:
        torch.randn_like = lambda a, **k: noise.clone().to(k.get("dtype", a.dtype))
        with torch.inference_mode():
            loss = pol.forward(input_ids, pixel_values, mask, vp, pp, ap, proprios, actions, t)
    finally:
        torch.randn_like = real
        h.remove()
    return loss.detach().clone(), seen["v_psi"]


def run_reference_generate(cfg, sd, input_ids, pixel_values, max_new_tokens: int):
    """Greedy decode of the chat VLM with the reference's modules: the merged prompt embeddings come from the reference's
    own `PiZero._forward_siglip_and_text_embedding` (pizero_internvl.py:718-796 -- the same extract_feature + masked scatter
    as `InternVLChatModel.generate`, V/.../modeling_internvl_chat.py:415-428), and the text is produced by the unmodified HF
    `Qwen2ForCausalLM.generate(inputs_embeds=..., do_sample=False, use_cache=True)` that `generate` calls at :430-437, one
    sample at a time (no padding). `sd` must hold `joint_model.mixtures.vlm.norm.weight` and optionally `lm_head.weight`
    (else tied to the embedding table). Returns (tokens [B, max_new_tokens], first_logits [B, vocab])."""
    parts = {}
    pol = build_reference_policy(cfg, sd, parts)
    llm = parts["llm"]
    with torch.no_grad():
        llm.model.embed_tokens.weight.copy_(sd["embed_tokens.weight"])
        llm.lm_head.weight.copy_(sd.get("lm_head.weight", sd["embed_tokens.weight"]))
    llm.eval()
    llm.generation_config.pad_token_id = cfg.pad_token_id
    llm.generation_config.eos_token_id = None
    valid = (input_ids != cfg.pad_token_id).sum(-1).tolist()
    toks, first = [], []
    with torch.inference_mode():
        emb = pol._forward_siglip_and_text_embedding(input_ids, pixel_values)
        for b in range(input_ids.shape[0]):
            e = emb[b:b + 1, : valid[b]]
            first.append(llm(inputs_embeds=e).logits[0, -1].clone())
            out = llm.generate(inputs_embeds=e, attention_mask=torch.ones(e.shape[:2], dtype=torch.int64),
                               max_new_tokens=max_new_tokens, min_new_tokens=max_new_tokens, do_sample=False, use_cache=True)
            toks.append(out[0, -max_new_tokens:].clone())
    return torch.stack(toks), torch.stack(first)
