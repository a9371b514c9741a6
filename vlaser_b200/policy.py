"""Drop-in policy object for the reference's `PiZeroInference`
(`Vlaser_VLA/Simpler/src/model/vla/pizero_internvl.py:1286-1307`), backed by the sm_100a library
through the C ABI (include/vlaser_b200.h). Same call surface as the reference class as used by
`EvalAgent` (`Vlaser_VLA/Simpler/src/agent/eval.py:33-56, 111-139`):

    model = PiZeroInference(cfg)                       # cfg: hydra-style cfg of the reference or VlaserConfig
    model.load_state_dict(sd, strict=False); model.freeze_all_weights()
    model.to(dtype); model.to(device); model.eval()
    mask, vp, pp, ap = model.build_causal_mask_and_position_ids(attention_mask, dtype)
    itp_mask, action_mask = model.split_full_mask_into_submasks(mask)
    actions = model(input_ids, pixel_values, itp_mask, action_mask, vp, pp, ap, proprios)   # [B, horizon, A]

PyTorch is used only for device memory / streams of the caller's tensors; there is no PyTorch
compute path and no fallback: a missing library or GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os
from collections import namedtuple
from typing import Dict, Iterable, Optional, Tuple

import torch

from . import _lib
from .config import VlaserConfig
from .weights import canonical_key, state_dict_spec

_IncompatibleKeys = namedtuple("IncompatibleKeys", ["missing_keys", "unexpected_keys"])


def _cfg_get(cfg, name, default=None):
    if hasattr(cfg, "get"):
        try:
            v = cfg.get(name, default)
            return default if v is None else v
        except TypeError:
            pass
    return getattr(cfg, name, default)


def _nested_get(cfg, path, default=None):
    cur = cfg
    for name in path:
        if cur is None:
            return default
        cur = _cfg_get(cur, name, None)
    return default if cur is None else cur


def check_supported_cfg(cfg) -> None:
    """Raises for cfg settings that change the numbers of `infer_action` and that this library does not implement,
    instead of silently returning actions of a different model:
      * `integration_method` other than "euler" (pizero_internvl.py:164, 910-922, 1309-1334)
      * a non-empty `action_expert_adaptive_mode` (adaLN / adaLN-Zero time conditioning, :292-305, 890-905)
      * LoRA / quantised mixtures (:`lora`, `quantize`)
      * `mixture.{proprio,action}.hidden_size` that disagree with each other (the two mixtures share one expert)"""
    if isinstance(cfg, VlaserConfig):
        return
    meth = str(_cfg_get(cfg, "integration_method", "euler"))
    if meth != "euler":
        raise NotImplementedError(f"integration_method={meth!r}: vlaser_b200 implements the Euler flow integration only "
                                  "(pizero_internvl.py:910-912)")
    mode = _cfg_get(cfg, "action_expert_adaptive_mode", None)
    if mode:
        raise NotImplementedError(f"action_expert_adaptive_mode={mode!r}: adaLN time conditioning is not implemented "
                                  "(released Vlaser-VLA checkpoints use None, config/eval/*internvl*.yaml)")
    for key in ("lora", "quantize"):
        if _cfg_get(cfg, key, False):
            raise NotImplementedError(f"{key}=True: LoRA / quantised mixtures are outside the inference hot path of vlaser_b200")
    hp = _nested_get(cfg, ("mixture", "proprio", "hidden_size"))
    ha = _nested_get(cfg, ("mixture", "action", "hidden_size"))
    if hp is not None and ha is not None and int(hp) != int(ha):
        raise ValueError(f"mixture.proprio.hidden_size={hp} != mixture.action.hidden_size={ha}: the proprio and action "
                         "mixtures share one expert (pizero_internvl.py:1213-1232)")


def config_from_reference_cfg(cfg) -> VlaserConfig:
    """Maps the hydra cfg keys `PiZero.__init__` reads (pizero_internvl.py:156-337) plus the
    `IMAGE_448` env switch (:180-183) and `model_size` (:116-135) onto a VlaserConfig. The LLM / expert widths of
    the InternVL path come from `model_size` (the `mixture.*.hidden_size` / `intermediate_size` keys of the yaml are
    the PaliGemma leftovers the reference overrides at :116-135); the expert hidden width is taken from
    `mixture.action.hidden_size` when present. `vocab_size` of the yaml (151667) is not the embedding row count --
    the reference resizes the table to len(tokenizer) (:88-89, 335) -- so embed_rows follows the checkpoint tensor
    (`PiZeroInference.load_state_dict`)."""
    if isinstance(cfg, VlaserConfig):
        return cfg
    check_supported_cfg(cfg)
    model_size = str(_cfg_get(cfg, "model_size", "2B"))
    kw = {}
    if model_size == "8B":  # InternVL3-8B: Qwen2.5-7B LLM, expert MLP 5600 (pizero_internvl.py:128-131)
        kw.update(hidden=3584, n_heads=28, n_kv_heads=4, mlp=18944, mlp_e=5600)
    he = _nested_get(cfg, ("mixture", "action", "hidden_size"))
    if he is not None:
        kw["hidden_e"] = int(he)
    clip = _cfg_get(cfg, "final_action_clip_value", None)
    return VlaserConfig(
        image_size=448 if "IMAGE_448" in os.environ else 224,
        max_image_text_tokens=int(_cfg_get(cfg, "max_image_text_tokens")),
        cond_steps=int(_cfg_get(cfg, "cond_steps")),
        horizon_steps=int(_cfg_get(cfg, "horizon_steps")),
        action_dim=int(_cfg_get(cfg, "action_dim")),
        proprio_dim=int(_cfg_get(cfg, "proprio_dim")),
        num_inference_steps=int(_cfg_get(cfg, "num_inference_steps")),
        final_action_clip_value=None if clip is None else float(clip),
        time_max_period=float(_cfg_get(cfg, "time_max_period", 10000.0)),
        image_token_index=int(_cfg_get(cfg, "image_token_index", 151667)),
        pad_token_id=int(_cfg_get(cfg, "pad_token_id", 151643)),
        **kw,
    )


def _to_c_cfg(cfg: VlaserConfig, max_batch: int) -> _lib.VlaserCfg:
    c = _lib.VlaserCfg()
    for name, _ in _lib.VlaserCfg._fields_:
        if name in ("clip_enabled", "final_action_clip_value", "max_batch"):
            continue
        setattr(c, name, getattr(cfg, name))
    c.clip_enabled = 0 if cfg.final_action_clip_value is None else 1
    c.final_action_clip_value = 0.0 if cfg.final_action_clip_value is None else float(cfg.final_action_clip_value)
    c.max_batch = int(max_batch)
    return c


class PiZeroInference:
    """B200-native `PiZeroInference` (inference only)."""

    def __init__(self, cfg, use_ddp: bool = False, max_batch: Optional[int] = None, device: Optional[int] = None):
        self.cfg = cfg
        self.vcfg = config_from_reference_cfg(cfg)
        self.use_ddp = use_ddp
        v = self.vcfg
        # attributes the reference exposes and callers read (pizero_internvl.py:207-224, 336)
        self.max_image_text_tokens = v.max_image_text_tokens
        self.num_proprio_tokens = 1
        self.num_action_tokens = v.num_action_tokens
        self.horizon_steps = v.num_action_tokens
        self.action_dim = v.action_dim
        self.proprio_dim = v.proprio_dim
        self.num_inference_steps = v.num_inference_steps
        self.final_action_clip_value = v.final_action_clip_value
        self.image_token_index = v.image_token_index
        self.pad_token_id = v.pad_token_id
        self.num_images = v.num_images
        self.training = False

        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("vlaser_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.max_batch = int(max_batch if max_batch is not None else _cfg_get(cfg, "max_batch", 64))
        self._handle = C.c_void_p()
        self._options: Dict[str, int] = {}
        self._finalized = False
        self._loaded = set()
        # The library computes in bf16 with fp32 accumulation / residual streams whatever dtype is asked for (see `to`)
        self.dtype = torch.bfloat16
        for env, opt in (("VLASER_PDL", "pdl"), ("VLASER_EXP_STAGES", "exp_stages"), ("VLASER_FUSE_NORM", "fuse_norm"),
                         ("VLASER_PAIR_GEMM", "pair_gemm"), ("VLASER_EXPERT_MEGA", "expert_mega"),
                         ("VLASER_EXPERT_STREAMS", "expert_streams"), ("VLASER_EXPERT_GEMV", "expert_gemv"),
                         ("VLASER_L2_PREFETCH", "l2_prefetch"), ("VLASER_SMALL_M_SPLIT", "small_m_split"),
                         ("VLASER_PAIR_MIN_TILES", "pair_min_tiles"), ("VLASER_PROP_STREAM", "prop_stream"),
                         ("VLASER_DET_SPLITK", "det_splitk"), ("VLASER_VIT_TC", "vit_tc"),
                         ("VLASER_PREFILL_TC_LONG", "prefill_tc_long")):  # measurement switches
            if env in os.environ:
                self._options[opt] = int(os.environ[env])
        self._create_handle()

    def _create_handle(self):
        """(Re)creates the library handle for self.vcfg and replays the options set so far."""
        if self._handle.value:
            self._lib.vlaser_destroy(self._handle)
            self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            ccfg = _to_c_cfg(self.vcfg, self.max_batch)
            _lib.check(self._lib.vlaser_create(C.byref(ccfg), C.byref(self._handle)), "vlaser_create")
        for k, val in self._options.items():
            _lib.check(self._lib.vlaser_set_option(self._handle, k.encode(), int(val)), f"set_option({k})")

    # ---- nn.Module-like surface used by EvalAgent ------------------------------------------------
    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = True):
        """eval.py:196-212 — feeds every tensor of the path to the library by its state-dict key."""
        if self._finalized:
            raise RuntimeError("weights are frozen after the first forward")
        # the embedding table has as many rows as the checkpoint's tokenizer had tokens (the reference resizes it to
        # len(tokenizer), pizero_internvl.py:88-89, 335): follow the tensor, not a constant
        for key, t in state_dict.items():
            if torch.is_tensor(t) and canonical_key(key) == "embed_tokens.weight" and t.dim() == 2 \
                    and int(t.shape[0]) != self.vcfg.embed_rows:
                if self._loaded:
                    raise RuntimeError(f"embed_tokens.weight has {t.shape[0]} rows but the handle was built for "
                                       f"{self.vcfg.embed_rows}: load the embedding table first")
                if int(t.shape[0]) <= max(self.vcfg.image_token_index, self.vcfg.pad_token_id):
                    raise ValueError(f"embed_tokens.weight has {t.shape[0]} rows: fewer than the image / pad token ids")
                self.vcfg = self.vcfg.with_(embed_rows=int(t.shape[0]))
                self._create_handle()
        unexpected = []
        with torch.cuda.device(self.device):
            for key, t in state_dict.items():
                if not torch.is_tensor(t):
                    continue
                t = t.detach()
                if t.dtype not in (torch.float32, torch.bfloat16):
                    t = t.to(torch.float32)
                t = t.contiguous()
                shape = (C.c_int64 * max(t.dim(), 1))(*(list(t.shape) or [1]))
                rc = self._lib.vlaser_set_weight(self._handle, key.encode(), C.c_void_p(t.data_ptr()), shape,
                                                 max(t.dim(), 1), _lib.BF16 if t.dtype == torch.bfloat16 else _lib.F32)
                _lib.check(rc, f"vlaser_set_weight({key})")
                if rc == 1:
                    unexpected.append(key)
                else:
                    self._loaded.add(canonical_key(key))
            torch.cuda.synchronize(self.device)
        missing = [k for k in state_dict_spec(self.vcfg) if k not in self._loaded]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing={missing[:5]} unexpected={unexpected[:5]}")
        return _IncompatibleKeys(missing, unexpected)

    def finalize(self):
        if not self._finalized:
            with torch.cuda.device(self.device):
                _lib.check(self._lib.vlaser_finalize(self._handle), "vlaser_finalize")
            self._finalized = True
        return self

    def freeze_all_weights(self):
        return self

    def tie_action_proprio_weights(self):
        return self

    def to(self, *args, **kwargs):
        """`.to(dtype)` / `.to(device)` of the reference (eval.py:43-45). The compute path is fixed: bf16 tensor-core
        GEMMs with fp32 accumulation, fp32 residual streams and norms, weights rounded to bf16 at load. Asking for
        float32 (the shipped evaluator configs set `use_bf16: False`) therefore does NOT give the reference's fp32
        arithmetic: it is accepted -- the I/O tensors may be float32 -- and a warning says so once. Actions stay within
        the 1e-2 bound of the fp32 reference (tests/test_gpu_parity.py, incl. un-rounded fp32 weights)."""
        import warnings
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.dtype):
                if a == torch.float32 and not getattr(self, "_warned_fp32", False):
                    warnings.warn("vlaser_b200 computes in bf16 (fp32 accumulation / residuals) even when float32 is "
                                  "requested; weights are rounded to bf16 at load", RuntimeWarning, stacklevel=2)
                    self._warned_fp32 = True
                elif a not in (torch.float32, torch.bfloat16):
                    raise TypeError(f"vlaser_b200: unsupported dtype {a} (float32 or bfloat16)")
                self.dtype = a
            elif isinstance(a, (torch.device, str, int)) and not isinstance(a, bool):
                dev = torch.device(a) if not isinstance(a, int) else torch.device("cuda", a)
                if dev.type != "cuda":
                    raise RuntimeError("vlaser_b200 has no CPU path")
                if dev.index is not None and dev.index != self.device.index:
                    raise RuntimeError(f"the handle lives on {self.device}; construct PiZeroInference(device={dev.index}) instead")
        return self

    def eval(self):
        return self

    def parameters(self) -> Iterable[torch.Tensor]:
        return iter(())

    def set_option(self, key: str, value: int):
        _lib.check(self._lib.vlaser_set_option(self._handle, key.encode(), int(value)), f"set_option({key})")
        if key not in ("stop_after", "profile_gemm", "profile_mtot"):
            self._options[key] = int(value)
        return self

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None and self._handle.value:
                self._lib.vlaser_destroy(self._handle)
                self._handle = C.c_void_p()
        except Exception:
            pass

    # ---- input preparation (kept for API compatibility; the kernels never read the dense mask) ----
    def build_causal_mask_and_position_ids(self, attention_mask: torch.Tensor, dtype: torch.dtype):
        """pizero_internvl.py:517-587 (block attention mask + per-block position ids starting at 1).
        Vectorised; same values as the reference's per-sample loop."""
        bsz, S = attention_mask.shape
        H = self.num_action_tokens
        L = S + H + 1
        cnt = attention_mask.sum(dim=1).view(bsz, 1, 1)
        dev = attention_mask.device
        r = torch.arange(L, device=dev).view(1, L, 1)
        c = torch.arange(L, device=dev).view(1, 1, L)
        allowed = ((r < cnt) & (c < cnt)) | ((r >= S) & (c < cnt)) | ((r == S) & (c == S)) | ((r > S) & (c >= S))
        mask = torch.full((bsz, L, L), torch.finfo(dtype).min, dtype=dtype, device=dev)
        mask.masked_fill_(allowed, 0)
        mask = mask.unsqueeze(1)
        vlm_position_ids = torch.arange(1, S + 1).repeat(bsz, 1)
        proprio_position_ids = torch.arange(1, 2).repeat(bsz, 1)
        action_position_ids = torch.arange(2, H + 2).repeat(bsz, 1)
        return mask, vlm_position_ids, proprio_position_ids, action_position_ids

    def split_full_mask_into_submasks(self, causal_mask: torch.Tensor):
        """pizero_internvl.py:589-603 (4-D branch)."""
        n = self.max_image_text_tokens + self.num_proprio_tokens
        return causal_mask[..., :n, :n], causal_mask[..., -self.num_action_tokens:, :]

    # ---- the hot path ------------------------------------------------------------------------------
    def forward(self, input_ids, pixel_values, image_text_proprio_mask=None, action_mask=None,
                vlm_position_ids=None, proprio_position_ids=None, action_position_ids=None, proprios=None,
                noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self.infer_action(input_ids, pixel_values, image_text_proprio_mask, action_mask, vlm_position_ids,
                                 proprio_position_ids, action_position_ids, proprios, noise=noise)

    __call__ = forward

    def infer_action(self, input_ids, pixel_values, image_text_proprio_mask=None, action_mask=None,
                     vlm_position_ids=None, proprio_position_ids=None, action_position_ids=None, proprios=None,
                     noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """pizero_internvl.py:798-936. Position ids are the ones `build_causal_mask_and_position_ids`
        produces (vlm 1..S, proprio 1, action 2..H+1); the id tensors are accepted for signature
        compatibility. `noise` (optional, [B,H,A]) replaces the internal `torch.randn` draw."""
        v = self.vcfg
        self.finalize()
        dev = self.device
        dtype = pixel_values.dtype
        if dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("pixel_values must be float32 or bfloat16")
        io = _lib.BF16 if dtype == torch.bfloat16 else _lib.F32
        B = pixel_values.size(0) // v.num_images
        S, H, A = v.max_image_text_tokens, v.num_action_tokens, v.action_dim
        input_ids = input_ids.to(dev, torch.int64).contiguous()
        pixel_values = pixel_values.to(dev).contiguous()
        proprios = proprios.to(dev, dtype).contiguous()
        assert input_ids.shape == (B, S), f"input_ids must be [B,{S}]"
        # The reference uses whatever position ids it is handed (pizero_internvl.py:854-855, 895); the kernels bake in
        # the ids build_causal_mask_and_position_ids produces. Anything else would silently change the RoPE phases.
        for name, ids_t, first, n in (("vlm_position_ids", vlm_position_ids, 1, S), ("proprio_position_ids", proprio_position_ids, 1, 1),
                                      ("action_position_ids", action_position_ids, 2, H)):
            if ids_t is None:
                continue
            want = torch.arange(first, first + n, device=ids_t.device, dtype=ids_t.dtype)
            if tuple(ids_t.shape[-1:]) != (n,) or not bool((ids_t.reshape(-1, n) == want).all()):
                raise NotImplementedError(f"{name} differs from the ids of build_causal_mask_and_position_ids "
                                          f"({first}..{first + n - 1}): custom position ids are not supported")
        valid_len = None
        if image_text_proprio_mask is not None:
            # number of keys the first prefix row may attend to == valid prefix length
            valid_len = (image_text_proprio_mask[:, 0, 0, :S] == 0).sum(-1).to(dev, torch.int32).contiguous()
        if noise is None:
            noise = torch.randn((B, H, A), device=dev, dtype=dtype)  # pizero_internvl.py:879-881
        noise = noise.to(dev, dtype).contiguous()
        assert noise.shape == (B, H, A)
        out = torch.empty((B, v.horizon_steps, A), device=dev, dtype=dtype)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = self._lib.vlaser_infer_action(
                self._handle, C.c_void_p(input_ids.data_ptr()), C.c_void_p(pixel_values.data_ptr()),
                C.c_void_p(valid_len.data_ptr()) if valid_len is not None else None,
                C.c_void_p(proprios.data_ptr()), C.c_void_p(noise.data_ptr()), C.c_void_p(out.data_ptr()),
                B, io, C.c_void_p(stream))
        _lib.check(rc, "vlaser_infer_action")
        return out

    def infer_action_naive(self, input_ids, pixel_values, causal_mask=None, vlm_position_ids=None,
                           proprio_position_ids=None, action_position_ids=None, proprios=None,
                           noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """pizero_internvl.py:938-1003: the same flow integration with the whole joint model re-run in every Euler step
        (cache_mode="no_append"). Under the block mask the prefix and proprio blocks never see the action block, so the
        result is the one `infer_action` returns; this entry runs the cached schedule. (With the InternVL backbone the
        reference's own method cannot run -- it does not hand `position_embeddings_all` to the joint model and fails at
        joint_model.py:549 -- so there is no golden vector for it; it exists here for signature compatibility.)"""
        itp = None
        if causal_mask is not None:
            itp, _ = self.split_full_mask_into_submasks(causal_mask)
        return self.infer_action(input_ids, pixel_values, itp, None, vlm_position_ids, proprio_position_ids,
                                 action_position_ids, proprios, noise=noise)

    def flow_matching_loss(self, input_ids, pixel_values, causal_mask=None, vlm_position_ids=None,
                           proprio_position_ids=None, action_position_ids=None, proprios=None, actions=None, t=None,
                           noise: Optional[torch.Tensor] = None, return_velocity: bool = False):
        """`PiZero.forward(..., is_vlm=False)` of the training step, forward only (pizero_internvl.py:1064-1197): the
        flow-matching loss mean((v_psi - (x1 - (1 - sig_min) x0))^2) for target chunks `actions` [B, H, A] at the
        per-sample times `t` [B]. `noise` (optional) replaces the `torch.randn_like` draw of :1094. Returns the loss as a
        0-d float32 tensor on the device (and v_psi [B, H, A] float32 with return_velocity). No gradients are produced."""
        v = self.vcfg
        self.finalize()
        dev = self.device
        dtype = pixel_values.dtype
        if dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("pixel_values must be float32 or bfloat16")
        io = _lib.BF16 if dtype == torch.bfloat16 else _lib.F32
        B = pixel_values.size(0) // v.num_images
        S, H, A = v.max_image_text_tokens, v.num_action_tokens, v.action_dim
        input_ids = input_ids.to(dev, torch.int64).contiguous()
        pixel_values = pixel_values.to(dev).contiguous()
        proprios = proprios.to(dev, dtype).contiguous()
        actions = actions.to(dev, dtype).contiguous()
        t = t.to(dev, dtype).contiguous()
        assert input_ids.shape == (B, S) and actions.shape == (B, H, A) and t.shape == (B,)
        valid_len = None
        if causal_mask is not None:
            valid_len = (causal_mask[:, 0, 0, :S] == 0).sum(-1).to(dev, torch.int32).contiguous()
        if noise is None:
            noise = torch.randn_like(actions)  # :1094
        noise = noise.to(dev, dtype).contiguous()
        loss = torch.empty((), device=dev, dtype=torch.float32)
        vel = torch.empty((B, H, A), device=dev, dtype=torch.float32) if return_velocity else None
        sig = float(_cfg_get(self.cfg, "flow_sig_min", 0.001))
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = self._lib.vlaser_flow_matching_loss(
                self._handle, C.c_void_p(input_ids.data_ptr()), C.c_void_p(pixel_values.data_ptr()),
                C.c_void_p(valid_len.data_ptr()) if valid_len is not None else None, C.c_void_p(proprios.data_ptr()),
                C.c_void_p(actions.data_ptr()), C.c_void_p(noise.data_ptr()), C.c_void_p(t.data_ptr()), sig,
                C.c_void_p(loss.data_ptr()), C.c_void_p(vel.data_ptr()) if vel is not None else None, B, io,
                C.c_void_p(stream))
        _lib.check(rc, "vlaser_flow_matching_loss")
        return (loss, vel) if return_velocity else loss

    def generate(self, input_ids, pixel_values=None, attention_mask=None, max_new_tokens: int = 32,
                 eos_token_id: Optional[int] = None, fill_token_id: Optional[int] = None, return_first_logits: bool = False):
        """Greedy text generation of the chat VLM -- `InternVLChatModel.generate(pixel_values, input_ids, attention_mask,
        do_sample=False)` (Vlaser_VLM/internvl_chat/internvl/model/internvl_chat/modeling_internvl_chat.py:401-439), also
        the decode loop around `PiZero.infer_text` (pizero_internvl.py:1412-1447). `input_ids` [B, max_image_text_tokens]
        right padded with pad_token_id, `pixel_values` [tiles, 3, px, px] the image tiles of the batch in order (None: a
        text-only prompt). Returns tokens [B, max_new_tokens] int64 (fill_token_id after a sample's eos; default: the pad
        token) and, with return_first_logits, the float32 logits [B, vocab] of the last prompt position."""
        v = self.vcfg
        self.finalize()
        dev = self.device
        dtype = pixel_values.dtype if pixel_values is not None else torch.bfloat16
        if dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("pixel_values must be float32 or bfloat16")
        io = _lib.BF16 if dtype == torch.bfloat16 else _lib.F32
        input_ids = input_ids.to(dev, torch.int64).contiguous()
        B, S = input_ids.shape
        assert S == v.max_image_text_tokens, f"input_ids must be [B,{v.max_image_text_tokens}]"
        if pixel_values is not None:
            pixel_values = pixel_values.to(dev).contiguous()
            assert pixel_values.size(0) == B * v.num_images, "one tile per (sample, image slot): B * cond_steps tiles"
        valid_len = None
        if attention_mask is not None:
            valid_len = attention_mask.sum(-1).to(dev, torch.int32).contiguous()
        eos = -1 if eos_token_id is None else int(eos_token_id)
        fill = v.pad_token_id if fill_token_id is None else int(fill_token_id)
        out = torch.empty((B, int(max_new_tokens)), device=dev, dtype=torch.int64)
        logits = torch.empty((B, v.embed_rows), device=dev, dtype=torch.float32) if return_first_logits else None
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = self._lib.vlaser_generate(
                self._handle, C.c_void_p(input_ids.data_ptr()),
                C.c_void_p(pixel_values.data_ptr()) if pixel_values is not None else None,
                C.c_void_p(valid_len.data_ptr()) if valid_len is not None else None, int(max_new_tokens), eos, fill,
                C.c_void_p(out.data_ptr()), C.c_void_p(logits.data_ptr()) if logits is not None else None, B, io,
                C.c_void_p(stream))
        _lib.check(rc, "vlaser_generate")
        return (out, logits) if return_first_logits else out

    def infer_action_host(self, input_ids, pixel_values, proprios, noise, valid_len=None) -> torch.Tensor:
        """Same call on HOST tensors (CPU torch tensors): H2D staging, compute, D2H, synchronise."""
        v = self.vcfg
        self.finalize()
        dtype = pixel_values.dtype
        io = _lib.BF16 if dtype == torch.bfloat16 else _lib.F32
        B = pixel_values.size(0) // v.num_images
        input_ids = input_ids.to(torch.int64).contiguous()
        pixel_values = pixel_values.contiguous()
        proprios = proprios.to(dtype).contiguous()
        noise = noise.to(dtype).contiguous()
        out = torch.empty((B, v.horizon_steps, v.action_dim), dtype=dtype)
        vl = valid_len.to(torch.int32).contiguous() if valid_len is not None else None
        with torch.cuda.device(self.device):
            rc = self._lib.vlaser_infer_action_host(
                self._handle, C.c_void_p(input_ids.data_ptr()), C.c_void_p(pixel_values.data_ptr()),
                C.c_void_p(vl.data_ptr()) if vl is not None else None, C.c_void_p(proprios.data_ptr()),
                C.c_void_p(noise.data_ptr()), C.c_void_p(out.data_ptr()), B, io)
        _lib.check(rc, "vlaser_infer_action_host")
        return out

    # ---- introspection (tests, bench) --------------------------------------------------------------
    def debug_buffer(self, name: str, dtype: torch.dtype) -> torch.Tensor:
        n = self._lib.vlaser_debug_copy(self._handle, name.encode(), None, 0)
        _lib.check(int(n), f"debug_copy({name})")
        buf = torch.empty(int(n), dtype=torch.uint8)
        _lib.check(int(self._lib.vlaser_debug_copy(self._handle, name.encode(), C.c_void_p(buf.data_ptr()), int(n))),
                   f"debug_copy({name})")
        return buf.view(dtype)

    def profile_read(self):
        """(gemm_ms, gemm_flops, gemm_launches) accumulated since "profile_gemm" was switched on."""
        ms, fl, n = C.c_double(), C.c_double(), C.c_int64()
        _lib.check(self._lib.vlaser_profile_read(self._handle, C.byref(ms), C.byref(fl), C.byref(n)), "profile_read")
        return ms.value, fl.value, n.value

    def profile_read_classes(self):
        """{"single": {...}, "pair": {...}, "expert": {...}} with ms / flops / bytes / launches since "profile_gemm" was
        switched on ("expert" = the persistent expert-stream kernel, one launch per pass)."""
        buf = (C.c_double * 12)()
        _lib.check(self._lib.vlaser_profile_read_classes(self._handle, buf), "profile_read_classes")
        keys = ("ms", "flops", "bytes", "launches")
        return {"single": dict(zip(keys, buf[0:4])), "pair": dict(zip(keys, buf[4:8])), "expert": dict(zip(keys, buf[8:12]))}

    def last_launch_count(self) -> int:
        return int(self._lib.vlaser_last_launch_count(self._handle))

    def flops_per_call(self, valid_len) -> float:
        vl = torch.as_tensor(valid_len, dtype=torch.int32).contiguous()
        return float(self._lib.vlaser_flops_per_call(self._handle, C.c_void_p(vl.data_ptr()), vl.numel()))
