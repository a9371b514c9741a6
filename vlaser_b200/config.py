"""Shape/config description of the Vlaser-VLA `infer_action` hot path.

Field names follow the reference's hydra config keys where one exists
(`Vlaser_VLA/Simpler/config/eval/bridge_internvl_448.yaml`, read by
`PiZero.__init__`, `Vlaser_VLA/Simpler/src/model/vla/pizero_internvl.py:156-337`) and the
InternVL3 / InternViT / Qwen2.5 config fields otherwise (SURVEY.md §8 dims table).
"""
from __future__ import annotations

from dataclasses import dataclass, asdict, replace


@dataclass(frozen=True)
class VlaserConfig:
    # --- InternViT (modeling_intern_vit.py:133-295) ---
    image_size: int = 224
    patch_size: int = 14
    vit_hidden: int = 1024
    vit_layers: int = 24
    vit_heads: int = 16            # head dim is vit_hidden // vit_heads and must be 64
    vit_mlp: int = 4096
    vit_eps: float = 1e-6
    # --- projector mlp1 (modeling_internvl_chat.py:89-94); nn.LayerNorm default eps ---
    proj_eps: float = 1e-5
    # --- LLM / "vlm" mixture (Qwen2.5) ---
    hidden: int = 1536
    layers: int = 28
    n_heads: int = 12
    n_kv_heads: int = 2
    head_dim: int = 128            # hard-coded 128 in joint_model.py:437
    mlp: int = 8960
    rms_eps: float = 1e-6
    rope_theta: float = 1e6
    # --- action expert (proprio + action mixtures share it; pizero_internvl.py:116-135) ---
    hidden_e: int = 768
    mlp_e: int = 8960
    # --- tokens ---
    embed_rows: int = 151930       # rows of embed_tokens after resize_token_embeddings
    image_token_index: int = 151667
    pad_token_id: int = 151643
    max_image_text_tokens: int = 384
    # --- action parameterisation ---
    cond_steps: int = 1            # == num_images (pizero_internvl.py:336)
    horizon_steps: int = 4         # cfg.horizon_steps (returned chunk length)
    action_dim: int = 7
    proprio_dim: int = 7
    num_inference_steps: int = 10
    final_action_clip_value: float | None = 1.0
    time_max_period: float = 10000.0

    # ---- derived ----
    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size

    @property
    def patches(self) -> int:
        return self.grid * self.grid

    @property
    def vit_tokens(self) -> int:
        return self.patches + 1

    @property
    def image_tokens(self) -> int:
        """T: image tokens per image after pixel_shuffle(0.5)."""
        return self.patches // 4

    @property
    def num_images(self) -> int:
        return self.cond_steps

    @property
    def num_action_tokens(self) -> int:
        """H = horizon_steps + cond_steps - 1 (pizero_internvl.py:210,223)."""
        return self.horizon_steps + self.cond_steps - 1

    @property
    def q_dim(self) -> int:
        return self.n_heads * self.head_dim

    @property
    def kv_dim(self) -> int:
        return self.n_kv_heads * self.head_dim

    def to_dict(self) -> dict:
        return asdict(self)

    def with_(self, **kw) -> "VlaserConfig":
        return replace(self, **kw)


def vlaser_2b_bridge() -> VlaserConfig:
    """BASELINE configs 1/2: Vlaser-2B-VLA, WidowX/Bridge, 1x224 px, S=384, H=4, A=7."""
    return VlaserConfig()


def vlaser_2b_fractal() -> VlaserConfig:
    """BASELINE config 3: Google Robot / fractal (Dp=8, time_max_period=100)."""
    return VlaserConfig(proprio_dim=8, time_max_period=100.0)


def vlaser_2b_robotwin() -> VlaserConfig:
    """BASELINE config 4: RoboTwin shapes on the same path: 3 cameras, 14-DoF, S=896."""
    return VlaserConfig(cond_steps=3, action_dim=14, proprio_dim=14, max_image_text_tokens=896)


def vlaser_8b_vla() -> VlaserConfig:
    """BASELINE config 5 dims: Vlaser-8B (InternViT-300M + Qwen2.5-7B: hidden 3584, 28/4 heads, MLP 18944) as the 8B VLA
    (`model_size="8B"`, pizero_internvl.py:128-131: expert MLP 5600), one 448 px tile per observation (256 image
    tokens). The prefix encode + bidirectional block prefill of config 5 is this preset with `stop_after=3`."""
    return VlaserConfig(image_size=448, hidden=3584, n_heads=28, n_kv_heads=4, mlp=18944, mlp_e=5600, embed_rows=151674)


def tiny() -> VlaserConfig:
    """Small config with the same structure (GQA 6:1, head_dim 128, ViT head 64, 16x16 patch grid
    -- the reference hard-codes the grid, pizero_internvl.py:739-743) for golden fixtures and fast
    parity tests."""
    return VlaserConfig(
        image_size=224, vit_hidden=128, vit_layers=2, vit_heads=2, vit_mlp=256,
        hidden=256, layers=3, n_heads=6, n_kv_heads=1, mlp=512,
        hidden_e=128, mlp_e=384,
        embed_rows=512, image_token_index=500, pad_token_id=499,
        max_image_text_tokens=96,
    )


def tiny_robotwin() -> VlaserConfig:
    """`tiny` with the RoboTwin shapes (3 cameras, 14-DoF chunks) for plugin tests."""
    return tiny().with_(cond_steps=3, action_dim=14, proprio_dim=14, max_image_text_tokens=224)


PRESETS = {
    "vlaser_2b_bridge": vlaser_2b_bridge,
    "vlaser_2b_fractal": vlaser_2b_fractal,
    "vlaser_2b_robotwin": vlaser_2b_robotwin,
    "vlaser_8b_vla": vlaser_8b_vla,
    "tiny": tiny,
    "tiny_robotwin": tiny_robotwin,
}
