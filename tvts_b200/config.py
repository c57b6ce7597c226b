"""Architecture + workload descriptions for the TVTS / TVTSv2 pre-training hot path.

Dimensions restate the constructor calls of the reference:
  B/16: v2/model/model_dist_TVTSv2_ViT_B_16.py:29-31,49   B/32: ..._B_32.py:29-31,49
  H/14: v2/model/model_dist_TVTSv2_ViT_H_14.py:43-45,85 + v2/OpenCLIP/model_configs/ViT-H-14.json
  text (B models): CLIP(512,224,12,768,p,77,49408,512,8,12)  v2/CLIP/clip/model.py:243-300
"""
from dataclasses import dataclass, replace


@dataclass(frozen=True)
class ArchConfig:
    name: str
    # video tower
    patch: int
    width: int
    layers: int
    heads: int
    embed_dim: int
    num_frames: int = 12          # size of temporal_embedding; caps T
    mask_ratio: float = 0.0
    resolution: int = 224
    act: str = "quick_gelu"       # 'quick_gelu' (B models) | 'gelu' (erf; H/14)
    ln_eps: float = 1e-5
    post_mode: str = "all"        # 'all': ln_post+proj on every token (B) | 'h14': ln_post on CLS only
    # text tower
    text_width: int = 512
    text_heads: int = 8
    text_layers: int = 12
    context: int = 77
    vocab: int = 49408
    text_act: str = "quick_gelu"
    # sort head
    sort_heads: int = 8
    sort_depth: int = 2
    n_trans: int = 4
    sort_ln_eps: float = 1e-6
    temperature: float = 0.05
    # input normalisation applied by the data pipeline (v2/video_transforms/videoaug.py:16); used only when uint8 clips are fed
    input_mean: tuple = (0.485, 0.456, 0.406)
    input_std: tuple = (0.229, 0.224, 0.225)

    @property
    def patches_per_frame(self):
        return (self.resolution // self.patch) ** 2

    @property
    def kept_per_frame(self):
        # v2/model/video_encoder_ViT_B_16.py:220  time_n = int(P * (1 - mask_ratio))
        return int(self.patches_per_frame * (1 - self.mask_ratio))

    def tokens(self, T):
        return 1 + T * self.kept_per_frame

    def small(self, **kw):
        return replace(self, **kw)


TVTSV2_B_32 = ArchConfig("TVTSv2_B_32", patch=32, width=768, layers=12, heads=12, embed_dim=512, mask_ratio=0.0)
TVTSV2_B_16 = ArchConfig("TVTSv2_B_16", patch=16, width=768, layers=12, heads=12, embed_dim=512, mask_ratio=0.5)
TVTSV2_H_14 = ArchConfig("TVTSv2_H_14", patch=14, width=1280, layers=32, heads=16, embed_dim=1024, mask_ratio=0.7,
                         act="gelu", post_mode="h14", text_width=1024, text_heads=16, text_layers=24,
                         text_act="gelu", sort_heads=16)

ARCHS = {c.name: c for c in (TVTSV2_B_32, TVTSV2_B_16, TVTSV2_H_14)}

# tiny variants used by parity tests (same code paths, seconds on CPU)
# (all head dims 64 like production; heads >= 2 everywhere: with 1 head the reference's in-place `q *= scale`
#  at sort_transformer.py:51 hits a view and raises under torch 2.x autograd)
TINY_B = ArchConfig("tiny_B", patch=32, width=128, layers=2, heads=2, embed_dim=128, mask_ratio=0.0,
                    text_width=128, text_heads=2, text_layers=2, vocab=512, sort_heads=2)
TINY_B_MASK = TINY_B.small(name="tiny_B_mask", patch=16, mask_ratio=0.5)
# H/14-shaped toys: patch 14, mask 0.7 (int(256*0.3) = 76 kept patches), head dim 80, exact GELU in both towers, ln_post on CLS only,
# sort head on the patch tokens.  TINY_H (width 160) runs on the CPU restatements only (the CUDA LayerNorm needs a width that is a
# multiple of 128); TINY_H640 (8 heads x 80) is the smallest H/14-shaped model the kernels take.
TINY_H = ArchConfig("tiny_H", patch=14, width=160, layers=2, heads=2, embed_dim=128, mask_ratio=0.7, act="gelu", post_mode="h14",
                    text_width=128, text_heads=2, text_layers=2, vocab=512, text_act="gelu", sort_heads=2)
TINY_H640 = TINY_H.small(name="tiny_H640", width=640, heads=8)


@dataclass(frozen=True)
class Workload:
    """One of BASELINE.json's configs, as runtime dims."""
    name: str
    arch: ArchConfig
    batch: int          # per GPU
    frames: int
    n_trans: int = 4


WORKLOADS = {
    # BASELINE.json configs[0..2]; c3 is quoted per GPU (global 256 over 8 GPUs)
    "c1": Workload("c1", TVTSV2_B_32, batch=4, frames=2),
    "c2": Workload("c2", TVTSV2_B_32, batch=64, frames=8),
    "c3": Workload("c3", TVTSV2_B_16, batch=32, frames=8),
    # BASELINE.json configs[3]: H/14, 16 frames (the temporal table is built with num_frames=16); per-GPU batch 8
    "c4": Workload("c4", TVTSV2_H_14.small(num_frames=16), batch=8, frames=16),
    # not a BASELINE config: the toy model, for dry runs of the bench script (tests/bench_dryrun.py)
    "tiny": Workload("tiny", TINY_B_MASK, batch=2, frames=3),
}
