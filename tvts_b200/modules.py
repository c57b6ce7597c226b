"""nn.Module mirror of the reference's model surface for the hot path, executing on the B200 kernels.

The modules below are PARAMETER CONTAINERS with the reference's exact parameter names / shapes / initialisers
(so `named_parameters()` substring policies, `state_dict()` round trips and DDP all behave as in the reference);
all arithmetic is delegated to tvts_b200.engine (one autograd node per tower, hand-written CUDA kernels).

Reference classes mirrored (paths relative to /root/reference):
  VisionTransformer / SpaceTimeTransformer / ResidualSpaceTimeAttentionBlock / VarAttention / LayerNorm / QuickGELU
      v2/model/video_encoder_ViT_B_16.py:18-235 (identical B_32 copy)
  CLIP text Transformer / ResidualAttentionBlock      v2/CLIP/clip/model.py:157-203, init :301-328
  SortTransformer / AttnBlock / SelfAttention / Mlp   v2/model/sort_transformer.py:16-142
  TVTSv2_B_16 / TVTSv2_B_32                           v2/model/model_dist_TVTSv2_ViT_B_16.py:10-116
  NormSoftmaxLoss                                     v2/model/loss.py:5-25
  sim_matrix                                          v2/model/model_dist_TVTSv2_ViT_B_16.py:119-127
"""
import math
import os
from collections import OrderedDict

import numpy as np
import torch
from torch import nn

from . import config as C
from . import engine as E


class BaseModel(nn.Module):
    """v2/base/base_model.py:6-26"""

    def forward(self, *inputs):
        raise NotImplementedError

    def __str__(self):
        params = sum(int(np.prod(p.size())) for p in self.parameters() if p.requires_grad)
        return super().__str__() + "\nTrainable parameters: {}".format(params)


class LayerNorm(nn.LayerNorm):
    """Parameter holder; the fp32 LayerNorm itself runs in tvts_layernorm_fwd/bwd."""


class QuickGELU(nn.Module):
    pass


def _no_direct_call(self, *a, **k):
    raise RuntimeError(f"{type(self).__name__} is a parameter container of the B200 engine; call the owning tower's forward()")


# ------------------------------------------------------------------------------------------------ video tower
class VarAttention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, initialize="random"):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        if initialize == "zeros":      # video_encoder_ViT_B_16.py:28-34: the temporal branch starts as an exact no-op
            self.qkv.weight.data.fill_(0)
            self.qkv.bias.data.fill_(0)
            self.proj.weight.data.fill_(1)
            self.proj.bias.data.fill_(0)

    forward = _no_direct_call


class ResidualSpaceTimeAttentionBlock(nn.Module):
    def __init__(self, d_model, n_head, time_init="zeros"):
        super().__init__()
        self.attn = VarAttention(d_model, num_heads=n_head, qkv_bias=True)
        self.timeattn = VarAttention(d_model, num_heads=n_head, qkv_bias=True, initialize=time_init)
        self.ln_3 = LayerNorm(d_model)          # registration order = the reference's (:99-108): optimizer state in its
        self.ln_1 = LayerNorm(d_model)          # checkpoints is keyed by position in named_parameters() order
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, d_model * 4)),
            ("gelu", QuickGELU()),
            ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = LayerNorm(d_model)

    forward = _no_direct_call


class SpaceTimeTransformer(nn.Module):
    def __init__(self, width, layers, heads):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.Sequential(*[ResidualSpaceTimeAttentionBlock(width, heads) for _ in range(layers)])

    forward = _no_direct_call


class VisionTransformer(nn.Module):
    """Video ViT with divided space-time attention and tube masking; forward(x [B,T,3,R,R], keep_ind [B,n]) -> [B,N,E]."""

    def __init__(self, input_resolution, patch_size, width, layers, heads, output_dim, num_frames=12, mask_ratio=0.):
        super().__init__()
        self.input_resolution = input_resolution
        self.output_dim = output_dim
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = SpaceTimeTransformer(width, layers, heads)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))
        self.patches_per_frame = (input_resolution // patch_size) ** 2
        self.temporal_embedding = nn.Parameter(scale * torch.randn(num_frames, width))
        self.mask_ratio = mask_ratio
        self.cfg = C.ArchConfig("video", patch=patch_size, width=width, layers=layers, heads=heads, embed_dim=output_dim,
                                num_frames=num_frames, mask_ratio=mask_ratio, resolution=input_resolution)
        self._ordered = None

    def _named(self):
        if self._ordered is None:
            have = dict(self.named_parameters())
            self._ordered = OrderedDict((k, have[k]) for k in E.video_param_names(self.cfg))
        return self._ordered

    def forward(self, x, keep_ind):
        return E.video_tower(self.cfg, self._named(), x, keep_ind)


class ResidualSpaceTimeAttentionBlockH14(nn.Module):
    """v2/model/video_encoder_ViT_H_14.py:210-254 (registration order ln_1, attn, timeattn, ln_3, ls_3, ls_1, ln_2, mlp, ls_2; the
    LayerScales are nn.Identity because ls_init_value is None in ViT-H-14.json)."""

    def __init__(self, d_model, n_head, mlp_ratio=4.0):
        super().__init__()
        self.ln_1 = LayerNorm(d_model)
        self.attn = VarAttention(d_model, num_heads=n_head, qkv_bias=True)
        self.timeattn = VarAttention(d_model, num_heads=n_head, qkv_bias=True, initialize="zeros")
        self.ln_3 = LayerNorm(d_model)
        self.ls_3 = nn.Identity()
        self.ls_1 = nn.Identity()
        self.ln_2 = LayerNorm(d_model)
        mlp_width = int(d_model * mlp_ratio)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, mlp_width)),
            ("gelu", nn.GELU()),
            ("c_proj", nn.Linear(mlp_width, d_model))]))
        self.ls_2 = nn.Identity()

    forward = _no_direct_call


class SpaceTimeTransformerH14(nn.Module):
    def __init__(self, width, layers, heads, mlp_ratio=4.0):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.ModuleList([ResidualSpaceTimeAttentionBlockH14(width, heads, mlp_ratio) for _ in range(layers)])

    forward = _no_direct_call


class VisionTransformerH14(nn.Module):
    """The modified OpenCLIP video ViT (v2/model/video_encoder_ViT_H_14.py:303-484): exact GELU, ln_post on the CLS token only;
    forward(x [B,T,3,R,R], keep_ind [B,n]) -> (pooled [B,E], tokens [B,N-1,E]).  Options the H/14 model never enables
    (LayerScale, patch dropout, patch-norm, attentional / average pooling) are rejected."""

    def __init__(self, image_size, patch_size, width, layers, heads, mlp_ratio=4.0, ls_init_value=None, global_average_pool=False,
                 attentional_pool=False, n_queries=256, attn_pooler_heads=8, output_dim=512, patch_dropout=0., input_patchnorm=False,
                 act_layer=nn.GELU, norm_layer=None, output_tokens=False, num_frames=12, mask_ratio=0.):
        super().__init__()
        if ls_init_value is not None or global_average_pool or attentional_pool or input_patchnorm or patch_dropout:
            raise NotImplementedError("VisionTransformerH14: only the configuration of model_dist_TVTSv2_ViT_H_14.py is built")
        if act_layer is not nn.GELU:
            raise NotImplementedError("VisionTransformerH14: act_layer must be nn.GELU")
        self.output_tokens = output_tokens
        self.image_size, self.patch_size = (image_size, image_size), (patch_size, patch_size)
        self.grid_size = (image_size // patch_size, image_size // patch_size)
        self.output_dim = output_dim
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn(self.grid_size[0] * self.grid_size[1] + 1, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = SpaceTimeTransformerH14(width, layers, heads, mlp_ratio)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))
        self.patches_per_frame = self.grid_size[0] * self.grid_size[1]
        self.temporal_embedding = nn.Parameter(scale * torch.randn(num_frames, width))
        self.mask_ratio = mask_ratio
        self.cfg = C.ArchConfig("video_h14", patch=patch_size, width=width, layers=layers, heads=heads, embed_dim=output_dim,
                                num_frames=num_frames, mask_ratio=mask_ratio, resolution=image_size, act="gelu", post_mode="h14")
        self._ordered = None

    def _named(self):
        if self._ordered is None:
            have = dict(self.named_parameters())
            self._ordered = OrderedDict((k, have[k]) for k in E.video_param_names(self.cfg))
        return self._ordered

    def forward(self, x, keep_ind):
        out = E.video_tower(self.cfg, self._named(), x, keep_ind)
        return out[:, 0], out[:, 1:]


# ------------------------------------------------------------------------------------------------ CLIP text tower
class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model, n_head, attn_mask=None, open_clip=False):
        super().__init__()
        if open_clip:                       # v2/OpenCLIP/transformer.py:189-214: ln_1, attn, ls_1, ln_2, mlp (nn.GELU), ls_2
            self.ln_1 = LayerNorm(d_model)
            self.attn = nn.MultiheadAttention(d_model, n_head)
            self.ls_1 = nn.Identity()
            self.ln_2 = LayerNorm(d_model)
        else:                               # v2/CLIP/clip/model.py:171-183: attn, ln_1, mlp (QuickGELU), ln_2
            self.attn = nn.MultiheadAttention(d_model, n_head)
            self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, d_model * 4)),
            ("gelu", nn.GELU() if open_clip else QuickGELU()),
            ("c_proj", nn.Linear(d_model * 4, d_model))]))
        if open_clip:
            self.ls_2 = nn.Identity()
        else:
            self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask

    forward = _no_direct_call


class Transformer(nn.Module):
    def __init__(self, width, layers, heads, attn_mask=None, open_clip=False):
        super().__init__()
        self.width, self.layers = width, layers
        blocks = [ResidualAttentionBlock(width, heads, attn_mask, open_clip) for _ in range(layers)]
        self.resblocks = nn.ModuleList(blocks) if open_clip else nn.Sequential(*blocks)

    forward = _no_direct_call


class CLIPTextParts(nn.Module):
    """The pieces of a CLIP model the TVTSv2 constructors pull out (transformer, token/positional embedding, ln_final,
    text_projection), initialised like CLIP.initialize_parameters (v2/CLIP/clip/model.py:301-328)."""

    def __init__(self, embed_dim=512, context_length=77, vocab_size=49408, width=512, heads=8, layers=12, open_clip=False):
        super().__init__()
        self.context_length = context_length
        self.transformer = Transformer(width, layers, heads, open_clip=open_clip)
        self.token_embedding = nn.Embedding(vocab_size, width)
        self.positional_embedding = nn.Parameter(torch.empty(context_length, width))
        self.ln_final = LayerNorm(width)
        self.text_projection = nn.Parameter(torch.empty(width, embed_dim))
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        proj_std = (width ** -0.5) * ((2 * layers) ** -0.5)
        attn_std = width ** -0.5
        fc_std = (2 * width) ** -0.5
        for block in self.transformer.resblocks:
            nn.init.normal_(block.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(block.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(block.mlp.c_fc.weight, std=fc_std)
            nn.init.normal_(block.mlp.c_proj.weight, std=proj_std)
        nn.init.normal_(self.text_projection, std=width ** -0.5)


# ------------------------------------------------------------------------------------------------ sort head
class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden_features, out_features)

    forward = _no_direct_call


class SelfAttention(nn.Module):
    def __init__(self, dim, num_heads=12, qkv_bias=True):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)

    forward = _no_direct_call


class AttnBlock(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=True):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = SelfAttention(dim, num_heads=num_heads, qkv_bias=qkv_bias)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio))

    forward = _no_direct_call


class SortTransformer(nn.Module):
    """forward(text [B,n_trans,E], x [B,N,E]) -> logits [B,n_trans,n_trans]   (sort_transformer.py:84-142).
    `_init_weights` exists in the reference but is never applied, so Linear/LayerNorm keep torch's default init."""

    def __init__(self, num_classes, embed_dim=768, depth=2, num_heads=12, mlp_ratio=4., qkv_bias=True):
        super().__init__()
        self.num_classes = num_classes
        self.embed_dim = embed_dim
        self.type_embed = nn.Parameter(torch.zeros(1, 2, embed_dim))
        self.blocks = nn.ModuleList([AttnBlock(embed_dim, num_heads, mlp_ratio, qkv_bias) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.head = nn.Linear(embed_dim, num_classes)
        self.cfg = C.ArchConfig("sort", patch=16, width=embed_dim, layers=0, heads=num_heads, embed_dim=embed_dim,
                                sort_heads=num_heads, sort_depth=depth, n_trans=num_classes)
        self._ordered = None

    def _named(self):
        if self._ordered is None:
            self._ordered = OrderedDict(self.named_parameters())
        return self._ordered

    def forward(self, text, x):
        """`text` is [B, n_trans, E]; the engine consumes the clip-major [n_trans*B, E] layout it came from."""
        B, nt, Edim = text.shape
        t = text.detach().permute(1, 0, 2).reshape(nt * B, Edim)
        return E.sort_head(self.cfg, self._named(), t, x)

    def forward_clip_major(self, t_clip_major, x):
        return E.sort_head(self.cfg, self._named(), t_clip_major.detach(), x)


# ------------------------------------------------------------------------------------------------ losses
def sim_matrix(a, b, eps=1e-8):
    return E.sim_matrix(a, b, eps)


class NormSoftmaxLoss(nn.Module):
    def __init__(self, temperature=0.05):
        super().__init__()
        self.temperature = temperature

    def forward(self, x):
        return E.norm_softmax_loss(x, self.temperature)


# ------------------------------------------------------------------------------------------------ top model
class TVTSv2Base(BaseModel):
    """Shared body of TVTSv2_B_16 / TVTSv2_B_32 (model_dist_TVTSv2_ViT_B_16.py:10-116)."""

    PATCH = 16
    MASK_RATIO = 0.5
    CLIP_FILE = "CLIP/models/ViT-B-16.pt"
    SORT_HEAD = True
    OPEN_CLIP = False

    def __init__(self, args, load_checkpoint=None, arch=None):
        super().__init__()
        self.args = args
        self.num_clips = 4
        from .clip_compat import load as clip_load
        # an explicit `arch=` (tests, bench: synthetic weights) or a checkpoint that overwrites every weight anyway may run without the
        # released CLIP file; the plain reference-style construction raises like clip.load does (v2/CLIP/clip/clip.py:120-127)
        # (the reference's ConfigParser.initialize fills every constructor parameter whose NAME is a top-level config key -- `arch` is
        # one: it would hand the {"type", "args"} dict of the config file over here; anything that is not an ArchConfig is ignored)
        if not isinstance(arch, C.ArchConfig):
            arch = None
        allow_missing = arch is not None or load_checkpoint not in ["", None] or os.environ.get("TVTS_ALLOW_RANDOM_CLIP") == "1"
        arch = arch or self._default_arch()
        self.arch = arch
        clip_model, clip_visual_sd = clip_load(self.CLIP_FILE, arch, open_clip=self.OPEN_CLIP, allow_missing=allow_missing)
        self.text_model = clip_model.transformer
        self.text_token_embedding = clip_model.token_embedding
        self.text_positional_embedding = clip_model.positional_embedding
        self.text_ln_final = clip_model.ln_final
        self.text_projection = clip_model.text_projection
        self.video_model = self._build_video_model(arch)
        if load_checkpoint in ["", None] and clip_visual_sd is not None:
            new_sd = {}
            for k, v in clip_visual_sd.items():                     # :36-44 CLIP -> space-time key remap
                k = k.replace("in_proj_", "qkv.").replace("out_proj", "proj")
                new_sd[k] = v
            self.video_model.load_state_dict(new_sd, strict=False)
            print("ViT initialized with CLIP weights.")
        if self.SORT_HEAD:
            self.n_trans = arch.n_trans
            self.pred_model = SortTransformer(num_classes=self.n_trans, embed_dim=arch.embed_dim, num_heads=arch.sort_heads)
        if load_checkpoint not in ["", None]:
            # reference checkpoints pickle their ConfigParser next to the weights (base_trainer.py:173-181): weights_only must be off
            checkpoint = torch.load(load_checkpoint, map_location=self._checkpoint_location(), weights_only=False)
            state_dict = checkpoint["state_dict"]
            from .compat import state_dict_data_parallel_fix
            self.load_state_dict(state_dict_data_parallel_fix(state_dict, self.state_dict()), strict=True)
            print("loading checkpoint from {}".format(load_checkpoint))
        self._text_named = None

    def _default_arch(self):
        return C.TVTSV2_B_16 if self.PATCH == 16 else C.TVTSV2_B_32

    def _build_video_model(self, arch):
        return VisionTransformer(input_resolution=arch.resolution, patch_size=arch.patch, width=arch.width, layers=arch.layers,
                                 heads=arch.heads, output_dim=arch.embed_dim, num_frames=arch.num_frames, mask_ratio=arch.mask_ratio)

    def _checkpoint_location(self):
        return "cuda:{}".format(self.args.local_rank)

    def set_device(self, device):
        self.device = device

    def _text_params(self):
        if self._text_named is None:
            d = OrderedDict()
            for k, p in self.named_parameters():
                if k.startswith("text_"):
                    d[k] = p
            self._text_named = d
        return self._text_named

    def compute_text_all(self, text):
        """[n_txt, ctx] int tokens -> [n_txt, E]   (:97-111)"""
        return E.text_tower(self.arch, self._text_params(), text)

    def compute_text(self, text):
        """-> (text_before_embeddings, text_embeddings): the same [n_txt, E] tensor twice, like the reference (:108-111)."""
        t = self.compute_text_all(text)
        return t, t

    def compute_video(self, video, keep_ind):
        """-> (video_before_embeddings [B, N, E], video_embeddings [B, E] = the CLS row)   (:113-116)"""
        out = self.video_model(video, keep_ind)
        return out, out[:, 0, :]

    def _tower_stream(self, ref):
        """The text tower and the video tower are independent until the losses (model_dist_TVTSv2_ViT_B_16.py:61-76), so on a GPU the text
        tower runs on a second stream: its small GEMMs (M = n_trans x B x ctx rows: one or two tile waves) fill the SMs that the video
        tower's persistent GEMMs leave idle in their last partial wave, instead of queueing behind them.  Autograd replays each node's
        backward on the stream its forward used, so the two backward passes overlap in the same way; inside a CUDA-graph capture the
        fork / join become parallel branches of the graph.  TVTS_TEXT_STREAM=0 keeps everything on one stream."""
        if not ref.is_cuda or os.environ.get("TVTS_TEXT_STREAM", "1") == "0":
            return None
        s = getattr(self, "_side_stream", None)
        if s is None or s.device != ref.device:
            s = torch.cuda.Stream(device=ref.device)
            self._side_stream = s
        return s

    def forward(self, data, return_embeds=True):
        text, video, keep_ind = data["text"], data["video"], data["keep_ind"]
        B = video.shape[0]
        side = self._tower_stream(video)
        if side is None:
            t, _ = self.compute_text(text)                            # [n_trans*B, E] clip-major
            video_order_embeddings, video_embeddings = self.compute_video(video, keep_ind)
        else:
            cur = torch.cuda.current_stream(video.device)
            side.wait_stream(cur)                                     # tokens / weights are ready on the caller's stream
            with torch.cuda.stream(side):
                t, _ = self.compute_text(text)
            video_order_embeddings, video_embeddings = self.compute_video(video, keep_ind)
            cur.wait_stream(side)
            t.record_stream(cur)                                      # consumed below on the caller's stream
        n_trans = t.shape[0] // B
        text_embeddings = E.group_mean(t, n_trans)                    # :74-76
        if n_trans != 1:
            predict_order = self.pred_model.forward_clip_major(t, video_order_embeddings)   # :69-70 text is detached
        else:
            predict_order = None
        if return_embeds:
            return text_embeddings, video_embeddings, predict_order
        return sim_matrix(text_embeddings, video_embeddings)


class TVTSv2_B_16(TVTSv2Base):
    PATCH, MASK_RATIO, CLIP_FILE = 16, 0.5, "CLIP/models/ViT-B-16.pt"


class TVTSv2_B_32(TVTSv2Base):
    PATCH, MASK_RATIO, CLIP_FILE = 32, 0.0, "CLIP/models/ViT-B-32.pt"


class TVTSv2_H_14(TVTSv2Base):
    """v2/model/model_dist_TVTSv2_ViT_H_14.py:13-158: OpenCLIP ViT-H/14 towers (video width 1280 / 32 layers / 16 heads of dim 80,
    text 1024 / 24 layers / 16 heads, embed 1024, exact GELU), mask ratio 0.7, sort head over the PATCH tokens (no CLS), 16 heads.
    The reference runs forward under fp16 autocast (:96); here the GEMM operands are bf16 with fp32 accumulation like the B models.
    The OpenCLIP checkpoint (`create_model('ViT-H-14', pretrained='laion2b_s32b_b79k', cache_dir='OpenCLIP/models')`, :22-24) is read
    from OpenCLIP/models/open_clip_pytorch_model.bin when that file exists."""

    PATCH, MASK_RATIO, CLIP_FILE = 14, 0.7, "OpenCLIP/models/open_clip_pytorch_model.bin"
    OPEN_CLIP = True

    def _default_arch(self):
        return C.TVTSV2_H_14

    def __init__(self, args, load_checkpoint=None, arch=None):
        super().__init__(args, load_checkpoint, arch=arch)
        ctx = self.arch.context
        self.text_attn_mask = torch.full((ctx, ctx), float("-inf")).triu_(1)     # OpenCLIP build_attention_mask; plain attribute (:36)

    def _build_video_model(self, arch):
        return VisionTransformerH14(image_size=arch.resolution, patch_size=arch.patch, width=arch.width, layers=arch.layers,
                                    heads=arch.heads, mlp_ratio=4.0, output_dim=arch.embed_dim, act_layer=nn.GELU,
                                    num_frames=arch.num_frames, mask_ratio=arch.mask_ratio)

    def compute_video(self, video, keep_ind):
        """-> (video_before_embeddings = patch tokens [B, N-1, E], video_embeddings = pooled CLS [B, E])   (:155-157)"""
        pooled, tokens = self.video_model(video, keep_ind)
        return tokens, pooled


# ------------------------------------------------------------------------------------------------ downstream (zero-shot / feature extraction)
class TVTSv2Downstream(TVTSv2Base):
    """v2/downstream/model_TVTSv2_ViT_B_16.py:10-98 (and _B_32): the pre-training towers with mask_ratio 0 and no sort head, used
    forward-only by zero_ret_* / zero_recognition_* / feature_extraction_*.  Constructor takes only `load_checkpoint`; the
    state_dict has no `pred_model.*` keys; forward returns (text_embeddings, video_embeddings)."""

    SORT_HEAD = False
    MEAN_OVER_CLIPS = True        # False in the *_mc variants (model_TVTSv2_ViT_B_16_mc.py:64: the mean over n_trans is commented out)

    def _default_arch(self):
        return {16: C.TVTSV2_B_16, 32: C.TVTSV2_B_32, 14: C.TVTSV2_H_14}[self.PATCH].small(mask_ratio=0.0)

    def __init__(self, load_checkpoint=None, arch=None):
        super().__init__(None, load_checkpoint, arch=arch.small(mask_ratio=0.0) if isinstance(arch, C.ArchConfig) else None)

    def _checkpoint_location(self):
        return None               # downstream/model_TVTSv2_ViT_B_16.py:42 torch.load(load_checkpoint)

    def forward(self, data, return_embeds=True):
        text, video, keep_ind = data["text"], data["video"], data["keep_ind"]
        B = video.shape[0]
        t, _ = self.compute_text(text)
        n_trans = t.shape[0] // B
        if self.MEAN_OVER_CLIPS:
            text_embeddings = E.group_mean(t, n_trans)                # :62-64
        else:
            text_embeddings = t.reshape(n_trans, B, t.shape[-1])      # _mc: [n_choices, B, E]
        _, video_embeddings = self.compute_video(video, keep_ind)
        if return_embeds:
            return text_embeddings, video_embeddings
        return sim_matrix(text_embeddings, video_embeddings)


class TVTSv2_B_16_downstream(TVTSv2Downstream):
    PATCH, CLIP_FILE = 16, "CLIP/models/ViT-B-16.pt"


class TVTSv2_B_32_downstream(TVTSv2Downstream):
    PATCH, CLIP_FILE = 32, "CLIP/models/ViT-B-32.pt"


class TVTSv2_B_16_downstream_mc(TVTSv2_B_16_downstream):
    MEAN_OVER_CLIPS = False


class TVTSv2_B_32_downstream_mc(TVTSv2_B_32_downstream):
    MEAN_OVER_CLIPS = False


class TVTSv2_H_14_downstream(TVTSv2Downstream):
    """v2/downstream/model_TVTSv2_ViT_H_14.py: the H/14 towers with mask_ratio 0 and no sort head."""
    PATCH, CLIP_FILE, OPEN_CLIP = 14, "OpenCLIP/models/open_clip_pytorch_model.bin", True
    _build_video_model = TVTSv2_H_14._build_video_model
    compute_video = TVTSv2_H_14.compute_video

    def __init__(self, load_checkpoint=None, arch=None):
        super().__init__(load_checkpoint, arch)
        ctx = self.arch.context
        self.text_attn_mask = torch.full((ctx, ctx), float("-inf")).triu_(1)


class TVTSv2_H_14_downstream_mc(TVTSv2_H_14_downstream):
    MEAN_OVER_CLIPS = False
