"""nn.Module mirror of the TVTS v1 model surface (v1/model/model_dist_TVTS.py, v1/model/video_encoder.py): same class names,
constructor / forward signatures and PARAMETER NAMES (state_dicts and checkpoints interchange with the reference and with HF
DistilBERT), with every FLOP routed to the CUDA kernels through tvts_b200.engine_v1.  Modules hold parameters only."""
import types
import warnings
from collections import OrderedDict
from functools import partial

import torch
from torch import nn

from . import engine as E
from . import engine_v1 as E1
from .modules import BaseModel, SortTransformer, _no_direct_call, sim_matrix  # noqa: F401


# ------------------------------------------------------------------------------------------------ video tower
class PatchEmbed(nn.Module):
    """v1/model/video_encoder.py:78-99: Conv3d(3, D, kernel = stride = (tubelet, p, p)) WITH bias."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, num_frames=16, tubelet_size=2):
        super().__init__()
        self.img_size, self.patch_size, self.tubelet_size = (img_size, img_size), (patch_size, patch_size), tubelet_size
        self.num_patches = (img_size // patch_size) ** 2 * (num_frames // tubelet_size)
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=(tubelet_size, patch_size, patch_size),
                              stride=(tubelet_size, patch_size, patch_size))

    forward = _no_direct_call


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden_features, in_features)

    forward = _no_direct_call


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=True):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)

    forward = _no_direct_call


class Block(nn.Module):
    """v1/model/video_encoder.py:59-75 (registration order norm1, attn, norm2, mlp)."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=True, norm_layer=None):
        super().__init__()
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))

    forward = _no_direct_call


class VisionTransformer(nn.Module):
    """v1/model/video_encoder.py:102-226: tubelet ViT with joint attention and a per-tube mask;
    forward(x [B,T,3,R,R], keep_ind [B,T/2,n]) -> [B, 1 + (T/2) n, D] (all tokens after the final norm; pre_logits = Identity)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=0, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.,
                 qkv_bias=True, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0., norm_layer=None, num_frames=16,
                 tubelet_size=2, representation_size=None):
        super().__init__()
        if in_chans != 3 or tubelet_size != 2 or not qkv_bias or qk_scale is not None or num_classes or representation_size:
            raise NotImplementedError("VisionTransformer (v1): only the configuration of v1/model/model_dist_TVTS.py:38-42 is built")
        if drop_rate or attn_drop_rate or drop_path_rate:
            raise NotImplementedError("VisionTransformer (v1): dropout / drop-path are 0 in the reference configuration")
        self.num_features = self.embed_dim = embed_dim
        self.tubelet_size = tubelet_size
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim, num_frames, tubelet_size)
        num_tubes = num_frames // tubelet_size
        self.patches_per_frame = self.patch_embed.num_patches // num_tubes
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patches_per_frame + 1, embed_dim))
        self.temporal_embed = nn.Parameter(torch.zeros(1, num_tubes, embed_dim))
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads, mlp_ratio, qkv_bias, norm_layer) for _ in range(depth)])
        self.norm = (norm_layer or partial(nn.LayerNorm, eps=1e-6))(embed_dim)
        self.pre_logits = nn.Identity()
        self.head = nn.Identity()
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        nn.init.trunc_normal_(self.cls_token, std=.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                nn.init.zeros_(m.bias)
        self.cfg = types.SimpleNamespace(patch=patch_size, width=embed_dim, heads=num_heads, layers=depth)
        self._ordered = None

    def _named(self):
        if self._ordered is None:
            have = dict(self.named_parameters())
            self._ordered = OrderedDict((k, have[k]) for k in E1.video_param_names(self.cfg.layers))
        return self._ordered

    def forward(self, x, keep_ind):
        return E1.video_tower(self.cfg, self._named(), x, keep_ind)


# ------------------------------------------------------------------------------------------------ DistilBERT shell
class _Embeddings(nn.Module):
    def __init__(self, vocab, dim, max_pos, pad_id=0):
        super().__init__()
        self.word_embeddings = nn.Embedding(vocab, dim, padding_idx=pad_id)
        self.position_embeddings = nn.Embedding(max_pos, dim)
        self.LayerNorm = nn.LayerNorm(dim, eps=1e-12)

    forward = _no_direct_call


class _SelfAttention(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.q_lin, self.k_lin, self.v_lin, self.out_lin = (nn.Linear(dim, dim) for _ in range(4))

    forward = _no_direct_call


class _FFN(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.lin1 = nn.Linear(dim, hidden)
        self.lin2 = nn.Linear(hidden, dim)

    forward = _no_direct_call


class _TransformerBlock(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.attention = _SelfAttention(dim)
        self.sa_layer_norm = nn.LayerNorm(dim, eps=1e-12)
        self.ffn = _FFN(dim, hidden)
        self.output_layer_norm = nn.LayerNorm(dim, eps=1e-12)

    forward = _no_direct_call


class _TransformerStack(nn.Module):
    def __init__(self, layers, dim, hidden):
        super().__init__()
        self.layer = nn.ModuleList([_TransformerBlock(dim, hidden) for _ in range(layers)])

    forward = _no_direct_call


class DistilBertShell(nn.Module):
    """Parameter tree of `transformers.DistilBertModel` (same names, shapes and registration order; dropout is not modelled: parity
    runs of the un-vendored text encoder use p = 0).  __call__(input_ids=..., attention_mask=...) returns an object with
    `.last_hidden_state_cls` = last_hidden_state[:, 0]: the only slice the v1 model consumes (model_dist_TVTS.py:126)."""

    def __init__(self, vocab_size=30522, dim=768, n_layers=6, n_heads=12, hidden_dim=3072, max_position_embeddings=512, pad_token_id=0):
        super().__init__()
        self.config = types.SimpleNamespace(hidden_size=dim, dim=dim, n_layers=n_layers, n_heads=n_heads, hidden_dim=hidden_dim,
                                            vocab_size=vocab_size, max_position_embeddings=max_position_embeddings)
        self.embeddings = _Embeddings(vocab_size, dim, max_position_embeddings, pad_token_id)
        self.transformer = _TransformerStack(n_layers, dim, hidden_dim)
        for m in self.modules():                                    # DistilBertPreTrainedModel._init_weights: N(0, 0.02), zero bias
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Embedding):
                nn.init.normal_(m.weight, std=0.02)
        self.cfg = types.SimpleNamespace(text_width=dim, text_heads=n_heads, text_layers=n_layers)
        self._ordered = None

    def _named(self):
        if self._ordered is None:
            have = dict(self.named_parameters())
            self._ordered = OrderedDict((k, have[k]) for k in E1.distil_param_names(self.cfg.text_layers))
        return self._ordered

    def cls_hidden(self, input_ids, attention_mask=None):
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        return E1.distil_cls(self.cfg, self._named(), input_ids, attention_mask)

    def forward(self, input_ids=None, attention_mask=None, **unused):
        return types.SimpleNamespace(last_hidden_state_cls=self.cls_hidden(input_ids, attention_mask))


def load_distilbert(name):
    """`AutoModel.from_pretrained(name)` stand-in: the HF weights are used when they are available locally (no network here);
    otherwise the shell keeps its random init (warned)."""
    shell = DistilBertShell()
    try:
        import transformers
        hf = transformers.AutoModel.from_pretrained(name, local_files_only=True)
        shell.load_state_dict({k: v for k, v in hf.state_dict().items() if k in dict(shell.named_parameters())}, strict=True)
        p_drop = max(float(getattr(hf.config, "dropout", 0.0)), float(getattr(hf.config, "attention_dropout", 0.0)))
        if p_drop > 0.0:         # the reference trains with text_model.train(): HF dropout is active there, and is NOT modelled here
            warnings.warn(f"'{name}' is configured with dropout {p_drop:g}; the tvts_b200 DistilBERT kernels run without dropout "
                          "(v1 pre-training here lacks that regulariser; parity with the reference is shown at p = 0)")
    except Exception as e:   # noqa: BLE001 -- missing cache / package: keep the random init
        warnings.warn(f"'{name}' weights not available locally ({type(e).__name__}): DistilBERT text encoder random-initialised")
    return shell


# ------------------------------------------------------------------------------------------------ top model
class _Proj(nn.Sequential):
    forward = _no_direct_call


class TVTS(BaseModel):
    """v1/model/model_dist_TVTS.py:18-141.  Extra keyword-only arguments build toy-sized models for the parity tests."""

    def __init__(self, args, video_params, text_params, projection_dim=256, load_checkpoint=None, projection="minimal", *,
                 text_model=None, video_model=None, sort_heads=12):
        super().__init__()
        self.args = args
        self.video_params = video_params
        self.text_params = text_params
        if not text_params["pretrained"]:
            raise NotImplementedError("Huggingface text models require pretrained init.")
        if not text_params["model"].startswith("distilbert"):
            raise NotImplementedError("only the DistilBERT text encoder of v1/configs/dist-yt-pt.json is built")
        self.text_model = text_model if text_model is not None else load_distilbert(text_params["model"])
        if video_model is None:
            if video_params.get("arch_config", "base_patch16_224") != "base_patch16_224":
                raise NotImplementedError
            video_model = VisionTransformer(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, qkv_bias=True,
                                            norm_layer=partial(nn.LayerNorm, eps=1e-6), num_frames=video_params.get("num_frames", 16))
            if load_checkpoint in ["", None]:
                import os
                if os.path.isfile("./mae_pretrain_vit_base.pth"):                    # :49-61 MAE IN-1K init, 2-D kernel repeated over dt
                    sd = torch.load("./mae_pretrain_vit_base.pth", map_location="cpu", weights_only=False)["model"]
                    if "patch_embed.proj.weight" in sd and sd["patch_embed.proj.weight"].dim() == 4:
                        sd["patch_embed.proj.weight"] = sd["patch_embed.proj.weight"].unsqueeze(2).repeat(1, 1, 2, 1, 1)
                    video_model.load_state_dict(sd, strict=False)
                    print("ViT initialized with MAE IN-1K weights.")
                else:
                    warnings.warn("./mae_pretrain_vit_base.pth not found: video tower keeps its fresh init")
        self.video_model = video_model
        ftr_dim = video_model.embed_dim
        if projection == "minimal":
            self.txt_proj = _Proj(nn.ReLU(), nn.Linear(self.text_model.config.hidden_size, projection_dim))
            self.vid_proj = _Proj(nn.Linear(ftr_dim, projection_dim))
        else:
            raise NotImplementedError("only projection='minimal' (v1/configs/dist-yt-pt.json) is built")
        self.n_trans = 4
        self.pred_model = SortTransformer(num_classes=self.n_trans, embed_dim=ftr_dim, num_heads=sort_heads)
        if load_checkpoint not in ["", None]:
            checkpoint = torch.load(load_checkpoint, map_location="cuda:{}".format(self.args.local_rank), weights_only=False)
            from .compat import state_dict_data_parallel_fix
            self.load_state_dict(state_dict_data_parallel_fix(checkpoint["state_dict"], self.state_dict()), strict=True)
            print("loading checkpoint from {}".format(load_checkpoint))

    def set_device(self, device):
        self.device = device

    def compute_text(self, text_data):
        text_before = self.text_model.cls_hidden(text_data["input_ids"], text_data.get("attention_mask"))
        lin = self.txt_proj[1]
        return text_before, E1.projection(text_before, lin.weight, lin.bias, relu=True)

    def compute_video(self, video_data, keep_ind):
        tokens = self.video_model(video_data, keep_ind)
        lin = self.vid_proj[0]
        return tokens, E1.projection(tokens[:, 0, :], lin.weight, lin.bias)

    def forward(self, data, return_embeds=True):
        text, video, keep_ind = data["text"], data["video"], data["keep_ind"]
        B = video.shape[0]
        text_before, t = self.compute_text(text)                            # [n_trans*B, W], [n_trans*B, proj] clip-major
        n_trans = t.shape[0] // B
        text_embeddings = E.group_mean(t, n_trans)                          # :107-110
        tokens, video_embeddings = self.compute_video(video, keep_ind)
        if n_trans != 1:
            predict_order = self.pred_model.forward_clip_major(text_before, tokens)     # :101-102 transcripts are detached; raw tokens incl. CLS
        else:
            predict_order = None
        if return_embeds:
            return text_embeddings, video_embeddings, predict_order
        return sim_matrix(text_embeddings, video_embeddings)
