"""`clip.load` stand-in for the TVTSv2 constructors (reference: v2/CLIP/clip/clip.py:93-195 as called from
v2/model/model_dist_TVTSv2_ViT_B_16.py:19).

The reference pulls ONLY the text transformer pieces and the visual state_dict out of the loaded CLIP model.  If the
checkpoint file exists it is read (TorchScript archive or plain state_dict, both as released by OpenAI); otherwise --
there is no network in the build/bench environment -- the text pieces are random-initialised exactly like
CLIP.initialize_parameters and no visual weights are returned (the video tower keeps its own fresh init).
"""
import os
import warnings

import torch

from . import modules as M


def _read_state_dict(path):
    try:
        jit = torch.jit.load(path, map_location="cpu")
        return jit.state_dict()
    except Exception:
        sd = torch.load(path, map_location="cpu", weights_only=False)
        return sd.get("state_dict", sd) if isinstance(sd, dict) else sd.state_dict()


def load(name, arch, device="cpu", open_clip=False, allow_missing=False):
    """-> (CLIPTextParts, visual_state_dict | None).  open_clip=True: the text blocks register their parameters in OpenCLIP's order
    (v2/OpenCLIP/transformer.py:189-214); an OpenCLIP `open_clip_pytorch_model.bin` has the same key layout as an OpenAI state_dict."""
    parts = M.CLIPTextParts(embed_dim=arch.embed_dim, context_length=arch.context, vocab_size=arch.vocab,
                            width=arch.text_width, heads=arch.text_heads, layers=arch.text_layers, open_clip=open_clip)
    visual_sd = None
    if name and os.path.isfile(name):
        sd = {k: v.float() for k, v in _read_state_dict(name).items() if torch.is_tensor(v)}
        text_sd = {k: v for k, v in sd.items()
                   if k.startswith(("transformer.", "token_embedding.", "ln_final.")) or k in ("positional_embedding", "text_projection")}
        missing = parts.load_state_dict(text_sd, strict=False)
        if missing.missing_keys:
            warnings.warn(f"CLIP checkpoint {name} lacks text keys: {missing.missing_keys[:4]}...")
        visual_sd = {k[len("visual."):]: v for k, v in sd.items() if k.startswith("visual.")}
    elif not allow_missing:
        raise RuntimeError(f"Model {name} not found (the reference's clip.load raises here too); pass arch= / load_checkpoint= or set "
                           "TVTS_ALLOW_RANDOM_CLIP=1 to build the towers from a random initialisation")
    else:
        warnings.warn(f"CLIP checkpoint '{name}' not found: text tower random-initialised (CLIP.initialize_parameters), "
                      "video tower keeps its fresh init")
    return parts.float(), visual_sd
