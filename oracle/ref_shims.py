"""Import scaffolding for running the UNMODIFIED reference (read-only at /root/reference) in the
build container.  TEST INFRASTRUCTURE ONLY: used by oracle/make_golden.py to pin the oracle
restatement (oracle/tvts_oracle.py) and to write tests/golden/*.npz.  /root/reference does not exist
on the GPU box, so nothing under tests -m gpu / smoke() / bench.py may import this module.

None of the shims touch arithmetic:
  * timm.models.layers  -> DropPath(identity at p=0), trunc_normal_, to_2tuple, StdConv2dSame
    (imported by v2/model/sort_transformer.py:6, v1/model/video_encoder.py:5)
  * ftfy.fix_text       -> identity (v2/CLIP/clip/simple_tokenizer.py)
  * base.BaseModel      -> loaded straight from v2/base/base_model.py (the real base/__init__.py
    drags in decord / tslearn / pims which are not installed)
  * utils.util.state_dict_data_parallel_fix -> loaded from source text of v2/utils/util.py:25-51
  * clip.load           -> returns a random-init CLIP(...) of the right architecture (no weights offline)
"""
import importlib.machinery
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("TVTS_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "v2", "model"))


def _mod(name, is_pkg=False):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None, is_package=is_pkg)
    if is_pkg:
        m.__path__ = []
    sys.modules[name] = m
    return m


def _load_file(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def install(version="v2"):
    """Make `import model.model_dist_TVTSv2_ViT_B_16` etc. work against the reference tree."""
    import torch
    from torch import nn

    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    root = os.path.join(REF_ROOT, version)

    if "timm" not in sys.modules:
        timm = _mod("timm", True)
        timm_models = _mod("timm.models", True)
        layers = _mod("timm.models.layers")

        class DropPath(nn.Module):
            def __init__(self, p=0.0):
                super().__init__()
                assert p == 0.0

            def forward(self, x):
                return x

        layers.DropPath = DropPath
        layers.trunc_normal_ = torch.nn.init.trunc_normal_
        layers.to_2tuple = lambda x: x if isinstance(x, tuple) else (x, x)
        layers.StdConv2dSame = nn.Conv2d
        timm.models = timm_models
        timm_models.layers = layers
        reg = _mod("timm.models.registry")
        reg.register_model = lambda f: f

    if "ftfy" not in sys.modules:
        ftfy = _mod("ftfy")
        ftfy.fix_text = lambda s: s

    # purge modules of a previously installed version (v1 <-> v2 share package names)
    for k in list(sys.modules):
        if k.split(".")[0] in ("model", "base", "utils", "CLIP", "OpenCLIP", "trainer"):
            del sys.modules[k]
    sys.path[:] = [p for p in sys.path if not p.startswith(REF_ROOT)]
    sys.path.insert(0, root)

    base = _mod("base", True)
    bm = _load_file("base.base_model", os.path.join(root, "base", "base_model.py"))
    base.BaseModel = bm.BaseModel

    utils = _mod("utils", True)
    util = _mod("utils.util")
    src = open(os.path.join(root, "utils", "util.py")).read()
    start = src.index("def state_dict_data_parallel_fix")
    end = src.index("\ndef ", start + 10)
    exec(compile(src[start:end], "utils/util.py", "exec"), util.__dict__)
    utils.util = util
    utils.inf_loop = lambda dl: dl
    return root


def patch_clip_load(patch_size, seed=0):
    """clip.load(...) -> (random-init CLIP with the ViT-B text tower, None).  v2/CLIP/clip/model.py:301-328."""
    import torch
    from CLIP import clip as clip_mod
    from CLIP.clip.model import CLIP

    def _load(name, device="cpu", jit=False, download_root=None):
        g = torch.random.get_rng_state()
        torch.manual_seed(seed)
        m = CLIP(512, 224, 12, 768, patch_size, 77, 49408, 512, 8, 12)
        torch.random.set_rng_state(g)
        return m, None

    clip_mod.load = _load
    return clip_mod
