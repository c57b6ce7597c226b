"""End-to-end drop-in check: the reference's UNMODIFIED entry script `v2/train_dist_TVTSv2_ViT_B_16.py` is executed (runpy, `__main__`)
with `tvts_b200/dropin` ahead of the reference tree on sys.path -- its own `ConfigParser`, `utils`, `logger` and `base` packages, its own
parameter-group / freeze code, `Trainer_TVTSv2_B_16(...).train()` -- for one epoch of two synthetic loaders (a transcript batch with
both losses, a caption batch with InfoNCE only) plus validation.

What has to be stubbed is the environment, not the path under test: packages that are not installed here (`sacred`, `humanize`, `dominate`; `ftfy` behind
`CLIP.clip.tokenize`), the data stack (`data_loader.data_loader`: video decoding), CUDA / NCCL (this test runs on the CPU: the C ABI is
routed to the torch restatement tests/emu.py, `nccl` becomes `gloo`), and `transformers.AdamW` (removed from current `transformers`;
INTEGRATION.md: swap in `tvts_b200.optim.AdamW`).  Skipped when the reference tree is not present (e.g. on the GPU box)."""
import json
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/v2"

RUNNER = r'''
import os, runpy, sys, types
root, ref, stub, cfg = sys.argv[1:5]
sys.path[:0] = [stub, os.path.join(root, "tvts_b200", "dropin"), root, os.path.join(root, "tests"), ref]
import torch
import torch.distributed as dist
import emu
emu.install()                                              # C ABI -> torch restatement (CPU)
torch.cuda.set_device = lambda *a, **k: None
_init = dist.init_process_group
dist.init_process_group = lambda backend=None, **k: _init("gloo", **k)
import transformers
from tvts_b200 import optim
transformers.AdamW = lambda params, **k: optim.AdamW(params, **k)
script = os.path.join(ref, "train_dist_TVTSv2_ViT_B_16.py")
sys.argv = [script, "-c", cfg]
runpy.run_path(script, run_name="__main__")
print("SCRIPT-FINISHED")
'''

STUBS = {
    "sacred/__init__.py": '''
        class Experiment:
            def __init__(self, name): self.name = name
            def main(self, fn): return fn
            def add_config(self, cfg): self.cfg = cfg
            def run(self): raise RuntimeError("neptune path is not exercised")
    ''',
    "humanize/__init__.py": "def naturalsize(x, *a, **k): return str(x)\n",
    "dominate/__init__.py": "class document:\n    def __init__(self, *a, **k): pass\n",          # utils/html.py (web visualiser, unused here)
    "dominate/tags.py": "meta = h3 = table = tr = td = p = a = img = br = video = source = attr = span = None\n",
    "CLIP/__init__.py": "",
    "CLIP/clip.py": '''
        import torch
        def tokenize(texts, context_length=77, truncate=False):
            """CLIP token rows [SOT, ids..., EOT, 0...] from a hash of the words (the BPE tokenizer needs ftfy / regex data files)."""
            sot, eot = 49406, 49407
            out = torch.zeros(len(texts), context_length, dtype=torch.int)
            for i, t in enumerate(texts):
                ids = [1 + (hash(w) % 40000) for w in t.split()][:context_length - 2]
                row = [sot] + ids + [eot]
                out[i, :len(row)] = torch.tensor(row)
            return out
    ''',
    "data_loader/__init__.py": "",
    "data_loader/data_loader.py": '''
        import torch
        class _Sampler:
            def set_epoch(self, e): pass
        class MultiDistTextVideoDataLoader:
            """one synthetic batch per epoch, shaped like the reference loaders' (YTTemporal: 4 transcripts per clip + labels; WebVid: 1 caption)"""
            def __init__(self, args=None, dataset_name="", batch_size=1, split="train", patches_per_frame=196, mask_ratio=0.5, **unused):
                self.dataset_name, self.batch_size, self.split = dataset_name, 1, split
                self.n_samples, self.train_sampler = 1, _Sampler()
                self.yt = dataset_name.startswith("YT")
                self.n_keep = int(patches_per_frame * (1 - mask_ratio))
                self.P = patches_per_frame
            def __len__(self): return 1
            def __iter__(self):
                g = torch.Generator().manual_seed(7 if self.yt else 8)
                T, nt = (2, 4) if self.yt else (3, 1)
                text = [[f"clip {c} of a {'lecture' if self.yt else 'web'} video number {b}" for b in range(self.batch_size)] for c in range(nt)]
                batch = {"video": torch.randn(self.batch_size, T, 3, 224, 224, generator=g), "text": text,
                         "keep_ind": torch.stack([torch.randperm(self.P, generator=g)[:self.n_keep] for _ in range(self.batch_size)]),
                         "label": torch.arange(nt).repeat(self.batch_size, 1)}
                yield batch
    ''',
}


@pytest.mark.timeout(1500)
def test_unmodified_reference_entry_script_runs_on_the_dropin_tree(tmp_path):
    if not os.path.isfile(os.path.join(REF, "train_dist_TVTSv2_ViT_B_16.py")):
        pytest.skip("reference tree not present")
    stub = tmp_path / "stubs"
    for rel, src in STUBS.items():
        f = stub / rel
        f.parent.mkdir(parents=True, exist_ok=True)
        f.write_text(textwrap.dedent(src))
    cfg = json.load(open(os.path.join(REF, "configs", "dist-yt-web-pt-vit-b-16.json")))
    cfg["n_gpu"] = 0
    cfg["trainer"].update(epochs=1, save_dir=str(tmp_path / "results"), save_period=0, monitor="off", init_val=False)
    cfg_path = tmp_path / "cfg.json"
    cfg_path.write_text(json.dumps(cfg))
    runner = tmp_path / "runner.py"
    runner.write_text(RUNNER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29800 + os.getpid() % 150), WORLD_SIZE="1", RANK="0", LOCAL_RANK="0",
               TVTS_ALLOW_RANDOM_CLIP="1", PYTHONHASHSEED="0", OMP_NUM_THREADS=str(os.cpu_count() or 1))
    r = subprocess.run([sys.executable, str(runner), ROOT, REF, str(stub), str(cfg_path)], capture_output=True, text=True, timeout=1400,
                       env=env, cwd=str(tmp_path))
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-4000:]
    assert "SCRIPT-FINISHED" in r.stdout
    # one step per loader: the transcript batch carries both losses, the caption batch InfoNCE only
    lines = [l for l in r.stdout.splitlines() if l.startswith("Train Epoch: 1 dl")]
    assert len(lines) == 2, out[-3000:]
    yt, web = lines
    assert "dl0" in yt and "dl1" in web
    loss_ce = lambda l: float(l.split("Loss_ce:")[1].split()[0])
    loss_ct = lambda l: float(l.split("Loss_ct:")[1].split()[0])
    assert loss_ce(yt) > 0.5 and loss_ce(web) == 0.0 and loss_ct(yt) >= 0.0 and loss_ct(web) >= 0.0
    assert "val_0_t2v_metrics_R1" in r.stdout            # the validation pass of both loaders ran and reported through the trainer's log


# ------------------------------------------------------------------------------------------------ TVTS v1
RUNNER_V1 = r'''
import os, runpy, sys, types
root, ref, stub, cfg = sys.argv[1:5]
sys.path[:0] = [stub, os.path.join(root, "tvts_b200", "dropin_v1"), root, os.path.join(root, "tests"), ref]
import torch
import torch.distributed as dist
import emu
emu.install()                                              # C ABI -> torch restatement (CPU)
torch.cuda.set_device = lambda *a, **k: None
_init = dist.init_process_group
dist.init_process_group = lambda backend=None, **k: _init("gloo", **k)
import transformers
from tvts_b200 import optim
transformers.AdamW = lambda params, **k: optim.AdamW(params, **k)
class _Tok:                                                # HF tokenizer call signature (v1/trainer/trainer.py:130-131)
    def __call__(self, texts, return_tensors="pt", padding=True, truncation=True, max_length=50):
        lens = [min(max_length, 2 + len(t.split())) for t in texts]
        Lc = max_length if padding == "max_length" else max(lens)
        ids = torch.zeros(len(texts), Lc, dtype=torch.long)
        for i, t in enumerate(texts):
            row = [101] + [1000 + (hash(w) % 20000) for w in t.split()][:lens[i] - 2] + [102]
            ids[i, :len(row)] = torch.tensor(row)
        return {"input_ids": ids, "attention_mask": (ids != 0).long()}
transformers.AutoTokenizer = types.SimpleNamespace(from_pretrained=lambda *a, **k: _Tok())
script = os.path.join(ref, "train_dist_TVTS.py")
sys.argv = [script, "-c", cfg]
runpy.run_path(script, run_name="__main__")
print("SCRIPT-FINISHED")
'''

STUBS_V1 = dict(STUBS)
STUBS_V1.pop("CLIP/__init__.py"), STUBS_V1.pop("CLIP/clip.py")
STUBS_V1.update({
    "neptunecontrib/__init__.py": "", "neptunecontrib/monitoring/__init__.py": "",
    "neptunecontrib/monitoring/sacred.py": "class NeptuneObserver:\n    def __init__(self, *a, **k): pass\n",
    "pims/__init__.py": "",
    "cv2/__init__.py": "def setNumThreads(n): pass\nclass ocl:\n    @staticmethod\n    def setUseOpenCL(flag): pass\n",
    "data_loader/data_loader.py": '''
        import torch
        class _Sampler:
            def set_epoch(self, e): pass
        class MultiDistTextVideoDataLoader:
            """one synthetic batch per epoch shaped like v1's YTTemporal loader: 4 transcripts per clip, an independent keep mask per tube"""
            def __init__(self, args=None, dataset_name="", batch_size=1, split="train", **unused):
                self.dataset_name, self.batch_size, self.split = dataset_name, 1, split
                self.n_samples, self.train_sampler = 1, _Sampler()
            def __len__(self): return 1
            def __iter__(self):
                g = torch.Generator().manual_seed(11)
                T, nt, n = 4, 4, 49
                text = [[f"transcript {c} of lecture number {b} with a few more words" for b in range(self.batch_size)] for c in range(nt)]
                keep = torch.stack([torch.stack([torch.randperm(196, generator=g)[:n] for _ in range(T // 2)]) for _ in range(self.batch_size)])
                yield {"video": torch.randn(self.batch_size, T, 3, 224, 224, generator=g), "text": text, "keep_ind": keep,
                       "label": torch.arange(nt).repeat(self.batch_size, 1)}
    ''',
})


@pytest.mark.timeout(1500)
def test_unmodified_v1_entry_script_runs_on_the_dropin_tree(tmp_path):
    """The same for TVTS v1: `v1/train_dist_TVTS.py` unmodified (its ConfigParser builds the model, the optimizer through
    `config.initialize('optimizer', transformers, ...)` and `Trainer_TVTS`) on `tvts_b200/dropin_v1`."""
    ref = "/root/reference/v1"
    if not os.path.isfile(os.path.join(ref, "train_dist_TVTS.py")):
        pytest.skip("reference tree not present")
    stub = tmp_path / "stubs"
    for rel, src in STUBS_V1.items():
        f = stub / rel
        f.parent.mkdir(parents=True, exist_ok=True)
        f.write_text(textwrap.dedent(src))
    cfg = json.load(open(os.path.join(ref, "configs", "dist-yt-pt.json")))
    cfg["n_gpu"] = 0
    cfg["trainer"].update(epochs=1, save_dir=str(tmp_path / "results"), save_period=0, monitor="off", init_val=False)
    cfg_path = tmp_path / "cfg.json"
    cfg_path.write_text(json.dumps(cfg))
    runner = tmp_path / "runner.py"
    runner.write_text(RUNNER_V1)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29960 + os.getpid() % 30), WORLD_SIZE="1", RANK="0", LOCAL_RANK="0",
               PYTHONHASHSEED="0", OMP_NUM_THREADS=str(os.cpu_count() or 1))
    r = subprocess.run([sys.executable, str(runner), ROOT, ref, str(stub), str(cfg_path)], capture_output=True, text=True, timeout=1400,
                       env=env, cwd=str(tmp_path))
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-4000:]
    assert "SCRIPT-FINISHED" in r.stdout
    lines = [l for l in r.stdout.splitlines() if l.startswith("Train Epoch: 1 dl")]
    assert len(lines) == 1, out[-3000:]
    assert float(lines[0].split("Loss_ce:")[1].split()[0]) > 0.5 and float(lines[0].split("Loss_ct:")[1].split()[0]) >= 0.0
