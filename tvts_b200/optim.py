"""Parameter-group policy and optimizer of the reference entry script, on a flat arena with one fused CUDA kernel.

Mirrors (paths relative to /root/reference):
  param groups / freezing    v2/train_dist_TVTSv2_ViT_B_16.py:66-124   (4 groups: new/CLIP x decay/no-decay; text layers 0-8 frozen)
  transformers.AdamW         v2/train_dist_TVTSv2_ViT_B_16.py:125      (transformers==4.10.2, un-vendored: betas (0.9,0.999),
                             eps 1e-6, correct_bias=True, decoupled weight decay applied after the update; parameters whose
                             grad is None are skipped and their step counter does not advance)
  per-epoch decay            v2/trainer/trainer.py:402-411             (param_group['lr'] *= 0.1: just mutate .param_groups)

FlatState owns ONE fp32 arena each for master weights, gradients and the two moments (same offsets, each tensor padded to
a whole chunk) plus a bf16 arena with the GEMM-operand copy of the weights.  The nn.Parameters are re-pointed at views of
the master arena, the engine writes gradients straight into the gradient arena (engine.ParamView.gbuf), the gradient
all-reduce is one NCCL call over the arena, and `AdamW.step()` is one kernel launch (tvts_adamw_flat) that also refreshes
the bf16 copy.
"""
import math

import torch

from . import _lib as L

CHUNK = 4096

_CURRENT = None


def current():
    return _CURRENT


class FlatState:
    def __init__(self, params, chunk=CHUNK):
        """params: list of trainable nn.Parameters (all on one CUDA/CPU device, fp32)."""
        global _CURRENT
        self.params = list(params)
        assert self.params, "FlatState needs at least one parameter"
        dev = self.params[0].device
        self.device = dev
        self.chunk = chunk
        self.offsets, self.index = [], {}
        off = 0
        chunk_tensor = []
        for i, p in enumerate(self.params):
            assert p.dtype == torch.float32 and p.device == dev
            self.offsets.append(off)
            self.index[id(p)] = i
            nch = (p.numel() + chunk - 1) // chunk
            chunk_tensor += [i] * nch
            off += nch * chunk
        self.total = off
        self.n_chunks = len(chunk_tensor)
        self.p = torch.zeros(off, dtype=torch.float32, device=dev)
        self.g = torch.zeros(off, dtype=torch.float32, device=dev)
        self.m = torch.zeros(off, dtype=torch.float32, device=dev)
        self.v = torch.zeros(off, dtype=torch.float32, device=dev)
        self.bf = torch.zeros(off, dtype=L.OPERAND_DTYPE, device=dev)
        self.chunk_tensor = torch.tensor(chunk_tensor, dtype=torch.int32, device=dev)
        self.synced = [-1] * len(self.params)       # parameter ._version for which self.bf holds the bf16 copy
        self.touched = set()
        with torch.no_grad():
            for i, p in enumerate(self.params):
                view = self.p[self.offsets[i]: self.offsets[i] + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                if p.grad is not None:
                    p.grad = None
        _CURRENT = self

    def release(self):
        global _CURRENT
        if _CURRENT is self:
            _CURRENT = None

    # ---- views -------------------------------------------------------------------------------------------------
    def has(self, p):
        i = self.index.get(id(p))           # the arena keeps its parameters alive, so a matching id is the same object
        return i is not None and self.params[i] is p

    def _view(self, arena, p):
        i = self.index[id(p)]
        return arena[self.offsets[i]: self.offsets[i] + p.numel()].view(p.shape)

    def grad_view(self, p):
        self.touched.add(self.index[id(p)])
        return self._view(self.g, p)

    def bf16_view(self, p):
        """bf16 operand copy of p, re-cast only if p changed outside the fused optimizer."""
        i = self.index[id(p)]
        view = self._view(self.bf, p)
        if self.synced[i] != p._version:
            L.call("cast_bf16", p.detach(), view, p.numel())
            self.synced[i] = p._version
        return view

    def refresh_bf16(self):
        """Re-cast the bf16 copy of every operand that changed outside the fused optimizer (load_state_dict, manual edits).  Eager
        steps get this lazily through bf16_view(); a captured CUDA graph reads the arena copy directly, so TrainStep calls this
        before every replay (a few hundred integer compares when nothing changed)."""
        n = 0
        for i, p in enumerate(self.params):
            if self.synced[i] != -1 and self.synced[i] != p._version:
                L.call("cast_bf16", p.detach(), self._view(self.bf, p), p.numel())
                self.synced[i] = p._version
                n += 1
        return n

    # ---- step bookkeeping ----------------------------------------------------------------------------------------
    def zero_grad(self):
        """One memset over the gradient arena; every parameter's .grad becomes None (torch 2.x zero_grad semantics)."""
        self.g.zero_()
        for i in self.touched:
            self.params[i].grad = None
        self.touched.clear()

    def active_mask(self):
        """Which parameters currently carry a gradient (what AdamW.prepare() keys on)."""
        return [p.grad is not None for p in self.params]

    def set_active(self, mask):
        """Re-create the .grad attachment pattern of an earlier step (CUDA-graph replay: the Python-side bookkeeping of the step
        that was captured -- which parameters received a gradient -- has to be restored before the optimizer looks at it)."""
        self.touched = set()
        for i, (p, on) in enumerate(zip(self.params, mask)):
            if on:
                p.grad = self._view(self.g, p)
                self.touched.add(i)
            else:
                p.grad = None

    def publish_grads(self, params=None):
        """Attach arena views as .grad of every parameter the backward wrote (what autograd's AccumulateGrad would do)."""
        for i in self.touched:
            p = self.params[i]
            if p.grad is None:
                p.grad = self._view(self.g, p)


def reference_param_groups(named_parameters, text_layers=12, tune_from=9):
    """The 4 optimizer groups of the reference entry script; freezes (requires_grad=False) the un-tuned text layers."""
    no_decay_names = ["bias", "LayerNorm", "ln_", "norm", "ls_", "LayerScale"]   # the last two: H_14 script :17 (no such parameters in B)
    text_tune_layers = ["resblocks.%d." % i for i in range(tune_from, text_layers)]
    decay_clip, no_decay_clip, decay_new, no_decay_new = [], [], [], []
    for name, param in named_parameters:
        nd = any(s in name for s in no_decay_names)
        if "video_model" in name:
            if "timeattn" in name or "ln_3" in name or "ls_3" in name:
                (no_decay_new if nd else decay_new).append(param)
            else:
                (no_decay_clip if nd else decay_clip).append(param)
        elif "text" in name:
            if "resblocks" in name:
                if any(tl in name for tl in text_tune_layers):
                    (no_decay_clip if nd else decay_clip).append(param)
                else:
                    param.requires_grad = False
            else:
                (no_decay_clip if nd else decay_clip).append(param)
        else:
            (no_decay_new if nd else decay_new).append(param)
    return [
        {"params": decay_new, "weight_decay": 0.05, "lr": 1e-4},
        {"params": no_decay_new, "weight_decay": 0.0, "lr": 1e-4},
        {"params": decay_clip, "weight_decay": 0.05, "lr": 1e-7},
        {"params": no_decay_clip, "weight_decay": 0.0, "lr": 1e-7},
    ]


class AdamW:
    """transformers.AdamW semantics (see module docstring) over a FlatState; `param_groups` is mutable like torch's."""

    GROWTH_INTERVAL = 2000          # torch.cuda.amp.GradScaler defaults (x2 after 2000 finite steps, x0.5 + skipped step on overflow)
    MAX_SCALE = 2.0 ** 24

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True, dynamic_scale=None,
                 init_scale=None):
        """dynamic_scale (default: on for the IEEE-half operand build, off for bf16): the backward runs under a loss scale that lives
        on the device (`scale_tensor`); launch() checks the gradient arena for non-finite values, skips the whole update when it finds
        one and adapts the scale -- all inside the captured step, no host synchronisation."""
        groups = list(params)
        if groups and not isinstance(groups[0], dict):
            groups = [{"params": groups}]
        self.param_groups = []
        for g in groups:
            d = {"lr": lr, "betas": betas, "eps": eps, "weight_decay": weight_decay, "correct_bias": correct_bias}
            d.update(g)
            d["params"] = [p for p in d["params"] if p.requires_grad]
            self.param_groups.append(d)
        flat = [p for g in self.param_groups for p in g["params"]]
        self.flat = FlatState(flat)
        self.steps = [0] * len(flat)
        self.group_of = [gi for gi, g in enumerate(self.param_groups) for _ in g["params"]]
        b, e = self.param_groups[0]["betas"], self.param_groups[0]["eps"]
        assert all(g["betas"] == b and g["eps"] == e for g in self.param_groups), "per-group betas/eps are not supported"
        self._ring = [torch.zeros(len(flat), 4, dtype=torch.float32).pin_memory() if self.flat.device.type == "cuda"
                      else torch.zeros(len(flat), 4, dtype=torch.float32) for _ in range(4)]
        self._ring_ev = [None] * 4
        self._slot = 0
        self.table = torch.zeros(len(flat), 4, dtype=torch.float32, device=self.flat.device)
        self.dynamic_scale = (L.OPERAND == "fp16") if dynamic_scale is None else bool(dynamic_scale)
        self._init_scale = float(init_scale if init_scale is not None else (L.DEFAULT_LOSS_SCALE if L.DEFAULT_LOSS_SCALE != 1.0 else 1024.0))
        self._make_scale_state()

    def _make_scale_state(self, scale=None, steps=None):
        dev = self.flat.device
        self.scale_state = torch.tensor([self._init_scale if scale is None else scale, 0.0, 0.0, 0.0], dtype=torch.float32, device=dev)
        self.steps_dev = torch.tensor(self.steps if steps is None else steps, dtype=torch.int32, device=dev)
        self.scale_tensor = self.scale_state[0]          # 0-d view: `loss * optimizer.scale_tensor` is what the backward starts from

    def sync_steps(self):
        """Device-side step counters -> self.steps (dynamic scaling only: skipped steps do not count, and only the device knows)."""
        if self.dynamic_scale:
            self.steps = [int(x) for x in self.steps_dev.tolist()]
        return self.steps

    @property
    def skipped_steps(self):
        return int(self.scale_state[3].item()) if self.dynamic_scale else 0

    def follow_parameters(self):
        """Re-create the flat arenas on the device the parameters live on NOW.  The reference entry scripts build the optimizer while
        the model is still on the CPU and let the trainer move the model afterwards (v2/base/base_trainer.py:19-21): `model.to(device)`
        gives every parameter a fresh storage, so the arenas (and the parameters' arena views) have to be rebuilt there.  Adam moments
        and step counters are carried over.  No-op when the parameters still are views of the arena."""
        fs = self.flat
        p0 = fs.params[0]
        if p0.device == fs.device and p0.data_ptr() == fs._view(fs.p, p0).data_ptr():
            return False
        old_m, old_v = fs.m, fs.v
        fs.release()
        self.flat = FlatState(fs.params, fs.chunk)
        self.flat.m.copy_(old_m)
        self.flat.v.copy_(old_v)
        dev = self.flat.device
        n = len(self.flat.params)
        self._ring = [torch.zeros(n, 4, dtype=torch.float32).pin_memory() if dev.type == "cuda" else torch.zeros(n, 4, dtype=torch.float32)
                      for _ in range(4)]
        self._ring_ev = [None] * 4
        self._slot = 0
        self.table = torch.zeros(n, 4, dtype=torch.float32, device=dev)
        self._make_scale_state(scale=float(self.scale_state[0].item()), steps=[int(x) for x in self.steps_dev.tolist()])
        return True

    def zero_grad(self, set_to_none=True):
        self.flat.zero_grad()

    def prepare(self):
        """Host half of step(): advance the step counters of the parameters that have a gradient, refresh the per-tensor
        hyper-parameter table and upload it (async, pinned ring).  Stream-ordered before launch()."""
        fs = self.flat
        slot = self._slot
        self._slot = (slot + 1) % len(self._ring)
        if self._ring_ev[slot] is not None:
            self._ring_ev[slot].synchronize()
        host = self._ring[slot]
        rows = host.numpy()
        for i, p in enumerate(fs.params):
            if p.grad is None:
                rows[i, 2] = 0.0
                continue
            g = self.param_groups[self.group_of[i]]
            rows[i, 1] = g["lr"] * g["weight_decay"] if g["weight_decay"] > 0.0 else 0.0
            rows[i, 2] = 1.0
            if self.dynamic_scale:              # the kernel evaluates the bias correction from its own step counters
                rows[i, 0] = g["lr"]
                rows[i, 3] = 1.0 if g["correct_bias"] else 0.0
                continue
            self.steps[i] += 1
            t = self.steps[i]
            b1, b2 = g["betas"]
            step_size = g["lr"]
            if g["correct_bias"]:
                step_size = step_size * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
            rows[i, 0] = step_size
        self.table.copy_(host, non_blocking=True)
        if fs.device.type == "cuda":
            ev = torch.cuda.Event()
            ev.record()
            self._ring_ev[slot] = ev

    def launch(self, grad_scale=1.0):
        """Device half of step(): ONE kernel over the arena (capturable in a CUDA graph: it only reads device-resident state)."""
        fs = self.flat
        b1, b2 = self.param_groups[0]["betas"]
        if self.dynamic_scale:
            L.call("adamw_flat_dyn", fs.p, fs.g, fs.m, fs.v, fs.bf, self.chunk_tensor_arg(), self.table, self.steps_dev, self.scale_state,
                   len(fs.params), fs.n_chunks, fs.chunk, float(b1), float(b2), float(self.param_groups[0]["eps"]),
                   float(self.GROWTH_INTERVAL), float(self.MAX_SCALE))
        else:
            L.call("adamw_flat", fs.p, fs.g, fs.m, fs.v, fs.bf, self.chunk_tensor_arg(), self.table, fs.n_chunks, fs.chunk, float(b1),
                   float(b2), float(self.param_groups[0]["eps"]), float(grad_scale))
        for i, p in enumerate(fs.params):
            if p.grad is not None:
                fs.synced[i] = p._version      # the kernel refreshed the bf16 copy of every active tensor

    # ---- launch() in phases: the data-parallel step updates the arena bucket by bucket while the next bucket's gradient all-reduce is
    # still on the wire (trainer.TrainStep, pipelined AdamW).  launch_check() -> [share scale_state[2] across ranks] -> launch_range() per
    # bucket of whole chunks -> launch_finish(); together they are exactly launch().
    def launch_check(self):
        """dynamic scale only: finite check over the (LOCAL) gradient arena -> scale_state[2]"""
        fs = self.flat
        if self.dynamic_scale:
            L.call("adamw_dyn_check", fs.g, fs.n_chunks * fs.chunk, self.scale_state)

    def launch_range(self, c0, c1, grad_scale=1.0):
        """the update of chunks [c0, c1) of the arena"""
        fs = self.flat
        if c1 <= c0:
            return
        b1, b2 = self.param_groups[0]["betas"]
        a, b = c0 * fs.chunk, c1 * fs.chunk
        ct = self.chunk_tensor_arg()[c0:c1]
        if self.dynamic_scale:
            L.call("adamw_dyn_apply", fs.p[a:b], fs.g[a:b], fs.m[a:b], fs.v[a:b], fs.bf[a:b], ct, self.table, self.steps_dev, self.scale_state,
                   c1 - c0, fs.chunk, float(b1), float(b2), float(self.param_groups[0]["eps"]))
        else:
            L.call("adamw_flat", fs.p[a:b], fs.g[a:b], fs.m[a:b], fs.v[a:b], fs.bf[a:b], ct, self.table, c1 - c0, fs.chunk, float(b1),
                   float(b2), float(self.param_groups[0]["eps"]), float(grad_scale))

    def launch_finish(self):
        fs = self.flat
        if self.dynamic_scale:
            L.call("adamw_dyn_finish", self.steps_dev, self.table, self.scale_state, len(fs.params), float(self.GROWTH_INTERVAL),
                   float(self.MAX_SCALE))
        for i, p in enumerate(fs.params):
            if p.grad is not None:
                fs.synced[i] = p._version      # the kernels refreshed the bf16 copy of every active tensor

    def step(self, grad_scale=1.0):
        self.prepare()
        self.launch(grad_scale)

    def chunk_tensor_arg(self):
        return self.flat.chunk_tensor

    def state_dict(self):
        """The `torch.optim.Optimizer.state_dict()` layout `transformers.AdamW` (4.10.2) writes into the reference's checkpoints
        (v2/base/base_trainer.py:173-181): param ids = positions in group order; per stepped parameter {'step' (int), 'exp_avg',
        'exp_avg_sq'} shaped like the parameter; groups carry lr / betas / eps / weight_decay / correct_bias."""
        fs = self.flat
        self.sync_steps()
        groups, idx = [], 0
        for g in self.param_groups:
            d = {k: v for k, v in g.items() if k != "params"}
            d["params"] = list(range(idx, idx + len(g["params"])))
            idx += len(g["params"])
            groups.append(d)
        state = {}
        for i, p in enumerate(fs.params):
            if self.steps[i] > 0:
                state[i] = {"step": self.steps[i], "exp_avg": fs._view(fs.m, p).clone(), "exp_avg_sq": fs._view(fs.v, p).clone()}
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        """Inverse of state_dict(); also accepts a checkpoint the reference's own optimizer wrote for the same model (same group
        sizes: the groups are built from named_parameters() in the same order)."""
        fs = self.flat
        sizes = [len(g["params"]) for g in sd["param_groups"]]
        if sizes != [len(g["params"]) for g in self.param_groups]:
            raise ValueError("loaded state dict has parameter groups of sizes %s, optimizer has %s"
                             % (sizes, [len(g["params"]) for g in self.param_groups]))
        ids = [i for g in sd["param_groups"] for i in g["params"]]
        self.steps = [0] * len(fs.params)
        fs.m.zero_()
        fs.v.zero_()
        with torch.no_grad():
            for pos, pid in enumerate(ids):
                st = sd["state"].get(pid)
                if st is None:
                    continue
                p = fs.params[pos]
                if tuple(st["exp_avg"].shape) != tuple(p.shape):
                    raise ValueError("optimizer state %d has shape %s, parameter has %s" % (pid, tuple(st["exp_avg"].shape), tuple(p.shape)))
                self.steps[pos] = int(st["step"])
                fs._view(fs.m, p).copy_(st["exp_avg"])
                fs._view(fs.v, p).copy_(st["exp_avg_sq"])
        for g, s_ in zip(self.param_groups, sd["param_groups"]):
            g.update({k: (tuple(v) if k == "betas" else v) for k, v in s_.items() if k != "params"})
        self.steps_dev.copy_(torch.tensor(self.steps, dtype=torch.int32))


def build_reference_optimizer(model, text_layers=None, tune_from=None):
    """What v2/train_dist_TVTSv2_ViT_B_16.py:66-125 builds for `model` (train_dist_TVTSv2_ViT_H_14.py:68-127: text layers 0-17 of 24
    frozen; same 4 groups)."""
    if text_layers is None:
        text_layers = model.arch.text_layers
    if tune_from is None:
        tune_from = (text_layers * 3) // 4
    groups = reference_param_groups(list(model.named_parameters()), text_layers, tune_from)
    return AdamW(groups)
