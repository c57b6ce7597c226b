"""The training step of the reference trainer on the B200 kernels (data-parallel, one process per GPU).

Mirrors (paths relative to /root/reference):
  AllGather_multi                      v2/trainer/trainer.py:41-57    all_gather forward, LOCAL-SLICE backward (no reduce)
  Trainer_TVTSv2_*._train_epoch body   v2/trainer/trainer.py:463-499  tokens/video -> model -> gathers -> sim_matrix ->
                                       NormSoftmaxLoss + 2*CE(sort) -> backward -> optimizer.step
  DDP gradient averaging               v2/base/base_trainer.py:23-25  (find_unused_parameters=True semantics: parameters that
                                       received no gradient on ANY rank keep grad=None and are skipped by the optimizer)

The two embedding gathers of the reference are fused into ONE NCCL all-gather of the concatenated [B_local, 2E] buffer;
the gradient all-reduce runs over a flat fp32 arena (one NCCL call per arena chunk), averaged by 1/W like DDP.
"""
import os
import types

import torch
import torch.distributed as dist

from . import _lib as L
from . import engine as E
from . import modules as M


_COLLECTIVES_OFF = False      # set around the pre-capture warm-up of TrainStep._capture: that run must not issue collectives


def _world():
    if _COLLECTIVES_OFF:
        return 1
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


class NativeComm:
    """The two collectives of the step through the C ABI's own NCCL communicator (include/tvts_b200.h: tvts_comm_*) instead of
    torch.distributed -- selected with TVTS_COMM=native.  torch.distributed is still what brings the ranks together: it carries the
    128-byte NCCL id from rank 0 to the others (the side channel tvts_comm_init asks for).  One instance per process."""

    def __init__(self, device):
        import ctypes
        self._ct = ctypes
        lib = L.lib()
        rank, world = dist.get_rank(), dist.get_world_size()
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (ctypes.c_ubyte * 128)()
            L.check(lib.tvts_comm_unique_id(buf), "comm_unique_id")
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        box = [ident.tolist()]
        dist.broadcast_object_list(box, src=0)
        buf = (ctypes.c_ubyte * 128)(*box[0])
        self.handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            L.check(lib.tvts_comm_init(ctypes.byref(self.handle), buf, ctypes.c_int64(rank), ctypes.c_int64(world)), "comm_init")
        self.rank, self.world, self.device = rank, world, device

    def _stream(self):
        return self._ct.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def all_gather(self, out, local):
        ct = self._ct
        L.check(L.lib().tvts_comm_allgather(self.handle, ct.c_void_p(local.data_ptr()), ct.c_void_p(out.data_ptr()),
                                            ct.c_int64(local.numel() * local.element_size()), self._stream()), "comm_allgather")

    def all_reduce_avg(self, t):
        ct = self._ct
        assert t.dtype == torch.float32 and t.is_contiguous()
        L.check(L.lib().tvts_comm_allreduce(self.handle, ct.c_void_p(t.data_ptr()), ct.c_int64(t.numel()), ct.c_int64(1), self._stream()),
                "comm_allreduce")

    def destroy(self):
        if self.handle:
            torch.cuda.synchronize(self.device)
            L.check(L.lib().tvts_comm_destroy(self.handle), "comm_destroy")
            self.handle = None


_NATIVE = None


def native_comm(device=None):
    """The process's NativeComm when TVTS_COMM=native, more than one rank and a CUDA device; else None (torch.distributed collectives)."""
    global _NATIVE
    if os.environ.get("TVTS_COMM", "torch") != "native" or _world() == 1:
        return None
    if _NATIVE is None:
        if device is None or torch.device(device).type != "cuda":
            return None
        _NATIVE = NativeComm(torch.device(device))
    return _NATIVE


def shutdown_native_comm():
    """Destroy the native communicator (after every CUDA graph that holds its kernels is gone, before the process group goes away)."""
    global _NATIVE
    if _NATIVE is not None:
        _NATIVE.destroy()
        _NATIVE = None


def _all_gather(out, local):
    nc = native_comm(local.device)
    if nc is not None:
        nc.all_gather(out, local)
    else:
        dist.all_gather_into_tensor(out, local)


class AllGather_multi(torch.autograd.Function):
    """apply(tensor, n_gpu, args) -> rank-ordered concatenation; backward = this rank's slice of grad_output.
    (`args.rank` is the GLOBAL rank, v2/trainer/trainer.py:53-57.)"""

    @staticmethod
    def forward(ctx, tensor, n_gpu, args):
        W = _world()
        tensor = tensor.contiguous()
        ctx.rank = getattr(args, "rank", _rank()) if args is not None else _rank()
        ctx.batch_size = tensor.shape[0]
        if W == 1:
            ctx.rank = 0
            return tensor.clone()
        out = torch.empty((W * tensor.shape[0],) + tuple(tensor.shape[1:]), dtype=tensor.dtype, device=tensor.device)
        _all_gather(out, tensor)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output[ctx.batch_size * ctx.rank: ctx.batch_size * (ctx.rank + 1)], None, None


class _GatherPair(torch.autograd.Function):
    """Both embedding gathers of trainer.py:481-482 as ONE collective: [B_local, E] x2 -> [Bg, E] x2."""

    @staticmethod
    def forward(ctx, video, text):
        W, r = _world(), _rank()
        B, Edim = video.shape
        ctx.B, ctx.r = B, (r if W > 1 else 0)
        if W == 1:
            return video.clone(), text.clone()
        local = torch.cat([video, text], 1).contiguous()
        out = torch.empty((W * B, 2 * Edim), dtype=local.dtype, device=local.device)
        _all_gather(out, local)
        return out[:, :Edim].contiguous(), out[:, Edim:].contiguous()

    @staticmethod
    def backward(ctx, gv, gt):
        s = slice(ctx.B * ctx.r, ctx.B * (ctx.r + 1))
        return gv[s], gt[s]


def gather_embeddings(video_embeds, text_embeds):
    return _GatherPair.apply(video_embeds, text_embeds)


def gather_for_fused_loss(video, text):
    """(video_all, text_all, first local row) for engine.fused_losses: the same single collective as _GatherPair, no autograd (the fused
    kernel returns the gradient of the local rows itself, which is what AllGather_multi.backward keeps)."""
    W, r = _world(), _rank()
    if W == 1:
        return video, text, 0
    B, Edim = video.shape
    local = torch.cat([video, text], 1).contiguous()
    out = torch.empty((W * B, 2 * Edim), dtype=local.dtype, device=local.device)
    _all_gather(out, local)
    return out[:, :Edim].contiguous(), out[:, Edim:].contiguous(), r * B


def average_flat(fs):
    """Gradient averaging over the flat arena: one NCCL all-reduce (AVG), no flatten / unflatten copies."""
    if _world() == 1:
        return 0
    nc = native_comm(fs.g.device)
    if nc is not None:
        nc.all_reduce_avg(fs.g)
    elif dist.get_backend() == "nccl":
        dist.all_reduce(fs.g, op=dist.ReduceOp.AVG)
    else:                                   # gloo (CPU tests) has no AVG
        dist.all_reduce(fs.g, op=dist.ReduceOp.SUM)
        fs.g.mul_(1.0 / _world())
    return fs.g.numel() * 4


def _average_tensor(t):
    nc = native_comm(t.device)
    if nc is not None:
        nc.all_reduce_avg(t)
    elif dist.get_backend() == "nccl":
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t.mul_(1.0 / _world())


def arena_ranges(fs, late):
    """Split the gradient arena into maximal contiguous runs of tensors: -> (early, late) lists of (start, end) element offsets, `late[i]`
    telling whether parameter i's gradient becomes final late in the backward (the text tower, whose backward runs last)."""
    early_r, late_r = [], []
    for i, p in enumerate(fs.params):
        start = fs.offsets[i]
        end = fs.offsets[i + 1] if i + 1 < len(fs.params) else fs.total
        dst = late_r if late[i] else early_r
        if dst and dst[-1][1] == start:
            dst[-1] = (dst[-1][0], end)
        else:
            dst.append((start, end))
    return early_r, late_r


def average_ranges(fs, ranges):
    if _world() == 1:
        return
    for a, b in ranges:
        _average_tensor(fs.g[a:b])


def average_gradients(params):
    """DDP semantics on explicit buffers: all-reduce(sum)/W of every parameter that has a gradient on this rank.
    All ranks run the same graph on the hot path (same n_trans per step), so the set of parameters with gradients is
    identical across ranks; parameters without one stay grad=None (find_unused_parameters behaviour under torch 2.x)."""
    W = _world()
    if W == 1:
        return 0
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.mul_(1.0 / W)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return flat.numel() * 4


def trim_text_context(tokens, multiple=8):
    """Cut a CLIP token matrix [n_txt, ctx] (EOT = the row maximum, zeros after it) down to the longest caption of the batch, rounded up
    to `multiple`.  EXACT for the causal text towers: positions after a sequence's EOT never influence its EOT row (the only row that is
    pooled, model_dist_TVTSv2_ViT_B_16.py:104-108) and receive zero gradient, so dropping trailing columns that are padding in EVERY
    sequence changes nothing but the amount of padding work (77 -> typically 24..48 columns: GEMM rows, LayerNorm rows and attention
    of the text tower shrink proportionally).  Intended for HOST tensors (the tokenizer's output); on a device tensor it costs a sync."""
    ctx = tokens.shape[1]
    longest = int(tokens.argmax(dim=-1).max().item()) + 1
    keep = min(ctx, max(multiple, -(-longest // multiple) * multiple))
    return tokens if keep == ctx else tokens[:, :keep].contiguous()


def _flatten(data, prefix=""):
    """{'video': t, 'text': {'input_ids': t, ...}} -> {'video': t, 'text.input_ids': t, ...} (v1 batches nest the tokenizer output)."""
    out = {}
    for k, v in data.items():
        if isinstance(v, dict) or (hasattr(v, "keys") and hasattr(v, "__getitem__") and not torch.is_tensor(v)):
            out.update(_flatten({kk: v[kk] for kk in v.keys()}, prefix + k + "."))
        else:
            out[prefix + k] = v
    return out


def _unflatten(flat):
    out = {}
    for k, v in flat.items():
        d = out
        parts = k.split(".")
        for part in parts[:-1]:
            d = d.setdefault(part, {})
        d[parts[-1]] = v
    return out


class TrainStep:
    """One optimizer step of Trainer_TVTSv2_*._train_epoch for a tokenised batch.

    data: {'video' [B,T,3,R,R] f32, 'text' [n_trans*B, ctx] int (clip-major), 'keep_ind' [B,n] int64, 'label' [B,n_trans] int64}
    (host or device tensors; host tensors are copied like trainer.py:474-475,489).
    Returns (loss1, loss2) as 0-d device tensors (no host sync here; the caller decides when to .item()).

    use_graph=True captures the whole step (memset, ~850 kernel launches, NCCL collectives, AdamW) into ONE CUDA graph per input
    signature and replays it: the Python / launch cost of a step (~30 ms of host time) collapses to a graph launch.  Inputs are
    copied into static device buffers; the optimizer's small host-side table upload stays outside the graph."""

    def __init__(self, model, optimizer=None, temperature=0.05, device=None, use_graph=False):
        self.model = model
        self.optimizer = optimizer
        self.loss = M.NormSoftmaxLoss(temperature)
        self.device = device if device is not None else next(model.parameters()).device
        self.params = [p for p in model.parameters() if p.requires_grad]   # after the optimizer applied the freeze policy
        self.use_graph = use_graph
        self.fused_loss = os.environ.get("TVTS_FUSED_LOSS", "1") != "0"      # the single-launch loss kernel (Bg <= 256, E <= 1024)
        self.loss_scale = L.DEFAULT_LOSS_SCALE     # 1 for bf16 operands; fp16 operands: static scale, undone inside the AdamW kernel
        # TVTS_OVERLAP_ALLREDUCE=1: gradient all-reduce in buckets on a communication stream while the backward is still running (see
        # _arm_early_allreduce).  Default OFF = ONE all-reduce of the whole arena after the backward: measured on 8 B200s (round 2,
        # profiles/r2_scaling.md) the bucketed variant is not faster (32.77 vs 32.60 ms / step) -- the backward is a chain of persistent
        # GEMM grids that own every SM's registers and shared memory, so the NCCL kernels cannot co-reside and only move the wait around.
        self.overlap = os.environ.get("TVTS_OVERLAP_ALLREDUCE", "0") == "1"
        # TVTS_PIPELINED_ADAMW=1 (world > 1): all-reduce in buckets, AdamW on bucket i while bucket i + 1 is on the wire
        # (_pipelined_allreduce_adamw).  Default OFF: identical weights (tests/test_dist_cpu.py, tests/test_dist_gpu.py) but measured not
        # faster on B200s (round 2, profiles/r2/call28_*, call29_*): 31.7 vs 31.4 ms at 2 GPUs, 33.26 vs 33.12 ms at 8 -- the AdamW pass and
        # the NCCL kernels both live on HBM bandwidth, and five small collectives replace one.
        self.pipelined = os.environ.get("TVTS_PIPELINED_ADAMW", "0") == "1"
        self._ranges = None
        self._comm_stream = None
        self._graphs = {}
        self.launches_per_graph = 0
        self._copy_stream = None
        self._staged = None
        if use_graph and _world() > 1:
            self._warm_collectives()

    def _warm_collectives(self):
        """Every rank constructs its TrainStep at the same point of the program, so this is where the communicator, its channels and
        its scratch buffers come to life (an all-gather and an arena-sized all-reduce, the two collectives of the step) -- NOT in the
        pre-capture warm-up of _capture(): ranks meet new input signatures (trimmed caption lengths, v1 `padding=True`) at different
        steps, and a warm-up that really ran the step's collectives would pair them with another rank's NEXT step."""
        flat = getattr(self.optimizer, "flat", None)
        t = torch.zeros(8, dtype=torch.float32, device=self.device)
        out = torch.empty(8 * _world(), dtype=torch.float32, device=self.device)
        native_comm(self.device)               # TVTS_COMM=native: the C ABI's communicator is created here, collectively
        _all_gather(out, t)
        if flat is not None:
            average_flat(flat)                 # gradients are zeroed at the start of every step: averaging them here is harmless
        else:
            _average_tensor(t)
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)

    def close(self):
        """Destroy the captured graphs (their executables hold the NCCL kernels of the step; the process group must not be torn down
        underneath them -- round 1's multi-rank exit hang) and the staging state.  Call before dist.destroy_process_group()."""
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)
        for ent in self._graphs.values():
            ent.graph.reset()
            ent.static.clear()
            ent.l1 = ent.l2 = None
        self._graphs.clear()
        self._last_key = None
        self._staged = None
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)

    def to_device(self, data):
        out = {}
        for k, v in data.items():
            if isinstance(v, dict):
                out[k] = self.to_device(v)
            else:
                out[k] = v.to(self.device, non_blocking=True) if torch.is_tensor(v) else v
        return out

    def _body(self, data, optimizer_launch_only=False, skip_optimizer=False):
        flat = getattr(self.optimizer, "flat", None)
        if flat is not None:
            self.optimizer.zero_grad()              # one memset over the gradient arena
        else:
            for p in self.params:                   # optimizer.zero_grad() (torch 2.x: set_to_none)
                p.grad = None
        overlapped = self.overlap and flat is not None and _world() > 1
        if overlapped:
            self._arm_early_allreduce(flat)
        text_embeds, video_embeds, pred_order = self.model(data)
        Bg = video_embeds.shape[0] * _world()
        if self.fused_loss and video_embeds.is_cuda and E.fused_losses_supported(Bg, video_embeds.shape[1]):
            # one launch: all-gather (NCCL) -> sim_matrix + NormSoftmaxLoss + local-slice gradient + 2 * sort CE (csrc/loss_fused.cu)
            loss1, loss2 = E.fused_losses(video_embeds, text_embeds, pred_order, data.get("label"), self.loss.temperature,
                                          gather_for_fused_loss)
            total = loss1 + loss2 if pred_order is not None else loss1
        else:
            video_all, text_all = gather_embeddings(video_embeds, text_embeds)
            output = M.sim_matrix(video_all, text_all)                  # rows videos, cols texts (trainer.py:484)
            loss1 = self.loss(output)
            if pred_order is not None:
                loss2 = E.sort_ce(pred_order, data["label"], 2.0)       # trainer.py:487-492
                total = loss1 + loss2
            else:
                loss2 = torch.zeros((), device=loss1.device)
                total = loss1
        dynamic = getattr(self.optimizer, "dynamic_scale", False)
        if dynamic:
            total = total * self.optimizer.scale_tensor                 # device-resident dynamic loss scale (IEEE-half operand build)
        elif self.loss_scale != 1.0:
            total = total * self.loss_scale                             # static scale: optimizer-less / stock-optimizer runs of that build
        total.backward()
        if (self.pipelined and not overlapped and flat is not None and _world() > 1 and hasattr(self.optimizer, "launch_range")
                and not skip_optimizer and (dynamic or hasattr(self.optimizer, "launch"))):
            self._pipelined_allreduce_adamw(flat, dynamic, optimizer_launch_only)
            return loss1.detach(), loss2.detach()
        if overlapped:
            self._finish_bucketed_allreduce(flat)
        elif flat is not None:
            average_flat(flat)
        else:
            average_gradients(self.params)
        if self.optimizer is not None and not skip_optimizer:
            if dynamic:                              # finite check + un-scale + update + scale policy: all on the device
                if optimizer_launch_only:
                    self.optimizer.launch()
                else:
                    self.optimizer.step()
            elif self.loss_scale != 1.0:
                if optimizer_launch_only:
                    self.optimizer.launch(1.0 / self.loss_scale)
                elif hasattr(self.optimizer, "launch"):
                    self.optimizer.step(1.0 / self.loss_scale)          # the fused kernel un-scales the gradients itself
                else:                                                   # stock torch / transformers optimizer: un-scale in place first
                    for p in self.params:
                        if p.grad is not None:
                            p.grad.mul_(1.0 / self.loss_scale)
                    self.optimizer.step()
            elif optimizer_launch_only:
                self.optimizer.launch()
            else:
                self.optimizer.step()
        return loss1.detach(), loss2.detach()

    PIPELINE_BUCKETS = 4    # gradient all-reduce / AdamW pipeline depth (see _pipelined_allreduce_adamw)

    def _pipelined_allreduce_adamw(self, flat, dynamic, launch_only):
        """world > 1: the gradient arena is all-reduced in PIPELINE_BUCKETS contiguous buckets of whole chunks on a communication stream,
        and the fused AdamW updates bucket i on the compute stream while bucket i + 1 is still on the wire -- the optimizer pass (HBM-bound,
        0.7 ms at B/16) hides under the all-reduce (NVLink-bound) instead of following it.  Same sums, same update: identical weights to
        average_flat() + optimizer.launch() (tests/test_dist_cpu.py, tests/test_dist_gpu.py).
        Dynamic loss scale (IEEE-half build): the whole-step skip decision has to exist before the first bucket is updated, so the finite
        check runs on the LOCAL gradients right after the backward and its flag is MAX-reduced across the ranks (4 bytes) -- a rank with a
        non-finite gradient makes every rank skip; what this cannot see, a sum of finite fp32 gradients overflowing 3.4e38, does not occur
        under a loss scale <= 2^24."""
        opt = self.optimizer
        if not launch_only:
            opt.prepare()
        scale = 1.0 if dynamic else 1.0 / self.loss_scale
        if dynamic:
            opt.launch_check()
            flag = opt.scale_state[2:3]
            if dist.get_backend() == "nccl" or flat.device.type == "cpu":
                dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        nb = max(1, min(self.PIPELINE_BUCKETS, flat.n_chunks))
        cuts = [flat.n_chunks * i // nb for i in range(nb + 1)]
        comm = None
        if flat.device.type == "cuda":
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=self.device)
            comm = self._comm_stream
        if comm is None:
            for c0, c1 in zip(cuts[:-1], cuts[1:]):
                _average_tensor(flat.g[c0 * flat.chunk: c1 * flat.chunk])
                opt.launch_range(c0, c1, scale)
        else:
            cur = torch.cuda.current_stream()
            comm.wait_stream(cur)                   # the backward (and the finite flag) are complete on the compute stream
            done = []
            with torch.cuda.stream(comm):
                for c0, c1 in zip(cuts[:-1], cuts[1:]):
                    _average_tensor(flat.g[c0 * flat.chunk: c1 * flat.chunk])
                    ev = torch.cuda.Event()
                    ev.record(comm)
                    done.append(ev)
            for ev, c0, c1 in zip(done, cuts[:-1], cuts[1:]):
                cur.wait_event(ev)
                opt.launch_range(c0, c1, scale)
        opt.launch_finish()

    BUCKET_BLOCKS = 3       # video blocks per all-reduce bucket (12 blocks -> 4 buckets of ~85 MB at B/16)

    def _arm_early_allreduce(self, flat):
        """DDP-style bucketed overlap on the flat arena: as soon as a bucket's gradients are final (engine.GRAD_READY tags) its arena
        ranges are all-reduced on a communication stream while the backward continues; whatever has not been sent when the backward
        returns (ln_pre / patch-embed / embeddings, anything without a tag) goes last.  Buckets: every BUCKET_BLOCKS video blocks (the
        backward walks the blocks downwards, so a bucket is complete when its LOWEST block is), the sort head, the text tower.  The arena
        is laid out by optimizer group in model order, hence a bucket is at most one contiguous run per group.  Every rank issues the same
        collectives in the same order (the autograd engine walks the same graph), and the sums are the same sums as the single
        all-reduce: identical gradients (tests/test_dist_cpu.py, tests/test_dist_gpu.py)."""
        if self._ranges is None:
            names = {id(p): n for n, p in self.model.named_parameters()}
            buckets = {}

            def bucket_of(name):
                if name.startswith("text_"):
                    return ("text",)
                if name.startswith("pred_model."):
                    return ("sort",)
                key = "video_model.transformer.resblocks."
                if name.startswith(key):
                    i = int(name[len(key):].split(".")[0])
                    return ("video_block", i - i % self.BUCKET_BLOCKS)
                return ("rest",)
            for i, p in enumerate(flat.params):
                start = flat.offsets[i]
                end = flat.offsets[i + 1] if i + 1 < len(flat.params) else flat.total
                runs = buckets.setdefault(bucket_of(names.get(id(p), "")), [])
                if runs and runs[-1][1] == start:
                    runs[-1] = (runs[-1][0], end)
                else:
                    runs.append((start, end))
            self._ranges = buckets
            if self.device.type == "cuda":
                self._comm_stream = torch.cuda.Stream(device=self.device)
        self._sent = set()
        comm = self._comm_stream

        def hook(tag):
            if tag not in self._ranges or tag in self._sent:
                return
            self._sent.add(tag)
            if comm is None:
                average_ranges(flat, self._ranges[tag])
                return
            comm.wait_stream(torch.cuda.current_stream())       # this bucket's gradients are complete on the producing stream
            with torch.cuda.stream(comm):
                average_ranges(flat, self._ranges[tag])
        E.GRAD_READY = hook

    def _finish_bucketed_allreduce(self, flat):
        E.GRAD_READY = None
        comm = self._comm_stream
        rest = [r for tag, runs in self._ranges.items() if tag not in self._sent for r in runs]
        if comm is None:
            average_ranges(flat, rest)
            return
        cur = torch.cuda.current_stream()
        comm.wait_stream(cur)
        with torch.cuda.stream(comm):
            average_ranges(flat, rest)
        cur.wait_stream(comm)

    def _capture(self, key, data):
        if not (self.optimizer is None or hasattr(self.optimizer, "launch")):
            raise RuntimeError("TrainStep(use_graph=True) needs tvts_b200.optim.AdamW (or no optimizer)")
        static = {k: v.to(self.device, copy=True) for k, v in data.items() if torch.is_tensor(v)}      # `data` is flat here
        nested = _unflatten(static)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        global _COLLECTIVES_OFF
        with torch.cuda.stream(side):               # warm-up outside capture: allocator pools, bf16 weight casts, func attributes.
            _COLLECTIVES_OFF = True                 # Collective-free (see _warm_collectives): the captured collectives have the same
            try:                                    # sizes for every input signature, so ranks may capture at different steps
                self._body(nested, skip_optimizer=True)
            finally:
                _COLLECTIVES_OFF = False
        cur.wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        n0 = L.launch_count()
        # further input signatures (caption vs transcript batches, trimmed text lengths) capture into the FIRST graph's memory pool: the
        # graphs replay one at a time, so their multi-GB activation workspaces can overlap instead of adding up
        first = next(iter(self._graphs.values()), None)
        ctx = torch.cuda.graph(graph) if first is None else torch.cuda.graph(graph, pool=first.graph.pool())
        with ctx:
            l1, l2 = self._body(nested, optimizer_launch_only=True)
        self.launches_per_graph = L.launch_count() - n0
        flat = getattr(self.optimizer, "flat", None)
        ent = types.SimpleNamespace(graph=graph, static=static, l1=l1, l2=l2, active=flat.active_mask() if flat is not None else None)
        self._graphs[key] = ent
        self._last_key = key
        return ent

    def prefetch(self, data):
        """Start the host->device copy of the NEXT batch on a dedicated copy stream so that it overlaps the step that is running
        (what a pipelined data loader does); the following __call__(None) consumes it.  Pinned host tensors copy asynchronously."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._stage_bufs = {}
        data = _flatten(data)
        cs = self._copy_stream
        if self._staged is not None:
            cs.wait_event(self._staged[2])           # the previous consumer has finished reading the staging buffers
        with torch.cuda.stream(cs):
            staged = {}
            for k, v in data.items():
                if not torch.is_tensor(v):
                    continue
                buf = self._stage_bufs.get(k)
                if buf is None or buf.shape != v.shape or buf.dtype != v.dtype:
                    buf = torch.empty(v.shape, dtype=v.dtype, device=self.device)
                    self._stage_bufs[k] = buf
                buf.copy_(v, non_blocking=True)
                staged[k] = buf
            ready = torch.cuda.Event()
            ready.record(cs)
        self._staged = (staged, ready, torch.cuda.Event())

    def _take_staged(self):
        staged, ready, consumed = self._staged
        torch.cuda.current_stream().wait_event(ready)
        return staged, consumed

    def __call__(self, data=None):
        consumed = None
        if data is None:
            if self._staged is None:
                raise ValueError("TrainStep(None) needs a batch staged with prefetch()")
            data, consumed = self._take_staged()
        if not self.use_graph:
            out = self._body(self.to_device(_unflatten(data) if consumed is not None else data))
            if consumed is not None:
                consumed.record(torch.cuda.current_stream())
            return out
        data = _flatten(data)
        key = tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(data.items()) if torch.is_tensor(v))
        ent = self._graphs.get(key)
        if ent is None:
            ent = self._capture(key, data)
        if getattr(self, "_last_key", None) != key:
            # another graph (e.g. the caption-batch graph, which leaves the sort head without gradients) ran last: restore the
            # gradient-attachment pattern this graph was captured with, so that AdamW.prepare() skips / steps the right tensors
            flat = getattr(self.optimizer, "flat", None)
            if flat is not None and ent.active is not None:
                flat.set_active(ent.active)
            self._last_key = key
        flat = getattr(self.optimizer, "flat", None)
        if flat is not None:
            flat.refresh_bf16()                     # weights edited outside the optimizer (checkpoint resume): the graph reads the
        E.WEIGHTS.refresh()                         # bf16 operand copies directly, so bring them up to date first
        for k, v in ent.static.items():
            v.copy_(data[k], non_blocking=True)
        if consumed is not None:
            consumed.record(torch.cuda.current_stream())   # staging buffers are free again: the next H2D copy may start now
        if self.optimizer is not None:
            self.optimizer.prepare()                # host table -> device, stream-ordered before the replay
        ent.graph.replay()
        return ent.l1, ent.l2


def validate(model, batches, device=None, metrics=None):
    """The retrieval / order-accuracy core of Trainer_TVTSv2_*._valid_epoch (v2/trainer/trainer.py:527-635) for tokenised
    batches: forward without grad, all-gather of the embeddings (and of argmax(pred_order) / labels), then
    sim_matrix(text, video) (note the transposed order of validation, :605) -> t2v / v2t metrics and the exact-order accuracy.
    Returns {'t2v_metrics': {...}, 'v2t_metrics': {...}, 'order_acc': float | None}."""
    from . import metrics as MT
    metrics = metrics or [MT.t2v_metrics, MT.v2t_metrics]
    device = device if device is not None else next(model.parameters()).device
    W = _world()
    was_training = model.training
    model.eval()
    text_arr, vid_arr = [], []
    hits = total = 0

    def gather(t):
        if W == 1:
            return t
        out = torch.empty((W * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous())
        return out

    with torch.no_grad():
        for data in batches:
            data = {k: (v.to(device) if torch.is_tensor(v) else
                        ({kk: vv.to(device) for kk, vv in v.items()} if isinstance(v, dict) else v)) for k, v in data.items()}
            text_embed, vid_embed, preds = model(data, return_embeds=True)
            vid_arr.append(gather(vid_embed).cpu())
            text_arr.append(gather(text_embed).cpu())
            if preds is not None:
                labels = gather(data["label"].to(device)).cpu().numpy()
                order = gather(torch.argmax(preds, dim=-1)).cpu().numpy()
                hits += int((order == labels).all(axis=1).sum())
                total += order.shape[0]
    sims = M.sim_matrix(torch.cat(text_arr).to(device), torch.cat(vid_arr).to(device)).detach().cpu().numpy()
    if was_training:
        model.train()
    res = {m.__name__: m(sims) for m in metrics}
    res["order_acc"] = (hits / total) if total else None
    return res


class Trainer_TVTSv2:
    """Epoch loop of `Trainer_TVTSv2_{B_16,B_32,H_14}` (v2/trainer/trainer.py:361-525) on top of `TrainStep` / `validate`:
    same constructor signature, `train()`, `_train_epoch`, `_valid_epoch`, `_adjust_learning_rate`.

    What it keeps from the reference: the loader interleaving (the loader whose length equals `len_epoch` drives, the others are
    cycled, :440-461), clip-major flattening of the caption lists + tokenisation (:465-473), one optimizer step per loader batch,
    per-loader running losses, per-epoch x0.1 decay at `args.schedule` milestones (:402-411).  What it does differently: the step
    itself runs through `TrainStep` (CUDA graph, arena all-reduce, fused AdamW) instead of DDP + `optimizer.step()`, and losses are
    read back once per `log_step` instead of 3-4 `.item()` syncs per step.  Checkpoint files keep the reference's layout
    (`save_checkpoint` / `resume_checkpoint`, `config.resume`); monitoring / early stopping stay with the reference's
    `base.Multi_BaseTrainer_dist` (out of scope): pass `on_epoch_end` to hook them in."""

    def __init__(self, args, model, loss, metrics, optimizer, config, data_loader, valid_data_loader=None, lr_scheduler=None,
                 len_epoch=None, writer=None, visualizer=None, tokenizer=None, max_samples_per_epoch=50000, use_graph=True,
                 on_epoch_end=None, trim_text=True):
        self.args, self.model, self.loss, self.metrics, self.optimizer, self.config = args, model, loss, metrics, optimizer, config
        self.data_loader = list(data_loader)
        self.valid_data_loader = valid_data_loader
        self.do_validation = valid_data_loader is not None
        self.lr_scheduler, self.writer, self.visualizer, self.tokenizer = lr_scheduler, writer, visualizer, tokenizer
        self.max_samples_per_epoch = max_samples_per_epoch
        self.trim_text = trim_text                   # drop token columns that are padding in every caption of the batch (exact: causal text tower)
        if len_epoch is None:
            self.len_epoch = None
            for x in self.data_loader:
                if getattr(x, "dataset_name", "").startswith("YT"):
                    self.len_epoch = len(x)
            if self.len_epoch is None:
                self.len_epoch = len(self.data_loader[0])
        else:
            self.len_epoch = len_epoch
        self.batch_size = getattr(self.data_loader[0], "batch_size", 1)
        self.log_step = max(1, int(self.batch_size ** 0.5))
        self.n_gpu = getattr(args, "world_size", _world())
        self.allgather = AllGather_multi.apply
        self.num_clips = 4
        self.n_trans = 4
        try:                                        # dict or the reference's ConfigParser (only __getitem__ is relied upon)
            trainer_cfg = config["trainer"]
        except (KeyError, TypeError, IndexError):
            trainer_cfg = {}
        self.epochs = trainer_cfg.get("epochs", 1)
        self.start_epoch = 1
        self.init_val = trainer_cfg.get("init_val", True)              # base_trainer.py:37
        # base_trainer.py:19-25: the entry scripts hand over a CPU model whose optimizer already exists; move it to this rank's GPU (no
        # DDP wrap: TrainStep averages the gradient arena itself) and let the optimizer re-create its arenas next to the parameters
        try:
            n_gpu_cfg = config["n_gpu"]
        except (KeyError, TypeError, IndexError):
            n_gpu_cfg = 1
        if torch.cuda.is_available() and n_gpu_cfg > 0 and next(model.parameters()).device.type != "cuda":
            model.to(torch.device("cuda", getattr(args, "local_rank", 0) or 0))
        if hasattr(optimizer, "follow_parameters"):
            optimizer.follow_parameters()
        self.device = next(model.parameters()).device
        model.device = self.device
        # checkpoint period / best-model monitor (base_trainer.py:35-53,116-143)
        self.save_period = trainer_cfg.get("save_period", 0)
        self.monitor = trainer_cfg.get("monitor", "off")
        self.checkpoint_dir = getattr(config, "save_dir", None)
        if self.monitor == "off":
            self.mnt_mode, self.mnt_metric, self.mnt_best = "off", None, 0
        else:
            self.mnt_mode, self.mnt_metric = self.monitor.split()
            assert self.mnt_mode in ("min", "max")
            self.mnt_best = float("inf") if self.mnt_mode == "min" else -float("inf")
        temperature = getattr(loss, "temperature", 0.05)
        # whole-step CUDA graphs need the fused flat optimizer (its update is one capturable launch); with a stock torch / transformers
        # optimizer (the unmodified entry script on transformers==4.10.2) the step runs launch by launch
        graphable = self.device.type == "cuda" and (optimizer is None or hasattr(optimizer, "launch"))
        self.step = TrainStep(model, optimizer, temperature, self.device, use_graph=use_graph and graphable)
        self.on_epoch_end = on_epoch_end
        self.history = []
        resume = getattr(config, "resume", None)
        if resume is not None:                                   # base_trainer.py:60-61
            self.resume_checkpoint(resume)

    # ---- helpers ------------------------------------------------------------------------------------------------
    def _tokenize(self, data):
        """caption lists [clip][sample] -> flat clip-major list (index t*B + b) -> token tensor (:465-473)."""
        if self.tokenizer is not None and not torch.is_tensor(data["text"]):
            text_all = []
            for clip in data["text"]:
                text_all = text_all + list(clip)
            data = dict(data)
            data["text"] = self.tokenizer(text_all, truncate=True)
        if self.trim_text and torch.is_tensor(data["text"]) and data["text"].dim() == 2:
            data = dict(data)
            data["text"] = trim_text_context(data["text"])
        return data

    def _adjust_learning_rate(self, optimizer, epoch, args):
        lr_rate = 1.0
        for milestone in getattr(args, "schedule", []):
            if epoch == milestone:
                lr_rate = 0.1
        for group in optimizer.param_groups:
            group["lr"] = group["lr"] * lr_rate
        return lr_rate

    # ---- checkpoints (file format of v2/base/base_trainer.py:165-247) -----------------------------------------------------------
    def save_checkpoint(self, path, epoch, ddp_prefix=True):
        """Write what `_save_checkpoint` writes: {'arch', 'epoch', 'state_dict', 'optimizer', 'monitor_best', 'config'}.  The reference
        saves the DDP-wrapped model, so its keys carry a `module.` prefix; `ddp_prefix=True` reproduces that (both the reference's
        `_resume_checkpoint` and the model constructors' `load_checkpoint=` accept either form)."""
        sd = self.model.state_dict()
        if ddp_prefix and not next(iter(sd)).startswith("module."):
            sd = type(sd)(("module." + k, v) for k, v in sd.items())
        state = {"arch": type(self.model).__name__, "epoch": epoch, "state_dict": sd, "optimizer": self.optimizer.state_dict(),
                 "monitor_best": self.mnt_best, "config": self.config}
        torch.save(state, path)
        return path

    def resume_checkpoint(self, path):
        """`_resume_checkpoint` (:196-247): model weights (with or without the `module.` prefix), optimizer state when the optimizer
        type is unchanged, start epoch = saved epoch + 1."""
        from .compat import state_dict_data_parallel_fix
        checkpoint = torch.load(str(path), map_location=self.device, weights_only=False)
        self.start_epoch = checkpoint["epoch"] + 1
        self.mnt_best = checkpoint["monitor_best"]
        with torch.no_grad():
            # in-place copy: the parameters are views of the optimizer's flat arena and must stay so
            self.model.load_state_dict(state_dict_data_parallel_fix(checkpoint["state_dict"], self.model.state_dict()))
        saved_cfg, cfg = checkpoint.get("config"), self.config
        same_opt = True
        try:
            same_opt = saved_cfg["optimizer"]["type"] == cfg["optimizer"]["type"]
        except (KeyError, TypeError):
            pass
        if same_opt:
            self.optimizer.load_state_dict(checkpoint["optimizer"])
        else:
            print("Warning: Optimizer type given in config file is different from that of checkpoint. Optimizer parameters not being resumed.")
        return checkpoint

    # ---- epoch ----------------------------------------------------------------------------------------------------
    def _train_epoch(self, epoch):
        self.model.train()
        total_loss = [None] * len(self.data_loader)
        for loader in self.data_loader:
            sampler = getattr(loader, "train_sampler", None)
            if sampler is not None:
                sampler.set_epoch(epoch)
        iter_dl = [None] * len(self.data_loader)
        loop_dl, loop_dl_idx = None, None
        for dl_idx, dl in enumerate(self.data_loader):
            if loop_dl is None and len(dl) == self.len_epoch:
                loop_dl, loop_dl_idx = dl, dl_idx
            else:
                iter_dl[dl_idx] = iter(dl)
        if loop_dl is None:
            loop_dl, loop_dl_idx = self.data_loader[0], 0
            iter_dl[0] = None
        def batches():
            """(batch_idx, loader index, tokenised batch) in the reference's order: per iteration one batch of every loader (:440-475)."""
            for batch_idx, loop_dl_data in enumerate(loop_dl):
                data_li = [None] * len(self.data_loader)
                for dl_idx in range(len(iter_dl)):
                    if dl_idx != loop_dl_idx:
                        try:
                            data_li[dl_idx] = next(iter_dl[dl_idx])
                        except StopIteration:
                            iter_dl[dl_idx] = iter(self.data_loader[dl_idx])
                            data_li[dl_idx] = next(iter_dl[dl_idx])
                data_li[loop_dl_idx] = loop_dl_data
                for dl_idx, data in enumerate(data_li):
                    data = self._tokenize(data)
                    yield batch_idx, dl_idx, {k: data[k] for k in ("text", "video", "keep_ind", "label") if k in data}

        # On a GPU the loop runs one batch ahead: step i is enqueued (asynchronously: one CUDA-graph launch), THEN batch i+1 is fetched
        # from its loader, tokenised and its host->device copy started on the copy stream -- all of it under the kernels of step i.
        stream = batches()
        pipelined = self.device.type == "cuda"
        cur = next(stream, None)
        if pipelined and cur is not None:
            self.step.prefetch(cur[2])
        while cur is not None:
            batch_idx, dl_idx, batch = cur
            loss1, loss2 = self.step(None) if pipelined else self.step(batch)
            cur = next(stream, None)
            if pipelined and cur is not None:
                self.step.prefetch(cur[2])
            step_loss = (loss1 + loss2).reshape(())
            total_loss[dl_idx] = step_loss.clone() if total_loss[dl_idx] is None else total_loss[dl_idx] + step_loss
            if batch_idx % self.log_step == 0 and getattr(self.args, "local_rank", 0) == 0:
                print("Train Epoch: {} dl{} [{}/{}] Loss_ct: {:.6f} Loss_ce: {:.6f} Loss: {:.6f}".format(
                    epoch, dl_idx, batch_idx, self.len_epoch, loss1.item(), loss2.item(), step_loss.item()))
        log = {f"loss_{i}": (float(t.item()) / self.len_epoch if t is not None else 0.0) for i, t in enumerate(total_loss)}
        if self.do_validation:
            val_log = self._valid_epoch(epoch)
            if getattr(self.args, "rank", 0) == 0:
                log.update(val_log)
        self._adjust_learning_rate(self.optimizer, epoch, self.args)
        return log

    def _valid_epoch(self, epoch):
        out = {}
        for dl_idx, dl in enumerate(self.valid_data_loader):
            res = validate(self.model, (self._tokenize(d) for d in dl), device=self.device, metrics=self.metrics or None)
            for name, vals in res.items():
                if isinstance(vals, dict):
                    for k, v in vals.items():
                        out[f"val_{dl_idx}_{name}_{k}"] = v
            out[f"val_loss_{dl_idx}"] = res["order_acc"] if res["order_acc"] is not None else 1.0
        return out

    def train(self):
        try:
            return self._train()
        finally:
            self.step.close()         # captured graphs hold NCCL kernels: release them before the script destroys the process group

    def _train(self):
        if self.init_val and self.do_validation:
            self._valid_epoch(-1)
        for epoch in range(self.start_epoch, self.epochs + 1):
            log = {"epoch": epoch}
            log.update(self._train_epoch(epoch))
            self.history.append(log)
            if getattr(self.args, "rank", 0) == 0:
                for k, v in log.items():
                    print("    {:15s}: {}".format(str(k), v))
            best = False
            if self.mnt_mode != "off" and getattr(self.args, "rank", 0) == 0:
                if self.mnt_metric in log:
                    improved = log[self.mnt_metric] <= self.mnt_best if self.mnt_mode == "min" else log[self.mnt_metric] >= self.mnt_best
                    if improved:
                        self.mnt_best, best = log[self.mnt_metric], True
                else:
                    print("Warning: Metric '{}' is not found. Model performance monitoring is disabled.".format(self.mnt_metric))
                    self.mnt_mode = "off"
            if self.checkpoint_dir is not None and getattr(self.args, "rank", 0) == 0 and \
                    ((self.save_period and epoch % self.save_period == 0) or best):
                self._save_checkpoint(epoch, save_best=best)
            if self.on_epoch_end is not None:
                self.on_epoch_end(self, epoch, log)
        return self.history

    def _save_checkpoint(self, epoch, save_best=False):
        """base_trainer.py:165-194: checkpoint-epoch{N}.pth in config.save_dir (+ model_best.pth when the monitored metric improved)."""
        import shutil
        os.makedirs(str(self.checkpoint_dir), exist_ok=True)
        path = self.save_checkpoint(os.path.join(str(self.checkpoint_dir), "checkpoint-epoch{}.pth".format(epoch)), epoch)
        print("Saving checkpoint: {} ...".format(path))
        if save_best:
            shutil.copyfile(path, os.path.join(str(self.checkpoint_dir), "model_best.pth"))
            print("Saving current best: model_best.pth ...")
        return path


Trainer_TVTSv2_B_16 = Trainer_TVTSv2
Trainer_TVTSv2_B_32 = Trainer_TVTSv2
Trainer_TVTSv2_H_14 = Trainer_TVTSv2      # v2/trainer/trainer.py:648-905: the same step (no GradScaler; autocast lives in the model)


def verbose(epoch, metrics, mode, name="TEST"):
    """One-line retrieval report (v2/trainer/trainer.py:942-947; imported by the downstream zero_ret_* scripts)."""
    print(f"[{mode}]{name:s} epoch {epoch}, R@1: {metrics['R1']:.1f}, R@5: {metrics['R5']:.1f}, R@10: {metrics['R10']:.1f}, "
          f"R@50: {metrics['R50']:.1f}, MedR: {metrics['MedR']:g}, MeanR: {metrics['MeanR']:.1f}")


def format_nested_metrics_for_writer(metrics, mode, name="TEST"):
    """{key: value} -> {'[mode]name_key': value}   (v2/trainer/trainer.py:950-955)"""
    return {f"[{mode}]{name}_{key}": val for key, val in metrics.items()}


class Trainer_TVTS(Trainer_TVTSv2):
    """v1/trainer/trainer.py:40-260: the same epoch loop and step; the HuggingFace tokenizer is called with
    `return_tensors='pt', padding=True, truncation=True, max_length=50` (:130-131) and the model receives its
    {'input_ids', 'attention_mask'} dict.  Learning rate: base_lr x 0.1 per milestone already passed, set absolutely (:80-84)."""

    MAX_LENGTH = 50

    def __init__(self, *a, pad_to_max_length=False, **k):
        if k.get("len_epoch") is None and len(a) < 10:
            # v1/trainer/trainer.py:50-54: the LONGEST loader sets the epoch length (and therefore drives; the others are cycled)
            loaders = list(k["data_loader"] if "data_loader" in k else a[6])
            k["len_epoch"] = max(len(x) for x in loaders)
        super().__init__(*a, **k)
        self.pad_to_max_length = pad_to_max_length      # one sequence length -> one CUDA graph (padded keys are masked: same results)
        self.base_lr = self.optimizer.param_groups[0]["lr"]                        # v1/base/base_trainer.py:30

    def _tokenize(self, data):
        if self.tokenizer is not None and not isinstance(data["text"], dict):
            text_all = []
            for clip in data["text"]:
                text_all = text_all + list(clip)
            enc = self.tokenizer(text_all, return_tensors="pt", padding="max_length" if self.pad_to_max_length else True,
                                 truncation=True, max_length=self.MAX_LENGTH)
            data = dict(data)
            data["text"] = {"input_ids": enc["input_ids"], "attention_mask": enc["attention_mask"]}
        return data

    def _adjust_learning_rate(self, optimizer, epoch, args):
        lr = self.base_lr
        for milestone in getattr(args, "schedule", []):
            lr *= 0.1 if epoch >= milestone else 1.0
        for group in optimizer.param_groups:
            group["lr"] = lr
        return lr
