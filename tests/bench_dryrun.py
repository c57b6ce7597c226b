"""TEST HELPER (run in its own process by tests/test_bench_cpu.py): dry-run of bench.run_ours on the CPU -- kernels = the torch
restatements of tests/emu.py, torch.cuda's device / stream / event / graph objects = stand-ins.  Checks the script's control flow and the
JSON line's keys; the numbers it prints are meaningless."""
import contextlib, json, sys, types, io
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')]
import torch
import emu
emu.install()
from tvts_b200 import trainer as TR, _lib

real_device = torch.device
class FakeStream:
    cuda_stream = 0
    def wait_stream(self, s): pass
    def wait_event(self, e): pass
class FakeEvent:
    def __init__(self, **k): pass
    def record(self, s=None): pass
    def synchronize(self): pass
    def elapsed_time(self, other): return 12.5
class FakeGraph:
    def __init__(self): self.fn = None; self.outs = None
    def pool(self): return "pool"
    def reset(self): self.fn = None
    def replay(self):
        for d, s in zip(self.outs, self.fn()): d.copy_(s)
state = types.SimpleNamespace(capturing=None)
@contextlib.contextmanager
def graph_ctx(g, pool=None):
    state.capturing = g
    try: yield
    finally: state.capturing = None
orig_body = TR.TrainStep._body
def body(self, data, optimizer_launch_only=False, skip_optimizer=False):
    out = orig_body(self, data, optimizer_launch_only=optimizer_launch_only, skip_optimizer=skip_optimizer)
    if state.capturing is not None:
        g = state.capturing; g.fn = lambda: orig_body(self, data, optimizer_launch_only=True); g.outs = out
    return out
TR.TrainStep._body = body
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda *a: None
torch.cuda.current_stream = lambda *a, **k: FakeStream()
torch.cuda.Stream = lambda *a, **k: FakeStream()
torch.cuda.stream = lambda s: contextlib.nullcontext()
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.Event = FakeEvent
torch.cuda.CUDAGraph = FakeGraph
torch.cuda.graph = graph_ctx
torch.device = lambda *a, **k: real_device("cpu")
torch.Tensor.pin_memory = lambda self, *a, **k: self
lib = _lib.lib()
import torch.distributed as dist
_real_init = dist.init_process_group
dist.init_process_group = lambda backend=None, **k: _real_init("gloo")      # multi-rank dry run (torchrun): gloo in place of nccl
import bench
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    # multi-rank dry run (launched by torchrun): one configuration, printed straight to stdout -- run_ours destroys its captured graphs,
    # then the process group, and RETURNS (round 1 left through os._exit because the teardown hung under live graphs)
    sys.argv = ["bench.py", "--gpus", os.environ["WORLD_SIZE"], "--workload", "tiny", "--steps", "2", "--warmup", "1", "--no-cpu-baseline"]
    bench.run_ours(bench.parse())
    assert not dist.is_initialized(), "run_ours must leave the process group destroyed"
    raise SystemExit(0)
for extra in ([], ["--u8-input", "--trim-text"]) + ((["--no-graph"],) if "--all" in sys.argv[1:] else ()):
    sys.argv = ["bench.py", "--workload", "tiny", "--steps", "2", "--warmup", "1", "--no-cpu-baseline"] + extra
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        bench.run_ours(bench.parse())
    line = json.loads(buf.getvalue().strip().splitlines()[-1])
    need = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
            "roofline", "e2e", "gpu_launches", "clocks"}
    print("DRYRUN", json.dumps({"flags": extra, "missing": sorted(need - set(line)), "dtype": line["dtype"], "h2d": line["e2e"]["h2d_bytes_per_step"],
                                "launch": line["config"]["launch"], "input": line["config"]["input"], "value": line["value"]}))
