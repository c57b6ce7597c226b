"""ctypes binding of libtvts_b200.so (the C ABI declared in include/tvts_b200.h).

The library is built in-tree by build.sh (`__graft_entry__.build()`); there is NO fallback: if it is missing
or a call fails, a RuntimeError is raised (the product path never routes through torch ops or the oracle).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))

# 16-bit operand format (include/tvts_b200.h: tvts_operand_format).  TVTS_OPERAND=fp16 (DEFAULT since round 2) -> libtvts_b200_fp16.so:
# IEEE half, 11-bit mantissa -- the format that holds the north star's 1e-3 loss-trajectory bound (measured on the B200 over 100 steps
# of BASELINE.json configs[0]: 2.1e-4 / 4.2e-4 against 1.8e-3 / 3.5e-3 with bf16 operands, profiles/r2_loss_trajectory.md); the
# backward runs under a loss scale (initial value TVTS_LOSS_SCALE, default 1024) that the fused optimizer adapts on the device
# (optim.AdamW dynamic_scale: finite check + skipped step + GradScaler policy).  TVTS_OPERAND=bf16 -> libtvts_b200.so (no loss scale).
OPERAND = os.environ.get("TVTS_OPERAND", "fp16").lower()
if OPERAND not in ("bf16", "fp16"):
    raise ValueError(f"TVTS_OPERAND={OPERAND!r}: expected 'bf16' or 'fp16'")
OPERAND_DTYPE = torch.float16 if OPERAND == "fp16" else torch.bfloat16
DEFAULT_LOSS_SCALE = float(os.environ.get("TVTS_LOSS_SCALE", "1024" if OPERAND == "fp16" else "1"))
LIB_PATH = os.path.join(_HERE, "lib", "libtvts_b200_fp16.so" if OPERAND == "fp16" else "libtvts_b200.so")
if os.environ.get("TVTS_LIB_PATH"):       # A/B measurements against an earlier build of the same library on the same box
    LIB_PATH = os.path.abspath(os.environ["TVTS_LIB_PATH"])

_lib = None

c_void_p, c_int, c_i64, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("a", c_void_p), ("b", c_void_p), ("out", c_void_p), ("out_pre", c_void_p),
        ("bias", c_void_p), ("residual", c_void_p), ("aux", c_void_p),
        ("M", c_i64), ("N", c_i64), ("K", c_i64),
        ("lda", c_i64), ("ldb", c_i64), ("ldo", c_i64), ("ldr", c_i64), ("ldaux", c_i64),
        ("a_mn", ctypes.c_int32), ("b_mn", ctypes.c_int32), ("out_dtype", ctypes.c_int32),
        ("act", ctypes.c_int32), ("dact", ctypes.c_int32), ("accumulate", ctypes.c_int32),
        ("splits", ctypes.c_int32), ("alpha", c_float),
    ]


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"tvts_b200: native library {LIB_PATH} not found -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or ./build.sh).  There is no CPU / PyTorch fallback for the hot path.")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.tvts_last_error.restype = ctypes.c_char_p
        _lib.tvts_launch_count.restype = ctypes.c_longlong
        if int(_lib.tvts_operand_format()) != int(OPERAND == "fp16"):
            fmt, _lib = int(_lib.tvts_operand_format()), None
            raise RuntimeError(f"tvts_b200: {LIB_PATH} was built for operand format {fmt}, TVTS_OPERAND={OPERAND} -- rebuild")
    return _lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"tvts_b200.{what} failed ({rc}): {lib().tvts_last_error().decode()}")


def stream_ptr():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return c_void_p(t.data_ptr())


def launch_count():
    return int(lib().tvts_launch_count())


def call(name, *args):
    """Call `tvts_<name>(*args, stream)`; tensors are passed as device pointers, ints as int64, floats as float."""
    fn = getattr(lib(), "tvts_" + name)
    conv = []
    for a in args:
        if a is None:
            conv.append(c_void_p(None))
        elif isinstance(a, torch.Tensor):
            conv.append(c_void_p(a.data_ptr()))
        elif isinstance(a, bool):
            conv.append(c_i64(int(a)))
        elif isinstance(a, int):
            conv.append(c_i64(a))
        elif isinstance(a, float):
            conv.append(c_float(a))
        else:
            conv.append(a)
    conv.append(stream_ptr())
    check(fn(*conv), name)


ACT = {None: 0, "none": 0, "quick_gelu": 1, "gelu": 2}


def gemm(a, b, out, *, M, N, K, lda, ldb, ldo=None, a_mn=False, b_mn=False, bias=None, residual=None, ldr=None,
         aux=None, ldaux=None, out_pre=None, act=None, dact=None, accumulate=False, splits=0, alpha=1.0):
    """out[M,N] = epilogue(alpha * A.B^T); see include/tvts_b200.h.  All tensors must live on the current CUDA device."""
    g = GemmArgs()
    g.a, g.b, g.out = a.data_ptr(), b.data_ptr(), out.data_ptr()
    g.out_pre = out_pre.data_ptr() if out_pre is not None else None
    g.bias = bias.data_ptr() if bias is not None else None
    g.residual = residual.data_ptr() if residual is not None else None
    g.aux = aux.data_ptr() if aux is not None else None
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldb = lda, ldb
    g.ldo = ldo if ldo is not None else N
    g.ldr = ldr if ldr is not None else N
    g.ldaux = ldaux if ldaux is not None else N
    g.a_mn, g.b_mn = int(a_mn), int(b_mn)
    if out.dtype == torch.float32:
        g.out_dtype = 0
    elif out.dtype == OPERAND_DTYPE:
        g.out_dtype = 1
    else:
        raise TypeError(f"gemm output must be float32 or {OPERAND_DTYPE}")
    assert a.dtype == OPERAND_DTYPE and b.dtype == OPERAND_DTYPE
    g.act, g.dact = ACT[act], ACT[dact]
    g.accumulate, g.splits, g.alpha = int(accumulate), splits, alpha
    check(lib().tvts_gemm(ctypes.byref(g), stream_ptr()), "gemm")
    return out
