"""The C-ABI library must load without a GPU and export every symbol include/tvts_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "tvts_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tvts_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from tvts_b200 import _lib
    lib = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 30, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.tvts_version() == 100


def test_argument_errors_are_reported_not_crashing():
    from tvts_b200 import _lib
    lib = _lib.lib()
    g = _lib.GemmArgs()
    assert lib.tvts_gemm(ctypes.byref(g), None) < 0
    assert b"empty problem" in lib.tvts_last_error()


def _err(lib):
    return lib.tvts_last_error().decode()


def test_entry_points_validate_arguments_before_touching_the_device():
    """Error convention of the C ABI (include/tvts_b200.h): negative return code + thread-local message, no launch, no crash --
    checked without a GPU because every entry point validates its arguments first."""
    from tvts_b200 import _lib
    lib = _lib.lib()
    i64, f32, vp = ctypes.c_int64, ctypes.c_float, ctypes.c_void_p
    dummy = ctypes.create_string_buffer(256)
    p = ctypes.cast(dummy, vp)
    # attention: head dims 64 (specialised kernels) and 80 (generic kernels, H/14) only
    rc = lib.tvts_attn_fwd(p, p, p, i64(1), i64(16), i64(2), i64(96), i64(0), i64(0), i64(0), i64(0), f32(0.1), None)
    assert rc < 0 and "head dim" in _err(lib)
    # divided modes need N == 1 + T*n
    rc = lib.tvts_attn_fwd(p, p, p, i64(1), i64(16), i64(2), i64(64), i64(1), i64(2), i64(5), i64(0), f32(0.1), None)
    assert rc < 0 and "1 + T*n" in _err(lib)
    # query window must lie inside the sequence
    rc = lib.tvts_attn_window_fwd(p, p, p, i64(1), i64(16), i64(2), i64(64), i64(14), i64(4), f32(0.1), None)
    assert rc < 0 and "attn_window_fwd" in _err(lib)
    # LayerNorm width must be a multiple of 128
    rc = lib.tvts_layernorm_fwd(p, p, p, p, i64(1), p, p, i64(4), i64(100), f32(1e-5), None)
    assert rc < 0 and "multiple of 128" in _err(lib)
    # GEMM: bf16 output needs N % 8 == 0; split-K needs accumulate
    g = _lib.GemmArgs()
    g.a = g.b = g.out = ctypes.addressof(dummy) & ~0xF
    g.M, g.N, g.K, g.lda, g.ldb, g.ldo, g.ldr, g.ldaux = 128, 12, 64, 64, 64, 16, 16, 16
    g.out_dtype = 1
    assert lib.tvts_gemm(ctypes.byref(g), None) < 0 and "multiples of 8" in _err(lib)
    g.N, g.out_dtype, g.splits = 16, 0, 4
    assert lib.tvts_gemm(ctypes.byref(g), None) < 0 and "split-K requires accumulate" in _err(lib)
    # optimizer: null arena
    rc = lib.tvts_adamw_flat(None, None, None, None, None, None, None, i64(4), i64(4096), f32(0.9), f32(0.999), f32(1e-6), f32(1.0), None)
    assert rc < 0 and "adamw_flat" in _err(lib)
    # communicator entry points: argument errors are reported, not crashed on; destroying nothing is fine (no NCCL call is made here)
    h = ctypes.c_void_p()
    assert lib.tvts_comm_init(ctypes.byref(h), None, i64(0), i64(2)) < 0 and "comm_init" in _err(lib)
    assert lib.tvts_comm_init(ctypes.byref(h), p, i64(3), i64(2)) < 0 and "rank 3 of 2" in _err(lib)
    assert lib.tvts_comm_allreduce(None, None, i64(4), i64(1), None) < 0 and "comm_allreduce" in _err(lib)
    assert lib.tvts_comm_allgather(None, p, p, i64(4), None) < 0 and "comm_allgather" in _err(lib)
    assert lib.tvts_comm_destroy(None) == 0
    # empty problems are no-ops, not errors
    assert lib.tvts_attn_fwd(p, p, p, i64(0), i64(16), i64(2), i64(64), i64(0), i64(0), i64(0), i64(0), f32(0.1), None) == 0
    assert lib.tvts_layernorm_fwd(p, p, p, p, i64(1), p, p, i64(0), i64(128), f32(1e-5), None) == 0


def test_fp16_operand_build_is_selected_by_env_and_checked():
    """TVTS_OPERAND=fp16 loads libtvts_b200_fp16.so (same C ABI, tvts_operand_format() == 1), switches the host-side operand dtype and
    the loss scale; a library of the other format under that name is refused."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import torch; from tvts_b200 import _lib as L, engine as E; lib = L.lib();"
            "assert lib.tvts_operand_format() == 1 and L.OPERAND_DTYPE is torch.float16 and E.BF16 is torch.float16;"
            "assert L.DEFAULT_LOSS_SCALE == 1024.0 and L.LIB_PATH.endswith('libtvts_b200_fp16.so');"
            "L._lib = None; L.LIB_PATH = L.LIB_PATH.replace('_fp16', '');\\n"
            "try:\\n    L.lib(); raise SystemExit('format mismatch was not detected')\\n"
            "except RuntimeError as e:\\n    assert 'operand format' in str(e)")
    subprocess.run([sys.executable, "-c", code.replace("\\n", "\n")], check=True, cwd=root, env=dict(os.environ, TVTS_OPERAND="fp16", PYTHONPATH=root))
    from tvts_b200 import _lib
    assert _lib.lib().tvts_operand_format() == int(_lib.OPERAND == "fp16")


def test_torch_restatements_take_the_arguments_of_the_c_entry_points():
    """tests/emu.py stands in for the library in every CPU host-logic test, so each restatement must take exactly the arguments of the
    C entry point it mirrors (minus the trailing stream): same count, checked against the prototypes in include/tvts_b200.h."""
    import inspect
    import emu
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "include", "tvts_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = dict(re.findall(r"\bint\s+tvts_(\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S))
    checked = 0
    for name, fn in emu.OPS.items():
        if name not in protos:
            continue
        params = [p.strip() for p in protos[name].split(",") if p.strip() and p.strip() != "void"]
        assert params and params[-1].replace(" ", "").startswith("void*stream"), (name, params[-1])
        arity = len([p for p in inspect.signature(fn).parameters.values() if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)])
        var = any(p.kind == p.VAR_POSITIONAL for p in inspect.signature(fn).parameters.values())
        assert var or arity == len(params) - 1, (name, arity, len(params) - 1)
        checked += 1
    assert checked >= 35, checked


def test_gemm_args_struct_layout_matches_the_header(tmp_path):
    """tvts_b200/_lib.GemmArgs (ctypes) must mirror `struct tvts_gemm_args` byte for byte: size and every field offset, as compiled by gcc."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        import pytest
        pytest.skip("needs gcc")
    from tvts_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fields = [f[0] for f in _lib.GemmArgs._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "tvts_b200.h"\nint main(void) {\n  printf("%zu\\n", sizeof(tvts_gemm_args));\n'
                   + "".join(f'  printf("%zu\\n", offsetof(tvts_gemm_args, {f}));\n' for f in fields) + "  return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert out[0] == ctypes.sizeof(_lib.GemmArgs)
    assert out[1:] == [getattr(_lib.GemmArgs, f).offset for f in fields]


def test_product_package_never_imports_the_oracle_or_the_test_restatements():
    """The oracle and tests/emu.py are test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import them."""
    import glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r"^\s*(from|import)\s+(emu|tvts_oracle|oracle|tests)\b", re.M)
    for path in glob.glob(os.path.join(root, "tvts_b200", "**", "*.py"), recursive=True):
        assert not pat.search(open(path).read()), path
    bench_src = open(os.path.join(root, "bench.py")).read()
    # the baseline legs only: the two CPU-oracle legs (v2, v1) and the stock-PyTorch-eager-on-the-GPU baseline; never the timed product path
    assert len(re.findall(r"import tvts_oracle", bench_src)) == 3
    for fn in ("def cpu_oracle_steps", "def v1_cpu_oracle_steps", "def eager_gpu_steps"):
        body = bench_src.split(fn)[1].split("\ndef ")[0]
        assert "import tvts_oracle" in body, fn
    assert "tvts_oracle" not in bench_src.split("def run_ours")[1].split("\ndef ")[0]
