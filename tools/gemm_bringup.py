"""GPU bring-up for the tcgen05 GEMM: each case runs in its own subprocess (a trap/hang in one variant
must not poison the rest).  Usage on the GPU box:  python tools/gemm_bringup.py [--perf]
Writes gpurun_out/gemm_bringup.log.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_case(spec):
    import torch
    from tvts_b200 import _lib as L
    torch.manual_seed(0)
    dev = "cuda"
    M, N, K = spec["M"], spec["N"], spec["K"]
    a_mn, b_mn = spec.get("a_mn", 0), spec.get("b_mn", 0)
    if "pair" in spec:
        L.lib().tvts_gemm_set_pair_mode(spec["pair"])
    if spec.get("epi"):
        L.lib().tvts_gemm_debug_epi(spec["epi"])
    if spec.get("dbg"):
        L.lib().tvts_gemm_debug_set(*spec["dbg"])
    A = torch.randn(M, K, device=dev) * 0.5
    B = torch.randn(N, K, device=dev) * 0.5
    Ab, Bb = A.bfloat16(), B.bfloat16()
    ref = Ab.float() @ Bb.float().t()
    a_store = Ab.t().contiguous() if a_mn else Ab
    b_store = Bb.t().contiguous() if b_mn else Bb
    lda = M if a_mn else K
    ldb = N if b_mn else K
    kw = {}
    mode = spec.get("mode", "plain")
    out_dtype = torch.bfloat16 if spec.get("bf16_out") else torch.float32
    out = torch.zeros(M, N, device=dev, dtype=out_dtype)
    if mode == "bias_act_res":
        bias = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev)
        pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        kw = dict(bias=bias, residual=res, act="quick_gelu")
        if out_dtype == torch.bfloat16:
            kw["out_pre"] = pre
        z = ref + bias
        ref_pre = z
        ref = z * torch.sigmoid(1.702 * z) + res
    elif mode == "dact":
        aux = (torch.randn(M, N, device=dev)).bfloat16()
        kw = dict(aux=aux, dact="gelu")
        x = aux.float()
        d = 0.5 * (1 + torch.erf(x / 2 ** 0.5)) + x * torch.exp(-0.5 * x * x) / (2 * 3.141592653589793) ** 0.5
        ref = ref * d
    elif mode == "splitk":
        out = torch.randn(M, N, device=dev)
        ref = ref + out
        kw = dict(accumulate=True, splits=spec.get("splits", 0))
    L.gemm(a_store, b_store, out, M=M, N=N, K=K, lda=lda, ldb=ldb, a_mn=a_mn, b_mn=b_mn, **kw)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    res = {"max_abs_err": err, "ref_max": scale, "rel": err / max(scale, 1e-9)}
    if mode == "bias_act_res" and "out_pre" in kw:
        res["pre_err"] = (kw["out_pre"].float() - ref_pre).abs().max().item()
    if spec.get("perf"):
        # time: rotate over enough distinct buffers to exceed L2 (126 MB)
        nbuf = max(2, int(300e6 // (M * K * 2 + M * N * out.element_size())) + 1)
        As = [a_store.clone() for _ in range(nbuf)]
        Os = [out.clone() for _ in range(nbuf)]
        for i in range(3):
            L.gemm(As[i % nbuf], b_store, Os[i % nbuf], M=M, N=N, K=K, lda=lda, ldb=ldb, a_mn=a_mn, b_mn=b_mn, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record()
        for i in range(iters):
            L.gemm(As[i % nbuf], b_store, Os[i % nbuf], M=M, N=N, K=K, lda=lda, ldb=ldb, a_mn=a_mn, b_mn=b_mn, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res["ms"] = ms
        res["tflops"] = 2.0 * M * N * K / ms / 1e9
        if not a_mn and not b_mn:
            Bt = Bb
            for i in range(3):
                torch.matmul(As[i % nbuf], Bt.t())
            torch.cuda.synchronize()
            e0.record()
            for i in range(iters):
                torch.matmul(As[i % nbuf], Bt.t())
            e1.record()
            torch.cuda.synchronize()
            res["cublas_tflops"] = 2.0 * M * N * K / (e0.elapsed_time(e1) / iters) / 1e9
    return res


CASES = [
    dict(name="kk_128x256x64", M=128, N=256, K=64),
    dict(name="kk_128x256x256", M=128, N=256, K=256),
    dict(name="kk_256x512x512", M=256, N=512, K=512),
    dict(name="kk_ragged", M=297, N=384, K=128),
    dict(name="kk_small_n", M=300, N=128, K=512),
    dict(name="kk_n4", M=64, N=4, K=128),
    dict(name="kk_bias_act_res", M=1000, N=768, K=768, mode="bias_act_res"),
    dict(name="kk_bf16_out", M=1000, N=2304, K=768, bf16_out=True),
    dict(name="kk_dact", M=515, N=512, K=2048, mode="dact", bf16_out=True),
    dict(name="kk_splitk", M=768, N=2304, K=8192, mode="splitk"),
    dict(name="mn_a", M=256, N=256, K=128, a_mn=1),
    dict(name="mn_b", M=256, N=256, K=128, b_mn=1),
    dict(name="mn_ab", M=256, N=512, K=256, a_mn=1, b_mn=1),
    dict(name="mn_ab_wgrad", M=768, N=2304, K=25152, a_mn=1, b_mn=1, mode="splitk"),
    dict(name="mn_ab_ragged", M=384, N=128, K=297 * 8, a_mn=1, b_mn=1, mode="splitk"),
    # alternates for the MN-major descriptor if the default is wrong: (lbo, sbo, kadv)
    dict(name="mn_ab_alt_swapped", M=256, N=512, K=256, a_mn=1, b_mn=1, dbg=(1024, 8192, 2048)),
    dict(name="mn_ab_alt_adv", M=256, N=512, K=256, a_mn=1, b_mn=1, dbg=(8192, 1024, 32)),
]
PAIR = [
    dict(name="pair_256x256x64", M=256, N=256, K=64, pair=1),
    dict(name="pair_512x512x256", M=512, N=512, K=256, pair=1),
    dict(name="pair_ragged", M=1000, N=384, K=128, pair=1),
    dict(name="pair_bias_act_res", M=1000, N=768, K=768, mode="bias_act_res", pair=1),
    dict(name="pair_dact_mnB", M=515, N=512, K=2048, b_mn=1, mode="dact", bf16_out=True, pair=1),
    dict(name="pair_mn_ab", M=512, N=512, K=256, a_mn=1, b_mn=1, pair=1),
    dict(name="pair_wgrad", M=768, N=2304, K=25152, a_mn=1, b_mn=1, mode="splitk", pair=1),
    dict(name="pair_perf_qkv", M=25152, N=2304, K=768, bf16_out=True, perf=True, pair=1),
    dict(name="solo_perf_qkv", M=25152, N=2304, K=768, bf16_out=True, perf=True, pair=0),
    dict(name="pair_perf_proj_res", M=25152, N=768, K=768, mode="bias_act_res", perf=True, pair=1),
    dict(name="pair_perf_fc", M=25152, N=3072, K=768, bf16_out=True, perf=True, pair=1),
    dict(name="pair_perf_cproj", M=25152, N=768, K=3072, perf=True, pair=1),
    dict(name="solo_perf_cproj", M=25152, N=768, K=3072, perf=True, pair=0),
    dict(name="pair_perf_wgrad", M=768, N=3072, K=25152, a_mn=1, b_mn=1, mode="splitk", perf=True, pair=1),
    dict(name="solo_perf_wgrad", M=768, N=3072, K=25152, a_mn=1, b_mn=1, mode="splitk", perf=True, pair=0),
    dict(name="pair_perf_8192", M=8192, N=8192, K=8192, bf16_out=True, perf=True, pair=1),
    dict(name="solo_perf_8192", M=8192, N=8192, K=8192, bf16_out=True, perf=True, pair=0),
]
EPI = [
    dict(name="epi0_qkv", M=25120, N=2304, K=768, bf16_out=True, perf=True, pair=1),
    dict(name="epi1_qkv_nostore", M=25120, N=2304, K=768, bf16_out=True, perf=True, epi=1, pair=1),
    dict(name="epi3_qkv_skip", M=25120, N=2304, K=768, bf16_out=True, perf=True, epi=3, pair=1),
    dict(name="epi0_qkv_f32", M=25120, N=2304, K=768, perf=True, pair=1),
    dict(name="epi0_text_qkv", M=9856, N=1536, K=512, bf16_out=True, perf=True, pair=1),
    dict(name="epi3_text_qkv_skip", M=9856, N=1536, K=512, bf16_out=True, perf=True, epi=3, pair=1),
    dict(name="epi0_text_out", M=9856, N=512, K=512, perf=True, pair=1),
    dict(name="epi3_text_out_skip", M=9856, N=512, K=512, perf=True, epi=3, pair=1),
    dict(name="epi0_k3072", M=25120, N=768, K=3072, perf=True, pair=1),
    dict(name="epi3_k3072_skip", M=25120, N=768, K=3072, perf=True, epi=3, pair=1),
    dict(name="epi3_8192_skip", M=8192, N=8192, K=8192, bf16_out=True, perf=True, epi=3, pair=1),
]
PERF = [
    dict(name="perf_qkv", M=25152, N=2304, K=768, bf16_out=True, perf=True),
    dict(name="perf_proj_res", M=25152, N=768, K=768, mode="bias_act_res", perf=True),
    dict(name="perf_fc", M=25152, N=3072, K=768, bf16_out=True, perf=True),
    dict(name="perf_cproj", M=25152, N=768, K=3072, perf=True),
    dict(name="perf_wgrad_mn", M=768, N=3072, K=25152, a_mn=1, b_mn=1, mode="splitk", perf=True),
    dict(name="perf_8192", M=8192, N=8192, K=8192, bf16_out=True, perf=True),
]


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        spec = json.loads(sys.argv[2])
        print("RESULT " + json.dumps(run_case(spec)))
        return
    cases = list(CASES)
    if "--perf" in sys.argv:
        cases += PERF
    if "--epi" in sys.argv:
        cases = list(EPI)
    if "--pair" in sys.argv:
        cases = list(PAIR)
    if "--one" in sys.argv:
        name = sys.argv[sys.argv.index("--one") + 1]
        spec = [c for c in CASES + PERF + EPI + PAIR if c["name"] == name][0]
        print(run_case(spec))
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "gemm_bringup.log"), "w")
    for spec in cases:
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", json.dumps(spec)],
                               capture_output=True, text=True, timeout=120)
            out = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            msg = out[-1][7:] if out else ("FAIL rc=%d %s" % (p.returncode, (p.stderr or "")[-300:].replace("\n", " | ")))
        except subprocess.TimeoutExpired:
            msg = "TIMEOUT"
        line = f"{spec['name']:24s} {msg}"
        print(line, flush=True)
        log.write(line + "\n")
        log.flush()


if __name__ == "__main__":
    main()
