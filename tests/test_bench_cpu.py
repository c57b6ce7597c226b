"""bench.py's reference arm (the CPU oracle port timed on the host cores) must run without a GPU and print the contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample-batch", "1", "--workload", "c1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
              "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_flop_model_reproduces_survey_table():
    """bench.algorithmic_flops_per_pair == SURVEY.md section 8d table (fwd+bwd GFLOP per pair, modes A / B)."""
    sys.path.insert(0, ROOT)
    import bench
    from tvts_b200 import config as C
    table = {("c1", C.TVTSV2_B_32, 2): (82.9, 127.1), ("c2", C.TVTSV2_B_32, 8): (289.6, 346.7), ("c3", C.TVTSV2_B_16, 8): (563.7, 641.4),
             ("c4", C.TVTSV2_H_14.small(num_frames=16), 16): (6317.0, 6855.0)}
    for (name, cfg, T), (a, b) in table.items():
        tol = 0.15 if a < 1000 else 1.0
        assert abs(bench.algorithmic_flops_per_pair(cfg, T, 1) / 1e9 - a) < tol, name
        assert abs(bench.algorithmic_flops_per_pair(cfg, T, 4) / 1e9 - b) < tol, name


def test_bench_control_flow_dry_run():
    """bench.run_ours end to end on the CPU (torch restatements for the kernels, stand-ins for torch.cuda's device / stream / event / graph
    objects; tests/bench_dryrun.py, own process because it patches torch globally): every key of the contract is on the JSON line, the
    CUDA-graph path is the default, and the uint8 / trimmed-text input options shrink the per-step host->device bytes."""
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "bench_dryrun.py")], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [json.loads(l[len("DRYRUN "):]) for l in r.stdout.splitlines() if l.startswith("DRYRUN ")]
    assert len(lines) == 2 and all(not l["missing"] for l in lines)
    assert lines[0]["launch"].startswith("one CUDA graph") and lines[0]["dtype"] in ("bf16", "fp16") and lines[0]["value"] > 0
    assert lines[1]["h2d"] < lines[0]["h2d"] / 3 and "uint8" in lines[1]["input"] and "trimmed" in lines[1]["input"]


def test_bench_two_rank_dry_run_prints_one_line_and_exits_cleanly():
    """The multi-rank path of bench.py under torchrun (2 gloo ranks, same stand-ins): rank 0 prints ONE JSON line with the aggregate
    value, and every rank leaves through the barrier + os._exit(0) path (no destroy_process_group: round 1's NCCL teardown hang)."""
    import json
    import subprocess
    port = 29700 + (os.getpid() % 200)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "bench_dryrun.py")], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1 and lines[0]["n_gpus"] == 2 and lines[0]["config"]["parallelism"] == "dp2" and lines[0]["value"] > 0
    assert "grad_allreduce" in lines[0]["config"]["step"]


def test_v1_flop_model_reproduces_survey_row_c5():
    """tools/bench_v1.py: SURVEY section 8d row c5 -- 76.2 GF video + 4 x ~4.3 GF text + 12.2 GF sort head per pair forward, x3 for fwd+bwd."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_v1
    fwd = bench_v1.flops_per_pair() / 3e9
    assert abs(fwd - (76.2 + 4 * 4.3 + 12.2)) < 0.5, fwd
    b = bench_v1.make_batch(2)
    assert b["video"].shape == (2, 16, 3, 224, 224) and b["keep_ind"].shape == (2, 8, 49) and b["text"]["input_ids"].shape == (8, 50)
    assert (b["text"]["input_ids"] * (1 - b["text"]["attention_mask"])).abs().sum() == 0          # right padding with id 0
