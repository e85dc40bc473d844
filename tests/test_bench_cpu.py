"""CPU tests of bench.py's host-side bookkeeping (no GPU, no oracle run)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def test_both_arms_print_the_same_config():
    c = bench.CONFIGS["cfg5"]
    ours = bench.config_dict("cfg5", c, 8, c["batch"], c["frames"], "bf16")
    ref = bench.config_dict("cfg5", c, 8, c["batch"], c["frames"], "bf16")
    assert ours == ref and ours["workload"] == "cfg5" and ours["utterances_total"] == 8 * c["batch"] == 256
    assert ours["frames"] == 1292 and ours["utterances_per_gpu"] == 32


def test_frame_bytes_estimate_matches_the_library_figure():
    # the library's exact per-frame weight stream, as recorded by a GPU run (roofline.frame_bytes)
    rec = json.load(open(os.path.join(ROOT, "profiles", "r02e_final_bench_n1.json")))
    exact = rec["roofline"]["frame_bytes"]
    est = bench.frame_weight_bytes_estimate("1.5", "bf16")
    assert abs(est - exact) / exact < 5e-3
    assert bench.frame_weight_bytes_estimate("1.5", "f32") == 2 * est


def test_recorded_lines_carry_the_contract_keys():
    for name in ("r02e_final_bench_n1.json", "r02e_final_bench_n2.json"):
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks"):
            assert k in d, (name, k)
        assert d["config"]["workload"] == "cfg5" and d["gpu_launches"] > 0
        r = d["roofline"]
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["bound"] == "hbm"
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    ref = json.load(open(os.path.join(ROOT, "profiles", "r02e_final_bench_reference_arm.json")))
    assert ref["impl"] == "reference" and ref["cpu_baseline"]["kind"] == "port"
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_arm_under_torchrun_prints_one_line_from_rank_0():
    """`bench.py --impl reference` launched like the driver launches it for N > 1: rank 0 alone times the oracle on the
    host cores and prints ONE JSON line with the GPU arm's config; the other rank exits 0 without work."""
    import socket
    import subprocess
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-sample-frames", "2"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
    c = bench.CONFIGS["cfg5"]
    assert d["config"] == bench.config_dict("cfg5", c, 2, c["batch"], c["frames"], "bf16")
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
