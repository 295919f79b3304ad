"""CPU-side checks of bench.py's contract: the reference arm (the CPU port of the reference loop; `--impl
reference`) runs without a GPU and prints ONE JSON line with the agreed keys, non-zero ranks under torchrun stay
silent, and the algorithmic per-output figures match SURVEY.md 8(d)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e,
                         timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    return res.stdout


def test_reference_arm_prints_one_json_line():
    out = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--config", "c1")
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "Gsamples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["vs_baseline"] is None


def test_reference_arm_other_ranks_are_silent():
    out = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--gpus", "2",
                    env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.strip() == ""


def test_algorithmic_figures_match_survey():
    sys.path.insert(0, ROOT)
    import bench
    # bytes/out and flop/out per config
    outs, nbytes, flops = bench.algorithmic(bench.CONFIGS["c2"], 1024)
    assert outs == 1024 * (1 << 20) and nbytes == outs * 8 and flops == outs * 126
    outs, nbytes, flops = bench.algorithmic(bench.CONFIGS["c3"], 256)
    assert flops == outs * 8194 and nbytes == outs * 8
    outs, nbytes, flops = bench.algorithmic(bench.CONFIGS["c4"], 2048)
    assert outs == 2048 * 1572864 and nbytes == 2048 * (1 << 20) * 4 + outs * 4 and flops == outs * 64
    outs, nbytes, flops = bench.algorithmic(bench.CONFIGS["c5"], 8192)
    assert flops == outs * 1020 and nbytes == outs * 8
    for name in ("c2", "c3", "c4", "c5"):
        assert bench.make_taps(bench.CONFIGS[name]).dtype.name == "float32"
