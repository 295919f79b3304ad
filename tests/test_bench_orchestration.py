"""CPU dry run of bench.py's orchestration (main): the measurement functions are replaced by stubs, the control flow --
sweep over c1..c5, strong split, e2e last, failures reported instead of losing the line -- is the real one.  The GPU
measurement functions themselves run in `bench.py` on the B200 box (their records are committed under profiles/)."""
import json
import os
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


class _FakeTorch:
    class cuda:
        @staticmethod
        def empty_cache():
            pass

    @staticmethod
    def empty(*a, **k):
        return object()


class _FakeEnv:
    def __init__(self, args):
        self.torch, self.dist = _FakeTorch, None
        self.rank, self.world, self.local_rank, self.dev = 0, 1, 0, "cpu"
        self.opts, self.variant, self.ffma_tflops = {}, args.variant, None
        self.sdist = types.SimpleNamespace(shard_rows=lambda b, w, r: (0, b))
        lib = types.SimpleNamespace(scir_b200_microbench_ffma=lambda h, it, out: 0)
        self.L = types.SimpleNamespace(lib=lambda: lib)
        self.gpu = types.SimpleNamespace(torch_context=lambda t: types.SimpleNamespace(handle=None))

    any_rank = bench.Env.any_rank
    guarded = bench.Env.guarded

    def max_over_ranks(self, v):
        return float(v)


def _record(name, rows):
    return {"workload": name, "rows_per_gpu": rows, "steps": 3, "ms_per_step": 2.0 if rows > 600 else 1.0, "ms_by_rank": [1.0],
            "step_ms_rank0": {}, "value": 1.0, "unit": bench.UNIT, "clocks": {"sm_mhz": 1965.0, "reasons": []},
            "roofline": {"bound": "hbm", "achieved": 1.0, "peak": 2.0, "unit": "GB/s", "frac": 0.5, "kernel_ms": 1.0},
            "parity": {"frac": 0.1, "ok": True}, "gpu_launches": 6, "tensor_core_launches": 3, "fft_launches": 0,
            "arithmetic": "stub", "l2": "stub"}


def _run(monkeypatch, capsys, argv, fail_config=None, fail_e2e=False):
    calls = []

    def fake_measure_config(env, name, cfg, rows, steps, warmup, want_parity=True, keep=False):
        calls.append((name, rows))
        if name == fail_config:
            raise RuntimeError("boom in " + name)
        return _record(name, rows), ("kept" if keep else None)

    def fake_e2e(env, cfg, rows, steps, kept):
        calls.append(("e2e", rows))
        if fail_e2e:
            raise MemoryError("no pinned memory")
        return {"value": 5.0, "unit": bench.UNIT, "h2d_bytes_per_step": 1, "d2h_bytes_per_step": 1}

    monkeypatch.setattr(bench, "Env", _FakeEnv)
    monkeypatch.setattr(bench, "measure_config", fake_measure_config)
    monkeypatch.setattr(bench, "measure_e2e", fake_e2e)
    monkeypatch.setattr(bench, "cpu_reference_rate", lambda cfg, taps, threads, budget: (0.05 * threads, threads, 0.1))
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    bench.main()
    out = [ln for ln in capsys.readouterr().out.splitlines() if ln.strip()]
    assert len(out) == 1                                         # ONE JSON line
    return json.loads(out[0]), calls


def test_default_run_measures_every_config_and_e2e_last(monkeypatch, capsys):
    d, calls = _run(monkeypatch, capsys, ["--steps", "3", "--warmup", "3", "--strong-div", "2"])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "configs", "strong", "parity"):
        assert key in d, key
    assert sorted(d["configs"]) == ["c1", "c2", "c3", "c4", "c5"] and d["config"]["config"] == "c2"
    assert [c for c, _ in calls][0] == "c2" and calls[-1][0] == "e2e"            # headline first, host-array leg last
    assert ("c2", 512) in calls and ("c5", 4096) in calls                       # the strong split: rows / 2
    assert d["strong"]["c2"]["rows_per_gpu"] == 512 and d["strong"]["c2"]["efficiency_vs_n1"] == pytest.approx(1.0)
    assert d["cpu_baseline"]["cores"] == 1 and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert d["config"] == bench.config_dict("c2", bench.CONFIGS["c2"], 1024, 1)   # the dict the reference arm prints too


def test_a_failing_config_is_reported_not_fatal(monkeypatch, capsys):
    d, _ = _run(monkeypatch, capsys, ["--steps", "3", "--strong-div", "2"], fail_config="c5")
    assert "boom in c5" in d["configs"]["c5"]["error"] and "ms_per_step" in d["configs"]["c3"]
    assert "error" in d["strong"]["c5"] and "ms" in d["strong"]["c2"]
    assert d["value"] == 1.0                                     # the headline survives


def test_a_failing_e2e_leg_is_reported_not_fatal(monkeypatch, capsys):
    d, _ = _run(monkeypatch, capsys, ["--steps", "3", "--no-sweep", "--no-cpu"], fail_e2e=True)
    assert d["e2e"]["value"] is None and "no pinned memory" in d["e2e"]["error"]
    assert d["configs"] is None and d["strong"] is None
