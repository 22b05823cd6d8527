"""bench.py --impl reference on the CPU: one JSON line, the contract's keys, the reference (or its restatement) as the
thing timed.  (The GPU arm needs a B200; its line is checked by the driver and committed under profiles/.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size-mb", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "MB/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"].startswith("encode MB/s") and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["bytes_per_gpu"] == 2000000


def test_gpu_arm_refuses_to_run_without_a_gpu():
    """no CPU fallback: without CUDA the GPU arm must fail loudly (on a GPU box this test is skipped)"""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("CUDA present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--size-mb", "1", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout) or "no CPU" in (r.stderr + r.stdout)


def test_traffic_figure_is_tied_to_the_sources(tmp_path, monkeypatch):
    """roofline.traffic comes from a committed ncu capture and is used only for the workload it was taken on and for the code it
    was taken from (sha of the library, or — the nvcc build is not bit-reproducible — sha of its sources)"""
    sys.path.insert(0, ROOT)
    import bench
    from libzling_b200 import build as zbuild
    src = zbuild.source_sha16()
    assert len(src) == 16 and src == zbuild.source_sha16()
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    monkeypatch.setattr(bench, "lib_sha16", lambda: "0" * 16)
    monkeypatch.setattr(bench, "src_sha16", lambda: src)
    os.makedirs(tmp_path / "profiles")
    rec = {"kernel": "k", "lib_sha16": "f" * 16, "src_sha16": src, "workload_bytes": 100, "level": 0, "dram_bytes_read": 7, "dram_bytes_write": 5, "source": "x"}
    (tmp_path / "profiles" / "traffic_k.json").write_text(json.dumps(rec))
    assert bench.load_traffic("k", 100, 0) == (12, "x")                       # same sources
    assert bench.load_traffic("k", 101, 0)[0] is None                         # another workload
    assert bench.load_traffic("k", 100, 4)[0] is None                         # another level
    monkeypatch.setattr(bench, "src_sha16", lambda: "1" * 16)
    assert bench.load_traffic("k", 100, 0)[0] is None                         # other sources
    monkeypatch.setattr(bench, "lib_sha16", lambda: "f" * 16)
    assert bench.load_traffic("k", 100, 0)[0] == 12                           # the very library of the capture
    assert bench.load_traffic("none", 100, 0) == (None, None)
