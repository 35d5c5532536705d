"""bench.py's reference arm and bookkeeping, runnable without a GPU: the JSON-line contract of `--impl reference`
(the reference's own CPU path = the oracle, the one place besides cpu_baseline where bench.py may execute it), the
kernel-source stamp that guards ncu-derived numbers, and the refusal of a stale roofline_traffic.json."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "patches/s" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["metric"] == "SUNet_B train patches/sec (256^2, bf16)" and d["config"]["global_batch"] == 128
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "4-patch" in cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_non_zero_rank_of_the_reference_arm_exits_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps",
                        "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_kernel_stamp_guards_traffic(tmp_path, monkeypatch):
    sys.path.insert(0, ROOT)
    import bench
    sha = bench.kernel_source_sha()
    assert len(sha) == 16 and sha == bench.kernel_source_sha()
    pdir = tmp_path / "profiles" / "r99"
    os.makedirs(pdir)
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    monkeypatch.setattr(bench, "kernel_source_sha", lambda: sha)
    (pdir / "roofline_traffic.json").write_text(json.dumps(
        {"dram_bytes_per_launch": 1.25e9, "kernel_sha": "0123456789abcdef", "source": "ncu launch list"}))
    val, why = bench.measured_traffic()
    assert val is None and "stale" in why
    (pdir / "roofline_traffic.json").write_text(json.dumps(
        {"dram_bytes_per_launch": 1.25e9, "kernel_sha": sha, "source": "ncu launch list"}))
    val, why = bench.measured_traffic()
    assert val == 1.25e9 and sha in why


def test_committed_traffic_is_from_this_build():
    """The ncu launch list under profiles/ that bench.py takes `roofline.traffic` from was captured on exactly the kernel
    sources in the tree (otherwise the bench line would say traffic: null — legal, but then the evidence is stale)."""
    sys.path.insert(0, ROOT)
    import bench
    val, why = bench.measured_traffic()
    assert val is not None and val > 1e8, why
    assert bench.kernel_source_sha() in why
