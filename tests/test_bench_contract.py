"""CPU (-m "not gpu"): the reference arm of bench.py runs without a GPU (it times the oracle port
on the host cores) -- check the one-JSON-line contract and its keys on a tiny grid."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--grid", "12",
                          "--steps", "1", "--warmup", "1", "--ref-iters", "2"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout                  # exactly one line on stdout
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "iterations/s" and d["value"] > 0
    assert d["dtype"] == "f64" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_our_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("CPU-only check")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--grid", "12", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0                           # no CPU fallback: the product path needs the GPU
    assert out.stdout.strip() == ""                      # and never prints a number it did not measure


def test_traffic_stamps_match_the_kernel_sources():
    """profiles/traffic.json holds ncu-measured DRAM bytes per launch; bench.py echoes an entry as
    `roofline.traffic` only while it is stamped with the fingerprint of the device headers its kernel
    class is compiled from.  A header edited after the capture (even a comment) makes the bench line
    report `traffic: null` -- this test says so before the round ends."""
    sys.path.insert(0, ROOT)
    import bench
    from new_cg_variants_b200 import build as b
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    assert "pr_fused" in t and "csr_sp_pr" in t
    for cls, ent in t.items():
        assert ent["kernel_sources_sha"] == b.kernel_fingerprint(cls), f"{cls}: captured for other kernel sources"
        assert all(os.path.exists(os.path.join(b.CSRC, f)) for p, lst in b.KERNEL_SOURCES.items() if cls.startswith(p) for f in lst)
    val, note = bench.read_traffic("pr_fused")
    assert val == t["pr_fused"]["dram_bytes_per_launch"] and "ncu" in note
    assert bench.read_traffic("no_such_kernel")[0] is None
    # a class's stamp ignores headers its kernel is not compiled from
    assert b.kernel_fingerprint("pr_fused") != b.kernel_fingerprint("csr_sp_pr") != b.kernel_fingerprint()
