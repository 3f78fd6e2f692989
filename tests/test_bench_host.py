"""Host-side checks of bench.py that need no GPU: the synthetic generators are deterministic, the
reference arm prints one well-formed JSON line, every workload is declared consistently."""
import json
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_generators_are_deterministic_and_well_formed():
    w = dict(bench.WORKLOADS["cfg2_dna100m"], n=200_000)
    a, b = bench.gen_text_for(w), bench.gen_text_for(w)
    assert torch.equal(a, b) and int(a[-1]) == 0 and int(a[:-1].min()) == 1 and int(a[:-1].max()) == 4
    p, st = bench.gen_patterns(a, 1000, 32, 4, 4)
    assert torch.equal(p, bench.gen_patterns(a, 1000, 32, 4, 4)[0])
    t = a.numpy()
    for k in range(0, 1000, 2):                               # even patterns are substrings at `starts`
        assert np.array_equal(t[int(st[k]):int(st[k]) + 32], p[k].numpy())
    r = dict(bench.WORKLOADS["cfg3_rlfm_256m"], n=16 * 20_000)
    tr = bench.gen_text_for(r).numpy()
    base, other = tr[:20_000], tr[20_000:40_000]
    assert 0 < int((base != other).sum()) < 200               # ~0.1 % substitutions per copy
    mp = dict(bench.WORKLOADS["cfg4_multi_480m"], n=500_000)
    tm = bench.gen_text_for(mp).numpy()
    assert int((tm == 0).sum()) == 24 and tm[-1] == 0 and tm[0] != 0
    by = dict(bench.WORKLOADS["cfg5_bytes1g"], n=100_000)
    tb = bench.gen_text_for(by)
    flat, off = bench.gen_ragged_patterns(tb, 500, 255, 4)
    lens = np.diff(off.numpy())
    assert lens.min() >= 8 and lens.max() <= 64 and int(off[-1]) == flat.numel() and int(flat.min()) >= 1


def test_workload_table_is_consistent():
    for name, w in bench.WORKLOADS.items():
        assert w["kind"] in (0, 1, 2) and w["mc"] >= w["sigma"] >= 1 and w["level"] >= 0, name
        assert (w["m"] == 0) == (name.startswith("cfg5")), name


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1_dna1m",
                          "--steps", "2", "--warmup", "3"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["config"]["workload"] == "cfg1_dna1m"


def test_ncu_capture_is_refused_when_stale(tmp_path, monkeypatch):
    """profiles/ncu_traffic.json entries are used only for exactly these sources, this workload, batch and mode"""
    h = bench.source_hash()
    assert len(h) == 16 and h == bench.source_hash()
    prof = tmp_path / "profiles"
    prof.mkdir()
    entry = {"source_hash": h, "npat": 1000, "dram_bytes_per_step": 1.0, "l2_read_requests_per_step": 2.0}
    (prof / "ncu_traffic.json").write_text(json.dumps({"target_dna1g:rich": entry, "target_dna1g:compact": dict(entry, source_hash="0" * 16)}))
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    monkeypatch.setattr(bench, "source_hash", lambda: h)
    assert bench.ncu_capture("target_dna1g", 1000, "rich") == entry
    assert bench.ncu_capture("target_dna1g", 2000, "rich") is None          # another batch size: never scaled
    assert bench.ncu_capture("target_dna1g", 1000, "compact") is None       # taken from other sources
    assert bench.ncu_capture("cfg2_dna100m", 1000, "rich") is None


def test_rank_affinity_helper_is_harmless_without_gpus():
    before = os.sched_getaffinity(0)
    try:
        n = bench.pin_rank_to_numa(0, 1)
        assert n is None or n >= 1
    finally:
        os.sched_setaffinity(0, before)
