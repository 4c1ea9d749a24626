"""The bench line's evidence files (CPU): bench.py looks `roofline.traffic` up in profiles/r2_traffic.json by the kernel name the
engine reports, the taxon count and the patterns per GPU -- a renamed kernel or a stale capture would silently turn it into null."""
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def test_traffic_captures_cover_every_shard_size_of_the_bench():
    bench = _load("bench_under_test", os.path.join(ROOT, "bench.py"))
    for pats in (1000000, 500000, 250000, 125000):          # 1, 2, 4, 8 GPUs of BASELINE configs[1]
        traffic, src = bench.measured_traffic("cl_tree_dna2_kernel<4,2,8,2>", 200, pats)
        assert traffic and src, pats
        # a whole-tree evaluation writes every internal CL once: 198 nodes x 16 rows x 8 bytes per pattern, and little more
        assert 0.99 * 198 * 128 * pats <= traffic <= 1.08 * 198 * 128 * pats, (pats, traffic)
    for taxa, pats, nodes in ((100, 199999, 98), (60, 50000, 58)):      # configs 3 and 4 (one part): 80 rows
        traffic, src = bench.measured_traffic("cl_tree_aa_kernel<8,8,2> x 4 categories", taxa, pats)
        assert traffic and src, (taxa, pats)
        assert 0.99 * nodes * 640 * pats <= traffic <= 1.08 * nodes * 640 * pats, (taxa, traffic)


def test_kernel_names_of_the_generator_match_the_engine():
    gen = _load("make_traffic_json", os.path.join(ROOT, "tools", "make_traffic_json.py"))
    assert gen.short_name("void p4b::cl_tree_dna2_kernel<(int)4, (int)2, (int)8, (int)2>(p4b::TreeArgs2)") == "cl_tree_dna2_kernel<4,2,8,2>"
    assert gen.short_name("void p4b::cl_tree_aa_kernel<(int)8, (int)8, (int)2>(p4b::TreeArgsAA)") == "cl_tree_aa_kernel<8,8,2>"
    d = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
    for e in d["entries"]:
        assert os.path.exists(os.path.join(ROOT, e["source"].split(" ")[0])), e["source"]


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (rank 0 only; the reference's own Pf engine on the host cores) prints ONE JSON line with the
    contract's keys, its `config` object equal to the one the B200 arm prints for the same arguments."""
    import subprocess
    import pytest
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_loader
    if not ref_loader.have_ref_pf():
        pytest.skip("oracle/_ref not built")
    args = ["--gpus", "1", "--steps", "2", "--warmup", "1", "--patterns", "4000", "--taxa", "12", "--cpu-sample", "256"]
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"] + args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1000.0) < 1e-6
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] == "reference"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    bench = _load("bench_under_test2", os.path.join(ROOT, "bench.py"))
    sys.argv, keep = ["bench.py"] + args, sys.argv
    try:
        assert d["config"] == bench.bench_config(bench.parse())
    finally:
        sys.argv = keep
    # under torchrun every other rank leaves without work and without a line
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"] + args, capture_output=True, text=True, timeout=60,
                       env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r.returncode == 0 and not r.stdout.strip()


def test_clock_feed_cuts_samples_to_the_timed_region(monkeypatch):
    """bench.py samples the SM clock and the throttle reasons from a process of its own and keeps the samples whose time stamps fall
    inside a timed region; without NVML (here) the feed yields nothing and the in-process sampler's result stands."""
    import time
    bench = _load("bench_under_test3", os.path.join(ROOT, "bench.py"))
    monkeypatch.setenv("P4B_BENCH_FAKE_CLOCKS", "1")
    feed = bench.ClockFeed(0)
    try:
        time.sleep(0.4)
        t0 = time.time()
        time.sleep(0.25)
        t1 = time.time()
        w = feed.window(t0, t1)
        assert w and 10 <= w["samples"] <= 60 and w["sm_mhz"] == 1800.0 and w["sm_max_mhz"] == 1965.0 and w["reasons"] == ["sw_power_cap"]
        assert feed.window(t1 + 100, t1 + 200) is None
        assert bench.better_clocks({"samples": 1, "sm_mhz": 1.0}, feed, t0, t1) == w
        assert bench.better_clocks({"samples": 10 ** 6}, feed, t0, t1) == {"samples": 10 ** 6}
    finally:
        pid = feed.proc.pid
        feed.close()
    assert not os.path.exists("/proc/%d" % pid)
    monkeypatch.delenv("P4B_BENCH_FAKE_CLOCKS")
    dead = bench.ClockFeed(0)            # no NVML device here: the child exits at once
    time.sleep(0.5)
    assert dead.window(0, time.time() + 1) is None and bench.better_clocks({"samples": 1}, dead, 0, time.time()) == {"samples": 1}
    dead.close()
