#!/usr/bin/env python
"""bench.py -- full-tree log-likelihood evaluations per second on BASELINE config 2.

Workload (BASELINE.json configs[1]): 200-taxon synthetic DNA alignment with
1,000,000 distinct site patterns, GTR+G4, one tree.  A *step* is one
``pf.p4_treeLogLike(cTree, 0)``: the conditional-likelihood recursion over all
198 internal nodes plus the per-part reduction (Pf/p4_tree.c:868-922) -- the
T_like timer of BASELINE.md section 3, the figure the ">= 50x" bar is stated
against.  Patterns are sharded across ranks (one process per GPU, strong
scaling: the alignment is fixed); the only collective is the NCCL all-reduce of
the partial lnL.

  value     evals/s with everything resident in HBM (device-timed, max over ranks)
  e2e       evals/s of ``Tree.calcLogLike()`` through the pf mirror: Python glue,
            host Q/eigen/gamma, host->device parameter blocks, P(t) kernel, CL
            recursion, reduction, device->host result
  roofline  the whole-tree CL kernel: compulsory bytes of the fused design (every internal CL stored once, tips and
            counts read once) / its CUDA-event time, against the measured HBM peak; measured DRAM traffic of the same
            kernel at the same shard size from this round's ncu capture (profiles/r2_traffic.json); SURVEY.md 8(d)'s
            per-node byte model under its own name
  sustained the same step back to back for >= 1.5 s (clocks under the power cap), beside the K-step burst
  configs   at N = 1: BASELINE configs 1, 3, 4 (full evaluation + dirty path) and 5 (generations/s, through the
            call-protocol mirror and through the reference's REAL p4 Mcmc), each device-timed, with roofline,
            single-core reference on a column sample and site likelihoods checked against the reference
  cpu_baseline / --impl reference
            the reference's own Pf engine (oracle/_ref, built from its unmodified
            sources) on the host cores, on a bounded sample of the same alignment,
            scaled linearly in patterns (every hot loop is linear in patterns)

Usage: python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

CFG = 2
N_TAX = 200
N_PATTERNS = 1000000
METRIC = "full-tree lnL evals/sec"
UNIT = "evals/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--patterns", type=int, default=N_PATTERNS, help="override the pattern count (testing only)")
    ap.add_argument("--taxa", type=int, default=N_TAX)
    ap.add_argument("--cpu-sample", type=int, default=16384, help="patterns in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--per-node", action="store_true", help="one CL launch per node instead of the whole-tree kernel")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configs (1, 3, 4, 5) that follow the headline at N=1")
    ap.add_argument("--no-real-p4", action="store_true", help="skip the run of the reference's real p4 Mcmc on this engine (config 5)")
    ap.add_argument("--sustained-s", type=float, default=1.5, help="length of the sustained sub-measurement")
    return ap.parse_args()


def workload_name(a):
    return "cfg2: %d-taxon synthetic DNA, %d patterns, GTR+G4, full-tree lnL (p4_treeLogLike)" % (a.taxa, a.patterns)


# ------------------------------------------------------------------------------
# algorithmic bytes
# ------------------------------------------------------------------------------
def bench_config(a):
    """The ``config`` object of the JSON line: the workload only, the same in both arms (what differs by arm is under ``run``)."""
    return {"workload": workload_name(a), "taxa": a.taxa, "patterns": a.patterns, "model": "GTR+G4", "parts": 1,
            "l2": "inputs larger than L2: an evaluation streams the whole CL working set (8*4*4 bytes per internal node and pattern; "
                  "25.5 GB over all GPUs at the full size), no flush between steps"}


def algorithmic_bytes_per_pattern(tree):
    """Per pattern and per full-tree evaluation: for every internal node,
    8*dim*nCat*(1 + k_internal_children) + k_leaf_children bytes."""
    total = 0
    for pNum, mp in enumerate(tree.model.parts):
        unit = 8 * mp.dim * mp.nGammaCat
        for n in tree.iterInternalsPostOrder():
            kids = list(n.iterChildren())
            k_int = sum(1 for c in kids if not c.isLeaf)
            total += unit * (1 + k_int) + (len(kids) - k_int)
    return total


# ------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an NVML thread every 5 ms (the timed region of a
    sharded run is well under the 100 ms floor of an `nvidia-smi -lms` loop); nvidia-smi as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.thread = None
        self.stopFlag = False
        self.sm, self.mx, self.reasons = [], [], set()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.nvml = pynvml
            self._sample()                        # the first NVML queries of a process are slow: pay for them outside the timed region
            self.sm, self.mx, self.reasons = [], [], set()
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[index])
            except (ValueError, IndexError):
                pass
        return index

    def _sample(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)))
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20)):
            if r & bit:
                self.reasons.add(name)

    def _loop(self):
        while not self.stopFlag:
            try:
                self._sample()
            except Exception:
                break
            time.sleep(0.005)

    def start(self):
        if self.nvml is not None:
            import threading
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self.stopFlag = True
            self.thread.join(timeout=2)
            if not self.sm:
                try:
                    self._sample()
                except Exception:
                    pass
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "samples": len(self.sm), "reasons": sorted(self.reasons), "source": "nvml, 5 ms period"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 20"}


_FEED_SRC = r"""
import os, sys, time
fake = os.environ.get("P4B_BENCH_FAKE_CLOCKS")          # tests/test_bench_evidence.py: no GPU there
parent = os.getppid()
out = open(sys.argv[2], "w")
if not fake:
    import pynvml as n
    n.nvmlInit()
    h = n.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
    mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
while os.getppid() == parent:
    if fake:
        sm, mx, r = 1800, 1965, 4
    else:
        sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
    out.write("%.6f %d %d %d\n" % (time.time(), sm, mx, r))
    out.flush()
    time.sleep(0.005)
"""


class ClockFeed:
    """The same NVML samples taken by a PROCESS OF ITS OWN for the whole run, time-stamped, and cut to a timed region afterwards:
    the bench loop spends its time inside blocking engine calls, and a sampling thread of this process only runs between them."""

    def __init__(self, index):
        import tempfile
        self.proc, self.path = None, None
        try:
            fd, self.path = tempfile.mkstemp(prefix="p4b_clocks_", suffix=".txt")
            os.close(fd)
            self.proc = subprocess.Popen([sys.executable, "-c", _FEED_SRC, str(ClockSampler._physical_index(index)), self.path],
                                         stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def window(self, t0, t1):
        """Samples with t0 <= time.time() <= t1, as the ``clocks`` object; None if the feed has none there."""
        if self.proc is None:
            return None
        sm, mx, reasons = [], [], set()
        try:
            with open(self.path) as f:
                for line in f:
                    w = line.split()
                    if len(w) != 4 or not (t0 <= float(w[0]) <= t1):
                        continue
                    sm.append(float(w[1]))
                    mx.append(float(w[2]))
                    for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20)):
                        if int(w[3]) & bit:
                            reasons.add(name)
        except (OSError, ValueError):
            return None
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_min_mhz": min(sm), "sm_max_mhz": max(mx), "samples": len(sm), "reasons": sorted(reasons),
                "source": "nvml in a process of its own, 5 ms period, samples inside the timed region by time stamp"}

    def close(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.proc = None
        if self.path:
            try:
                os.unlink(self.path)
            except OSError:
                pass


def better_clocks(in_process, feed, t0, t1):
    """The feed's samples of the region when it has more of them than the in-process thread got."""
    w = feed.window(t0, t1) if feed is not None else None
    if w and w["samples"] > ((in_process or {}).get("samples") or 0):
        return w
    return in_process


# ------------------------------------------------------------------------------
# reference engine on the host cores
# ------------------------------------------------------------------------------
def _ref_worker(args):
    """One process: the reference Pf engine on a slice of the sample alignment."""
    taxa, sample, lo, hi, reps, warm = args
    import numpy as np
    import ref_loader
    import p4_phylogenetics_b200 as P
    rpf = ref_loader.load_ref_pf()
    tree = build_tree(P, rpf, taxa, sample, site_slice=(lo, hi), repeat=False)
    tree._commonCStuff()
    nPat = rpf.partPatternCount(tree.data.parts[0].cPart)
    for _ in range(max(1, warm)):
        lnL = rpf.p4_treeLogLike(tree.cTree, 0)   # warm-up
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        lnL = rpf.p4_treeLogLike(tree.cTree, 0)
        times.append(time.perf_counter() - t0)
    return nPat, times, lnL


def build_tree(P, pf, taxa, patterns, site_slice=None, repeat=True):
    """Config-2 tree/model/alignment; ``site_slice`` keeps a contiguous range of
    alignment columns (used to split the CPU sample over host processes)."""
    import numpy as np
    host, synth = P.host, P.synth
    rng = np.random.Generator(np.random.PCG64(20240 + CFG))
    tree = synth.random_tree(pf, taxa, rng)
    mp = synth.dna_model_part(0, rng, 4, pInvar=0.0)
    cache = os.environ.get("P4B_BENCH_CACHE")     # tuning aid: reuse the generated alignment between runs on one box
    key = None
    if cache:
        key = os.path.join(cache, "cfg2_%d_%d_%d.npy" % (taxa, patterns, int(repeat)))
    if key and os.path.exists(key):
        arr = np.load(key)
        aln = host.Alignment(pf, [arr[i].tobytes() for i in range(arr.shape[0])], host.DNA_SYMBOLS, host.DNA_EQUATES)
    else:
        aln = synth.make_alignment(pf, tree, mp, patterns, rng, "dna", repeat=repeat)
        if key:
            os.makedirs(cache, exist_ok=True)
            np.save(key, np.stack([np.frombuffer(x, dtype=np.uint8) for x in aln.sequences]))
    if site_slice is not None:
        lo, hi = site_slice
        aln = host.Alignment(pf, [s[lo:hi] for s in aln.sequences], aln.symbols, aln.equates)
    tree.attach(host.Data(pf, [aln]), host.Model(pf, [mp]))
    return tree


def time_reference(taxa, full_patterns, sample, nproc, reps, warm=1):
    """Reference Pf evals/s extrapolated to ``full_patterns``: the sample alignment
    is cut into ``nproc`` column ranges, one process each (Pf itself is
    single-threaded); a whole-sample evaluation takes as long as
    the slowest slice."""
    import multiprocessing as mp
    import numpy as np
    import p4_phylogenetics_b200 as P
    # length of the sample alignment (sites, not patterns), to cut it into ranges
    rng = np.random.Generator(np.random.PCG64(20240 + CFG))
    tree = P.synth.random_tree(None, taxa, rng)
    mpart = P.synth.dna_model_part(0, rng, 4, pInvar=0.0)
    aln = P.synth.make_alignment(None, tree, mpart, sample, rng, "dna", repeat=False)
    nSites = aln.length
    cuts = [(nSites * i) // nproc for i in range(nproc + 1)]
    jobs = [(taxa, sample, cuts[i], cuts[i + 1], reps, warm) for i in range(nproc)]
    if nproc == 1:
        res = [_ref_worker(jobs[0])]
    else:
        ctx = mp.get_context("fork")
        with ctx.Pool(nproc) as pool:
            res = pool.map(_ref_worker, jobs)
    nPat = sum(r[0] for r in res)                      # patterns actually evaluated (slices compress separately)
    per_rep = [max(r[1][k] for r in res) for k in range(reps)]
    t = sum(per_rep) / len(per_rep)                    # the mean over the timed steps, like the GPU arm
    evals_per_s_sample = 1.0 / t
    return evals_per_s_sample * (nPat / float(full_patterns)), nPat, t


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import ref_loader
    if not ref_loader.have_ref_pf():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/pf.so was not built (needs /root/reference at build time)"}))
        return 0
    nproc = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    nproc = max(1, min(nproc, 64))
    sample = a.cpu_sample * nproc
    sample = min(sample, a.patterns)
    reps = max(1, min(a.steps, 200))                   # exactly --steps timed steps (each ~0.3 s of CPU work), after --warmup untimed ones
    t0 = time.perf_counter()
    value, nPat, t = time_reference(a.taxa, a.patterns, sample, nproc, reps, max(1, a.warmup))
    wall = time.perf_counter() - t0
    sample_desc = ("%d patterns of the %d (%d column ranges, one process each); %.3f s per evaluation of the sample; "
                   "scaled by patterns" % (nPat, a.patterns, nproc, t))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": reps, "warmup": max(1, a.warmup),
        "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": bench_config(a),
        "run": {"engine": "reference Pf (Pf/*.c, gcc -O2) on host cores", "wall_s": round(wall, 1)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nproc, "kind": "reference", "sample": sample_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------
# byte / flop models, the reference on a column sample, the other BASELINE configs
# ------------------------------------------------------------------------------
def work_model(pf, tree, shard_of=None):
    """Per full-tree evaluation on this rank.
      compulsory  what the fused whole-tree design MUST move through HBM: every internal node's CL written once (the state
                  p4_treeLogLike leaves behind, as the reference does), every tip code and every pattern count read once.
                  Child CLs are not in it: a node's running CL stays in registers / shared memory / L2 between steps.
      algorithmic SURVEY.md 8(d): the same plus one read of every non-root internal CL (what a per-node design moves).
      flops       2*dim^2*nCat per internal-child edge and pattern."""
    comp = alg = flops = 0
    for pNum, mp in enumerate(tree.model.parts):
        if shard_of is not None:
            lo, hi = shard_of(pNum)
            nPat = hi - lo
        else:
            nPat = pf.partPatternCount(tree.data.parts[pNum].cPart)
        unit = 8 * mp.dim * mp.nGammaCat
        nLeaves = sum(1 for n in tree.nodes if n.isLeaf)
        for n in tree.iterInternalsPostOrder():
            kids = list(n.iterChildren())
            k_int = sum(1 for c in kids if not c.isLeaf)
            comp += unit * nPat
            alg += (unit * (1 + k_int) + (len(kids) - k_int)) * nPat
            flops += 2 * mp.dim * mp.dim * mp.nGammaCat * k_int * nPat
        comp += (nLeaves + 4) * nPat
    return float(comp), float(alg), float(flops)


def load_peaks():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


FP64_DMMA_TFLOPS = 37.0   # FP64 mma.sync m8n8k4, measured on this pool with tools/microbench/membw.cu (profiles/r2_membw.txt); FMA pipe 34.7


def measured_traffic(kernel_key, taxa, patterns_per_gpu):
    """DRAM bytes per launch of the shipped kernel at this exact shard size, from this round's ncu --set full captures
    (profiles/r2_traffic.json); None when no capture of this kernel at this size exists."""
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(tpath):
        return None, None
    for e in json.load(open(tpath)).get("entries", []):
        if e.get("kernel") == kernel_key and e.get("taxa") == taxa and e.get("patterns_per_gpu") == patterns_per_gpu:
            return e["dram_bytes_read"] + e["dram_bytes_write"], e.get("source")
    return None, None


def sample_twin(P, rpf, tree, nCols, seed):
    """The reference engine on nCols sampled columns of every alignment of `tree` (site likelihoods do not depend on the
    other columns, so the reference only has to evaluate the sample)."""
    import numpy as np
    H = P.host
    rng = np.random.default_rng(seed)
    alns, colsPer = [], []
    for aln in tree.data.alignments:
        cols = np.sort(rng.choice(aln.length, size=min(nCols, aln.length), replace=False))
        seqs = [np.frombuffer(x if isinstance(x, (bytes, bytearray)) else x.encode(), dtype=np.uint8)[cols].tobytes() for x in aln.sequences]
        alns.append(H.Alignment(rpf, seqs, aln.symbols, aln.equates))
        colsPer.append(cols)
    twin = tree.dupe()
    twin.pf = rpf
    twin.data = H.Data(rpf, alns)
    twin.model = H.clone_model(tree.model, rpf)
    return twin, colsPer


def reference_on_sample(P, pf, tree, nCols, seed=3, reps=3):
    """(max relative difference of the site likelihoods on the sample, reference evals/s of ONE core scaled to the full
    pattern count, description) -- the reference's own Pf engine (oracle/_ref)."""
    import numpy as np
    import ref_loader
    if not ref_loader.have_ref_pf():
        return None, None, "oracle/_ref not built"
    rpf = ref_loader.load_ref_pf()
    site = np.array(tree.getSiteLikes())
    twin, colsPer = sample_twin(P, rpf, tree, nCols, seed)
    want = np.array(twin.getSiteLikes())
    off = woff = 0
    worst = 0.0
    for aln, cols in zip(tree.data.alignments, colsPer):
        got = site[off + cols]
        w = want[woff:woff + len(cols)]
        worst = max(worst, float(np.max(np.abs(got - w) / w)))
        off += aln.length
        woff += len(cols)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        rpf.p4_treeLogLike(twin.cTree, 0)
        ts.append(time.perf_counter() - t0)
    t = min(ts)
    nSample = sum(rpf.partPatternCount(p.cPart) for p in twin.data.parts)
    nFull = sum(pf.partPatternCount(p.cPart) for p in tree.data.parts)
    value = 1.0 / (t * nFull / float(nSample))
    desc = "%d patterns of the %d (a column sample per part); %.3f s per evaluation of the sample; scaled by patterns" % (nSample, nFull, t)
    twin.deleteCStuff()
    twin.model.free()
    twin.data.free()
    return worst, value, desc


def free_tree(tree):
    tree.deleteCStuff()
    tree.model.free()
    tree.data.free()


CFG_NAMES = {
    1: "cfg1: 32-taxon synthetic DNA, 10k patterns, GTR+I+G4, Tree.calcLogLike",
    3: "cfg3: 100-taxon synthetic protein, 200k patterns, LG+G4 (20-state FP64 DMMA contraction)",
    4: "cfg4: 60-taxon protein NDCH2 (a composition per node), 4 parts x 50k patterns, full evaluation + dirty path",
    5: "cfg5: Mcmc.run, 8 Metropolis-coupled chains, 100-taxon DNA, 500k patterns, GTR+G4, generations/s",
}
CFG_SAMPLE = {1: 10000, 3: 1500, 4: 300, 5: 4096}


def config_block(P, pf, cfg, peak, steps, real_p4=True):
    """One BASELINE config that is not the headline: device-timed evaluations, roofline, single-core reference on a
    sample, site likelihoods checked against the reference.  Returns a dict for the bench line's `configs` block."""
    import numpy as np
    t0 = time.perf_counter()
    tree = P.synth.build_config(pf, cfg)
    setup = time.perf_counter() - t0
    lnL = tree.calcLogLike()
    for _ in range(3):
        pf.p4_treeLogLike(tree.cTree, 0)
    pf.treeTimerBegin(tree.cTree)
    for _ in range(steps):
        pf.p4_treeLogLike(tree.cTree, 0)
    ms = pf.treeTimerEnd(tree.cTree) / steps
    e0 = time.perf_counter()
    for _ in range(steps):
        tree.calcLogLike()
    e2e_ms = (time.perf_counter() - e0) * 1e3 / steps
    comp, alg, flops = work_model(pf, tree)
    dim = tree.model.parts[0].dim
    kernel = pf.lastCLKernelName()
    nPat0 = pf.partPatternCount(tree.data.parts[0].cPart)
    traffic, traffic_src = measured_traffic(kernel, sum(1 for n in tree.nodes if n.isLeaf), nPat0)
    if traffic and tree.model.nParts > 1:      # one capture = one launch = one part; the parts are the same size to a pattern
        traffic, traffic_src = traffic * tree.model.nParts, traffic_src + " x %d parts" % tree.model.nParts
    lean_ms = None
    if dim in (4, 20):      # lnL-only evaluations (p4b_setTreeStoresCL(0)): what the optimisers' objective runs
        pf.setTreeStoresCL(tree.cTree, 0)
        for _ in range(3):
            pf.p4_treeLogLike(tree.cTree, 0)
        pf.treeTimerBegin(tree.cTree)
        for _ in range(steps):
            pf.p4_treeLogLike(tree.cTree, 0)
        lean_ms = pf.treeTimerEnd(tree.cTree) / steps
        pf.setTreeStoresCL(tree.cTree, 1)
        tree.calcLogLike()
    out = {"workload": CFG_NAMES[cfg], "taxa": sum(1 for n in tree.nodes if n.isLeaf), "parts": tree.model.nParts,
           "patterns": [pf.partPatternCount(p.cPart) for p in tree.data.parts], "lnL": lnL, "setup_s": round(setup, 1),
           "ms_per_eval": ms, "evals_per_s": 1000.0 / ms, "e2e_calcLogLike_evals_per_s": 1000.0 / e2e_ms,
           "roofline": {"bound": "hbm", "achieved": comp / ms / 1e6, "peak": peak, "unit": "GB/s", "frac": comp / ms / 1e6 / peak,
                        "basis": "compulsory bytes: every internal CL stored once + tips + counts", "compulsory_GB": comp / 1e9,
                        "algorithmic_8d_GB": alg / 1e9, "algorithmic_8d_GBps": alg / ms / 1e6,
                        "kernel": kernel, "traffic": traffic, "traffic_source": traffic_src},
           "lnl_only_ms_per_eval": lean_ms}
    if dim == 20:
        out["roofline"]["tensor"] = {"GFLOP_per_eval": flops / 1e9, "TFLOPs": flops / ms / 1e9, "peak_TFLOPs": FP64_DMMA_TFLOPS,
                                     "frac": flops / ms / 1e9 / FP64_DMMA_TFLOPS, "peak_source": "FP64 mma.sync m8n8k4 microbenchmark (profiles/r2_membw.txt)"}
    worst, cpu_value, desc = reference_on_sample(P, pf, tree, CFG_SAMPLE[cfg])
    out["reference_check"] = {"site_likelihoods_max_rel_diff": worst, "tolerance": 1e-9, "ok": (worst is not None and worst <= 1e-9)}
    out["cpu_baseline"] = {"value": cpu_value, "unit": UNIT, "cores": 1, "kind": "reference", "sample": desc}
    if cfg == 4:
        # the dirty path: ONE composition changed, p4's whole-part protocol (p4/chain.py:305-380: p4_setPrams(pNum), every
        # internal node, p4_partLogLike); the engine finds by content what the change reaches (memoisation, its default)
        rng = np.random.default_rng(0)
        pf.setMemoize(1)
        tree.calcLogLike()
        leaves = [n for n in tree.nodes if n.isLeaf]
        internals = [n for n in tree.nodes if not n.isLeaf and n is not tree.root]
        res = {}
        for name, pool in (("one_leaf_composition", leaves), ("one_internal_composition", internals)):
            ts = []
            for k in range(steps):
                pNum = k % tree.model.nParts
                c = tree.model.parts[pNum].comps[pool[k % len(pool)].parts[pNum].compNum]
                c.val[:] = P.synth.normalise_comp(c.val * np.exp(rng.normal(0.0, 0.05, size=c.val.shape)))
                t1 = time.perf_counter()
                pf.p4_setPrams(tree.cTree, pNum)
                for n in tree.iterInternalsPostOrder():
                    pf.p4_setConditionalLikelihoodsOfInternalNodePart(n.cNode, pNum)
                v = pf.p4_partLogLike(tree.cTree, tree.data.parts[pNum].cPart, pNum, 0)
                ts.append((time.perf_counter() - t1) * 1e3)
            res[name + "_ms"] = sum(ts) / len(ts)
        dirty = sum(tree.partLikes)
        full = tree.calcLogLike()
        res["dirty_equals_full_rel_diff"] = abs(dirty - full) / abs(full)
        pf.setMemoize(0)
        out["dirty_path"] = res
    if cfg == 5:
        free_tree(tree)
        tree = P.synth.build_config(pf, 5)
        tree.bulkSetCStuff = True
        m = P.mcmc.Mcmc(tree, nChains=8, seed=1)
        pf.setMemoize(1)
        m.run(10, batched="pipelined")
        gens = 100
        n0 = pf.kernelLaunchCount()
        t1 = time.perf_counter()
        m.run(gens, batched="pipelined")
        pf.treeSync(m.chains[0].curTree.cTree)
        dt = time.perf_counter() - t1
        pf.setMemoize(0)
        out["mcmc_mirror"] = {"gens_per_s": gens / dt, "ms_per_gen": 1e3 * dt / gens, "chains": 8, "gens": gens,
                              "launches_per_gen": (pf.kernelLaunchCount() - n0) / gens,
                              "driver": "p4-phylogenetics_b200/mcmc.py (the call protocol of Chain.proposeSp / gen with nine proposals; pipelined chains)"}
        for c in m.chains:
            for t in (c.curTree, c.propTree):
                if t is not tree:
                    t.deleteCStuff()
                    t.model.free()
    free_tree(tree)
    if cfg == 5 and real_p4:
        out["mcmc_real_p4"] = real_p4_mcmc()
    return out


def codon61_block(P, pf, peak, nTax=32, nSites=60000, nCat=4):
    """north_star's "61-state codon" case (not a BASELINE config; the reference has no codon model: a 61-symbol 'standard'
    datatype through its generic-dim loop): the generic tensor-core whole-tree kernel (csrc/tree_dmma.cuh), device-timed,
    with its FP64 tensor throughput and the site likelihoods checked against the reference on a column sample."""
    t0 = time.perf_counter()
    tree = P.synth.build_generic(pf, P.synth.SYMBOLS_61, nTax, nSites, nCat, 6161, equates={"!": "abcd"})
    setup = time.perf_counter() - t0
    lnL = tree.calcLogLike()
    for _ in range(3):
        pf.p4_treeLogLike(tree.cTree, 0)
    steps = 10
    pf.treeTimerBegin(tree.cTree)
    for _ in range(steps):
        pf.p4_treeLogLike(tree.cTree, 0)
    ms = pf.treeTimerEnd(tree.cTree) / steps
    comp, alg, flops = work_model(pf, tree)
    out = {"workload": "61-state 'standard' data (codon-like), %d taxa, %d patterns, %d rate categories, full-tree lnL" %
                       (nTax, pf.partPatternCount(tree.data.parts[0].cPart), nCat),
           "kernel": pf.lastCLKernelName(), "lnL": lnL, "setup_s": round(setup, 1), "ms_per_eval": ms, "evals_per_s": 1000.0 / ms,
           "roofline": {"bound": "tensor", "achieved": flops / ms / 1e9, "peak": FP64_DMMA_TFLOPS, "unit": "TFLOP/s", "frac": flops / ms / 1e9 / FP64_DMMA_TFLOPS,
                        "GFLOP_per_eval": flops / 1e9, "peak_source": "FP64 mma.sync m8n8k4 microbenchmark (profiles/r2_membw.txt)",
                        "hbm_compulsory_GBps": comp / ms / 1e6, "hbm_frac": comp / ms / 1e6 / peak}}
    worst, cpu_value, desc = reference_on_sample(P, pf, tree, 400)
    out["reference_check"] = {"site_likelihoods_max_rel_diff": worst, "tolerance": 1e-9, "ok": (worst is not None and worst <= 1e-9)}
    out["cpu_baseline"] = {"value": cpu_value, "unit": UNIT, "cores": 1, "kind": "reference", "sample": desc}
    free_tree(tree)
    return out


def real_p4_mcmc(taxa=100, patterns=500000, chains=8, gens=160):
    """Config 5 through the reference's REAL p4 package (its own Mcmc.run / Chain code, staged under oracle/_ref/p4) with
    this repository's pf module as p4.pf -- a process of its own (tests/dropin/p4_like_side.py)."""
    import ref_loader
    if not ref_loader.have_ref_p4():
        return {"unavailable": "oracle/_ref/p4 not staged"}
    cmd = [sys.executable, os.path.join(ROOT, "tests", "dropin", "p4_like_side.py"), "mine", "--taxa", str(taxa), "--patterns", str(patterns),
           "--gens", str(gens), "--chains", str(chains), "--skip-opt", "--continue-gens", str(gens)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    except subprocess.TimeoutExpired:
        return {"unavailable": "timed out"}
    lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
    if r.returncode != 0 or not lines:
        return {"unavailable": "failed: " + (r.stderr or r.stdout)[-300:]}
    d = json.loads(lines[-1][len("RESULT"):])
    return {"gens_per_s": d.get("gens_per_s_continued", d["gens_per_s"]), "gens_per_s_first_run_with_chain_construction": d["gens_per_s"],
            "gens": gens, "chains": chains, "lnL0": d["lnL0"], "calcLogLike_s": d["calc_again_s"],
            "note": "gens_per_s: a second Mcmc.run(n) of the same object (the first call also builds the chains: two trees each, their device state)",
            "driver": "p4.Mcmc(t, nChains=8).run(n) -- p4/mcmc.py:2496, p4/chain.py proposals, unmodified -- on this pf module"}


# ------------------------------------------------------------------------------
# the B200 engine
# ------------------------------------------------------------------------------
def run_b200(a):
    import numpy as np
    import torch
    import p4_phylogenetics_b200 as P
    pf = P.pf
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        log("warning: WORLD_SIZE=%d but --gpus %d" % (world, a.gpus))
    if not torch.cuda.is_available() or pf.deviceCount() < 1:
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local)
    pf.setDevice(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        uid = [pf.commGetUniqueId() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        pf.commInitRank(uid[0], rank, world)     # also sets the pattern shard of this rank

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    feed = ClockFeed(local) if rank == 0 else None
    pf.setMemoize(0)      # every call does its full work: nothing in the timed regions is answered from an earlier result
    if a.per_node:
        pf.setFusedTreeKernel(0)
    t0 = time.perf_counter()
    tree = build_tree(P, pf, a.taxa, a.patterns)
    nPat = pf.partPatternCount(tree.data.parts[0].cPart)
    t_setup = time.perf_counter() - t0
    lnL0 = tree.calcLogLike()     # allocates device state, uploads the shard, first evaluation
    lo, hi = pf.treeShardRange(tree.cTree, 0)
    if rank == 0:
        log("setup %.1fs: %d taxa, %d sites, %d patterns (shard %d..%d), lnL %.6f, device bytes %.2f GB"
            % (t_setup, a.taxa, tree.data.parts[0].nChar, nPat, lo, hi, lnL0, pf.treeDeviceBytes(tree.cTree) / 1e9))
    bpp = algorithmic_bytes_per_pattern(tree)
    n_internal = sum(1 for _ in tree.iterInternalsPostOrder())

    # ---- value: resident, device-timed ------------------------------------------
    for _ in range(a.warmup):
        pf.p4_treeLogLike(tree.cTree, 0)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = pf.kernelLaunchCount()
    cl_ms, cl_launches = [], 0
    pf.treeTimerBegin(tree.cTree)
    w0 = time.perf_counter()
    c0 = time.time()
    for _ in range(a.steps):
        lnL = pf.p4_treeLogLike(tree.cTree, 0)
        ms, cl_launches = pf.treeLastCLTiming(tree.cTree)
        cl_ms.append(ms)
    dev_ms = pf.treeTimerEnd(tree.cTree)
    barrier()
    wall_ms = (time.perf_counter() - w0) * 1e3
    c1 = time.time()
    launches = pf.kernelLaunchCount() - launches0
    clocks = better_clocks(sampler.stop(), feed, c0, c1) if rank == 0 else None
    tmax = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max = [float(x) for x in tmax.tolist()]
    ms_per_step = dev_ms_max / a.steps
    value = 1000.0 / ms_per_step

    # ---- lnL-only evaluations (opt-in mode, reported beside the headline) -----------
    pf.setTreeStoresCL(tree.cTree, 0)
    for _ in range(a.warmup):
        pf.p4_treeLogLike(tree.cTree, 0)
    barrier()
    pf.treeTimerBegin(tree.cTree)
    for _ in range(a.steps):
        lnL_lean = pf.p4_treeLogLike(tree.cTree, 0)
    lean_ms = pf.treeTimerEnd(tree.cTree)
    barrier()
    pf.setTreeStoresCL(tree.cTree, 1)
    tl = torch.tensor([lean_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tl, op=dist.ReduceOp.MAX)
    lean_value = 1000.0 * a.steps / float(tl.item())

    # ---- sustained: the same step back to back for >= a.sustained_s seconds (clocks settle under the power cap) -----
    n_sus = max(a.steps, int(a.sustained_s * 1000.0 / ms_per_step) + 1)
    sampler2 = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler2.start()
    pf.treeTimerBegin(tree.cTree)
    c0 = time.time()
    for _ in range(n_sus):
        pf.p4_treeLogLike(tree.cTree, 0)
    sus_ms = pf.treeTimerEnd(tree.cTree)
    barrier()
    c1 = time.time()
    clocks_sus = better_clocks(sampler2.stop(), feed, c0, c1) if rank == 0 else None
    if feed is not None:
        feed.close()
    ts = torch.tensor([sus_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
    sustained = {"value": 1000.0 * n_sus / float(ts.item()), "unit": UNIT, "steps": n_sus, "seconds": float(ts.item()) / 1e3, "clocks": clocks_sus}

    # ---- e2e: Tree.calcLogLike() through the pf mirror ----------------------------
    rng = np.random.default_rng(1)
    for _ in range(2):
        tree.calcLogLike()
    barrier()
    e0 = time.perf_counter()
    for k in range(a.steps):
        # new branch lengths every step, so that the host really re-sends parameters
        n = tree.nodes[1 + (k % (len(tree.nodes) - 1))]
        n.br.len = float(np.clip(n.br.len * (1.0 + 0.01 * rng.standard_normal()), 1e-4, 0.5))
        lnL_e2e = tree.calcLogLike()
    barrier()
    e2e_ms = (time.perf_counter() - e0) * 1e3
    te = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = 1000.0 * a.steps / float(te.item())
    # ---- host-side pieces of one Tree.calcLogLike(), timed separately (SURVEY.md 8d timers i-iii) --------
    def avg_ms(fn, n=20):
        pf.treeSync(tree.cTree)
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        pf.treeSync(tree.cTree)
        return (time.perf_counter() - t0) * 1e3 / n
    host_ms = {"Tree.setCStuff (4 pf calls per node)": avg_ms(tree.setCStuff),
               "Model.setCStuff": avg_ms(tree.model.setCStuff),
               "pf.p4_setPrams (host Q/eigen/gamma + P(t) kernel)": avg_ms(lambda: pf.p4_setPrams(tree.cTree, -1)),
               "fast_bindings": bool(getattr(pf, "_hot", None))}
    n_nonroot = len(tree.nodes) - 1
    h2d = n_nonroot * (56 + 8 * 4) + (2 * 16 + 4) * 8      # P jobs + effective branch lengths + one eigensystem
    d2h = 16

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the CL kernel --------------------------------------------------
    # achieved = COMPULSORY bytes of the fused design / the kernel's CUDA-event time: every internal CL written once (the
    # state p4_treeLogLike leaves behind), tips and counts read once.  SURVEY.md 8(d)'s per-node figure, which also counts a
    # re-read of every internal child, is reported beside it under its own name: a whole-tree kernel keeps children in
    # registers / shared memory / L2, so that figure divided by the time would exceed the HBM peak without meaning anything.
    peak, peak_src = load_peaks()
    shard = hi - lo
    comp_bytes, alg_bytes, flops = work_model(pf, tree, shard_of=lambda pNum: pf.treeShardRange(tree.cTree, pNum))
    cl_ms_avg = sum(cl_ms) / len(cl_ms)
    achieved = comp_bytes / (cl_ms_avg * 1e-3) / 1e9
    kernel = pf.lastCLKernelName() if cl_launches == 1 else "cl_dna_kernel<4> (all %d CL launches of one evaluation)" % cl_launches
    traffic, traffic_src = measured_traffic(kernel, a.taxa, shard)
    roofline = {"bound": "hbm", "kernel": kernel + " (whole-tree CL recursion + site likelihoods + lnL fold, one launch)" if cl_launches == 1 else kernel,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "basis": "compulsory bytes per launch: 8*dim*nCat bytes stored per internal node and pattern + tips + counts",
                "compulsory_bytes_per_pattern": comp_bytes / shard,
                "traffic": traffic, "traffic_source": traffic_src,
                "traffic_GBps": (traffic / (cl_ms_avg * 1e-3) / 1e9) if traffic else None,
                "traffic_frac_of_peak": (traffic / (cl_ms_avg * 1e-3) / 1e9 / peak) if traffic else None,
                "algorithmic_8d": {"bytes_per_pattern": bpp, "GBps": float(bpp) * shard / (cl_ms_avg * 1e-3) / 1e9,
                                   "note": "SURVEY.md 8(d) per-node byte model (every child CL re-read from HBM); kept for comparison, not a roofline fraction"},
                "launches_per_eval": cl_launches, "avg_launch_us": 1e3 * cl_ms_avg / max(cl_launches, 1), "cl_ms_per_eval": cl_ms_avg}

    cpu = None
    if not a.no_cpu_baseline:
        import ref_loader
        if ref_loader.have_ref_pf():
            sample = min(a.cpu_sample, a.patterns)
            v, nP, t = time_reference(a.taxa, a.patterns, sample, 1, 3)
            cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "reference",
                   "sample": "%d patterns of the %d; %.3f s per evaluation of the sample; scaled by patterns" % (nP, a.patterns, t)}
        else:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": bench_config(a),
        "run": {"engine": "libp4b200.so on %d B200" % world, "patterns_per_gpu": shard, "internal_nodes": n_internal, "setup_s": round(t_setup, 1),
                "cl_working_set_GB_per_gpu": round(pf.treeDeviceBytes(tree.cTree) / 1e9, 2),
                "timing": "CUDA events on the engine stream, max over ranks"},
        "pattern_updates_per_s": value * n_internal * nPat,
        "lnL": lnL, "wall_ms_per_step": wall_ms_max / a.steps,
        "lnl_only": {"value": lean_value, "unit": UNIT, "lnL": lnL_lean,
                     "note": "opt-in p4b_setTreeStoresCL(0): only the CLs the evaluation re-reads are written; not the headline"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "call": "Tree.calcLogLike(): model.setCStuff + tree.setCStuff + pf.p4_setPrams + pf.p4_treeLogLike", "lnL": lnL_e2e,
                "host_ms": host_ms},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "clocks": clocks,
        "sustained": sustained,
        "host": "torch is imported for the rendezvous (torch.distributed), the timing barrier and the max over ranks only; the engine is libp4b200.so (nvcc, no torch)",
    }
    # ---- the other BASELINE configs (N = 1 only; parity-test cases, reported beside the headline) --------------------
    if world == 1 and not a.no_configs and a.patterns == N_PATTERNS and a.taxa == N_TAX:
        free_tree(tree)
        tree = None
        configs = {}
        for cfg in (1, 3, 4, 5):
            try:
                configs["cfg%d" % cfg] = config_block(P, pf, cfg, peak, 10, real_p4=not a.no_real_p4)
            except SystemExit as e:
                configs["cfg%d" % cfg] = {"failed": "engine error: %s" % (e,)}
            except Exception as e:      # the headline line must still be printed
                configs["cfg%d" % cfg] = {"failed": "%s: %s" % (type(e).__name__, e)}
            log("cfg%d: %s" % (cfg, json.dumps(configs["cfg%d" % cfg])[:400]))
        try:
            configs["codon61"] = codon61_block(P, pf, peak)
        except SystemExit as e:
            configs["codon61"] = {"failed": "engine error: %s" % (e,)}
        except Exception as e:
            configs["codon61"] = {"failed": "%s: %s" % (type(e).__name__, e)}
        log("codon61: %s" % json.dumps(configs["codon61"])[:400])
        line["configs"] = configs
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)
    return run_b200(a)


if __name__ == "__main__":
    sys.exit(main())
