#!/usr/bin/env python
"""bench.py -- benchmark of the B200 tree-scoring engine (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # CPU arm: oracle/_ref + oracle port, all cores

A "step" is one full evaluation of the hot path on one batch of synthetic input. The ONE JSON line
rank 0 prints carries the headline workload at the top level and every other BASELINE
configuration under "workloads" (each with value / ms_per_step / roofline / cpu_baseline / e2e /
check), all measured in the same process, one after the other:
  dna     (headline; BASELINE config 3) DNA GTR+G4 full-tree pruning + root lnL, 256 taxa x 4M
          site patterns, sharded contiguously over the ranks (total fixed => "strong").
  fitch   (BASELINE config 2) Fitch down-pass length, 64 taxa x 1M bit-packed DNA characters.
  fitch64 the same tree over 64 Mi characters (the bandwidth regime).
  aa / codon (BASELINE configs 4 / 5) 20-state +G4 and 61-state pruning (fp64 tensor cores); codon
          also runs config 5's branch-length re-evaluation loop.
  cfg1    (BASELINE config 1) DNA GTR+G4, 16 taxa x 10k sites: the latency regime.
`--workload X` makes X the headline; `--workloads a,b|all|none` picks the secondary set (default: all
when the headline is dna, none otherwise). PyTorch is used for events, NCCL and nothing else.

L2 rule: a working set below 4 x the 126 MB L2 (configs 1 and 2, lnL-only mode on small shards)
gets 512 MB written between timed iterations, outside each iteration's own event pair; larger ones
are timed with one event pair around all K steps. `config.l2` says which applied.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GTR_CO = [1.0, 2.5, 0.8, 1.2, 3.0]
GTR_PI = [0.30, 0.20, 0.25, 0.25]
BASE_PATTERNS = 65536  # evolved once, tiled to the workload size
FITCH_BASE = 1 << 20   # random characters generated once, tiled to the workload size

WORKLOADS = {
    # kind: "lk" (pruning) | "fitch" | "compress"; cpu_sample = patterns / characters of the CPU leg
    "dna": dict(kind="lk", T=256, N=4_000_000, S=4, K=4, cpu_sample=409_600, cpu_sample_1core=32_768,
                name="DNA GTR+G4 full-tree pruning, 256 taxa x 4M site patterns"),
    "aa": dict(kind="lk", T=128, N=500_000, S=20, K=4, cpu_sample=16_384, cpu_sample_1core=2_048,
               name="AA 20-state +G4 pruning, 128 taxa x 500k patterns"),
    "codon": dict(kind="lk", T=64, N=200_000, S=61, K=1, cpu_sample=8_192, cpu_sample_1core=1_024,
                  name="Codon 61-state pruning, 64 taxa x 200k patterns"),
    "cfg1": dict(kind="lk", T=16, N=10_000, S=4, K=4, cpu_sample=10_000, cpu_sample_1core=10_000,
                 name="DNA GTR+G4 log-likelihood, 16 taxa x 10k sites"),
    "fitch": dict(kind="fitch", T=64, N=1_000_000, S=4, K=1, cpu_sample=1_000_000, cpu_sample_1core=1_000_000,
                  name="Fitch/non-additive parsimony length, 64 taxa x 1M bit-packed DNA characters"),
    "fitch64": dict(kind="fitch", T=64, N=64 * FITCH_BASE, S=4, K=1, cpu_sample=FITCH_BASE, cpu_sample_1core=FITCH_BASE,
                    name="Fitch/non-additive parsimony length, 64 taxa x 64Mi bit-packed DNA characters"),
    "compress": dict(kind="compress", T=256, N=4_000_000, S=4, K=4, cpu_sample=500_000,
                     name="site-pattern compression (the step before the path), 256 taxa x 4M raw DNA sites"),
}
SECONDARY = ["fitch", "fitch64", "aa", "codon", "cfg1"]


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_model(wl, diagonalize=None):
    """The workload's MlModel.t record. diagonalize: the eigen-solver (default: the product's own
    phylo_diagonalize_*; the CPU arm passes the reference's, so that it never loads the product)."""
    from phylocaml_b200 import mlmodel

    if wl["S"] == 4:
        return mlmodel.create(("GTR", GTR_CO), 4, pi=GTR_PI, site_var=("gamma", wl["K"], 0.5), diagonalize=diagonalize)
    if wl["S"] == 20:
        R, pi = mlmodel.synthetic_reversible(20, 4)
        return mlmodel.create(("Const", R), 20, pi=pi, site_var=("gamma", wl["K"], 0.5), diagonalize=diagonalize)
    R, pi = mlmodel.gy94(2.0, 0.5, 5)
    return mlmodel.create(("Const", R), 61, pi=pi, diagonalize=diagonalize)


def mask_dtype(S):
    return np.uint8 if S <= 8 else (np.uint32 if S <= 32 else np.uint64)


def shard_bounds(N, world, rank, align=1024):
    blocks = (N + align - 1) // align
    per, rem = divmod(blocks, world)
    lo_b = rank * per + min(rank, rem)
    hi_b = lo_b + per + (1 if rank < rem else 0)
    return min(lo_b * align, N), min(hi_b * align, N)


def lk_bytes(T, S, K, tip_bytes, mode="pernode"):
    """Algorithmic bytes per pattern per launch. Per-node kernels: SURVEY.md 8(d) streaming
    model. Tree-fused kernel: compulsory traffic -- every interior CLV (+ scale counter)
    written once and every tip cell read once when CLVs are retained; tips only otherwise."""
    C = S * K * 8
    d = dict(prune_inner_inner=3 * C + 12, prune_tip_inner=2 * C + tip_bytes + 8,
             prune_tip_tip=C + 2 * tip_bytes + 4, root_lnl=2 * C + 8,
             tree=(2 * T - 3) * C + T * tip_bytes + (2 * T - 3) * 4)
    d["tree_fused"] = (T - 2) * (C + 4) + T * tip_bytes if mode == "fused" else T * tip_bytes
    return d


# fp64 tensor-core (mma.sync.m8n8k4.f64) peak of this pool's B200s measured by tools/peaks.cu
# (k_dmma; profiles/README.md): the denominator for the 20/61-state contraction kernels.
# MEASURED_PEAKS.json only carries the bf16 figure, which no fp64 kernel can be held against.
FP64_DMMA_TFLOPS = 37.1
L2_BYTES = 126 << 20      # B200 L2
FLUSH_BYTES = 512 << 20   # written between timed iterations when a working set could stay in L2


def device_footprint(workload, T, S, K, tip_bytes, mode, n_local):
    """Bytes one evaluation touches in HBM on this rank (decides the L2 rule): Fitch = every node's
    bit-sliced set (0.5 B per character for DNA); likelihood = tips (nibble-packed for 4 states) plus
    every interior CLV + scale counter when they are kept."""
    if workload == "fitch":
        return (2 * T - 1) * 0.5 * n_local
    tips = T * (0.5 if S == 4 else tip_bytes) * n_local
    return tips if mode == "fused-lnl" else tips + (T - 2) * (S * K * 8 + 4) * n_local


def parse_cpulist(text):
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11] (sysfs local_cpulist format)."""
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.extend(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_local_cpus(torch, local, sysfs="/sys/bus/pci/devices", min_cpus=8):
    """Multi-rank runs only: keep this rank's main thread (which allocates and first-touches the
    pinned tip buffer of the end-to-end leg) on the CPUs the kernel lists as local to its GPU, so
    that 8 ranks do not all upload across the socket interconnect. Best effort: returns a short
    description for the JSON line, or None when anything needed is missing."""
    try:
        prop = torch.cuda.get_device_properties(local)
        dom, bus, dev = getattr(prop, "pci_domain_id", 0), getattr(prop, "pci_bus_id", None), getattr(prop, "pci_device_id", None)
        if bus is None or dev is None:
            return None
        path = os.path.join(sysfs, "%04x:%02x:%02x.0" % (dom, bus, dev))
        with open(os.path.join(path, "local_cpulist")) as f:
            cpus = set(parse_cpulist(f.read()))
        allowed = cpus & set(os.sched_getaffinity(0))
        if len(allowed) < min_cpus or allowed == set(os.sched_getaffinity(0)):
            return None  # nothing to gain, or too few CPUs for this rank's helper threads
        os.sched_setaffinity(0, allowed)
        node = "?"
        try:
            with open(os.path.join(path, "numa_node")) as f:
                node = f.read().strip()
        except OSError:
            pass
        return "main thread bound to the %d CPUs local to the GPU (NUMA node %s)" % (len(allowed), node)
    except Exception:  # noqa: BLE001 -- never let placement tuning break a measurement
        return None


def host_placement(torch, local, world, sysfs="/sys/bus/pci/devices"):
    """What is true about where this rank's host side runs, for the JSON line: the CPUs and NUMA
    node the kernel lists as local to the GPU, whether the main thread was bound to them
    (bind_to_gpu_local_cpus: multi-rank runs only, and only when that is a proper subset of the
    CPUs this process may use), and the CPU count the ranks share."""
    info = {"cpus_visible": len(os.sched_getaffinity(0)), "ranks_sharing_them": world, "bound": False}
    try:
        prop = torch.cuda.get_device_properties(local)
        path = os.path.join(sysfs, "%04x:%02x:%02x.0" % (getattr(prop, "pci_domain_id", 0), prop.pci_bus_id, prop.pci_device_id))
        with open(os.path.join(path, "local_cpulist")) as f:
            info["gpu_local_cpulist"] = f.read().strip()
        with open(os.path.join(path, "numa_node")) as f:
            info["gpu_numa_node"] = int(f.read().strip())
    except Exception:  # noqa: BLE001
        info["gpu_local_cpulist"] = None
    note = bind_to_gpu_local_cpus(torch, local, sysfs) if world > 1 else None
    if note:
        info["bound"], info["note"] = True, note
    elif world > 1:
        info["note"] = ("not bound: the GPU-local CPU list is every CPU this process may use (all GPUs of the box hang "
                        "off one NUMA node), so %d ranks share %d CPUs for their upload threads" % (world, info["cpus_visible"]))
    return info


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(dev), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def count(self):
        """samples written so far"""
        if self.p is None:
            return 1 << 30
        try:
            with open(self.f.name) as g:
                return sum(1 for _ in g)
        except OSError:
            return 0

    def stop(self):
        if self.p is None:
            return None
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        self.f.close()
        os.unlink(self.f.name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "samples": len(sm),
                "reasons": sorted(reasons)}


def build_tips(tree_mod, tr, model, T, n_local, S, lo=0):
    """Evolve BASE_PATTERNS realistic patterns down the tree once; the global alignment is that
    block tiled to the workload size, and a rank holds columns [lo, lo + n_local) of it (so every
    rank count scores the same alignment), in page-locked host memory (the e2e leg uploads from
    there)."""
    from phylocaml_b200 import engine

    base_n = BASE_PATTERNS
    base = tree_mod.evolve_tips(tr, model, base_n, seed=3, dtype=mask_dtype(S))
    tips = engine.pinned_empty((T, n_local), mask_dtype(S))
    pos = 0
    while pos < n_local:
        off = (lo + pos) % base_n
        n = min(base_n - off, n_local - pos)
        tips[:, pos:pos + n] = base[:, off:off + n]
        pos += n
    return tips


def tile_cols(base, lo, n, out=None, unit=1):
    """Columns [lo, lo + n) of `base` repeated for ever (both in units of `unit` base columns per
    array column: packed layouts hold several alignment columns per element)."""
    bn = base.shape[1]
    if out is None:
        out = np.empty((base.shape[0], n), dtype=base.dtype)
    pos = 0
    while pos < n:
        off = (lo + pos) % bn
        k = min(bn - off, n - pos)
        out[:, pos:pos + k] = base[:, off:off + k]
        pos += k
    return out


def fitch_tips(tree_mod, T, n_total, lo, n_local, pinned=True):
    """The global Fitch alignment is one random block of FITCH_BASE characters tiled to n_total; a
    rank holds columns [lo, lo + n_local) of it (one character per byte, the reference's W = 8
    layout, lib/bitvector/bv.h:29-55)."""
    from phylocaml_b200 import engine

    base = tree_mod.random_fitch_chars(T, min(n_total, FITCH_BASE), 4, seed=5)
    tips = engine.pinned_empty((T, n_local), np.uint8) if pinned else np.empty((T, n_local), np.uint8)
    pos = 0
    while pos < n_local:
        off = (lo + pos) % base.shape[1]
        n = min(base.shape[1] - off, n_local - pos)
        tips[:, pos:pos + n] = base[:, off:off + n]
        pos += n
    return tips, base


def cpu_legs(wl, kind, ops, ra, rb, rt, n_nodes, model, sample, budget_s=8.0):
    """CPU baseline legs on `sample` (the first patterns / characters of rank 0's shard), timed on
    this box's host cores: the reference's own C where it exists (Fitch: bv_fitch + bv_distance of
    lib/bitvector/bv.c:46-55,148-160 out of oracle/_ref, kind "reference"), otherwise the oracle
    port (pruning: the reference has no pruning loop, lib/likelihood_c.ml:1-33; kind "port").
    Returns (headline leg = all threads at the reference's -O2, variants, last result)."""
    from oracle.oracle import Oracle, Ref

    cores = os.cpu_count() or 1
    T = wl["T"]
    unit = "char-ops/s" if kind == "fitch" else "site-updates/s"
    use_ref = kind == "fitch" and Ref.available()

    def runner(variant, nthreads, smp):
        if use_ref:
            ref = Ref(variant)
            return lambda: {"length": ref.fitch_score_tree(smp, ops, n_nodes, ra, rb, nthreads=nthreads)}
        orc = Oracle(variant)
        if kind == "fitch":
            return lambda: orc.fitch_score_tree(smp, None, ops, n_nodes, ra, rb, nthreads=nthreads)
        return lambda: orc.lk_score_tree(model, smp, None, ops, n_nodes, ra, rb, rt, nthreads=nthreads)

    def best_of(fn, budget):
        best, res, t_all = None, None, time.perf_counter()
        for _ in range(3):
            t0 = time.perf_counter()
            res = fn()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            if time.perf_counter() - t_all > budget:
                break
        return best, res

    ns = sample.shape[1]
    what = "characters" if kind == "fitch" else "patterns"
    code = ("reference bv_fitch per node + bv_distance (lib/bitvector/bv.c via oracle/_ref, W=8)" if use_ref
            else "oracle C port (oracle/phylo_oracle.c)")
    best, res = best_of(runner("o2", cores, sample), budget_s)
    main = {"value": (T - 1) * ns / best, "unit": unit, "cores": cores, "kind": "reference" if use_ref else "port",
            "sample": "first %d of this rank's %s, best of <=3 passes, %s, gcc -O2 (the reference's flag, "
                      "myocamlbuild.ml:6)%s, %d host threads" % (ns, what, code, "" if use_ref else " -ffp-contract=off", cores)}
    variants = {}
    n1 = min(ns, wl.get("cpu_sample_1core", ns))
    s1 = np.ascontiguousarray(sample[:, :n1])
    b1, _ = best_of(runner("o2", 1, s1), budget_s / 2)
    variants["O2_1_thread"] = {"value": (T - 1) * n1 / b1, "unit": unit, "cores": 1, "sample": "first %d %s" % (n1, what)}
    fast_ok = (Ref.available("o3") if use_ref else Oracle.available("o3"))
    if fast_ok:
        b3, _ = best_of(runner("o3", cores, sample), budget_s / 2)
        variants["O3_x86-64-v3_all_threads"] = {"value": (T - 1) * ns / b3, "unit": unit, "cores": cores,
                                               "sample": "first %d %s; gcc -O3 -march=x86-64-v3 (AVX2 + FMA; "
                                                         "-march=native of the build container would not be "
                                                         "portable to this box)" % (ns, what)}
        b31, _ = best_of(runner("o3", 1, s1), budget_s / 2)
        variants["O3_x86-64-v3_1_thread"] = {"value": (T - 1) * n1 / b31, "unit": unit, "cores": 1,
                                            "sample": "first %d %s" % (n1, what)}
    main["variants"] = variants
    return main, res


def reference_model(wl):
    """Model record for the CPU arm built WITHOUT the product library: the eigensystem comes from the
    reference's own diagonalize_* (lib/mlmodel.c:163-262, oracle/_ref) or, where that is not built,
    from numpy.linalg."""
    from oracle.oracle import Ref, numpy_diagonalize

    diag = Ref().diagonalize if Ref.available() else numpy_diagonalize
    return make_model(wl, diagonalize=diag), ("reference diagonalize_* (oracle/_ref, LAPACK)" if Ref.available()
                                              else "numpy.linalg")


def reference_leg(key, wl, steps, warmup):
    """One workload of the CPU arm: `steps` timed passes over a bounded sample, all host threads."""
    from oracle.oracle import Oracle, Ref
    from phylocaml_b200 import tree as tree_mod

    cores = os.cpu_count() or 1
    T, S, K = wl["T"], wl["S"], wl["K"]
    tr = tree_mod.random_tree(T, seed=1)
    ops, ra, rb, rt, n_nodes = tree_mod.schedule(tr)
    ns = min(wl["cpu_sample"], wl["N"])
    if wl["kind"] == "fitch":
        chars, _ = fitch_tips(tree_mod, T, wl["N"], 0, ns, pinned=False)
        if Ref.available():  # the reference's own bv_fitch / bv_distance, compiled unmodified
            ref = Ref()
            fn = lambda: ref.fitch_score_tree(chars, ops, n_nodes, ra, rb, nthreads=cores)
            kind = "reference"
            sample = ("%d taxa x %d chars (of %d), W=8 one char per byte; reference bv_fitch per node + bv_distance "
                      "(lib/bitvector/bv.c, -O2), characters in %d slabs on %d host threads" % (T, ns, wl["N"], cores, cores))
        else:
            orc = Oracle()
            fn = lambda: orc.fitch_score_tree(chars, None, ops, n_nodes, ra, rb, nthreads=cores)["length"]
            kind = "port"
            sample = "%d taxa x %d chars (of %d), W=8 one char per byte (bv.c layout), %d pthreads" % (T, ns, wl["N"], cores)
        metric, unit, dtype, extra = "fitch_char_ops_per_s", "char-ops/s", "u8", {}
    else:
        orc = Oracle()
        model, diag_src = reference_model(wl)
        tips = tree_mod.evolve_tips(tr, model, min(ns, BASE_PATTERNS), seed=3, dtype=mask_dtype(S))
        tips = np.tile(tips, (1, (ns + tips.shape[1] - 1) // tips.shape[1]))[:, :ns].copy()
        fn = lambda: orc.lk_score_tree(model, tips, None, ops, n_nodes, ra, rb, rt, nthreads=cores)["lnl"]
        metric, unit, dtype, kind = "clv_site_updates_per_s", "site-updates/s", "f64", "port"
        sample = ("%d taxa x %d patterns (of %d), oracle C port -O2 -ffp-contract=off (the reference has no pruning "
                  "loop, lib/likelihood_c.ml:1-33), %d pthreads; eigensystem: %s" % (T, ns, wl["N"], cores, diag_src))
        extra = {}
    times, out = [], None
    for _ in range(warmup):
        out = fn()
    for _ in range(steps):
        t0 = time.perf_counter()
        out = fn()
        times.append(time.perf_counter() - t0)
    total = sum(times)
    value = (T - 1) * ns * steps / total
    rec = {"metric": metric, "value": value, "unit": unit, "steps": steps, "warmup": warmup,
           "ms_per_step": 1e3 * total / steps, "dtype": dtype,
           "config": workload_config(key, wl, wl["N"], wl["N"], 1, None, None),
           "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "check": {"result_on_sample": out}}
    rec.update(extra)
    return rec


def run_reference(args, wl):
    """CPU arm (`--impl reference`): the reference's own C where it exists (bv_fitch / bv_distance
    from oracle/_ref for Fitch), otherwise the oracle port of the path, with all host threads, each
    step a bounded sample of the workload. Nothing of the product library is loaded here: the model
    record comes from the reference's diagonalize_* (or numpy)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rec = reference_leg(args.workload, wl, args.steps, args.warmup)
    line = {"impl": "reference", "metric": rec["metric"], "value": rec["value"], "unit": rec["unit"],
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": rec["ms_per_step"],
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": rec["dtype"],
            "data": "synthetic", "config": rec["config"], "cpu_baseline": rec["cpu_baseline"], "e2e": rec["e2e"],
            "gpu_launches": 0, "check": rec["check"]}
    others = secondary_list(args)
    if others:
        line["workloads"] = {}
        for key in others:
            line["workloads"][key] = reference_leg(key, dict(WORKLOADS[key]), 3, 1)
    print(json.dumps(line))


def secondary_list(args):
    if args.workloads == "none" or (args.workloads == "default" and args.workload != "dna"):
        return []
    if args.workloads in ("all", "default"):
        return [k for k in SECONDARY if k != args.workload]
    return [k for k in args.workloads.split(",") if k and k != args.workload]


def workload_config(key, wl, n_total, n_local, world, l2_note, numa_note):
    cfg = {"workload": wl["name"], "taxa": wl["T"], "patterns_total": n_total, "states": wl["S"],
           "rate_classes": wl["K"], "tree": "random topology seed 1, Exp(0.1) branch lengths",
           "tips": "random DNA singletons + 2% two-state ambiguity (2^20 characters, tiled)" if wl["kind"] == "fitch"
           else "evolved under the model (65536 patterns, tiled), 1% missing"}
    if l2_note is not None:
        cfg.update({"patterns_per_gpu": n_local, "l2": l2_note,
                    "sharding": "contiguous 1024-aligned pattern slabs, one process per GPU",
                    "collective": "allreduce of one scalar per step" if world > 1 else "none",
                    "host_placement": numa_note})
    return cfg


class Ctx:
    """What every workload of one bench.py process shares: torch, the process group, the peaks."""

    def __init__(self, torch, dist, world, rank, local, numa_note):
        self.torch, self.dist, self.world, self.rank, self.local, self.numa_note = torch, dist, world, rank, local, numa_note
        self.hbm_peak, self.peak_src = peaks()
        self.flush_buf = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed_steps(self, fn, n, flush):
        """Device time of exactly n calls of fn (ms, max over ranks) and the last result. Large
        working sets: one event pair around all n, barrier + synchronize on both sides. Working sets
        that could survive in the 126 MB L2: 512 MB are written before every iteration and each
        iteration has its own event pair (the flush is outside the timed region); the pairs are summed."""
        torch = self.torch
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        out, total = None, 0.0
        if not flush:
            self.barrier()
            a.record()
            for _ in range(n):
                out = fn()
            b.record()
            self.barrier()
            total = a.elapsed_time(b)
        else:
            if self.flush_buf is None:
                self.flush_buf = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device="cuda")
            self.barrier()
            for i in range(n):
                self.flush_buf.fill_(i & 0xFF)
                a.record()
                out = fn()
                b.record()
                b.synchronize()
                total += a.elapsed_time(b)
            self.barrier()
        tms = torch.tensor([total], dtype=torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(tms, op=self.dist.ReduceOp.MAX)
        return float(tms.item()), out


def run_compress(args, wl, torch, engine, tree_mod, local, hbm_peak, peak_src):
    """raw alignment (host) -> unique site patterns + weights (host), through
    phylo_compress_patterns. A step = one whole alignment on one GPU; with several GPUs visible the record
    gains a `group` block (phylo_group_compress_patterns: slabs of sites per device, tables merged on device 0)."""
    T, N = wl["T"], args.patterns or wl["N"]
    if args.taxa:
        T = args.taxa
    tr = tree_mod.random_tree(T, seed=1)
    model = make_model(WORKLOADS["dna"])
    base = tree_mod.evolve_tips(tr, model, BASE_PATTERNS, seed=3)
    rng = np.random.default_rng(11)
    raw = engine.pinned_empty((T, N), np.uint8)  # page-locked, like the Bigarray staging buffers (phylo_host_alloc)
    raw[:] = base[:, rng.integers(0, base.shape[1], N)]  # every column ~61 times, shuffled
    eng = engine.Engine(local)
    for _ in range(args.warmup):
        pats, wt, s2p = eng.compress_patterns(raw)
    eng.profile(True, reset=True)
    sampler = ClockSampler(local)
    l0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    steps = max(1, args.steps // 4)
    for _ in range(steps):
        pats, wt, s2p = eng.compress_patterns(raw)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    clocks = sampler.stop()
    prof = eng.profile_get()
    eng.profile(False)
    kms, kn = prof.get("compress", (0.0, 0))
    kms_step = kms / max(1, steps)
    P = pats.shape[1]
    comp_bytes = T * N + T * P + 8 * P + 4 * N  # read the alignment, write patterns + weights + site map
    cpu = None
    if not args.no_cpu_baseline:
        from oracle.oracle import compress_patterns as cpu_compress

        ns = min(N, wl["cpu_sample"])
        t0 = time.perf_counter()
        opats, owt, os2p = cpu_compress(raw[:, :ns])
        dt = time.perf_counter() - t0
        cpu = {"value": ns / dt, "unit": "sites/s", "cores": 1, "kind": "port",
               "sample": "first %d sites, numpy void-view unique (oracle/oracle.py)" % ns}
    full_ok = None
    if N <= 2_000_000 or args.no_cpu_baseline is False:
        from oracle.oracle import compress_patterns as cpu_compress

        ns = min(N, 1_000_000)
        a, b, c = eng.compress_patterns(raw[:, :ns])
        oa, ob, oc = cpu_compress(raw[:, :ns])
        full_ok = bool(np.array_equal(a, oa) and np.array_equal(b, ob) and np.array_equal(c, oc))
    line = {
        "metric": "compress_sites_per_s", "value": N / (kms_step * 1e-3) if kms_step else None, "unit": "sites/s",
        "n_gpus": 1, "steps": steps, "warmup": args.warmup, "ms_per_step": kms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": wl["name"], "taxa": T, "sites": N, "patterns_found": int(P),
                   "l2": "inputs exceed L2 (%.2f GB alignment)" % (T * N / 1e9)},
        "roofline": {"bound": "hbm", "kernel": "compress (transpose + hash + insert + verify + scan + gather)",
                     "achieved": comp_bytes / (kms_step * 1e-3) / 1e9 if kms_step else None, "peak": hbm_peak,
                     "unit": "GB/s", "frac": comp_bytes / (kms_step * 1e-3) / 1e9 / hbm_peak if kms_step else None,
                     "traffic": None, "peak_source": peak_src,
                     "bytes_model": "compulsory: alignment read once, patterns + weights + site map written once"},
        "cpu_baseline": cpu,
        "e2e": {"value": N / (ms * 1e-3), "unit": "sites/s", "h2d_bytes_per_step": int(raw.nbytes),
                "d2h_bytes_per_step": int(T * P + 8 * P + 4 * N), "ms_per_step": ms, "steps": steps},
        "gpu_launches": int(eng.launch_count - l0), "clocks": clocks,
        "check": {"patterns": int(P), "weights_sum": float(wt.sum()), "equals_oracle_on_first_1M_sites": full_ok},
    }
    # several GPUs in this process: the same alignment through phylo_group_compress_patterns (slabs of sites on
    # every device, per-slab tables merged on device 0); must reproduce the single-device result exactly
    ndev = torch.cuda.device_count()
    if ndev > 1:
        g = engine.Group(list(range(ndev)))
        gp, gw, gs = g.compress_patterns(raw)
        t0 = time.perf_counter()
        for _ in range(steps):
            gp, gw, gs = g.compress_patterns(raw)
        gms = 1e3 * (time.perf_counter() - t0) / steps
        line["group"] = {"devices": ndev, "ms_per_step_end_to_end": gms, "sites_per_s": N / (gms * 1e-3),
                         "equals_single_device": bool(np.array_equal(gp, pats) and np.array_equal(gw, wt) and np.array_equal(gs, s2p))}
        g.close()
    print(json.dumps(line))
    eng.close()


def run_group(ctx, args, headline):
    """N > 1 only, after every rank has released its engine: rank 0 drives ALL N GPUs through one
    phylo_group handle (the path a single OCaml process would use, include/phylo_engine.h last section) on the
    headline workload -- host wall clock around the synchronous group call (there is no single stream to put
    events on) -- while the other ranks wait at a barrier. lnL must equal the per-rank run's to the last bit."""
    from phylocaml_b200 import engine, tree as tree_mod

    wl = WORKLOADS["dna"]
    T, S, K = wl["T"], wl["S"], wl["K"]
    n_total = args.patterns or wl["N"]
    rec = None
    ctx.barrier()
    if ctx.rank == 0:
        try:
            tr = tree_mod.random_tree(T, seed=1)
            ops, ra, rb, rt, n_nodes = tree_mod.schedule(tr)
            model = make_model(wl)
            base = tree_mod.evolve_tips(tr, model, BASE_PATTERNS, seed=3, dtype=mask_dtype(S))
            tips = engine.pinned_empty((T, n_total), mask_dtype(S))
            tile_cols(base, 0, n_total, out=tips)
            g = engine.Group(list(range(ctx.world)))
            g.lk_set_model(model)
            t0 = time.perf_counter()
            g.lk_set_tips(tips, capacity=n_nodes)
            t_up = time.perf_counter() - t0
            for _ in range(3):
                lnl = g.lk_score_tree(ops, ra, rb, rt)
            reps = max(5, min(args.steps, 20))
            t0 = time.perf_counter()
            for _ in range(reps):
                lnl = g.lk_score_tree(ops, ra, rb, rt)
            dt = (time.perf_counter() - t0) / reps
            rec = {"what": "one process, one phylo_group handle over all %d GPUs (one engine + host worker thread per GPU)" % ctx.world,
                   "devices": ctx.world, "workload": wl["name"], "patterns_total": n_total,
                   "ms_per_step": 1e3 * dt, "value": (T - 1) * n_total / dt, "unit": "site-updates/s",
                   "set_tips_ms": 1e3 * t_up, "upload_format": "%d-byte state masks, column slabs straight out of the caller's matrix" % tips.dtype.itemsize,
                   "lnl": lnl, "lnl_equals_per_rank_run": bool(headline is not None and lnl == headline.get("check", {}).get("result")),
                   "timing": "host perf_counter around synchronous group calls (max over devices by construction)",
                   "kernel_launches": int(g.launch_count)}
            g.close()
            engine.pinned_free(tips)
        except Exception as ex:  # noqa: BLE001 -- a missing record must not cost the headline line
            rec = {"error": str(ex)[:300]}
    # the other ranks wait on the HOST (gloo): an NCCL barrier would park a spinning kernel on their GPUs and
    # time-slice them against the group's engines
    ctx.dist.barrier(group=ctx.cpu_group)
    return rec


def run_directions(torch, engine, tree_mod, local, wl, model, base, ops, ra, rb, rt, n_nodes, n_sample=131072):
    """phylo_lk_uppass + edge joins + phylo_lk_param_gradient on the first n_sample patterns of the workload
    (own engine; one GPU). candidate joins/s: one phylo_lk_edge_lnl per branch of the tree -- what an SPR / TBR
    neighbourhood costs per candidate once the directional CLVs exist (lib/tree.ml:299-494)."""
    from phylocaml_b200 import mlmodel

    T, S, K = wl["T"], wl["S"], wl["K"]
    try:
        up_slot, cap, _, edges = tree_mod.uppass_plan(ops, ra, rb, rt, n_nodes)
        tips = tile_cols(base, 0, n_sample)
        e2 = engine.Engine(local)
        e2.lk_set_model(model)
        e2.lk_set_tips(tips, capacity=cap)
        lnl = e2.lk_score_tree(ops, ra, rb, rt)

        def timed(fn, reps):
            torch.cuda.synchronize()
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0.record()
            for _ in range(reps):
                out = fn()
            b1.record()
            torch.cuda.synchronize()
            return b0.elapsed_time(b1) / reps, out

        e2.lk_uppass(ops, ra, rb, rt, up_slot)
        up_ms, _ = timed(lambda: e2.lk_uppass(ops, ra, rb, rt, up_slot), 2)
        down_ms, _ = timed(lambda: e2.lk_score_tree(ops, ra, rb, rt), 2)
        ea, eb, et = [e[0] for e in edges], [e[1] for e in edges], [e[2] for e in edges]
        vals = e2.lk_edge_lnl_batch(ea, eb, et)
        join_ms, vals = timed(lambda: e2.lk_edge_lnl_batch(ea, eb, et), 2)
        worst = max(abs(x - lnl) / abs(lnl) for x in vals)
        rec = {"patterns": n_sample, "taxa": T, "edges": len(edges), "down_pass_ms": down_ms, "up_pass_ms": up_ms,
               "all_edge_joins_ms": join_ms, "candidate_joins_per_s": len(edges) / (join_ms * 1e-3),
               "join_site_updates_per_s": len(edges) * n_sample / (join_ms * 1e-3),
               "max_rel_diff_of_edge_lnl_vs_root_edge": worst, "tolerance": 1e-11,
               "note": "up pass = 2T-4 pruning updates (one per directional CLV), one launch per tree level for 4 states; each join reads two CLVs; gradient: every branch in one DMMA launch for 4 states"}
        if S == 4:  # GTR+G: 5 exchangeabilities + the Gamma shape, against central differences of full evaluations
            h, alpha = 1e-6, 0.5
            co = np.array(GTR_CO, dtype=float)

            def mk(c, a):
                return mlmodel.create(("GTR", list(c)), 4, pi=GTR_PI, site_var=("gamma", K, a))

            dQ, drates = np.zeros((6, 4, 4)), np.zeros((6, K))
            for p in range(5):
                e = np.zeros(5)
                e[p] = h
                dQ[p] = (mk(co + e, alpha)["Q"] - mk(co - e, alpha)["Q"]) / (2 * h)
            drates[5] = (mk(co, alpha + h)["rates"] - mk(co, alpha - h)["rates"]) / (2 * h)
            g = e2.lk_param_gradient(ops, ra, rb, rt, up_slot, dQ=dQ, drates=drates)
            # (best of three single calls: the call has a host part -- dP/dtheta per branch -- and one slow pass on
            # a busy host core once read 97 ms against the usual 29)
            grad_ms, g = min((timed(lambda: e2.lk_param_gradient(ops, ra, rb, rt, up_slot, dQ=dQ, drates=drates), 1)
                              for _ in range(3)), key=lambda t: t[0])

            def f(c, a):
                e2.lk_set_model(mk(c, a))
                e2.lk_set_tips(tips, capacity=cap)
                return e2.lk_score_tree(ops, ra, rb, rt)

            hh = 1e-4
            t0 = time.perf_counter()
            fd = []
            for p in range(6):
                e = np.zeros(5)
                if p < 5:
                    e[p] = hh
                    fd.append((f(co + e, alpha) - f(co - e, alpha)) / (2 * hh))
                else:
                    fd.append((f(co, alpha + hh) - f(co, alpha - hh)) / (2 * hh))
            fd_ms = 1e3 * (time.perf_counter() - t0)
            rec["param_gradient"] = {"parameters": "5 GTR exchangeabilities + Gamma shape", "gradient_ms": grad_ms,
                                     "central_differences_ms_incl_model_setup": fd_ms, "gradient": [float(x) for x in g],
                                     "max_rel_diff_vs_central_differences": float(max(abs(a - b) / max(1.0, abs(b)) for a, b in zip(g, fd)))}
        e2.close()
        return rec
    except Exception as ex:  # noqa: BLE001 -- a side record must not cost the headline line
        return {"error": str(ex)[:300]}


def setup_exchange(ctx, args, eng):
    """Peer-mapped mailboxes between the ranks' engines (cudaIpc handles travel over the process group).
    Returns a note for the JSON line; ctx-independent state lives in the engine. Every rank must take the
    same branch, so the outcome is agreed on with an all-reduce."""
    if ctx.world == 1:
        return None
    if args.exchange != "p2p":
        return "nccl all-reduce of one scalar (host -> device -> NCCL -> host)"
    torch, dist = ctx.torch, ctx.dist
    ok, why = 1, ""
    try:
        box, handle = eng.exchange_alloc()
        handles = [None] * ctx.world
        dist.all_gather_object(handles, handle)
        boxes = [box if r == ctx.rank else eng.exchange_open(handles[r]) for r in range(ctx.world)]
        eng.exchange_set(ctx.rank, boxes)
    except Exception as ex:  # noqa: BLE001 -- fall back, but say so
        ok, why = 0, str(ex)
    flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        return "nccl all-reduce (peer-mapped exchange unavailable on some rank: %s)" % (why or "see other ranks")
    eng.set_option(eng.OPT_DEFER_SCALAR, 1)
    return ("device-side: block partials into peer-mapped mailboxes (cudaIpc over NVLink), canonical fold in rank "
            "order by one CTA per rank, no NCCL call and no host round trip before the final scalar")


def run_workload(ctx, args, key, wl, primary):
    """One workload on this process's GPU (all ranks call it together). Returns the JSON record on
    rank 0 (None elsewhere). `primary`: the headline -- full step counts, the other likelihood modes,
    the branch-length loop; secondary workloads run fewer steps and skip the mode comparison (codon
    keeps the branch-length loop: it is half of BASELINE config 5)."""
    from phylocaml_b200 import engine, tree as tree_mod

    torch, dist, world, rank, local = ctx.torch, ctx.dist, ctx.world, ctx.rank, ctx.local
    hbm_peak, peak_src = ctx.hbm_peak, ctx.peak_src
    kind = wl["kind"]
    steps = args.steps if primary else max(5, min(args.steps, 10))
    warmup = args.warmup if primary else max(3, min(args.warmup, 3))
    e2e_steps = args.e2e_steps if primary else 3
    other_modes = primary and not args.no_other_modes
    want_branch_loop = (primary and not args.no_other_modes) or key == "codon"
    mode = args.mode if kind == "lk" else None
    T, S, K = wl["T"], wl["S"], wl["K"]
    n_total = wl["N"] * (world if args.scaling == "weak" else 1)
    lo, hi = shard_bounds(n_total, world, rank)
    n_local = hi - lo
    tr = tree_mod.random_tree(T, seed=1)
    ops, ra, rb, rt, n_nodes = tree_mod.schedule(tr)
    eng = engine.Engine(local)
    launches0 = eng.launch_count
    xnote = setup_exchange(ctx, args, eng)
    p2p = xnote is not None and xnote.startswith("device-side")

    def needs_flush(m):
        return device_footprint("fitch" if kind == "fitch" else key, T, S, K, mask_dtype(S)().itemsize, m, n_local) < 4 * L2_BYTES

    base = None
    if kind == "fitch":
        model = None
        # the host alignment is kept in the engine's compact upload format: bit-sliced planes (0.5 B per
        # DNA character; phylo_fitch_pack_planes), built from the tiled base block. The one-byte-per-character
        # form (the reference's W = 8 layout) is only materialised for the CPU leg's sample.
        base = tree_mod.random_fitch_chars(T, min(n_total, FITCH_BASE), 4, seed=5)
        planes_base = engine.fitch_pack_planes(base, 4)          # T x (base/32 * 4) uint32
        if args.upload == "compact" and (base.shape[1] % 32 == 0 or n_local <= base.shape[1]):
            wpl = 4  # uint32 words per 32 characters
            tips = engine.pinned_empty((T, ((n_local + 31) // 32) * wpl), np.uint32)
            tile_cols(planes_base, (lo // 32) * wpl, tips.shape[1], out=tips)
            set_tips = lambda: eng.fitch_set_tips_planes(tips, n_local, 4, capacity=n_nodes)
            upload_note = "bit-sliced planes, 0.5 B per character (elt_bytes = 0)"
        else:
            tips = engine.pinned_empty((T, n_local), np.uint8)
            tile_cols(base, lo, n_local, out=tips)
            set_tips = lambda: eng.fitch_set_tips(tips, 4, capacity=n_nodes)
            upload_note = "one byte per character (the reference's W = 8 layout), transcoded on the device"
        sample_of = lambda ns: tile_cols(base, lo, ns)
        eng.set_option(eng.OPT_FITCH_WALK, {"auto": 1, "tile": 3, "regwalk": 2, "l2": 0}[args.fitch_kernel])
        set_tips()
        acc = torch.zeros(1, dtype=torch.int64, device="cuda")

        def step():
            v = eng.fitch_score_tree(ops, ra, rb)
            if p2p:
                return eng.exchange_sum_u64(v)  # the path's only exchange: one exact integer
            if world > 1:
                acc[0] = v
                dist.all_reduce(acc)
                return int(acc.item())
            return v

        def e2e_step():
            set_tips()
            return step()

        metric, unit, dtype = "fitch_char_ops_per_s", "char-ops/s", "u32 (bit-sliced state planes)"
        kbytes = {"fitch_tree": (2 * T - 1) * 0.5}  # compulsory: every node set moved once
        h2d = tips.nbytes
    else:
        model = make_model(wl)
        base = tree_mod.evolve_tips(tr, model, BASE_PATTERNS, seed=3, dtype=mask_dtype(S))
        packed_base = engine.pack_nibbles(base) if S == 4 and args.upload == "compact" else None
        # 20 / 61 states, compact upload: one byte per cell -- the alphabet symbol (state index, or S for a missing
        # cell) -- translated to state masks on the device through the engine's symbol table
        # (phylo_engine_set_symbol_table, what an Alphabet.t supplies); 4 / 8 times fewer PCIe bytes than the masks
        sym_base, sym_table = None, None
        if S > 4 and S < 255 and args.upload == "compact":
            b64 = base.astype(np.uint64)
            full = np.uint64((1 << S) - 1)
            idx = np.full(b64.shape, 255, dtype=np.uint8)
            for st in range(S):
                idx[b64 == np.uint64(1 << st)] = st
            idx[b64 == full] = S
            if not np.any(idx == 255):  # no partial ambiguity codes in this alignment
                sym_base = idx
                sym_table = np.zeros(256, dtype=np.uint64)
                for st in range(S):
                    sym_table[st] = np.uint64(1 << st)
                sym_table[S] = full

        def host_shard(lo_, n_):
            """this rank's slab of the host alignment in the upload format (pinned)"""
            if packed_base is not None and lo_ % 2 == 0:
                # compact upload format: two 4-bit masks per byte (mask_bytes = 0; phylo_pack_nibbles)
                t = engine.pinned_empty((T, (n_ + 1) // 2), np.uint8)
                tile_cols(packed_base, lo_ // 2, t.shape[1], out=t)
                return t, n_, "packed 4-bit masks, 0.5 B per cell (mask_bytes = 0)"
            if sym_base is not None:
                t = engine.pinned_empty((T, n_), np.uint8)
                tile_cols(sym_base, lo_, n_, out=t)
                return t, None, "1-byte alphabet symbols, translated by the engine's symbol table (phylo_engine_set_symbol_table)"
            t = engine.pinned_empty((T, n_), mask_dtype(S))
            tile_cols(base, lo_, n_, out=t)
            return t, None, "%d-byte state masks" % t.dtype.itemsize

        tips, packed_n, upload_note = host_shard(lo, n_local)
        sample_of = lambda ns: tile_cols(base, lo, ns)
        eng.lk_set_model(model)
        if sym_table is not None:
            eng.set_symbol_table(sym_table)

        def set_mode(m):
            eng.set_option(eng.OPT_FUSED_TREE, 0 if m == "pernode" else 1)
            eng.set_option(eng.OPT_RETAIN_CLV, 0 if m == "fused-lnl" else 1)

        set_mode(mode)
        eng.lk_set_tips(tips, capacity=n_nodes, packed_n=packed_n)
        acc = torch.zeros(1, dtype=torch.float64, device="cuda")

        def step():
            v = eng.lk_score_tree(ops, ra, rb, rt)
            if p2p:
                return eng.lk_exchange_reduce()  # the path's only exchange; bit-identical for any N
            if world > 1:
                acc[0] = v
                dist.all_reduce(acc)  # one fp64 scalar
                return float(acc.item())
            return v

        def e2e_step():
            # host buffers in, scalar out: upload (pinned -> HBM, overlapped slab by slab with
            # the scoring) + full evaluation + lnL read-back, through one C-ABI call
            v = eng.lk_score_alignment(tips, ops, ra, rb, rt, capacity=n_nodes, packed_n=packed_n)
            if p2p:
                return eng.lk_exchange_reduce()
            if world > 1:
                acc[0] = v
                dist.all_reduce(acc)
                return float(acc.item())
            return v

        metric, unit, dtype = "clv_site_updates_per_s", "site-updates/s", "f64"
        kbytes = lk_bytes(T, S, K, mask_dtype(S)().itemsize, mode)
        h2d = tips.nbytes + ops.nbytes
    units_per_step = (T - 1) * n_total

    for _ in range(warmup):
        result = step()
    # ---- N > 1, --balance (an experiment, off by default): slabs sized to each GPU's measured speed. The ranks
    # meet at the scalar exchange of every step, so a step costs what the slowest GPU needs, and under the 1 kW
    # power cap the GPUs of a box run at different clocks (1780-1920 MHz seen). Shares proportional to patterns /
    # kernel-time of three more warm-up steps, still whole 1024-pattern blocks (lnL stays bit-identical: the
    # canonical fold does not care where the cuts are). Result on 8 GPUs: no gain (3.45 vs 3.40 ms) -- which GPU
    # is slow changes from one second to the next.
    balance = None
    if world > 1 and kind == "lk" and args.balance and n_total >= 1024 * 64 * world:
        eng.profile(True, reset=True)
        for _ in range(3):
            step()
        pr = eng.profile_get()
        eng.profile(False)
        t_mine = sum(v[0] for k, v in pr.items() if k not in ("reduce1024",)) / 3.0  # the exchange wait is not work
        every = [None] * world
        dist.all_gather_object(every, (n_local, t_mine))
        speed = np.array([n / max(t, 1e-9) for n, t in every])
        blocks_total = (n_total + 1023) // 1024
        want = speed / speed.sum() * blocks_total
        nb = np.maximum(1, np.floor(want).astype(np.int64))
        for i in np.argsort(-(want - np.floor(want)))[: int(blocks_total - nb.sum())]:
            nb[i] += 1
        cuts = np.concatenate([[0], np.cumsum(nb)]) * 1024
        cuts[-1] = n_total
        new_lo, new_hi = int(cuts[rank]), int(cuts[rank + 1])
        balance = {"patterns_per_rank_before": [int(n) for n, _ in every], "kernel_ms_per_rank_before": [float(t) for _, t in every],
                   "patterns_per_rank": [int(cuts[i + 1] - cuts[i]) for i in range(world)]}
        if (new_lo, new_hi) != (lo, hi):
            engine.pinned_free(tips)
            lo, hi, n_local = new_lo, new_hi, new_hi - new_lo
            tips, packed_n, upload_note = host_shard(lo, n_local)
            eng.lk_set_tips(tips, capacity=n_nodes, packed_n=packed_n)
            h2d = tips.nbytes + ops.nbytes
        for _ in range(2):
            result = step()
    eng.profile(True, reset=True)
    sampler = ClockSampler(local) if rank == 0 else None
    flush = needs_flush(mode)
    ms, result = ctx.timed_steps(step, steps, flush)
    prof = eng.profile_get()
    eng.profile(False)
    ms_with_kernel_events = None
    if flush:
        # microsecond steps (working set near the L2 size: one or two launches per step): the engine profiler's
        # event pair around every launch costs ~3 us per event boundary on the device, a fifth of such a step.
        # The per-kernel durations come from the pass above; the step time from a second pass of the same K
        # steps without those inner events (both reported).
        ms_with_kernel_events = ms / steps
        ms, result = ctx.timed_steps(step, steps, flush)
    launches = eng.launch_count - launches0
    step_launches = (eng.launch_count - launches0) // ((2 if flush else 1) * steps + warmup)
    # a timed region shorter than a few nvidia-smi samples (50 ms apart): keep the same load running,
    # untimed, so that the clocks / throttle reasons are sampled under it (the count is derived from
    # the all-reduced time, so every rank runs the same number of steps)
    want_ms = 600.0 if primary else 300.0
    extra_steps = 0 if ms >= want_ms else min(20000, int(np.ceil((want_ms - ms) / max(ms / steps, 1e-3))))
    for _ in range(extra_steps):
        step()
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["untimed_steps_run_for_sampling"] = extra_steps
    value = units_per_step * steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel class (CUDA events on the launching stream)
    roof, kernels = None, {}
    tot_kernel_ms = sum(v[0] for v in prof.values()) or 1.0
    for name, (kms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        per_launch_ms = kms / n
        entry = {"launches_per_step": n / steps, "avg_us": 1e3 * per_launch_ms,
                 "share_of_kernel_time": kms / tot_kernel_ms}
        if name in kbytes:
            by = kbytes[name] * n_local
            entry["algorithmic_bytes_per_launch"] = by
            entry["achieved_gbs"] = by / (per_launch_ms * 1e-3) / 1e9
            entry["frac"] = entry["achieved_gbs"] / hbm_peak
            if roof is None:
                traffic = None
                tpath = os.path.join(ROOT, "profiles", "traffic.json")
                if os.path.exists(tpath):
                    with open(tpath) as f:
                        tj = json.load(f)
                    tkey = "fitch" if kind == "fitch" else key
                    for t in (tj.get(name), tj.get(name + "_small"), tj.get(name + "_" + key)):
                        if (t and t.get("patterns") and t.get("workload", "dna") in (tkey, key) and
                                t.get("taxa") in (None, T) and
                                t.get("min_patterns", 0) <= n_local <= t.get("max_patterns", 1 << 62)):
                            traffic = t["dram_bytes_per_launch"] * n_local / t["patterns"]
                roof = {"bound": "hbm", "kernel": name, "achieved": entry["achieved_gbs"], "peak": hbm_peak,
                        "unit": "GB/s", "frac": entry["frac"], "traffic": traffic, "peak_source": peak_src,
                        "bytes_model": ("compulsory (tree-fused: each interior CLV + scale counter written once, "
                                        "tips read once)" if name == "tree_fused" else
                                        "SURVEY 8(d) per-node streaming") if kind != "fitch"
                        else "compulsory (tree-fused: every node set moved once)"}
        kernels[name] = entry
    # 20 / 61 states: the update is a dense contraction (K S (2 (2S - 1) + 1) flops per
    # update); the inner+inner kernel is held against the measured fp64 DMMA peak as well. For 61
    # states that is the governing roofline (15 flop/B), for 20 states the two rooflines meet.
    roof_tensor = None
    if kind == "lk" and S >= 20 and "prune_inner_inner" in kernels:
        fl = K * S * (2 * (2 * S - 1) + 1) * n_local
        ent = kernels["prune_inner_inner"]
        tf = fl / (ent["avg_us"] * 1e-6) / 1e12
        ent["algorithmic_flops_per_launch"] = fl
        ent["achieved_tflops"] = tf
        roof_tensor = {"bound": "tensor", "kernel": "prune_inner_inner", "achieved": tf, "peak": FP64_DMMA_TFLOPS,
                       "unit": "TFLOP/s", "frac": tf / FP64_DMMA_TFLOPS, "traffic": None,
                       "peak_source": "fp64 DMMA (mma.sync.m8n8k4.f64) measured by tools/peaks.cu on this pool",
                       "step_level": {"algorithmic_flops_per_step": fl * (T - 1),
                                      "achieved_tflops": fl * (T - 1) / (ms * 1e-3 / steps) / 1e12,
                                      "note": "SURVEY's dense count for every update; tip sides are table "
                                              "lookups in the engine, so this exceeds the executed flops"}}
        roof_tensor["step_level"]["frac"] = roof_tensor["step_level"]["achieved_tflops"] / FP64_DMMA_TFLOPS
    if kind == "lk" and S >= 20 and roof_tensor is None and "tree_fused" in kernels:
        # tree-fused DMMA kernel (lk_treem_kernel): only CLV operands go through the tensor cores (an observed
        # tip contributes a column of P by table lookup), root edge: one side. Counted twice: the useful
        # contraction flops (2 S^2 per side, class and pattern) and the DMMA work issued (row tiles padded to 8)
        sides = int(np.sum(ops["left"] >= T) + np.sum(ops["right"] >= T)) + (1 if rb >= T else 0)
        mt, ks = (S + 7) // 8, (S + 3) // 4
        useful = 2.0 * S * S * K * sides * n_local
        issued = 512.0 * mt * ks * K * sides * ((n_local + 7) // 8)
        ent = kernels["tree_fused"]
        t_s = ent["avg_us"] * 1e-6
        ent["algorithmic_flops_per_launch"] = useful
        ent["issued_dmma_flops_per_launch"] = issued
        roof_tensor = {"bound": "tensor", "kernel": "tree_fused", "achieved": useful / t_s / 1e12, "peak": FP64_DMMA_TFLOPS,
                       "unit": "TFLOP/s", "frac": useful / t_s / 1e12 / FP64_DMMA_TFLOPS, "traffic": None,
                       "issued_tflops": issued / t_s / 1e12, "issued_frac": issued / t_s / 1e12 / FP64_DMMA_TFLOPS,
                       "clv_operand_sides": sides,
                       "peak_source": "fp64 DMMA (mma.sync.m8n8k4.f64) measured by tools/peaks.cu on this pool",
                       "note": "tip operands are table lookups (no flops counted); the same launch is also held against "
                               "the HBM roofline (compulsory bytes: every interior CLV written once)"}
    if kind == "lk":
        step_bytes = kbytes["tree"] * n_local
        roof_step = {"bytes_model": "SURVEY 8(d) per-node streaming (what a kernel-per-node engine must move)",
                     "algorithmic_bytes_per_step": step_bytes,
                     "achieved_gbs": step_bytes / (ms * 1e-3 / steps) / 1e9}
        roof_step["frac"] = roof_step["achieved_gbs"] / hbm_peak
    else:
        step_bytes = kbytes["fitch_tree"] * n_local
        roof_step = {"bytes_model": "compulsory, against the whole step (host schedule work + launch + result read-back "
                                    "included), not the kernel alone",
                     "algorithmic_bytes_per_step": step_bytes,
                     "achieved_gbs": step_bytes / (ms * 1e-3 / steps) / 1e9}
        roof_step["frac"] = roof_step["achieved_gbs"] / hbm_peak

    # ---- e2e: host buffers in, scalar out, through the C ABI, every step
    e2e_step()
    e2e_ms, result_e2e = ctx.timed_steps(e2e_step, e2e_steps, flush)
    e2e_val = units_per_step * e2e_steps / (e2e_ms * 1e-3)
    e2e = {"value": e2e_val, "unit": unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 8,
           "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps}

    # ---- the other likelihood modes, briefly (same inputs, same timing method)
    modes = None
    if kind == "lk" and other_modes:
        modes = {mode: {"value": value, "ms_per_step": ms / steps, "lnl": result}}
        for m in ("fused", "fused-lnl", "pernode"):
            if m == mode:
                continue
            set_mode(m)
            for _ in range(2):
                r_m = step()
            nm = max(3, steps // 4)
            tm, r_m = ctx.timed_steps(step, nm, needs_flush(m))
            modes[m] = {"value": units_per_step * nm / (tm * 1e-3), "ms_per_step": tm / nm, "lnl": r_m,
                        "l2_flushed_between_steps": needs_flush(m)}
        set_mode(mode)
        step()  # leave the engine's site lnL / CLVs in the primary mode's state

    # ---- Fitch, length only (PHYLO_OPT_RETAIN_CLV = 0): what config 2 literally asks for -- no interior set is
    # written (compulsory bytes: the tips once) and the tree is evaluated from its centre edge (phylo_fitch_reroot)
    if kind == "fitch" and not args.no_other_modes:
        modes = {"retain": {"value": value, "ms_per_step": ms / steps, "length": result,
                            "note": "every interior set written once (the primary figures of this record)"}}
        eng.set_option(eng.OPT_RETAIN_CLV, 0)
        for _ in range(3):
            r_m = step()
        fl = device_footprint("fitch", T, S, K, mask_dtype(S)().itemsize, mode, n_local) < 8 * L2_BYTES
        tm, r_m = ctx.timed_steps(step, steps, fl)
        eng.set_option(eng.OPT_RETAIN_CLV, 1)
        by = T * 0.5 * n_local
        modes["length-only"] = {"value": units_per_step * steps / (tm * 1e-3), "ms_per_step": tm / steps, "length": r_m,
                                "l2_flushed_between_steps": fl, "algorithmic_bytes_per_step": by,
                                "achieved_gbs_step": by / (tm / steps * 1e-3) / 1e9,
                                "frac_step": by / (tm / steps * 1e-3) / 1e9 / hbm_peak,
                                "note": "whole step (launch + result read-back included) against the tips read once"}
        step()  # leave the interior sets resident again

    if p2p:
        eng.set_option(eng.OPT_DEFER_SCALAR, 0)  # what follows is per-rank work that reads local values
    # ---- branch-length re-evaluation loop (BASELINE config 5's second half):
    # 50 evaluations of lnL(t) on the root edge as a Brent/Newton driver would issue them (one
    # length per call: each depends on the previous result), then 10 re-prunes of the path to
    # the root after one branch changed, all other CLVs staying resident.
    branch_loop = None
    if kind == "lk" and mode == "fused" and want_branch_loop:
        def timed(fn, reps):
            torch.cuda.synchronize()
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0.record()
            for i in range(reps):
                out = fn(i)
            b1.record()
            torch.cuda.synchronize()
            return b0.elapsed_time(b1) / reps, out

        step()  # every interior CLV resident
        ts = [rt * f for f in np.geomspace(0.05, 20.0, 50)]
        eng.lk_edge_prepare(ra, rb)  # first call allocates the table and loads the kernels
        prep_ms, _ = timed(lambda i: eng.lk_edge_prepare(ra, rb), 3)
        eval_ms, last = timed(lambda i: eng.lk_edge_eval([ts[i]]), 50)
        batch_ms, _ = timed(lambda i: eng.lk_edge_eval(ts[:16]), 3)
        direct_ms, direct = timed(lambda i: eng.lk_edge_lnl(ra, rb, [ts[min(i, 49)]]), 5)
        opt_ms, opt = timed(lambda i: eng.lk_optimize_branch(ra, rb, t0=rt), 1)
        # path from a deep op up to the root end
        parent_of = {int(o["left"]): j for j, o in enumerate(ops)}
        parent_of.update({int(o["right"]): j for j, o in enumerate(ops)})
        depth_of = lambda j: 0 if int(ops[j]["parent"]) not in parent_of else 1 + depth_of(parent_of[int(ops[j]["parent"])])
        i0 = max(range(len(ops)), key=depth_of)
        path, j = [i0], i0
        while int(ops[j]["parent"]) in parent_of:
            j = parent_of[int(ops[j]["parent"])]
            path.append(j)
        sub = ops[sorted(path)].copy()

        def reprune(i):
            sub[0]["t_left"] = float(ops[i0]["t_left"]) * (1.0 + 0.01 * (i + 1))
            return eng.lk_score_tree(sub, ra, rb, rt)

        rep_ms, rep_lnl = timed(reprune, 10)
        branch_loop = {
            "edge": "root edge", "sumtable_prepare_ms": prep_ms, "lnl_d1_d2_eval_ms": eval_ms, "evals": 50,
            "eval_16_lengths_one_pass_ms": batch_ms, "p_matrix_edge_lnl_ms": direct_ms,
            "newton_optimize_ms": opt_ms, "newton_iters": opt[2], "t_opt": opt[0],
            "reprune_path_ops": len(path), "reprune_ms": rep_ms, "reprunes": 10,
            "loop_total_ms": prep_ms + 50 * eval_ms + 10 * rep_ms,
            "note": "per-rank times (no allreduce); sum table = one CLV-sized pass per evaluation; no L2 flush "
                    "inside this loop: consecutive evaluations re-read the same %.0f MB table, as the "
                    "optimiser's loop does" % (n_local * S * K * 8 / 1e6),
        }
        eng.lk_set_tips(tips, capacity=n_nodes, packed_n=packed_n)
        step()  # restore the unmodified tree's state

    # ---- every edge as a root edge (3-directional CLVs) and the model-parameter gradient
    # (rank 2), on a pattern sample small enough to hold the 2 T - 4 extra CLVs next to the engine above
    directions = None
    if kind == "lk" and primary and not args.no_other_modes and rank == 0:
        directions = run_directions(torch, engine, tree_mod, local, wl, model, base, ops, ra, rb, rt, n_nodes)

    # ---- CPU baseline + correctness check against the oracle (rank 0, N=1 only)
    cpu, check = None, {"result": result, "result_e2e": result_e2e}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ns = min(wl["cpu_sample"], n_local)
        sample = sample_of(ns)
        cpu, res = cpu_legs(wl, kind, ops, ra, rb, rt, n_nodes, model, sample, budget_s=8.0 if primary else 4.0)
        if kind == "fitch":
            check["oracle_length_of_sample"] = int(res["length"])
            if ns == n_local:
                check["bit_exact"] = bool(res["length"] == result == result_e2e)
            elif ns == FITCH_BASE and n_local % FITCH_BASE == 0:
                # the alignment is the sample tiled: the length of the whole is an exact multiple
                want = int(res["length"]) * (n_local // FITCH_BASE)
                check["expected_from_tiling"] = want
                check["bit_exact"] = bool(want == result == result_e2e)
        else:
            site = eng.lk_get_site_lnl()
            ref_site = res["site_lnl"]
            check["site_lnl_max_rel_err_vs_oracle_on_sample"] = float(np.max(np.abs(site[:ns] - ref_site)) / np.max(np.abs(ref_site)))
            check["tolerance"] = 1e-9
            if ns == n_local:
                check["lnl_rel_err_vs_oracle"] = float(abs(result - res["lnl"]) / abs(res["lnl"]))
            check["e2e_equals_resident"] = bool(result == result_e2e)

    foot = device_footprint("fitch" if kind == "fitch" else key, T, S, K, mask_dtype(S)().itemsize, mode, n_local)
    l2_note = ("L2 flushed between timed iterations (%d MB written before each, outside its event pair; "
               "working set %.3f GB per GPU could stay in the 126 MB L2)" % (FLUSH_BYTES >> 20, foot / 1e9)
               if flush else "inputs exceed L2 (working set %.1f GB per GPU, 126 MB L2); no flush" % (foot / 1e9))
    line = None
    if rank == 0:
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": dict(workload_config(key, wl, n_total, n_local, world, l2_note, ctx.numa_note), upload_format=upload_note,
                           **({"collective": xnote} if xnote else {}),
                           **({"sharding": "contiguous 1024-aligned pattern slabs sized to each GPU's measured speed "
                                           "(warm-up kernel times), one process per GPU", "balance": balance} if balance else {})),
            "mode": mode, "modes": modes,
            "roofline": (roof_tensor if roof_tensor and S > 32 else roof),
            "roofline_hbm": roof if roof_tensor and S > 32 else None,
            "roofline_tensor": roof_tensor if roof_tensor and S <= 32 else None,
            "roofline_step": roof_step, "kernels": kernels, "cpu_baseline": cpu,
            "e2e": e2e, "ms_per_step_with_kernel_events": ms_with_kernel_events, "branch_loop": branch_loop, "directions": directions, "gpu_launches": int(step_launches * steps),
            "gpu_launches_total_incl_warmup": int(launches), "clocks": clocks, "check": check,
        }
    eng.close()
    engine.pinned_free(tips)
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="dna", choices=sorted(WORKLOADS), help="the headline workload")
    ap.add_argument("--workloads", default="default",
                    help="secondary workloads reported under \"workloads\": all | none | comma list "
                         "(default: all when the headline is dna)")
    ap.add_argument("--patterns", type=int, default=0, help="override the headline's total pattern count")
    ap.add_argument("--taxa", type=int, default=0, help="override the headline's number of taxa")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--upload", default="compact", choices=["compact", "bytes"],
                    help="host alignment format of the end-to-end leg: the engine's compact formats (packed 4-bit "
                         "masks / bit-sliced planes) or one element per cell as the reference holds it")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="fused", choices=["fused", "fused-lnl", "pernode"],
                    help="likelihood path: tree-fused kernel keeping every CLV (default), tree-fused "
                         "lnL-only (no CLV written), or one streaming kernel per node")
    ap.add_argument("--no-other-modes", action="store_true", help="skip the short runs of the other modes")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="how the ranks' scalars meet (N > 1): on the device through peer-mapped mailboxes "
                         "(phylo_lk_exchange_reduce; default) or host -> NCCL all-reduce -> host")
    ap.add_argument("--balance", action="store_true",
                    help="N > 1: slabs proportional to each GPU's warm-up kernel speed instead of equal ones "
                         "(measured on 8 B200s: 3.45 ms against 3.40 ms for equal slabs -- the GPUs' speeds wander "
                         "under the power cap instead of differing persistently; off by default)")
    ap.add_argument("--no-group", action="store_true", help="skip the single-process phylo_group record (N > 1)")
    ap.add_argument("--fitch-kernel", default="auto", choices=["auto", "tile", "regwalk", "l2"],
                    help="whole-tree Fitch kernel (PHYLO_OPT_FITCH_WALK)")
    args = ap.parse_args()
    assert args.warmup >= 0 and args.steps >= 1
    wl = dict(WORKLOADS[args.workload])
    if args.patterns:
        wl["N"] = args.patterns
    if args.taxa:
        wl["T"] = args.taxa
    if args.patterns or args.taxa:
        wl["name"] += " [overridden: %d taxa x %d patterns]" % (wl["T"], wl["N"])

    if args.impl == "reference":
        run_reference(args, wl)
        return

    import torch

    from phylocaml_b200 import engine, tree as tree_mod

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa_note = host_placement(torch, local, world)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Ctx(torch, dist, world, rank, local, numa_note)
    ctx.cpu_group = dist.new_group(backend="gloo") if world > 1 else None

    if args.workload == "compress":
        run_compress(args, wl, torch, engine, tree_mod, local, ctx.hbm_peak, ctx.peak_src)
        return

    t_all = time.perf_counter()
    line = run_workload(ctx, args, args.workload, wl, primary=True)
    others = secondary_list(args)
    if others:
        recs = {}
        for key in others:
            t0 = time.perf_counter()
            rec = run_workload(ctx, args, key, dict(WORKLOADS[key]), primary=False)
            if rec is not None:
                rec["bench_wall_s"] = time.perf_counter() - t0
                recs[key] = rec
        if line is not None:
            line["workloads"] = recs
    if world > 1 and not args.no_group and args.workload == "dna":
        grp = run_group(ctx, args, line)
        if line is not None:
            line["group"] = grp
    if line is not None:
        line["bench_wall_s"] = time.perf_counter() - t_all
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
