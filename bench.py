#!/usr/bin/env python
"""Benchmark of the tree-scoring hot path (BASELINE.json metric: trees x sites scored / second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload lnl_target|lnl_c2|fitch_target|trees_c4|aa_c5] [--no-cpu] [--no-extra]

Default workload = the north_star target shape: synthetic DNA, 256 taxa x 100,000 sites (i.i.d. uniform
ACGT, so patterns = sites), GTR+Gamma4 log-likelihood at fixed branch lengths of EVERY distinct SPR
neighbour (255,530) of a random tree.  One "step" = one pass of the hot path over that batch: plan
(host, O(n)) + score (device) + the per-tree scores and the move table back on the host.  The other
workloads are BASELINE.json's remaining configs, each a driver-parsable headline of its own
(`--workload`), and -- in the default run at N = 1 -- compact legs under "other_workloads".

N > 1 (one process per GPU, torchrun): weak scaling -- the job is the pooled list of N such batches
(rank r scores batch r: the neighbours of base tree r, or block r of the random-tree list); every
GPU holds the full alignment, there is no data-path collective, and the per-tree scores are written
by the library into a device tensor that goes straight into ONE NCCL all-gather
(pylogeny_b200/distributed.py).  "one_neighbourhood_sharded" (N > 1) additionally splits ONE
neighbourhood across the ranks: strong scaling.

One JSON line on stdout (rank 0).  `roofline` names the dominant kernel and the resource that binds
it: "fp64" (FP64 pipe: the DMMA / DFMA peak is MEASURED IN THIS RUN by plg_fp64_peak) for the
likelihood walks, "int" (INT32 logic pipe, measured in this run) for the Fitch walks, "hbm"
(MEASURED_PEAKS.json copy bandwidth) for the level kernels; algorithmic bytes and
DRAM traffic (ncu) are separate fields.  `cpu_baseline` / `--impl reference` = the CPU side on
sample trees that the REFERENCE's own rearrangement code enumerated (tests/golden/bench_ref_trees.npz,
tests/golden/make_bench_fixture.py): that arm imports nothing of pylogeny_b200.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "tree-sites/s"
EXCH = np.array([1.2, 3.4, 0.8, 1.1, 4.1, 1.0])
FREQS = np.array([0.28, 0.22, 0.24, 0.26])
ALPHA = 0.7
FLOP_PER_UNIT_DNA = 640.0   # per neighbour x pattern: CLV update 240 + 3-way evaluation ~400
FLOP_PER_NODEOP_DNA = 240.0  # per node-op x pattern of a full traversal (R (2 K (2K - 1) + K))
FLOP_PER_NODEOP_AA = 6320.0
INTOPS_PER_FITCH_UNIT = 20.0 / 32.0  # ~20 logic ops per 32 patterns per node-op (SURVEY.md 8d caveat i)

WORKLOADS = {
    "lnl_target": dict(kind="spr", score="lnl", datatype="dna", n_taxa=256, n_sites=100_000, seed=1002, ref_sample=48,
                       metric="trees x sites scored / sec (GTR+G4 logL, all SPR neighbours, 256 taxa x 100k sites)",
                       label="north_star target shape (BASELINE configs[2] alignment): synthetic DNA 256 taxa x 100k sites, "
                             "GTR+G4 lnL (fixed branch lengths) of all %d distinct SPR neighbours of a random tree"),
    "lnl_c2": dict(kind="spr", score="lnl", datatype="dna", n_taxa=64, n_sites=10_000, seed=1001, ref_sample=192,
                   metric="trees x sites scored / sec (GTR+G4 logL, all SPR neighbours, 64 taxa x 10k sites)",
                   label="BASELINE configs[1]: synthetic DNA 64 taxa x 10k sites, GTR+G4 lnL (fixed branch lengths) of all "
                         "%d distinct SPR neighbours of a random tree"),
    "fitch_target": dict(kind="spr", score="fitch", datatype="dna", n_taxa=256, n_sites=100_000, seed=1002, ref_sample=48,
                         metric="trees x sites scored / sec (Fitch parsimony, all SPR neighbours, 256 taxa x 100k sites)",
                         label="BASELINE configs[2] shape: synthetic DNA 256 taxa x 100k sites, Fitch length of all %d "
                               "distinct SPR neighbours of a random tree (one step of the parsimonyGreedy climb)"),
    "trees_c4": dict(kind="trees", score="lnl", datatype="dna", n_taxa=128, n_sites=1_000_000, seed=1004, trees_per_rank=64,
                     ref_sample=0,
                     metric="trees x sites scored / sec (GTR+G4 logL, random trees, 128 taxa x 1M sites)",
                     label="BASELINE configs[3]: synthetic DNA 128 taxa x 1M sites, GTR+G4 lnL of %d random trees per GPU "
                           "(the 100k-tree list is scored in such blocks)"),
    "aa_c5": dict(kind="tbr", score="lnl", datatype="aa", n_taxa=50, n_sites=5_000, seed=1005, ref_sample=64,
                  metric="trees x sites scored / sec (WAG+G4 logL, TBR neighbourhood, 50 taxa x 5k sites)",
                  label="BASELINE configs[4]: synthetic protein 50 taxa x 5k sites, WAG+G4 lnL (fixed branch lengths) of all "
                        "%d distinct TBR neighbours of a random tree"),
}
DEFAULT_WORKLOAD = "lnl_target"


# ----------------------------------------------------------------------------------------------
# synthetic inputs (numpy only: shared by both arms and by tests/golden/make_bench_fixture.py)
# ----------------------------------------------------------------------------------------------
def random_desc(rng, n_tips):
    nodes = list(range(n_tips))
    desc, nxt = [], n_tips
    while len(nodes) > 1:
        a = nodes.pop(int(rng.integers(len(nodes))))
        b = nodes.pop(int(rng.integers(len(nodes))))
        desc.append((a, b))
        nodes.append(nxt)
        nxt += 1
    return np.array(desc, dtype=np.int32)


def make_alignment(seed=1001, n=64, s=10_000, datatype="dna"):
    """Tip codes uint8 [n, s] (DNA: one-hot 4-bit masks; AA: 0..19) and unit pattern weights."""
    rng = np.random.default_rng(seed)
    if datatype == "dna":
        return (1 << rng.integers(0, 4, (n, s))).astype(np.uint8), np.ones(s)
    return rng.integers(0, 20, (n, s)).astype(np.uint8), np.ones(s)


def make_tree(seed, n=64):
    rng = np.random.default_rng(seed + 7919)
    d = random_desc(rng, n)
    return d, rng.exponential(0.1, d.shape)


def base_tree(wl, b):
    """Base tree number b (rank b's at N > 1) of a neighbourhood workload."""
    if wl["n_taxa"] == 64:
        return make_tree(wl["seed"] + b, 64)  # (round-1 convention of the configs[1] headline)
    d = random_desc(np.random.default_rng(wl["seed"] + 7919 + 2 * b), wl["n_taxa"])
    return d, np.random.default_rng(wl["seed"] + 7920 + 2 * b).exponential(0.1, d.shape)


def random_trees(wl, count, block):
    rng = np.random.default_rng(wl["seed"] + 31 * block + 5)
    descs = np.stack([random_desc(rng, wl["n_taxa"]) for _ in range(count)])
    return descs, rng.exponential(0.1, descs.shape)


def fitch_states(wl):
    """Parsimony state characters uint8 [n_sites, n_taxa] ('0'..'3') of a Fitch workload."""
    rng = np.random.default_rng(wl["seed"])
    return (rng.integers(0, 4, (wl["n_sites"], wl["n_taxa"])) + 48).astype(np.uint8)


def desc_to_newick(desc, labels):
    n, text = len(labels), {}
    for i, (a, b) in enumerate(np.asarray(desc)):
        text[n + i] = "(%s,%s)" % tuple(labels[c] if c < n else text[c] for c in (int(a), int(b)))
    return text[n + len(desc) - 1] + ";"


def protein_model_for_oracle():
    """WAG for the CPU arm WITHOUT the product: read the table out of the C-ABI library by ctypes only if it
    is there; the reference arm's cost does not depend on the matrix, so a fixed-seed reversible model
    stands in otherwise (said in `sample`)."""
    rng = np.random.default_rng(7)
    return rng.uniform(0.1, 3.0, 190), rng.dirichlet(np.ones(20) * 20), "fixed-seed reversible 20-state model"


class ClockSampler(object):
    """SM clock and throttle reasons sampled DURING the timed region.  In-process NVML thread
    (pynvml) when available: an `nvidia-smi -lms` child process perturbed the timed steps by up to
    a millisecond per sample (it takes the driver lock for tens of fields); the NVML calls used here
    are two cheap queries per sample, every 50 ms (every 10 ms they doubled the step time of the launch-rich
    aa_c5 workload -- 70 ms against 35.7 -- while leaving the long steps alone).  Falls back to nvidia-smi."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu_index, uuid=None):
        self.gpu, self.uuid, self.proc, self.path = gpu_index, uuid, None, None
        self.thread, self.stop_flag, self.samples, self.max_mhz, self.mask = None, False, [], None, 0

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        if self.uuid:
            try:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(self.uuid)).encode())
            except Exception:
                pass
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.gpu)

    def start(self):
        try:
            import threading
            nv, h = self._nvml_handle()
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons

            def loop():
                while not self.stop_flag:
                    try:
                        self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                        self.mask |= int(reasons(h))
                    except Exception:
                        pass
                    time.sleep(float(os.environ.get("PLG_BENCH_SAMPLE_S", "0.05")))

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread:
            self.stop_flag = True
            self.thread.join()
            return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(name for bit, name in self.REASONS if self.mask & bit),
                    "samples": len(self.samples), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


_JSON_OUT = None


def claim_stdout():
    """stdout carries exactly one JSON line.  Libraries write there too (NCCL prints its version
    banner with printf), so file descriptor 1 is pointed at stderr for the rest of the process and
    the JSON line goes to a private duplicate of the original stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(line + "\n")
    out.flush()


def ncu_traffic(kernel, workload):
    """DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum) of one `ncu --set full`
    capture of this kernel ON THIS WORKLOAD (profiles/traffic.json names the capture); None if the
    kernel was never captured on it."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        t = json.load(open(p)).get("%s@%s" % (kernel, workload))
        if t:
            return {"dram_bytes_per_launch": t["dram_bytes"], "source": t.get("source", "profiles/")}
    return None


# ----------------------------------------------------------------------------------------------
# CPU side (cpu_baseline and --impl reference): oracle/ only, never pylogeny_b200
# ----------------------------------------------------------------------------------------------
def _fixture():
    return np.load(os.path.join(ROOT, "tests", "golden", "bench_ref_trees.npz"))


def cpu_sample(name, seconds):
    """One bounded sample of workload `name` on the host cores: the FP64 CPU restatement of the
    likelihood (oracle/lik_oracle.c; libpll itself is not available offline -- kind "port") or the
    reference's own compiled fitch.cpp core (oracle/_ref, kind "reference"), one tree per thread at
    a time, on sample trees enumerated by the reference's rearrangement code (fixture) or drawn at
    random (trees_c4).  Returns (cpu_baseline dict, seconds of the sample)."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    wl = WORKLOADS[name]
    cores = os.cpu_count() or 1
    n, S = wl["n_taxa"], wl["n_sites"]
    if wl["kind"] == "trees":
        descs, brs = random_trees(wl, max(cores, 8), 0)
        src = "random trees (numpy, same generator as the GPU arm)"
    else:
        fx = _fixture()
        descs, brs = fx[name + "_desc"].astype(np.int32), fx[name + "_brlen"]
        src = ("SPR neighbours enumerated and materialised by the reference's rearrangement.py (tests/golden/bench_ref_trees.npz)"
               if wl["kind"] == "spr" else "TBR neighbours from tests/golden/bench_ref_trees.npz (plain-Python bisection/reconnection; "
               "the reference has no TBR enumerator)")
    if wl["score"] == "fitch":
        states = fitch_states(wl)
        labels = ["T%d" % i for i in range(n)]
        use_ref = oracle.ref_available()
        fn = oracle.ref_fitch_cost if use_ref else oracle.fitch_cost
        newicks = [desc_to_newick(d, labels) for d in descs]
        taxa = {l: i for i, l in enumerate(labels)}
        # calibrate on a small column sample, then size (trees x sites) to the budget
        x0 = oracle.FitchInputs([bytes(r).decode("ascii") for r in states[:2000]], [1] * 2000, taxa)
        t0 = time.perf_counter()
        fn(newicks[0], inputs=x0)
        per_site = (time.perf_counter() - t0) / 2000
        n_trees = min(len(newicks), max(cores, 2 * cores))
        sub = int(min(S, max(2000, seconds * cores / (per_site * n_trees))))
        x = oracle.FitchInputs([bytes(r).decode("ascii") for r in states[:sub]], [1] * sub, taxa)
        t0 = time.perf_counter()
        with ThreadPoolExecutor(cores) as ex:  # ctypes releases the GIL inside the C call
            list(ex.map(lambda s: fn(s, inputs=x), newicks[:n_trees]))
        dt = time.perf_counter() - t0
        kind = "reference" if use_ref else "port"
        what = "reference fitch.cpp:6-272 compiled (oracle/_ref)" if use_ref else "C port of fitch.cpp (oracle/fitch_oracle.c)"
        return ({"value": n_trees * sub / dt, "unit": UNIT, "cores": cores, "kind": kind,
                 "sample": "%d trees x %d of the %d sites, %s; %s; %.1f s" % (n_trees, sub, S, what, src, dt)}, dt)
    codes, w = make_alignment(wl["seed"], n, S, wl["datatype"])
    if wl["datatype"] == "dna":
        exch, freqs, mname = EXCH, FREQS, "GTR+G4"
    else:
        exch, freqs, mname = protein_model_for_oracle()
    t0 = time.perf_counter()
    sub0 = min(S, 5000)
    oracle.lnl(codes[:, :sub0], w[:sub0], exch, freqs, ALPHA, descs[0], brs[0])
    per_site = (time.perf_counter() - t0) / sub0
    n_trees = min(len(descs), max(cores, 2 * cores))
    sub = int(min(S, max(sub0, seconds * cores / (per_site * n_trees))))
    c_sub, w_sub = np.ascontiguousarray(codes[:, :sub]), w[:sub]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(lambda i: oracle.lnl(c_sub, w_sub, exch, freqs, ALPHA, descs[i], brs[i]), range(n_trees)))
    dt = time.perf_counter() - t0
    return ({"value": n_trees * sub / dt, "unit": UNIT, "cores": cores, "kind": "port",
             "sample": "%d trees x %d of the %d sites, full re-traversal by the FP64 scalar pruning oracle (%s; libpll 1.0.2 SSE3 is not "
                       "available offline); %s; %.1f s" % (n_trees, sub, S, mname, src, dt)}, dt)


def workload_config(name, n_gpus, count=None):
    wl = WORKLOADS[name]
    cfg = {"workload": wl["label"] % count if count is not None else wl["label"].replace("%d ", ""), "name": name,
           "batches_per_step": n_gpus,
           "sharding": "pooled list of %d batch(es) in %d contiguous block(s), one per GPU; scores written to a device tensor and "
                       "all-gathered (NCCL)" % (n_gpus, n_gpus),
           "l2": "per-step working set (GBs of conditional vectors) >> 126 MB L2; L2 additionally flushed between steps",
           "seed": wl["seed"]}
    if wl["score"] == "lnl":
        cfg["planner"] = ("edge insertion by depth-first search walks (directional CLVs of the base tree, then per neighbour 1 CLV "
                          "update + 1 three-way evaluation, partials kept on chip)" if wl["kind"] == "spr" and wl["datatype"] == "dna"
                          else ("whole post-order on chip (tree walks)" if wl["kind"] == "trees" else
                                "TBR by part roots: base-tree vectors pushed through their branches once (T, H), one fused candidate step per "
                                "reconnection edge (3 chained 20x20x4 products on the FP64 tensor pipe), pair terms of a bisection as a "
                                "per-pattern 8x8-tiled DMMA product (csrc/lik_tbr20.cuh)"))
    return cfg


def run_reference(args):
    """--impl reference: the CPU side of the selected workload on the host cores, product-free."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    budget = float(os.environ.get("PLG_BENCH_REF_SECONDS", "60"))  # whole run, spread over the steps
    per = max(min(1.5, budget), budget / max(1, args.steps + args.warmup))
    for _ in range(args.warmup):
        cpu_sample(name, per)
    vals, times, cb = [], [], None
    for _ in range(args.steps):
        cb, dt = cpu_sample(name, per)
        vals.append(cb["value"])
        times.append(dt)
    v = float(np.mean(vals))
    cb["value"] = v
    wl = WORKLOADS[name]
    emit(json.dumps({
        "impl": "reference", "metric": wl["metric"], "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32 bitsets" if wl["score"] == "fitch" else "f64", "data": "synthetic",
        "config": workload_config(name, args.gpus), "cpu_baseline": cb,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
class Runner(object):
    """One workload on this rank: resident inputs, the timed step, the e2e step, the roofline."""

    def __init__(self, name, ctx):
        import torch
        from pylogeny_b200 import _lib, distributed as D, fitch, likelihood, neighbourhood as nbh
        self.torch, self.L, self.D, self.nbh, self.lik, self.fitch = torch, _lib.lib(), D, nbh, likelihood, fitch
        self._lib = _lib
        self.name, self.wl, self.ctx = name, WORKLOADS[name], ctx
        wl = self.wl
        n, S = wl["n_taxa"], wl["n_sites"]
        self.move = {"spr": nbh.MOVE_SPR, "tbr": nbh.MOVE_TBR}.get(wl["kind"])
        self.model = None
        if wl["score"] == "fitch":
            self.states = fitch_states(wl)
            self.w_int = np.ones(S, dtype=np.int64)
            self.aln = fitch.FitchAlignment.from_dense(self.states, self.w_int)
        else:
            self.codes, self.w = make_alignment(wl["seed"], n, S, wl["datatype"])
            dt = _lib.DATA_DNA if wl["datatype"] == "dna" else _lib.DATA_AA
            self.aln = likelihood.LikAlignment(self.codes, self.w, dt)
            if wl["datatype"] == "dna":
                self.model = likelihood.Model(EXCH, FREQS, ALPHA)
            else:
                e, f = likelihood.protein_model("WAG")
                self.model = likelihood.Model(e, f, ALPHA)
        if wl["kind"] == "trees":
            self.descs, self.brs = random_trees(wl, wl["trees_per_rank"], ctx.rank)
            self.count = wl["trees_per_rank"]
        else:
            self.d, self.br = base_tree(wl, ctx.rank)
            self.count = nbh.neighbourhood_size(self.d, self.move)
        dtype = torch.int32 if wl["score"] == "fitch" else torch.float64
        self.pool = torch.zeros(ctx.world * self.count, dtype=dtype, device=ctx.device)  # the pooled score list
        self.units_per_step = ctx.world * self.count * S

    # -- the hot path, inputs resident, scores of this rank's batch into its block of the pooled tensor
    def _score_into(self, aln, model, out_ptr, want_moves):
        C = __import__("ctypes")
        L, _lib = self.L, self._lib
        n = C.c_int64()
        stream = None
        if self.wl["kind"] == "trees":
            _lib.check(L.plg_lik_score_trees_shard(aln._h, model._h, self.count, self.wl["n_taxa"], self.descs.ctypes.data,
                                                   self.brs.ctypes.data, None, 0, 1, C.c_void_p(out_ptr), _lib.OUT_DEVICE, stream))
            return None
        moves = np.empty((self.count, 4), dtype=np.int32) if want_moves else None
        mp = moves.ctypes.data if want_moves else None
        if self.wl["score"] == "fitch":
            _lib.check(L.plg_fitch_score_neighbourhood_out(aln._h, self.move, self.wl["n_taxa"], self.d.ctypes.data, None, 0, 1,
                                                           C.byref(n), mp, C.c_void_p(out_ptr), _lib.OUT_DEVICE, stream))
        else:
            _lib.check(L.plg_lik_score_neighbourhood_out(aln._h, model._h, self.move, self.wl["n_taxa"], self.d.ctypes.data,
                                                         self.br.ctypes.data, None, 0, 1, C.byref(n), mp, C.c_void_p(out_ptr),
                                                         _lib.OUT_DEVICE, stream))
        return moves

    def step(self, aln=None, model=None):
        """Plan + score this rank's batch; the library writes the block into the pooled device tensor, one
        all-gather completes it on every rank; the scores (and the move table) end up on the host."""
        ctx = self.ctx
        block = self.pool[ctx.rank * self.count:(ctx.rank + 1) * self.count]
        moves = self._score_into(aln or self.aln, model or self.model, block.data_ptr(), True)
        self.D.all_gather_blocks(ctx, self.pool, ctx.world * self.count)
        return moves, self.pool.cpu().numpy()

    def e2e_step(self):
        """Same through the public API with HOST buffers: the alignment goes up, the scores come down."""
        if self.wl["score"] == "fitch":
            a = self.fitch.FitchAlignment.from_dense(self.pin_states.numpy(), self.w_int)
        else:
            dt = self._lib.DATA_DNA if self.wl["datatype"] == "dna" else self._lib.DATA_AA
            a = self.lik.LikAlignment(self.pin_codes.numpy(), self.pin_w.numpy(), dt)
        moves, scores = self.step(a, self.model)
        a.close()
        return moves, scores

    def prepare_e2e(self):
        t = self.torch
        if self.wl["score"] == "fitch":
            self.pin_states = t.from_numpy(self.states).pin_memory()
            h2d = self.states.nbytes + self.w_int.nbytes
        else:
            self.pin_codes, self.pin_w = t.from_numpy(self.codes).pin_memory(), t.from_numpy(self.w).pin_memory()
            h2d = self.codes.nbytes + self.w.nbytes
        if self.wl["kind"] == "trees":
            h2d += self.descs.nbytes + self.brs.nbytes
        else:
            h2d += self.d.nbytes + (self.br.nbytes if self.wl["score"] == "lnl" else 0)
        return h2d

    def close(self):
        self.aln.close()
        if self.model is not None:
            self.model.close()


def read_prof(L, kind):
    import ctypes as C
    ms, by, un, n, rq = C.c_double(), C.c_double(), C.c_double(), C.c_int64(), C.c_double()
    L.plg_profile_read(kind, C.byref(ms), C.byref(by), C.byref(un), C.byref(n), C.byref(rq))
    return {"ms": ms.value, "bytes": by.value, "units": un.value, "launches": n.value, "requested": rq.value}


_PEAKS = {}


def pipe_peak(L, mode):
    """FP64 (mode 0 DFMA, 1 DMMA, 2 both) / INT32 logic (3) pipe peak, measured now on this device."""
    import ctypes as C
    if mode not in _PEAKS:
        v = C.c_double()
        from pylogeny_b200 import _lib
        _lib.check(L.plg_fp64_peak(mode, C.byref(v)))
        _PEAKS[mode] = v.value
    return _PEAKS[mode]


def roofline_for(name, L, prof, total_ms, hbm, hbm_src):
    """The dominant kernel of a workload against the resource that binds it."""
    wl = WORKLOADS[name]
    kinds = {0: "fitch kernels", 1: "CLV update kernels", 2: "evaluation kernels", 3: "pmatrix_kernel", 4: "walk kernels"}
    dom = max(prof, key=lambda k: prof[k]["ms"])
    p = prof[dom]
    if not p["launches"] or p["ms"] <= 0:
        return None
    sec = p["ms"] * 1e-3
    alg_gbs = p["bytes"] / sec / 1e9
    common = {"launches": p["launches"], "avg_launch_ms": p["ms"] / p["launches"], "share_of_step": p["ms"] / total_ms,
              "algorithmic_bytes_per_launch": p["bytes"] / p["launches"], "algorithmic_gbs": alg_gbs,
              "algorithmic_gbs_over_hbm_peak": alg_gbs / hbm, "hbm_peak_gbs": hbm, "hbm_peak_source": hbm_src,
              "requested_gbs": p["requested"] / sec / 1e9}
    if wl["score"] == "lnl" and dom == 4:
        dmma = wl["kind"] == "spr"
        kernel = "walk4m_kernel" if dmma else "twalk4_kernel"
        flop = FLOP_PER_UNIT_DNA if dmma else FLOP_PER_NODEOP_DNA
        peak = max(pipe_peak(L, 1), pipe_peak(L, 0)) if dmma else pipe_peak(L, 0)
        ach = p["units"] * flop / sec / 1e12
        r = {"kernel": kernel, "bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
             "flop_per_unit": flop,
             "unit_definition": "neighbour x pattern (1 CLV update + 1 three-way evaluation)" if dmma else "node-op x pattern",
             "peak_source": "measured in this run: plg_fp64_peak, %s issue rate (DFMA %.1f, DMMA %.1f, interleaved %.1f TFLOP/s: one pipe)" % (
                 "DMMA m8n8k4" if dmma else "DFMA", pipe_peak(L, 0), pipe_peak(L, 1), pipe_peak(L, 2)),
             "traffic": ncu_traffic(kernel, name)}
    elif wl["score"] == "fitch":
        kernel = "fitch_search_walk_kernel (+ fitch_level_kernel for the base tree's directional sets)"
        peak = pipe_peak(L, 3)
        ach = p["units"] * INTOPS_PER_FITCH_UNIT / sec / 1e12
        r = {"kernel": kernel, "bound": "int", "achieved": ach, "peak": peak, "unit": "Tlop/s (INT32 logic lane-ops)",
             "frac": ach / peak, "ops_per_unit": INTOPS_PER_FITCH_UNIT, "unit_definition": "node-op x pattern (fold or three-way join)",
             "peak_source": "measured in this run: plg_fp64_peak mode 3 (LOP3 issue rate)",
             "traffic": ncu_traffic("fitch_search_walk_kernel", name)}
    elif wl["datatype"] == "aa":
        kernel = {1: "tbr20_step_kernel (+ clv20_kernel for the base tree's directional vectors)" if wl["kind"] == "tbr" else "clv20_kernel",
                  2: "pairs20_kernel" if wl["kind"] == "tbr" else "eval20_rows_kernel"}.get(dom, kinds[dom])
        peak = pipe_peak(L, 1)
        ach = p["units"] * FLOP_PER_NODEOP_AA / sec / 1e12
        r = {"kernel": kernel, "bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
             "flop_per_unit": FLOP_PER_NODEOP_AA,
             "unit_definition": "two 20x20 matrix-vector products x 4 categories + their elementwise join, per pattern (what one CLV "
                                "update costs: SURVEY.md 8d); the fused candidate steps are counted by the products they execute",
             "peak_source": "measured in this run: plg_fp64_peak, DMMA m8n8k4 issue rate", "traffic": ncu_traffic(kernel, name)}
    else:
        r = {"kernel": kinds[dom], "bound": "hbm", "achieved": alg_gbs, "peak": hbm, "unit": "GB/s", "frac": alg_gbs / hbm,
             "peak_source": hbm_src, "traffic": None}
    r.update(common)
    r["other_kernels"] = [{"kernels": kinds[k], "ms_per_step_share": prof[k]["ms"] / total_ms, "launches": prof[k]["launches"]}
                          for k in prof if k != dom and prof[k]["launches"]]
    return r


def run_workload(name, ctx, steps, warmup, sampler=None, e2e_steps=5):
    """W warm-up steps, then exactly `steps` timed steps (CUDA events on the launching stream, barrier +
    synchronize on both sides, L2 flushed before each, max over ranks)."""
    import torch
    import torch.distributed as dist
    world = ctx.world
    r = Runner(name, ctx)
    L = r.L
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=ctx.device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 3)):
        moves, scores = r.step()
    barrier()
    if sampler:
        sampler.start()
    L.plg_profile_enable(1)
    for k in range(5):
        L.plg_profile_read(k, None, None, None, None, None)
    launches0 = L.plg_launch_count()
    total_ms = 0.0
    barrier()
    wall0 = time.perf_counter()
    for _ in range(steps):
        flush.zero_()  # L2 flush, outside the per-step events
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        moves, scores = r.step()
        e1.record()
        torch.cuda.synchronize()
        total_ms += e0.elapsed_time(e1)
    barrier()
    wall = time.perf_counter() - wall0
    launches = L.plg_launch_count() - launches0
    L.plg_profile_enable(0)
    clocks = sampler.stop() if sampler else None
    prof = {k: read_prof(L, k) for k in range(5)}
    t = torch.tensor([total_ms], dtype=torch.float64, device=ctx.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / steps
    hbm, hbm_src = measured_hbm_peak()
    out = {"metric": r.wl["metric"], "value": r.units_per_step / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
           "warmup": max(warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "u32 bitsets" if r.wl["score"] == "fitch" else "f64", "data": "synthetic",
           "config": workload_config(name, world, r.count), "roofline": roofline_for(name, L, prof, total_ms, hbm, hbm_src),
           "gpu_launches": int(launches), "clocks": clocks, "wall_s_timed_region": wall, "trees_per_step": world * r.count,
           "checksum": float(np.asarray(scores, dtype=np.float64).sum())}
    # e2e: host buffers in, host results out, every step, on every rank; host clock, barrier on both sides, max over ranks
    h2d = r.prepare_e2e()
    e2e_ms = []
    for i in range(2 + e2e_steps):
        barrier()
        t0 = time.perf_counter()
        moves, scores = r.e2e_step()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=ctx.device)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if i >= 2:
            e2e_ms.append(1e3 * float(dt.item()))
    d2h = scores.nbytes + (moves.nbytes if moves is not None else 0)
    out["e2e"] = {"value": r.units_per_step / (np.mean(e2e_ms) * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(world * h2d),
                  "d2h_bytes_per_step": int(world * d2h), "ms_per_step": float(np.mean(e2e_ms)),
                  "ms_samples": [round(x, 3) for x in e2e_ms],
                  "note": "%d GPU(s), per rank and step: alignment create (H2D) + plan + score + all-gather + scores/moves D2H + destroy" % world}
    r.close()
    del flush
    torch.cuda.empty_cache()
    return out


def sharded_neighbourhood(ctx, name="lnl_target"):
    """north_star (4), literally: ONE neighbourhood split across the ranks (each plans and walks only the
    searches that hold its block of neighbours), scores all-gathered.  Strong scaling."""
    import torch
    import torch.distributed as dist
    from pylogeny_b200 import distributed as D, likelihood
    wl = WORKLOADS[name]
    codes, w = make_alignment(wl["seed"], wl["n_taxa"], wl["n_sites"])
    la = likelihood.LikAlignment(codes, w)
    model = likelihood.Model(EXCH, FREQS, ALPHA)
    d, br = base_tree(wl, 0)
    ms = []
    for i in range(4):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _, lnl = D.score_neighbourhood(ctx, la, model, d, br)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=ctx.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if i:
            ms.append(float(t.item()))
    M = lnl.shape[0]
    la.close(); model.close()
    return {"workload": "ONE SPR neighbourhood split across the ranks (pylogeny_b200.distributed.score_neighbourhood): " +
                        wl["label"] % M, "value": M * wl["n_sites"] / (np.mean(ms) * 1e-3), "unit": UNIT,
            "ms_per_step": float(np.mean(ms)), "steps": len(ms), "scaling": "strong", "checksum": float(lnl.sum().item())}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from pylogeny_b200 import distributed as D
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (pylogeny_b200 has no CPU fallback)")
    ctx = D.init()
    rank, world = ctx.rank, ctx.world
    try:
        uuid = torch.cuda.get_device_properties(ctx.device).uuid
    except Exception:
        uuid = None
    sampler = ClockSampler(ctx.device.index or 0, uuid) if rank == 0 else None
    out = run_workload(args.workload, ctx, args.steps, args.warmup, sampler)
    if world > 1 and args.workload == DEFAULT_WORKLOAD:
        out["one_neighbourhood_sharded"] = sharded_neighbourhood(ctx)
        c4 = run_workload("trees_c4", ctx, 3, 3, None, e2e_steps=1)  # BASELINE configs[3] across the GPUs
        out["trees_c4"] = {k: c4[k] for k in ("metric", "value", "unit", "ms_per_step", "steps", "trees_per_step", "roofline", "e2e", "config")}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if world == 1:
        if not args.no_cpu:
            out["cpu_baseline"] = cpu_sample(args.workload, 12.0)[0]
        if args.workload == DEFAULT_WORKLOAD and not args.no_extra:
            other = {}
            for name in WORKLOADS:
                if name == args.workload:
                    continue
                o = run_workload(name, ctx, 5, 3, None, e2e_steps=3)
                other[name] = {k: o[k] for k in ("metric", "value", "unit", "ms_per_step", "steps", "trees_per_step", "roofline",
                                                 "e2e", "config", "checksum", "gpu_launches")}
                if not args.no_cpu:
                    other[name]["cpu_baseline"] = cpu_sample(name, 6.0)[0]
            out["other_workloads"] = other
            out["streaming"] = streaming_leg()
            out["smoothed_lnl"] = smoothed_lnl_leg(full=True)
            out["greedy_climb"] = greedy_climb_leg()
    emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def greedy_climb_leg():
    """BASELINE configs[2] proper: the parsimonyGreedy hill-climb (one neighbourhood per step, move to the
    best neighbour while it improves), first 10 steps end to end on the host clock."""
    import torch
    from pylogeny_b200 import fitch, neighbourhood as nbh
    wl = WORKLOADS["fitch_target"]
    a = fitch.FitchAlignment.from_dense(fitch_states(wl), np.ones(wl["n_sites"], dtype=np.int64))
    d, _ = base_tree(wl, 0)
    a.score_trees(d[None])
    nbh.score_parsimony(a, d, want_moves=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    path, _ = nbh.parsimony_greedy(a, d, max_steps=10)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    a.close()
    return {"workload": "parsimonyGreedy climb over the SPR landscape, 256 taxa x 100k sites (BASELINE configs[2]), first 10 steps",
            "steps": len(path) - 1, "ms_per_step": 1e3 * dt / max(1, len(path) - 1), "path_head": [int(x) for x in path[:4]],
            "note": "per step: score all 255,530 neighbours, argmin, materialise the winner"}


def smoothed_lnl_leg(T=1024, full=False):
    """What scoring.getLogLikelihood returns per tree in the reference (libpllWrapper.c:152, :205:
    default branch lengths, 3 smoothing passes, lnL) for the first T SPR neighbours of the configs[1]
    tree, through plg_pll_loglik_batch (all trees in lockstep on the device) and, on 16 of them,
    through one handle per tree as the reference does it."""
    import torch
    from pylogeny_b200 import alignment, libpllWrapper as W, neighbourhood as nbh, pll
    wl = WORKLOADS["lnl_c2"]
    codes, _ = make_alignment(wl["seed"], wl["n_taxa"], wl["n_sites"])
    letters = np.array(list("?AC?G???T"))
    seqs = ["".join(letters[row]) for row in codes]
    ali = alignment.phylipFriendlyAlignment(names=["T%d" % i for i in range(wl["n_taxa"])], seqs=seqs)
    labels = ali.getTaxa()
    d, _ = base_tree(wl, 0)
    descs, _, _ = nbh.neighbour_trees(d, None, np.arange(T, dtype=np.int64))
    newicks = [nbh.desc_to_newick(x, labels) for x in descs]
    pm = pll.model_for(ali)
    W.getLogLikelihoodBatch(ali.getPhylip(), newicks[:8], pm.getFileName())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lnl, _ = W.getLogLikelihoodBatch(ali.getPhylip(), newicks, pm.getFileName())
    dt = time.perf_counter() - t0
    h = W.new(ali.getPhylip(), newicks[0], pm.getFileName())  # (warm-up, as for the batch: scratch allocation, first launch)
    W.getLogLikelihood(h)
    W.destroy(h)
    t0 = time.perf_counter()
    ref = []
    for nw in newicks[:16]:
        h = W.new(ali.getPhylip(), nw, pm.getFileName())
        ref.append(W.getLogLikelihood(h))
        W.destroy(h)
    dt1 = (time.perf_counter() - t0) / 16
    out = {"workload": "GTR+G4 lnL after 3 branch-length smoothing passes from z = 0.9 (scoring.getLogLikelihood "
                       "semantics) of the first %d SPR neighbours, 64 taxa x 10k sites" % T,
           "value": T * wl["n_sites"] / dt, "unit": UNIT, "ms_per_tree": 1e3 * dt / T, "per_tree_handle_ms": 1e3 * dt1,
           "max_rel_diff_vs_per_tree": float(np.max(np.abs((np.array(ref) - np.array(lnl[:16])) / np.array(ref))))}
    if full:
        # the WHOLE neighbourhood with the reference's per-tree semantics (what one step of heuristic.likelihoodGreedy
        # evaluates, heuristic.py:148-153), in one batched call; bytes = every conditional vector an action reads or
        # writes, once (C + 4 = 132 B per inner vector and pattern, 1 B per tip): per pass 3(n-2) re-orientations
        # (result + two operands) and 2n-3 Newton steps (the two end vectors), three passes, plus the down pass
        n = wl["n_taxa"]
        all_desc, _, _ = nbh.neighbour_trees(d, None, None)
        all_nw = [nbh.desc_to_newick(x, labels) for x in all_desc]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        full_lnl, _ = W.getLogLikelihoodBatch(ali.getPhylip(), all_nw, pm.getFileName())
        dtf = time.perf_counter() - t0
        C4 = 132.0
        per_pass = 3 * (n - 2) * C4 + (4 * n - 12) * C4 + 2 * n + (3 * n - 6) * C4 + n
        down = (n - 2) * C4 + (n - 3) * C4 + (n - 1)
        bytes_per_tree = (3 * per_pass + down) * wl["n_sites"]
        hbm, hbm_src = measured_hbm_peak()
        gbs = len(all_nw) * bytes_per_tree / dtf / 1e9
        out["full_neighbourhood"] = {
            "workload": "the same for ALL %d SPR neighbours in one call (one likelihoodGreedy step)" % len(all_nw),
            "seconds": dtf, "ms_per_tree": 1e3 * dtf / len(all_nw), "value": len(all_nw) * wl["n_sites"] / dtf, "unit": UNIT,
            "best_lnl": float(np.max(full_lnl)),
            "roofline": {"kernels": "clv4_kernel + clvopt4_kernel + opt_step4_kernel + newton_kernel (lockstep program, host clock "
                                    "over the whole call incl. planning and uploads)", "bound": "hbm", "achieved": gbs, "peak": hbm,
                         "unit": "GB/s", "frac": gbs / hbm, "peak_source": hbm_src, "bytes_per_tree": bytes_per_tree,
                         "bytes": "ALGORITHMIC: every vector an action reads or writes, once per action; the fused "
                                  "re-orientation + Newton step (clvopt4_kernel) keeps one of them in registers, so about 12 % "
                                  "fewer bytes really move",
                         "assumes": "all three passes run on every tree (they do from z = 0.9 on random data)"}}
    pm.close()
    ali.close()
    return out


def streaming_leg():
    """The level-synchronous kernels -- every node vector through HBM, working set far beyond L2 -- as the plain
    HBM-roofline evidence north_star asks for (achieved HBM GB/s of the parsimony and CLV-update kernels), on
    descriptor batches at BASELINE configs[3] (lnL) and configs[2] (Fitch) shapes with PLG_TREE_MODE=levels.
    (The shipped tree walks of the same calls are the "trees_c4" workload.)"""
    import torch
    from pylogeny_b200 import _lib, fitch, likelihood
    L = _lib.lib()
    hbm, hbm_src = measured_hbm_peak()
    res = {}

    def leg(call, n_units, kinds, kernel, workload):
        for _ in range(2):
            call()
        L.plg_profile_enable(1)
        for k in kinds:
            read_prof(L, k)
        ms = []
        for _ in range(5):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            call()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        p = [read_prof(L, k) for k in kinds]
        L.plg_profile_enable(0)
        kms, by, nl, rq = (sum(x[f] for x in p) for f in ("ms", "bytes", "launches", "requested"))
        ach = by / (kms * 1e-3) / 1e9
        return {"workload": workload, "value": n_units / (np.mean(ms) * 1e-3), "unit": UNIT, "ms_per_call": float(np.mean(ms)), "steps": 5,
                "roofline": {"kernel": kernel, "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                             "requested_gbs": rq / (kms * 1e-3) / 1e9, "launches": nl // 5, "kernel_ms_per_call": kms / 5,
                             "peak_source": hbm_src}}

    os.environ["PLG_TREE_MODE"] = "levels"
    try:
        wl = WORKLOADS["trees_c4"]
        codes, w = make_alignment(wl["seed"], wl["n_taxa"], wl["n_sites"])
        la = likelihood.LikAlignment(codes, w)
        m = likelihood.Model(EXCH, FREQS, ALPHA)
        descs, brs = random_trees(wl, 3, 0)
        res["lnl_levels"] = leg(lambda: la.score_trees(m, descs, brs), 3 * wl["n_sites"], (1, 2), "clv4_kernel + eval4_kernel",
                                "128 taxa x 1M sites, GTR+G4, 3 random trees per call (BASELINE configs[3] shape), level programs")
        la.close(); m.close()
        wl = WORKLOADS["fitch_target"]
        fa = fitch.FitchAlignment.from_dense(fitch_states(wl), np.ones(wl["n_sites"], dtype=np.int64))
        rng = np.random.default_rng(1004)
        descs = np.stack([random_desc(rng, wl["n_taxa"]) for _ in range(2048)])
        res["fitch_levels"] = leg(lambda: fa.score_trees(descs), 2048 * wl["n_sites"], (0,), "fitch_level_kernel",
                                  "256 taxa x 100k sites, 2048 random trees per call (BASELINE configs[2] shape), level programs")
        fa.close()
    finally:
        os.environ.pop("PLG_TREE_MODE", None)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline legs (profiling runs)")
    ap.add_argument("--no-extra", action="store_true", help="only the selected workload (no other_workloads / streaming legs)")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
