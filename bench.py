#!/usr/bin/env python
"""Benchmark of the tree-scoring hot path (BASELINE.json metric: trees x sites scored / second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload at every N: BASELINE.json configs[1] -- synthetic DNA, 64 taxa x 10,000 sites (i.i.d.
uniform ACGT, seed 1001, so patterns = sites), GTR+Gamma4 log-likelihood at fixed branch
lengths of every distinct SPR neighbour (14,762) of a random tree.  One "step" = one pass of
the hot path over that batch: plan (host) + score (device) + read back the per-tree scores.
At N > 1 the job is the pooled neighbour-tree list of N such neighbourhoods (base trees seeded
1001+b): N x 14,762 trees, sharded across the ranks in contiguous blocks (rank r scores block r
= the neighbours of base tree r; full alignment on every GPU, no data-path collective) and the
per-tree scores come back through one NCCL all-gather -- weak scaling.  Splitting ONE
neighbourhood across ranks (shard=(rank, world) in the C-ABI: each rank plans and walks only its
block of neighbours) is measured too at N > 1, on the north_star target shape (256 taxa x 100k
sites, all 255,530 SPR neighbours): the "one_neighbourhood_sharded" object, strong scaling.

One JSON line on stdout (rank 0).  Extra objects: "roofline" (dominant kernel, CUDA-event timed
inside the timed region), "cpu_baseline" (CPU oracle on a bounded sample of the same workload),
"e2e" (same metric through the public API with host buffers, H2D of the alignment and D2H of
the scores inside the timed region), "parsimony" (BASELINE config 3 shape, 256 taxa x 100k
sites, Fitch scoring of all 255,530 SPR neighbours; N=1 only).
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_TAXA, N_SITES, SEED = 64, 10_000, 1001
EXCH = np.array([1.2, 3.4, 0.8, 1.1, 4.1, 1.0])
FREQS = np.array([0.28, 0.22, 0.24, 0.26])
ALPHA = 0.7
METRIC = "trees x sites scored / sec (GTR+G4 logL, all SPR neighbours, 64 taxa x 10k sites)"
UNIT = "tree-sites/s"


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


def make_alignment(seed=SEED, n=N_TAXA, s=N_SITES):
    rng = np.random.default_rng(seed)
    codes = (1 << rng.integers(0, 4, (n, s))).astype(np.uint8)
    return codes, np.ones(s)


def make_tree(seed, n=N_TAXA):
    rng = np.random.default_rng(seed + 7919)
    d = random_desc(rng, n)
    return d, rng.exponential(0.1, d.shape)


class ClockSampler(object):
    """SM clock and throttle reasons sampled DURING the timed region.  In-process NVML thread
    (pynvml) when available: an `nvidia-smi -lms` child process perturbed the timed steps by up to
    a millisecond per sample (it takes the driver lock for tens of fields); the NVML calls used here
    are two cheap queries per sample.  Falls back to nvidia-smi."""
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
                    time.sleep(0.01)

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


def measured_peaks():
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


def ncu_traffic(kernel, units_per_launch):
    """DRAM bytes (ncu dram__bytes_read+write) of one captured launch, scaled per node-op x pattern
    unit to this run's average launch -- 'per launch like achieved'.  None if never captured."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        t = json.load(open(p)).get(kernel)
        if t and t.get("units"):
            return t["dram_bytes"] / t["units"] * units_per_launch
    return None


# ----------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py touches oracle/)
# ----------------------------------------------------------------------------------------------
def cpu_lnl_sample(seconds=12.0):
    """FP64 CPU restatement (oracle/lik_oracle.c; libpll itself is not available) on a bounded
    sample of the workload's neighbour trees, all host cores (one tree per thread at a time)."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    from pylogeny_b200 import neighbourhood as nbh
    codes, w = make_alignment()
    d, br = make_tree(SEED)
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    descs, brs, _ = nbh.neighbour_trees(d, br, np.arange(1))
    oracle.lnl(codes, w, EXCH, FREQS, ALPHA, descs[0], brs[0])
    t1 = time.perf_counter() - t0
    n_sample = int(max(cores, min(4096, seconds * cores / max(t1, 1e-3))))
    n_sample -= n_sample % cores
    sel = np.linspace(0, nbh.neighbourhood_size(d) - 1, n_sample).astype(np.int64)
    descs, brs, _ = nbh.neighbour_trees(d, br, sel)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:  # ctypes releases the GIL inside the C call
        out = list(ex.map(lambda i: oracle.lnl(codes, w, EXCH, FREQS, ALPHA, descs[i], brs[i]), range(n_sample)))
    dt = time.perf_counter() - t0
    return {"value": n_sample * N_SITES / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d of the 14762 neighbour trees x all 10000 sites, FP64 scalar pruning oracle "
                      "(libpll 1.0.2 SSE3 is not available offline), %.1f s" % (n_sample, dt)}, dt, out


def cpu_fitch_sample(states, desc, seconds=10.0):
    """The reference's own compiled fitch.cpp core (oracle/_ref) when present, else the C port."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    from pylogeny_b200 import neighbourhood as nbh
    cores = os.cpu_count() or 1
    n, S = states.shape[1], states.shape[0]
    sub, per_core = 25000, 4
    profs = [bytes(r).decode("ascii") for r in states[:sub]]
    labels = ["T%d" % i for i in range(n)]
    x = oracle.FitchInputs(profs, [1] * sub, {l: i for i, l in enumerate(labels)})
    use_ref = oracle.ref_available()
    fn = oracle.ref_fitch_cost if use_ref else oracle.fitch_cost
    descs, _, _ = nbh.neighbour_trees(desc, None, np.linspace(0, nbh.neighbourhood_size(desc) - 1, cores * per_core).astype(np.int64))
    nws = [nbh.desc_to_newick(dd, labels) for dd in descs]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(lambda s: fn(s, inputs=x), nws))
    dt = time.perf_counter() - t0
    return {"value": cores * per_core * sub / dt, "unit": UNIT, "cores": cores,
            "kind": "reference" if use_ref else "port",
            "sample": "%d neighbour trees x %d of the %d sites, %s, %.1f s" % (
                cores * per_core, sub, S, "reference fitch.cpp:6-272 compiled (oracle/_ref)" if use_ref else "C port", dt)}


def run_reference(args):
    """--impl reference: the reference-side CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for _ in range(args.warmup):
        cpu_lnl_sample(seconds=1.0)
    vals, times = [], []
    for _ in range(args.steps):
        budget = float(os.environ.get("PLG_BENCH_REF_SECONDS", "40"))  # whole run, spread over the steps
        cb, dt, _ = cpu_lnl_sample(seconds=max(min(2.0, budget), budget / args.steps))
        vals.append(cb["value"])
        times.append(dt)
    v = float(np.mean(vals))
    cb["value"] = v
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus), "cpu_baseline": cb,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(n_gpus):
    return {"workload": "BASELINE configs[1]: synthetic DNA 64 taxa x 10k sites, GTR+G4 lnL (fixed branch lengths) "
                        "of all 14762 distinct SPR neighbours of a random tree",
            "planner": "edge insertion by depth-first search walks (directional CLVs of the base tree, then per "
                       "neighbour 1 CLV update + 1 three-way evaluation, partials kept on chip)",
            "neighbourhoods_per_step": n_gpus,
            "sharding": "pooled list of %d x 14762 neighbour trees in %d contiguous blocks, scores all-gathered (NCCL)" % (
                n_gpus, n_gpus),
            "l2": "per-step working set (~19 GB of CLVs) >> 126 MB L2; L2 additionally flushed between steps",
            "seed": SEED}


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from pylogeny_b200 import _lib, fitch, likelihood, neighbourhood as nbh
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (pylogeny_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    _lib.check(_lib.lib().plg_set_device(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.lib()
    dev = torch.device("cuda", local)
    codes, w = make_alignment()
    trees = [make_tree(SEED + b) for b in range(world)]
    model = likelihood.Model(EXCH, FREQS, ALPHA)
    aln = likelihood.LikAlignment(codes, w)
    M = nbh.neighbourhood_size(trees[0][0])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gathered = torch.empty(world * M, dtype=torch.float64, device=dev) if world > 1 else None

    my_d, my_br = trees[rank]

    def step():
        _, lnl = nbh.score_likelihood(aln, model, my_d, my_br)
        if world > 1:  # tiny all-gather of the per-tree scores over NCCL
            dist.all_gather_into_tensor(gathered, torch.from_numpy(lnl).to(dev, non_blocking=True))
            return gathered
        return lnl

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    try:
        uuid = torch.cuda.get_device_properties(local).uuid
    except Exception:
        uuid = None
    sampler = ClockSampler(local, uuid)
    if rank == 0:
        sampler.start()
    L.plg_profile_enable(1)
    for k in range(5):
        L.plg_profile_read(k, None, None, None, None, None)
    launches0 = L.plg_launch_count()
    total_ms = 0.0
    barrier()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()  # L2 flush, outside the per-step events
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        res = step()
        e1.record()
        torch.cuda.synchronize()
        total_ms += e0.elapsed_time(e1)
    barrier()
    wall = time.perf_counter() - wall0
    launches = L.plg_launch_count() - launches0
    L.plg_profile_enable(0)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * M * N_SITES / (ms_per_step * 1e-3)

    # per-kernel roofline (events recorded by the library around every launch of the timed steps)
    import ctypes as C
    prof = {}
    for k, name in enumerate(("fitch_level_kernel", "clv4_kernel", "eval4_kernel", "pmatrix_kernel", "walk4m_kernel")):
        ms, by, un, n, rq = C.c_double(), C.c_double(), C.c_double(), C.c_int64(), C.c_double()
        L.plg_profile_read(k, C.byref(ms), C.byref(by), C.byref(un), C.byref(n), C.byref(rq))
        prof[name] = {"ms": ms.value, "bytes": by.value, "units": un.value, "launches": n.value, "requested": rq.value}
    peak, peak_src = measured_peaks()
    dom = max(("walk4m_kernel", "clv4_kernel", "eval4_kernel"), key=lambda k: prof[k]["ms"])

    def roof(name):
        p = prof[name]
        if not p["launches"] or p["ms"] <= 0:
            return None
        ach = p["bytes"] / (p["ms"] * 1e-3) / 1e9
        return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": ncu_traffic(name, p["units"] / p["launches"]), "peak_source": peak_src,
                "launches": p["launches"],
                "avg_launch_ms": p["ms"] / p["launches"], "algorithmic_bytes_per_launch": p["bytes"] / p["launches"],
                "requested_bytes_per_launch": p["requested"] / p["launches"],
                "requested_gbs": p["requested"] / (p["ms"] * 1e-3) / 1e9,
                "share_of_step": p["ms"] / total_ms}

    roofline = roof(dom)
    # FP64 view of the same launches: ~640 flop per neighbour x pattern (CLV update 240, SURVEY.md 8d;
    # 3-way evaluation ~400) against the DMMA peak measured on this pool (profiles/r1c_fp64_peak.txt)
    if dom == "walk4m_kernel" and prof[dom]["ms"] > 0:
        tf = prof[dom]["units"] * 640.0 / (prof[dom]["ms"] * 1e-3) / 1e12
        roofline["fp64"] = {"achieved_tflops": tf, "peak_tflops": 37.1, "frac": tf / 37.1,
                            "peak_source": "measured DMMA m8n8k4 (profiles/micro/fp64_peak.cu)",
                            "flop_per_unit": 640}
    roofline["other_kernels"] = [r for r in (roof(k) for k in prof if k != dom) if r]
    # e2e: same metric through the public API with HOST buffers; the alignment goes up and the
    # scores come down inside the timed region, every step -- on every rank (its own neighbourhood),
    # host clock, barrier on both sides, max over ranks
    pinned_codes = torch.from_numpy(codes).pin_memory()
    pinned_w = torch.from_numpy(w).pin_memory()
    e2e_ms = []
    for i in range(2 + min(args.steps, 5)):
        barrier()
        t0 = time.perf_counter()
        a2 = likelihood.LikAlignment(pinned_codes.numpy(), pinned_w.numpy())
        moves, lnl = nbh.score_likelihood(a2, model, my_d, my_br)
        a2.close()
        if world > 1:
            dist.all_gather_into_tensor(gathered, torch.from_numpy(lnl).to(dev, non_blocking=True))
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if i >= 2:
            e2e_ms.append(1e3 * float(dt.item()))
    h2d = world * (codes.nbytes + w.nbytes + my_d.nbytes + my_br.nbytes)
    e2e = {"value": world * M * N_SITES / (np.mean(e2e_ms) * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(world * (lnl.nbytes + moves.nbytes)), "ms_per_step": float(np.mean(e2e_ms)),
           "ms_samples": [round(x, 3) for x in e2e_ms],
           "note": "%d GPU(s), per rank and step: plg_lik_aln_create + plg_lik_score_neighbourhood + destroy" % world}
    sharded = None
    if world > 1:
        # north_star (4), literally: ONE neighbourhood -- the target shape, 256 taxa x 100k sites, all
        # 255,530 SPR neighbours by GTR+G4 -- sharded across the ranks (each plans and walks only the
        # searches that hold its block of neighbours), scores combined over NCCL.  Strong scaling.
        n2, s2 = 256, 100_000
        rng2 = np.random.default_rng(1002)
        codes2 = (1 << rng2.integers(0, 4, (n2, s2))).astype(np.uint8)
        la2 = likelihood.LikAlignment(codes2, np.ones(s2))
        d2 = random_desc(np.random.default_rng(1002 + 7919), n2)
        br2 = np.random.default_rng(1002 + 7920).exponential(0.1, d2.shape)
        M2 = nbh.neighbourhood_size(d2)
        buf = torch.zeros(M2, dtype=torch.float64, device=dev)

        def shard_step():
            _, part = nbh.score_likelihood(la2, model, d2, br2, shard=(rank, world), want_moves=False)
            lo, hi = M2 * rank // world, M2 * (rank + 1) // world
            buf.zero_()
            buf[lo:hi] = torch.from_numpy(part[lo:hi]).to(dev)
            dist.all_reduce(buf)  # disjoint blocks: the sum is the concatenation
            return buf

        shard_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res2 = shard_step()
        e1.record()
        torch.cuda.synchronize()
        t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        sharded = {"workload": "ONE SPR neighbourhood sharded across the ranks: synthetic DNA 256 taxa x 100k sites, "
                               "GTR+G4 lnL of all %d neighbours" % M2,
                   "value": M2 * s2 / (float(t2.item()) * 1e-3), "unit": UNIT, "ms_per_step": float(t2.item()),
                   "scaling": "strong", "checksum": float(res2.sum().item())}
        la2.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rank 0 only from here: e2e, CPU baseline, parsimony leg --------------------------------
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(world),
           "roofline": roofline, "gpu_launches": int(launches), "clocks": clocks,
           "wall_s_timed_region": wall, "trees_per_step": world * M}
    if sharded:
        out["one_neighbourhood_sharded"] = sharded

    out["e2e"] = e2e
    if world == 1:
        if not args.no_cpu:
            out["cpu_baseline"] = cpu_lnl_sample()[0]
        out["parsimony"] = parsimony_leg(fitch, nbh, L, peak, peak_src, not args.no_cpu)
        out["streaming"] = streaming_leg(fitch, likelihood, L, peak, peak_src)
        out["target_shape_lnl"] = target_lnl_leg(likelihood, nbh, L, peak, peak_src)
        out["smoothed_lnl"] = smoothed_lnl_leg(nbh)
    emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def parsimony_leg(fitch, nbh, L, peak, peak_src, cpu=True):
    """BASELINE config 3 shape: 256 taxa x 100k sites, Fitch length of all 255,530 SPR neighbours."""
    import ctypes as C
    import torch
    n, S = 256, 100_000
    rng = np.random.default_rng(1002)
    states = (rng.integers(0, 4, (S, n)) + 48).astype(np.uint8)
    a = fitch.FitchAlignment.from_dense(states, np.ones(S, dtype=np.int64))
    d = random_desc(np.random.default_rng(1002 + 7919), n)
    for _ in range(2):
        moves, costs = nbh.score_parsimony(a, d)
    L.plg_profile_enable(1)
    L.plg_profile_read(0, None, None, None, None, None)
    ms = []
    for _ in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        moves, costs = nbh.score_parsimony(a, d)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    L.plg_profile_enable(0)
    kms, by, un, nl, rq = C.c_double(), C.c_double(), C.c_double(), C.c_int64(), C.c_double()
    L.plg_profile_read(0, C.byref(kms), C.byref(by), C.byref(un), C.byref(nl), C.byref(rq))
    M = len(costs)
    ach = by.value / (kms.value * 1e-3) / 1e9
    res = {"workload": "BASELINE configs[2] shape: synthetic DNA 256 taxa x 100k sites, Fitch length of all %d "
                       "distinct SPR neighbours of a random tree (edge insertion)" % M,
           "value": M * S / (np.mean(ms) * 1e-3), "unit": UNIT, "ms_per_step": float(np.mean(ms)), "dtype": "u32 bitsets",
           "roofline": {"kernel": "fitch_search_walk_kernel (+ fitch_level_kernel for the base tree's directional sets)",
                        "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                        "frac": ach / peak, "traffic": ncu_traffic("fitch_search_walk_kernel", un.value / max(nl.value, 1)),
                        "algorithmic_bytes_per_launch": by.value / max(nl.value, 1),
                        "requested_gbs": rq.value / (kms.value * 1e-3) / 1e9, "peak_source": peak_src,
                        "launches": nl.value, "avg_launch_ms": kms.value / max(nl.value, 1),
                        "share_of_step": kms.value / sum(ms)},
           "checksum": int(costs.astype(np.int64).sum()), "min_cost": int(costs.min())}
    # BASELINE configs[2] proper: the greedy hill-climb (one neighbourhood per step, move to the best
    # neighbour while it improves), first 10 steps timed end to end on the host clock
    a.score_trees(d[None])  # (the climb scores its start tree once: first use of the tree-walk path allocates)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    path, _ = nbh.parsimony_greedy(a, d, max_steps=10)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res["greedy_climb"] = {"steps": len(path) - 1, "ms_per_step": 1e3 * dt / max(1, len(path) - 1),
                           "path_head": [int(x) for x in path[:4]],
                           "note": "parsimony_greedy: score all neighbours, argmin, materialise the winner"}
    if cpu:
        res["cpu_baseline"] = cpu_fitch_sample(states, d)
    a.close()
    return res


def target_lnl_leg(likelihood, nbh, L, peak, peak_src):
    """north_star target shape for likelihood: GTR+G4 lnL of all 255,530 SPR neighbours of a
    256-taxon x 100k-site tree (the parsimony leg is the same shape).  One call, ~1 s."""
    import ctypes as C
    import torch
    n, S = 256, 100_000
    rng = np.random.default_rng(1002)
    codes = (1 << rng.integers(0, 4, (n, S))).astype(np.uint8)
    la = likelihood.LikAlignment(codes, np.ones(S))
    m = likelihood.Model(EXCH, FREQS, ALPHA)
    d = random_desc(np.random.default_rng(1002 + 7919), n)
    br = np.random.default_rng(1002 + 7920).exponential(0.1, d.shape)
    nbh.score_likelihood(la, m, d, br, want_moves=False)
    L.plg_profile_enable(1)
    L.plg_profile_read(4, None, None, None, None, None)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, lnl = nbh.score_likelihood(la, m, d, br, want_moves=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    kms, by, un, nl, rq = C.c_double(), C.c_double(), C.c_double(), C.c_int64(), C.c_double()
    L.plg_profile_read(4, C.byref(kms), C.byref(by), C.byref(un), C.byref(nl), C.byref(rq))
    L.plg_profile_enable(0)
    la.close()
    ach = by.value / (kms.value * 1e-3) / 1e9
    return {"workload": "synthetic DNA 256 taxa x 100k sites, GTR+G4 lnL (fixed branch lengths) of all %d distinct "
                        "SPR neighbours of a random tree" % len(lnl),
            "value": len(lnl) * S / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "checksum": float(lnl.sum()),
            "roofline": {"kernel": "walk4m_kernel", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "requested_gbs": rq.value / (kms.value * 1e-3) / 1e9,
                         "launches": nl.value, "avg_launch_ms": kms.value / max(nl.value, 1), "peak_source": peak_src,
                         "fp64_tflops": un.value * 640.0 / (kms.value * 1e-3) / 1e12}}


def smoothed_lnl_leg(nbh, T=1024):
    """What scoring.getLogLikelihood returns per tree in the reference (libpllWrapper.c:152, :205:
    default branch lengths, 3 smoothing passes, lnL) for the first T SPR neighbours of the bench
    tree, through plg_pll_loglik_batch (all trees in lockstep on the device) and, on 8 of them,
    through one handle per tree as the reference does it."""
    import torch
    from pylogeny_b200 import alignment, libpllWrapper as W, pll
    codes, _ = make_alignment()
    letters = np.array(list("?AC?G???T"))
    seqs = ["".join(letters[row]) for row in codes]
    ali = alignment.phylipFriendlyAlignment(names=["T%d" % i for i in range(N_TAXA)], seqs=seqs)
    labels = ali.getTaxa()
    d, _ = make_tree(SEED)
    descs, _, _ = nbh.neighbour_trees(d, None, np.arange(T, dtype=np.int64))
    newicks = [nbh.desc_to_newick(x, labels) for x in descs]
    pm = pll.partitionModel(ali)
    pm.createSimpleModel()
    W.getLogLikelihoodBatch(ali.getPhylip(), newicks[:8], pm.getFileName())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lnl, _ = W.getLogLikelihoodBatch(ali.getPhylip(), newicks, pm.getFileName())
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    ref = []
    for nw in newicks[:8]:
        h = W.new(ali.getPhylip(), nw, pm.getFileName())
        ref.append(W.getLogLikelihood(h))
        W.destroy(h)
    dt1 = (time.perf_counter() - t0) / 8
    pm.close()
    ali.close()
    return {"workload": "GTR+G4 lnL after 3 branch-length smoothing passes from z = 0.9 (scoring.getLogLikelihood "
                        "semantics) of the first %d SPR neighbours, 64 taxa x 10k sites" % T,
            "value": T * N_SITES / dt, "unit": UNIT, "ms_per_tree": 1e3 * dt / T,
            "per_tree_handle_ms": 1e3 * dt1,
            "max_rel_diff_vs_per_tree": float(np.max(np.abs((np.array(ref) - np.array(lnl[:8])) / np.array(ref))))}


def streaming_leg(fitch, likelihood, L, peak, peak_src):
    """Descriptor batches (plg_*_score_trees) at BASELINE config-4 shape for lnL (128 taxa x 1M sites)
    and config-3 shape for Fitch (256 taxa x 100k sites, 2048 random trees).  Two formulations each:
    `tree_walk` (shipped: whole post-order on chip, HBM only supplies the tips -- FP64 / integer
    pipe bound, algorithmic-bytes fraction > 1) and `levels` (PLG_TREE_MODE=levels: one launch per
    tree level, every node vector through HBM -- the plain HBM-roofline evidence for the streaming
    kernels, working set far beyond L2)."""
    import ctypes as C
    import torch
    res = {}

    def read(kind):
        ms, by, un, nl, rq = C.c_double(), C.c_double(), C.c_double(), C.c_int64(), C.c_double()
        L.plg_profile_read(kind, C.byref(ms), C.byref(by), C.byref(un), C.byref(nl), C.byref(rq))
        return ms.value, by.value, nl.value, rq.value

    def leg(call, n_units, kinds, kernel, workload):
        for _ in range(2):
            call()
        L.plg_profile_enable(1)
        for k in kinds:
            read(k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        kms = by = nl = rq = 0
        for k in kinds:
            a, b, c, d = read(k)
            kms += a; by += b; nl += c; rq += d
        L.plg_profile_enable(0)
        ach = by / (kms * 1e-3) / 1e9
        return {"workload": workload, "value": n_units / dt, "unit": UNIT, "ms_per_call": dt * 1e3,
                "roofline": {"kernel": kernel, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                             "frac": ach / peak, "requested_gbs": rq / (kms * 1e-3) / 1e9, "launches": nl // 3,
                             "kernel_ms_per_call": kms / 3, "peak_source": peak_src}}

    rng = np.random.default_rng(1004)
    n, S = 128, 1_000_000
    codes = (1 << rng.integers(0, 4, (n, S))).astype(np.uint8)
    la = likelihood.LikAlignment(codes, np.ones(S))
    m = likelihood.Model(EXCH, FREQS, ALPHA)
    descs = np.stack([random_desc(rng, n) for _ in range(16)])
    brs = rng.exponential(0.1, descs.shape)
    for mode, T, kinds, kernel in (("walk", 16, (4,), "twalk4_kernel"), ("levels", 3, (1, 2), "clv4_kernel + eval4_kernel")):
        os.environ["PLG_TREE_MODE"] = mode
        res["lnl_" + ("tree_walk" if mode == "walk" else "levels")] = leg(
            lambda: la.score_trees(m, descs[:T], brs[:T]), T * S, kinds, kernel,
            "128 taxa x 1M sites, GTR+G4, %d random trees per call (BASELINE configs[3] shape)" % T)
    la.close()
    n, S, T = 256, 100_000, 2048
    states = (rng.integers(0, 4, (S, n)) + 48).astype(np.uint8)
    fa = fitch.FitchAlignment.from_dense(states, np.ones(S, dtype=np.int64))
    descs = np.stack([random_desc(rng, n) for _ in range(T)])
    for mode, kernel in (("walk", "fitch_walk_kernel"), ("levels", "fitch_level_kernel")):
        os.environ["PLG_TREE_MODE"] = mode
        res["fitch_" + ("tree_walk" if mode == "walk" else "levels")] = leg(
            lambda: fa.score_trees(descs), T * S, (0,), kernel,
            "256 taxa x 100k sites, %d random trees per call (BASELINE configs[2] shape)" % T)
    os.environ.pop("PLG_TREE_MODE", None)
    fa.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline legs (profiling runs)")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
