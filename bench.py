#!/usr/bin/env python
"""Headline benchmark: VBR SpMM effective TFLOP/s on nonzero-block FLOPs (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # our sm_100a path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU VBR::multiply

A "step" is one pass of the hot path (C = A*B, one kernel launch) over the workload.  The
default workload is BASELINE config #3: a seeded R-MAT 65536 x 65536 matrix (density 1e-3,
pattern-only), blocked with the reference's `-a 5 -b 64 -B 64 -t 0.6` row clustering into VBR,
times a dense 65536 x 2048 B, bf16 operands, fp32 accumulation.

  value      whole-job TFLOP/s = 2 * nztot * n * K / t, operands resident in HBM, t = CUDA-event
             time of K back-to-back launches on the launching stream (max over ranks).
  e2e        the same metric through the one-shot reference-facing C-ABI call
             (sparta_vbr_spmm: host VBR arrays + host B in, host C out; every H2D/D2H copy,
             the device-side repack and the kernel are inside the timed region).
  roofline   tensor bound: achieved = nonzero-block FLOPs per launch / mean launch time,
             peak = MEASURED_PEAKS.json bf16 (burst when the timed region is < 1 s).
  cpu_baseline  the reference's own serial VBR::multiply (oracle/_ref, built from the unmodified
             sources) on a bounded sample of the same workload, 1 thread.

Multi-GPU (torchrun, one rank per GPU): A is sharded by contiguous block-row ranges balanced on
nonzero-block area, B is replicated with one NCCL broadcast (outside the timed region), C stays
row-partitioned; no collective runs inside the timed region.  Total work is fixed => "strong".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: scale, density, w, B rows per group, tau, n, blocking algo
    "rmat16_a5": dict(kind="rmat", scale=16, density=1e-3, w=64, rb=64, tau=0.6, n=2048, algo=5,
                      desc="R-MAT 65536x65536 (a,b,c,d=.57,.19,.19,.05; 4.29M draws, dedup; seed 1), "
                           "-P 1 -a 5 -b 64 -B 64 -t 0.6 (Jaccard), B 65536x2048 uniform(0,1) seed 2"),
    "er14_fixed": dict(kind="er", scale=14, density=2.57e-5, w=64, rb=64, tau=0.0, n=1024, algo=2,
                       desc="ER 16384x16384 p=2.57e-5 seed 1, -a 2 -F 1 -b 64 -B 64 (10% block density), "
                            "B 16384x1024"),
    "rmat16_a4": dict(kind="rmat", scale=16, density=1e-3, w=64, rb=64, tau=0.6, n=4096, algo=4,
                      desc="BASELINE config #4 at a quarter of its side: R-MAT 65536x65536 (same generator), "
                           "-P 1 -a 4 -b 64 -t 0.6 (variable-height VBR: 33 605 block-rows, 90% of height 1, tallest 13 043), "
                           "B 65536x4096"),
    "rmat18_a4": dict(kind="rmat", scale=18, density=2.5e-4, w=64, rb=64, tau=0.6, n=4096, algo=4, precision="tf32",
                      desc="BASELINE config #4 at its stated size: R-MAT 262144x262144 (same generator; 17.18M draws = "
                           "65.5 per row like config #3, i.e. density 2.5e-4 -- BASELINE does not fix it; 14.83M distinct "
                           "nonzeros), -P 1 -a 4 -b 64 -t 0.6 (variable-height VBR: 137 137 block-rows, 127 586 of height 1, "
                           "7.87M nonzero blocks, nztot 773.7M), B 262144x4096, tf32; the grouping (17 min of CPU) ships in "
                           "cache/ through the library's grouping cache"),
    "rmat12_a5": dict(kind="rmat", scale=12, density=4e-3, w=64, rb=64, tau=0.6, n=512, algo=5,
                      desc="small R-MAT 4096x4096 for quick checks"),
}
WORKLOADS["rmat16_a4"]["precision"] = "tf32"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------- workload

def make_matrix(wl):
    from sparta_b200 import synth
    N = 1 << wl["scale"]
    if wl["kind"] == "rmat":
        r, c = synth.rmat_edges(wl["scale"], int(wl["density"] * N * N), seed=1)
    else:
        r, c = synth.er_edges(N, N, wl["density"], seed=1)
    r, c = synth.pin_shape(r, c, N, N)
    rowptr, colind, _ = synth.csr_from_edges(r, c, N)
    return N, rowptr, colind


def make_grouping(wl, N, rowptr, colind):
    """Row clustering through the product's host layer (bit-exact with the reference's
    BlockingEngine::GetGrouping, src/general/blocking.cpp:633), behind the LIBRARY's grouping cache
    (sparta_host_blocking_cached: the reference's `.g` file format keyed on the CSR pattern and the
    flags, include/sparta_b200.h).  cache/ in the repo holds the groupings of the bench workloads so
    that GPU minutes are not spent on CPU blocking (-a 4 at 2^18 rows: 17 minutes)."""
    from sparta_b200 import lib
    if wl["algo"] == 2:
        return np.arange(N, dtype=np.int64) // wl["rb"]
    cache = os.environ.get("SPARTA_BENCH_CACHE", os.path.join(ROOT, "cache"))
    g, hit = lib.host_blocking_cached(cache, N, N, rowptr, colind, algo=wl["algo"], tau=wl["tau"],
                                      block_col_size=wl["w"], row_block_size=wl["rb"], sim_measure=1,
                                      use_pattern=True, use_group=False)
    log(f"[bench] grouping: {'cache hit' if hit else 'computed and cached'} ({cache})")
    return g


def build_vbr(wl, N, rowptr, colind, grouping, weighted=False):
    """VBR::fill_from_CSR_inplace through the product's host layer.  Pattern-only (-P 1, what every
    reference batch script uses) or, with `weighted`, uniform(-1, 1) values (seed 3): the blocking only
    looks at the pattern, so the structure is the same and the A operand's rounding is exercised."""
    from sparta_b200 import lib
    val = None
    if weighted:
        val = np.random.default_rng(3).uniform(-1.0, 1.0, size=len(colind)).astype(np.float32)
    return lib.host_vbr_fill(N, N, rowptr, colind, val, grouping, wl["w"], wl["rb"],
                             force_fixed_size=(wl["algo"] == 2), pattern_only=not weighted)


# --------------------------------------------------------------------------- clocks

class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        rows = [l for (t, l) in self.lines if t0 - 0.05 <= t <= t1 + 0.1] or [l for _, l in self.lines]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in rows:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- CPU reference legs

def cpu_reference_sample(v, Bm, n, target_gflop, threads, variant=None, n_cols=None, min_block_rows=16):
    """Times the reference's serial VBR::multiply (oracle/_ref; else the oracle port) on a sample of
    the workload's block-rows: at least `min_block_rows` of them, EVENLY SPACED over the nonempty
    block-rows (the matrix is sorted, dense groups first, so the first block-rows alone would not
    be representative), more until ~target_gflop of nonzero-block FLOPs are covered.  n_cols < n
    restricts B to its first n_cols columns (the slow -O0 build).  threads > 1 runs that many
    independent copies of the serial routine on disjoint parts of the sample."""
    from oracle.oracle_py import Oracle, Reference
    kind = "reference" if Reference.available(variant) else "port"
    impl = Reference(variant) if kind == "reference" else Oracle()
    nn = n if n_cols is None else min(n, n_cols)
    rp, nz = v["row_part"], v["nzcount"]
    w = v["block_col_size"]
    hts = np.diff(rp)
    area = nz * hts * w                              # elements of mab per block-row
    flops = 2.0 * area * nn
    live = np.nonzero(area > 0)[0]
    want = max(min_block_rows, threads)
    while True:
        take = live[np.unique(np.linspace(0, len(live) - 1, num=min(want, len(live))).round().astype(np.int64))]
        if flops[take].sum() >= target_gflop * 1e9 or len(take) == len(live):
            break
        want *= 2
    # the smallest allowed sample is still too much CPU work at full n: the columns of B are independent in
    # VBR::multiply (vbr.cpp:342-368), so the sample takes a leading subset of them (multiples of 64)
    if n_cols is None and flops[take].sum() > 1.25 * target_gflop * 1e9:
        nn = int(max(64, min(n, (n * target_gflop * 1e9 / flops[take].sum()) // 64 * 64)))
        flops = 2.0 * area * nn
    jab_off = np.concatenate([[0], np.cumsum(nz)])
    mab_off = np.concatenate([[0], np.cumsum(area)])
    # the sample as `threads` small VBR matrices (block-rows dealt round-robin by descending work)
    order = take[np.argsort(-flops[take], kind="stable")]
    parts = [order[t::threads] for t in range(threads)]
    subs = []
    for part in parts:
        if len(part) == 0:
            continue
        part = np.sort(part)
        subs.append({
            "rows": int(hts[part].sum()), "cols": v["cols"], "block_col_size": w,
            "row_part": np.concatenate([[0], np.cumsum(hts[part])]).astype(np.int64), "nzcount": nz[part].copy(),
            "jab": np.concatenate([v["jab"][jab_off[b]:jab_off[b + 1]] for b in part]).astype(np.int64),
            "mab": np.concatenate([v["mab"][mab_off[b]:mab_off[b + 1]] for b in part]),
        })
    Bs = np.ascontiguousarray(Bm[:nn])
    done = [None] * len(subs)

    def work(i):
        done[i] = impl.vbr_multiply(subs[i], Bs, nn)

    t0 = time.perf_counter()
    if len(subs) == 1:
        work(0)
    else:
        ths = [threading.Thread(target=work, args=(i,)) for i in range(len(subs))]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
    dt = time.perf_counter() - t0
    total = float(flops[take].sum())
    build = {"O0": "no -O flag (the reference's `make serial`, makefile:2)", "O3": "-O3 (makefile.MARZOLA:2)",
             None: "-O2"}[variant]
    return {"seconds": dt, "flops": total, "tflops": total / dt / 1e12, "kind": kind,
            "cores": len(subs), "block_rows": int(len(take)),
            "sample": f"{len(take)} of {len(live)} nonempty block-rows, evenly spaced ({total / 1e9:.1f} GFLOP of "
                      f"nonzero-block work) at n={nn}" + ("" if nn == n else f" of {n}") + f"; serial VBR::multiply, {build}"
                      + (f", {len(subs)} independent copies on disjoint block-rows" if len(subs) > 1 else "")}


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from sparta_b200 import synth
    N, rowptr, colind = make_matrix(wl)
    grouping = make_grouping(wl, N, rowptr, colind)
    # only the sampled block-rows are needed; the fill is cheap enough to do whole
    v = build_vbr(wl, N, rowptr, colind, grouping, args.weighted)
    n = wl["n"]
    Bm = synth.seeded_B(v["cols"], n, seed=2)
    threads = args.cpu_threads or (os.cpu_count() or 1)
    # every step is a bounded sample: sized so that warmup + steps end in about 2.5 minutes at the ~0.5 GFLOP/s
    # a thread of this routine sustains when all cores run it (measured: 8.2 GFLOP/s on 16 threads)
    per_step = args.cpu_gflop_per_step
    if per_step <= 0:
        per_step = 0.5 * min(30.0, max(2.0, 150.0 / (args.warmup + args.steps)))
    times, flops, info = [], 0.0, None
    for i in range(args.warmup + args.steps):
        info = cpu_reference_sample(v, Bm, n, per_step * threads, threads)
        if i >= args.warmup:
            times.append(info["seconds"])
            flops = info["flops"]
    t = sum(times)
    value = flops * len(times) / t / 1e12
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, wl, v),
        "cpu_baseline": {"value": value, "unit": "TFLOP/s", "cores": info["cores"], "kind": info["kind"],
                         "sample": info["sample"]},
        "e2e": {"value": value, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm

METRIC = "VBR SpMM effective TFLOP/s (nonzero-block FLOPs)"


def workload_config(args, wl, v):
    return {"workload": args.workload, "description": wl["desc"],
            "values": "uniform(-1,1) seed 3 (weighted)" if args.weighted else "pattern-only (-P 1): A in {0, 1}",
            "rows": int(v["rows"]), "cols": int(v["cols"]),
            "block_col_size": wl["w"], "block_rows": int(v["block_rows"]), "nz_blocks": int(len(v["jab"])),
            "nztot": int(v["nztot"]), "B_cols": wl["n"], "flop_per_step": 2.0 * v["nztot"] * wl["n"],
            "l2_policy": "inputs larger than L2 (packed A alone exceeds 126 MB); no flush",
            "parallelism": f"row-block shards x{args.gpus} ({args.partition}-balanced), B replicated by one NCCL broadcast"}


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"burst": d.get("bf16_tflops"), "sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def lib_sha16():
    import hashlib
    from sparta_b200 import lib as L
    with open(L.LIB_PATH, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()[:16]


def source_sha16():
    """Hash of the kernel-side sources: what a profile is stamped with (the .so itself differs from
    build to build by embedded paths and timestamps of the toolchain)."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "sparta_b200", "csrc")
    # the kernels, the tile scheduler and the layer that picks pipeline / tiles / launch attributes; the host-side
    # builders (blocking, VBR fill, grouping cache) and the multi-GPU driver do not change what a launch does
    kernel_side = ("abi.cu", "csr_kernel.cu", "csr_kernel.h", "pack_kernels.cu", "pack_kernels.h", "sched_types.h",
                   "schedule.cpp", "schedule.h", "spmm_kernel.cu", "spmm_kernel.h")
    for name in kernel_side:
        with open(os.path.join(d, name), "rb") as f:
            h.update(name.encode())
            h.update(f.read())
    return h.hexdigest()[:16]


def read_traffic(workload, precision):
    """DRAM bytes per launch (and tensor-pipe activity) of the SpMM kernel from the committed ncu
    --set full capture -- only if profiles/traffic.json was taken from THIS build of the kernel
    sources (its `source_sha16` stamp); a stale capture is reported as null, not as a number."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None, "no profiles/traffic.json"
    with open(p) as f:
        d = json.load(f)
    rec = d.get(f"{workload}:{precision}")
    if not rec:
        return None, None, "no capture of this workload"
    if rec.get("source_sha16") != source_sha16():
        return None, None, f"capture is of another build ({rec.get('source_sha16')} != {source_sha16()})"
    return rec.get("dram_bytes"), rec.get("pipe_tensor_active_pct"), rec.get("source", "profiles/traffic.json")


def measure_tf32_peak(dev):
    """cuBLAS tf32 GEMM rate on this GPU in this run (SURVEY section 7 "which peak"): torch.matmul of
    two 8192^2 fp32 matrices with TF32 tensor cores allowed, best of 10, CUDA events."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn((8192, 8192), device=dev, dtype=torch.float32)
        b = torch.randn((8192, 8192), device=dev, dtype=torch.float32)
        for _ in range(3):
            torch.matmul(a, b)
        best = 1e30
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    import sparta_b200
    from sparta_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    lib = sparta_b200.load()
    if lib.sparta_device_count() < 1:
        raise SystemExit("bench.py needs a compute-capability 10.x GPU (there is no CPU path); "
                         "use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- host side: matrix -> grouping (rank 0, broadcast) -> VBR
    t0 = time.perf_counter()
    N, rowptr, colind = make_matrix(wl)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    if world > 1:
        g = torch.empty(N, dtype=torch.int64, device=dev)
        if rank == 0:
            g.copy_(torch.from_numpy(make_grouping(wl, N, rowptr, colind)))
        dist.broadcast(g, 0)
        grouping = g.cpu().numpy()
    else:
        grouping = make_grouping(wl, N, rowptr, colind)
    t_block = time.perf_counter() - t0
    t0 = time.perf_counter()
    v = build_vbr(wl, N, rowptr, colind, grouping, args.weighted)
    t_fill = time.perf_counter() - t0
    n = wl["n"]
    if rank == 0:
        if args.partition == "model":
            # contiguous block-row ranges balanced on the scheduler's modelled kernel time per shard
            cuts = sparta_b200.partition_block_rows_modelled(v["rows"], v["cols"], wl["w"], v["row_part"], v["nzcount"],
                                                             v["jab"], n, world, precision=args.precision,
                                                             **tuning_opts(args))
        else:
            cuts = sparta_b200.partition_block_rows(v["row_part"], v["nzcount"], world)
    if world > 1:   # one rank cuts (a few seconds of host threads at 10^5 block-rows), everybody gets the cuts
        ct = torch.zeros(world + 1, dtype=torch.int64, device=dev)
        if rank == 0:
            ct.copy_(torch.from_numpy(np.asarray(cuts, dtype=np.int64)))
        dist.broadcast(ct, 0)
        cuts = ct.cpu().numpy()
    lo, hi = int(cuts[rank]), int(cuts[rank + 1])
    if rank == 0:
        log(f"[bench] matrix {N}x{N} nnz={len(colind)} gen {t_gen:.1f}s blocking {t_block:.1f}s fill {t_fill:.1f}s "
            f"block_rows={v['block_rows']} nz_blocks={len(v['jab'])} nztot={v['nztot']}")

    # ---- device side: A shard resident, B replicated by ONE NCCL broadcast
    def make_handle(lo_, hi_):
        return sparta_b200.Handle.from_vbr(v["rows"], v["cols"], wl["w"], v["row_part"], v["nzcount"], v["jab"],
                                           v["mab"], precision=args.precision, device=local,
                                           block_row_begin=lo_, block_row_end=hi_, n_hint=n, **tuning_opts(args))
    h = make_handle(lo, hi)
    Bd = torch.empty((n, v["cols"]), dtype=torch.float32, device=dev)
    Bm = None
    if rank == 0:
        Bm = synth.seeded_B(v["cols"], n, seed=2)
        Bd.copy_(torch.from_numpy(Bm))
    if world > 1:
        torch.cuda.synchronize()
        tb0 = time.perf_counter()
        dist.broadcast(Bd, 0)
        torch.cuda.synchronize()
        t_bcast = time.perf_counter() - tb0
    else:
        t_bcast = 0.0
    h.set_B_device(Bd.data_ptr(), v["cols"], n)

    # ---- the modelled partition corrected by measured shard times (at most twice), where the slowest rank is more
    # than 3 % above the mean; a re-cut that does not lower the slowest shard's time is undone.  Setup, outside
    # every timed region.
    rebalance = None
    if world > 1 and args.partition == "model" and args.rebalance:
        def shard_times(hh):
            for _ in range(3):
                hh.run_async()
            hh.synchronize()
            ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_ = torch.cuda.ExternalStream(hh.stream, device=dev)
            ev_a.record(s_)
            for _ in range(5):
                hh.run_async()
            ev_b.record(s_)
            ev_b.synchronize()
            mine = torch.tensor([ev_a.elapsed_time(ev_b) / 5], dtype=torch.float64, device=dev)
            allms = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
            dist.all_gather(allms, mine)
            return np.array([float(x.item()) for x in allms])

        def rebuild(new_cuts):
            nonlocal h, lo, hi
            lo, hi = int(new_cuts[rank]), int(new_cuts[rank + 1])
            h.close()
            h = make_handle(lo, hi)
            h.set_B_device(Bd.data_ptr(), v["cols"], n)

        measured = shard_times(h)
        for _ in range(2):
            if measured.max() <= 1.03 * measured.mean():
                break
            ct = torch.zeros(world + 1, dtype=torch.int64, device=dev)
            if rank == 0:
                modelled = sparta_b200.partition_model_times(v["rows"], v["cols"], wl["w"], v["row_part"], v["nzcount"],
                                                             v["jab"], n, cuts, precision=args.precision, **tuning_opts(args))
                new_cuts = sparta_b200.partition_block_rows_measured(v["rows"], v["cols"], wl["w"], v["row_part"],
                                                                     v["nzcount"], v["jab"], n, world, cuts, measured,
                                                                     modelled, precision=args.precision, **tuning_opts(args))
                ct.copy_(torch.from_numpy(np.asarray(new_cuts, dtype=np.int64)))
            dist.broadcast(ct, 0)
            old_cuts, new_cuts = cuts, ct.cpu().numpy()
            if np.array_equal(np.asarray(old_cuts), new_cuts):
                break
            rebuild(new_cuts)
            again = shard_times(h)
            step = {"cuts": [int(c) for c in old_cuts], "ms_per_rank": measured.tolist(),
                    "recut": [int(c) for c in new_cuts], "recut_ms_per_rank": again.tolist()}
            rebalance = (rebalance or []) + [step]
            if again.max() < measured.max():
                cuts, measured = new_cuts, again
                step["kept"] = True
            else:
                rebuild(old_cuts)
                step["kept"] = False
                break
    del Bd
    st = h.stats()
    stream = torch.cuda.ExternalStream(h.stream, device=dev)
    my_flops = 2.0 * st["nztot"] * n
    total_flops = 2.0 * v["nztot"] * n

    def barrier():
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        h.run_async()
    h.synchronize()

    launches_before = h.stats()["kernel_launches"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    tm0 = sampler.mark()
    ev0.record(stream)
    for _ in range(args.steps):
        h.run_async()
    ev1.record(stream)
    ev1.synchronize()
    torch.cuda.synchronize()
    tm1 = sampler.mark()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(tm0, tm1) if rank == 0 else None
    launches = h.stats()["kernel_launches"] - launches_before
    t_all = torch.tensor([ms], dtype=torch.float64, device=dev)
    per_rank_ms = [ms / args.steps]
    if world > 1:
        gathered = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(gathered, t_all)
        per_rank_ms = [float(g.item()) / args.steps for g in gathered]
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    ms_max = float(t_all.item())
    value = total_flops * args.steps / (ms_max * 1e-3) / 1e12

    # ---- correctness spot check inside the bench (fp64 recomputation of sampled rows)
    check = spot_check(h, v, lo, hi, n, args.precision, Bm, dist if world > 1 else None, dev, rank)

    # ---- optional: assemble the row-partitioned C on every rank with one NCCL all-gather
    gather = None
    if args.gather_c and world > 1:
        gather = gather_c(h, v, cuts, n, rank, world, dev, dist)

    # ---- e2e: the one-shot reference-facing call, host buffers in and out (rank-local shard)
    e2e = None
    if not args.no_e2e:
        rows_shard = int(v["row_part"][hi] - v["row_part"][lo])
        c_resident = None
        if rows_shard:
            c_resident = np.zeros((n, rows_shard), dtype=np.float32)
            h.get_C(c_resident, rows_shard)
        val_csr = None
        if args.weighted:
            val_csr = np.random.default_rng(3).uniform(-1.0, 1.0, size=len(colind)).astype(np.float32)
        e2e = run_e2e(args, wl, v, lo, hi, n, Bm, world, rank, dev, dist if world > 1 else None, total_flops,
                      c_resident, (N, rowptr, colind, val_csr), grouping)
    h.close()

    # ---- CPU baseline beside it (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        info = cpu_reference_sample(v, Bm, n, args.cpu_gflop, 1)
        cpu = {"value": info["tflops"], "unit": "TFLOP/s", "cores": 1, "kind": info["kind"],
               "sample": info["sample"] + f"; {info['seconds']:.1f} s", "host_cores_available": os.cpu_count()}
        # the reference's own optimisation levels on a smaller sample (the -O0 build is 8x slower)
        variants = {}
        for var, cols_sub, gf in (("O0", 64, 0.5), ("O3", None, args.cpu_gflop / 3)):
            from oracle.oracle_py import Reference
            if Reference.available(var):
                vi = cpu_reference_sample(v, Bm, n, gf, 1, variant=var, n_cols=cols_sub)
                variants[var] = {"value": vi["tflops"], "sample": vi["sample"] + f"; {vi['seconds']:.1f} s"}
        cpu["variants"] = variants

    if rank == 0:
        peaks = read_peaks()
        per_launch_ms = ms_max / args.steps
        timed_s = ms_max * 1e-3
        peak_kind = "burst" if timed_s < 1.0 else "sustained"
        if args.precision == "tf32":
            # no tf32 figure in MEASURED_PEAKS.json: measured here, in this run, on this GPU
            tf32_peak = measure_tf32_peak(dev)
            peak = tf32_peak * world
            peak_desc = f"cuBLAS tf32 8192^3 measured in this run ({tf32_peak:.1f} TFLOP/s per GPU, burst)"
        else:
            peak = peaks[peak_kind] * world
            peak_desc = f"{peak_kind} bf16 cuBLAS, {peaks['source']}"
        if world > 1:
            peak_desc += f" x{world} GPUs"
        achieved = total_flops / (per_launch_ms * 1e-3) / 1e12
        bytes_min = st_bytes_min(v, n, args.precision)
        traffic, pipe_active, traffic_src = read_traffic(args.workload, args.precision)
        line = {
            "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per_launch_ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": workload_config(args, wl, v),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "pipe_tensor_active_pct": pipe_active,
                         "peak_kind": peak_desc,
                         "frac_of_nominal_dense": achieved / ((1125.0 if args.precision == "tf32" else 2250.0) * world),
                         "kernel": "spmm_vbr_sm100", "algorithmic_bytes": bytes_min,
                         "hbm_peak_gbs": peaks["hbm_gbs"] * world,
                         "hbm_frac_of_measured": bytes_min / (per_launch_ms * 1e-3) / 1e9 / (peaks["hbm_gbs"] * world)},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "check": check,
            "gather_c": gather,
            "setup": {"blocking_s": t_block, "vbr_fill_s": t_fill, "matrix_gen_s": t_gen,
                      "a_upload_pack_ms": st["upload_ms"], "b_broadcast_s": t_bcast,
                      "sched_imbalance": st["sched_imbalance"], "grid": st["grid"], "items": st["items"],
                      "team": st["team"], "split_pieces": st["split_pieces"], "zero_tiles": st["zero_tiles"],
                      "gather_rows": st["gather_rows"], "gather_nnz": st["gather_nnz"], "chunks": st["chunks"],
                      "wide_tiles": st["wide_tiles"], "super_rows": st["super_rows"], "smem_bytes": st["smem_bytes"],
                      "shard_block_rows": [int(c) for c in cuts], "per_rank_ms_per_step": per_rank_ms,
                      "rebalanced_from": rebalance},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def tuning_opts(args):
    o = {}
    for k in ("seg_rows", "acc_cols", "panel_stages", "num_ctas", "cta_pair", "row_order", "l2_slab_mb", "max_chain", "split_k", "copy_warps", "fuse_rows", "pipeline", "gather_max_height", "gather_passes", "wide_tiles"):
        val = getattr(args, k)
        if val:
            o[k] = val
    return o


def st_bytes_min(v, n, precision):
    """A once + B once (distinct column blocks touched) + C written once."""
    es = 4 if precision == "tf32" else 2
    touched = len(np.unique(v["jab"]))
    return float(v["nztot"] * es + touched * v["block_col_size"] * n * es + v["rows"] * n * 4)


def spot_check(h, v, lo, hi, n, precision, Bm, dist, dev, rank, n_check=64):
    """fp64 recomputation of a few block-rows of this rank's shard from the operands rounded to
    the kernel's input precision; returns the max relative error (norm of SURVEY 8(c))."""
    import torch
    from sparta_b200.synth import round_to
    rp, nz, w = v["row_part"], v["nzcount"], v["block_col_size"]
    if dist is not None:
        # every rank needs B on the host for the check: fetch it from rank 0
        Bt = torch.empty((n, v["cols"]), dtype=torch.float32, device=dev)
        if rank == 0:
            Bt.copy_(torch.from_numpy(Bm))
        dist.broadcast(Bt, 0)
        Bm = Bt.cpu().numpy()
    rows_shard = int(rp[hi] - rp[lo])
    if rows_shard == 0:
        return {"max_rel_err": 0.0, "block_rows_checked": 0}
    out = np.zeros((n, rows_shard), dtype=np.float32)
    h.get_C(out, rows_shard)
    jab_off = np.concatenate([[0], np.cumsum(nz)])
    mab_off = np.concatenate([[0], np.cumsum(nz * np.diff(rp) * w)])
    rng = np.random.default_rng(7 + rank)
    cand = np.arange(lo, hi)
    # the first, the last, the tallest, the one with most blocks and random ones: >= 64 block-rows
    tall = lo + int(np.argmax(np.diff(rp)[lo:hi]))
    wide = lo + int(np.argmax(nz[lo:hi]))
    pick = np.unique(np.concatenate([[lo, hi - 1, tall, wide],
                                     rng.choice(cand, size=min(n_check, len(cand)), replace=False)]))
    Br = round_to(Bm, precision).astype(np.float64)
    Bf = Bm.astype(np.float64)
    worst, worst_r, scale = 0.0, 0.0, 0.0
    rows_checked = 0
    for ib in pick:
        hgt = int(rp[ib + 1] - rp[ib])
        use = min(hgt, 256)             # of a very tall block-row the first 256 rows
        rows_checked += use
        acc = np.zeros((n, use))
        acc_r = np.zeros((n, use))
        for q in range(int(nz[ib])):
            jb = int(v["jab"][jab_off[ib] + q])
            blk = v["mab"][mab_off[ib] + q * hgt * w: mab_off[ib] + (q + 1) * hgt * w].reshape(w, hgt)[:, :use]
            k0, k1 = jb * w, min((jb + 1) * w, v["cols"])
            acc += Bf[:, k0:k1] @ blk.astype(np.float64)[:k1 - k0]
            acc_r += Br[:, k0:k1] @ round_to(blk, precision).astype(np.float64)[:k1 - k0]
        got = out[:, rp[ib] - rp[lo]: rp[ib] - rp[lo] + use]
        worst = max(worst, float(np.abs(got - acc).max()))
        worst_r = max(worst_r, float(np.abs(got - acc_r).max()))
        scale = max(scale, float(np.abs(acc).max()))
    tol = 1e-5 if precision == "tf32" else 2e-2
    err = worst / max(scale, 1e-30)
    err_r = worst_r / max(scale, 1e-30)
    # tf32: the <= 1e-5 bound is on the arithmetic (operands as the kernel sees them, SURVEY 8c option i);
    # the error against the unrounded fp32 operands is reported next to it
    ok = err_r <= tol if precision == "tf32" else err <= tol
    return {"max_rel_err": err, "tolerance": tol, "ok": bool(ok), "block_rows_checked": int(len(pick)),
            "rows_checked": int(rows_checked),
            "tolerance_applies_to": "max_rel_err_vs_rounded_operands" if precision == "tf32" else "max_rel_err",
            "against": "fp64 recomputation of sampled block-rows from the fp32 operands (norm max|dC|/max|C|)",
            "max_rel_err_vs_rounded_operands": worst_r / max(scale, 1e-30)}


def gather_c(h, v, cuts, n, rank, world, dev, dist):
    """C stays row-partitioned by default; this is the optional all-gather (outside the timed
    region of `value`): the ragged slabs are padded to the tallest, gathered over NCCL and
    checked -- every rank must end up with every other rank's slab bit for bit."""
    import torch
    from sparta_b200 import dist as sd
    rp = v["row_part"]
    rows_per_rank = [int(rp[int(cuts[r + 1])] - rp[int(cuts[r])]) for r in range(world)]
    mine = rows_per_rank[rank]
    slab = torch.zeros((n, max(mine, 1)), dtype=torch.float32, device=dev)
    if mine:
        h.get_C_device(slab.data_ptr(), mine)
    slab = slab[:, :mine]
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    full = sd.all_gather_C(slab, rows_per_rank, n)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # checksum of every slab as its owner computed it vs as this rank received it
    own = torch.zeros(world, dtype=torch.float64, device=dev)
    own[rank] = slab.double().sum() if mine else 0.0
    dist.all_reduce(own)
    offs = np.concatenate([[0], np.cumsum(rows_per_rank)])
    got = torch.stack([full[:, offs[r]:offs[r + 1]].double().sum() for r in range(world)])
    same = bool(torch.equal(own, got)) and bool(torch.equal(full[:, offs[rank]:offs[rank + 1]], slab))
    ok = torch.tensor([1.0 if same else 0.0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"ms": float(t.item()), "bytes_per_rank_out": int(full.numel() * 4), "ok": bool(ok.item() == 1.0),
            "how": "slabs padded to the tallest, one NCCL all_gather, padding cut away (sparta_b200/dist.py)"}


def run_e2e(args, wl, v, lo, hi, n, Bm, world, rank, dev, dist, total_flops, c_resident, csr, grouping):
    """End to end through the public API with HOST buffers, every copy inside the timed region.

    N = 1: the one-shot sparta_csr_vbr_spmm -- host CSR + grouping + host B in, host C out.  It
    replaces fill_from_CSR_inplace + cublas_fixed_blocks_multiply (cuda_multiply.cpp:129-137): index
    arrays on the host, dense blocks rebuilt on the device from the nonzeros, kernel, download.
    N > 1: the same flow sharded -- every rank builds the handle of ITS block-rows from the CSR
    (Handle.from_csr_grouping), rank 0 alone uploads B and ONE NCCL broadcast replicates it, every rank
    multiplies and downloads its slab of C.  `vbr_arrays` times the drop-in sparta_vbr_spmm call (host
    VBR arrays incl. the fp32 mab in) on the rank's shard beside it."""
    import ctypes as C
    import torch
    import sparta_b200
    from sparta_b200 import lib as L
    lib = sparta_b200.load()
    rp, nz, w = v["row_part"], v["nzcount"], v["block_col_size"]
    N_rows, rowptr, colind, val = csr
    rows_s = int(rp[hi] - rp[lo])

    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()
    k_rowptr, a_rowptr = pin(rowptr.astype(np.int64))
    k_colind, a_colind = pin(colind.astype(np.int64))
    k_grp, a_grp = pin(np.asarray(grouping, dtype=np.int64))
    a_val = None
    if val is not None:
        k_val, a_val = pin(val.astype(np.float32))
    k_C = torch.zeros((n, max(rows_s, 1)), dtype=torch.float32).pin_memory()
    a_C = k_C.numpy()
    if rank == 0:
        k_B, a_B = pin(Bm)
    steps = max(1, min(args.steps, args.e2e_steps))
    dt = C.c_float(0)
    fixed = wl["algo"] == 2
    Bd = torch.empty((n, v["cols"]), dtype=torch.float32, device=dev) if world > 1 else None
    stream = None

    def once():
        if world == 1:
            L._check(lib.sparta_csr_vbr_spmm(N_rows, N_rows, L._ptr(a_rowptr), L._ptr(a_colind),
                                             None if a_val is None else L._ptr(a_val), L._ptr(a_grp), w, wl["rb"],
                                             int(fixed), L._ptr(a_B), v["cols"], n, L._ptr(a_C), max(rows_s, 1),
                                             L.PRECISIONS[args.precision], C.byref(dt)))
            return
        # B first: its upload (rank 0) and the NCCL broadcast are enqueued, then every rank builds the handle of
        # its shard on the host while B crosses PCIe and NVLink
        if rank == 0:
            Bd.copy_(k_B, non_blocking=True)          # the ONE upload of B
        dist.broadcast(Bd, 0)                          # NCCL over NVLink (returns once enqueued)
        hh = sparta_b200.Handle.from_csr_grouping(N_rows, N_rows, a_rowptr, a_colind, a_val, a_grp, w, wl["rb"], fixed,
                                                  precision=args.precision, device=dev.index, block_row_begin=lo,
                                                  block_row_end=hi, n_hint=n, **tuning_opts(args))
        try:
            torch.cuda.current_stream().synchronize()
            if rows_s:
                hh.set_B_device(Bd.data_ptr(), v["cols"], n)
                hh.run()
                hh.get_C(a_C[:, :rows_s] if a_C.shape[1] == rows_s else a_C, max(rows_s, 1))
        finally:
            hh.close()

    for _ in range(2):   # warm-up (context, stream-ordered allocator pools of two streams, NCCL channel)
        once()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    each = []
    for _ in range(steps):
        ts = time.perf_counter()
        once()
        each.append(1e3 * (time.perf_counter() - ts))
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    sec = time.perf_counter() - t0
    diff = 0.0
    if rows_s and c_resident is not None:
        scale = float(np.abs(c_resident).max()) or 1.0
        diff = float(np.abs(a_C[:, :rows_s] - c_resident).max()) / scale
    h2d = a_rowptr.nbytes + a_colind.nbytes + a_grp.nbytes + (a_val.nbytes if a_val is not None else 0)
    h2d_b = Bm.nbytes if rank == 0 else 0
    d2h = rows_s * n * 4
    each_all = [[round(x, 2) for x in each]]
    if dist is not None:
        et = torch.tensor(each, dtype=torch.float64, device=dev)
        gathered = [torch.empty_like(et) for _ in range(world)]
        dist.all_gather(gathered, et)
        each_all = [[round(float(x), 2) for x in g.tolist()] for g in gathered]
    t_all = torch.tensor([sec, float(h2d_b), float(d2h), float(diff)], dtype=torch.float64, device=dev)
    if dist is not None:
        tmax = t_all.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_all, op=dist.ReduceOp.SUM)
        sec, diff = float(tmax[0].item()), float(tmax[3].item())
        h2d_b, d2h = float(t_all[1].item()), float(t_all[2].item())
    # nonzeros actually sent: this rank's (offset, value) pairs; the index arrays are host-side input
    out = {"value": total_flops * steps / sec / 1e12, "unit": "TFLOP/s",
           "h2d_bytes_per_step": int(h2d_b + len(colind) * 12 + 0), "d2h_bytes_per_step": int(d2h), "steps": steps,
           "ms_per_step": 1e3 * sec / steps, "ms_each_step_per_rank": each_all,
           "max_rel_diff_vs_resident_handle": diff,
           "same_result": bool(diff <= 1e-6), "host_input_bytes": int(h2d + Bm.nbytes if rank == 0 else h2d),
           "call": ("sparta_csr_vbr_spmm (host CSR + grouping + host B -> host C: index build, nonzeros and B up, "
                    "device-side block rebuild + pack, kernel, C down; pinned host buffers; wall clock)" if world == 1 else
                    "per rank Handle.from_csr_grouping on its block-rows + B uploaded by rank 0 and ONE NCCL broadcast + "
                    "run + download of the rank's C slab; wall clock, max over ranks"),
           "h2d_note": "B (fp32) + 12 bytes per nonzero (element offset + value); the host-side index build is inside the timed region"}
    # the drop-in call on host VBR arrays (fp32 mab included), this rank's shard, one repetition
    jab_off = np.concatenate([[0], np.cumsum(nz)])
    mab_off = np.concatenate([[0], np.cumsum(nz * np.diff(rp) * w)])
    if rows_s and not args.no_e2e_vbr:
        if dist is not None:
            Bt = torch.empty((n, v["cols"]), dtype=torch.float32, device=dev)
            if rank == 0:
                Bt.copy_(torch.from_numpy(Bm))
            dist.broadcast(Bt, 0)
            Bm = Bt.cpu().numpy()
            del Bt
        k1, a_rp = pin((rp[lo:hi + 1] - rp[lo]).astype(np.int64))
        k2, a_nz = pin(nz[lo:hi].astype(np.int64))
        k3, a_jab = pin(v["jab"][jab_off[lo]:jab_off[hi]].astype(np.int64))
        k4, a_mab = pin(v["mab"][mab_off[lo]:mab_off[hi]])
        k5, a_B2 = pin(Bm)

        def once_vbr():
            L._check(lib.sparta_vbr_spmm(rows_s, v["cols"], hi - lo, w, L._ptr(a_rp), L._ptr(a_nz), L._ptr(a_jab),
                                         L._ptr(a_mab), L._ptr(a_B2), v["cols"], n, L._ptr(a_C), max(rows_s, 1),
                                         L.PRECISIONS[args.precision], C.byref(dt)))
        once_vbr()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        once_vbr()
        torch.cuda.synchronize()
        sec_v = time.perf_counter() - t0
        d2 = float(np.abs(a_C[:, :rows_s] - c_resident).max()) / (float(np.abs(c_resident).max()) or 1.0) if c_resident is not None else 0.0
        out["vbr_arrays"] = {"ms_per_step_this_rank": 1e3 * sec_v, "h2d_bytes": int(a_rp.nbytes + a_nz.nbytes + a_jab.nbytes + a_mab.nbytes + a_B2.nbytes),
                             "max_rel_diff_vs_resident_handle": d2,
                             "call": "sparta_vbr_spmm on the rank's shard (host VBR arrays incl. the fp32 mab + host B -> host C)"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rmat16_a5", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default=None, choices=["bf16", "fp16", "tf32"],
                    help="operand precision (default: the workload's, bf16 unless BASELINE names another)")
    ap.add_argument("--weighted", action="store_true", help="uniform(-1,1) values instead of the pattern-only matrix")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-e2e-vbr", action="store_true", help="skip the extra timing of the sparta_vbr_spmm call on host VBR arrays")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather-c", action="store_true", help="multi-GPU: also all-gather C over NCCL and verify it")
    ap.add_argument("--cpu-gflop", type=float, default=24.0, help="size of the cpu_baseline sample (about 25 s of the reference's -O2 build)")
    ap.add_argument("--cpu-gflop-per-step", type=float, default=0.0,
                    help="--impl reference: nonzero-block GFLOP per thread and step (0: sized so that the run ends in "
                         "about 2.5 minutes)")
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--rebalance", type=int, default=1,
                    help="multi-GPU: re-cut the modelled partition once with measured shard times when the slowest rank "
                         "is more than 3 %% above the mean, at most twice, a re-cut that does not help is undone (0: never)")
    ap.add_argument("--partition", default="model", choices=["model", "area"],
                    help="multi-GPU block-row partition: balanced on modelled shard time (default) or on nonzero-block area")
    for k in ("seg_rows", "acc_cols", "panel_stages", "num_ctas", "cta_pair", "row_order", "l2_slab_mb", "max_chain", "split_k", "copy_warps", "fuse_rows", "pipeline", "gather_max_height", "gather_passes", "wide_tiles"):
        ap.add_argument("--" + k.replace("_", "-"), dest=k, type=int, default=0)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("[bench] warmup < 3 breaks the timing rules; using 3")
        args.warmup = 3
    wl = WORKLOADS[args.workload]
    if args.precision is None:
        args.precision = wl.get("precision", "bf16")
    if args.impl == "reference":
        run_reference_arm(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
