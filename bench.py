#!/usr/bin/env python
"""
bench.py -- fp64 cell-updates/s of the 2-D WENO5 MHD hot path on 1/2/4/8 B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun)
    python bench.py --impl reference [...]                    CPU arm: the oracle on host cores

A "step" is one full RK4 time step (4 stages x 2 sweeps, constrained transport, boundary
fill and the CFL reduction for dt) of every interior cell: one cell-update per cell per step
(SURVEY.md 8d).  Workload: Orszag-Tang vortex, 4096 x 4096 cells PER GPU (weak scaling:
global grid 4096 x 4096*N, y-slabs, NCCL halo exchange per stage + max-allreduce per step).

value   device-resident throughput: state already in HBM, K steps of weno5_advance(), timed
        with CUDA events on the solver's stream, barrier + synchronize on both sides, max
        over ranks.
e2e     the same metric through the reference-shaped call sequence with HOST buffers:
        upload(q,bxb,byb) -> timestep() -> classical_RK4_step(dt) -> download(), every step,
        pinned host memory, copies inside the timed region.
roofline  the dominant kernel is the flux sweep (flux_ws_kernel<DIR>, 8 launches per step); algorithmic
        bytes per launch = 1048 B/cell-update / 8 launches * cells; duration from CUDA events
        recorded around every flux launch inside the timed region.  The binding roof is the FP64 pipe
        (bound = "fp64"); achieved/peak/frac stay the HBM figures the metric asks for.
cpu_baseline  the C oracle (oracle/, "port": the Julia reference cannot run here), OpenMP on
        all host cores, on a bounded sample (rank 0, N=1 only).
strong  BASELINE config 4 (Orszag-Tang 8192^2 global, 8192/N rows per GPU) timed for a few steps in the same
        run; at N=8 also config 5 (blast wave 16384^2).  N>1 runs start with the slab bit-identity
        pre-check (julia_b200/slabcheck.py) whose outcome is part of the line.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

NX_PER_GPU = 4096
NY_PER_GPU = 4096
ALGO_BYTES_PER_CELL_UPDATE = 1048.0
ALGO_FLOPS_PER_CELL_UPDATE = 22.0e3          # SURVEY.md 8d (FMA = 2)
FLUX_LAUNCHES_PER_STEP = 8
FP64_DFMA_PER_CLK = 148 * 64                 # B200: 148 SMs x 64 DFMA per clock
FP64_PEAK_TFLOPS_MEASURED = 37.0             # profiles/fp64_peak.txt (scripts/microbench/fp64_peak.cu on this pool's B200)
# the CPU arm times an n x n Orszag-Tang sample, the same at every N (the environment knob exists for the host-side test)
REF_SAMPLE_N = int(os.environ.get("WENO5_REF_SAMPLE_N", "2048"))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, mx, reasons, power = [], [], set(), []
        for t, line in self.samples:
            if t < t0 or t > t1:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                power.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm
def oracle_rate(n, steps, fast=True):
    from julia_b200 import problems as pr
    from oracle import oracle_c as O
    g, q, bx, by = pr.orszag_tang(n)
    t0 = time.time()
    O.advance_2d(g, q, bx, by, steps, fast=fast)
    dt = time.time() - t0
    return n * n * steps / dt, dt


def cpu_baseline(budget_s=15.0):
    host_threads()
    from oracle import oracle_c as O
    O.build()
    rate, _ = oracle_rate(256, 1)
    n = int(max(256, min(2048, (rate * budget_s / 2) ** 0.5 // 64 * 64)))
    steps = int(max(2, min(12, round(budget_s * rate / (n * n)))))
    rate, secs = oracle_rate(n, steps)
    out = {"value": rate, "unit": "cell-updates/s", "cores": O.num_threads(True), "kind": "port",
           "host_cpus": os.cpu_count(),
           "sample": f"orszag_tang {n}x{n}, {steps} RK4 steps incl. timestep, C oracle -O3 OpenMP, {secs:.1f} s"}
    # the same port on ONE host thread (the reference itself is single-threaded)
    try:
        gomp = C.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(1)
        r1, s1 = oracle_rate(512, 1)
        gomp.omp_set_num_threads(int(out["cores"]))
        out["single_thread"] = {"value": r1, "unit": "cell-updates/s", "sample": f"orszag_tang 512x512, 1 RK4 step, {s1:.1f} s"}
    except OSError:
        pass
    return out


def run_ext3d(n, steps, with_cpu):
    """3-D extension beside the headline: weno5_advance_3d on the 3-D Orszag-Tang vortex n^3,
    device-resident, dt recomputed every step; rate from the wall clock around the call (it returns after a stream
    synchronize).  CPU arm: the 3-D oracle (OpenMP) on a 64^3 sample."""
    import time
    from julia_b200 import problems3d as p3
    from julia_b200.solver3d import Solver3D
    g, q, bx, by, bz = p3.orszag_tang_3d(n)
    with Solver3D(g) as s:
        s.upload(q, bx, by, bz)
        s.advance(2)
        t0 = time.perf_counter()
        _, nd = s.advance(steps)
        wall = time.perf_counter() - t0
        s.profile(True)
        s.advance(1)
        ms, nl = s.profile_read()
        divb = float(abs(s.divb()[3:-3, 3:-3, 3:-3]).max())
        launches = s.launch_count()
    out = {"workload": f"orszag_tang_3d {n}^3, periodic, one GPU", "steps": int(nd), "ms_per_step": 1e3 * wall / steps,
           "value": n ** 3 * steps / wall, "unit": "cell-updates/s",
           "sweep_ms_per_launch_xyz": [m / 4 for m in ms], "sweep_share_of_step": sum(ms) / (1e3 * wall / steps),
           "algorithmic_bytes_per_cell_update": 1152, "max_divb": divb, "gpu_launches": int(launches),
           "timing": "host clock around weno5_advance_3d (synchronises on return)"}
    if with_cpu:
        from oracle import oracle_c as O
        g, q, bx, by, bz = p3.orszag_tang_3d(64)
        t0 = time.perf_counter()
        O.advance_3d(g, q, bx, by, bz, 1, fast=True)
        sec = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 64 ** 3 / sec, "unit": "cell-updates/s", "cores": O.num_threads(True), "kind": "port",
                               "sample": f"orszag_tang_3d 64^3, 1 RK4 step incl. timestep, 3-D C oracle -O3 OpenMP, {sec:.1f} s"}
    return out


def host_threads():
    """All host cores for the CPU arm, at every N: torchrun exports OMP_NUM_THREADS=1 to its workers, which would
    leave the oracle on one thread (round-1 SCALE record).  Set before the oracle's OpenMP runtime starts, and
    again through libgomp in case it already has."""
    n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except OSError:
        pass
    return n


def run_reference(args):
    """The reference's own CPU implementation cannot run natively (Julia 0.6 source, no julia binary): this
    arm times the oracle port of it (bit-identical numerics, tests/test_reference_pin.py) on all host cores, on a
    bounded sample of the workload -- the same sample and the same thread count whatever N is."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    want = host_threads()
    from oracle import oracle_c as O
    O.build()
    from julia_b200 import problems as pr
    n = REF_SAMPLE_N
    g, q, bx, by = pr.orszag_tang(n)
    q, bx, by, _ = O.advance_2d(g, q, bx, by, args.warmup, fast=True)
    t0 = time.time()
    O.advance_2d(g, q, bx, by, args.steps, fast=True)
    dt = time.time() - t0
    value = n * n * args.steps / dt
    cores = O.num_threads(True)
    sample = (f"orszag_tang {n}x{n} (bounded sample of the {NX_PER_GPU}x{NY_PER_GPU}-per-GPU workload; the rate is per "
              f"cell-update, so it is size-normalised), {args.steps} RK4 steps incl. timestep, {cores} OpenMP threads")
    line = {
        "impl": "reference", "metric": "cell_updates_per_sec", "value": value, "unit": "cell-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "ms_per_step_of": f"the {n}x{n} sample", "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": cores, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count(), "threads_requested": want},
        "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "no julia binary exists here (Julia 0.6 source); the reference only runs through the mechanical "
                "translation used for the golden vectors (scripts/make_reference_golden.py, ~1e4x slower than native "
                "Julia, not a meaningful timing).  Timed instead: the C port of its E-branch numerics, which "
                "tests/test_reference_pin.py shows to be bit-identical to it, OpenMP over lines, all host threads",
    }
    if cores <= 1 and (os.cpu_count() or 1) > 1:
        line["warning"] = "the CPU arm ran on ONE thread of a multi-core host: this line is not comparable"
    print(json.dumps(line), flush=True)


def workload_config(ngpu):
    return {"workload": f"orszag_tang {NX_PER_GPU}x{NY_PER_GPU} cells per GPU, global {NX_PER_GPU}x{NY_PER_GPU * ngpu}, "
                        "periodic, gamma=5/3, CFL 0.8, dt recomputed every step",
            "nx": NX_PER_GPU, "ny_global": NY_PER_GPU * ngpu, "partition": f"y-slabs x{ngpu}",
            "l2": "no flush needed: ~6.5 GB of resident planes per GPU >> 126 MB L2",
            "es_roundtrip": 1,
            "reference_arm_sample": f"--impl reference times an orszag_tang {REF_SAMPLE_N}x{REF_SAMPLE_N} sample of this "
                                    "workload on all host threads, at every N (rate per cell-update)",
            "note": "E <- S2E(E2S(E)) after every stage, exactly as the 2-D reference's ES_Switch does "
                    "(Weno5_2D.jl:1738,1765); --es-roundtrip 0 selects the 1-D reference's form (keeps E, ~3.5% faster)"}


# ----------------------------------------------------------------------------- GPU arm
def pinned_array(lib, shape):
    from julia_b200 import lib as L
    n = int(np.prod(shape))
    p = C.c_void_p()
    L.check(lib.weno5_host_alloc(C.byref(p), n * 8))
    buf = (C.c_double * n).from_address(p.value)
    a = np.frombuffer(buf, dtype=np.float64).reshape(shape, order="F")
    return a, p


def run_e2e(args, lib, L, s, q, bx, by, gglob, local, world, cells_global, nyl, sync_all, torch, dist):
    """End-to-end arm: host arrays in, host arrays out, every step (see module docstring)."""
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    hq, pq = pinned_array(lib, q.shape)
    hbx, pbx = pinned_array(lib, bx.shape)
    hby, pby = pinned_array(lib, by.shape)
    hq[...] = q
    hbx[...] = bx
    hby[...] = by
    for _ in range(1):                                   # warm the path once
        s.upload(hq, hbx, hby)
        dt = s.timestep()[0]
        s.rk4_step(dt)
        L.check(lib.weno5_download(s.h, L.dptr(hq), L.dptr(hbx), L.dptr(hby)))
    sync_all()
    t0 = time.time()
    for _ in range(e2e_steps):
        s.upload(hq, hbx, hby)
        dt = s.timestep()[0]
        s.rk4_step(dt)
        L.check(lib.weno5_download(s.h, L.dptr(hq), L.dptr(hbx), L.dptr(hby)))
    sync_all()
    e2e_s = time.time() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    bytes_dir = 11 * q.shape[0] * q.shape[1] * 8
    e2e_seq = {"value": cells_global * e2e_steps / e2e_s, "unit": "cell-updates/s", "h2d_bytes_per_step": bytes_dir,
               "d2h_bytes_per_step": bytes_dir, "steps": e2e_steps, "path": "sequential",
               "calls": "weno5_upload -> weno5_timestep -> weno5_rk4_step -> weno5_download, pinned host arrays"}
    e2e = e2e_seq
    if world > 1:
        e2e["path_note"] = ("N > 1: every rank moves its own slab host<->device around each step (sequential calls, NCCL "
                            "halo exchange inside the step); the pipelined single-call path (weno5_pipe_step, the N = 1 "
                            "figure) streams ONE grid through ONE GPU and has no multi-GPU form -- compare this number "
                            "with e2e.sequential of the N = 1 line")

    # single GPU: the pipelined host step (y-slabs with overlap rows; H2D, compute and D2H overlap)
    if world == 1 and nyl >= 1024:
        from julia_b200.solver import Pipe
        rows = 128
        oq, poq = pinned_array(lib, q.shape)
        obx, pobx = pinned_array(lib, bx.shape)
        oby, poby = pinned_array(lib, by.shape)
        dt0 = s.timestep()[0]                             # time step of the state in hq (it is the resident one)
        with Pipe(gglob, device=local, rows_per_slab=rows, es_roundtrip=bool(args.es_roundtrip)) as pipe:
            a_in, a_out = (hq, hbx, hby), (oq, obx, oby)
            _, _, _, dtn = pipe.step(*a_in, dt0, out=a_out)          # warm-up
            a_in, a_out = a_out, a_in
            psteps = max(1, min(args.steps, 2 * args.e2e_steps))
            torch.cuda.synchronize()
            t0 = time.time()
            for _ in range(psteps):
                _, _, _, dtn = pipe.step(*a_in, dtn, out=a_out)
                a_in, a_out = a_out, a_in
            torch.cuda.synchronize()
            pipe_s = time.time() - t0
        nslab = (nyl + rows - 1) // rows
        up_rows = q.shape[1] + 24                        # every public row once + the 24 wrap-around rows sent ahead
        e2e = {"value": cells_global * psteps / pipe_s, "unit": "cell-updates/s",
               "h2d_bytes_per_step": 10 * q.shape[0] * up_rows * 8, "d2h_bytes_per_step": bytes_dir, "steps": psteps,
               "path": "pipelined",
               "calls": "weno5_pipe_step(q0,bxb0,byb0,dt) -> (q4,bxb4,byb4,dt_next): host arrays in, host arrays out, "
                        f"{nslab} y-slabs of {rows}+32 rows on 3 compute streams; the input arrays cross PCIe once into an image "
                        "on the device (upload stream, one event per row range) from which the slabs gather their windows, "
                        "the result rows leave through a result image on a download stream; pinned host arrays; the S plane "
                        "is not uploaded (output only)",
               "sequential": e2e_seq}
        for p_ in (poq, pobx, poby):
            lib.weno5_host_free(p_)
    for p in (pq, pbx, pby):
        lib.weno5_host_free(p)
    return e2e


STRONG_N1_CACHE = os.path.join(ROOT, "gpurun_out", "strong_n1.json")


def run_strong(problem, nx, ny_global, steps, rank, world, local, lib, torch, dist, sync_all, es_roundtrip):
    """BASELINE configs 4/5: a FIXED global grid cut into world y-slabs, timed like the main leg (CUDA events on the
    solver's stream, barrier + synchronize on both sides, max over ranks).  The N = 1 time of the same grid is left
    in gpurun_out/strong_n1.json so that the N > 1 runs of the same session can state their efficiency."""
    from julia_b200 import partition as pt
    from julia_b200 import problems as pr
    from julia_b200.solver import Solver
    nyl = ny_global // world
    ysz = float(ny_global) / nx
    gglob = pr.Grid(nx=nx, ny=ny_global, xsize=1.0, ysize=ysz, bc_x=pr.PERIODIC, bc_y=pr.PERIODIC)
    _, q, bx, by = getattr(pr, problem)(nx, ny=nyl, ysize=float(nyl) / nx, y0=float(rank) * nyl / nx)
    s = Solver(gglob, device=local, ny_local=nyl, j_offset=rank * nyl, es_roundtrip=bool(es_roundtrip))
    try:
        if world > 1:
            uid = pt.broadcast_unique_id(Solver.comm_unique_id, rank, dist)
            s.comm_init(rank, world, uid)
        s.upload(q, bx, by)
        del q, bx, by
        stream = torch.cuda.ExternalStream(lib.weno5_stream(s.h), device=torch.device("cuda", local))
        s.advance(3)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        _, ndone = s.advance(steps)
        e1.record(stream)
        sync_all()
        ms = e0.elapsed_time(e1)
        assert ndone == steps
    finally:
        s.close()
    if world > 1:
        tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    key = f"{problem}_{nx}x{ny_global}"
    out = {"workload": f"{problem} {nx}x{ny_global} global, {nyl} rows per GPU, periodic", "steps": steps, "warmup": 3,
           "ms_per_step": ms / steps, "value": nx * ny_global * steps / (ms * 1e-3), "unit": "cell-updates/s",
           "efficiency_vs_n1": None}
    if rank == 0:
        cache = {}
        try:
            with open(STRONG_N1_CACHE) as f:
                cache = json.load(f)
        except (OSError, ValueError):
            pass
        if world == 1:
            cache[key] = {"ms_per_step": ms / steps, "when": time.time()}
            try:
                os.makedirs(os.path.dirname(STRONG_N1_CACHE), exist_ok=True)
                with open(STRONG_N1_CACHE, "w") as f:
                    json.dump(cache, f)
            except OSError:
                pass
            out["efficiency_vs_n1"] = 1.0
        elif key in cache and time.time() - cache[key].get("when", 0) < 6 * 3600:
            t1 = cache[key]["ms_per_step"]
            out["efficiency_vs_n1"] = t1 / (world * ms / steps)
            out["n1_ms_per_step"] = t1
            out["n1_source"] = "gpurun_out/strong_n1.json, written by the N=1 run of this session"
        else:
            # no N=1 run on this box in this session: the committed N=1 measurement of the same leg (another box of
            # the same pool), labelled as such
            try:
                with open(os.path.join(ROOT, "profiles", "r2_logs", "strong_n1_committed.json")) as f:
                    ref = json.load(f).get(key)
            except (OSError, ValueError):
                ref = None
            if ref:
                out["efficiency_vs_committed_n1"] = ref["ms_per_step"] / (world * ms / steps)
                out["n1_ms_per_step"] = ref["ms_per_step"]
                out["n1_source"] = "profiles/r2_logs/strong_n1_committed.json (N=1 measured in an earlier run, not on this box)"
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world != 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if rank == 0:
        ge.build()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group(backend="cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    from julia_b200 import lib as L
    from julia_b200 import partition as pt
    from julia_b200 import problems as pr
    from julia_b200.solver import Solver
    lib = L.load()

    # N > 1: slab-decomposed results must equal the one-GPU results bit for bit (4 cases, see julia_b200/slabcheck.py)
    slab_check = None
    if world > 1 and not args.no_slab_check:
        from julia_b200 import slabcheck
        slab_check = slabcheck.run_all(rank, world, local, dist)
        dist.barrier()

    nx, nyl = args.nx, args.ny
    gglob = pr.Grid(nx=nx, ny=nyl * world, xsize=1.0, ysize=float(world) * nyl / nx, bc_x=pr.PERIODIC, bc_y=pr.PERIODIC)
    gs, q, bx, by = getattr(pr, args.problem)(nx, ny=nyl, ysize=float(nyl) / nx, y0=float(rank) * nyl / nx)
    s = Solver(gglob, device=local, ny_local=nyl, j_offset=rank * nyl, es_roundtrip=bool(args.es_roundtrip))
    if world > 1:
        uid = pt.broadcast_unique_id(Solver.comm_unique_id, rank, dist)
        s.comm_init(rank, world, uid)
    s.upload(q, bx, by)
    halo_mode = {0: "none (one GPU)", 1: "nccl send/recv, packed messages; nccl all-reduce",
                 2: "peer-memory push (CUDA IPC), one kernel per stage; nccl all-reduce",
                 3: "peer-memory push (CUDA IPC), one kernel per stage, waited for inside the next x sweep; "
                    "all-reduce over peer memory inside the dt kernel"}[s.halo_mode()]
    stream = torch.cuda.ExternalStream(lib.weno5_stream(s.h), device=torch.device("cuda", local))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- device-resident: W warm-up steps, then exactly K timed steps
    s.advance(args.warmup)
    sync_all()
    s.profile(True)
    s.profile_read(reset=True)
    l0 = s.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record(stream)
    t_end, ndone = s.advance(args.steps)
    e1.record(stream)
    sync_all()
    t_wall1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = s.launch_count() - l0
    flux_ms, flux_n = s.profile_read(reset=True)
    s.profile(False)
    assert ndone == args.steps, (ndone, args.steps)
    if world > 1:
        tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    cells_local = nx * nyl
    cells_global = cells_local * world
    value = cells_global * args.steps / (ms * 1e-3)

    # ---- end to end through host buffers (reference-shaped call sequence)
    e2e = None
    if args.e2e_steps > 0:
        e2e = run_e2e(args, lib, L, s, q, bx, by, gglob, local, world, cells_global, nyl, sync_all, torch, dist)


    if rank == 0:
        sampler.stop()
    s.close()

    strong = None
    if args.strong_steps > 0 and (nx, nyl, args.problem) == (NX_PER_GPU, NY_PER_GPU, "orszag_tang"):
        strong = {"config4": run_strong("orszag_tang", 8192, 8192, args.strong_steps, rank, world, local, lib, torch, dist,
                                        sync_all, args.es_roundtrip)}
        if world == 8:
            strong["config5"] = run_strong("blast", 16384, 16384, args.strong_steps, rank, world, local, lib, torch, dist,
                                           sync_all, args.es_roundtrip)

    if rank == 0:
        clocks = sampler.summary(t_wall0, t_wall1)
        peak, how = measured_peaks()
        flux_avg_ms = flux_ms / max(flux_n, 1)
        algo_bytes_launch = ALGO_BYTES_PER_CELL_UPDATE / FLUX_LAUNCHES_PER_STEP * cells_local
        achieved = algo_bytes_launch / (flux_avg_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "flux_kernel_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        step_traffic, pipe_pct = None, None
        if os.path.exists(tp):
            with open(tp) as f:
                tj = json.load(f)
            step_traffic, pipe_pct = tj.get("step_dram_bytes"), tj.get("fp64_pipe_pct")
        fp64_tflops = ALGO_FLOPS_PER_CELL_UPDATE * cells_local * args.steps / (ms * 1e-3) / 1e12
        fp64_peak_clock = (FP64_DFMA_PER_CLK * 2 * clocks["sm_mhz"] * 1e6 / 1e12) if clocks.get("sm_mhz") else None
        line = {
            "metric": "cell_updates_per_sec", "value": value, "unit": "cell-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world) if (nx, nyl, args.problem) == (NX_PER_GPU, NY_PER_GPU, "orszag_tang") else
            {"workload": f"{args.problem} {nx}x{nyl} per GPU (non-default workload)", "nx": nx, "ny_global": nyl * world},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": "profiles/flux_kernel_traffic.json: dram__bytes_read+write per "
                                                               "launch of the committed ncu --set full capture (not of this run)",
                         "peak_source": how, "kernel": "flux_ws_kernel<0|1> (x and y sweeps)",
                         "avg_launch_ms": flux_avg_ms, "launches_timed": int(flux_n),
                         "algorithmic_bytes_per_launch": algo_bytes_launch,
                         "share_of_step": flux_ms / ms,
                         "whole_step_frac": ALGO_BYTES_PER_CELL_UPDATE * cells_local * args.steps / (ms * 1e-3) / 1e9 / peak,
                         "step_traffic": step_traffic,
                         "fp64": {"achieved_tflops": fp64_tflops, "peak_tflops": FP64_PEAK_TFLOPS_MEASURED,
                                  "peak_source": "measured, profiles/fp64_peak.txt",
                                  "peak_tflops_at_sampled_clock": fp64_peak_clock,
                                  "frac": fp64_tflops / FP64_PEAK_TFLOPS_MEASURED,
                                  "pipe_pct_ncu": pipe_pct,
                                  "note": "FP64 pipe is the binding roof (SURVEY 8d): 22 kflop vs 1048 B per cell-update; "
                                          "achieved/peak/frac above are the HBM figures the metric asks for; pipe_pct_ncu = "
                                          "sm__inst_executed_pipe_fp64 of the committed ncu capture (x, y sweep)"}},
            "sim": {"t_end": t_end},
            "halo_exchange": halo_mode,
        }
        if strong is not None:
            line["strong"] = strong
        if slab_check is not None:
            line["slab_bit_identical"] = slab_check["slab_bit_identical"]
            line["slab_check"] = slab_check
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline()
        else:
            line["cpu_baseline"] = None
        if world == 1 and args.ext3d_n > 0:
            try:
                line["extension_3d"] = run_ext3d(args.ext3d_n, 5, not args.no_cpu)
            except Exception as e:                       # the headline line must not depend on the extension
                line["extension_3d"] = {"error": f"{type(e).__name__}: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=NX_PER_GPU)
    ap.add_argument("--ny", type=int, default=NY_PER_GPU, help="rows per GPU")
    ap.add_argument("--problem", default="orszag_tang", choices=["orszag_tang", "blast"],
                    help="synthetic workload (BASELINE configs 3/4: orszag_tang, config 5: blast)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--strong-steps", type=int, default=4, help="timed steps of the strong-scaling leg (0: skip)")
    ap.add_argument("--ext3d-n", type=int, default=192, help="N = 1: also time the 3-D extension on n^3 cells (0: skip)")
    ap.add_argument("--no-slab-check", action="store_true", help="N > 1: skip the slab bit-identity pre-check")
    ap.add_argument("--es-roundtrip", type=int, default=1, help="1: E <- S2E(E2S(E)) after every stage (2-D reference form)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
