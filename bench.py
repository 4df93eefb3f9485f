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
roofline  the dominant kernel is flux_kernel<DIR> (8 launches per step); algorithmic bytes per
        launch = 1048 B/cell-update / 8 launches * cells; duration from CUDA events
        recorded around every flux launch inside the timed region.
cpu_baseline  the C oracle (oracle/, "port": the Julia reference cannot run here), OpenMP on
        all host cores, on a bounded sample (rank 0, N=1 only).
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
FP64_PEAK_TFLOPS_NOMINAL = 37.2              # 148 SM x 64 DFMA/clk x 1.965 GHz


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


def run_reference(args):
    """The reference's own CPU implementation cannot run natively (Julia 0.6 source, no julia binary): this
    arm times the oracle port of it (bit-identical numerics, tests/test_reference_pin.py) on all host cores, on a
    bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_c as O
    O.build()
    rate0, _ = oracle_rate(256, 1)
    total_steps = args.steps + args.warmup
    budget = 90.0
    n = int(max(256, min(2048, (rate0 * budget / max(total_steps, 1)) ** 0.5 // 64 * 64)))
    from julia_b200 import problems as pr
    g, q, bx, by = pr.orszag_tang(n)
    q, bx, by, _ = O.advance_2d(g, q, bx, by, args.warmup, fast=True)
    t0 = time.time()
    O.advance_2d(g, q, bx, by, args.steps, fast=True)
    dt = time.time() - t0
    value = n * n * args.steps / dt
    cores = O.num_threads(True)
    sample = f"orszag_tang {n}x{n} (sample of the 4096x4096-per-GPU workload), {args.steps} RK4 steps incl. timestep"
    line = {
        "impl": "reference", "metric": "cell_updates_per_sec", "value": value, "unit": "cell-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "no julia binary exists here (Julia 0.6 source); the reference only runs through the mechanical "
                "translation used for the golden vectors (scripts/make_reference_golden.py, ~1e4x slower than native "
                "Julia, not a meaningful timing).  Timed instead: the C port of its E-branch numerics, which "
                "tests/test_reference_pin.py shows to be bit-identical to it, OpenMP over lines, all host threads",
    }
    print(json.dumps(line), flush=True)


def workload_config(ngpu):
    return {"workload": f"orszag_tang {NX_PER_GPU}x{NY_PER_GPU} cells per GPU, global {NX_PER_GPU}x{NY_PER_GPU * ngpu}, "
                        "periodic, gamma=5/3, CFL 0.8, dt recomputed every step",
            "nx": NX_PER_GPU, "ny_global": NY_PER_GPU * ngpu, "partition": f"y-slabs x{ngpu}",
            "l2": "no flush needed: ~6.5 GB of resident planes per GPU >> 126 MB L2",
            "es_roundtrip": 1,
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
               "d2h_bytes_per_step": bytes_dir, "steps": e2e_steps,
               "calls": "weno5_upload -> weno5_timestep -> weno5_rk4_step -> weno5_download, pinned host arrays"}
    e2e = e2e_seq

    # single GPU: the pipelined host step (y-slabs with overlap rows; H2D, compute and D2H overlap)
    if world == 1 and nyl >= 1024:
        from julia_b200.solver import Pipe
        rows = 512
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
        up_rows = nslab * (rows + 2 * 16 + 8)
        e2e = {"value": cells_global * psteps / pipe_s, "unit": "cell-updates/s",
               "h2d_bytes_per_step": 10 * q.shape[0] * up_rows * 8, "d2h_bytes_per_step": bytes_dir, "steps": psteps,
               "calls": "weno5_pipe_step(q0,bxb0,byb0,dt) -> (q4,bxb4,byb4,dt_next): host arrays in, host arrays out, "
                        f"{nslab} y-slabs of {rows}+32 rows on 3 streams (copies and compute overlap), pinned host arrays; "
                        "the S plane is not uploaded (output only)",
               "sequential": e2e_seq}
        for p_ in (poq, pobx, poby):
            lib.weno5_host_free(p_)
    for p in (pq, pbx, pby):
        lib.weno5_host_free(p)
    return e2e


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

    nx, nyl = args.nx, args.ny
    gglob = pr.Grid(nx=nx, ny=nyl * world, xsize=1.0, ysize=float(world) * nyl / nx, bc_x=pr.PERIODIC, bc_y=pr.PERIODIC)
    gs, q, bx, by = getattr(pr, args.problem)(nx, ny=nyl, ysize=float(nyl) / nx, y0=float(rank) * nyl / nx)
    s = Solver(gglob, device=local, ny_local=nyl, j_offset=rank * nyl, es_roundtrip=bool(args.es_roundtrip))
    if world > 1:
        uid = pt.broadcast_unique_id(Solver.comm_unique_id, rank, dist)
        s.comm_init(rank, world, uid)
    s.upload(q, bx, by)
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
        fp64_tflops = ALGO_FLOPS_PER_CELL_UPDATE * cells_local * args.steps / (ms * 1e-3) / 1e12
        line = {
            "metric": "cell_updates_per_sec", "value": value, "unit": "cell-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world) if (nx, nyl, args.problem) == (NX_PER_GPU, NY_PER_GPU, "orszag_tang") else
            {"workload": f"{args.problem} {nx}x{nyl} per GPU (non-default workload)", "nx": nx, "ny_global": nyl * world},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": how, "kernel": "flux_kernel<0|1> (x and y sweeps)",
                         "avg_launch_ms": flux_avg_ms, "launches_timed": int(flux_n),
                         "algorithmic_bytes_per_launch": algo_bytes_launch,
                         "share_of_step": flux_ms / ms,
                         "whole_step_frac": ALGO_BYTES_PER_CELL_UPDATE * cells_local * args.steps / (ms * 1e-3) / 1e9 / peak,
                         "fp64": {"achieved_tflops": fp64_tflops, "peak_tflops_nominal": FP64_PEAK_TFLOPS_NOMINAL,
                                  "frac": fp64_tflops / FP64_PEAK_TFLOPS_NOMINAL,
                                  "note": "FP64 pipe is the binding roof (SURVEY 8d): 22 kflop vs 1048 B per cell-update"}},
            "sim": {"t_end": t_end},
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline()
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    s.close()
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
