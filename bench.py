#!/usr/bin/env python
"""
bench.py -- M DG-dof updates/s of the explicit P1DG shallow-water SSPRK33 step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1..5] [--k 19] [--no-wd]

Default workload = BASELINE.json config 5, the configuration the metric is quoted on: the reference's North Sea mesh
k-sectioned to 3 942 120 triangles (35.5 M dofs), nonlinear SWE, Lax-Friedrichs, Manning drag, Coriolis, tidal
elevation on the open boundary, wetting-drying on; synthetic bathymetry / initial state (harness/workloads.py).
One "step" = one SSPRK33 step = 3 fused stage-kernel launches over every triangle.  `--config 1..4` runs the other
BASELINE configurations through the same driver (their lines are committed under profiles/).

One JSON line on stdout (rank 0); everything else goes to stderr.
`value`: state resident in HBM, CUDA-event timed, max over ranks.
`e2e`: the same metric through the reference-facing API (FlowSolver2d mirror -> SSPRK33.advance(t, update_forcings))
with host-side forcings: every stage the tidal elevation is computed on the host and copied H2D from pinned memory,
every step the print_state norms are reduced on the device and read back D2H.
`no_wd`: config 5 only -- the same measurement with wetting-drying off (the explicit W&D step is a defined extension
without a reference code path; both figures are reported side by side).
`parity` (N > 1): before anything is timed, a small instance of the same workload is advanced 3 steps on the
distributed mesh and on rank 0 alone and the owned records are compared bit for bit; the run aborts on a mismatch.
`--impl reference`: the CPU restatement of the reference (oracle/swe_oracle.c, OpenMP, all host threads) on the
same workload; Firedrake itself cannot be installed here.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_BASE = 228          # B / triangle-stage, SURVEY.md 8d / BASELINE.md section 3
METRIC = "M DG-dof updates/sec (2D SWE SSPRK33)"
# extra P1 coefficient columns (4 B / triangle-stage each) and the launched specialisation, per BASELINE config
CONFIG_KERNEL = {1: (0, "swe_stage_kernel<1, 1>"), 2: (0, "swe_stage_kernel<0, 1>"), 3: (12, "swe_stage_kernel<0, 4>"),
                 4: (0, "swe_stage_kernel<1, 1>"), 5: (8, "swe_stage_kernel<1, 3>")}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=5, choices=[1, 2, 3, 4, 5], help="BASELINE.json configuration")
    ap.add_argument("--k", type=int, default=19, help="config 5: k-section refinement of the 10 920-triangle North Sea mesh")
    ap.add_argument("--no-wd", action="store_true", help="config 5: wetting-drying off in the main measurement")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-second-leg", action="store_true", help="config 5: skip the wetting-drying-off leg (`no_wd`)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the bit-identity check against rank 0 alone")
    ap.add_argument("--no-overlap", action="store_true", help="multi-GPU, unfused transport: no boundary / interior overlap")
    ap.add_argument("--no-fused", action="store_true", help="multi-GPU: boundary launch + push kernel + barrier instead of "
                    "the fused compute + halo-push launch")
    ap.add_argument("--no-graph", action="store_true", help="do not capture the resident step in a CUDA graph")
    ap.add_argument("--e2e-graph", default="step", choices=["step", "stage"], help="e2e path: one CUDA graph per step "
                    "(boundary data of the three stages in three banks) or one per RK stage")
    ap.add_argument("--no-l2-flush", action="store_true", help="never flush L2 between steps (default: flush when the "
                    "per-GPU working set is smaller than 1.5 x L2)")
    ap.add_argument("--transport", default="auto", choices=["auto", "nccl", "symm"], help="multi-GPU halo transport")
    return ap.parse_args()


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the load phase (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.samples:
            if ts < t0 or ts > t1 + 0.1:
                continue
            w = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(w[0]))
                mx.append(float(w[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), w[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ CPU baseline / reference arm
def make_cpu_oracle(cfg, k, wd):
    """oracle/swe_oracle.c set up for the SWE part of BASELINE config `cfg`; returns (oracle, records, dt, mesh, note)."""
    from oracle.c_oracle import COracle, records_from_nodal, host_threads
    from harness.workloads import north_sea_mesh, north_sea_setup, tide_values, config_mesh, _cfl_dt
    nthreads = host_threads()
    note = ""
    if cfg == 5:
        mesh = north_sea_mesh(k)
        setup = north_sea_setup(mesh, wetting_drying=wd)
        co = COracle(mesh, setup["bath"], nonlinear=True, lf_on=True, coriolis=setup["coriolis"], manning=setup["manning"],
                     bnd={100: {"elev": 0.0, "uv": (0.0, 0.0)}}, bf_elev=tide_values(setup, 0.0), wd_on=wd,
                     wd_alpha=setup["wd_alpha"], threads=nthreads)
        return co, records_from_nodal(setup["uv0"], setup["eta0"]), setup["dt"], mesh, note
    mesh = config_mesh(cfg)
    x = mesh.coords[mesh.cells]
    X, Y = x[..., 0], x[..., 1]
    uv = np.zeros(x.shape)
    if cfg == 1:
        co = COracle(mesh, 20.0, nonlinear=True, lf_on=True, threads=nthreads)
        eta, dt = 2.0 * np.exp(-((X - 20e3) / 4e3) ** 2), _cfl_dt(mesh, 20.0, 0.1)
    elif cfg == 2:
        L = 44294.46
        co = COracle(mesh, 50.0, nonlinear=False, threads=nthreads)
        eta, dt = -np.cos(2 * np.pi * X / L), _cfl_dt(mesh, 50.0, 0.0)
    elif cfg == 3:
        L = 1.0e6
        yv = mesh.coords[:, 1]
        co = COracle(mesh, 1000.0, nonlinear=False, coriolis=1.0e-4 + 2.0e-11 * yv, linear_drag=1.0e-6,
                     wind_stress=np.stack([0.1 * np.sin(np.pi * (yv / L - 0.5)), 0.0 * yv], -1), threads=nthreads)
        eta, dt = 1.0e-3 * np.sin(2 * np.pi * X / L) * np.sin(np.pi * Y / L), _cfl_dt(mesh, 1000.0, 0.0)
    else:
        co = COracle(mesh, 1.0, nonlinear=True, lf_on=True, threads=nthreads)
        uv = np.stack([0.5 - Y, X - 0.5], -1)
        eta, dt = 0.0 * X, _cfl_dt(mesh, 1.0, 0.75)
        note = " (SWE part only: the tracer equation and the limiter are not in the C port)"
    return co, records_from_nodal(uv, eta), dt, mesh, note


def cpu_port(cfg, k, wd, target_seconds, max_steps=50):
    """Times oracle/swe_oracle.c (C + OpenMP, all host threads) on a bounded number of SSPRK33 steps."""
    co, rec, dt, mesh, note = make_cpu_oracle(cfg, k, wd)
    t0 = time.perf_counter()
    co.ssprk33(rec, dt, 1)                 # warm-up step (page faults, thread start)
    t1 = time.perf_counter() - t0
    n = int(max(1, min(max_steps, round(target_seconds / max(t1, 1e-3)))))
    t0 = time.perf_counter()
    co.ssprk33(rec, dt, n)
    el = time.perf_counter() - t0
    val = 9.0 * mesh.n_cells * n / el / 1e6
    cores = co.threads()
    assert cores > 1 or (os.cpu_count() or 1) == 1, "CPU baseline must use every host core"
    return val, cores, n, el, mesh.n_cells, note


def workload_name(a, wd):
    from harness.workloads import CONFIG_NAMES
    if a.config == 5:
        return (f"config 5: north_sea.msh k={a.k} refined, nonlinear SWE + LF + Manning + Coriolis + tidal elev BC, "
                f"wetting_drying={'on' if wd else 'off'}")
    return f"config {a.config}: " + CONFIG_NAMES[a.config]


def reference_arm(a):
    wd = not a.no_wd
    co, rec, dt, mesh, note = make_cpu_oracle(a.config, a.k, wd)
    # bounded sample: each timed "step" advances the whole mesh by one SSPRK33 step; K capped so the run ends in minutes
    t0 = time.perf_counter()
    co.ssprk33(rec, dt, 1)
    t_one = time.perf_counter() - t0
    W = min(a.warmup, 2)
    K = int(max(1, min(a.steps, round(120.0 / max(t_one, 1e-3)))))
    co.ssprk33(rec, dt, W)
    t0 = time.perf_counter()
    co.ssprk33(rec, dt, K)
    el = time.perf_counter() - t0
    val = 9.0 * mesh.n_cells * K / el / 1e6
    cores = co.threads()
    assert cores > 1 or (os.cpu_count() or 1) == 1, "reference arm must use every host core"
    sample = f"{K} SSPRK33 steps of the full {mesh.n_cells}-triangle workload{note}"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "M dof-updates/s", "n_gpus": a.gpus,
            "steps": K, "warmup": W, "ms_per_step": el / K * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a, wd), "triangles": int(mesh.n_cells), "dofs": int(9 * mesh.n_cells),
                       "note": "CPU restatement of the reference discretisation (oracle/swe_oracle.c, OpenMP, "
                               f"{cores} threads); Firedrake/PETSc cannot be installed offline"},
            "cpu_baseline": {"value": val, "unit": "M dof-updates/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "M dof-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ runs
def make_run(a, rank, world, wd, small=False):
    """The workload driver (harness/runs.py).  ``small``: the reduced instance used by the N > 1 parity check."""
    from harness.workloads import north_sea_mesh, north_sea_setup
    fused = not a.no_fused
    if a.config == 5:
        mesh = north_sea_mesh(3 if small else a.k)
        setup = north_sea_setup(mesh, wetting_drying=wd)
        if world > 1:
            from harness.runs import PartitionedSWE
            run = PartitionedSWE(mesh, setup, rank, world, wd=wd, transport=a.transport, overlap=not a.no_overlap,
                                 fused=fused)
        else:
            from harness.runs import SingleSWE
            run = SingleSWE(mesh, setup, wd=wd)
        run.n_global = mesh.n_cells
        run.dofs_per_cell = 9
        return run
    from harness.runs import ConfigRun
    scale = {1: 1.0, 2: 0.125, 3: 0.1, 4: 0.06}[a.config] if small else 1.0
    return ConfigRun(a.config, rank, world, scale=scale, transport=a.transport, fused=fused)


def parity_check(a, rank, world, wd):
    """N > 1: 3 steps of a small instance, distributed vs rank 0 alone, owned records compared bit for bit."""
    import torch
    import torch.distributed as dist
    nsteps = 3
    drun = make_run(a, rank, world, wd, small=True)
    for _ in range(nsteps):
        drun.step_e2e()
    torch.cuda.synchronize()
    fields = drun.owned_nodal()
    owned = drun.part.owned_global.copy()
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((owned, [np.ascontiguousarray(f) for f in fields]), gathered, dst=0)
    res = None
    if rank == 0:
        srun = make_run(a, 0, 1, wd, small=True)      # rank 0 alone: same workload, world = 1
        for _ in range(nsteps):
            srun.step_e2e()
        torch.cuda.synchronize()
        ref = srun.state_nodal() if hasattr(srun, "state_nodal") else srun.owned_nodal()
        md, ncells = 0.0, 0
        for own, fl in gathered:
            ncells += own.shape[0]
            for f, r in zip(fl, ref):
                md = max(md, float(np.abs(f - r[own]).max()))
        ok = md == 0.0 and ncells == srun.n_global and all(np.isfinite(r).all() for r in ref)
        res = {"n_gpu_bit_identical": bool(ok), "max_abs_diff": md, "steps": nsteps, "triangles": int(srun.n_global),
               "what": f"{world}-rank run vs rank 0 alone, every owned cell record after {nsteps} SSPRK33 steps "
                       "through the reference-facing API"}
        del srun
    flag = torch.tensor([1 if (res is None or res["n_gpu_bit_identical"]) else 0], device="cuda")
    dist.broadcast(flag, src=0)
    del drun
    torch.cuda.empty_cache()
    if int(flag.item()) != 1:
        raise RuntimeError(f"multi-GPU parity check FAILED: {res}")
    return res


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wd = not a.no_wd
    if a.impl == "reference":
        if rank == 0:
            reference_arm(a)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from thetis_b200 import solver2d
    from thetis_b200.build import build_library
    solver2d.print_output = log            # stdout carries the JSON line only
    if rank == 0:
        build_library()
    if world > 1:
        dist.barrier()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    parity = None
    if world > 1 and not a.no_parity:
        parity = parity_check(a, rank, world, wd)
        barrier()

    l2_bytes = int(getattr(torch.cuda.get_device_properties(local_rank), "L2_cache_size", 126 * 2 ** 20))
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)

    def measure(run, steps, with_clocks):
        """Resident timing of `run.step_resident` (CUDA events on the launching stream, max over ranks)."""
        if not a.no_graph and hasattr(run, "enable_graph"):
            run.enable_graph()
        sampler = ClockSampler(local_rank)
        if rank == 0 and with_clocks:
            sampler.start()
        for _ in range(max(a.warmup, 3)):
            run.step_resident()
        barrier()
        # load phase so that the clock samples see the same kernel mix even when K is small
        t_load0 = time.time()
        burn_until = time.time() + 1.5
        while True:
            for _ in range(10):
                run.step_resident()
            torch.cuda.synchronize()
            # EVERY rank must issue the same number of steps (the fused halo exchange counts launches; a rank that
            # leaves this time-based loop one iteration earlier than its peers would starve them): rank 0 decides
            go = torch.tensor([1 if time.time() < burn_until else 0], dtype=torch.int32, device="cuda")
            if world > 1:
                dist.broadcast(go, src=0)
            if int(go.item()) == 0:
                break
        barrier()
        if world > 1 and getattr(run, "plan", None) is not None and run.plan.fused:
            ep, err = run.eng.halo_fused_status()
            if err:
                raise RuntimeError("a halo flag wait timed out during warm-up: the ranks did not issue the same "
                                   "sequence of fused launches")
        l0 = run.launches()
        # L2 policy (timing rules): when the per-GPU working set (3 rotating state arrays + the static patch blocks)
        # is not clearly larger than L2 (N = 8: 3 x 36 MB), L2 is flushed between timed steps by writing a buffer
        # twice the L2 size, and every step is timed on its own with CUDA events (the flush is not timed).
        work_bytes = 3 * run.n_owned() * 8 * run.dofs_per_cell + run.n_owned() * 40
        flush = (work_bytes < 1.5 * l2_bytes) and not a.no_l2_flush
        flush_buf = torch.empty(2 * l2_bytes // 8, dtype=torch.float64, device="cuda") if flush else None

        def timed(step_fn, nsteps):
            if not flush:
                e0.record()
                for _ in range(nsteps):
                    step_fn()
                e1.record()
                barrier()
                ms_ = e0.elapsed_time(e1)
            else:
                evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nsteps)]
                for a_, b_ in evs:
                    flush_buf.fill_(0.0)
                    a_.record()
                    step_fn()
                    b_.record()
                barrier()
                ms_ = sum(a_.elapsed_time(b_) for a_, b_ in evs)
            if world > 1:
                t = torch.tensor([ms_], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_ = float(t.item())
            return ms_

        ms = timed(run.step_resident, steps)
        t_load1 = time.time()
        launches = run.launches() - l0
        clocks = sampler.stop(t_load0, t_load1) if (rank == 0 and with_clocks) else None
        return ms, launches, clocks, timed, work_bytes, flush

    run = make_run(a, rank, world, wd)
    n_tri_global = run.n_global
    dpc = run.dofs_per_cell
    ms, launches, clocks, timed, work_bytes, flush = measure(run, a.steps, True)
    value = float(dpc) * n_tri_global * a.steps / (ms * 1e-3) / 1e6

    # the stage kernel's average launch duration: the timed region holds only SWE stage launches (+ halo waits when
    # N > 1) except in config 4, where the SWE part of the step is timed on its own right after
    if a.config == 4:
        sw = run.swe
        for _ in range(3):
            sw.advance_device()
        barrier()
        e0.record()
        for _ in range(a.steps):
            sw.advance_device()
        e1.record()
        barrier()
        kernel_ms = e0.elapsed_time(e1) / (3 * a.steps)
    else:
        kernel_ms = ms / (3 * a.steps)

    # ---------------- end-to-end through the reference-facing API
    e2e = None
    if not a.no_e2e:
        run.use_fused_norms(True)
        if not a.no_graph and a.e2e_graph == "step" and hasattr(run, "enable_step_graph"):
            run.enable_step_graph()        # one graph per step, the forcing of stage i in bank i of the boundary arrays
        elif not a.no_graph and hasattr(run, "enable_stage_graphs"):
            run.enable_stage_graphs()
        for _ in range(max(a.warmup, 3)):
            run.step_e2e()
        barrier()
        ms2 = timed(run.step_e2e, a.steps)
        e2e = {"value": float(dpc) * n_tri_global * a.steps / (ms2 * 1e-3) / 1e6, "unit": "M dof-updates/s",
               "h2d_bytes_per_step": int(run.h2d_bytes_per_step()), "d2h_bytes_per_step": int(run.d2h_bytes_per_step()),
               "ms_per_step": ms2 / a.steps,
               "path": run.e2e_path()}
    n_tri_local = run.n_owned()
    transport, overlap, fusedp = getattr(run, "transport", None), getattr(run, "overlap", None), None
    if world > 1:
        fusedp = bool(run.plan.fused)
        if fusedp:
            ep, err = run.eng.halo_fused_status()
            if err:
                raise RuntimeError("a halo flag wait timed out during the run")
    dt = run.dt
    graph_on = hasattr(run, "_graph") or hasattr(run, "_graphs")

    # ---------------- config 5: the same measurement with wetting-drying off, reported side by side
    no_wd = None
    if a.config == 5 and wd and not a.no_second_leg:
        del run
        torch.cuda.empty_cache()
        run2 = make_run(a, rank, world, False)
        k2 = max(10, min(a.steps, 50))
        ms_b, _, _, _, _, _ = measure(run2, k2, False)
        no_wd = {"value": 9.0 * n_tri_global * k2 / (ms_b * 1e-3) / 1e6, "unit": "M dof-updates/s", "steps": k2,
                 "ms_per_step": ms_b / k2, "kernel": "swe_stage_kernel<1, 2>",
                 "note": "wetting-drying off: the configuration whose every term has a reference code path"}
        del run2
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (the fused stage kernel)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    extra, kname = CONFIG_KERNEL[a.config]
    if a.config == 5 and not wd:
        kname = "swe_stage_kernel<1, 2>"
    alg_bytes = ALG_BYTES_BASE + extra          # + 4 B / triangle-stage per P1 coefficient column
    achieved = alg_bytes * n_tri_local / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": kname, "algorithmic_bytes_per_triangle_stage": alg_bytes,
                "triangles_per_launch": int(n_tri_local), "avg_launch_ms": kernel_ms, "peak_source": peak_src,
                "note": "the kernel is fp64-pipe bound and clock sensitive: `value` is timed right after a 1.5 s burn "
                        "under the board power cap (SM clock in `clocks`), `e2e` later in the run"}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if a.config == 5 and world == 1 and os.path.exists(tr):
        try:
            tj = json.load(open(tr))
            roofline["traffic"] = tj.get("dram_bytes_per_launch")
            roofline["traffic_source"] = tj.get("source", "profiles/traffic.json (ncu --set full capture of this kernel, not this run)")
        except Exception:
            pass
    if no_wd is not None:
        no_wd["roofline_frac"] = alg_bytes * n_tri_local / (no_wd["ms_per_step"] / 3 * 1e-3) / 1e9 / peak

    cpu = None
    if not a.no_cpu_baseline and world == 1:
        val, cores, n, el, ntri, note = cpu_port(a.config, a.k, wd, a.cpu_seconds)
        cpu = {"value": val, "unit": "M dof-updates/s", "cores": cores, "kind": "port",
               "sample": f"{n} SSPRK33 steps of the full {ntri}-triangle workload{note} ({el:.1f} s, oracle/swe_oracle.c, OpenMP)"}

    par = "single GPU"
    if world > 1:
        par = (f"domain decomposition x{world}, halo transport {transport}, "
               + ("fused compute + halo-push launch (per-peer epoch flags)" if fusedp else f"push kernel + barrier, overlap {overlap}")
               + f", cuda graph {graph_on}")
    line = {"metric": METRIC, "value": value, "unit": "M dof-updates/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a, wd), "triangles": int(n_tri_global), "dofs": int(dpc * n_tri_global),
                       "dt": dt,
                       "wetting_drying_step": ("plain-mass explicit extension (DESIGN.md section 6): the reference's depth "
                                               "formulation, no reference code path for the step; `no_wd` is the leg "
                                               "whose every term and step has one") if (a.config == 5 and wd) else None,
                       "l2": ("per-GPU working set %.0f MB vs L2 %.0f MB: " % (work_bytes / 1e6, l2_bytes / 1e6))
                             + ("L2 flushed between timed steps (2 x L2 buffer written, untimed), steps timed one by one"
                                if flush else "inputs larger than L2, no flush"),
                       "parallelism": par},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "stage_kernel_launches": int(3 * a.steps),
            "roofline": roofline, "cpu_baseline": cpu, "no_wd": no_wd, "parity": parity}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
