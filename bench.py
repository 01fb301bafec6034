#!/usr/bin/env python
"""
bench.py -- M DG-dof updates/s of the explicit P1DG shallow-water SSPRK33 step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--k 19] [--no-wd]

Workload (BASELINE.json config 5, the configuration the metric is quoted on): the reference's North Sea mesh
k-sectioned to 3 942 120 triangles (35.5 M dofs), nonlinear SWE, Lax-Friedrichs, Manning drag, Coriolis, tidal
elevation on the open boundary, wetting-drying on; synthetic bathymetry / initial state (thetis_b200/workloads.py).
One "step" = one SSPRK33 step = 3 fused stage-kernel launches over every triangle.

One JSON line on stdout (rank 0).  `value`: state resident in HBM, CUDA-event timed, max over ranks.
`e2e`: the same metric through the reference-facing API (FlowSolver2d mirror -> SSPRK33.advance(t, update_forcings))
with host-side forcings: every stage the tidal elevation is computed on the host and copied H2D from pinned memory,
every step the print_state norms are reduced on the device and read back D2H.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--k", type=int, default=19, help="k-section refinement of the 10 920-triangle North Sea mesh")
    ap.add_argument("--no-wd", action="store_true", help="switch wetting-drying off (config 5 has it on)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="multi-GPU: do not overlap the halo push with interior patches")
    ap.add_argument("--no-graph", action="store_true", help="do not capture the resident step in a CUDA graph")
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
def cpu_port(mesh, setup, wd, target_seconds, tide):
    """Times oracle/swe_oracle.c (C + OpenMP, all host threads) on a bounded number of SSPRK33 steps."""
    from oracle.c_oracle import COracle, records_from_nodal
    co = COracle(mesh, setup["bath"], nonlinear=True, lf_on=True, coriolis=setup["coriolis"], manning=setup["manning"],
                 bnd={100: {"elev": 0.0, "uv": (0.0, 0.0)}}, bf_elev=tide, wd_on=wd, wd_alpha=setup["wd_alpha"])
    rec = records_from_nodal(setup["uv0"], setup["eta0"])
    dt = setup["dt"]
    t0 = time.perf_counter()
    co.ssprk33(rec, dt, 1)                 # warm-up step (page faults, thread start)
    t1 = time.perf_counter() - t0
    n = int(max(1, min(50, round(target_seconds / max(t1, 1e-3)))))
    t0 = time.perf_counter()
    co.ssprk33(rec, dt, n)
    el = time.perf_counter() - t0
    val = 9.0 * mesh.n_cells * n / el / 1e6
    return val, co.threads(), n, el


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wd = not a.no_wd
    from thetis_b200.workloads import north_sea_mesh, north_sea_setup, tide_values

    workload = f"north_sea.msh k={a.k} refined, nonlinear SWE + LF + Manning + Coriolis + tidal elev BC, wetting_drying={'on' if wd else 'off'}"
    if a.impl == "reference":
        if rank != 0:
            return
        mesh = north_sea_mesh(a.k)
        setup = north_sea_setup(mesh, wetting_drying=wd)
        tide = tide_values(setup, 0.0)
        from oracle.c_oracle import COracle, records_from_nodal
        co = COracle(mesh, setup["bath"], coriolis=setup["coriolis"], manning=setup["manning"],
                     bnd={100: {"elev": 0.0, "uv": (0.0, 0.0)}}, bf_elev=tide, wd_on=wd, wd_alpha=setup["wd_alpha"])
        rec = records_from_nodal(setup["uv0"], setup["eta0"])
        dt = setup["dt"]
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
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "M dof-updates/s", "n_gpus": a.gpus,
                "steps": K, "warmup": W, "ms_per_step": el / K * 1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "triangles": int(mesh.n_cells), "dofs": int(9 * mesh.n_cells),
                           "note": "CPU restatement of the reference discretisation (oracle/swe_oracle.c, OpenMP); "
                                   "Firedrake/PETSc cannot be installed offline"},
                "cpu_baseline": {"value": val, "unit": "M dof-updates/s", "cores": co.threads(), "kind": "port",
                                 "sample": f"{K} SSPRK33 steps of the full {mesh.n_cells}-triangle workload"},
                "e2e": {"value": val, "unit": "M dof-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from thetis_b200 import _lib as L
    from thetis_b200.build import build_library
    if rank == 0:
        build_library()
    if world > 1:
        dist.barrier()

    mesh = north_sea_mesh(a.k)
    setup = north_sea_setup(mesh, wetting_drying=wd)
    n_tri_global = mesh.n_cells
    dt = setup["dt"]

    if world > 1:
        from thetis_b200.parallel import PartitionedSWE
        run = PartitionedSWE(mesh, setup, rank, world, wd=wd, transport=a.transport, overlap=not a.no_overlap)
    else:
        from thetis_b200.parallel import SingleSWE
        run = SingleSWE(mesh, setup, wd=wd)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing
    if not a.no_graph and hasattr(run, "enable_graph"):
        run.enable_graph()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(a.warmup, 3)):
        run.step_resident()
    barrier()
    # load phase so that the clock samples see the same kernel mix even when K is small
    t_load0 = time.time()
    burn_until = time.time() + 1.5
    while time.time() < burn_until:
        for _ in range(10):
            run.step_resident()
        torch.cuda.synchronize()
    barrier()
    l0 = run.launches()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    # L2 policy (timing rules): the per-GPU working set (3 rotating state arrays + the static patch blocks) is
    # larger than L2 at N <= 4; when it is not (N = 8: 3 x 36 MB), L2 is flushed between timed steps by writing a
    # buffer twice the L2 size, and every step is timed on its own with CUDA events (the flush is not timed).
    l2_bytes = int(getattr(torch.cuda.get_device_properties(local_rank), "L2_cache_size", 126 * 2 ** 20))
    work_bytes = 3 * run.n_owned() * 72 + run.n_owned() * 40
    flush = (work_bytes < 1.5 * l2_bytes) and not a.no_l2_flush
    flush_buf = torch.empty(2 * l2_bytes // 8, dtype=torch.float64, device="cuda") if flush else None

    def timed(step_fn, nsteps):
        if not flush:
            e0.record()
            for _ in range(nsteps):
                step_fn()
            e1.record()
            barrier()
            return e0.elapsed_time(e1)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nsteps)]
        for a_, b_ in evs:
            flush_buf.fill_(0.0)
            a_.record()
            step_fn()
            b_.record()
        barrier()
        return sum(a_.elapsed_time(b_) for a_, b_ in evs)

    ms = timed(run.step_resident, a.steps)
    t_load1 = time.time()
    launches = run.launches() - l0
    stage_launches = run.stage_launches_per_step() * a.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop(t_load0, t_load1) if rank == 0 else None
    value = 9.0 * n_tri_global * a.steps / (ms * 1e-3) / 1e6

    # ---------------- end-to-end through the reference-facing API
    e2e = None
    if not a.no_e2e:
        run.use_fused_norms(True)
        if not a.no_graph and hasattr(run, "enable_stage_graphs"):
            run.enable_stage_graphs()
        for _ in range(max(a.warmup, 3)):
            run.step_e2e()
        barrier()
        ms2 = timed(run.step_e2e, a.steps)
        if world > 1:
            t = torch.tensor([ms2], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms2 = float(t.item())
        e2e = {"value": 9.0 * n_tri_global * a.steps / (ms2 * 1e-3) / 1e6, "unit": "M dof-updates/s",
               "h2d_bytes_per_step": int(run.h2d_bytes_per_step()), "d2h_bytes_per_step": int(run.d2h_bytes_per_step()),
               "ms_per_step": ms2 / a.steps,
               "path": run.e2e_path()}

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
    alg_bytes = ALG_BYTES_BASE + 4 + 4          # + Manning + Coriolis P1 coefficient fields (4 B/triangle-stage each)
    n_tri_local = run.n_owned()
    kernel_ms = ms / (3 * a.steps)               # the timed region holds only stage kernels (+ halo traffic when N > 1)
    achieved = alg_bytes * n_tri_local / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "swe_stage_kernel<true>", "algorithmic_bytes_per_triangle_stage": alg_bytes,
                "triangles_per_launch": int(n_tri_local), "avg_launch_ms": kernel_ms, "peak_source": peak_src}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            roofline["traffic"] = json.load(open(tr)).get("dram_bytes_per_launch")
        except Exception:
            pass

    cpu = None
    if not a.no_cpu_baseline:
        val, cores, n, el = cpu_port(mesh, setup, wd, a.cpu_seconds, tide_values(setup, 0.0))
        cpu = {"value": val, "unit": "M dof-updates/s", "cores": cores, "kind": "port",
               "sample": f"{n} SSPRK33 steps of the full {mesh.n_cells}-triangle workload ({el:.1f} s, oracle/swe_oracle.c, OpenMP)"}

    line = {"metric": METRIC, "value": value, "unit": "M dof-updates/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "triangles": int(n_tri_global), "dofs": int(9 * n_tri_global),
                       "dt": dt,
                       "l2": ("per-GPU working set %.0f MB vs L2 %.0f MB: " % (work_bytes / 1e6, l2_bytes / 1e6))
                             + ("L2 flushed between timed steps (2 x L2 buffer written, untimed), steps timed one by one"
                                if flush else "inputs larger than L2, no flush"),
                       "parallelism": (f"domain decomposition x{world}, halo transport {run.transport}, overlap {run.overlap}, "
                                       f"cuda graph {hasattr(run, '_graph')}" if world > 1 else "single GPU")},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "stage_kernel_launches": int(stage_launches),
            "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
