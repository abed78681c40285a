"""Benchmark of the thermoelastic hot path on B200 (contract: see the task statement / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--cells C] [--impl tefem|reference]

A "step" is ONE implicit transient time step (reference solvers.py:164-186) of BASELINE.json config 4: a
unit box of C^3 cells x 6 Kuhn tets (default C = 190: 41 154 000 tets, 6 967 871 nodes, 27 871 484 DOF), single
config-2 material, config-2 Dirichlet sets / time-varying surface load / Rayleigh damping, dt = 1800/999 s,
Krylov tolerance 1e-10 on the per-field scaled true residual.  Every step = load vector, Newmark predictor,
right-hand side (2 plane SpMVs), BiCGSTAB solve on the block-Jacobi-scaled system with the aggregation coarse
correction as right preconditioner (two 4x4-block SpMVs + two preconditioner passes per iteration), corrector.
Assembly of M, D, K (kernel K1) is timed separately on the same mesh and reported under "assembly".

One JSON line is printed by rank 0; see DESIGN.md for the definition of every key.  Besides the contract's keys:
  prepare_s        everything before the time loop on this mesh (assembly, Dirichlet elimination, block-Jacobi scaling,
                   preconditioner set-up, halo set-up), max over ranks;  rhs_planes_spmv: bandwidth of the two plane SpMVs
                   of the Newmark right-hand side;
  parity_vs_1gpu   (N > 1) the row-partitioned solve against the single-GPU solve of a small box, in this very run;
  same_config      (N = 1) GPU arm and the reference's CPU implementation on the SAME mesh and steps, both measured:
                   BASELINE config 2 (examples/example_transient_2.py on data/sandwich.msh);
  cpu_baseline     the reference's algorithm on a bounded sample, SCALED to the 28 M-DOF metric ("modelled": true).
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
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"      # stdout carries exactly one JSON line (NCCL prints its version banner there)

MAT2 = dict(rho=2300, Y=64e9, nu=0.1, k=1., c=1., alpha=4e-6, T0=20.)     # examples/example_transient_2.py:29
T_END, N_T = 1800e0, 1000                                                    # example_transient_2.py:20-21
DU = {4: [('x', 0.), ('y', 0.), ('z', 0.)], 5: [('y', 0.), ('z', 0.)]}       # example_transient_2.py:38-41
DT = {4: 0., 5: 0.}
ALPHA_M = ALPHA_K = 0.2
RTOL = 1e-10
ML_OPTIONS = None    # --ml "n_pre=3,n_post=2,cycle=classic,...": keyword arguments of multilevel.TwoLevelPreconditioner (experiments)
HALO = "peer"        # --halo nccl: the NCCL send/recv halo exchange (measured alternative to the peer-memory halo)


def surface_load():
    vec_t = np.linspace(0., T_END, N_T)
    arr = np.zeros((3, N_T))
    arr[0, :] = np.sin(2 * np.pi * vec_t / 900.) * -1e9                      # example_transient_2.py:53-59
    return arr


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def build_case(cells, rank=0, world=1):
    from thermoelasticity_fem_b200 import LinearThermoElastic, Mesh, Model, LinearTransient
    from thermoelasticity_fem_b200.synthetic import kuhn_box
    if world > 1:
        from thermoelasticity_fem_b200.distributed import build_distributed_case
        return build_distributed_case(cells, rank, world, MAT2, DU, DT, surface_load(), ALPHA_M, ALPHA_K,
                                      T_END, N_T, RTOL, halo=HALO, precond_options=ML_OPTIONS)
    pts, tets, tris, nodes = kuhn_box(cells, cells, cells, 1.0, 1.0, 1.0, n_layers=1, index_dtype=np.int32)
    mesh = Mesh().from_arrays(pts, tets, tris, nodes)
    mesh.set_materials({1: LinearThermoElastic(**MAT2)})
    mesh.make_elements()
    model = Model(mesh, dict_dirichlet_U=DU, dict_dirichlet_theta=DT, dict_surface_forces={5: surface_load()},
                  alpha_M=ALPHA_M, alpha_K=ALPHA_K)
    n = mesh.n_nodes
    solver = LinearTransient(model, np.zeros(3 * n), np.zeros(3 * n), np.zeros(n), np.zeros(n), T_END, N_T, 0.5, 0.25,
                             rtol=RTOL, history='last', progress=False, check_every=16, precond_options=ML_OPTIONS)
    return mesh, model, solver


def run_tefem(args):
    import torch
    import torch.distributed as dist
    from thermoelasticity_fem_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.load()
    hbm_peak, peak_src = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    t0 = time.perf_counter()
    mesh, model, solver = build_case(args.cells, rank, world)
    dm = mesh.device_mesh()
    S = dm.schedule
    t_setup = time.perf_counter() - t0
    E_loc, P_loc, N_own = dm.n_elems, S.n_blocks, S.n_rows
    n_dofs_total = 4 * (args.cells + 1) ** 3
    E_total = 6 * args.cells ** 3

    # ---- assembly (K1): M + D + K in one launch, and K alone; CUDA events, outputs (>= L2) rewritten every launch
    asm = {}
    for ops, planes in (("MKD", 29), ("K", 13)):
        out = dm.assemble(ops, ALPHA_M, ALPHA_K)
        for _ in range(3):
            dm.assemble(ops, ALPHA_M, ALPHA_K, out=out, check=False)
        reps = 5
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(reps):
            dm.assemble(ops, ALPHA_M, ALPHA_K, out=out, check=False)
        ev1.record()
        torch.cuda.synchronize()
        ms = max_over_ranks(ev0.elapsed_time(ev1) / reps)
        alg = 8 * planes * P_loc + 20 * E_loc + 24 * dm.n_nodes + 4 * 16 * E_loc
        asm[ops] = {"ms": ms, "elements_per_s": E_total / (ms * 1e-3),
                    "algorithmic_bytes": alg, "achieved_gbs": alg / ms / 1e6, "frac": alg / ms / 1e6 / hbm_peak}
        del out
    dm.check_jacobian()
    asm["schedule"] = {"tiles": int(S.n_cta), "max_items_per_tile": int(S.max_cta_items), "max_pieces_per_tile": int(S.max_cta_slots),
                       "pieces_first": bool(S.pieces_first), "bank_aware_numbering": bool(S.bank_optimized),
                       "lists_pitched": bool(S.lists_pitched), "smem_bytes_per_cta": int(S.smem_bytes())}

    # ---- everything before the time loop (assembly again, elimination, scaling, preconditioner, halo set-up)
    barrier()
    t0 = time.perf_counter()
    solver.prepare()
    barrier()
    prepare_s = max_over_ranks(time.perf_counter() - t0)
    kry = solver._ss.krylov
    # ---- the two plane SpMVs of the Newmark right-hand side (D and K in plane storage, once per step)
    planes = solver.planes if world > 1 else model._planes
    rhs_spmv = {}
    xin, yout = solver._w1, solver._rhs
    for nm, pmap in (("D", solver._mapD), ("K", solver._mapK)):
        dm.planes_spmv(planes[nm], pmap, xin, yout, alpha=-1.0, beta_y=1.0)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            dm.planes_spmv(planes[nm], pmap, xin, yout, alpha=-1.0, beta_y=1.0)
        e1.record()
        torch.cuda.synchronize()
        ms = max_over_ranks(e0.elapsed_time(e1) / 3)
        nbytes = 8 * int(planes[nm].shape[0]) * P_loc + 4 * P_loc + 4 * (N_own + 1) + 3 * 8 * 4 * N_own
        rhs_spmv[nm] = {"ms": ms, "planes": int(planes[nm].shape[0]), "algorithmic_bytes": nbytes,
                        "achieved_gbs": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / hbm_peak}
    for i in range(1, args.warmup + 1):
        solver.step(i)
    # device time of the Krylov solve inside every step (events around the driver call): step = solve + everything else
    # (load vector, predictor, right-hand side, scaling, halo of the new acceleration, corrector, host bookkeeping)
    solve_events = []
    orig_bicgstab = kry.bicgstab

    def timed_bicgstab(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig_bicgstab(*a, **k)
        e1.record()
        solve_events.append((e0, e1))
        return out
    kry.bicgstab = timed_bicgstab
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    kry.profile(True)
    launches0 = lib.tefem_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    infos = []
    barrier()
    ev0.record()
    for i in range(args.warmup + 1, args.warmup + args.steps + 1):
        infos.append(solver.step(i))
    ev1.record()
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = lib.tefem_launch_count() - launches0
    kry.bicgstab = orig_bicgstab
    solve_ms = max_over_ranks(sum(e0.elapsed_time(e1) for e0, e1 in solve_events) / max(len(solve_events), 1))
    spmv_ms, spmv_count = kry.profile(False)
    clocks = sampler.stop() if rank == 0 else None
    iters = [inf["iterations"] for inf in infos]
    pc = getattr(solver._ss, "precond", None) or getattr(solver, "precond", None)
    pc = pc if hasattr(pc, "n1") else None
    precond_desc = "block-Jacobi scaling only" if pc is None else (
        f"block-Jacobi scaling + additive aggregation correction with rigid-body coarse spaces: level 1 = {pc.n1} "
        f"aggregates (cubes of {pc.factor1:g} h, Galerkin, operator stored in {pc.level1_dtype}, {pc.n_pre}+{pc.n_post} damped-Jacobi "
        f"sweeps {pc.omega_c:g}, {pc.cycle} cycle), "
        f"level 2 = {pc.n2} aggregates (dense inverse), fine weight {pc.omega:g}; coarse cycle: "
        + ("one persistent kernel per application, rows of levels 1 and 2 partitioned over the ranks, coarse vectors exchanged "
           "through peer memory" if pc.coarse == "persistent" else "separate launches on replicated levels"))
    ms_per_step = ms_total / args.steps
    steps_per_s = 1e3 / ms_per_step

    # ---- SpMV roofline (dominant kernel): algorithmic bytes, node-block accounting
    n_halo = getattr(solver, "n_halo", 0)
    spmv_bytes = 8 * 16 * P_loc + 4 * P_loc + 4 * (N_own + 1) + 16 * 4 * N_own
    spmv_avg_ms = max_over_ranks(spmv_ms / max(spmv_count, 1))
    spmv_gbs = spmv_bytes / spmv_avg_ms / 1e6
    traffic = None      # DRAM bytes per launch from the committed ncu --set full capture of this very command
    tpath = os.path.join(ROOT, "profiles", "r2_spmv_traffic.json")      # ncu --set full of the CURRENT kernel (profiles/README.md)
    if os.path.exists(tpath):
        with open(tpath) as fh:
            tj = json.load(fh)
        if tj["workload"] == {"cells": args.cells, "n_gpus": world}:
            traffic = tj["traffic_bytes_per_launch_mean"]
    roofline = {"bound": "hbm", "kernel": "spmv_kernel (4x4-block SpMV of the Krylov iteration, fused dot partials)",
                "achieved": spmv_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": spmv_gbs / hbm_peak,
                "peak_source": peak_src, "traffic": traffic, "launches_timed": int(spmv_count),
                "avg_launch_ms": spmv_avg_ms, "algorithmic_bytes_per_launch": spmv_bytes,
                "share_of_step": spmv_ms / ms_total}

    # ---- end to end through the public API: host-resident inputs and results every step
    e2e = None
    if args.e2e_steps > 0:
        n_loc = 4 * dm.n_nodes if world == 1 else solver._x.numel()
        host = [torch.empty(n_loc, dtype=torch.float64).pin_memory() for _ in range(3)]
        # H2D: the step's load amplitudes (4 doubles per load term) live in pinned host memory and are copied to the device
        # every step; the load kernel reads them from there (tefem_load_axpy_dev) - no amplitude travels as a kernel argument
        n_terms = solver.n_load_terms()
        coef_host = torch.zeros((max(n_terms, 1), 4), dtype=torch.float64).pin_memory()
        coef_dev = torch.zeros((max(n_terms, 1), 4), dtype=torch.float64, device="cuda")
        first = args.warmup + args.steps + 1
        # the state of step i is snapshotted on the device (3 vector copies) and read back to pinned host memory on
        # a second stream while step i + 1 runs; every read-back completes inside the timed region
        stage = [torch.empty_like(d) for d in (solver._x, solver._v, solver._a)]
        copy_stream = torch.cuda.Stream()
        ev_stage, ev_copied = torch.cuda.Event(), torch.cuda.Event()
        main = torch.cuda.current_stream()
        barrier()
        t0 = time.perf_counter()
        for k, i in enumerate(range(first, first + args.e2e_steps)):
            if n_terms:                                                    # (a rank may own no loaded node)
                coef_host[:n_terms].copy_(torch.as_tensor(solver.load_coefficients(i), dtype=torch.float64).reshape(-1, 4))
            coef_dev.copy_(coef_host, non_blocking=True)                   # H2D: this step's load amplitudes
            solver.step(i, coef_dev=coef_dev)
            if k:
                main.wait_event(ev_copied)                                 # previous read-back has left the staging copy
            for sdev, d in zip(stage, (solver._x, solver._v, solver._a)):
                sdev.copy_(d, non_blocking=True)
            ev_stage.record(main)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev_stage)
                for h, sdev in zip(host, stage):                           # D2H: X, Xdot, Xdotdot of this step
                    h.copy_(sdev, non_blocking=True)
                ev_copied.record(copy_stream)
        torch.cuda.synchronize()
        barrier()
        dt_e2e = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": args.e2e_steps / dt_e2e, "unit": "steps/s", "h2d_bytes_per_step": int(sum_over_ranks(32 * n_terms)),
               "d2h_bytes_per_step": int(sum_over_ranks(3 * 8 * n_loc)), "steps": args.e2e_steps,
               "note": "H2D = the step's load amplitudes from pinned memory (read by the load kernel on the device); "
                       "read-back of step i overlapped with step i + 1 (device snapshot + copy stream)"}

    # ---- N > 1: the partitioned solve against the single-GPU solve of a small box, in this run
    parity = None
    if world > 1 and args.parity_cells > 0:
        parity = parity_vs_single_gpu(args.parity_cells, rank, world)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = cpu_baseline(n_dofs_total, budget_s=args.cpu_budget) if (world == 1 and args.cpu_budget > 0) else None
    same = same_config_sandwich(args.same_config_steps) if (world == 1 and args.same_config_steps > 0) else None
    n_free = int(n_dofs_total - (len(model.dirichlet_dofs_U) + len(model.dirichlet_dofs_theta))) if world == 1 else None
    line = {
        "metric": "transient steps/s (implicit Newmark step of the coupled thermoelastic model, 28M DOF box)",
        "value": steps_per_s, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"BASELINE config 4: Kuhn box {args.cells}^3 cells = {E_total} tets, {n_dofs_total} DOF, "
                               "transient (config-2 material/BCs/load/Rayleigh), BiCGSTAB rtol 1e-10 per field, "
                               "block-Jacobi scaling + 3-level aggregation preconditioner",
                   "cells": args.cells, "n_tets": E_total, "n_dofs": n_dofs_total, "n_free_dofs": n_free,
                   "partition": "single GPU" if world == 1 else f"x-slabs of node rows over {world} GPUs",
                   "halo": None if world == 1 else {
                       "peer": "fused: ONE SpMV kernel stores its boundary values into the neighbours' vectors over NVLink, multiplies "
                               "the interior rows, waits for the neighbours' flags, multiplies the boundary rows",
                       "peer-push": "peer memory: separate push launch, the SpMV waits for the flags before its first row",
                       "nccl": "NCCL send/recv, overlapped with the interior rows"}[HALO],
                   "l2_policy": "operands larger than L2 (system matrix %.1f GB per GPU)" % (8 * 16 * P_loc / 1e9)},
        "solver": {"preconditioner": precond_desc, "iterations_per_step": float(np.mean(iters)), "iterations": iters,
                   "dof_iterations_per_s": n_dofs_total * float(np.sum(iters)) / (ms_total * 1e-3),
                   "ms_per_iteration": ms_total / max(float(np.sum(iters)), 1.0),
                   "solve_ms_per_step": solve_ms, "other_ms_per_step": ms_total / args.steps - solve_ms,
                   "solve_ms_per_iteration": solve_ms * args.steps / max(float(np.sum(iters)), 1.0),
                   "relres": [inf["relres"] for inf in infos], "restarts": [inf["restarts"] for inf in infos]},
        "assembly": asm, "roofline": roofline, "cpu_baseline": cpu, "same_config": same, "e2e": e2e,
        "gpu_launches": int(launches), "clocks": clocks, "setup_s": t_setup, "prepare_s": prepare_s,
        "rhs_planes_spmv": rhs_spmv, "parity_vs_1gpu": parity,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def parity_vs_single_gpu(cells, rank, world, steps=2):
    """The row-partitioned transient solve of a cells^3 box on `world` GPUs against the single-GPU solve of the same box
    (computed by rank 0 while the others wait): largest relative difference per field of x, v, a after `steps` steps."""
    import torch
    import torch.distributed as dist
    mesh, model, solver = build_case(cells, rank, world)
    solver.prepare()
    infos = [solver.step(i) for i in range(1, steps + 1)]
    lp = solver.lp
    sizes = [int(4 * (lp.bounds[r + 1] - lp.bounds[r])) for r in range(world)]
    full = []
    for t in solver.owned_state():
        pad = torch.zeros(max(sizes), dtype=torch.float64, device="cuda")
        pad[:t.numel()] = t
        out = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(out, pad)
        full.append(torch.cat([o[:s_] for o, s_ in zip(out, sizes)]).cpu().numpy())
    res = None
    if rank == 0:
        _, _, ref = build_case(cells, 0, 1)
        ref.prepare()
        ref_infos = [ref.step(i) for i in range(1, steps + 1)]
        want = [(ref._x + ref.model._g_fill).cpu().numpy(), ref._v.cpu().numpy(), ref._a.cpu().numpy()]
        worst = {}
        for name, got, w in zip(("x", "v", "a"), full, want):
            idx = np.arange(w.size)
            for fld, m in (("u", idx % 4 != 3), ("theta", idx % 4 == 3)):
                worst[f"{name}_{fld}"] = float(np.linalg.norm(got[m] - w[m]) / max(np.linalg.norm(w[m]), 1e-300))
        res = {"cells": cells, "n_dofs": int(want[0].size), "steps": steps, "rel_diff": worst, "max_rel_diff": max(worst.values()),
               "ok": bool(max(worst.values()) < 1e-7), "tolerance": 1e-7,
               "iterations": {"partitioned": [i["iterations"] for i in infos], "single_gpu": [i["iterations"] for i in ref_infos]}}
        del ref
    del solver
    torch.cuda.synchronize()
    dist.barrier()
    return res


# ---------------------------------------------------------------------------------------------------
# same mesh, same steps, both measured: BASELINE config 2 (examples/example_transient_2.py on data/sandwich.msh)
# ---------------------------------------------------------------------------------------------------
def _sandwich_config2(ns, n_steps, mesh_path=None, blocks=None):
    """Model + LinearTransient of example_transient_2.py truncated to n_steps steps (same dt = 1800 / 999) for the package
    `ns` (the reference's modules, or this repo's)."""
    dt = T_END / (N_T - 1)
    n_t = n_steps + 1
    vec_t = np.linspace(0., dt * n_steps, n_t)
    mesh = ns.Mesh()
    if blocks is not None:
        mesh.from_blocks(*blocks)
    else:
        mesh.load_mesh(mesh_path)
    mat = ns.LinearThermoElastic(**MAT2)
    mesh.set_materials({1: mat, 2: mat, 3: mat})
    mesh.make_elements()
    arr = np.zeros((3, n_t))
    arr[0, :] = np.sin(2 * np.pi * vec_t / 900.) * -1e9
    model = ns.Model(mesh, dict_dirichlet_U=DU, dict_dirichlet_theta=DT, dict_surface_forces={5: arr},
                     alpha_M=ALPHA_M, alpha_K=ALPHA_K)
    n = mesh.n_nodes
    return model, (model, np.zeros(3 * n), np.zeros(3 * n), np.zeros(n), np.zeros(n), dt * n_steps, n_t, 0.5, 0.25)


def same_config_sandwich(ref_steps=5, gpu_steps=200):
    """BASELINE config 2 on BOTH arms, measured in this run: the unmodified reference (its own classes, byte code under
    oracle/_ref or the sources when present; SuperLU re-factorised every step, one core) and this repo through the same
    public API.  Both are timed over the bodies of their time loops (the reference's through the iterator it wraps its loop
    in, solvers.py:163); set-up (assembly, DOF maps) is reported beside each."""
    import types
    import torch
    import thermoelasticity_fem_b200 as te
    from oracle import reference_shim as rs
    out = {"workload": "BASELINE config 2: examples/example_transient_2.py on data/sandwich.msh (3 119 nodes, 13 712 tets, "
                       "12 476 DOF), dt = 1800/999 s, Rayleigh 0.2 / 0.2", "metric": "transient steps/s", "same_config": True}
    # ---- GPU arm
    g = np.load(rs.SANDWICH_NPZ)
    blocks = [(str(g[f"b{i}_type"]), int(g[f"b{i}_tag"]), g[f"b{i}_conn"].astype(np.int64)) for i in range(int(g["n_blocks"]))]
    ns_gpu = types.SimpleNamespace(Mesh=te.Mesh, LinearThermoElastic=te.LinearThermoElastic, Model=te.Model)
    model, targs = _sandwich_config2(ns_gpu, gpu_steps, blocks=(g["points"], blocks))
    solver = te.LinearTransient(*targs, progress=False, history='last')
    t0 = time.perf_counter()
    solver.prepare()
    torch.cuda.synchronize()
    t_prep = time.perf_counter() - t0
    for i in range(1, 4):
        solver.step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    its = [solver.step(i)["iterations"] for i in range(4, gpu_steps + 1)]
    torch.cuda.synchronize()
    dt_gpu = time.perf_counter() - t0
    out["gpu"] = {"steps_per_s": (gpu_steps - 3) / dt_gpu, "steps_timed": gpu_steps - 3, "prepare_s": t_prep,
                  "iterations_per_step": float(np.mean(its))}
    del solver, model
    out["reference"] = same_config_reference(ref_steps)
    if "steps_per_s" in out["reference"]:
        out["ratio_measured"] = out["gpu"]["steps_per_s"] / out["reference"]["steps_per_s"]
    return out


def same_config_reference(ref_steps=5):
    """The unmodified reference on BASELINE config 2, timed on the host cores (see same_config_sandwich)."""
    import types
    from oracle import reference_shim as rs
    if not rs.available():
        return {"unavailable": "neither /root/reference nor oracle/_ref byte code present"}
    ref = rs.load_reference()
    ns_ref = types.SimpleNamespace(Mesh=ref.mesh.Mesh, LinearThermoElastic=ref.materials.LinearThermoElastic, Model=ref.model.Model)
    # the reference wraps its time loop in tqdm (solvers.py:163): substituting a timing iterator for the progress bar gives
    # the start time of every loop body without touching the reference's code
    stamps = []

    def timed(iterable, *a, **k):
        for item in iterable:
            stamps.append(time.perf_counter())
            yield item
        stamps.append(time.perf_counter())

    _, rargs = _sandwich_config2(ns_ref, ref_steps, mesh_path=rs.sandwich_path())
    rsolver = ref.solvers.LinearTransient(*rargs)
    saved = ref.solvers.tqdm
    ref.solvers.tqdm = timed
    try:
        t0 = time.perf_counter()
        rsolver.solve()
        t_total = time.perf_counter() - t0
    finally:
        ref.solvers.tqdm = saved
    per_step = (stamps[-1] - stamps[0]) / ref_steps
    return {"steps_per_s": 1.0 / per_step, "s_per_step": per_step, "steps_timed": ref_steps, "cores": 1,
            "kind": "reference", "source": "sources" if rs.sources_available() else "byte code (oracle/_ref)",
            "what_is_timed": "the bodies of the reference's own time loop (solvers.py:163-186: assemble_F, apply_dirichlet_F, "
                             "2 SpMV, spsolve re-factorised, updates), via the iterator it wraps the loop in",
            "solve_call_s_incl_assembly_and_postprocessing": t_total, "host_cpu_count": os.cpu_count()}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's algorithm on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_case(cells):
    from oracle import tefem_oracle as orc
    from thermoelasticity_fem_b200.synthetic import kuhn_box
    pts, tets, tris, nodes = kuhn_box(cells, cells, cells, 1.0, 1.0, 1.0, n_layers=1)
    groups = dict(nodes=nodes, tri=tris, tet=tets)
    mats = {1: orc.Material(**MAT2)}
    return orc, pts, groups, mats


def cpu_assembly_rate(n_elems=4000):
    """elements/s of the per-element port (the reference's Python loop, model.py:87-118): M + K + D."""
    orc, pts, groups, mats = cpu_case(10)
    conn = groups["tet"][1][:n_elems]
    t0 = time.perf_counter()
    for e in range(conn.shape[0]):
        orc.element_matrices_percall(pts[conn[e]], mats[1], e, which=("Muu",))                      # assemble_M
        orc.element_matrices_percall(pts[conn[e]], mats[1], e, which=("Kuu", "Kut", "Ktt"))        # assemble_K
        orc.element_matrices_percall(pts[conn[e]], mats[1], e, which=("Dtu", "Dtt", "Muu", "Kuu"))  # assemble_D (Rayleigh)
    return conn.shape[0] / (time.perf_counter() - t0)


class CpuTransient:
    """Reference algorithm per step (solvers.py:164-186): F(i), lifting, 2 SpMV, spsolve(K_dyn, rhs), updates."""

    def __init__(self, cells):
        import scipy.sparse.linalg as spla
        orc, pts, groups, mats = cpu_case(cells)
        self.orc, self.pts, self.groups, self.spla = orc, pts, groups, spla
        n_nodes = pts.shape[0]
        conn, mat_id, rows = orc.element_table(groups["tet"], mats)
        ops = orc.assemble(pts, conn, rows, mat_id, ALPHA_M, ALPHA_K)
        maps = orc.free_dofs_lists(n_nodes, groups["tri"], DU, DT)
        self.free = maps["free_dofs"]
        self.K = ops["K"]
        self.Mff, self.Dff, self.Kff = (orc.reduce_matrix(ops[k], self.free) for k in "MDK")
        self.dt = T_END / (N_T - 1)
        self.Kdyn = (self.Mff + 0.5 * self.dt * self.Dff + 0.25 * self.dt ** 2 * self.Kff).tocsc()
        self.x = np.zeros(self.free.shape[0])
        self.v = np.zeros_like(self.x)
        self.a = np.zeros_like(self.x)
        self.loads = dict(surface={5: surface_load()})
        self.n_dofs = 4 * n_nodes
        self.i = 0

    def step(self):
        self.i += 1
        orc, dt = self.orc, self.dt
        F = orc.assemble_F(self.pts, self.groups, self.loads, idx_t=self.i)
        Ff = orc.lift_rhs(F, self.K, self.free, self.n_dofs, self.groups["tri"], DU, DT)
        rhs = Ff - self.Dff @ (self.v + 0.5 * dt * self.a) - self.Kff @ (self.x + dt * self.v + 0.25 * dt ** 2 * self.a)
        a_new = self.spla.spsolve(self.Kdyn, rhs)                         # re-factorised every step, like the reference
        v_new = self.v + 0.5 * dt * self.a + 0.5 * dt * a_new
        x_new = self.x + dt * self.v + 0.5 * dt ** 2 * (0.5 * self.a + 0.5 * a_new)
        self.x, self.v, self.a = x_new, v_new, a_new


SCALING_RULE = ("sample steps/s x (sample DOF / workload DOF): LINEAR in the DOF count, optimistic for the reference - "
                "measured in the survey container the step costs 1.4 / 8.3 / 31.7 s at 8 788 / 19 652 / 37 044 DOF (~ n^2.2), "
                "the sparse LU of a 3-D mesh is super-linear and cannot be held in host memory at 28 M DOF")


def cpu_baseline(n_dofs_full, budget_s=20.0, cells=16):
    t0 = time.perf_counter()
    case = CpuTransient(cells)
    t_setup = time.perf_counter() - t0
    n, t1 = 0, time.perf_counter()
    while n < 2 or (time.perf_counter() - t1 < budget_s * 0.6 and n < 50):
        case.step()
        n += 1
    dt = time.perf_counter() - t1
    raw = n / dt
    return {"value": raw * case.n_dofs / n_dofs_full, "unit": "steps/s", "cores": 1, "kind": "port", "modelled": True,
            "sample": f"oracle port of the reference step (F, lifting, 2 SpMV, scipy spsolve re-factorised every step) on a "
                      f"{cells}^3-cell Kuhn box = {case.n_dofs} DOF ({n} steps timed, {raw:.4f} steps/s there); "
                      f"assembly port = per-element Python loop",
            "sample_value": raw, "sample_n_dofs": case.n_dofs, "scaled_to": SCALING_RULE,
            "assembly_elements_per_s": cpu_assembly_rate(), "setup_s": t_setup, "host_cpu_count": os.cpu_count()}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm on the host cores, same metric / unit.  The reference cannot run
    the 28 M-DOF workload at all (dense n_dofs^2 assembly, model.py:88), so `value` is MODELLED ("modelled": true): the
    oracle port of its step on a bounded sample box, scaled linearly in the DOF count.  The MEASURED like-for-like figure is
    in "same_config": the unmodified reference (oracle/_ref byte code, or its sources) on BASELINE config 2, the mesh it
    ships - bench.py's own arm reports the GPU side of the same block."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cells = args.ref_cells
    case = CpuTransient(cells)
    for _ in range(args.warmup):
        case.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        case.step()
    dt = time.perf_counter() - t0
    raw = args.steps / dt
    E_total = 6 * args.cells ** 3
    n_dofs_total = 4 * (args.cells + 1) ** 3
    v = raw * case.n_dofs / n_dofs_total          # scaled to the metric's unit: steps/s of the 28 M-DOF workload
    sample = (f"bounded sample: the same transient step (config-2 physics) on a {cells}^3-cell Kuhn box = {case.n_dofs} DOF "
              f"({raw:.4f} steps/s there); scipy spsolve re-factorised every step as in the reference (solvers.py:172); "
              "single-threaded (Python + sequential SuperLU)")
    line = {"impl": "reference", "metric": "transient steps/s (implicit Newmark step of the coupled thermoelastic model, 28M DOF box)",
            "value": v, "unit": "steps/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"BASELINE config 4: Kuhn box {args.cells}^3 cells = {E_total} tets, {n_dofs_total} DOF, "
                                   "transient (config-2 material/BCs/load/Rayleigh), BiCGSTAB rtol 1e-10 per field, "
                               "block-Jacobi scaling + 3-level aggregation preconditioner",
                       "cells": args.cells, "n_tets": E_total, "n_dofs": n_dofs_total},
            "modelled": True,
            "cpu_baseline": {"value": v, "unit": "steps/s", "cores": 1, "kind": "port", "modelled": True, "sample": sample,
                             "sample_value": raw, "sample_n_dofs": case.n_dofs, "scaled_to": SCALING_RULE,
                             "host_cpu_count": os.cpu_count()},
            "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if args.same_config_steps > 0:
        line["same_config"] = {"workload": "BASELINE config 2: examples/example_transient_2.py on data/sandwich.msh (12 476 DOF)",
                               "metric": "transient steps/s", "same_config": True,
                               "reference": same_config_reference(args.same_config_steps)}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cells", type=int, default=190, help="cells per edge of the Kuhn box (190 = BASELINE config 4)")
    ap.add_argument("--impl", default="tefem", choices=["tefem", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for cpu_baseline (0 = skip)")
    ap.add_argument("--ref-cells", type=int, default=16)
    ap.add_argument("--parity-cells", type=int, default=24, help="N > 1: box for the partitioned-vs-single-GPU check (0 = skip)")
    ap.add_argument("--same-config-steps", type=int, default=5,
                    help="N = 1: reference steps of BASELINE config 2 timed beside the GPU arm (0 = skip; ~25 s of CPU)")
    ap.add_argument("--halo", default="peer", choices=["peer", "peer-push", "nccl"],
                    help="N > 1: halo of the SpMV - direct stores into the neighbours' vectors (default) or NCCL send/recv")
    ap.add_argument("--ml", default="", help="experiments: keyword arguments of the preconditioner, e.g. n_pre=3,n_post=2,cycle=classic")
    args = ap.parse_args()
    global HALO, ML_OPTIONS
    HALO = args.halo
    if args.ml:
        ML_OPTIONS = {}
        for kv in args.ml.split(","):
            k, v = kv.split("=")
            ML_OPTIONS[k] = v if k in ("cycle", "level1_dtype", "coarse") else (int(v) if k in ("n_pre", "n_post", "dense_limit") else float(v))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_tefem(args)


if __name__ == "__main__":
    main()
