#!/usr/bin/env python
"""
bench.py — staggered load steps / second on the BASELINE config, one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--n 1000]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[1] — single-edge notched shear 2D, AT2, Miehe
spectral split, n×n squares split into 2 P1 triangles (n = 1000 -> 2.0 M triangles, 1.0 M nodes),
loading u_x(top) = 1e-4·t, stagger 2/500/1e-8 (reference examples/PhaseFieldFracture/plot_1712.py).
A "step" is one load step = the whole stagger loop (u-solve, projection/history, phi-solve).
Load steps t = 0..W-1 are warm-up (t = 0 is the reference's trivial all-zero step), the next K
steps are timed.  N > 1: the same mesh is partitioned over the ranks (strong scaling).

One JSON line on stdout (rank 0).  `value` = K / device time of the K `pfx_load_step` calls
(fields resident in HBM); `e2e` = the same steps through the public `solve()` drop-in with host
buffers (Dirichlet values H2D, reactions/energies D2H, result files) timed on the host.
`roofline`: the dominant kernel (J_u·v apply) timed in isolation with CUDA events on the
library's stream, L2 flushed between launches, algorithmic bytes of SURVEY §8d.
`cpu_baseline`: the CPU oracle (scipy/SuperLU restatement of the reference algorithm — the
reference's dolfinx+PETSc path cannot run in this image) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "staggered_load_steps_per_second"
UNIT = "steps/s"
SAMPLE_N = int(os.environ.get("PFX_BENCH_SAMPLE_N", "160"))   # CPU sample: same workload on a SAMPLE_N x SAMPLE_N grid (env: tests shrink it)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index, self.proc, self.path = index, None, None

    def __enter__(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
                 "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        try:
            rows = [r.split(",") for r in open(self.path).read().strip().splitlines() if r.strip()]
            sm = [float(r[1]) for r in rows]
            out["sm_mhz"] = float(np.median(sm)) if sm else None
            out["sm_max_mhz"] = float(rows[0][2]) if rows else None
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            reasons = set()
            for r in rows:
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            out["reasons"] = sorted(reasons)
            os.unlink(self.path)
        except Exception:
            pass
        return out


# ------------------------------------------------------------------------------- CPU oracle leg
def run_oracle_sample(steps, warmup, n=SAMPLE_N, n_full=1000):
    """Times the CPU oracle (the only place bench.py executes oracle/) on the same workload at a
    reduced grid; returns (steps/s on the sample, steps/s scaled to the full grid ∝ 1/cells,
    description).  Scaling ∝ cells is generous to the CPU (sparse LU grows super-linearly)."""
    from oracle.solver import BC, Oracle, Params
    from phasefieldx_b200 import workloads as W
    w = W.sen_shear(n)
    P = Params(**w.params)
    o = Oracle(w.msh.geometry.x, w.msh.cells, w.cell_type, P)
    bcs = [BC(d) for _, d in w.bc_sets]
    times = []
    for s in range(warmup + steps):
        for bc, v in zip(bcs, w.bc_values(float(s))):
            bc.values = v
        t0 = time.perf_counter()
        o.load_step(bcs, [], dt=w.dt, min_stagger_iter=w.min_stagger_iter, max_stagger_iter=w.max_stagger_iter,
                    tol=w.stagger_tol)
        times.append(time.perf_counter() - t0)
    dt = sum(times[warmup:])
    v_sample = steps / dt
    scale = (2.0 * n * n) / (2.0 * n_full * n_full)
    desc = (f"oracle (numpy/scipy SuperLU restatement, 1 process) on the same workload at {n}x{n}x2 triangles "
            f"({2 * n * n} cells = 1/{int(round(1 / scale))} of the bench mesh), load steps t={warmup}..{warmup + steps - 1}: "
            f"{v_sample:.4f} steps/s on the sample; value = sample rate x cells ratio")
    return v_sample, v_sample * scale, desc, dt / steps * 1e3


def main_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    v_sample, v_full, desc, ms = run_oracle_sample(args.steps, min(args.warmup, 2), n_full=args.n)
    line = {"metric": METRIC, "value": v_full, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": min(args.warmup, 2), "ms_per_step": 1e3 / v_full, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"sen_shear_2d_spectral_n{args.n}", "cpu_sample_grid": SAMPLE_N,
                       "note": "dolfinx+PETSc cannot be installed in this image; oracle port timed instead"},
            "cpu_baseline": {"value": v_full, "unit": UNIT, "cores": 1, "host_cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": v_full, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------- GPU leg
def build_rank_engine(w, rank, world, device, opts):
    """Engine of this rank: whole mesh (world == 1) or its partition + NCCL / NVLink peer attach."""
    from phasefieldx_b200 import _lib, workloads as W
    if world == 1:
        return W.make_engine(w, device=device, **opts), None
    import torch.distributed as dist
    from phasefieldx_b200.distributed import RankMesh, attach
    from phasefieldx_b200.Materials.conversion import get_lambda_lame, get_mu_lame
    rm = RankMesh(w.msh.geometry.x, w.msh.cells, w.cell_type, world, rank)
    p = w.params
    e = _lib.Engine(rm.coords, rm.cells, w.cell_type, get_lambda_lame(p["E"], p["nu"]), get_mu_lame(p["E"], p["nu"]),
                    p["Gc"], p["l"], k=p["k"], model=_lib.model_code(p["degradation"], None if p["split_energy"] == "no" else p["split_energy"]),
                    irreversibility=(p["irreversibility"] == "miehe"), fatigue=p["fatigue"],
                    fatigue_val=p.get("fatigue_val", 0.05625), device=device, n_owned=rm.n_owned,
                    cell_primary=rm.cell_primary, **opts)
    local_sets = []
    for i, (_, dofs) in enumerate(w.bc_sets):
        keep, ldofs = rm.local_dofs(dofs, w.msh.geometry.dim)
        local_sets.append(keep)
        e.set_bc_u(i, ldofs)
    attach(e, rm, dist)
    return e, local_sets


def set_bc_values(e, w, t, local_sets):
    nbytes = 0
    for i, v in enumerate(w.bc_values(float(t))):
        if local_sets is not None:
            v = v[local_sets[i]]
        e.set_bc_values_u(i, v)
        nbytes += v.nbytes
    return nbytes


def e2e_solve(w, steps, warmup, device):
    """Same load steps through the public drop-in `solve()` (host buffers in, result files out)."""
    from phasefieldx_b200 import fem
    from phasefieldx_b200.Element.Phase_Field_Fracture.Input import Input
    from phasefieldx_b200.Element.Phase_Field_Fracture.solver.solver import solve
    import torch.distributed as dist
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    box = [tempfile.mkdtemp(prefix="pfx_bench_") if (not multi or dist.get_rank() == 0) else None]
    if multi:
        dist.broadcast_object_list(box, src=0)
    tmp = box[0]
    Data = Input(save_solution_vtu=False, results_folder_name=os.path.join(tmp, "results"), **w.params)
    V_u = fem.functionspace(w.msh, ("Lagrange", 1, (w.msh.geometry.dim,)))
    V_phi = fem.functionspace(w.msh, ("Lagrange", 1))
    bcs = []
    for _, dofs in w.bc_sets:
        bc = fem.DirichletBC(0.0, np.zeros(0, np.int64), V_phi)
        bc._dofs, bc._comp = dofs.astype(np.int32), None
        bc._vals = np.zeros(len(dofs))
        bc.values_per_dof = (lambda b=bc: b._vals)
        bcs.append(bc)
    stamps = []

    def update(bc_list, t):
        stamps.append(time.perf_counter())
        for bc, v in zip(bc_list, w.bc_values(float(t))):
            bc._vals = v
        return 1e-4 * t, 0, 0
    import contextlib
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        # the reference's file helpers print: stdout carries only the JSON line.
        # One extra load step: the timed region ends when it starts, i.e. after `steps` complete steps
        # including their result files, and excludes the engine tear-down at the end of solve()
        with contextlib.redirect_stdout(sys.stderr):
            solve(Data, w.msh, float(warmup + steps + 1), V_u, V_phi, bcs, [], update, None, None, None, None, w.dt, path=tmp,
                  bcs_list_u_names=[n for n, _ in w.bc_sets], min_stagger_iter=w.min_stagger_iter,
                  max_stagger_iter=w.max_stagger_iter, stagger_error_tol=w.stagger_tol, device=device)
    finally:
        os.chdir(cwd)
    t_end = time.perf_counter()
    import shutil
    if multi:
        dist.barrier()
    if not multi or dist.get_rank() == 0:
        shutil.rmtree(tmp, ignore_errors=True)
    h2d = sum(len(d) * 8 for _, d in w.bc_sets)
    d2h = len(w.bc_sets) * 3 * 8 + 11 * 8 + 256 * 6 * 8
    step_ms = [1e3 * (b - a) for a, b in zip(stamps[warmup:warmup + steps], stamps[warmup + 1:warmup + steps + 1])]
    sys.stderr.write(f"e2e step ms {step_ms}; tear-down + extra step {1e3 * (t_end - stamps[warmup + steps]):.1f} ms\n")
    return steps / (stamps[warmup + steps] - stamps[warmup]), h2d, d2h


def main_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from phasefieldx_b200 import workloads as W
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: phasefieldx_b200 has no CPU fallback")
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")        # keep NCCL's version banner off stdout (one JSON line)
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    w = W.sen_shear(args.n)
    opts = {}
    if args.cg_rtol:
        opts["cg_rtol"] = args.cg_rtol
    e, local_sets = build_rank_engine(w, rank, world, local_rank, opts)
    K, Wm = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    reports = []
    for s in range(Wm):
        set_bc_values(e, w, s, local_sets)
        e.load_step(dt=w.dt, min_stagger_iter=w.min_stagger_iter, max_stagger_iter=w.max_stagger_iter,
                    stagger_tol=w.stagger_tol)
    l0 = e.launch_count()
    barrier()
    with ClockSampler(local_rank) as clk:
        t0 = time.perf_counter()
        gpu_ms = 0.0
        for s in range(Wm, Wm + K):
            set_bc_values(e, w, s, local_sets)
            r = e.load_step(dt=w.dt, min_stagger_iter=w.min_stagger_iter, max_stagger_iter=w.max_stagger_iter,
                            stagger_tol=w.stagger_tol)
            gpu_ms += r["gpu_ms"]
            reports.append(r)
        barrier()
        wall = time.perf_counter() - t0
    launches = e.launch_count() - l0
    tt = torch.tensor([gpu_ms, wall * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    gpu_ms, wall_ms = tt.tolist()

    # ---- roofline of the dominant kernel (J_u·v), state of the last step
    peak, peak_kind = measured_peak()
    roof = None
    if rank == 0:
        u, phi, v = W.operator_state(w) if world == 1 else (None, None, None)
        if world == 1:
            e.apply_u(v, use_bc=True)
        ms = e.time_operator(0, reps=23, flush_l2=True)[3:]
        ms_nf = e.time_operator(0, reps=13, flush_l2=False)[3:]
        ms_it = e.time_operator(4, reps=13, flush_l2=True)[3:] if world == 1 else None
        bytes_apply = e.algorithmic_bytes(0)
        traffic = None
        try:    # DRAM bytes of one launch from the committed `ncu --set full` capture of this kernel/workload
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_apply_u_traffic.json")))
            if tj.get("workload") == w.name and world == 1:
                traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
        ach = bytes_apply / (ms.mean() * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "apply_u_tri_v3_kernel<spectral> (J_u·v on P1 triangles; persistent, TMA-staged)", "achieved": ach, "peak": peak,
                "peak_kind": peak_kind, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "algorithmic_bytes_per_launch": bytes_apply, "ms_per_launch": float(ms.mean()),
                "ms_per_launch_l2_warm": float(ms_nf.mean()),
                "note": "fp64 ALU work of the spectral split also bounds this kernel (see DESIGN.md)"}
        if ms_it is not None:
            bi = e.algorithmic_bytes(4)
            roof["pcg_iteration"] = {"achieved": bi / (ms_it.mean() * 1e-3) / 1e9, "ms": float(ms_it.mean()),
                                     "algorithmic_bytes": bi, "frac": bi / (ms_it.mean() * 1e-3) / 1e9 / peak}
    stats = e.block_stats()
    cst = e.coarse_stats()
    solver = ("matrix-free PCG, block-Jacobi + aggregate rigid-body-mode coarse space "
              f"({cst['aggregates']} aggregates per rank, {cst['setups']} set-ups so far)") if cst["active"] else "matrix-free PCG, Jacobi"
    if world > 1:
        dist.barrier()          # peers keep each other's buffers mapped (CUDA IPC): close together
    e.close()

    # end to end through the public drop-in solve(): on several GPUs it runs collectively (one process per
    # GPU, the same torch.distributed group), rank 0 owns the result files and the clock
    e2e_val, h2d, d2h = (None, 0, 0)
    if not args.no_e2e:
        e2e_val, h2d, d2h = e2e_solve(w, K, Wm, local_rank)
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu:
        v_s, v_f, desc, _ = run_oracle_sample(max(1, min(K, 2)), 1, n_full=args.n)
        cpu = {"value": v_f, "unit": UNIT, "cores": 1, "host_cores": os.cpu_count(), "kind": "port", "sample": desc}
    value = K / (gpu_ms * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": gpu_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": w.name, "nodes": int(w.msh.num_nodes), "cells": int(w.msh.num_cells),
                       "dofs_u_plus_phi": int(3 * w.msh.num_nodes), "load_steps_timed": [Wm, Wm + K - 1],
                       "stagger": [w.min_stagger_iter, w.max_stagger_iter, w.stagger_tol],
                       "l2": "working set of one PCG iteration (273 MB) exceeds L2; operator timing flushes L2",
                       "parallelism": f"mesh partitioned over {world} GPU(s)", "linear_solver": solver, "blocks": stats},
            "timing": {"device_ms": gpu_ms, "host_wall_ms": wall_ms,
                       "stagger_iters": [r["stagger"] for r in reports],
                       "cg_its_u": [int(r["cg_its_u"]) for r in reports],
                       "cg_its_phi": [int(r["cg_its_phi"]) for r in reports],
                       "cg_its_proj": [int(r["cg_its_proj"]) for r in reports]},
            "clocks": clk.summary(), "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
            "e2e": ({"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}
                    if e2e_val is not None else None)}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=1000, help="grid size of the shear workload (n x n x 2 triangles)")
    ap.add_argument("--cg-rtol", type=float, default=0.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        main_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        os.execv(sys.executable, cmd)
    main_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
