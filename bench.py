#!/usr/bin/env python
"""
bench.py — staggered load steps / second on the BASELINE configurations, one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                  [--workload auto|beam3d|sen_shear|sent|fatigue] [--beam-nx 480] [--n 1000]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P bench.py --gpus N --steps K --warmup W

Workloads (phasefieldx_b200/workloads.py; SURVEY §8d "Synthetic inputs per config"):
  beam3d     BASELINE.json configs[4]: 3D notched beam (reference examples/PhaseFieldFracture/plot_1718.py), P1
             tetrahedra, Miehe spectral split; nx x nx/3 x nx/6 cubes split into 6 tets, nx = 480 -> 6.28 M nodes,
             36.9 M tets, 25.1 M dofs (u + phi).  The north-star scaling workload and, with `--workload auto`
             (default), the PRIMARY line at every N.
  sen_shear  configs[1]: single-edge notched shear 2D (plot_1712.py), spectral split, n = 1000 -> 2.0 M P1 triangles.
             With `--workload auto` it is measured in the same run and reported under the key "config2".
  sent / fatigue   configs[0] / [2] on the reference mesh (5 089 nodes; committed fixture tests/golden/c1_mesh.npz);
             small meshes: one GPU only (replicas), reported as steps/s.
A "step" is one load step = the whole stagger loop (u-solve, projection/history, phi-solve).  Load steps
t = 0..W-1 are warm-up (t = 0 is the reference's trivial all-zero step), the next K steps are timed.  N > 1: the
same mesh is partitioned over the ranks (strong scaling).

One JSON line on stdout (rank 0).  `value` = K / device time of the K `pfx_load_step` calls (fields resident in
HBM, CUDA events on the library's stream, max over ranks); `e2e` = the same steps through the public `solve()`
drop-in with host buffers: Dirichlet values H2D, reactions / energies AND the u, phi fields D2H every step, result
files including the per-step VTU written (reference default save_solution_vtu=True), timed on the host.
`parity`: reactions and the 11 energies of the last timed step against tests/golden/bench_golden.json (made by
`--make-golden`: one GPU, every PCG solve to 1e-13 of its own right-hand side, no adaptive floor; Jacobi-PCG without
the coarse space where that is affordable).  `roofline`: the dominant kernel (J_u·v) timed in isolation with CUDA
events on the library's stream, L2 flushed between launches, algorithmic bytes of SURVEY §8d.  `cpu_baseline`: the
CPU oracle (scipy/SuperLU restatement of the reference algorithm — the reference's dolfinx+PETSc path cannot run in
this image) timed at several grid sizes, fitted (time per step ~ cells^p) and extrapolated to the bench mesh.
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

METRIC = "staggered_load_steps_per_second"
UNIT = "steps/s"
GOLDEN = os.path.join(ROOT, "tests", "golden", "bench_golden.json")
TOL_REACTION, TOL_ENERGY = 1e-6, 1e-8          # parity block: relative, against the tight-tolerance golden run


def log(*a):
    sys.stderr.write(" ".join(str(x) for x in a) + "\n")
    sys.stderr.flush()


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


# ------------------------------------------------------------------------------- workloads
def workload_key(name, args):
    """name of the instance a workload switch builds (without building the mesh)"""
    if name == "beam3d":
        nx = args.beam_nx
        return f"notched_beam_3d_spectral_{nx}x{nx // 3}x{nx // 6}"
    if name == "sen_shear":
        return f"sen_shear_2d_spectral_n{args.n}"
    return {"sent": "sent_2d_isotropic", "fatigue": "fatigue_sent_2d"}[name]


def build_workload(name, args, size=None):
    from phasefieldx_b200 import workloads as W
    if name == "beam3d":
        nx = size or args.beam_nx
        return W.notched_beam_3d(nx, nx // 3, nx // 6)
    if name == "sen_shear":
        return W.sen_shear(size or args.n)
    if name in ("sent", "fatigue"):
        from phasefieldx_b200.mesh import Mesh
        d = np.load(os.path.join(ROOT, "tests", "golden", "c1_mesh.npz"))
        msh = Mesh(d["points"], d["cells"], "triangle", gdim=2)
        return W.sen_tension(msh, d["bottom_nodes"], d["top_nodes"], fatigue=(name == "fatigue"))
    raise ValueError(name)


# ------------------------------------------------------------------------------- parity against the golden run
def load_golden():
    try:
        with open(GOLDEN) as fh:
            return json.load(fh)
    except Exception:
        return {}


def parity_block(wname, step, reactions, energies, stagger, golden=None):
    """Relative differences of the reactions (per Dirichlet set, scaled by the largest component of that set) and
    of the energies of load step `step` against the committed golden run of the same workload."""
    golden = load_golden() if golden is None else golden
    rec = golden.get(wname, {}).get("steps", {}).get(str(step))
    if rec is None:
        return {"available": False, "why": f"no golden record for workload {wname} step {step} in tests/golden/bench_golden.json"}
    rmax, emax = 0.0, 0.0
    worst_e = None
    for Rg, Ro in zip(reactions, rec["reactions"]):
        Rg, Ro = np.asarray(Rg, dtype=float), np.asarray(Ro, dtype=float)
        scale = np.abs(Ro).max()
        if scale > 0:
            rmax = max(rmax, float(np.abs(Rg - Ro).max() / scale))
    for key, ref in rec["energies"].items():
        if abs(ref) > 1e-300:
            d = abs(energies[key] - ref) / abs(ref)
            if d > emax:
                emax, worst_e = d, key
    return {"available": True, "step": int(step), "reaction_rel_diff": rmax, "energy_rel_diff": emax, "worst_energy": worst_e,
            "tol_reaction": TOL_REACTION, "tol_energy": TOL_ENERGY, "stagger": int(stagger), "stagger_golden": int(rec["stagger"]),
            "ok": bool(rmax <= TOL_REACTION and emax <= TOL_ENERGY and int(stagger) == int(rec["stagger"])),
            "golden": golden[wname].get("how", "")}


# ------------------------------------------------------------------------------- CPU oracle leg
def oracle_steps(name, args, size, steps, warmup):
    """Times the CPU oracle (the only place bench.py executes oracle/) on the workload at a reduced grid;
    returns (seconds per timed step, cells)."""
    from oracle.solver import BC, Oracle, Params
    w = build_workload(name, args, size=size)
    o = Oracle(w.msh.geometry.x, w.msh.cells, w.cell_type, Params(**w.params))
    bcs = [BC(d) for _, d in w.bc_sets]
    times = []
    for s in range(warmup + steps):
        for bc, v in zip(bcs, w.bc_values(float(s))):
            bc.values = v
        t0 = time.perf_counter()
        o.load_step(bcs, [], dt=w.dt, min_stagger_iter=w.min_stagger_iter, max_stagger_iter=w.max_stagger_iter, tol=w.stagger_tol)
        times.append(time.perf_counter() - t0)
    return sum(times[warmup:]) / steps, int(w.msh.num_cells)


def cpu_baseline(name, args, steps, warmup, full_cells):
    """Oracle at >= 3 grid sizes -> least-squares fit of log(time per step) against log(cells) -> rate at the bench
    mesh.  cores = 1: scipy / SuperLU factorise on one thread (torch's intra-op threads only serve the autograd
    tangent of the 3D split)."""
    env = os.environ.get("PFX_BENCH_SAMPLE_SIZES")
    if env:
        sizes = [int(s) for s in env.split(",")]
    elif name == "sen_shear":
        sizes = [80, 160, 240]
    elif name == "beam3d":
        # one tiny grid: the oracle's 3D spectral tangent is torch autograd of the reference formula (~35 s per load
        # step however small the mesh), so neither several sizes nor a fit are affordable or meaningful here
        sizes = [12]
    else:
        sizes = [None]
    pts = []
    for sz in sizes:
        sec, cells = oracle_steps(name, args, sz, steps, warmup)
        pts.append((sz, cells, sec))
        log(f"cpu oracle {name} size {sz}: {cells} cells, {sec:.3f} s per load step")
    if len(pts) >= 2:
        lx, ly = np.log([p[1] for p in pts]), np.log([p[2] for p in pts])
        p, c = np.polyfit(lx, ly, 1)
        sec_full = float(np.exp(c + p * np.log(full_cells)))
    else:
        p, sec_full = 1.0, pts[0][2] * max(1.0, full_cells / pts[0][1])      # linear in cells: generous to the CPU
    extrap = full_cells != pts[-1][1]
    desc = (f"oracle (numpy/scipy SuperLU restatement of the reference algorithm, 1 process) on the same workload at grid sizes "
            f"{[q[0] for q in pts]} = {[q[1] for q in pts]} cells: {[round(q[2], 3) for q in pts]} s per load step "
            f"(load steps t={warmup}..{warmup + steps - 1}); " + (f"fitted time ~ cells^{p:.2f}" if len(pts) >= 2 else "time taken as linear in cells")
            + (f", extrapolated to the {full_cells} cells of the bench mesh" if extrap else ""))
    if name == "beam3d":
        desc += ("; NB the oracle's 3D spectral tangent is torch autograd of the reference formula (a checker, not a tuned CPU "
                 "code), so this figure flatters the GPU and is context only")
    return {"value": 1.0 / sec_full, "unit": UNIT, "cores": 1, "host_cores": os.cpu_count(), "kind": "port", "sample": desc,
            "fit_exponent": float(p), "extrapolated": bool(extrap),
            "points": [{"size": q[0], "cells": q[1], "s_per_step": q[2]} for q in pts]}


def main_reference(args, rank):
    if rank != 0:
        return
    primary = "beam3d" if args.workload == "auto" else args.workload
    wu = min(args.warmup, 1)
    # bounded sample: at most 2 timed steps per grid size however large K is (the driver asks for K = 20)
    k = max(1, min(args.steps, 2))

    def arm(name):
        full = {"beam3d": 6 * args.beam_nx * (args.beam_nx // 3) * (args.beam_nx // 6), "sen_shear": 2 * args.n * args.n}.get(name, 9990)
        return cpu_baseline(name, args, k, wu, full)
    cb = arm(primary)
    line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / cb["value"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_key(primary, args), "timed_steps_per_grid_size": k, "warmup_steps_per_grid_size": wu,
                       "note": "dolfinx+PETSc cannot be installed in this image; the oracle port is timed instead, on reduced "
                               "grids of the same workload, and extrapolated with the fitted exponent (cpu_baseline.sample)"},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if args.workload == "auto":
        c2 = arm("sen_shear")
        line["config2"] = {"workload": workload_key("sen_shear", args), "value": c2["value"], "unit": UNIT, "cpu_baseline": c2}
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


def e2e_solve(w, steps, warmup, device, vtu_format="ascii", vtu_keep=None):
    """Same load steps through the public drop-in `solve()` (host buffers in; result files, per-step VTU with the
    u and phi fields out).  Returns (steps/s, h2d bytes, d2h bytes, parity record of the last timed step)."""
    from phasefieldx_b200 import fem
    from phasefieldx_b200.Element.Phase_Field_Fracture.Input import Input
    from phasefieldx_b200.Element.Phase_Field_Fracture.solver.solver import solve
    import torch.distributed as dist
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    box = [tempfile.mkdtemp(prefix="pfx_bench_") if (not multi or dist.get_rank() == 0) else None]
    if multi:
        dist.broadcast_object_list(box, src=0)
    tmp = box[0]
    Data = Input(save_solution_vtu=True, results_folder_name=os.path.join(tmp, "results"), **w.params)
    V_u = fem.functionspace(w.msh, ("Lagrange", 1, (w.msh.geometry.dim,)))
    V_phi = fem.functionspace(w.msh, ("Lagrange", 1))
    bcs = []
    for _, dofs in w.bc_sets:
        bc = fem.DirichletBC(0.0, np.zeros(0, np.int64), V_phi)
        bc._dofs, bc._comp = dofs.astype(np.int32), None
        bc._vals = np.zeros(len(dofs))
        bc.values_per_dof = (lambda b=bc: b._vals)
        bcs.append(bc)
    stamps, last = [], {}

    def update(bc_list, t):
        stamps.append(time.perf_counter())
        for bc, v in zip(bc_list, w.bc_values(float(t))):
            bc._vals = v
        return 1e-4 * t, 0, 0

    def after(step, rep, reactions, en):
        if step == warmup + steps - 1:
            last.update(step=step, stagger=rep["stagger"], reactions=[np.asarray(r).tolist() for r in reactions], energies=dict(en))
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
                  max_stagger_iter=w.max_stagger_iter, stagger_error_tol=w.stagger_tol, device=device,
                  vtu_format=vtu_format, vtu_keep=vtu_keep, vtu_parallel=multi, step_callback=after)
    finally:
        os.chdir(cwd)
    t_end = time.perf_counter()
    import shutil
    vtu_bytes = 0
    if not multi or dist.get_rank() == 0:
        vdir = os.path.join(tmp, "results", "paraview-solutions_vtu")
        if os.path.isdir(vdir):
            sizes = [os.path.getsize(os.path.join(vdir, f)) for f in os.listdir(vdir) if f.endswith(".vtu")]
            vtu_bytes = max(sizes) if sizes else 0
    if multi:
        dist.barrier()
    if not multi or dist.get_rank() == 0:
        shutil.rmtree(tmp, ignore_errors=True)
    d = w.msh.geometry.dim
    h2d = sum(len(dd) * 8 for _, dd in w.bc_sets)
    d2h = int(w.msh.num_nodes) * (d + 1) * 8 + len(w.bc_sets) * 3 * 8 + 11 * 8 + 256 * 6 * 8
    step_ms = [1e3 * (b - a) for a, b in zip(stamps[warmup:warmup + steps], stamps[warmup + 1:warmup + steps + 1])]
    log(f"e2e ({vtu_format}) step ms {[round(x, 1) for x in step_ms]}; tear-down + extra step {1e3 * (t_end - stamps[warmup + steps]):.1f} ms")
    return steps / (stamps[warmup + steps] - stamps[warmup]), h2d, d2h, last, vtu_bytes


PHASES = ["u_setup", "u_coarse_setup", "u_pcg", "projection", "fatigue", "phi_setup", "phi_coarse_setup", "phi_pcg", "norms", "other"]


def run_b200(name, args, rank, world, local_rank, primary):
    """Device-timed load steps + roofline + parity + e2e for one workload; returns the result dict (rank 0) or None."""
    import torch
    import torch.distributed as dist
    from phasefieldx_b200 import workloads as W
    t_setup = time.perf_counter()
    w = build_workload(name, args)
    opts = {}
    if args.cg_rtol:
        opts["cg_rtol"] = args.cg_rtol
    e, local_sets = build_rank_engine(w, rank, world, local_rank, opts)
    t_setup = time.perf_counter() - t_setup
    K, Wm = args.steps, args.warmup
    tets = w.cell_type == "tetrahedron"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    reports = []
    for s in range(Wm):
        set_bc_values(e, w, s, local_sets)
        e.load_step(dt=w.dt, min_stagger_iter=w.min_stagger_iter, max_stagger_iter=w.max_stagger_iter, stagger_tol=w.stagger_tol)
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
    # parity of the device-resident run: global reactions / energies of the last timed step (collective calls)
    reactions = [e.reaction(i).tolist() for i in range(len(w.bc_sets))]
    energies = e.energies()
    par = parity_block(w.name, Wm + K - 1, reactions, energies, reports[-1]["stagger"]) if rank == 0 else None

    # ---- roofline of the dominant kernel (J_u·v), state of the last step
    peak, peak_kind = measured_peak()
    roof = None
    if rank == 0:
        if world == 1:
            # operator micro-benchmark input of SURVEY §8d: v ~ U(-1, 1), seed 0, zero on Dirichlet dofs (the fields
            # u, phi, H are those of the last timed load step)
            v = np.random.default_rng(0).uniform(-1, 1, size=int(w.msh.num_nodes) * w.msh.geometry.dim)
            e.apply_u(v, use_bc=True)
        ms = e.time_operator(0, reps=23, flush_l2=True)[3:]
        ms_nf = e.time_operator(0, reps=13, flush_l2=False)[3:]
        bytes_apply = e.algorithmic_bytes(0)
        ach = bytes_apply / (ms.mean() * 1e-3) / 1e9
        if tets:
            kernel = "spmv_u_tet_kernel (J_u·v on P1 tetrahedra: 3x3-block sliced-ELL SpMV of the tangent assembled once per Newton linearisation)"
            traffic_file = "r02_spmv_u_tet_traffic.json"
        else:
            kernel = "apply_u_tri_v3_kernel<%s> (J_u·v on P1 triangles; matrix-free, persistent, TMA-staged)" % w.params["split_energy"]
            traffic_file = "r01_apply_u_traffic.json"
        traffic = None
        try:    # DRAM bytes of one launch from the committed `ncu --set full` capture of this kernel/workload
            tj = json.load(open(os.path.join(ROOT, "profiles", traffic_file)))
            if tj.get("workload") == w.name and world == 1:
                traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic, "algorithmic_bytes_per_launch": bytes_apply,
                "ms_per_launch": float(ms.mean()), "ms_per_launch_l2_warm": float(ms_nf.mean())}
        if tets:
            sb = e.algorithmic_bytes(5)
            roof["streamed_bytes_per_launch"] = sb
            roof["streamed_GBps"] = sb / (ms.mean() * 1e-3) / 1e9
            roof["streamed_frac"] = roof["streamed_GBps"] / peak
            roof["note"] = ("achieved/frac use the matrix-free ALGORITHMIC bytes of SURVEY §8d (200 B/node); the kernel itself streams the "
                            "assembled block matrix (streamed_*): an on-the-fly 3x3 spectral tangent is fp64-ALU-bound (SURVEY H4)")
        else:
            roof["note"] = "fp64 issue of the spectral split also bounds this kernel (see DESIGN.md)"
        if world == 1:
            ms_it = e.time_operator(4, reps=13, flush_l2=True)[3:]
            bi = e.algorithmic_bytes(4)
            roof["pcg_iteration"] = {"achieved": bi / (ms_it.mean() * 1e-3) / 1e9, "ms": float(ms_it.mean()),
                                     "algorithmic_bytes": bi, "frac": bi / (ms_it.mean() * 1e-3) / 1e9 / peak}
            ms_phi = e.time_operator(1, reps=13, flush_l2=True)[3:]
            bp = e.algorithmic_bytes(1)
            roof["apply_phi"] = {"achieved": bp / (ms_phi.mean() * 1e-3) / 1e9, "ms": float(ms_phi.mean()), "algorithmic_bytes": bp,
                                 "frac": bp / (ms_phi.mean() * 1e-3) / 1e9 / peak}
    stats = e.block_stats()
    cst = e.coarse_stats()
    solver = ("PCG, block-Jacobi + aggregate rigid-body-mode coarse space "
              f"({cst['aggregates']} aggregates per rank, {cst['setups']} set-ups so far)") if cst["active"] else "PCG, Jacobi"
    solver = ("assembled sliced-ELL operators, " if tets else "matrix-free ") + solver
    if world > 1:
        dist.barrier()          # peers keep each other's buffers mapped (CUDA IPC): close together
    e.close()

    # end to end through the public drop-in solve(): on several GPUs it runs collectively (one process per
    # GPU, the same torch.distributed group), rank 0 owns the result files and the clock
    e2e, e2e_bin = None, None
    if not args.no_e2e:
        # big 3D meshes: raw-appended VTU (the ASCII file of a 37 M-tet mesh is ~3 GB per step) and only the two most
        # recent step files kept on disk; 2D: the reference's ASCII layout, every file kept
        big = int(w.msh.num_cells) > 8_000_000
        fmt, keep = ("binary", 2) if big else ("ascii", None)
        v, h2d, d2h, last, vb = e2e_solve(w, K, Wm, local_rank, vtu_format=fmt, vtu_keep=keep)
        if rank == 0:
            e2e = {"value": v, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "vtu": fmt,
                   "vtu_bytes_per_step": vb, "vtu_files_kept": keep or "all",
                   "parity": parity_block(w.name, last.get("step", -1), last.get("reactions", []), last.get("energies", {}),
                                          last.get("stagger", -1)) if last else None}
        if not big and not args.no_e2e_binary:
            v2, _, _, _, vb2 = e2e_solve(w, K, Wm, local_rank, vtu_format="binary")
            if rank == 0:
                e2e_bin = {"value": v2, "unit": UNIT, "vtu": "binary", "vtu_bytes_per_step": vb2}
    if rank != 0:
        return None
    value = K / (gpu_ms * 1e-3)
    phase = {p: float(sum(r["phase_ms"][p] for r in reports)) for p in PHASES}
    st_its = sum(r["stagger"] for r in reports)
    cg_all = sum(int(r["cg_its_u"]) + int(r["cg_its_phi"]) + int(r["cg_its_proj"]) for r in reports)
    res = {"value": value, "ms_per_step": gpu_ms / K,
           "config": {"workload": w.name, "nodes": int(w.msh.num_nodes), "cells": int(w.msh.num_cells),
                      "dofs_u_plus_phi": int((w.msh.geometry.dim + 1) * w.msh.num_nodes), "load_steps_timed": [Wm, Wm + K - 1],
                      "stagger": [w.min_stagger_iter, w.max_stagger_iter, w.stagger_tol],
                      "l2": "working set of one PCG iteration exceeds the 126 MB L2; operator timing flushes L2",
                      "parallelism": f"mesh partitioned over {world} GPU(s)", "linear_solver": solver, "blocks": stats,
                      "setup_s": t_setup},
           "timing": {"device_ms": gpu_ms, "host_wall_ms": wall_ms, "phase_ms": phase,
                      "stagger_iters": [r["stagger"] for r in reports],
                      "cg_its_u": [int(r["cg_its_u"]) for r in reports],
                      "cg_its_phi": [int(r["cg_its_phi"]) for r in reports],
                      "cg_its_proj": [int(r["cg_its_proj"]) for r in reports],
                      "stagger_iterations_per_s": st_its / (gpu_ms * 1e-3), "pcg_iterations_per_s": cg_all / (gpu_ms * 1e-3)},
           "parity": par, "clocks": clk.summary(), "gpu_launches": int(launches), "roofline": roof, "e2e": e2e}
    if e2e_bin is not None:
        res["e2e_binary_vtu"] = e2e_bin
    return res


def make_golden(args):
    """Golden records of the parity block: one GPU, every PCG solve to 1e-13 of its own right-hand side
    (PFX_CG_ADAPTIVE=0); triangles: Jacobi-PCG without the coarse space (PFX_NO_COARSE=1); tetrahedra: the two-level
    PCG (a Jacobi-PCG needs ~1e4 iterations per solve at 25 M dofs)."""
    from phasefieldx_b200 import workloads as W
    names = ["beam3d", "sen_shear"] if args.workload == "auto" else [args.workload]
    gold = load_golden()
    for name in names:
        os.environ["PFX_CG_ADAPTIVE"] = "0"
        jac = name != "beam3d" or args.golden_jacobi
        if jac:
            os.environ["PFX_NO_COARSE"] = "1"
        else:
            os.environ.pop("PFX_NO_COARSE", None)
        w = build_workload(name, args)
        e = W.make_engine(w, cg_rtol=1e-13)
        steps = {}
        t0 = time.perf_counter()
        for s in range(args.warmup + args.steps):
            set_bc_values(e, w, s, None)
            r = e.load_step(dt=w.dt, min_stagger_iter=w.min_stagger_iter, max_stagger_iter=w.max_stagger_iter, stagger_tol=w.stagger_tol)
            steps[str(s)] = {"stagger": int(r["stagger"]), "reactions": [e.reaction(i).tolist() for i in range(len(w.bc_sets))],
                             "energies": e.energies(), "cg_its_u": int(r["cg_its_u"])}
            log(f"golden {w.name} step {s}: stagger {r['stagger']}, u-PCG its {r['cg_its_u']}, {time.perf_counter() - t0:.1f} s")
        e.close()
        gold[w.name] = {"how": ("one B200, cg_rtol 1e-13 relative to every right-hand side (PFX_CG_ADAPTIVE=0), "
                                + ("Jacobi-PCG without coarse space (PFX_NO_COARSE=1)" if jac else "two-level PCG")),
                        "steps": steps}
    out = args.golden_out or GOLDEN
    with open(out, "w") as fh:
        json.dump(gold, fh, indent=0, sort_keys=True)
    log("golden written to", out)


def main_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: phasefieldx_b200 has no CPU fallback")
    if args.make_golden:
        return make_golden(args)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")        # keep NCCL's version banner off stdout (one JSON line)
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    primary = "beam3d" if args.workload == "auto" else args.workload
    if primary in ("sent", "fatigue") and world > 1:
        raise RuntimeError("configs 1 / 3 (5 089-node mesh) do not shard: replicas only, run them with --gpus 1")
    res = run_b200(primary, args, rank, world, local_rank, True)
    res2 = run_b200("sen_shear", args, rank, world, local_rank, False) if args.workload == "auto" else None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            full = int(res["config"]["cells"])
            cpu = cpu_baseline(primary, args, 1, 1, full)
            if res2 is not None:
                res2["cpu_baseline"] = cpu_baseline("sen_shear", args, 1, 1, int(res2["config"]["cells"]))
        line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": res["config"], "timing": res["timing"], "parity": res["parity"],
                "clocks": res["clocks"], "gpu_launches": res["gpu_launches"], "roofline": res["roofline"], "cpu_baseline": cpu,
                "e2e": res["e2e"]}
        if "e2e_binary_vtu" in res:
            line["e2e_binary_vtu"] = res["e2e_binary_vtu"]
        if res2 is not None:
            res2["metric"], res2["unit"] = METRIC, UNIT
            line["config2"] = res2
        for blk in (line["parity"], (line["e2e"] or {}).get("parity")):
            if blk and blk.get("available") and not blk["ok"]:
                log("PARITY FAILURE against the golden run:", json.dumps(blk))
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
    ap.add_argument("--workload", default="auto", choices=["auto", "beam3d", "sen_shear", "sent", "fatigue"])
    ap.add_argument("--beam-nx", type=int, default=480, help="cubes along the beam (nx x nx/3 x nx/6 x 6 tetrahedra)")
    ap.add_argument("--n", type=int, default=1000, help="grid size of the shear workload (n x n x 2 triangles)")
    ap.add_argument("--cg-rtol", type=float, default=0.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-binary", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--make-golden", action="store_true", help="write tests/golden/bench_golden.json records (tight solver settings)")
    ap.add_argument("--golden-jacobi", action="store_true", help="--make-golden: Jacobi-PCG for the 3D workload too")
    ap.add_argument("--golden-out", default=None)
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
