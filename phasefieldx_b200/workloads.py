"""
Synthetic instances of the BASELINE.json configurations (SURVEY §8d "Synthetic inputs per
config"): deterministic meshes, Dirichlet sets and loading histories that need no mesher.

Each builder returns a `Workload`: mesh, material `Input` keyword set, a list of Dirichlet sets
(name, unrolled dofs, per-dof component) and `bc_values(t)` giving the per-dof values of every
set at pseudo-time t.  Used by bench.py, __graft_entry__.smoke() and the tests.
"""
import numpy as np

from . import mesh as M


class Workload:
    def __init__(self, name, msh, params, bc_sets, bc_values, stagger=(2, 500, 1e-8), dt=1.0, final_time=150.0):
        self.name, self.msh, self.params = name, msh, params
        self.bc_sets, self.bc_values = bc_sets, bc_values
        self.min_stagger_iter, self.max_stagger_iter, self.stagger_tol = stagger
        self.dt, self.final_time = dt, final_time

    @property
    def cell_type(self):
        return self.msh.cell_type


def _dofs(nodes, d, comps):
    nodes = np.asarray(nodes, dtype=np.int64)
    return (nodes[:, None] * d + np.asarray(comps)[None, :]).ravel().astype(np.int32)


def sen_shear(n=1000, l=0.06):
    """Config 2: single-edge notched shear, unit square [-.5,.5]², n×n squares split into 2 P1
    triangles, slit y=0, x<0 by node duplication; anisotropic + spectral
    (reference examples/PhaseFieldFracture/plot_1712.py:96-110, 186-196, 233-237)."""
    msh = M.create_rectangle(None, [[-0.5, -0.5], [0.5, 0.5]], [n, n])
    msh = M.cut_slit(msh, lambda x: np.isclose(x[1], 0.0) & (x[0] < -1e-9), lambda c: c[1] > 0)
    x = msh.geometry.x
    top = np.nonzero(np.isclose(x[:, 1], 0.5))[0]
    bot = np.nonzero(np.isclose(x[:, 1], -0.5))[0]
    inner = np.abs(x[:, 1]) < 0.5 - 1e-9
    left = np.nonzero(np.isclose(x[:, 0], -0.5) & inner)[0]
    right = np.nonzero(np.isclose(x[:, 0], 0.5) & inner)[0]
    sets = [("top", _dofs(top, 2, [0, 1])), ("bottom", _dofs(bot, 2, [0, 1])),
            ("left", _dofs(left, 2, [1])), ("right", _dofs(right, 2, [1]))]

    def bc_values(t):
        vals = [np.zeros(len(s[1])) for s in sets]
        vals[0][0::2] = 1e-4 * t
        return vals
    params = dict(E=210.0, nu=0.3, Gc=0.0027, l=l, degradation="anisotropic", split_energy="spectral",
                  degradation_function="quadratic", irreversibility="miehe", fatigue=False, k=0.0)
    return Workload(f"sen_shear_2d_spectral_n{n}", msh, params, sets, bc_values)


def sen_tension(msh, bottom_nodes, top_nodes, fatigue=False):
    """Configs 1 / 3 on the reference mesh (examples/PhaseFieldFracture/plot_1711.py:94-108,
    189-243; examples/Fatigue/plot_1800.py:102-116, 223-228)."""
    sets = [("top", _dofs(top_nodes, 2, [1])), ("bottom", _dofs(bottom_nodes, 2, [0, 1]))]
    if fatigue:
        def bc_values(t):
            v = (2.0 / np.pi) * 0.002 * np.arcsin(np.sin(2 * np.pi * t / 8.0))
            return [np.full(len(sets[0][1]), v), np.zeros(len(sets[1][1]))]
        params = dict(E=210.0, nu=0.3, Gc=0.0027, l=0.004, degradation="isotropic", split_energy="no",
                      degradation_function="quadratic", irreversibility="miehe", fatigue=True,
                      fatigue_degradation_function="asymptotic", fatigue_val=0.05625, k=0.0)
        return Workload("fatigue_sent_2d", msh, params, sets, bc_values, final_time=8 * 1000 + 1)

    def bc_values(t):
        v = 1e-4 * t if t <= 50 else 5e-3 + 1e-5 * (t - 50)
        return [np.full(len(sets[0][1]), v), np.zeros(len(sets[1][1]))]
    params = dict(E=210.0, nu=0.3, Gc=0.0027, l=0.015, degradation="isotropic", split_energy="no",
                  degradation_function="quadratic", irreversibility="miehe", fatigue=False, k=0.0)
    return Workload("sent_2d_isotropic", msh, params, sets, bc_values)


def notched_beam_3d(nx=96, ny=32, nz=16, l=None):
    """Config 5: Damage-Mechanics-Challenge beam 76.2×25.4×12.7 (reference
    examples/PhaseFieldFracture/plot_1718.py:114-128, 175-248, 286-290) on a Kuhn-split grid with a
    vertical slit notch of depth 5.08 at mid-span made by node duplication."""
    L, Hh, W = 76.2, 25.4, 12.7
    msh = M.create_box(None, [[0, 0, 0], [L, Hh, W]], [nx, ny, nz])
    hx = L / nx
    xm = round(38.1 / hx) * hx
    msh = M.cut_slit(msh, lambda x: np.isclose(x[0], xm) & (x[1] < 5.08 - 1e-9), lambda c: c[0] > xm)
    x = msh.geometry.x
    on = lambda c, w: (x[:, 0] > c - w) & (x[:, 0] < c + w)
    bl = np.nonzero(on(7.62, 4.0) & np.isclose(x[:, 1], 0.0))[0]
    br = np.nonzero(on(68.58, 4.0) & np.isclose(x[:, 1], 0.0))[0]
    tp = np.nonzero(on(38.1, 2.0) & np.isclose(x[:, 1], Hh))[0]
    sets = [("top", _dofs(tp, 3, [1])), ("bottom_left", _dofs(bl, 3, [0, 1, 2])), ("bottom_right", _dofs(br, 3, [1]))]

    def bc_values(t):
        return [np.full(len(sets[0][1]), -1e-2 * t), np.zeros(len(sets[1][1])), np.zeros(len(sets[2][1]))]
    h = max(L / nx, Hh / ny, W / nz)
    params = dict(E=0.6, nu=0.2, Gc=0.00013, l=(max(1.5, 4 * h) if l is None else l), degradation="anisotropic",
                  split_energy="spectral", degradation_function="quadratic", irreversibility="miehe", fatigue=False, k=0.0)
    return Workload(f"notched_beam_3d_spectral_{nx}x{ny}x{nz}", msh, params, sets, bc_values, final_time=40.0)


def make_engine(w, device=0, **opts):
    """Engine for a workload with its Dirichlet sets registered (values of t = 0)."""
    from . import _lib
    from .Materials.conversion import get_lambda_lame, get_mu_lame
    p = w.params
    lam, mu = get_lambda_lame(p["E"], p["nu"]), get_mu_lame(p["E"], p["nu"])
    model = _lib.model_code(p["degradation"], None if p["split_energy"] == "no" else p["split_energy"])
    e = _lib.Engine(w.msh.geometry.x, w.msh.cells, w.cell_type, lam, mu, p["Gc"], p["l"], k=p["k"], model=model,
                    irreversibility=(p["irreversibility"] == "miehe"), fatigue=p["fatigue"],
                    fatigue_val=p.get("fatigue_val", 0.05625), device=device,
                    degradation_function=p.get("degradation_function", "quadratic"), **opts)
    for i, (_, dofs) in enumerate(w.bc_sets):
        e.set_bc_u(i, dofs)
    for i, v in enumerate(w.bc_values(0.0)):
        e.set_bc_values_u(i, v)
    return e


def operator_state(w):
    """Operator micro-benchmark state of SURVEY §8d: φ = exp(−dist(x, notch line)/l) clipped to
    [0,1]; u = (1e-3·y·(1+x), 5e-4·x·y[, 2e-4·z]); v ~ U(−1,1), seed 0."""
    x = w.msh.geometry.x
    d = w.msh.geometry.dim
    l = w.params["l"]
    if d == 2:
        dist = np.where(x[:, 0] < 0, np.abs(x[:, 1]), np.hypot(x[:, 0], x[:, 1]))
        u = np.stack([1e-3 * x[:, 1] * (1 + x[:, 0]), 5e-4 * x[:, 0] * x[:, 1]], 1)
    else:
        dist = np.abs(x[:, 0] - 38.1) + np.maximum(x[:, 1] - 5.08, 0.0)
        u = np.stack([1e-3 * x[:, 1] * (1 + x[:, 0] / 76.2), 5e-4 * x[:, 0] * x[:, 1] / 76.2, 2e-4 * x[:, 2]], 1)
    phi = np.clip(np.exp(-dist / l), 0.0, 1.0)
    v = np.random.default_rng(0).uniform(-1, 1, size=u.size)
    return u.ravel(), phi, v


def three_point_bending(divx=400, divy=100, variational=True):
    """Config 4: energy-controlled three-point bending, symmetric half model [-4,0]x[0,1] of Q1 cells
    (reference examples/PhaseFieldFractureEnergyControlled/plot_three_point_var.py:96-131, 229-271 and
    plot_three_point_non_var.py): bc_y on the bottom-left support, bc_x on the symmetry line above the notch,
    traction (0, -1/0.075) on the loaded patch of the top edge.  Returns (mesh, Input keywords, Dirichlet sets,
    (traction nodes, nodal weights, T), solver keywords)."""
    from . import fem
    lx, ly, a0, surface = 4.0, 1.0, 0.2, 0.075
    msh = M.create_rectangle(None, [np.array([-lx, 0.0]), np.array([0.0, ly])], [divx, divy], M.CellType.quadrilateral)
    fb = M.locate_entities_boundary(msh, 1, lambda x: np.logical_and(np.isclose(x[1], 0), np.less(x[0], -lx + surface)))
    fc = M.locate_entities_boundary(msh, 1, lambda x: np.logical_and(np.isclose(x[0], 0), np.greater_equal(x[1], a0)))
    ft = M.locate_entities_boundary(msh, 1, lambda x: np.logical_and(np.isclose(x[1], ly), np.greater_equal(x[0], -surface)))
    sets = [("bottom_left", _dofs(M.facet_nodes(msh, fb), 2, [1])), ("center", _dofs(M.facet_nodes(msh, fc), 2, [0]))]
    nodes, w = fem.Measure(msh, ft).nodal_weights()
    params = dict(E=20.8, nu=0.3, Gc=0.0005, l=0.03, degradation="isotropic", split_energy="no",
                  degradation_function="quadratic", irreversibility="no", fatigue=False, k=0.0)
    kw = dict(final_gamma=0.5, dtau=0.001, dtau_min=1e-12, dtau_max=1.0, c1=35.0 if variational else 1.0, c2=1.0)
    return msh, params, sets, (nodes, w, np.array([0.0, -1.0 / surface])), kw
