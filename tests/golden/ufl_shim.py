"""
Torch-fp64 stand-in for the handful of `ufl` symbols the reference's pointwise constitutive
layer uses, so that the reference files run UNMODIFIED (from /root/reference, build container
only) and their outputs can be committed as golden vectors (SURVEY §8c "partial exception").
`ufl.diff` ≡ torch.autograd.grad(create_graph=True); d|x|/dx = sign x as in UFL.

Used only by tests/golden/make_golden.py.  Nothing here is product code and nothing from the
reference is copied: the reference modules are loaded by path at generation time.
"""
import importlib.util
import math
import os
import sys
import types

import numpy as np
import torch

REF_SRC = "/root/reference/src/phasefieldx"


class T:
    """Batched tensor-valued expression: data [n, *shape]."""

    def __init__(self, t, shape=()):
        self.t, self.shape = t, tuple(shape)

    # ---- helpers
    @staticmethod
    def lift(x, like=None):
        if isinstance(x, T):
            return x
        return T(torch.as_tensor(float(x), dtype=torch.float64), ())

    def _b(self, other_rank):
        """expand a scalar for broadcasting against a tensor of the given rank"""
        t = self.t
        for _ in range(other_rank):
            t = t[..., None]
        return t

    def __add__(self, o):
        o = T.lift(o)
        if self.shape == o.shape:
            return T(self.t + o.t, self.shape)
        if self.shape == ():
            return T(self._b(len(o.shape)) + o.t, o.shape)
        return T(self.t + o._b(len(self.shape)), self.shape)

    __radd__ = __add__

    def __neg__(self):
        return T(-self.t, self.shape)

    def __sub__(self, o):
        return self + (-T.lift(o))

    def __rsub__(self, o):
        return T.lift(o) + (-self)

    def __mul__(self, o):
        o = T.lift(o)
        if len(self.shape) == 2 and len(o.shape) == 2:          # UFL: A * B is the matrix product
            return T(self.t @ o.t, (self.shape[0], o.shape[1]))
        if self.shape == ():
            return T(self._b(len(o.shape)) * o.t, o.shape)
        if o.shape == ():
            return T(self.t * o._b(len(self.shape)), self.shape)
        raise NotImplementedError

    def __rmul__(self, o):
        return T.lift(o) * self

    def __truediv__(self, o):
        o = T.lift(o)
        assert o.shape == ()
        return T(self.t / o._b(len(self.shape)), self.shape)

    def __rtruediv__(self, o):
        return T(T.lift(o).t / self.t, ())

    def __pow__(self, p):
        assert self.shape == ()
        return T(self.t ** p, ())

    def __abs__(self):
        return T(torch.abs(self.t), self.shape)

    def __getitem__(self, idx):
        i, j = idx
        return T(self.t[..., i, j], ())

    @property
    def T(self):
        return T(self.t.transpose(-1, -2), self.shape[::-1])


class _U:
    """Fake displacement: carries the batched gradient [n,d,d] (requires_grad leaf)."""

    def __init__(self, G):
        self.G = G
        self.d = G.shape[-1]

    def __len__(self):
        return self.d

    def ufl_domain(self):
        d = self.d
        return types.SimpleNamespace(geometric_dimension=lambda: d)


def make_ufl():
    u = types.ModuleType("ufl")
    u.pi = math.pi
    u.tr = lambda A: T(A.t.diagonal(dim1=-2, dim2=-1).sum(-1), ())
    u.det = lambda A: T(torch.linalg.det(A.t), ())
    u.sqrt = lambda x: T(torch.sqrt(T.lift(x).t), ())
    u.cos = lambda x: T(torch.cos(x.t), ())
    u.acos = lambda x: T(torch.acos(x.t), ())
    u.atan2 = lambda y, x: T(torch.atan2(T.lift(y).t, T.lift(x).t), ())
    u.shape = lambda A: A.shape
    u.variable = lambda A: A
    u.diff = lambda f, A: T(torch.autograd.grad(f.t.sum(), A.t, create_graph=True)[0], A.shape)
    u.Identity = lambda d: T(torch.eye(d, dtype=torch.float64), (d, d))
    u.grad = lambda w: T(w.G, (w.d, w.d))
    u.sym = lambda A: T(0.5 * (A.t + A.t.transpose(-1, -2)), A.shape)
    u.inner = lambda A, B: T((A.t * B.t).sum((-1, -2)), ())
    u.dev = lambda A: A - (u.tr(A) / A.shape[0]) * u.Identity(A.shape[0])
    u.max_value = lambda a, b: T(torch.maximum(T.lift(a).t, T.lift(b).t), ())
    u.min_value = lambda a, b: T(torch.minimum(T.lift(a).t, T.lift(b).t), ())
    u.zero = lambda shape: T(torch.zeros(shape, dtype=torch.float64), shape)
    return u


def load_reference_modules():
    sys.modules["ufl"] = make_ufl()
    for pkg in ("phasefieldx", "phasefieldx.Math", "phasefieldx.Materials", "phasefieldx.Element",
                "phasefieldx.Element.Phase_Field_Fracture"):
        sys.modules[pkg] = types.ModuleType(pkg)
    mods = {}
    for name, rel in (("phasefieldx.Math.functions", "Math/functions.py"),
                      ("phasefieldx.Math.invariants", "Math/invariants.py"),
                      ("phasefieldx.Math.tensor_decomposition", "Math/tensor_decomposition.py"),
                      ("phasefieldx.Materials.elastic_isotropic", "Materials/elastic_isotropic.py"),
                      ("phasefieldx.Element.Phase_Field_Fracture.split_energy_stress_tangent_functions",
                       "Element/Phase_Field_Fracture/split_energy_stress_tangent_functions.py")):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_SRC, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        mods[name.rsplit(".", 1)[1]] = m
    return mods


def reference_constitutive(seed=1234, n=64):
    """ψ_a, ψ_b, σ_a, σ_b and the tangent action dσ_a[dε], dσ_b[dε] of the reference code at seeded
    random displacement gradients (several strain scales, incl. degenerate states)."""
    mods = load_reference_modules()
    S = mods["split_energy_stress_tangent_functions"]
    rng = np.random.default_rng(seed)
    out = {}
    lam, mu = 121.15384615384615, 80.76923076923077
    out["lam_mu"] = np.array([lam, mu])
    for d in (2, 3):
        G = rng.normal(size=(n, d, d))
        scale = np.repeat([1e-2, 1e-3, 1e-5, 1e-7], n // 4)
        G = G * scale[:, None, None]
        G[0] = 0.0                                                 # zero strain
        G[1] = np.eye(d) * 1e-3                                    # hydrostatic
        G[2] = np.diag([1.0, -0.3, -0.3][:d]) * 1e-3               # uniaxial-like (degenerate pair in 3D)
        dG = rng.normal(size=(n, d, d))
        out[f"G{d}"], out[f"dG{d}"] = G, dG
        for deg, split in (("isotropic", "no"), ("anisotropic", "spectral"), ("anisotropic", "deviatoric"),
                           ("hybrid", "spectral"), ("hybrid", "deviatoric")):
            Data = types.SimpleNamespace(degradation=deg, split_energy=split, lambda_=lam, mu=mu)
            Gt = torch.tensor(G, dtype=torch.float64, requires_grad=True)
            u = _U(Gt)
            key = f"{d}_{deg}_{split}"
            pa, pb = S.psi_a(u, Data), S.psi_b(u, Data)
            sa, sb = S.sigma_a(u, Data), S.sigma_b(u, Data)
            out[f"psi_a_{key}"] = pa.t.detach().numpy() * np.ones(n)
            out[f"psi_b_{key}"] = pb.t.detach().numpy() * np.ones(n)
            dGt = torch.tensor(dG)
            for nm, s in (("a", sa), ("b", sb)):
                st = s.t if s.t.dim() == 3 else s.t.expand(n, d, d)
                out[f"sig_{nm}_{key}"] = st.detach().numpy().copy()
                ds = np.zeros((n, d, d))
                if st.requires_grad:
                    for i in range(d):
                        for j in range(d):
                            g, = torch.autograd.grad(st[:, i, j].sum(), Gt, retain_graph=True, allow_unused=True)
                            if g is not None:
                                ds[:, i, j] = (g * dGt).sum((-1, -2)).detach().numpy()
                out[f"dsig_{nm}_{key}"] = ds
    return out
