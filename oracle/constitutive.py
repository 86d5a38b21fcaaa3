"""
ORACLE (test infrastructure, not product code) — pointwise constitutive layer.

CPU restatement of the reference's strain-energy splits:
  * isotropic law ψ, σ                       reference Materials/elastic_isotropic.py:26-70
  * Macaulay brackets ⟨x⟩± = ½(x ± |x|)       reference Math/functions.py:10-41
  * regularised closed-form eigen-states      reference Math/invariants.py:85-204 (eps = 3e-16)
  * spectral ± parts                          reference Math/tensor_decomposition.py:62-132
  * ψ_a/ψ_b/σ_a/σ_b spectral and vol-dev      reference .../split_energy_stress_tangent_functions.py:34-231
  * dispatch incl. `hybrid`                   same file :234-354
The consistent tangent is "AD of AD" in the reference (E_k = (∂λ_k/∂A)ᵀ through `ufl.diff`,
Jacobian through `ufl.derivative`, solver.py:186).  Two restatements live here:
  - `split_torch`: formulas in torch fp64, tangent by `torch.autograd` (mirrors UFL's AD,
    d|x|/dx = sign x); used for 3D spectral and as cross-check for everything else;
  - `split_numpy`: hand-derived closed forms (isotropic, vol-dev, 2×2 spectral), fast path.
All tensors are full d×d; tangents are C[ij,kl] = ∂σ_ij/∂A_kl with the d² entries independent.
"""
import numpy as np

EPS = 3.0e-16      # reference Math/invariants.py:96,185


def mac_p(x):
    return 0.5 * (x + np.abs(x))


def mac_n(x):
    return 0.5 * (x - np.abs(x))


def _hp(x):
    """d⟨x⟩⁺/dx with UFL's derivative of abs (sign, zero at zero)."""
    return 0.5 * (1.0 + np.sign(x))


# --------------------------------------------------------------------------- numpy closed forms
def _iso_np(A, lam, mu):
    d = A.shape[-1]
    I = np.eye(d)
    tr = np.trace(A, axis1=-2, axis2=-1)
    psi = 0.5 * lam * tr**2 + mu * np.einsum("...ij,...ij", A, A)
    sig = lam * tr[..., None, None] * I + 2 * mu * A
    C = lam * np.einsum("ij,kl->ijkl", I, I) + 2 * mu * np.einsum("ik,jl->ijkl", I, I)
    C = np.broadcast_to(C, A.shape[:-2] + (d, d, d, d))
    return psi, sig, C


def _voldev_np(A, lam, mu):
    d = A.shape[-1]
    I = np.eye(d)
    tr = np.trace(A, axis1=-2, axis2=-1)
    k0 = lam + 2.0 / d * mu
    dev = A - tr[..., None, None] / d * I
    psi_a = 0.5 * k0 * mac_p(tr) ** 2 + mu * np.einsum("...ij,...ij", dev, dev)
    psi_b = 0.5 * k0 * mac_n(tr) ** 2
    sig_a = k0 * mac_p(tr)[..., None, None] * I + 2 * mu * dev
    II = np.einsum("ij,kl->ijkl", I, I)
    Id = np.einsum("ik,jl->ijkl", I, I)
    C_a = k0 * _hp(tr)[..., None, None, None, None] * II + 2 * mu * (Id - II / d)
    return psi_a, psi_b, sig_a, C_a


def _spectral2_np(A, lam, mu):
    """2×2 regularised spectral split, hand-derived from eigenstate2 (invariants.py:174-204)."""
    a, b01, b10, c = A[..., 0, 0], A[..., 0, 1], A[..., 1, 0], A[..., 1, 1]
    tr = a + c
    Dl = (a - c) ** 2 + 4 * b01 * b10 + EPS**2
    s = np.sqrt(Dl)
    l0, l1 = 0.5 * (tr - s), 0.5 * (tr + s)
    I = np.eye(2)
    # M = (∂s/∂A)ᵀ·s : [[a-c, 2 A01],[2 A10, c-a]] after the transpose in E = diff(λ, A).T
    M = np.stack([np.stack([a - c, 2 * b01], -1), np.stack([2 * b10, c - a], -1)], -2)
    Ms = M / s[..., None, None]
    E0, E1 = 0.5 * (I - Ms), 0.5 * (I + Ms)
    p0, p1, n0, n1 = mac_p(l0), mac_p(l1), mac_n(l0), mac_n(l1)
    eP = p0[..., None, None] * E0 + p1[..., None, None] * E1
    eN = n0[..., None, None] * E0 + n1[..., None, None] * E1
    psi_a = 0.5 * lam * mac_p(tr) ** 2 + mu * np.einsum("...ij,...ij", eP, eP)
    psi_b = 0.5 * lam * mac_n(tr) ** 2 + mu * np.einsum("...ij,...ij", eN, eN)
    sig_a = lam * mac_p(tr)[..., None, None] * I + 2 * mu * eP
    # tangent: d eP = Σ_k [ H(λk) dλk E_k + ⟨λk⟩⁺ dE_k ],  dλ = ½(dtr ∓ ds), dE = ∓½ d(M/s)
    # ds/dA_kl = Mg_kl / s with Mg = ∂Δ/∂A /2 = [[a-c, 2 A10],[2 A01, c-a]] (un-transposed)
    Mg = np.stack([np.stack([a - c, 2 * b10], -1), np.stack([2 * b01, c - a], -1)], -2)
    ds = Mg / s[..., None, None]                                    # [..,k,l]
    dtr = np.broadcast_to(I, ds.shape)
    dl0, dl1 = 0.5 * (dtr - ds), 0.5 * (dtr + ds)
    # dM_ij/dA_kl (constant): M00 = A00-A11, M01 = 2A01, M10 = 2A10, M11 = A11-A00
    dM = np.zeros((2, 2, 2, 2))
    dM[0, 0, 0, 0], dM[0, 0, 1, 1] = 1, -1
    dM[1, 1, 1, 1], dM[1, 1, 0, 0] = 1, -1
    dM[0, 1, 0, 1] = 2
    dM[1, 0, 1, 0] = 2
    dMs = dM / s[..., None, None, None, None] - np.einsum("...ij,...kl->...ijkl", Ms, ds) / s[..., None, None, None, None]
    dE1 = 0.5 * dMs
    dE0 = -dE1
    deP = (np.einsum("...,...kl,...ij->...ijkl", _hp(l0), dl0, E0) + p0[..., None, None, None, None] * dE0
           + np.einsum("...,...kl,...ij->...ijkl", _hp(l1), dl1, E1) + p1[..., None, None, None, None] * dE1)
    II = np.einsum("ij,kl->ijkl", I, I)
    C_a = lam * _hp(tr)[..., None, None, None, None] * II + 2 * mu * deP
    return psi_a, psi_b, sig_a, C_a


def split_numpy(A, lam, mu, degradation, split_energy, want_tangent=True):
    """Returns dict(psi_a, psi_b, sig_a, sig_b, C_a, C_b) following the reference dispatch
    (split_energy_stress_tangent_functions.py:234-354): isotropic ⇒ ψ_a=ψ, σ_a=σ, σ_b=0;
    anisotropic ⇒ both split; hybrid ⇒ σ unsplit but ψ split."""
    d = A.shape[-1]
    psi, sig, C = _iso_np(A, lam, mu)
    zero_s, zero_C = np.zeros_like(sig), np.zeros_like(C)
    if degradation == "isotropic":
        return dict(psi_a=psi, psi_b=psi - psi, sig_a=sig, sig_b=zero_s, C_a=C, C_b=zero_C)
    if degradation not in ("anisotropic", "hybrid"):
        raise ValueError("Invalid")
    if split_energy == "deviatoric":
        psi_a, psi_b, sig_a, C_a = _voldev_np(A, lam, mu)
    elif split_energy == "spectral":
        if d == 2:
            psi_a, psi_b, sig_a, C_a = _spectral2_np(A, lam, mu)
        else:
            r = split_torch(A, lam, mu, "anisotropic", "spectral", want_tangent)
            psi_a, psi_b, sig_a, C_a = r["psi_a"], r["psi_b"], r["sig_a"], r.get("C_a", np.zeros_like(C))
    else:
        raise ValueError("Invalid")
    if degradation == "hybrid":
        return dict(psi_a=psi_a, psi_b=psi_b, sig_a=sig, sig_b=zero_s, C_a=C, C_b=zero_C)
    return dict(psi_a=psi_a, psi_b=psi_b, sig_a=sig_a, sig_b=sig - sig_a, C_a=C_a, C_b=C - C_a)


# --------------------------------------------------------------------------- torch "AD of AD"
# Sum-of-products form of the 3×3 discriminant (reference Math/invariants.py:105-144).  Each
# string lists signed monomials of entries "ij"; Δy is Δx with every index pair transposed.
_DX3 = [
    "+01.12.20 -02.10.21",
    "+01.01.12 -01.02.11 +01.02.22 -02.02.21",
    "+00.01.21 -01.01.20 -01.21.22 +02.21.21",
    "+00.02.12 +01.12.12 -02.02.10 -02.11.12",
    "+00.01.12 -01.02.10 -01.12.22 +02.12.21",
    "+00.02.21 -01.02.20 +01.12.21 -02.11.21",
    "+01.10.12 -02.10.11 +02.10.22 -02.12.20",
    "+00.00.12 -00.02.10 -00.11.12 -00.12.22 +01.10.12 +02.10.22 +11.12.22 -12.12.21",
    "+00.00.12 -00.02.10 -00.11.12 -00.12.22 +02.10.11 +02.12.20 +11.12.22 -12.12.21",
    "+00.01.11 -00.01.22 -01.01.10 +01.02.20 -01.11.22 +01.22.22 +02.11.21 -02.21.22",
    "+00.01.11 -00.01.22 +00.02.21 -01.01.10 -01.11.22 +01.12.21 +01.22.22 -02.21.22",
    "+00.01.12 -00.02.11 +00.02.22 -01.11.12 -02.02.20 +02.11.11 -02.11.22 +02.12.21",
    "+00.02.11 -00.02.22 -01.02.10 +01.11.12 -01.12.22 +02.02.20 -02.11.11 +02.11.22",
    "+00.00.11 -00.00.22 -00.01.10 +00.02.20 -00.11.11 +00.22.22 +01.10.11 -02.20.22 +11.11.22 -11.12.21 -11.22.22 +12.21.22",
]
_DD3 = [9, 6, 6, 6, 8, 8, 8, 2, 2, 2, 2, 2, 2, 1]


def _poly(A, spec, transpose):
    tot = 0
    for mono in spec.split():
        sgn = -1.0 if mono[0] == "-" else 1.0
        term = sgn
        for f in mono[1:].split("."):
            i, j = int(f[0]), int(f[1])
            term = term * (A[..., j, i] if transpose else A[..., i, j])
        tot = tot + term
    return tot


def eigenstate_torch(A):
    """Eigenvalues λ_k [..] and eigenprojectors E_k = (∂λ_k/∂A)ᵀ of a batch of d×d tensors,
    by the reference's regularised closed forms; differentiable (create_graph=True)."""
    import torch
    d = A.shape[-1]
    tr = A[..., 0, 0] + A[..., 1, 1] + (A[..., 2, 2] if d == 3 else 0.0)
    if d == 2:
        Dl = (A[..., 0, 0] - A[..., 1, 1]) ** 2 + 4 * A[..., 0, 1] * A[..., 1, 0] + EPS**2
        s = torch.sqrt(Dl)
        lams = [(tr - s) / 2, (tr + s) / 2]
    else:
        AA = A @ A
        I2 = (tr**2 - (AA[..., 0, 0] + AA[..., 1, 1] + AA[..., 2, 2])) / 2
        I3 = torch.linalg.det(A) if False else (
            A[..., 0, 0] * (A[..., 1, 1] * A[..., 2, 2] - A[..., 1, 2] * A[..., 2, 1])
            - A[..., 0, 1] * (A[..., 1, 0] * A[..., 2, 2] - A[..., 1, 2] * A[..., 2, 0])
            + A[..., 0, 2] * (A[..., 1, 0] * A[..., 2, 1] - A[..., 1, 1] * A[..., 2, 0]))
        dq = 2 * tr**3 - 9 * tr * I2 + 27 * I3
        Dl = 0
        for spec, w in zip(_DX3, _DD3):
            Dl = Dl + _poly(A, spec, False) * w * _poly(A, spec, True)
        xp = [A[..., 1, 0], A[..., 2, 0], A[..., 2, 1], -A[..., 0, 0] + A[..., 1, 1],
              -A[..., 0, 0] + A[..., 2, 2], -A[..., 1, 1] + A[..., 2, 2]]
        yp = [A[..., 0, 1], A[..., 0, 2], A[..., 1, 2], xp[3], xp[4], xp[5]]
        dp = 0
        for x_, w, y_ in zip(xp, [6, 6, 6, 1, 1, 1], yp):
            dp = dp + 0.5 * x_ * w * y_
        dp = dp + EPS**2
        Dl = Dl + EPS**2
        phi3 = torch.atan2(np.sqrt(27.0) * torch.sqrt(Dl), dq)
        lams = [(tr + 2 * torch.sqrt(dp) * torch.cos((phi3 + 2 * np.pi * k) / 3)) / 3 for k in (1, 2, 3)]
    Es = [torch.autograd.grad(l.sum(), A, create_graph=True)[0].transpose(-1, -2) for l in lams]
    return lams, Es


def split_torch(A_np, lam, mu, degradation, split_energy, want_tangent=True):
    """Same contract as `split_numpy`, every tangent by autograd of the stress formula."""
    import torch
    A = torch.tensor(np.ascontiguousarray(A_np), dtype=torch.float64, requires_grad=True)
    d = A.shape[-1]
    I = torch.eye(d, dtype=torch.float64)
    tr = A[..., 0, 0] + A[..., 1, 1] + (A[..., 2, 2] if d == 3 else 0.0)
    mp = lambda x: 0.5 * (x + torch.abs(x))
    mn = lambda x: 0.5 * (x - torch.abs(x))
    psi = 0.5 * lam * tr**2 + mu * (A * A).sum((-1, -2))
    sig = lam * tr[..., None, None] * I + 2 * mu * A
    if degradation == "isotropic":
        psi_a, psi_b, sig_a = psi, psi - psi, sig
    elif split_energy == "deviatoric":
        k0 = lam + 2.0 / d * mu
        dev = A - tr[..., None, None] / d * I
        psi_a = 0.5 * k0 * mp(tr) ** 2 + mu * (dev * dev).sum((-1, -2))
        psi_b = 0.5 * k0 * mn(tr) ** 2
        sig_a = k0 * mp(tr)[..., None, None] * I + 2 * mu * dev
    elif split_energy == "spectral":
        lams, Es = eigenstate_torch(A)
        eP = sum(mp(l)[..., None, None] * E for l, E in zip(lams, Es))
        eN = sum(mn(l)[..., None, None] * E for l, E in zip(lams, Es))
        psi_a = 0.5 * lam * mp(tr) ** 2 + mu * (eP * eP).sum((-1, -2))
        psi_b = 0.5 * lam * mn(tr) ** 2 + mu * (eN * eN).sum((-1, -2))
        sig_a = lam * mp(tr)[..., None, None] * I + 2 * mu * eP
    else:
        raise ValueError("Invalid")
    if degradation == "hybrid":
        sig_a = sig
    sig_b = sig - sig_a

    def tangent(S):
        C = torch.zeros(A.shape[:-2] + (d, d, d, d), dtype=torch.float64)
        for i in range(d):
            for j in range(d):
                g, = torch.autograd.grad(S[..., i, j].sum(), A, retain_graph=True, allow_unused=True)
                if g is not None:
                    C[..., i, j, :, :] = g
        return C.numpy()

    out = dict(psi_a=psi_a.detach().numpy(), psi_b=psi_b.detach().numpy(),
               sig_a=sig_a.detach().numpy(), sig_b=sig_b.detach().numpy())
    if want_tangent:
        out["C_a"] = tangent(sig_a)
        # σ_b = σ − σ_a with σ linear in A, so its tangent is the constant elastic one minus C_a
        Ce = lam * torch.einsum("ij,kl->ijkl", I, I) + 2 * mu * torch.einsum("ik,jl->ijkl", I, I)
        out["C_b"] = Ce.numpy() - out["C_a"] if degradation != "isotropic" else np.zeros_like(out["C_a"])
    return out


# --------------------------------------------------------------------------- scalar helpers
def g_quadratic(phi):
    """reference g_degradation_functions.py:12-35"""
    return (1.0 - phi) ** 2.0


def dg_quadratic(phi):
    return -2.0 * (1.0 - phi)


def g_borden(phi):
    """reference g_degradation_functions.py:38-50 AS CODED (s = 0, although the docs say s = 1)"""
    t = 1.0 - phi
    return 3.0 * t * t - 2.0 * t * t * t


def dg_borden(phi):
    """reference g_degradation_functions.py:53-64"""
    t = 1.0 - phi
    return -6.0 * t + 6.0 * t * t


def ddg_borden(phi):
    """d/dphi of dg_borden (what ufl.derivative(F_phi, phi) differentiates, solver.py:224)"""
    return 6.0 - 12.0 * (1.0 - phi)


def degradation(name):
    """(g, dg, ddg) of a degradation function; sargado / alessi are not restated."""
    if name == "quadratic":
        return g_quadratic, dg_quadratic, (lambda phi: 2.0 + 0.0 * phi)
    if name == "borden":
        return g_borden, dg_borden, ddg_borden
    raise NotImplementedError(f"degradation_function {name!r}")


def fatigue_asymptotic(alpha_bar, alpha_T):
    """reference fatigue_degradation_functions.py:19-38"""
    aux = (2.0 * alpha_T / (alpha_bar + alpha_T)) ** 2
    return np.where(alpha_bar >= alpha_T, aux, 1.0)
