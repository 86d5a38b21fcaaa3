// pfx_physics.cuh — pointwise constitutive layer, host+device, fp64, registers only.
//
// Restates (does not copy) the reference's symbolic UFL layer as closed-form scalar code:
//   isotropic law              pfx/Materials/elastic_isotropic.py:26-70
//   Macaulay brackets          pfx/Math/functions.py:10-41          <x>± = ½(x ± |x|)
//   2×2 / 3×3 eigen-states     pfx/Math/invariants.py:85-204        (regulariser eps = 3e-16)
//   spectral ± parts           pfx/Math/tensor_decomposition.py:62-132
//   spectral / vol-dev splits  pfx/Element/Phase_Field_Fracture/split_energy_stress_tangent_functions.py:34-231
//   dispatch incl. hybrid      same file :234-354
// The consistent tangent is UFL's derivative of the stress formula (solver.py:186); here it
// is obtained by evaluating the same formula on forward-mode dual numbers (d|x|/dx = sign x).
//
// Symmetric tensors are stored as S2 = (xx,yy,xy) / S3 = (xx,yy,zz,xy,xz,yz).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define PFX_HD __host__ __device__ __forceinline__
#else
#define PFX_HD inline
#endif

namespace pfx {

constexpr double kEps = 3.0e-16;   // pfx/Math/invariants.py:96,185

enum Model { M_ISO = 0, M_SPECTRAL = 1, M_VOLDEV = 2 };

struct Material {
    double lam, mu, Gc, l, k;
    int smodel;   // split used by the stress  (hybrid => M_ISO)
    int emodel;   // split used by psi_a/psi_b
    int fatigue;
    double alpha_T;
    int dfun;     // degradation function: 0 quadratic (1-phi)^2, 1 borden 3t^2 - 2t^3 with t = 1-phi
};

// degradation function g, dg/dphi, d2g/dphi2 (reference g_degradation_functions.py:12-64 as coded: borden with
// s = 0)
PFX_HD double g_fun(int dfun, double phi) {
    const double t = 1.0 - phi;
    return dfun == 1 ? t * t * (3.0 - 2.0 * t) : t * t;
}
PFX_HD double dg_fun(int dfun, double phi) {
    const double t = 1.0 - phi;
    return dfun == 1 ? 6.0 * t * (t - 1.0) : -2.0 * t;
}
PFX_HD double ddg_fun(int dfun, double phi) {
    return dfun == 1 ? 6.0 - 12.0 * (1.0 - phi) : 2.0;
}

// ----------------------------------------------------------------------------- dual numbers
struct Dual {
    double v, d;
    PFX_HD Dual() : v(0), d(0) {}
    PFX_HD Dual(double v_) : v(v_), d(0) {}
    PFX_HD Dual(double v_, double d_) : v(v_), d(d_) {}
};
PFX_HD Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
PFX_HD Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
PFX_HD Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
PFX_HD Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, a.d * b.v + a.v * b.d); }
PFX_HD Dual operator/(Dual a, Dual b) { double q = a.v / b.v; return Dual(q, (a.d - q * b.d) / b.v); }
PFX_HD Dual operator+(Dual a, double b) { return Dual(a.v + b, a.d); }
PFX_HD Dual operator+(double a, Dual b) { return Dual(a + b.v, b.d); }
PFX_HD Dual operator-(Dual a, double b) { return Dual(a.v - b, a.d); }
PFX_HD Dual operator-(double a, Dual b) { return Dual(a - b.v, -b.d); }
PFX_HD Dual operator*(Dual a, double b) { return Dual(a.v * b, a.d * b); }
PFX_HD Dual operator*(double a, Dual b) { return Dual(a * b.v, a * b.d); }
PFX_HD Dual operator/(Dual a, double b) { return Dual(a.v / b, a.d / b); }
PFX_HD Dual operator/(double a, Dual b) { double q = a / b.v; return Dual(q, -q * b.d / b.v); }
PFX_HD Dual m_sqrt(Dual a) { double s = ::sqrt(a.v); return Dual(s, 0.5 * a.d / s); }
PFX_HD Dual m_cos(Dual a) { return Dual(::cos(a.v), -::sin(a.v) * a.d); }
PFX_HD Dual m_sin(Dual a) { return Dual(::sin(a.v), ::cos(a.v) * a.d); }
PFX_HD Dual m_atan2(Dual y, Dual x) {
    double r2 = x.v * x.v + y.v * y.v;
    return Dual(::atan2(y.v, x.v), (x.v * y.d - y.v * x.d) / r2);
}
PFX_HD double sgn(double x) { return (x > 0.0) - (x < 0.0); }
// double overloads of the m_* math wrappers used by the templated formulas
PFX_HD double m_sqrt(double a) { return ::sqrt(a); }
PFX_HD double m_cos(double a) { return ::cos(a); }
PFX_HD double m_sin(double a) { return ::sin(a); }
PFX_HD double m_atan2(double y, double x) { return ::atan2(y, x); }
PFX_HD double m_abs(double a) { return ::fabs(a); }
PFX_HD Dual m_abs(Dual a) { return Dual(::fabs(a.v), sgn(a.v) * a.d); }
PFX_HD double val(double a) { return a; }
PFX_HD double val(Dual a) { return a.v; }

template <class T> PFX_HD T mac_p(T x) { return 0.5 * (x + m_abs(x)); }
template <class T> PFX_HD T mac_n(T x) { return 0.5 * (x - m_abs(x)); }
template <class T> PFX_HD T sq(T x) { return x * x; }

// ----------------------------------------------------------------------------- 2×2 spectral
// eigenstate2 (invariants.py:174-204) for symmetric A = (a, c, b): E∓ = ½(I ∓ M/s),
// M = [[a-c, 2b],[2b, c-a]], s = sqrt((a-c)² + 4b² + eps²); returns ε⁺ and ε⁻.
template <class T>
PFX_HD void spectral_parts2(const T e[3], T eP[3], T eN[3]) {
    T a = e[0], c = e[1], b = e[2];
    T tr = a + c;
    T amc = a - c;
    T s = m_sqrt(amc * amc + 4.0 * (b * b) + kEps * kEps);
    T l0 = 0.5 * (tr - s), l1 = 0.5 * (tr + s);
    T m00 = amc / s, m01 = (2.0 * b) / s;            // M/s (m11 = -m00)
    T E0[3] = {0.5 * (1.0 - m00), 0.5 * (1.0 + m00), -0.5 * m01};
    T E1[3] = {0.5 * (1.0 + m00), 0.5 * (1.0 - m00), 0.5 * m01};
    T p0 = mac_p(l0), p1 = mac_p(l1), n0 = mac_n(l0), n1 = mac_n(l1);
    for (int i = 0; i < 3; ++i) {
        eP[i] = p0 * E0[i] + p1 * E1[i];
        eN[i] = n0 * E0[i] + n1 * E1[i];
    }
}

#include "pfx_eig3_gen.cuh"   // generated: disc3_and_grad<T>(A[9], Dl, G[9]) (invariants.py:105-144)

// eigenstate3 (invariants.py:85-171): λ_k = (I1 + 2 sqrt(dp) cos((φ3 + 2πk)/3))/3, k = 1..3,
// E_k = (∂λ_k/∂A)ᵀ by the chain rule; symmetric input A (sym6), symmetric output.
template <class T>
PFX_HD void spectral_parts3(const T e[6], T eP[6], T eN[6]) {
    // full matrix, row-major
    T A[9] = {e[0], e[3], e[4], e[3], e[1], e[5], e[4], e[5], e[2]};
    T I1 = A[0] + A[4] + A[8];
    T trAA = A[0] * A[0] + A[4] * A[4] + A[8] * A[8] + 2.0 * (A[1] * A[3] + A[2] * A[6] + A[5] * A[7]);
    T I2 = (I1 * I1 - trAA) * 0.5;
    // cofactors = ∂det/∂A_kl
    T cof[9] = {A[4] * A[8] - A[5] * A[7], A[5] * A[6] - A[3] * A[8], A[3] * A[7] - A[4] * A[6],
                A[2] * A[7] - A[1] * A[8], A[0] * A[8] - A[2] * A[6], A[1] * A[6] - A[0] * A[7],
                A[1] * A[5] - A[2] * A[4], A[2] * A[3] - A[0] * A[5], A[0] * A[4] - A[1] * A[3]};
    T I3 = A[0] * cof[0] + A[1] * cof[1] + A[2] * cof[2];
    T dq = 2.0 * (I1 * I1 * I1) - 9.0 * (I1 * I2) + 27.0 * I3;
    T Dl, GD[9];
    disc3_and_grad<T>(A, Dl, GD);
    T d10 = A[4] - A[0], d20 = A[8] - A[0], d21 = A[8] - A[4];
    T dp = 3.0 * (A[3] * A[1] + A[6] * A[2] + A[7] * A[5]) + 0.5 * (d10 * d10 + d20 * d20 + d21 * d21);
    dp = dp + kEps * kEps;
    Dl = Dl + kEps * kEps;
    const double s27 = 5.196152422706632;   // sqrt(27)
    T sD = m_sqrt(Dl);
    T y = s27 * sD;
    T phi3 = m_atan2(y, dq);
    T sdp = m_sqrt(dp);
    // gradients wrt A_kl (row-major kl)
    T gI2[9], gdq[9], gdp[9], gph[9];
    T r2 = dq * dq + y * y;
    T cy = (s27 * 0.5) / sD;
    T cI1 = 6.0 * (I1 * I1) - 9.0 * I2;
    for (int k = 0; k < 3; ++k)
        for (int l = 0; l < 3; ++l) {
            T dI1 = (k == l) ? T(1.0) : T(0.0);
            gI2[3 * k + l] = I1 * dI1 - A[3 * l + k];
            gdq[3 * k + l] = cI1 * dI1 - 9.0 * (I1 * gI2[3 * k + l]) + 27.0 * cof[3 * k + l];
        }
    gdp[0] = -(d10 + d20); gdp[4] = d10 - d21; gdp[8] = d20 + d21;
    gdp[1] = 3.0 * A[3]; gdp[3] = 3.0 * A[1];
    gdp[2] = 3.0 * A[6]; gdp[6] = 3.0 * A[2];
    gdp[5] = 3.0 * A[7]; gdp[7] = 3.0 * A[5];
    for (int i = 0; i < 9; ++i) gph[i] = (dq * (cy * GD[i]) - y * gdq[i]) / r2;
    for (int i = 0; i < 6; ++i) { eP[i] = T(0.0); eN[i] = T(0.0); }
    const int sym_idx[6] = {0, 4, 8, 1, 2, 5};      // (k,l) of xx,yy,zz,xy,xz,yz
    const int sym_tr[6] = {0, 4, 8, 3, 6, 7};       // transposed partner (E = grad^T)
    for (int k = 1; k <= 3; ++k) {
        T th = (phi3 + (2.0 * 3.14159265358979323846 * k)) / 3.0;
        T ck = m_cos(th), sk = m_sin(th);
        T lk = (I1 + 2.0 * (sdp * ck)) / 3.0;
        T c1 = ck / sdp;
        T c2 = (2.0 / 3.0) * (sdp * sk);
        T pk = mac_p(lk), nk = mac_n(lk);
        for (int i = 0; i < 6; ++i) {
            int t = sym_tr[i];            // E_ij = ∂λ/∂A_ji
            T dI1 = (i < 3) ? T(1.0) : T(0.0);
            T Ek = (dI1 + c1 * gdp[t] - c2 * gph[t]) / 3.0;
            eP[i] = eP[i] + pk * Ek;
            eN[i] = eN[i] + nk * Ek;
            (void)sym_idx;
        }
    }
}

// ----------------------------------------------------------------------------- stress / energy
template <int D> struct SymN { static constexpr int N = (D == 2) ? 3 : 6; };

template <int D, class T> PFX_HD T trace(const T* e) { if (D == 2) return e[0] + e[1]; else return e[0] + e[1] + e[2]; }
template <int D, class T> PFX_HD T ddot(const T* a, const T* b) {
    T s = T(0.0);
    for (int i = 0; i < D; ++i) s = s + a[i] * b[i];
    for (int i = D; i < SymN<D>::N; ++i) s = s + 2.0 * (a[i] * b[i]);
    return s;
}
template <int D, class T> PFX_HD void spectral_parts(const T* e, T* eP, T* eN) {
    if (D == 2) spectral_parts2<T>(e, eP, eN); else spectral_parts3<T>(e, eP, eN);
}

// σ_a and σ_b (sigma_a/sigma_b dispatch, split_energy_stress_tangent_functions.py:295-354)
template <int D, class T>
PFX_HD void stress_split(const Material& m, int model, const T* e, T* sa, T* sb) {
    constexpr int N = SymN<D>::N;
    T tr = trace<D, T>(e);
    T s[N];
    for (int i = 0; i < N; ++i) s[i] = (2.0 * m.mu) * e[i];
    for (int i = 0; i < D; ++i) s[i] = s[i] + m.lam * tr;
    if (model == M_ISO) {
        for (int i = 0; i < N; ++i) { sa[i] = s[i]; sb[i] = T(0.0); }
        return;
    }
    if (model == M_VOLDEV) {
        double k0 = m.lam + 2.0 / D * m.mu;
        T trp = mac_p(tr);
        for (int i = 0; i < N; ++i) sa[i] = (2.0 * m.mu) * e[i];
        for (int i = 0; i < D; ++i) sa[i] = sa[i] - (2.0 * m.mu / D) * tr + k0 * trp;
    } else {
        T eP[N], eN[N];
        spectral_parts<D, T>(e, eP, eN);
        T trp = mac_p(tr);
        for (int i = 0; i < N; ++i) sa[i] = (2.0 * m.mu) * eP[i];
        for (int i = 0; i < D; ++i) sa[i] = sa[i] + m.lam * trp;
    }
    for (int i = 0; i < N; ++i) sb[i] = s[i] - sa[i];
}

// ψ_a and ψ_b (psi_a/psi_b dispatch, same file :234-292)
template <int D>
PFX_HD void energy_split(const Material& m, int model, const double* e, double& pa, double& pb) {
    constexpr int N = SymN<D>::N;
    double tr = trace<D, double>(e);
    if (model == M_ISO) {
        pa = 0.5 * m.lam * tr * tr + m.mu * ddot<D, double>(e, e);
        pb = pa - pa;
        return;
    }
    if (model == M_VOLDEV) {
        double k0 = m.lam + 2.0 / D * m.mu;
        double dev[N];
        for (int i = 0; i < N; ++i) dev[i] = e[i];
        for (int i = 0; i < D; ++i) dev[i] -= tr / D;
        pa = 0.5 * k0 * sq(mac_p(tr)) + m.mu * ddot<D, double>(dev, dev);
        pb = 0.5 * k0 * sq(mac_n(tr));
        return;
    }
    double eP[N], eN[N];
    spectral_parts<D, double>(e, eP, eN);
    pa = 0.5 * m.lam * sq(mac_p(tr)) + m.mu * ddot<D, double>(eP, eP);
    pb = 0.5 * m.lam * sq(mac_n(tr)) + m.mu * ddot<D, double>(eN, eN);
}

// Closed-form tangent action of the 2×2 spectral split (hand-derived from eigenstate2; equals
// the dual-number evaluation to round-off, tests/test_elements_host.py): one sqrt, one divide.
PFX_HD void dstress_spectral2(const Material& m, const double* e, const double* de, double* dsa, double* dsb) {
    const double a = e[0], c = e[1], b = e[2], da = de[0], dc = de[1], db = de[2];
    const double amc = a - c, tr = a + c, dtr = da + dc, dd = da - dc;
    const double s = ::sqrt(amc * amc + 4.0 * b * b + kEps * kEps), is = 1.0 / s;
    const double l0 = 0.5 * (tr - s), l1 = 0.5 * (tr + s);
    const double m0 = amc * is, m1 = 2.0 * b * is;
    const double ds = m0 * dd + 2.0 * m1 * db;
    const double h0 = 0.5 * (1.0 + sgn(l0)), h1 = 0.5 * (1.0 + sgn(l1));
    const double w0 = h0 * 0.5 * (dtr - ds), w1 = h1 * 0.5 * (dtr + ds);       // H(λk) dλk
    const double pd = 0.5 * (mac_p(l1) - mac_p(l0)) * is;                      // (p1-p0)/2s
    const double dm0 = dd - m0 * ds, dm1 = 2.0 * db - m1 * ds;
    // E1 = ½(I + M/s), E0 = I − E1
    const double e1x = 0.5 * (1.0 + m0), e1y = 0.5 * (1.0 - m0), e1s = 0.5 * m1;
    const double dPx = w0 * (1.0 - e1x) + w1 * e1x + pd * dm0;
    const double dPy = w0 * (1.0 - e1y) + w1 * e1y - pd * dm0;
    const double dPs = (w1 - w0) * e1s + pd * dm1;
    const double lt = m.lam * 0.5 * (1.0 + sgn(tr)) * dtr, mu2 = 2.0 * m.mu;
    dsa[0] = lt + mu2 * dPx; dsa[1] = lt + mu2 * dPy; dsa[2] = mu2 * dPs;
    const double lf = m.lam * dtr;
    dsb[0] = lf + mu2 * da - dsa[0]; dsb[1] = lf + mu2 * dc - dsa[1]; dsb[2] = mu2 * db - dsa[2];
}

// directional derivative of (σ_a, σ_b) at e along de
template <int D>
PFX_HD void dstress_split(const Material& m, int model, const double* e, const double* de, double* dsa, double* dsb) {
    constexpr int N = SymN<D>::N;
    double tr = trace<D, double>(de);
    if (model == M_ISO) {
        for (int i = 0; i < N; ++i) { dsa[i] = 2.0 * m.mu * de[i]; dsb[i] = 0.0; }
        for (int i = 0; i < D; ++i) dsa[i] += m.lam * tr;
        return;
    }
    if (model == M_VOLDEV) {
        // σ_a = k0<tr>+ I + 2μ dev ε  (split_energy_stress_tangent_functions.py:190-212)
        double k0 = m.lam + 2.0 / D * m.mu;
        double hp = 0.5 * (1.0 + sgn(trace<D, double>(e)));
        for (int i = 0; i < N; ++i) { dsa[i] = 2.0 * m.mu * de[i]; dsb[i] = 0.0; }
        for (int i = 0; i < D; ++i) {
            dsa[i] += (k0 * hp - 2.0 * m.mu / D) * tr;
            dsb[i] = (m.lam + 2.0 * m.mu / D - k0 * hp) * tr;
        }
        return;
    }
    if (D == 2) { dstress_spectral2(m, e, de, dsa, dsb); return; }
    Dual ed[N], sa[N], sb[N];
    for (int i = 0; i < N; ++i) ed[i] = Dual(e[i], de[i]);
    stress_split<D, Dual>(m, model, ed, sa, sb);
    for (int i = 0; i < N; ++i) { dsa[i] = sa[i].d; dsb[i] = sb[i].d; }
}

// reference-faithful dual-number evaluation (used by the tests to pin the closed forms)
template <int D>
PFX_HD void dstress_split_dual(const Material& m, int model, const double* e, const double* de, double* dsa, double* dsb) {
    constexpr int N = SymN<D>::N;
    Dual ed[N], sa[N], sb[N];
    for (int i = 0; i < N; ++i) ed[i] = Dual(e[i], de[i]);
    stress_split<D, Dual>(m, model, ed, sa, sb);
    for (int i = 0; i < N; ++i) { dsa[i] = sa[i].d; dsb[i] = sb[i].d; }
}

}  // namespace pfx
