// pfc_twist.cuh -- Jacobian mode of a pair whose transform carries no partials (the seeds sit on velocities): Float64 geometry, the
// twist-dependent part on Dual<N>.  Used by pfc_dual_chunked.cu; compiled for the host by tests/test_device_narrow_on_host.py.
#pragma once
#include "pfc_patch.cuh"

namespace pfc {

// ---- instructions whose TRANSFORM does not depend on the seeds (the seeds sit on velocities): only the twist carries partials ----------
// The contact polygon, its normal, the quadrature points and the elastic pressure are then plain Float64 (one clip per pair instead of
// one per chunk of partials); the damping term, the slip velocity, the friction coefficient and the wrench sums run on Dual<N>.  Same
// expressions, in the same order, as quad_point / Accum::point (pfc_patch.cuh) with the geometric operands demoted to double.
template <int N> struct TwistAcc {
    Dual<N> a[6];
    Vec3<Dual<N>> w_ang, w_lin;
    const double* fp;
    int n_points;
};

template <int N> PFC_D void quad_point_tw(const Vec3<double>& v1, const Vec3<double>& v2, const Vec3<double>& cen, const Vec3<double>& n, double g0, double g1,
                                          double g2, double g3, double za, double zb, double zc, double dA, double chi, double Ebar2, TwistAcc<N>& acc) {
    typedef Dual<N> D;
    const Vec3<double> r = mk<double>(v1.x * za + v2.x * zb + cen.x * zc, v1.y * za + v2.y * zb + cen.y * zc, v1.z * za + v2.z * zb + cen.z * zc);
    double eps = fma(g0, r.x, g3);
    eps = fma(g1, r.y, eps);
    eps = fma(g2, r.z, eps);
    const D vx = acc.w_lin.x + (acc.w_ang.y * r.z - acc.w_ang.z * r.y);   // w_lin + w_ang x r
    const D vy = acc.w_lin.y + (acc.w_ang.z * r.x - acc.w_ang.x * r.z);
    const D vz = acc.w_lin.z + (acc.w_ang.x * r.y - acc.w_ang.y * r.x);
    const D ee = -(g0 * vx + g1 * vy + g2 * vz);
    D damp = 1.0 + chi * ee;
    if (damp.v < 0.0) damp = D(0.0);
    const D p = (eps * Ebar2) * damp;
    if (!(0.0 < p.v)) return;
    const D p_dA = p * dA;
    const D t = -(vx * n.x + vy * n.y + vz * n.z);
    const D tx = t * n.x + vx, ty = t * n.y + vy, tz = t * n.z + vz;   // slip velocity in the contact plane
    const D mag2 = tx * tx + ty * ty + tz * tz;
    const double* fp = acc.fp;
    const double mu_s = fp[0], v_c = fp[2];
    D coef;
    if (mag2.v < v_c * v_c) {
        coef = (-mu_s * fp[6]) * p_dA;
    } else {
        const D mag = sqrt_(mag2);
        coef = (clamped_piecewise(mag, fp[3], fp[5], mu_s, fp[1]) / mag) * (-p_dA);
    }
    const D kx = p_dA * n.x + tx * coef, ky = p_dA * n.y + ty * coef, kz = p_dA * n.z + tz * coef;
    acc.a[0] += r.y * kz - r.z * ky; acc.a[1] += r.z * kx - r.x * kz; acc.a[2] += r.x * ky - r.y * kx;
    acc.a[3] += kx; acc.a[4] += ky; acc.a[5] += kz;
    ++acc.n_points;
}

template <int N> PFC_D void integrate_subtri_tw(const Vec3<double>& v1, const Vec3<double>& v2, const Vec3<double>& cen, const Vec3<double>& nrm,
                                                const double* eps_r, double chi, double Ebar2, int n_quad, TwistAcc<N>& acc) {
    const double area = dot(nrm, cross(v2 - v1, cen - v2) * 0.5);
    if (!(0.0 < area)) return;
    const double g0 = eps_r[0], g1 = eps_r[1], g2 = eps_r[2], g3 = eps_r[3];
    if (n_quad == 1) { quad_point_tw(v1, v2, cen, nrm, g0, g1, g2, g3, PFC_Q1, PFC_Q1, PFC_Q1, area * 1.0, chi, Ebar2, acc); return; }
    const double dA = area * PFC_Q1;
    quad_point_tw(v1, v2, cen, nrm, g0, g1, g2, g3, PFC_QA, PFC_QB, PFC_QA, dA, chi, Ebar2, acc);
    quad_point_tw(v1, v2, cen, nrm, g0, g1, g2, g3, PFC_QB, PFC_QA, PFC_QA, dA, chi, Ebar2, acc);
    quad_point_tw(v1, v2, cen, nrm, g0, g1, g2, g3, PFC_QA, PFC_QA, PFC_QB, dA, chi, Ebar2, acc);
}

}  // namespace pfc
