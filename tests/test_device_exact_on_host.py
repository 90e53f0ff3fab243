"""CPU-only: the DEVICE reference-order source (csrc/pfc_exact.cuh -- what bristle-friction instructions run on the GPU) compiled for
the host with g++ (-DPFC_HOST_CHECK, -ffp-contract=off, CUDA keywords shimmed away) and compared with the oracle BIT FOR BIT:

  * per candidate pair (random triangle-tetrahedron and tetrahedron-tetrahedron pairs, random poses and twists, both quadrature
    rules): the traction points -- normal, position, dA, pressure -- of pair_points_tri_tet / pair_points_tet_tet against the
    oracle's integrate_over_tri_tet / integrate_over_tet_tet (/root/reference/src/contact_algorithms_non_friction.jl:164-265),
    in Float64 and in Jacobian mode (values and all six partials);
  * per patch: several pairs form one TractionCache list; the three sequential passes (terms_cop -> terms_stiffness ->
    bristle_after_stiffness -> terms_friction -> bristle_finish) against the oracle's yes_contact!(::Bristle)
    (src/contact_algorithms_friction.jl:119-201): wrench and s-dot, Float64 and Dual, every bit equal.

Bitwise equality is the point: K^(-1/2) of the bristle model amplifies rounding-level differences of K by up to 1e8
(pfc_exact.cuh header), so the GPU's bristle results match the reference to 1e-9 only if K is reproduced exactly."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r"""
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>
#include "pfc_oracle.hpp"
#define __host__
#define __device__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __shfl_xor_sync(unsigned, double v, int, int = 32) { return v; }
static inline double __longlong_as_double(long long x) { double d; std::memcpy(&d, &x, 8); return d; }
#define PFC_HOST_CHECK 1
#include "pfc_exact.cuh"

using orc::V3; using orc::V4; using orc::M4;
namespace ex = pfc::ex;
typedef orc::Dual<6> OD;
typedef ex::XD<6> XD6;

static bool biteq(double a, double b) { return std::memcmp(&a, &b, 8) == 0 || (a == 0.0 && b == 0.0); }   // (+0 == -0: sign of zero never reaches a result)
static bool biteq(const XD6& a, const OD& b) { if (!biteq(a.v, b.v)) return false; for (int k = 0; k < 6; ++k) if (!biteq(a.p[k], b.p[k])) return false; return true; }

template <class T> struct Rec { ex::X3<T> n; ex::ExPoint<T> pt; };

static void fill_tet(const orc::Mesh& m, int first, pfc::TetRec& r, double* eps4) {
    V3<double> v[4]; V4<double> e4;
    for (int k = 0; k < 4; ++k) { v[k] = m.point[m.idx[first + k]]; e4[k] = m.eps[m.idx[first + k]]; eps4[k] = e4[k]; for (int i = 0; i < 3; ++i) r.v[3 * k + i] = v[k][i]; }
    const M4<double> inv = orc::inv44(orc::asMatOnePad4(v));
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.inv[4 * i + j] = inv(i, j);
    const V4<double> er = orc::mulrow4<double, double, double>(e4, inv);
    for (int j = 0; j < 4; ++j) r.eps_r[j] = er[j];
}

int main(int argc, char** argv) {
    const long n_case = argc > 1 ? atol(argv[1]) : 20000;
    std::mt19937_64 g(777);
    std::uniform_real_distribution<double> u(-1.0, 1.0), u01(0.0, 1.0);
    long bad_pts = 0, bad_pts_dual = 0, bad_patch = 0, bad_patch_dual = 0, n_patch = 0, n_points = 0, n_tet_tet = 0;
    for (long t = 0; t < n_case; ++t) {
        const int kind1 = (t % 3 == 0) ? 1 : 0;
        const int n_quad_rule = (t % 2) ? 2 : 1;
        const int n_prim = 1 + int(t % 5);   // a patch of up to 5 pairs: mesh 1 has n_prim primitives, mesh 2 one tetrahedron
        orc::Mesh m1, m2;
        m1.kind = kind1; m2.kind = 1;
        const int nv1 = kind1 ? 4 : 3;
        for (int pr = 0; pr < n_prim; ++pr) {
            for (;;) {
                std::vector<V3<double>> pts;
                for (int k = 0; k < nv1; ++k) pts.push_back(orc::mk3<double>(u(g), u(g), u(g)));
                if (kind1) {
                    double vol = orc::tet_volume(pts[0], pts[1], pts[2], pts[3]);
                    if (std::fabs(vol) < 0.02) continue;
                    if (vol < 0) std::swap(pts[0], pts[1]);
                }
                for (int k = 0; k < nv1; ++k) { m1.idx.push_back((int)m1.point.size()); m1.point.push_back(pts[k]); m1.eps.push_back(kind1 ? u01(g) : 0.0); }
                break;
            }
        }
        for (;;) {
            m2.point.clear(); m2.idx.clear(); m2.eps.clear();
            for (int k = 0; k < 4; ++k) { m2.point.push_back(orc::mk3<double>(u(g), u(g), u(g))); m2.idx.push_back(k); m2.eps.push_back(u01(g)); }
            double vol = orc::tet_volume(m2.point[0], m2.point[1], m2.point[2], m2.point[3]);
            if (std::fabs(vol) < 0.02) continue;
            if (vol < 0) std::swap(m2.point[0], m2.point[1]);
            break;
        }
        m1.Ebar = kind1 ? 0.5 + u01(g) : 0.0; m2.Ebar = 0.5 + u01(g);
        double q[4] = {u(g), u(g), u(g), u(g)};
        const double qn = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        for (double& x : q) x /= qn;
        const double R[9] = {1 - 2 * (q[2] * q[2] + q[3] * q[3]), 2 * (q[1] * q[2] - q[0] * q[3]), 2 * (q[1] * q[3] + q[0] * q[2]),
                             2 * (q[1] * q[2] + q[0] * q[3]), 1 - 2 * (q[1] * q[1] + q[3] * q[3]), 2 * (q[2] * q[3] - q[0] * q[1]),
                             2 * (q[1] * q[3] - q[0] * q[2]), 2 * (q[2] * q[3] + q[0] * q[1]), 1 - 2 * (q[1] * q[1] + q[2] * q[2])};
        const double tr[3] = {0.3 * u(g), 0.3 * u(g), 0.3 * u(g)};
        const double tw[6] = {u(g), u(g), u(g), u(g), u(g), u(g)};
        const double chi = 0.3 * u01(g);
        // ---- oracle, Float64 and Dual
        orc::BodyBodyCache<double> b;
        orc::BodyBodyCache<OD> bd;
        b.quad = bd.quad = orc::getTriQuadRule(n_quad_rule);
        b.mesh_1 = bd.mesh_1 = &m1; b.mesh_2 = bd.mesh_2 = &m2;
        ex::ExCtx<double> cx;
        ex::ExCtx<XD6> cd;
        auto seeded = [&](double v, OD& o, XD6& d) { o = OD(v); d = XD6(v); for (int k = 0; k < 6; ++k) { const double s = u(g); o.p[k] = s; d.p[k] = s; } };
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) { b.x_r2_r1(i, j) = R[3 * i + j]; cx.X21[4 * j + i] = R[3 * i + j]; seeded(R[3 * i + j], bd.x_r2_r1(i, j), cd.X21[4 * j + i]); }
            b.x_r2_r1(i, 3) = tr[i]; cx.X21[12 + i] = tr[i]; seeded(tr[i], bd.x_r2_r1(i, 3), cd.X21[12 + i]);
            b.x_r2_r1(3, i) = 0.0; cx.X21[4 * i + 3] = 0.0; bd.x_r2_r1(3, i) = OD(0.0); cd.X21[4 * i + 3] = XD6(0.0);
        }
        b.x_r2_r1(3, 3) = 1.0; cx.X21[15] = 1.0; bd.x_r2_r1(3, 3) = OD(1.0); cd.X21[15] = XD6(1.0);
        b.x_r1_r2 = orc::inv_transform(b.x_r2_r1); bd.x_r1_r2 = orc::inv_transform(bd.x_r2_r1);
        ex::make_x12(cx); ex::make_x12(cd);
        for (int i = 0; i < 6; ++i) { b.twist_r2_r1_r2[i] = tw[i]; cx.twist[i] = tw[i]; seeded(tw[i], bd.twist_r2_r1_r2[i], cd.twist[i]); }
        b.chi = bd.chi = chi; b.Ebar = bd.Ebar = m2.Ebar;
        cx.chi = cd.chi = chi; cx.Ebar1 = cd.Ebar1 = m1.Ebar; cx.Ebar2 = cd.Ebar2 = m2.Ebar; cx.n_quad = cd.n_quad = b.quad.n;
        pfc::TetRec t2; double eps2[4];
        fill_tet(m2, 0, t2, eps2);
        std::vector<Rec<double>> got;
        std::vector<Rec<XD6>> gotd;
        for (int pr = 0; pr < n_prim; ++pr) {
            if (kind1) { orc::integrate_over_tet_tet(pr, 0, b); orc::integrate_over_tet_tet(pr, 0, bd); }
            else { orc::integrate_over_tri_tet(pr, 0, b); orc::integrate_over_tri_tet(pr, 0, bd); }
            ex::ExPoint<double> pts[ex::kExMaxPoints]; ex::X3<double> n2;
            ex::ExPoint<XD6> ptd[ex::kExMaxPoints]; ex::X3<XD6> n2d;
            int np, npd, fl = 0;
            if (kind1) {
                pfc::TetRec t1; double eps1[4];
                fill_tet(m1, 4 * pr, t1, eps1);
                np = ex::pair_points_tet_tet(t1, eps1, t2, eps2, cx, n2, pts, fl);
                npd = ex::pair_points_tet_tet(t1, eps1, t2, eps2, cd, n2d, ptd, fl);
            } else {
                pfc::TriRec tri;
                for (int k = 0; k < 3; ++k) for (int i = 0; i < 3; ++i) tri.v[3 * k + i] = m1.point[3 * pr + k][i];
                const V3<double> n = orc::triangleNormal(m1.point[3 * pr], m1.point[3 * pr + 1], m1.point[3 * pr + 2]);
                for (int i = 0; i < 3; ++i) tri.n[i] = n[i];
                np = ex::pair_points_tri_tet(tri, t2, cx, n2, pts, fl);
                npd = ex::pair_points_tri_tet(tri, t2, cd, n2d, ptd, fl);
            }
            for (int k = 0; k < np; ++k) got.push_back({n2, pts[k]});
            for (int k = 0; k < npd; ++k) gotd.push_back({n2d, ptd[k]});
        }
        // ---- traction points, bit for bit
        bool ok = got.size() == b.traction.size(), okd = gotd.size() == bd.traction.size();
        for (size_t k = 0; ok && k < got.size(); ++k) {
            for (int i = 0; i < 3; ++i) ok = ok && biteq(got[k].n[i], b.traction[k].n[i]) && biteq(got[k].pt.r[i], b.traction[k].r_cart[i]);
            ok = ok && biteq(got[k].pt.dA, b.traction[k].dA) && biteq(got[k].pt.p, b.traction[k].p);
        }
        for (size_t k = 0; okd && k < gotd.size(); ++k) {
            for (int i = 0; i < 3; ++i) okd = okd && biteq(gotd[k].n[i], bd.traction[k].n[i]) && biteq(gotd[k].pt.r[i], bd.traction[k].r_cart[i]);
            okd = okd && biteq(gotd[k].pt.dA, bd.traction[k].dA) && biteq(gotd[k].pt.p, bd.traction[k].p);
        }
        if (!ok && bad_pts++ < 3) std::printf("points differ in case %ld (kind1 %d): %zu vs %zu\n", t, kind1, got.size(), b.traction.size());
        if (!okd && bad_pts_dual++ < 3) std::printf("Dual points differ in case %ld (kind1 %d): %zu vs %zu\n", t, kind1, gotd.size(), bd.traction.size());
        n_points += (long)b.traction.size();
        if (b.traction.empty() || !ok || !okd) continue;
        n_tet_tet += kind1;
        // ---- the bristle patch: three sequential passes
        const orc::Bristle bf = orc::make_bristle(0, 0.02 + 0.05 * u01(g), 1.0e3 + 1.0e5 * u01(g), 0.3 + 0.4 * u01(g), 0.2, 1.0e-3);
        const ex::BristleP bp = {bf.tau, bf.k_bar, bf.mu_s, bf.mu_d, bf.Ts_mu_s, bf.Ts_mu_d, bf.magic};
        const double smag = (t % 4 == 0) ? 1.0 : 1.0e-4;   // large bristle deformations reach the sliding branch of the friction law
        {
            orc::V6<double> s, sd; double sx[6];
            for (int k = 0; k < 6; ++k) { s[k] = smag * u(g); sx[k] = s[k]; }
            const orc::V6<double> w_ref = orc::yes_contact_bristle(bf, b, s, sd);
            double s10[10] = {0}, s27[27] = {0}, s6[6] = {0}, tt[27];
            for (auto& r : got) { ex::terms_cop(r.n, ex::x3(r.pt.r[0], r.pt.r[1], r.pt.r[2]), r.pt.dA, r.pt.p, tt); for (int k = 0; k < 10; ++k) s10[k] = s10[k] + tt[k]; }
            const ex::X3<double> cop = ex::xdivide(ex::x3(s10[7], s10[8], s10[9]), s10[6]);
            for (auto& r : got) { ex::terms_stiffness(r.n, ex::x3(r.pt.r[0], r.pt.r[1], r.pt.r[2]), r.pt.dA, r.pt.p, cop, tt);
                for (int k = 0; k < 18; ++k) s27[k] = s27[k] + tt[k]; for (int k = 18; k < 27; ++k) s27[k] = s27[k] - tt[k]; }
            double Sinv[6], Kh[6][6], D2[6], w[6], sdot[6];
            ex::PatchScratch<double> scr;
            ex::bristle_after_stiffness(s27, bp, sx, Sinv, Kh, D2, scr, ex::Coop());
            for (auto& r : got) { ex::terms_friction(bp, r.n, ex::x3(r.pt.r[0], r.pt.r[1], r.pt.r[2]), r.pt.dA, r.pt.p, cop, D2, cx.twist, tt); for (int k = 0; k < 6; ++k) s6[k] = s6[k] + tt[k]; }
            ex::bristle_finish(s10, s6, cop, bp, sx, Sinv, Kh, w, sdot);
            bool okp = true;
            for (int k = 0; k < 6; ++k) okp = okp && biteq(w[k], w_ref[k]) && biteq(sdot[k], sd[k]);
            if (!okp && bad_patch++ < 3) { std::printf("bristle patch differs in case %ld:", t); for (int k = 0; k < 6; ++k) std::printf(" %.17g/%.17g", w[k], w_ref[k]); std::printf("\n"); }
        }
        {
            orc::V6<OD> s, sd; XD6 sx[6];
            for (int k = 0; k < 6; ++k) seeded(smag * u(g), s[k], sx[k]);
            const orc::V6<OD> w_ref = orc::yes_contact_bristle(bf, bd, s, sd);
            XD6 s10[10], s27[27], s6[6], tt[27];
            for (auto& x : s10) x = XD6(0.0); for (auto& x : s27) x = XD6(0.0); for (auto& x : s6) x = XD6(0.0);
            for (auto& r : gotd) { ex::terms_cop(r.n, ex::x3(r.pt.r[0], r.pt.r[1], r.pt.r[2]), r.pt.dA, r.pt.p, tt); for (int k = 0; k < 10; ++k) s10[k] = s10[k] + tt[k]; }
            const ex::X3<XD6> cop = ex::xdivide(ex::x3(s10[7], s10[8], s10[9]), s10[6]);
            for (auto& r : gotd) { ex::terms_stiffness(r.n, ex::x3(r.pt.r[0], r.pt.r[1], r.pt.r[2]), r.pt.dA, r.pt.p, cop, tt);
                for (int k = 0; k < 18; ++k) s27[k] = s27[k] + tt[k]; for (int k = 18; k < 27; ++k) s27[k] = s27[k] - tt[k]; }
            XD6 Sinv[6], Kh[6][6], D2[6], w[6], sdot[6];
            ex::PatchScratch<XD6> scr;
            ex::bristle_after_stiffness(s27, bp, sx, Sinv, Kh, D2, scr, ex::Coop());
            for (auto& r : gotd) { ex::terms_friction(bp, r.n, ex::x3(r.pt.r[0], r.pt.r[1], r.pt.r[2]), r.pt.dA, r.pt.p, cop, D2, cd.twist, tt); for (int k = 0; k < 6; ++k) s6[k] = s6[k] + tt[k]; }
            ex::bristle_finish(s10, s6, cop, bp, sx, Sinv, Kh, w, sdot);
            bool okp = true;
            for (int k = 0; k < 6; ++k) okp = okp && biteq(w[k], w_ref[k]) && biteq(sdot[k], sd[k]);
            if (!okp && bad_patch_dual++ < 3) std::printf("Dual bristle patch differs in case %ld\n", t);
        }
        ++n_patch;
    }
    std::printf("cases %ld bad_pts %ld bad_pts_dual %ld bad_patch %ld bad_patch_dual %ld patches %ld points %ld tet_tet_patches %ld\n", n_case, bad_pts, bad_pts_dual,
                bad_patch, bad_patch_dual, n_patch, n_points, n_tet_tet);
    return (bad_pts || bad_pts_dual || bad_patch || bad_patch_dual) ? 1 : 0;
}
"""


def test_device_reference_order_source_matches_oracle_bit_for_bit(tmp_path):
    cpp = tmp_path / "exact_host.cpp"
    cpp.write_text(HARNESS)
    exe = tmp_path / "exact_host"
    inc = ["-I", os.path.join(ROOT, "oracle"), "-I", os.path.join(ROOT, "pressurefieldcontact.jl_b200", "csrc"), "-I", "/usr/local/cuda/include"]
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-w", *inc, "-o", str(exe), str(cpp)])
    out = subprocess.run([str(exe), "60000"], capture_output=True, text=True)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0, out.stdout[-2000:]
    f = out.stdout.split()
    keys = ("bad_pts", "bad_pts_dual", "bad_patch", "bad_patch_dual", "patches", "points", "tet_tet_patches")
    stats = {f[i]: int(f[i + 1]) for i in range(0, len(f) - 1) if f[i] in keys}
    assert stats["bad_pts"] == 0 and stats["bad_pts_dual"] == 0 and stats["bad_patch"] == 0 and stats["bad_patch_dual"] == 0
    assert stats["patches"] > 3000 and stats["points"] > 20000 and stats["tet_tet_patches"] > 300
