"""CPU-only: the DEVICE narrow-phase source (csrc/pfc_patch.cuh + pfc_clip.cuh + pfc_math.cuh) compiled for the host with g++
(-DPFC_HOST_CHECK, CUDA keywords shimmed away) and run pair by pair against the oracle's integrate_over_tri_tet /
integrate_over_tet_tet (oracle/pfc_oracle.hpp, following /root/reference/src/contact_algorithms_non_friction.jl:164-265) on
random triangle-tetrahedron and tetrahedron-tetrahedron pairs under random transforms and twists.

Both device routes are run:
  * the tile kernel's route (start_polygon_zeta -> slot -> clip_tet_inplace -> pair_normal -> finish_polygon_slot ->
    integrate_subtri per edge), and
  * the one-thread-per-pair route of the large path / warp kernels (integrate_pair: clip_pair -> integrate_subtri),
each with the accumulator in its traction-dump mode, and compared with the oracle's TractionCache list point by point: same
number of points, normal / position / dA / pressure within 1e-10 (relative to the larger of 1 and the value; the device
hoists per-tet inverses and uses one reciprocal in weightPoly, so the agreement is to rounding, not bitwise).
The regularized-Coulomb wrench of every pair that touches (Accum::point in its regularized mode: both branches of the friction
law) is compared with the oracle's yes_contact!(::Regularized) (src/contact_algorithms_friction.jl:13-30, 50-72) to 1e-10 of the
torque / force magnitude -- and once more in Jacobian mode: pose and twist carry six random partials each, the device templates run on
pfc::Dual<6>, the oracle on its own Dual<6>, and values and all partials of the wrench must agree to 1e-9 of the group's magnitude.
A third variant seeds the twist only (the velocity chunks of a Jacobian): the device's Float64-polygon / Dual-twist route
(pfc_twist.cuh) against the oracle's all-Dual evaluation with zero pose partials, same bar.
This is the GPU parity test's TractionCache check (tests/test_gpu_parity.py) restated where no GPU is needed."""
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
// ---- CUDA keywords and intrinsics the device headers use, for g++
#define __host__
#define __device__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
static inline double __shfl_xor_sync(unsigned, double v, int, int = 32) { return v; }
static inline int __shfl_xor_sync(unsigned, int v, int, int = 32) { return v; }
static inline double __longlong_as_double(long long x) { double d; std::memcpy(&d, &x, 8); return d; }
static inline long long __double_as_longlong(double x) { long long d; std::memcpy(&d, &x, 8); return d; }
#define PFC_HOST_CHECK 1
#include "pfc_patch.cuh"
#include "pfc_twist.cuh"

using orc::V3; using orc::V4; using orc::M4;

static bool close(double a, double b) { return std::fabs(a - b) <= 1.0e-10 * std::fmax(1.0, std::fmax(std::fabs(a), std::fabs(b))); }

struct Points { std::vector<double> d; int n = 0; };   // 8 doubles per point: n(3) r(3) dA p

static bool same(const Points& got, const std::vector<orc::TractionCache<double>>& ref) {
    if (got.n != (int)ref.size()) return false;
    for (int k = 0; k < got.n; ++k) {
        const double* o = &got.d[8 * k];
        for (int i = 0; i < 3; ++i) if (!close(o[i], ref[k].n[i]) || !close(o[3 + i], ref[k].r_cart[i])) return false;
        if (!close(o[6], ref[k].dA) || !close(o[7], ref[k].p)) return false;
    }
    return true;
}

int main(int argc, char** argv) {
    const long n_case = argc > 1 ? atol(argv[1]) : 100000;
    std::mt19937_64 g(4096);
    std::uniform_real_distribution<double> u(-1.0, 1.0), u01(0.0, 1.0);
    long bad_tile = 0, bad_pair = 0, with_points = 0, n_points = 0, tet_tet = 0, bad_wrench = 0, n_wrench = 0, bad_dual = 0, bad_twist = 0;
    for (long t = 0; t < n_case; ++t) {
        const int kind1 = (t % 3 == 0) ? 1 : 0;   // a third of the cases tet-tet
        const int n_quad_rule = (t % 2) ? 2 : 1;
        // ---- the two primitives (mesh frames) and the pose
        orc::Mesh m1, m2;
        m1.kind = kind1; m2.kind = 1;
        const int nv1 = kind1 ? 4 : 3;
        for (int k = 0; k < nv1; ++k) { m1.point.push_back(orc::mk3<double>(u(g), u(g), u(g))); m1.idx.push_back(k); m1.eps.push_back(kind1 ? u01(g) : 0.0); }
        for (int k = 0; k < 4; ++k) { m2.point.push_back(orc::mk3<double>(u(g), u(g), u(g))); m2.idx.push_back(k); m2.eps.push_back(u01(g)); }
        if (std::fabs(orc::tet_volume(m2.point[0], m2.point[1], m2.point[2], m2.point[3])) < 0.02) continue;
        if (kind1 && std::fabs(orc::tet_volume(m1.point[0], m1.point[1], m1.point[2], m1.point[3])) < 0.02) continue;
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
        // ---- oracle
        orc::BodyBodyCache<double> b;
        b.quad = orc::getTriQuadRule(n_quad_rule);
        b.mesh_1 = &m1; b.mesh_2 = &m2;
        for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) b.x_r2_r1(i, j) = R[3 * i + j]; b.x_r2_r1(i, 3) = tr[i]; b.x_r2_r1(3, i) = 0.0; }
        b.x_r2_r1(3, 3) = 1.0;
        b.x_r1_r2 = orc::inv_transform(b.x_r2_r1);
        for (int i = 0; i < 6; ++i) b.twist_r2_r1_r2[i] = tw[i];
        b.chi = chi; b.Ebar = m2.Ebar;
        if (kind1) orc::integrate_over_tet_tet(0, 0, b); else orc::integrate_over_tri_tet(0, 0, b);
        // ---- device records, as pfc_finalize lays them out (csrc/pfc_api.cu)
        pfc::TetRec tets[2];
        pfc::TriRec tri;
        auto fill_tet = [&](const orc::Mesh& m, pfc::TetRec& r) {
            V3<double> v[4]; V4<double> e4;
            for (int k = 0; k < 4; ++k) { v[k] = m.point[k]; e4[k] = m.eps[k]; for (int i = 0; i < 3; ++i) r.v[3 * k + i] = v[k][i]; }
            const M4<double> inv = orc::inv44(orc::asMatOnePad4(v));
            for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.inv[4 * i + j] = inv(i, j);
            const V4<double> er = orc::mulrow4<double, double, double>(e4, inv);
            for (int j = 0; j < 4; ++j) r.eps_r[j] = er[j];
        };
        fill_tet(m2, tets[0]);
        if (kind1) fill_tet(m1, tets[1]);
        else {
            for (int k = 0; k < 3; ++k) for (int i = 0; i < 3; ++i) tri.v[3 * k + i] = m1.point[k][i];
            const V3<double> n = orc::triangleNormal(m1.point[0], m1.point[1], m1.point[2]);
            for (int i = 0; i < 3; ++i) tri.n[i] = n[i];
        }
        pfc::InsDev ins;
        std::memset(&ins, 0, sizeof ins);
        ins.kind1 = kind1; ins.n_quad = b.quad.n; ins.prim_base1 = kind1 ? 1 : 0; ins.prim_base2 = 0;
        ins.chi = chi; ins.Ebar1 = kind1 ? m1.Ebar : 0.0; ins.Ebar2 = m2.Ebar;
        pfc::SceneDev sc;
        std::memset(&sc, 0, sizeof sc);
        sc.tets = tets; sc.tris = &tri; sc.ins = &ins; sc.n_ins = 1;
        pfc::PatchCtx<double> cx;
        for (int i = 0; i < 9; ++i) cx.x21.r[i] = R[i];
        for (int i = 0; i < 3; ++i) cx.x21.t[i] = tr[i];
        cx.x12 = pfc::inverse(cx.x21);
        cx.w_ang = pfc::mk<double>(tw[0], tw[1], tw[2]); cx.w_lin = pfc::mk<double>(tw[3], tw[4], tw[5]);
        cx.chi = chi; cx.Ebar1 = ins.Ebar1; cx.Ebar2 = ins.Ebar2; cx.n_quad = ins.n_quad;
        auto fresh = [&](pfc::Accum<double>& acc, Points& pts) {
            pts.d.assign(8 * 64, 0.0);
            acc.fp = ins.p; acc.w_ang = cx.w_ang; acc.w_lin = cx.w_lin; acc.dump = pts.d.data(); acc.dump_cap = 64;
            acc.reset(pfc::ACC_DUMP);
        };
        // ---- route 1: the tile kernel's compacted clip
        Points p_tile;
        {
            pfc::Accum<double> acc;
            fresh(acc, p_tile);
            double zr[16];
            const int n0 = pfc::start_polygon_zeta(sc, ins, 0, 0, cx, zr);
            if (n0 > 0) {
                pfc::PolyRec<double> slot;
                double* z = reinterpret_cast<double*>(&slot);
                for (int j = 0; j < 4 * n0; ++j) z[j] = zr[j];
                int flags = 0;
                const int n = pfc::clip_tet_inplace(z, n0, flags);
                if (n >= 3) {
                    pfc::finish_polygon_slot(n, tets[0], pfc::pair_normal(sc, ins, 0, 0, cx), slot);
                    for (int k = 0; k < n; ++k) {
                        const int kp = (k == 0) ? n - 1 : k - 1;
                        pfc::integrate_subtri(slot.v[kp], slot.v[k], slot.cen, slot.nrm, slot.eps_r, cx, acc);
                    }
                }
            }
            p_tile.n = acc.n_points;
        }
        // ---- route 2: one thread per pair
        Points p_pair;
        {
            pfc::Accum<double> acc;
            fresh(acc, p_pair);
            int flags = 0;
            pfc::integrate_pair(sc, ins, 0, 0, cx, acc, flags);
            p_pair.n = acc.n_points;
        }
        // ---- regularized Coulomb wrench of the pair (Accum::point in ACC_REGULARIZED) against yes_contact!(::Regularized)
        if (!b.traction.empty()) {
            const double v_c = (t % 4 == 1) ? 10.0 : 0.01 + 1.5 * u01(g);   // every point below v_c in a quarter of the cases
            const orc::Regularized reg = orc::make_regularized(v_c, 0.3 + 0.5 * u01(g), 0.2);
            const orc::V6<double> w_ref = orc::yes_contact_regularized(reg, b);
            const double fp[8] = {reg.mu_s, reg.mu_d, reg.v_c, reg.v_mu_s, reg.v_mu_d, (reg.mu_d - reg.mu_s) / (reg.v_mu_d - reg.v_mu_s), 1.0 / reg.v_c, 0.0};
            pfc::Accum<double, 6> acc;
            acc.fp = fp; acc.w_ang = cx.w_ang; acc.w_lin = cx.w_lin; acc.dump = nullptr; acc.dump_cap = 0;
            acc.reset(pfc::ACC_REGULARIZED);
            int flags = 0;
            pfc::integrate_pair(sc, ins, 0, 0, cx, acc, flags);
            double scale_ang = 0.0, scale_lin = 0.0;
            for (int i = 0; i < 3; ++i) { scale_ang = std::fmax(scale_ang, std::fabs(w_ref[i])); scale_lin = std::fmax(scale_lin, std::fabs(w_ref[3 + i])); }
            bool ok_w = true;
            for (int i = 0; i < 3; ++i) {
                if (std::fabs(acc.a[i] - w_ref[i]) > 1.0e-10 * std::fmax(scale_ang, 1.0e-6)) ok_w = false;
                if (std::fabs(acc.a[3 + i] - w_ref[3 + i]) > 1.0e-10 * std::fmax(scale_lin, 1.0e-6)) ok_w = false;
            }
            if (!ok_w && bad_wrench++ < 3) std::printf("regularized wrench differs in case %ld\n", t);
            ++n_wrench;
            // ---- the same pair in Jacobian mode: pose and twist carry 6 random partials each (Dual<6> on both sides)
            typedef orc::Dual<6> OD;
            typedef pfc::Dual<6> PD;
            orc::BodyBodyCache<OD> bd;
            pfc::PatchCtx<PD> cd;
            bd.quad = b.quad; bd.mesh_1 = &m1; bd.mesh_2 = &m2; bd.chi = chi; bd.Ebar = m2.Ebar;
            auto seeded = [&](double v, OD& o, PD& d) { o = OD(v); d = PD(v); for (int k = 0; k < 6; ++k) { const double s = u(g); o.p[k] = s; d.p[k] = s; } };
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) seeded(R[3 * i + j], bd.x_r2_r1(i, j), cd.x21.r[3 * i + j]);
                seeded(tr[i], bd.x_r2_r1(i, 3), cd.x21.t[i]);
                bd.x_r2_r1(3, i) = OD(0.0);
            }
            bd.x_r2_r1(3, 3) = OD(1.0);
            bd.x_r1_r2 = orc::inv_transform(bd.x_r2_r1);
            cd.x12 = pfc::inverse(cd.x21);
            for (int i = 0; i < 3; ++i) { seeded(tw[i], bd.twist_r2_r1_r2[i], (&cd.w_ang.x)[i]); seeded(tw[3 + i], bd.twist_r2_r1_r2[3 + i], (&cd.w_lin.x)[i]); }
            cd.chi = chi; cd.Ebar1 = ins.Ebar1; cd.Ebar2 = ins.Ebar2; cd.n_quad = ins.n_quad;
            if (kind1) orc::integrate_over_tet_tet(0, 0, bd); else orc::integrate_over_tri_tet(0, 0, bd);
            const orc::V6<OD> wd_ref = orc::yes_contact_regularized(reg, bd);
            pfc::Accum<PD, 6> accd;
            accd.fp = fp; accd.w_ang = cd.w_ang; accd.w_lin = cd.w_lin; accd.dump = nullptr; accd.dump_cap = 0;
            accd.reset(pfc::ACC_REGULARIZED);
            pfc::integrate_pair(sc, ins, 0, 0, cd, accd, flags);
            bool ok_d = (accd.n_points == (int)bd.traction.size());
            for (int grp = 0; grp < 2 && ok_d; ++grp) {
                double scale = 1.0e-6;
                for (int i = 0; i < 3; ++i) { scale = std::fmax(scale, std::fabs(wd_ref[3 * grp + i].v)); for (int k = 0; k < 6; ++k) scale = std::fmax(scale, std::fabs(wd_ref[3 * grp + i].p[k])); }
                for (int i = 0; i < 3; ++i) {
                    if (std::fabs(accd.a[3 * grp + i].v - wd_ref[3 * grp + i].v) > 1.0e-9 * scale) ok_d = false;
                    for (int k = 0; k < 6; ++k) if (std::fabs(accd.a[3 * grp + i].p[k] - wd_ref[3 * grp + i].p[k]) > 1.0e-9 * scale) ok_d = false;
                }
            }
            if (!ok_d && bad_dual++ < 3) std::printf("Dual-6 wrench differs in case %ld (kind1 %d, %d points vs %zu)\n", t, kind1, accd.n_points, bd.traction.size());
            // ---- and with seeds on the twist only (velocity seeds of the Jacobian): the device keeps the polygon in Float64 and runs
            // the twist-dependent part on Duals (pfc_twist.cuh); the oracle runs everything on Dual<6> with zero pose partials
            {
                orc::BodyBodyCache<OD> bt;
                bt.quad = b.quad; bt.mesh_1 = &m1; bt.mesh_2 = &m2; bt.chi = chi; bt.Ebar = m2.Ebar;
                for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) bt.x_r2_r1(i, j) = OD(bd.x_r2_r1(i, j).v);
                bt.x_r1_r2 = orc::inv_transform(bt.x_r2_r1);
                pfc::TwistAcc<6> tacc;
                for (int i = 0; i < 3; ++i) { seeded(tw[i], bt.twist_r2_r1_r2[i], (&tacc.w_ang.x)[i]); seeded(tw[3 + i], bt.twist_r2_r1_r2[3 + i], (&tacc.w_lin.x)[i]); }
                if (kind1) orc::integrate_over_tet_tet(0, 0, bt); else orc::integrate_over_tri_tet(0, 0, bt);
                const orc::V6<OD> wt_ref = orc::yes_contact_regularized(reg, bt);
                tacc.fp = fp; tacc.n_points = 0;
                for (int j = 0; j < 6; ++j) tacc.a[j] = PD(0.0);
                pfc::PolyRec<double> pr;
                int fl2 = 0;
                if (pfc::clip_pair(sc, ins, 0, 0, cx, pr, fl2)) {
                    pfc::Vec3<double> v2 = pr.v[pr.n - 1];
                    for (int k = 0; k < pr.n; ++k) { const pfc::Vec3<double> v1 = v2; v2 = pr.v[k]; pfc::integrate_subtri_tw(v1, v2, pr.cen, pr.nrm, pr.eps_r, cx.chi, cx.Ebar2, cx.n_quad, tacc); }
                }
                bool ok_t = (tacc.n_points == (int)bt.traction.size());
                for (int grp = 0; grp < 2 && ok_t; ++grp) {
                    double scale = 1.0e-6;
                    for (int i = 0; i < 3; ++i) { scale = std::fmax(scale, std::fabs(wt_ref[3 * grp + i].v)); for (int k = 0; k < 6; ++k) scale = std::fmax(scale, std::fabs(wt_ref[3 * grp + i].p[k])); }
                    for (int i = 0; i < 3; ++i) {
                        if (std::fabs(tacc.a[3 * grp + i].v - wt_ref[3 * grp + i].v) > 1.0e-9 * scale) ok_t = false;
                        for (int k = 0; k < 6; ++k) if (std::fabs(tacc.a[3 * grp + i].p[k] - wt_ref[3 * grp + i].p[k]) > 1.0e-9 * scale) ok_t = false;
                    }
                }
                if (!ok_t && bad_twist++ < 3) std::printf("twist-seeded wrench differs in case %ld (kind1 %d, %d points vs %zu)\n", t, kind1, tacc.n_points, bt.traction.size());
            }
        }
        const bool ok_tile = same(p_tile, b.traction), ok_pair = same(p_pair, b.traction);
        if (!ok_tile && bad_tile++ < 3) std::printf("tile route differs in case %ld (kind1 %d): %d points vs %zu\n", t, kind1, p_tile.n, b.traction.size());
        if (!ok_pair && bad_pair++ < 3) std::printf("pair route differs in case %ld (kind1 %d): %d points vs %zu\n", t, kind1, p_pair.n, b.traction.size());
        with_points += !b.traction.empty();
        n_points += (long)b.traction.size();
        tet_tet += kind1 && !b.traction.empty();
    }
    std::printf("cases %ld bad_tile %ld bad_pair %ld with_points %ld points %ld tet_tet_with_points %ld wrenches %ld bad_wrench %ld bad_dual %ld bad_twist %ld\n", n_case, bad_tile, bad_pair,
                with_points, n_points, tet_tet, n_wrench, bad_wrench, bad_dual, bad_twist);
    return (bad_tile || bad_pair || bad_wrench || bad_dual || bad_twist) ? 1 : 0;
}
"""


def test_device_narrow_phase_matches_oracle_point_by_point(tmp_path):
    cpp = tmp_path / "narrow_host.cpp"
    cpp.write_text(HARNESS)
    exe = tmp_path / "narrow_host"
    inc = ["-I", os.path.join(ROOT, "oracle"), "-I", os.path.join(ROOT, "pressurefieldcontact.jl_b200", "csrc"), "-I", "/usr/local/cuda/include"]
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-w", *inc, "-o", str(exe), str(cpp)])
    out = subprocess.run([str(exe), "200000"], capture_output=True, text=True)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0, out.stdout[-2000:]
    f = out.stdout.split()
    stats = {f[i]: int(f[i + 1]) for i in range(0, len(f) - 1) if f[i] in ("bad_tile", "bad_pair", "with_points", "points", "tet_tet_with_points", "wrenches", "bad_wrench", "bad_dual", "bad_twist")}
    assert stats["bad_tile"] == 0 and stats["bad_pair"] == 0 and stats["bad_wrench"] == 0 and stats["bad_dual"] == 0 and stats["bad_twist"] == 0 and stats["wrenches"] > 5000
    assert stats["with_points"] > 5000 and stats["tet_tet_with_points"] > 500   # the cases do produce contact polygons of both kinds
