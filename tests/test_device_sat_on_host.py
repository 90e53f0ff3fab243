"""CPU-only: the DEVICE source of the oriented-box overlap test (csrc/pfc_sat.cuh: sat_prepare_a + sat_test, built from
never-contracted __dmul_rn / __dadd_rn in the reference's evaluation order) compiled for the host with g++ and compared with
the oracle's BB_BB_intersect (oracle/pfc_oracle.hpp, following /root/reference/src/obb/bb_intersection.jl:2-74).

Candidate-pair lists are bit-exact only if this boolean is: every case must give the SAME answer, including cases placed on
the decision boundary -- for each random box pair the separation along a random direction is bisected down to neighbouring
doubles where the oracle's answer flips, and both implementations are evaluated there and a few ulps to either side."""
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
#include "pfc_oracle.hpp"
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
#include "pfc_sat.cuh"

static std::mt19937_64 g(148);
static std::uniform_real_distribution<double> u(-1.0, 1.0), u01(0.0, 1.0);

static void random_rotation(double* R) {   // row-major
    double q[4] = {u(g), u(g), u(g), u(g)};
    const double qn = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (double& x : q) x /= qn;
    const double M[9] = {1 - 2 * (q[2] * q[2] + q[3] * q[3]), 2 * (q[1] * q[2] - q[0] * q[3]), 2 * (q[1] * q[3] + q[0] * q[2]),
                         2 * (q[1] * q[2] + q[0] * q[3]), 1 - 2 * (q[1] * q[1] + q[3] * q[3]), 2 * (q[2] * q[3] - q[0] * q[1]),
                         2 * (q[1] * q[3] - q[0] * q[2]), 2 * (q[2] * q[3] + q[0] * q[1]), 1 - 2 * (q[1] * q[1] + q[2] * q[2])};
    std::memcpy(R, M, sizeof M);
}

struct Case { pfc::NodeRec a, b; double Rab[9], dir[3]; };

static bool oracle_says(const Case& c, double s) {
    orc::OBB A, B;
    orc::M3<double> R;
    orc::V3<double> t;
    for (int i = 0; i < 3; ++i) {
        A.c[i] = c.a.c[i]; A.e[i] = c.a.e[i]; B.c[i] = c.b.c[i]; B.e[i] = c.b.e[i]; t[i] = s * c.dir[i];
        for (int j = 0; j < 3; ++j) { A.R(i, j) = c.a.R[3 * i + j]; B.R(i, j) = c.b.R[3 * i + j]; R(i, j) = c.Rab[3 * i + j]; }
    }
    return orc::BB_BB_intersect<double>(R, t, A, B);
}
static bool device_says(const Case& c, double s) {
    const double tab[3] = {s * c.dir[0], s * c.dir[1], s * c.dir[2]};
    pfc::SatA A;
    pfc::sat_prepare_a(c.a, c.Rab, tab, A);      // nodes flagged kNodeInternalAabb take the path that skips the identity products
    return pfc::sat_test(A, c.b);
}
static bool device_general_says(const Case& c, double s) {   // the same boxes through the general path (every product evaluated)
    const double tab[3] = {s * c.dir[0], s * c.dir[1], s * c.dir[2]};
    pfc::SatA A;
    pfc::sat_prepare_a<false>(c.a, c.Rab, tab, A);
    return pfc::sat_test<false>(A, c.b);
}

int main(int argc, char** argv) {
    const long n_case = argc > 1 ? atol(argv[1]) : 20000;
    long bad = 0, n_eval = 0, n_boundary = 0, n_true = 0, n_aabb = 0, bad_special = 0;
    for (long k = 0; k < n_case; ++k) {
        Case c;
        std::memset(&c, 0, sizeof c);
        random_rotation(c.a.R); random_rotation(c.b.R); random_rotation(c.Rab);
        if (k % 5 == 0) { const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; std::memcpy(c.b.R, I, sizeof I); std::memcpy(c.Rab, I, sizeof I); std::memcpy(c.a.R, I, sizeof I); }   // axis-aligned: parallel edges, the 1e-14 guard decides
        for (int i = 0; i < 3; ++i) { c.a.c[i] = 0.2 * u(g); c.b.c[i] = 0.2 * u(g); c.a.e[i] = 0.05 + u01(g); c.b.e[i] = 0.05 + u01(g); c.dir[i] = u(g); }
        if (k % 7 == 0) c.b.e[k % 3] = 0.0;   // flat box (a triangle's leaf box)
        // axis-aligned internal nodes (what the trees' internal boxes are): a, b or both carry R = I and the kind that lets the device skip
        // the identity products; centres with +-0.0 and denormal components, a translation direction with a zero component
        const double I9[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        if (k % 3 == 1 || k % 5 == 0) { std::memcpy(c.a.R, I9, sizeof I9); c.a.kind = pfc::kNodeInternalAabb; }
        if (k % 3 == 2 || k % 5 == 0) { std::memcpy(c.b.R, I9, sizeof I9); c.b.kind = pfc::kNodeInternalAabb; }
        if (k % 11 == 0) { c.a.c[k % 3] = (k % 2) ? 0.0 : -0.0; c.b.c[(k + 1) % 3] = 4.9e-324; c.dir[(k + 2) % 3] = 0.0; }
        n_aabb += (c.a.kind == pfc::kNodeInternalAabb) || (c.b.kind == pfc::kNodeInternalAabb);
        auto check = [&](double s) {
            const bool o = oracle_says(c, s), d = device_says(c, s);
            ++n_eval; n_true += o;
            if (o != d && bad++ < 5) std::printf("case %ld at s = %.17g: oracle %d device %d\n", k, s, (int)o, (int)d);
            if (d != device_general_says(c, s) && bad_special++ < 5) std::printf("case %ld at s = %.17g: the axis-aligned path and the general path differ\n", k, s);
        };
        for (int r = 0; r < 4; ++r) check(4.0 * u01(g));
        // bisect the flip along dir
        double lo = 0.0, hi = 16.0;
        if (!oracle_says(c, lo) || oracle_says(c, hi)) continue;
        for (int it = 0; it < 80 && std::nextafter(lo, hi) < hi; ++it) { const double mid = 0.5 * (lo + hi); (oracle_says(c, mid) ? lo : hi) = mid; }
        ++n_boundary;
        double s = lo;
        for (int r = 0; r < 4; ++r) s = std::nextafter(s, 0.0);
        for (int r = 0; r < 9; ++r) { check(s); s = std::nextafter(s, 32.0); }
    }
    std::printf("cases %ld evaluations %ld boundaries %ld overlapping %ld bad %ld axis_aligned_cases %ld bad_special %ld\n", n_case, n_eval, n_boundary, n_true, bad, n_aabb,
                bad_special);
    return bad != 0 || bad_special != 0;
}
"""


def test_device_sat_boolean_equals_oracle_everywhere(tmp_path):
    cpp = tmp_path / "sat_host.cpp"
    cpp.write_text(HARNESS)
    exe = tmp_path / "sat_host"
    inc = ["-I", os.path.join(ROOT, "oracle"), "-I", os.path.join(ROOT, "pressurefieldcontact.jl_b200", "csrc"), "-I", "/usr/local/cuda/include"]
    # -ffp-contract=off: __dmul_rn / __dadd_rn never contract on the device, and neither may their host stand-ins
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-w", *inc, "-o", str(exe), str(cpp)])
    out = subprocess.run([str(exe), "20000"], capture_output=True, text=True)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0, out.stdout[-2000:]
    f = out.stdout.split()
    stats = {f[i]: int(f[i + 1]) for i in range(0, len(f) - 1) if f[i] in ("evaluations", "boundaries", "overlapping", "bad", "axis_aligned_cases", "bad_special")}
    assert stats["bad"] == 0 and stats["bad_special"] == 0 and stats["axis_aligned_cases"] > 10000
    assert stats["boundaries"] > 10000 and stats["evaluations"] > 150000 and 0 < stats["overlapping"] < stats["evaluations"]
