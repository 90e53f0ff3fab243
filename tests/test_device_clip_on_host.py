"""CPU-only: the DEVICE source of the polygon clips (csrc/pfc_clip.cuh: clip_node_inplace, clip_tet_inplace of the tile kernel;
clip_node, clip_tet<T> of the large path and the Dual mode -- the latter must give the same bits as the in-place one), compiled
for the host with g++ and run against the oracle's clip_in_tet_coordinates (oracle/pfc_oracle.hpp, which follows
/root/reference/src/clip/static_clip.jl:7-201) on random 3- and 4-gons in tetrahedral coordinates -- vertices on faces (+-0.0),
polygons fully inside / outside, NaNs.  Vertex counts, flags and every output coordinate must agree BIT FOR BIT: the device
version works in place on sign masks and rotates through registers -- a different control structure than the recursion it
restates -- so this pins its logic without a GPU.

One deliberate deviation: weightPoly divides twice (w1 / (w1 - w2), w2 / (w1 - w2), src/math_kernel/utility.jl:21-26), the device
multiplies by one reciprocal.  The harness therefore runs the device source twice: with the reciprocal replaced by the
reference's two divisions (text substitution) the outputs must be bit-identical; as shipped, vertex counts and flags must be
the same and the coordinates within 1e-12 (relative to the larger of 1 and the value)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r"""
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include "pfc_oracle.hpp"
#define PFC_D static inline
#define __ffs __builtin_ffs
enum { kFlagNonFinite = 2 };
namespace dev {
%s
}
namespace dev_div {
%s
}
#define __device__
#define __noinline__
namespace dev_generic {   // the thread-private clip_tet<T> of the large path / Dual mode, T = double
static inline double val(double x) { return x; }
%s
}
int main(int argc, char** argv) {
    const long n_case = argc > 1 ? atol(argv[1]) : 1000000;
    std::mt19937_64 g(20261018);
    std::uniform_real_distribution<double> u(-1.0, 1.5);
    long bad = 0, survivors = 0, cut7 = 0, nonfinite = 0;
    for (long t = 0; t < n_case; ++t) {
        const int n0 = 3 + (int)(g() %% 2);
        double z[32] = {0};
        orc::Poly<4, double> p;
        p.n = n0;
        for (int k = 0; k < n0; ++k)
            for (int i = 0; i < 4; ++i) {
                double v = u(g);
                const unsigned r = g() %% 16;
                if (r == 0) v = 0.0;
                if (r == 1) v = -0.0;
                if (r == 2) v = 1.0e-300;
                z[4 * k + i] = v;
            }
        if (t %% 1000 == 7) z[g() %% (4 * n0)] = std::nan("");
        if (t %% 1000 == 8) for (int k = 0; k < n0; ++k) z[4 * k + 1] = std::nan("");
        for (int k = 0; k < n0; ++k) for (int i = 0; i < 4; ++i) p.v[k][i] = z[4 * k + i];
        orc::ClipStatus st;
        const orc::Poly<4, double> ref = orc::clip_in_tet_coordinates(p, st);
        double z_div[32];
        std::memcpy(z_div, z, sizeof z);
        int flags = 0, flags_div = 0;
        const int n = dev::clip_tet_inplace(z, n0, flags);
        const int n_div = dev_div::clip_tet_inplace(z_div, n0, flags_div);
        dev_generic::Zeta<double> zg[9];
        for (int k = 0; k < n0; ++k) for (int i = 0; i < 4; ++i) zg[k].c[i] = p.v[k][i];
        int flags_g = 0;
        const int n_g = dev_generic::clip_tet<double>(zg, n0, flags_g);
        bool ok = (n == ref.n) && (((flags & kFlagNonFinite) != 0) == st.non_finite) && (n_div == ref.n) && (flags_div == flags) && (n_g == n) && (flags_g == flags);
        for (int k = 0; ok && k < n; ++k)   // both device clips use the same reciprocal form: identical bits
            for (int i = 0; i < 4; ++i)
                if (std::memcmp(&zg[k].c[i], &z[4 * k + i], sizeof(double)) != 0 && !(zg[k].c[i] != zg[k].c[i] && z[4 * k + i] != z[4 * k + i])) ok = false;
        for (int k = 0; ok && k < n; ++k)
            for (int i = 0; i < 4; ++i) {
                const double a = z[4 * k + i], a_div = z_div[4 * k + i], b = ref.v[k][i];
                if (std::memcmp(&a_div, &b, sizeof b) != 0 && !(a_div != a_div && b != b)) ok = false;   // same divisions: bit for bit
                const double scale = std::fmax(1.0, std::fmax(std::fabs(a), std::fabs(b)));
                if (!(std::fabs(a - b) <= 1.0e-12 * scale) && !(a != a && b != b)) ok = false;           // one reciprocal: rounding only
            }
        if (!ok && bad++ < 5) std::printf("mismatch in case %%ld: n %%d vs %%d, flags %%d vs %%d\n", t, n, ref.n, flags, (int)st.non_finite);
        survivors += n >= 3;
        cut7 += n >= 7;
        nonfinite += st.non_finite;
    }
    std::printf("cases %%ld bad %%ld survivors %%ld seven_or_more %%ld non_finite %%ld\n", n_case, bad, survivors, cut7, nonfinite);
    return bad != 0;
}
"""


def _device_clip_source():
    src = open(os.path.join(ROOT, "pressurefieldcontact.jl_b200", "csrc", "pfc_clip.cuh")).read()
    a = src.index("PFC_D void clip_node_inplace(")
    b = src.index("// zero_small_coordinates")
    return src[a:b]


def _device_generic_clip_source():
    src = open(os.path.join(ROOT, "pressurefieldcontact.jl_b200", "csrc", "pfc_clip.cuh")).read()
    a = src.index("template <class T> struct Zeta")
    b = src.index("// The same clip, working IN PLACE")
    return src[a:b]


def test_device_inplace_clip_matches_oracle_bitwise(tmp_path):
    cpp = tmp_path / "clip_host.cpp"
    dev = _device_clip_source()
    recip = "const double inv = 1.0 / (w1 - w2);\n    const double c1 = w1 * inv, c2 = w2 * inv;"
    assert recip in dev
    dev_div = dev.replace(recip, "const double sum_weight = w1 - w2;\n    const double c1 = w1 / sum_weight, c2 = w2 / sum_weight;")
    cpp.write_text(HARNESS % (dev, dev_div, _device_generic_clip_source()))
    exe = tmp_path / "clip_host"
    # -ffp-contract=off as in oracle/Makefile: the clip has no muladd site in the reference
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-w", "-I", os.path.join(ROOT, "oracle"), "-o", str(exe), str(cpp)])
    out = subprocess.run([str(exe), "1000000"], capture_output=True, text=True)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0, out.stdout[-2000:]
    fields = out.stdout.split()
    stats = {fields[i]: int(fields[i + 1]) for i in range(0, len(fields) - 1, 2) if fields[i] in ("bad", "survivors", "seven_or_more", "non_finite")}
    assert stats["bad"] == 0
    assert stats["survivors"] > 100000 and stats["non_finite"] > 0   # the cases exercise the cut paths and the error path
