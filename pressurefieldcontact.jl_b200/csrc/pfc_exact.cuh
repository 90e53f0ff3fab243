// pfc_exact.cuh -- the narrow phase and the bristle friction model in the REFERENCE'S OPERATION ORDER.
//
// Why this exists.  The bristle model (src/contact_algorithms_friction.jl:85-143, paths relative to /root/reference) multiplies by
// K̄^(-1/2) = V diag(1 / sqrt(max(lambda, 1e-16 lambda_max))) V'.  decompose_K! scales the rotational block of K by magic^2 = 1e-6 and
// contact patches are flat or small, so K̄ is rank deficient to working precision: its noise eigenvalues sit AT the 1e-16 clamp, and
// rounding-level changes of K (a different summation order, a fused multiply-add, a hoisted reciprocal) move the result by up to
// 1e8 eps.  The bristle wrench and s-dot therefore only reproduce (to the 1e-9 the path is held to) when K is reproduced BIT FOR
// BIT.  This header follows the reference operation by operation:
//   * products of small matrices as StaticArrays evaluates them (row-by-column, summed left to right, including the terms that
//     multiply the constant 0/1 bottom row), x_zeta2_r1 = x_zeta2_r2 * x_r2_r1 formed before it is applied to the vertices
//     (src/contact_algorithms_non_friction.jl:196-215), the tet-tet plane as (E eps) * (x_zeta1_r1 * x_r1_r2) (:164-194);
//   * weightPoly with its two divisions (src/math_kernel/utility.jl:21-26), centroid / triangle_area / pressure law as written
//     (src/clip/poly_eight.jl:35-52, src/math_kernel/geometry_kernel.jl:3-10, non_friction.jl:217-265), muladd only where the
//     reference says muladd (src/math_kernel/vector_projections.jl);
//   * every sum over traction points SEQUENTIAL in TractionCache order -- pair, polygon edge, quadrature point
//     (src/contact_algorithms_normal.jl:17-34, friction.jl:147-201);
//   * Dual numbers with ForwardDiff's rules stated term by term (XD<N> below).
// The translation unit that includes this header is compiled with -fmad=false (nvcc) / -ffp-contract=off (g++ host check): no
// contraction other than the explicit fma() calls.  tests/test_device_exact_on_host.py compiles it for the host and compares it
// with the CPU oracle bit for bit.
// Per-tetrahedron constants (inv([V;1]), eps * inv) are taken from TetRec: pfc_finalize computes them with the arithmetic the
// reference applies per pair (the adjugate formula, /root/repo/pressurefieldcontact.jl_b200/csrc/pfc_api.cu::invert_tet_matrix).
#pragma once
#include <cmath>

#include "pfc_math.cuh"
#include "pfc_types.cuh"

namespace pfc {
namespace ex {

#ifdef PFC_HOST_CHECK
#define PFC_XHD inline
#define PFC_XD_ inline
#else
#define PFC_XHD __host__ __device__ inline
#define PFC_XD_ __device__ inline
#endif

// ---- ForwardDiff.Dual{Nothing,Float64,N}, rule by rule --------------------------------------------------------------------------
template <int N> struct XD {
    double v;
    double p[N];
    PFC_XHD XD() {}
    PFC_XHD XD(double x) : v(x) { for (int i = 0; i < N; ++i) p[i] = 0.0; }
};
PFC_XHD double xval(double x) { return x; }
template <int N> PFC_XHD double xval(const XD<N>& x) { return x.v; }
#define PFC_XD_LOOP for (int i = 0; i < N; ++i)
template <int N> PFC_XHD XD<N> operator-(const XD<N>& a) { XD<N> r; r.v = -a.v; PFC_XD_LOOP r.p[i] = -a.p[i]; return r; }
template <int N> PFC_XHD XD<N> operator+(const XD<N>& a, const XD<N>& b) { XD<N> r; r.v = a.v + b.v; PFC_XD_LOOP r.p[i] = a.p[i] + b.p[i]; return r; }
template <int N> PFC_XHD XD<N> operator-(const XD<N>& a, const XD<N>& b) { XD<N> r; r.v = a.v - b.v; PFC_XD_LOOP r.p[i] = a.p[i] - b.p[i]; return r; }
template <int N> PFC_XHD XD<N> operator*(const XD<N>& a, const XD<N>& b) { XD<N> r; r.v = a.v * b.v; PFC_XD_LOOP r.p[i] = a.p[i] * b.v + a.v * b.p[i]; return r; }
template <int N> PFC_XHD XD<N> operator/(const XD<N>& a, const XD<N>& b) {
    XD<N> r; r.v = a.v / b.v;
    PFC_XD_LOOP r.p[i] = (a.p[i] - r.v * b.p[i]) / b.v;
    return r;
}
template <int N> PFC_XHD XD<N> operator+(const XD<N>& a, double b) { XD<N> r = a; r.v = a.v + b; return r; }
template <int N> PFC_XHD XD<N> operator+(double a, const XD<N>& b) { XD<N> r = b; r.v = a + b.v; return r; }
template <int N> PFC_XHD XD<N> operator-(const XD<N>& a, double b) { XD<N> r = a; r.v = a.v - b; return r; }
template <int N> PFC_XHD XD<N> operator-(double a, const XD<N>& b) { XD<N> r = -b; r.v = a - b.v; return r; }
template <int N> PFC_XHD XD<N> operator*(const XD<N>& a, double b) { XD<N> r; r.v = a.v * b; PFC_XD_LOOP r.p[i] = a.p[i] * b; return r; }
template <int N> PFC_XHD XD<N> operator*(double a, const XD<N>& b) { return b * a; }
template <int N> PFC_XHD XD<N> operator/(const XD<N>& a, double b) { XD<N> r; r.v = a.v / b; PFC_XD_LOOP r.p[i] = a.p[i] / b; return r; }
template <int N> PFC_XHD XD<N> operator/(double a, const XD<N>& b) { return XD<N>(a) / b; }

PFC_XHD double xsqrt(double x) { return sqrt(x); }
template <int N> PFC_XHD XD<N> xsqrt(const XD<N>& a) { XD<N> r; r.v = sqrt(a.v); PFC_XD_LOOP r.p[i] = a.p[i] / (2.0 * r.v); return r; }
// muladd: a hardware fma on the value part (Julia emits one on FMA-capable x86), ForwardDiff's product rule on the partials
PFC_XHD double xmuladd(double a, double b, double c) { return fma(a, b, c); }
template <int N> PFC_XHD XD<N> xmuladd(const XD<N>& a, const XD<N>& b, const XD<N>& c) {
    XD<N> r; r.v = fma(a.v, b.v, c.v); PFC_XD_LOOP r.p[i] = a.p[i] * b.v + a.v * b.p[i] + c.p[i]; return r; }
template <int N> PFC_XHD XD<N> xmuladd(double a, const XD<N>& b, double c) { XD<N> r; r.v = fma(a, b.v, c); PFC_XD_LOOP r.p[i] = a * b.p[i]; return r; }
template <int N> PFC_XHD XD<N> xmuladd(double a, const XD<N>& b, const XD<N>& c) { XD<N> r; r.v = fma(a, b.v, c.v); PFC_XD_LOOP r.p[i] = a * b.p[i] + c.p[i]; return r; }
template <class T> PFC_XHD T xmax(const T& x, const T& y) { return (xval(y) < xval(x)) ? x : y; }   // Julia: max(x, y) = ifelse(y < x, x, y)
template <class T> PFC_XHD T xclamp(const T& x, double lo, double hi) {                              // constants lose their partials
    if (xval(x) > hi) return T(hi);
    if (xval(x) < lo) return T(lo);
    return x;
}

// ---- small static vectors --------------------------------------------------------------------------------------------------------
template <class T> struct X3 { T x[3]; PFC_XHD T& operator[](int i) { return x[i]; } PFC_XHD const T& operator[](int i) const { return x[i]; } };
template <class T> struct X4 { T x[4]; PFC_XHD T& operator[](int i) { return x[i]; } PFC_XHD const T& operator[](int i) const { return x[i]; } };
template <class T> PFC_XHD X3<T> x3(const T& a, const T& b, const T& c) { X3<T> r; r[0] = a; r[1] = b; r[2] = c; return r; }
template <class T> PFC_XHD X3<T> operator+(const X3<T>& a, const X3<T>& b) { return x3<T>(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
template <class T> PFC_XHD X3<T> operator-(const X3<T>& a, const X3<T>& b) { return x3<T>(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
template <class T, class S> PFC_XHD X3<T> xscale(const X3<T>& a, const S& s) { return x3<T>(a[0] * s, a[1] * s, a[2] * s); }
template <class T, class S> PFC_XHD X3<T> xdivide(const X3<T>& a, const S& s) { return x3<T>(a[0] / s, a[1] / s, a[2] / s); }
template <class T> PFC_XHD T xdot(const X3<T>& a, const X3<T>& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <class T> PFC_XHD X3<T> xcross(const X3<T>& a, const X3<T>& b) {
    return x3<T>(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
// src/math_kernel/geometry_kernel.jl:3-10
template <class T> PFC_XHD T triangle_area(const X3<T>& v1, const X3<T>& v2, const X3<T>& v3, const X3<T>& n) { return xdot(n, xscale(xcross(v2 - v1, v3 - v2), 0.5)); }
// src/math_kernel/utility.jl:21-26
template <class T> PFC_XHD X4<T> weight_poly4(const X4<T>& p1, const X4<T>& p2, const T& w1, const T& w2) {
    const T sum_weight = w1 - w2;
    const T c1 = w1 / sum_weight;
    const T c2 = w2 / sum_weight;
    X4<T> r;
    for (int i = 0; i < 4; ++i) r[i] = c1 * p2[i] - c2 * p1[i];
    return r;
}
template <class T> PFC_XHD X3<T> weight_poly3(const X3<T>& p1, const X3<T>& p2, const T& w1, const T& w2) {
    const T sum_weight = w1 - w2;
    const T c1 = w1 / sum_weight;
    const T c2 = w2 / sum_weight;
    X3<T> r;
    for (int i = 0; i < 3; ++i) r[i] = c1 * p2[i] - c2 * p1[i];
    return r;
}

// ---- clip_in_tet_coordinates (src/clip/static_clip.jl:7-201) ---------------------------------------------------------------------
// One loop over the four faces on a thread-private vertex array; decisions as in pfc_clip.cuh::clip_tet (rotation to the first
// non-positive -> positive transition, cut_clip's arity-reducing recursion, the strict / non-strict asymmetry between 3-5 and 6-7
// vertices, the 7-vertex cut returning at once), new vertices by weightPoly as written.
template <class T> __device__ __noinline__ int clip_tet_exact(X4<T>* z, int n, int& flags) {
    for (int i = 0; i < 4; ++i) {
        bool all_non_pos = true, all_non_neg = true;
        unsigned non_pos = 0;
        for (int k = 0; k < n; ++k) {
            const double s = xval(z[k][i]);
            const bool np = (s <= 0.0);
            non_pos |= (np ? 1u : 0u) << k;
            all_non_pos = all_non_pos && np;
            all_non_neg = all_non_neg && (0.0 <= s);
        }
        if (all_non_pos) return 0;
        if (all_non_neg) continue;
        int k0 = -1;
        for (int k = 0; k < n; ++k) {
            const int k1 = (k + 1 == n) ? 0 : k + 1;
            if (((non_pos >> k) & 1u) && !((non_pos >> k1) & 1u)) { k0 = k; break; }
        }
        if (k0 < 0) { flags |= kFlagNonFinite; return 0; }
        X4<T> w[8];
        for (int j = 0; j < n; ++j) { int k = k0 + j; if (k >= n) k -= n; w[j] = z[k]; }
        int m = n;
        while (m > 3 && xval(w[m - 2][i]) <= 0.0) --m;
        const X4<T> z_start = weight_poly4(w[0], w[1], w[0][i], w[1][i]);
        const double last = xval(w[m - 1][i]);
        const bool last_inside = (m <= 5) ? (0.0 < last) : (0.0 <= last);
        if (last_inside) {
            const X4<T> z_end = weight_poly4(w[0], w[m - 1], w[0][i], w[m - 1][i]);
            z[0] = z_start;
            for (int k = 1; k < m; ++k) z[k] = w[k];
            z[m] = z_end;
            n = m + 1;
        } else {
            const X4<T> z_end = weight_poly4(w[m - 1], w[m - 2], w[m - 1][i], w[m - 2][i]);
            z[0] = z_start;
            for (int k = 1; k < m - 1; ++k) z[k] = w[k];
            z[m - 1] = z_end;
            n = m;
        }
        if (m == 7) return n;
    }
    return n;
}

// ---- clip_plane_tet (src/clip/plane_tet_intersection.jl:9-106) -------------------------------------------------------------------
// plane: 4 coefficients; v: the 4 vertices of tetrahedron 1 in r2 (columns of x_r2_zeta1, whose bottom row is exactly 1)
template <class T> __device__ __noinline__ int plane_tet_exact(const X4<T>& plane, const X3<T>* v, X3<T>* out) {
    T proj[4];
    int n_neg = 0, n_pos = 0;
    unsigned pos = 0, neg = 0;
    for (int k = 0; k < 4; ++k) {
        proj[k] = plane[0] * v[k][0] + plane[1] * v[k][1] + plane[2] * v[k][2] + plane[3] * T(1.0);
        const double p = xval(proj[k]);
        if (p < 0.0) { ++n_neg; neg |= 1u << k; }
        if (0.0 < p) { ++n_pos; pos |= 1u << k; }
    }
    if (n_pos == 0 || n_neg == 0) return 0;
    const signed char TRI[4][3][2] = {{{1, 0}, {3, 0}, {2, 0}}, {{0, 1}, {2, 1}, {3, 1}}, {{0, 2}, {3, 2}, {1, 2}}, {{0, 3}, {1, 3}, {2, 3}}};
    const signed char QUAD[3][4][2] = {{{1, 2}, {1, 3}, {0, 3}, {0, 2}}, {{0, 1}, {0, 3}, {2, 3}, {2, 1}}, {{0, 2}, {0, 1}, {3, 1}, {3, 2}}};
    int cnt, sel;
    bool forward;
    if (n_pos == 1) { sel = __ffs(pos) - 1; cnt = 3; forward = true; }
    else if (n_neg == 1) { sel = __ffs(neg) - 1; cnt = 3; forward = false; }
    else {
        const unsigned p0 = pos & 1u;
        if (((pos >> 1) & 1u) == p0) sel = 0;
        else if (((pos >> 2) & 1u) == p0) sel = 1;
        else if (((pos >> 3) & 1u) == p0) sel = 2;
        else return 0;
        cnt = 4;
        forward = (0.0 < xval(proj[0]));
    }
    for (int k = 0; k < cnt; ++k) {
        const int i1 = (cnt == 3) ? TRI[sel][k][0] : QUAD[sel][k][0];
        const int i2 = (cnt == 3) ? TRI[sel][k][1] : QUAD[sel][k][1];
        out[forward ? k : cnt - 1 - k] = weight_poly3(v[i1], v[i2], proj[i1], proj[i2]);
    }
    return cnt;
}

// ---- per (environment, instruction) inputs in mode T ------------------------------------------------------------------------------
template <class T> struct ExCtx {
    T X21[16];     // x_r2_r1.mat, column-major as handed over by the caller (bottom row included)
    T X12[16];     // inv(x_r2_r1): R', -(R' t) (RigidBodyDynamics inv(::Transform3D)), bottom row 0 0 0 1
    T twist[6];    // angular, linear
    double chi, Ebar1, Ebar2;
    int n_quad;    // 1 or 3 points per sub-triangle
};
template <class T> PFC_XHD void make_x12(ExCtx<T>& c) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) c.X12[4 * j + i] = c.X21[4 * i + j];
    for (int i = 0; i < 3; ++i) c.X12[12 + i] = -(c.X12[i] * c.X21[12] + c.X12[4 + i] * c.X21[13] + c.X12[8 + i] * c.X21[14]);
    c.X12[3] = T(0.0); c.X12[7] = T(0.0); c.X12[11] = T(0.0); c.X12[15] = T(1.0);
}

// One traction point (TractionCache, src/mechanism_scenario.jl:51-58) without its normal, which is shared by the pair's points.
template <class T> struct ExPoint { T r[3]; T dA; T p; };
constexpr int kExMaxPoints = 24;   // 8 polygon edges x 3 quadrature points

// XiaoGimbutas triangle rules 1 and 2 (src/clip/quadrature.jl:21-41), literals as in the reference
#define PFC_XQ1 0.33333333333333331483
#define PFC_XQA 0.16666666666666674068
#define PFC_XQB 0.66666666666666651864

// integrate_over_polygon_patch! (non_friction.jl:217-265): polygon in tetrahedral coordinates of tet 2 -> traction points
template <class T> PFC_XD_ int polygon_points(const X4<T>* z, int n, const TetRec& t2, const X3<T>& n2, const ExCtx<T>& cx, ExPoint<T>* out) {
    X3<T> pv[8];
    for (int k = 0; k < n; ++k)   // mul_then_un_pad(x_r2_zeta2, .)
        for (int i = 0; i < 3; ++i) pv[k][i] = t2.v[i] * z[k][0] + t2.v[3 + i] * z[k][1] + t2.v[6 + i] * z[k][2] + t2.v[9 + i] * z[k][3];
    // centroid (src/clip/poly_eight.jl:35-52)
    X3<T> cen;
    {
        const X3<T> a = pv[0];
        X3<T> c = pv[1];
        T cum_sum = T(0.0);
        X3<T> cum_prod = x3<T>(T(0.0), T(0.0), T(0.0));
        for (int k = 2; k < n; ++k) {
            const X3<T> b = c;
            c = pv[k];
            const T area = triangle_area(a, b, c, n2);
            cum_prod = cum_prod + xscale(xscale(a + b + c, double(1.0 / 3.0)), area);
            cum_sum = cum_sum + area;
        }
        cen = (xval(cum_sum) == 0.0) ? a : xdivide(cum_prod, cum_sum);
    }
    const X3<T> ang = x3<T>(cx.twist[0], cx.twist[1], cx.twist[2]), lin = x3<T>(cx.twist[3], cx.twist[4], cx.twist[5]);
    int np = 0;
    X3<T> v2 = pv[n - 1];
    for (int k = 0; k < n; ++k) {
        const X3<T> v1 = v2;
        v2 = pv[k];
        const T area = triangle_area(v1, v2, cen, n2);
        if (!(0.0 < xval(area))) continue;
        for (int q = 0; q < cx.n_quad; ++q) {
            double za, zb, zc, w;
            if (cx.n_quad == 1) { za = zb = zc = PFC_XQ1; w = 1.0; }
            else { za = (q == 1) ? PFC_XQB : PFC_XQA; zb = (q == 0) ? PFC_XQB : PFC_XQA; zc = (q == 2) ? PFC_XQB : PFC_XQA; w = PFC_XQ1; }
            X3<T> r;
            for (int i = 0; i < 3; ++i) r[i] = v1[i] * za + v2[i] * zb + cen[i] * zc;
            T eps = xmuladd(t2.eps_r[0], r[0], t2.eps_r[3]);   // a_dot_one_pad_b
            eps = xmuladd(t2.eps_r[1], r[1], eps);
            eps = xmuladd(t2.eps_r[2], r[2], eps);
            const X3<T> rd = lin + xcross(ang, r);
            const T ee = -(t2.eps_r[0] * rd[0] + t2.eps_r[1] * rd[1] + t2.eps_r[2] * rd[2]);
            const T damp = xmax(T(0.0), 1.0 + cx.chi * ee);
            const T p = eps * cx.Ebar2 * damp;
            if (0.0 < xval(p)) {
                ExPoint<T>& o = out[np++];
                o.r[0] = r[0]; o.r[1] = r[1]; o.r[2] = r[2];
                o.dA = w * area;
                o.p = p;
            }
        }
    }
    return np;
}

// integrate_over! for one candidate pair: traction points + their common normal; returns the point count
// tri-tet: non_friction.jl:196-215
template <class T> PFC_XD_ int pair_points_tri_tet(const TriRec& tri, const TetRec& t2, const ExCtx<T>& cx, X3<T>& n2, ExPoint<T>* out, int& flags) {
    T A[16];   // x_zeta2_r1 = x_zeta2_r2 * x_r2_r1, row-major A[4 i + j]
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i)
            A[4 * i + j] = t2.inv[4 * i] * cx.X21[4 * j] + t2.inv[4 * i + 1] * cx.X21[4 * j + 1] + t2.inv[4 * i + 2] * cx.X21[4 * j + 2] + t2.inv[4 * i + 3] * cx.X21[4 * j + 3];
    X4<T> z[8];
    for (int k = 0; k < 3; ++k) {
        const T o0 = T(tri.v[3 * k]), o1 = T(tri.v[3 * k + 1]), o2 = T(tri.v[3 * k + 2]), o3 = T(1.0);
        for (int i = 0; i < 4; ++i) z[k][i] = A[4 * i] * o0 + A[4 * i + 1] * o1 + A[4 * i + 2] * o2 + A[4 * i + 3] * o3;
    }
    const int n = clip_tet_exact(z, 3, flags);
    if (n < 3) return 0;
    const T m0 = T(tri.n[0]), m1 = T(tri.n[1]), m2 = T(tri.n[2]);
    for (int i = 0; i < 3; ++i) n2[i] = cx.X21[i] * m0 + cx.X21[4 + i] * m1 + cx.X21[8 + i] * m2;   // R(x_r2_r1) * triangleNormal
    return polygon_points(z, n, t2, n2, cx, out);
}
// tet-tet: non_friction.jl:164-194.  eps1 / eps2: the pressure-field values at the 4 vertices of each tetrahedron
template <class T> PFC_XD_ int pair_points_tet_tet(const TetRec& t1, const double* eps1, const TetRec& t2, const double* eps2, const ExCtx<T>& cx, X3<T>& n2,
                                                 ExPoint<T>* out, int& flags) {
    X4<T> plane;
    {
        double E1e[4], E2e[4];
        for (int k = 0; k < 4; ++k) { E1e[k] = cx.Ebar1 * eps1[k]; E2e[k] = cx.Ebar2 * eps2[k]; }
        T B[16];   // x_zeta1_r2 = x_zeta1_r1 * x_r1_r2, row-major
        for (int j = 0; j < 4; ++j)
            for (int i = 0; i < 4; ++i)
                B[4 * i + j] = t1.inv[4 * i] * cx.X12[4 * j] + t1.inv[4 * i + 1] * cx.X12[4 * j + 1] + t1.inv[4 * i + 2] * cx.X12[4 * j + 2] + t1.inv[4 * i + 3] * cx.X12[4 * j + 3];
        for (int j = 0; j < 4; ++j) {
            const T p1 = E1e[0] * B[j] + E1e[1] * B[4 + j] + E1e[2] * B[8 + j] + E1e[3] * B[12 + j];
            const double p2 = E2e[0] * t2.inv[j] + E2e[1] * t2.inv[4 + j] + E2e[2] * t2.inv[8 + j] + E2e[3] * t2.inv[12 + j];
            plane[j] = p2 - p1;
        }
    }
    X3<T> v[4], poly[4];
    for (int k = 0; k < 4; ++k)   // columns of x_r2_zeta1 = x_r2_r1 * [V1; 1]
        for (int i = 0; i < 3; ++i) v[k][i] = cx.X21[i] * t1.v[3 * k] + cx.X21[4 + i] * t1.v[3 * k + 1] + cx.X21[8 + i] * t1.v[3 * k + 2] + cx.X21[12 + i] * 1.0;
    const int n0 = plane_tet_exact(plane, v, poly);
    if (n0 < 3) return 0;
    X4<T> z[8];
    for (int k = 0; k < n0; ++k)   // one_pad_then_mul(x_zeta2_r2, .) then zero_small_coordinates
        for (int i = 0; i < 4; ++i) {
            const T zz = t2.inv[4 * i] * poly[k][0] + t2.inv[4 * i + 1] * poly[k][1] + t2.inv[4 * i + 2] * poly[k][2] + t2.inv[4 * i + 3] * T(1.0);
            z[k][i] = zz * ((1.0e-14 < fabs(xval(zz))) ? 1.0 : 0.0);
        }
    const int n = clip_tet_exact(z, n0, flags);
    if (n < 3) return 0;
    {
        const X3<T> pn = x3<T>(plane[0], plane[1], plane[2]);
        n2 = xdivide(pn, xsqrt(xdot(pn, pn)));
    }
    return polygon_points(z, n, t2, n2, cx, out);
}

// ---- friction laws (friction.jl:2-48) -----------------------------------------------------------------------------------------------
template <class T> PFC_XHD T clamped_piecewise_exact(const T& x, double x_1, double x_2, double y_1, double y_2) {
    const double k = (y_2 - y_1) / (x_2 - x_1);
    const T y = y_1 + (x - x_1) * k;
    return xclamp(y, y_2, y_1);
}
template <class T> PFC_XHD X3<T> vec_sub_vec_proj(const X3<T>& v, const X3<T>& n) {
    const T t = -xdot(v, n);
    return x3<T>(xmuladd(t, n[0], v[0]), xmuladd(t, n[1], v[1]), xmuladd(t, n[2], v[2]));
}
struct BristleP { double tau, k_bar, mu_s, mu_d, Ts_mu_s, Ts_mu_d, magic; };
template <class T> PFC_XHD X3<T> traction_bristle(const BristleP& bf, const X3<T>& Ts, const T& p_dA) {
    const T mag2 = xdot(Ts, Ts);
    X3<T> Tc;
    if (xval(mag2) < bf.mu_s * bf.mu_s) Tc = Ts;
    else {
        const T mag = xsqrt(mag2);
        const T mu = clamped_piecewise_exact(mag, bf.Ts_mu_s, bf.Ts_mu_d, bf.mu_s, bf.mu_d);
        Tc = xdivide(x3<T>(mu * Ts[0], mu * Ts[1], mu * Ts[2]), mag);
    }
    return xscale(Tc, p_dA);
}
template <class T> PFC_XHD X3<T> spatial_vel(const T* v6, const X3<T>& r) { return x3<T>(v6[3], v6[4], v6[5]) + xcross(x3<T>(v6[0], v6[1], v6[2]), r); }

// ---- per-point terms of the three passes; slot order = the order the reference adds them in --------------------------------------
// pass 1, normal_wrench_cop (normal.jl:17-34): lin(3) ang(3) int_p_dA int_p_dA_cop(3)
template <class T> PFC_XHD void terms_cop(const X3<T>& n, const X3<T>& r, const T& dA, const T& p, T* t) {
    const T p_dA = p * dA;
    const X3<T> lam = xscale(n, p_dA);
    const X3<T> m = xcross(r, lam);
    const X3<T> pr = xscale(r, p_dA);
    t[0] = lam[0]; t[1] = lam[1]; t[2] = lam[2]; t[3] = m[0]; t[4] = m[1]; t[5] = m[2]; t[6] = p_dA; t[7] = pr[0]; t[8] = pr[1]; t[9] = pr[2];
}
// pass 2, calc_patch_spatial_stiffness! (friction.jl:147-169): K22 (9, ADDED), K12 (9, ADDED), K11 (9, SUBTRACTED), all row-major
template <class T> PFC_XHD void terms_stiffness(const X3<T>& n, const X3<T>& r_cart, const T& dA, const T& p, const X3<T>& cop, T* t) {
    const T p_dA = p * dA;
    const X3<T> r = r_cart - cop;
    const X3<T> rxn = xcross(r, n);
    const T zero = T(0.0);
    const T skew[9] = {zero, -r[2], r[1], r[2], zero, -r[0], -r[1], r[0], zero};   // vector_to_skew_symmetric (RigidBodyDynamics.Spatial)
    const T a0 = r[0] * r[0], a1 = r[1] * r[1], a2 = r[2] * r[2];
    const T b12 = r[0] * r[1], b13 = r[0] * r[2], b23 = r[1] * r[2];
    const T sk2[9] = {-a1 - a2, b12, b13, b12, -a0 - a2, b23, b13, b23, -a0 - a1};   // vector_to_skew_symmetric_squared
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            const T eye = T(i == j ? 1.0 : 0.0);
            t[3 * i + j] = p_dA * (eye - n[i] * n[j]);
            t[9 + 3 * i + j] = p_dA * (skew[3 * i + j] - rxn[i] * n[j]);
            t[18 + 3 * i + j] = p_dA * (sk2[3 * i + j] + rxn[i] * rxn[j]);
        }
}
// pass 3, calc_spatial_bristle_force (friction.jl:171-201): lin(3) ang(3) about the centre of pressure
template <class T> PFC_XHD void terms_friction(const BristleP& bf, const X3<T>& n, const X3<T>& r, const T& dA, const T& p, const X3<T>& cop, const T* Delta2,
                                             const T* twist, T* t) {
    const X3<T> x2 = r - cop;
    const X3<T> delta2 = spatial_vel(Delta2, x2);
    const X3<T> rdot = spatial_vel(twist, r);
    const T p_dA = p * dA;
    X3<T> Ts = xscale(delta2 + xscale(rdot, bf.tau), -bf.k_bar);
    Ts = vec_sub_vec_proj(Ts, n);
    const X3<T> Tc = traction_bristle(bf, Ts, p_dA);
    const X3<T> m = xcross(x2, Tc);
    t[0] = Tc[0]; t[1] = Tc[1]; t[2] = Tc[2]; t[3] = m[0]; t[4] = m[1]; t[5] = m[2];
}

// ---- warp-cooperative execution of the patch-level steps ---------------------------------------------------------------------------
// The 6 x 6 eigen-decomposition is a few thousand dependent FP64 operations: on one lane it was most of a bristle evaluation.  The loops
// over independent matrix elements are dealt to the lanes of the warp that owns the patch (`for (k = co.lane; k < 6; k += co.n)`),
// which changes who computes an element, never how: the host check compiles the same source with one "lane" doing every index.
#ifdef PFC_HOST_CHECK
struct Coop { static constexpr int lane = 0; static constexpr int n = 1; void sync() const {} };
#else
struct Coop { int lane; static constexpr int n = 32; __device__ void sync() const { __syncwarp(); } };
#endif
template <class T> struct PatchScratch {   // shared by the lanes (shared memory on the device)
    double A[6][6], V[6][6], lam[6], g[6], f[6], fp[6], floor_;
    int clamped[6], m;
    T Kbar[6][6];
};

// 6 x 6 symmetric eigen-decomposition: cyclic Jacobi, run to convergence
// (the reference calls LAPACK / GenericLinearAlgebra, neither under /root/reference; only V f(L) V' is consumed, friction.jl:85-96)
template <class C> PFC_XD_ void jacobi6_exact(double (*A)[6], double (*V)[6], double* lam, const C& co) {
    for (int e = co.lane; e < 36; e += C::n) V[e / 6][e % 6] = (e / 6 == e % 6 ? 1.0 : 0.0);
    co.sync();
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = 0.0, diag = 0.0;   // (every lane: same operands, same order, same bits)
        for (int i = 0; i < 6; ++i) { diag += A[i][i] * A[i][i]; for (int j = i + 1; j < 6; ++j) off += A[i][j] * A[i][j]; }
        if (off == 0.0 || off <= 1.0e-44 * diag) break;
        for (int p = 0; p < 5; ++p)
            for (int q = p + 1; q < 6; ++q) {
                const double apq = A[p][q];
                if (apq == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
                const double t = (theta >= 0.0) ? 1.0 / (theta + sqrt(1.0 + theta * theta)) : -1.0 / (-theta + sqrt(1.0 + theta * theta));
                const double c = 1.0 / sqrt(1.0 + t * t);
                const double s = t * c;
                co.sync();   // every lane has read A[p][q], A[p][p], A[q][q]
                for (int k = co.lane; k < 6; k += C::n) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
                co.sync();
                for (int k = co.lane; k < 6; k += C::n) {
                    const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk;
                    const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
                }
                co.sync();
            }
    }
    for (int i = co.lane; i < 6; i += C::n) lam[i] = A[i][i];
    co.sync();
}

// calc_K̄_sqrt_inv (friction.jl:85-96).  scr.Kbar: upper triangle read (Hermitian wrapper).  Float64.
template <class C> PFC_XD_ void kbar_sqrt_inv(PatchScratch<double>& scr, double (*out)[6], const C& co) {
    for (int e = co.lane; e < 36; e += C::n) { const int i = e / 6, j = e % 6; scr.A[i][j] = (i <= j) ? scr.Kbar[i][j] : scr.Kbar[j][i]; }
    co.sync();
    jacobi6_exact(scr.A, scr.V, scr.lam, co);
    if (co.lane == 0) {
        double max_sig = scr.lam[0];
        for (int k = 1; k < 6; ++k) max_sig = fmax(max_sig, scr.lam[k]);
        for (int k = 0; k < 6; ++k) scr.f[k] = 1.0 / sqrt(xmax(scr.lam[k], max_sig * 1.0e-16));
    }
    co.sync();
    for (int e = co.lane; e < 36; e += C::n) {
        const int i = e / 6, j = e % 6;
        double acc = 0.0;
        for (int k = 0; k < 6; ++k) acc += (scr.V[i][k] * scr.f[k]) * scr.V[j][k];
        out[i][j] = acc;
    }
    co.sync();
}
// Dual mode: the partials of V f(L) V' are the first-order perturbation of that matrix function (Daleckii-Krein), which is what
// differentiating through a converged generic eigen-solver yields wherever the result is differentiable.  One lane per partial.
template <int N, class C> PFC_XD_ void kbar_sqrt_inv(PatchScratch<XD<N>>& scr, XD<N> (*out)[6], const C& co) {
    for (int e = co.lane; e < 36; e += C::n) { const int i = e / 6, j = e % 6; scr.A[i][j] = (i <= j) ? scr.Kbar[i][j].v : scr.Kbar[j][i].v; }
    co.sync();
    jacobi6_exact(scr.A, scr.V, scr.lam, co);
    if (co.lane == 0) {
        int m = 0;
        for (int k = 1; k < 6; ++k) if (scr.lam[k] > scr.lam[m]) m = k;
        scr.m = m;
        scr.floor_ = scr.lam[m] * 1.0e-16;
        for (int k = 0; k < 6; ++k) {
            scr.clamped[k] = !(scr.floor_ < scr.lam[k]);
            scr.g[k] = scr.clamped[k] ? scr.floor_ : scr.lam[k];
            scr.f[k] = 1.0 / sqrt(scr.g[k]);
            scr.fp[k] = -0.5 * scr.f[k] / scr.g[k];
        }
    }
    co.sync();
    for (int e = co.lane; e < 36; e += C::n) {
        const int i = e / 6, j = e % 6;
        double acc = 0.0;
        for (int k = 0; k < 6; ++k) acc += (scr.V[i][k] * scr.f[k]) * scr.V[j][k];
        out[i][j] = XD<N>(acc);
    }
    co.sync();
    const double (*V)[6] = scr.V;
    const double* lam = scr.lam; const double* f = scr.f; const double* fp = scr.fp; const int* clamped = scr.clamped; const int m = scr.m;
    for (int d = co.lane; d < N; d += C::n) {
        double dA[6][6], tmp[6][6], B[6][6], G[6][6];
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) dA[i][j] = (i <= j) ? scr.Kbar[i][j].p[d] : scr.Kbar[j][i].p[d];
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double a = 0; for (int k = 0; k < 6; ++k) a += dA[i][k] * V[k][j]; tmp[i][j] = a; }
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double a = 0; for (int k = 0; k < 6; ++k) a += V[k][i] * tmp[k][j]; B[i][j] = a; }
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) {
                if (i == j) { G[i][j] = fp[i] * (clamped[i] ? 1.0e-16 * B[m][m] : B[i][i]); continue; }
                double F;
                if (clamped[i] && clamped[j]) F = 0.0;
                else if (!clamped[i] && !clamped[j]) { const double si = sqrt(lam[i]), sj = sqrt(lam[j]); F = -1.0 / (si * sj * (si + sj)); }
                else F = (lam[i] == lam[j]) ? 0.0 : (f[i] - f[j]) / (lam[i] - lam[j]);
                G[i][j] = F * B[i][j];
            }
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double a = 0; for (int k = 0; k < 6; ++k) a += V[i][k] * G[k][j]; tmp[i][j] = a; }
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double a = 0; for (int k = 0; k < 6; ++k) a += tmp[i][k] * V[j][k]; out[i][j].p[d] = a; }
    }
    co.sync();
}

// The patch-level part of yes_contact!(::Bristle) (friction.jl:119-143) between the passes; called by every lane of the patch's warp.
//   after pass 2:  K (from the 27 sums) -> decompose_K! -> Sinv, Kh = K̄^(-1/2), Delta2 = Sinv .* (Kh * s)
template <class T, class C> PFC_XD_ void bristle_after_stiffness(const T* sum27, const BristleP& bf, const T* s, T* Sinv, T (*Kh)[6], T* Delta2, PatchScratch<T>& scr,
                                                                const C& co) {
    if (co.lane == 0) {
        T K[6][6];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                K[i][j] = sum27[18 + 3 * i + j] * bf.k_bar;          // K11
                K[3 + i][j] = sum27[9 + 3 * j + i] * bf.k_bar;       // K12'
                K[i][3 + j] = sum27[9 + 3 * i + j] * bf.k_bar;       // K12
                K[3 + i][3 + j] = sum27[3 * i + j] * bf.k_bar;       // K22
            }
        T Kf[6][6];   // Hermitian wrapper: the upper triangle is what is read
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) Kf[i][j] = (i <= j) ? K[i][j] : K[j][i];
        const T t_1 = Kf[0][0] + Kf[1][1] + Kf[2][2];
        const T t_2 = Kf[3][3] + Kf[4][4] + Kf[5][5];
        const T s_1 = 1.0 / xsqrt(t_1);
        for (int k = 0; k < 3; ++k) Sinv[k] = s_1 * bf.magic;
        const T s_2 = 1.0 / xsqrt(t_2);
        for (int k = 3; k < 6; ++k) Sinv[k] = s_2;
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) scr.Kbar[i][j] = Sinv[i] * Kf[i][j] * Sinv[j];
    }
    co.sync();
    kbar_sqrt_inv(scr, Kh, co);
    for (int i = co.lane; i < 6; i += C::n) {
        T acc = T(0.0);
        for (int j = 0; j < 6; ++j) acc = acc + Kh[i][j] * s[j];
        Delta2[i] = Sinv[i] * acc;
    }
    co.sync();
}
//   after pass 3: friction wrench about the cop (lin, ang sums) -> wrench about the r2 origin, s-dot
template <class T> __device__ __noinline__ void bristle_finish(const T* sum10, const T* sum6, const X3<T>& cop, const BristleP& bf, const T* s, const T* Sinv,
                                                             const T (*Kh)[6], T* wrench, T* sdot) {
    const X3<T> lin = x3<T>(sum6[0], sum6[1], sum6[2]), ang = x3<T>(sum6[3], sum6[4], sum6[5]);
    const X3<T> ang2 = ang + xcross(cop, lin);
    const T w_cop[6] = {ang[0], ang[1], ang[2], lin[0], lin[1], lin[2]};
    const double tau_inv = 1 / bf.tau;
    T sw[6];
    for (int i = 0; i < 6; ++i) sw[i] = Sinv[i] * w_cop[i];
    for (int i = 0; i < 6; ++i) {
        T acc = T(0.0);
        for (int j = 0; j < 6; ++j) acc = acc + Kh[i][j] * sw[j];
        sdot[i] = (-tau_inv) * (acc + s[i]);
    }
    for (int k = 0; k < 3; ++k) { wrench[k] = sum10[3 + k] + ang2[k]; wrench[3 + k] = sum10[k] + lin[k]; }
}

}  // namespace ex
}  // namespace pfc
