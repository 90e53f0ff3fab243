// pfc_oracle.hpp -- CPU restatement of PressureFieldContact.jl's contact-wrench evaluation.
//
// THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may build, link or call it.  The product
// (pressurefieldcontact.jl_b200/csrc) shares no code with this file.
//
// Parity status: the reference is Julia and Julia is not installed here, so the reference
// itself cannot be run.  The oracle is pinned by the reference's own analytic / known-answer
// tests (test/test_normal.jl, test/test_friction.jl, test/test_clip/*.jl, test/test_obb/
// test_intersection.jl) restated in tests/test_oracle_*.py.  No stored golden vectors exist in
// the reference.
//
// Style: a literal, recursive, scalar restatement templated on the scalar type T (double or
// Dual<N>), citing the reference file:line each function follows (paths relative to
// /root/reference).  Compile with -ffp-contract=off: Julia never contracts a*b+c unless the
// source says muladd, and where it does say muladd we call std::fma explicitly.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <type_traits>
#include <string>
#include <utility>
#include <vector>

namespace orc {

// ----------------------------------------------------------------------------------------------
// Dual numbers (ForwardDiff.Dual{Nothing,Float64,N}; ForwardDiff 0.10.3, Manifest.toml:129-133)
// ----------------------------------------------------------------------------------------------
template <int N>
struct Dual {
    double v;
    double p[N];
    Dual() : v(0.0) { for (int i = 0; i < N; ++i) p[i] = 0.0; }
    Dual(double x) : v(x) { for (int i = 0; i < N; ++i) p[i] = 0.0; }
};

inline double value(double x) { return x; }
template <int N> inline double value(const Dual<N>& x) { return x.v; }

template <int N> inline Dual<N> operator-(const Dual<N>& a) { Dual<N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.p[i] = -a.p[i]; return r; }
template <int N> inline Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.p[i] = a.p[i] + b.p[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.p[i] = a.p[i] - b.p[i]; return r; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.p[i] = a.p[i] * b.v + a.v * b.p[i]; return r; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; r.v = a.v / b.v;
    for (int i = 0; i < N; ++i) r.p[i] = (a.p[i] - r.v * b.p[i]) / b.v;
    return r;
}
template <int N> inline Dual<N> operator+(const Dual<N>& a, double b) { Dual<N> r = a; r.v = a.v + b; return r; }
template <int N> inline Dual<N> operator+(double a, const Dual<N>& b) { Dual<N> r = b; r.v = a + b.v; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, double b) { Dual<N> r = a; r.v = a.v - b; return r; }
template <int N> inline Dual<N> operator-(double a, const Dual<N>& b) { Dual<N> r = -b; r.v = a - b.v; return r; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, double b) { Dual<N> r; r.v = a.v * b; for (int i = 0; i < N; ++i) r.p[i] = a.p[i] * b; return r; }
template <int N> inline Dual<N> operator*(double a, const Dual<N>& b) { return b * a; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, double b) { Dual<N> r; r.v = a.v / b; for (int i = 0; i < N; ++i) r.p[i] = a.p[i] / b; return r; }
template <int N> inline Dual<N> operator/(double a, const Dual<N>& b) { return Dual<N>(a) / b; }
template <int N> inline Dual<N>& operator+=(Dual<N>& a, const Dual<N>& b) { a = a + b; return a; }
template <int N> inline Dual<N>& operator-=(Dual<N>& a, const Dual<N>& b) { a = a - b; return a; }

inline double sqrt_(double x) { return std::sqrt(x); }
template <int N> inline Dual<N> sqrt_(const Dual<N>& a) {
    Dual<N> r; r.v = std::sqrt(a.v);
    for (int i = 0; i < N; ++i) r.p[i] = a.p[i] / (2.0 * r.v);
    return r;
}
inline double muladd_(double a, double b, double c) { return std::fma(a, b, c); }
template <int N> inline Dual<N> muladd_(const Dual<N>& a, const Dual<N>& b, const Dual<N>& c) {
    Dual<N> r; r.v = std::fma(a.v, b.v, c.v);
    for (int i = 0; i < N; ++i) r.p[i] = a.p[i] * b.v + a.v * b.p[i] + c.p[i];
    return r;
}
template <int N> inline Dual<N> muladd_(double a, const Dual<N>& b, double c) {
    Dual<N> r; r.v = std::fma(a, b.v, c);
    for (int i = 0; i < N; ++i) r.p[i] = a * b.p[i];
    return r;
}
template <int N> inline Dual<N> muladd_(double a, const Dual<N>& b, const Dual<N>& c) {
    Dual<N> r; r.v = std::fma(a, b.v, c.v);
    for (int i = 0; i < N; ++i) r.p[i] = a * b.p[i] + c.p[i];
    return r;
}
// Julia: max(x, y) = ifelse(y < x, x, y) compared on values
template <class T> inline T max_(const T& x, const T& y) { return (value(y) < value(x)) ? x : y; }
// Julia: clamp(x, lo, hi) = ifelse(x > hi, hi, ifelse(x < lo, lo, x)) (constants lose partials)
template <class T> inline T clamp_(const T& x, double lo, double hi) {
    if (value(x) > hi) return T(hi);
    if (value(x) < lo) return T(lo);
    return x;
}

// ----------------------------------------------------------------------------------------------
// Counted: a double that counts floating-point operations (+ - * / sqrt = 1, muladd = 2;
// negation, abs, comparisons and conversions are free).  Used by orc_count_flops to publish the
// ALGORITHMIC FLOP count of the reference algorithm per evaluation (SURVEY.md section 8d asks for
// the constants to be re-counted from the oracle with an instrumented scalar type).
// ----------------------------------------------------------------------------------------------
inline int64_t& flop_counter() { static thread_local int64_t c = 0; return c; }
inline void count_flops(int64_t n) { flop_counter() += n; }
struct Counted {
    double v;
    Counted() : v(0.0) {}
    Counted(double x) : v(x) {}
};
inline double value(const Counted& x) { return x.v; }
inline Counted operator-(const Counted& a) { return Counted(-a.v); }
#define ORC_COUNTED_BINOP(op)                                                                              \
    inline Counted operator op(const Counted& a, const Counted& b) { count_flops(1); return Counted(a.v op b.v); } \
    inline Counted operator op(const Counted& a, double b) { count_flops(1); return Counted(a.v op b); }           \
    inline Counted operator op(double a, const Counted& b) { count_flops(1); return Counted(a op b.v); }
ORC_COUNTED_BINOP(+)
ORC_COUNTED_BINOP(-)
ORC_COUNTED_BINOP(*)
ORC_COUNTED_BINOP(/)
#undef ORC_COUNTED_BINOP
inline Counted sqrt_(const Counted& a) { count_flops(1); return Counted(std::sqrt(a.v)); }
inline Counted muladd_(const Counted& a, const Counted& b, const Counted& c) { count_flops(2); return Counted(std::fma(a.v, b.v, c.v)); }
inline Counted muladd_(double a, const Counted& b, double c) { count_flops(2); return Counted(std::fma(a, b.v, c)); }
inline Counted muladd_(double a, const Counted& b, const Counted& c) { count_flops(2); return Counted(std::fma(a, b.v, c.v)); }
inline double abs_(double x) { return std::fabs(x); }
inline Counted abs_(const Counted& x) { return Counted(std::fabs(x.v)); }

// ----------------------------------------------------------------------------------------------
// Small static vectors / matrices (StaticArrays 0.10.3 semantics: column-major, products are
// unrolled row-by-column sums evaluated left to right, never fused)
// ----------------------------------------------------------------------------------------------
template <class T> struct V3 { T x[3]; T& operator[](int i) { return x[i]; } const T& operator[](int i) const { return x[i]; } };
template <class T> struct V4 { T x[4]; T& operator[](int i) { return x[i]; } const T& operator[](int i) const { return x[i]; } };
template <class T> struct V6 { T x[6]; T& operator[](int i) { return x[i]; } const T& operator[](int i) const { return x[i]; } };
template <class T> struct M3 { T m[9];  T& operator()(int r, int c) { return m[c * 3 + r]; } const T& operator()(int r, int c) const { return m[c * 3 + r]; } };
template <class T> struct M4 { T m[16]; T& operator()(int r, int c) { return m[c * 4 + r]; } const T& operator()(int r, int c) const { return m[c * 4 + r]; } };

template <class T> inline V3<T> mk3(const T& a, const T& b, const T& c) { V3<T> r; r[0] = a; r[1] = b; r[2] = c; return r; }
template <class T> inline V3<T> zero3() { return mk3<T>(T(0.0), T(0.0), T(0.0)); }
template <class T> inline V3<T> operator+(const V3<T>& a, const V3<T>& b) { return mk3<T>(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
template <class T> inline V3<T> operator-(const V3<T>& a, const V3<T>& b) { return mk3<T>(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
template <class T> inline V3<T> operator-(const V3<T>& a) { return mk3<T>(-a[0], -a[1], -a[2]); }
template <class T, class S> inline V3<T> scale(const V3<T>& a, const S& s) { return mk3<T>(a[0] * s, a[1] * s, a[2] * s); }
template <class T, class S> inline V3<T> divide(const V3<T>& a, const S& s) { return mk3<T>(a[0] / s, a[1] / s, a[2] / s); }
template <class T> inline T dot(const V3<T>& a, const V3<T>& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <class T> inline V3<T> cross(const V3<T>& a, const V3<T>& b) {
    return mk3<T>(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
template <class T> inline V3<T> normalize(const V3<T>& a) { T n = sqrt_(dot(a, a)); return divide(a, n); }
template <class T> inline V3<T> lift3(const V3<double>& a) { return mk3<T>(T(a[0]), T(a[1]), T(a[2])); }

template <class TA, class TB, class TR> inline M4<TR> mul44(const M4<TA>& a, const M4<TB>& b) {
    M4<TR> r;
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i)
            r(i, j) = a(i, 0) * b(0, j) + a(i, 1) * b(1, j) + a(i, 2) * b(2, j) + a(i, 3) * b(3, j);
    return r;
}
template <class TA, class TB, class TR> inline V4<TR> mul4v(const M4<TA>& a, const V4<TB>& b) {
    V4<TR> r;
    for (int i = 0; i < 4; ++i) r[i] = a(i, 0) * b[0] + a(i, 1) * b[1] + a(i, 2) * b[2] + a(i, 3) * b[3];
    return r;
}
// 1x4 row times 4x4
template <class TA, class TB, class TR> inline V4<TR> mulrow4(const V4<TA>& a, const M4<TB>& b) {
    V4<TR> r;
    for (int j = 0; j < 4; ++j) r[j] = a[0] * b(0, j) + a[1] * b(1, j) + a[2] * b(2, j) + a[3] * b(3, j);
    return r;
}
template <class T> inline V3<T> mul3v(const M3<T>& a, const V3<T>& b) {
    V3<T> r;
    for (int i = 0; i < 3; ++i) r[i] = a(i, 0) * b[0] + a(i, 1) * b[1] + a(i, 2) * b[2];
    return r;
}
template <class T> inline V3<T> mul3tv(const M3<T>& a, const V3<T>& b) {  // transpose(a) * b
    V3<T> r;
    for (int i = 0; i < 3; ++i) r[i] = a(0, i) * b[0] + a(1, i) * b[1] + a(2, i) * b[2];
    return r;
}

// inv(::SMatrix{4,4}) -- StaticArrays is not vendored under /root/reference; the published
// algorithm for 4x4 is the explicit adjugate / determinant formula.  Parity on this routine is
// tolerance-level only.
inline M4<double> inv44(const M4<double>& A) {
    const double* m = A.m;
    double inv[16];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    count_flops(296);  // 16 cofactors x 17 + determinant 7 + 1 division + 16 scalings
    double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    double idet = 1.0 / det;
    M4<double> R;
    for (int i = 0; i < 16; ++i) R.m[i] = inv[i] * idet;
    return R;
}

// ----------------------------------------------------------------------------------------------
// MathKernel  (src/math_kernel/*.jl)
// ----------------------------------------------------------------------------------------------
// src/math_kernel/utility.jl:21-26
template <int NN, class T> struct VN { T x[NN]; };
template <class T> inline V4<T> weightPoly(const V4<T>& p1, const V4<T>& p2, const T& w1, const T& w2) {
    T sum_weight = w1 - w2;
    T c1 = w1 / sum_weight;
    T c2 = w2 / sum_weight;
    V4<T> r;
    for (int i = 0; i < 4; ++i) r[i] = c1 * p2[i] - c2 * p1[i];
    return r;
}
template <class T> inline V3<T> weightPoly(const V3<T>& p1, const V3<T>& p2, const T& w1, const T& w2) {
    T sum_weight = w1 - w2;
    T c1 = w1 / sum_weight;
    T c2 = w2 / sum_weight;
    V3<T> r;
    for (int i = 0; i < 3; ++i) r[i] = c1 * p2[i] - c2 * p1[i];
    return r;
}
// src/math_kernel/vector_projections.jl:2-7
template <class T> inline V3<T> vec_sub_vec_proj(const V3<T>& v, const V3<T>& n) {
    T t = -dot(v, n);
    return mk3<T>(muladd_(t, n[0], v[0]), muladd_(t, n[1], v[1]), muladd_(t, n[2], v[2]));
}
// src/math_kernel/vector_projections.jl:9-13   (a is Float64 1x4, b is T)
template <class T> inline T a_dot_one_pad_b(const V4<double>& a, const V3<T>& b) {
    T d = muladd_(a[0], b[0], a[3]);
    d = muladd_(a[1], b[1], d);
    return muladd_(a[2], b[2], d);
}
// src/math_kernel/geometry_kernel.jl:3-10
template <class T> inline V3<T> centroid3(const V3<T>& a, const V3<T>& b, const V3<T>& c) { return scale(a + b + c, double(1.0 / 3.0)); }
template <class T> inline V3<T> vector_area(const V3<T>& v1, const V3<T>& v2, const V3<T>& v3) { return scale(cross(v2 - v1, v3 - v2), 0.5); }
template <class T> inline T triangle_area(const V3<T>& v1, const V3<T>& v2, const V3<T>& v3, const V3<T>& n) { return dot(n, vector_area(v1, v2, v3)); }
template <class T> inline V3<T> triangleNormal(const V3<T>& v1, const V3<T>& v2, const V3<T>& v3) { return normalize(vector_area(v1, v2, v3)); }
// src/math_kernel/geometry_kernel.jl:27-39
inline double tet_volume(const V3<double>& v1, const V3<double>& v2, const V3<double>& v3, const V3<double>& v4) {
    double a1 = v1[0], a2 = v1[1], a3 = v1[2], b1 = v2[0], b2 = v2[1], b3 = v2[2];
    double c1 = v3[0], c2 = v3[1], c3 = v3[2], d1 = v4[0], d2 = v4[1], d3 = v4[2];
    double V = (b1 - a1) * (c2 * d3 - c3 * d2);
    V = std::fma(b2 - a2, c3 * d1 - c1 * d3, V);
    V = std::fma(b3 - a3, c1 * d2 - c2 * d1, V);
    V = std::fma(c1 - d1, a3 * b2 - a2 * b3, V);
    V = std::fma(c2 - d2, a1 * b3 - a3 * b1, V);
    V = std::fma(c3 - d3, a2 * b1 - a1 * b2, V);
    return V * double(1.0 / 6.0);
}

// ----------------------------------------------------------------------------------------------
// Clip  (src/clip/*.jl)
// ----------------------------------------------------------------------------------------------
// src/clip/poly_eight.jl:8-27
template <int NN, class T> struct Poly;
template <class T> struct Poly<4, T> { int n = 0; V4<T> v[8]; };
template <class T> struct Poly<3, T> { int n = 0; V3<T> v[8]; };

struct ClipStatus { bool non_finite = false; bool bad_arity = false; };

template <class T> Poly<4, T> clip(const V4<T>* z, int n, int i, ClipStatus& st);

// src/clip/static_clip.jl:197-201
template <class T> inline V4<T> clip_node(const V4<T>& z_non, const V4<T>& z_pos, int i) {
    return weightPoly(z_non, z_pos, z_non[i], z_pos[i]);
}

// src/clip/static_clip.jl:135-195 (the five arities folded into one recursion on n; the
// comparison asymmetry -- `0.0 < z_last` for 3,4,5 vertices and `0.0 <= z_last` for 6,7 -- and
// the direct return of the 7-vertex cut are preserved)
template <class T> Poly<4, T> cut_clip(const V4<T>* z, int n, int i, ClipStatus& st) {
    if (n > 3 && value(z[n - 2][i]) <= 0.0) return cut_clip(z, n - 1, i, st);
    V4<T> z_start = clip_node(z[0], z[1], i);
    bool last_inside = (n <= 5) ? (0.0 < value(z[n - 1][i])) : (0.0 <= value(z[n - 1][i]));
    V4<T> out[8];
    int m;
    if (last_inside) {
        V4<T> z_end = clip_node(z[0], z[n - 1], i);
        out[0] = z_start;
        for (int k = 1; k < n; ++k) out[k] = z[k];
        out[n] = z_end;
        m = n + 1;
    } else {
        V4<T> z_end = clip_node(z[n - 1], z[n - 2], i);
        out[0] = z_start;
        for (int k = 1; k < n - 1; ++k) out[k] = z[k];
        out[n - 1] = z_end;
        m = n;
    }
    if (n == 7) {  // static_clip.jl:185-195: returned as-is, remaining planes (if any) are not applied
        Poly<4, T> p;
        p.n = m;
        for (int k = 0; k < m; ++k) p.v[k] = out[k];
        for (int k = m; k < 8; ++k) p.v[k] = out[0];
        return p;
    }
    return clip(out, m, i + 1, st);
}

// src/clip/static_clip.jl:34-128
template <class T> Poly<4, T> clip(const V4<T>* z, int n, int i, ClipStatus& st) {
    if (i == 4) {  // "there is no 5th plane" (i is 0-based here)
        Poly<4, T> p;
        p.n = n;
        for (int k = 0; k < n; ++k) p.v[k] = z[k];
        for (int k = n; k < 8; ++k) p.v[k] = z[0];
        return p;
    }
    bool is_non_pos[8];
    bool all_non_pos = true, all_non_neg = true;
    for (int k = 0; k < n; ++k) {
        double s = value(z[k][i]);
        is_non_pos[k] = (s <= 0.0);
        all_non_pos = all_non_pos && is_non_pos[k];
        all_non_neg = all_non_neg && (0.0 <= s);
    }
    if (all_non_pos) return Poly<4, T>();
    if (all_non_neg) return clip(z, n, i + 1, st);
    for (int k = 0; k < n; ++k) {
        int k1 = (k + 1) % n;
        if (is_non_pos[k] && !is_non_pos[k1]) {
            V4<T> rot[8];
            for (int j = 0; j < n; ++j) rot[j] = z[(k + j) % n];
            return cut_clip(rot, n, i, st);
        }
    }
    st.non_finite = true;  // error("Non-finite vertex likely")
    return Poly<4, T>();
}

// src/clip/static_clip.jl:7-23
template <class T> Poly<4, T> clip_in_tet_coordinates(const Poly<4, T>& p, ClipStatus& st) {
    if (p.n == 3 || p.n == 4) return clip(p.v, p.n, 0, st);
    st.bad_arity = true;  // error("something is wrong")
    return Poly<4, T>();
}
template <class T> Poly<4, T> clip_in_tet_coordinates(const V4<T>& z1, const V4<T>& z2, const V4<T>& z3, ClipStatus& st) {
    V4<T> z[3] = {z1, z2, z3};
    return clip(z, 3, 0, st);
}

// src/clip/plane_tet_intersection.jl:9-106
template <class T> Poly<3, T> clip_plane_tet(const V4<T>& plane, const M4<T>& tet) {
    V3<T> v[4];
    for (int k = 0; k < 4; ++k) v[k] = mk3<T>(tet(0, k), tet(1, k), tet(2, k));
    V4<T> proj = mulrow4<T, T, T>(plane, tet);
    bool bool_neg[4], bool_pos[4];
    int n_neg = 0, n_pos = 0;
    for (int k = 0; k < 4; ++k) {
        bool_neg[k] = value(proj[k]) < 0.0; n_neg += bool_neg[k];
        bool_pos[k] = 0.0 < value(proj[k]); n_pos += bool_pos[k];
    }
    Poly<3, T> out;
    if (n_pos == 0 || n_neg == 0) return out;
    auto pwp = [&](int i1, int i2) { return weightPoly(v[i1 - 1], v[i2 - 1], proj[i1 - 1], proj[i2 - 1]); };
    auto tri = [&](int lone, const V3<T>& a, const V3<T>& b, const V3<T>& c) {
        out.n = 3;
        if (0.0 < value(proj[lone - 1])) { out.v[0] = a; out.v[1] = b; out.v[2] = c; }
        else { out.v[0] = c; out.v[1] = b; out.v[2] = a; }
        for (int k = 3; k < 8; ++k) out.v[k] = out.v[0];
    };
    auto quad = [&](const V3<T>& a, const V3<T>& b, const V3<T>& c, const V3<T>& d) {
        out.n = 4;
        if (0.0 < value(proj[0])) { out.v[0] = a; out.v[1] = b; out.v[2] = c; out.v[3] = d; }
        else { out.v[0] = d; out.v[1] = c; out.v[2] = b; out.v[3] = a; }
        for (int k = 4; k < 8; ++k) out.v[k] = out.v[0];
    };
    auto tet_k = [&](int k) {
        if (k == 1) tri(1, pwp(2, 1), pwp(4, 1), pwp(3, 1));
        if (k == 2) tri(2, pwp(1, 2), pwp(3, 2), pwp(4, 2));
        if (k == 3) tri(3, pwp(1, 3), pwp(4, 3), pwp(2, 3));
        if (k == 4) tri(4, pwp(1, 4), pwp(2, 4), pwp(3, 4));
    };
    if (n_pos == 1) {
        for (int k = 0; k < 4; ++k) if (bool_pos[k]) { tet_k(k + 1); return out; }
    } else if (n_neg == 1) {
        for (int k = 0; k < 4; ++k) if (bool_neg[k]) { tet_k(k + 1); return out; }
    } else {
        if (bool_pos[0] == bool_pos[1]) { quad(pwp(2, 3), pwp(2, 4), pwp(1, 4), pwp(1, 3)); return out; }
        if (bool_pos[0] == bool_pos[2]) { quad(pwp(1, 2), pwp(1, 4), pwp(3, 4), pwp(3, 2)); return out; }
        if (bool_pos[0] == bool_pos[3]) { quad(pwp(1, 3), pwp(1, 2), pwp(4, 2), pwp(4, 3)); return out; }
    }
    return out;
}

// src/clip/poly_eight.jl:35-52
template <class T> std::pair<T, V3<T>> poly_centroid(const Poly<3, T>& p, const V3<T>& n) {
    V3<T> cart_a = p.v[0];
    V3<T> cart_c = p.v[1];
    T cum_sum = T(0.0);
    V3<T> cum_prod = zero3<T>();
    for (int k = 2; k < p.n; ++k) {
        V3<T> cart_b = cart_c;
        cart_c = p.v[k];
        T area_ = triangle_area(cart_a, cart_b, cart_c, n);
        cum_prod = cum_prod + scale(centroid3(cart_a, cart_b, cart_c), area_);
        cum_sum = cum_sum + area_;
    }
    if (value(cum_sum) == 0.0) return {cum_sum, cart_a};
    return {cum_sum, divide(cum_prod, cum_sum)};
}

// src/clip/poly_eight.jl:60-75  (m is Float64, p is T)
template <class T> Poly<4, T> one_pad_then_mul(const M4<double>& m, const Poly<3, T>& p) {
    Poly<4, T> r;
    r.n = p.n;
    int cnt = (p.n <= 4) ? 4 : 8;
    for (int k = 0; k < cnt; ++k) {
        V4<T> o; o[0] = p.v[k][0]; o[1] = p.v[k][1]; o[2] = p.v[k][2]; o[3] = T(1.0);
        r.v[k] = mul4v<double, T, T>(m, o);
    }
    for (int k = cnt; k < 8; ++k) r.v[k] = r.v[0];
    return r;
}
// src/clip/poly_eight.jl:83-98
template <class T> Poly<3, T> mul_then_un_pad(const M4<double>& m, const Poly<4, T>& p) {
    Poly<3, T> r;
    r.n = p.n;
    int cnt = (p.n <= 4) ? 4 : 8;
    for (int k = 0; k < cnt; ++k) {
        V4<T> o = mul4v<double, T, T>(m, p.v[k]);
        r.v[k] = mk3<T>(o[0], o[1], o[2]);
    }
    for (int k = cnt; k < 8; ++k) r.v[k] = r.v[0];
    return r;
}
// src/clip/poly_eight.jl:106-126
template <class T> Poly<4, T> zero_small_coordinates(const Poly<4, T>& p) {
    Poly<4, T> r = p;
    int cnt = (p.n <= 4) ? 4 : 8;
    for (int k = 0; k < cnt; ++k)
        for (int i = 0; i < 4; ++i) {
            bool keep = 1.0e-14 < std::fabs(value(p.v[k][i]));
            r.v[k][i] = p.v[k][i] * (keep ? 1.0 : 0.0);
        }
    for (int k = cnt; k < 8; ++k) r.v[k] = r.v[0];
    return r;
}

// src/clip/quadrature.jl:21-41 (triangle rules 1 and 2; literals as in the reference)
struct TriQuadRule { int n; double zeta[3][3]; double w[3]; };
inline TriQuadRule getTriQuadRule(int n_rule) {
    TriQuadRule q{};
    if (n_rule == 1) {
        q.n = 1;
        q.zeta[0][0] = q.zeta[0][1] = q.zeta[0][2] = 0.33333333333333331483;
        q.w[0] = 1.0;
    } else {
        const double a = 0.16666666666666674068, b = 0.66666666666666651864;
        q.n = 3;
        double z[3][3] = {{a, b, a}, {b, a, a}, {a, a, b}};
        for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) q.zeta[i][j] = z[i][j]; q.w[i] = 0.33333333333333331483; }
    }
    return q;
}

// ----------------------------------------------------------------------------------------------
// Binary_BB_Trees  (src/obb/*.jl)
// ----------------------------------------------------------------------------------------------
struct OBB { V3<double> c, e; M3<double> R; };

// src/obb/bb_intersection.jl:17-74
template <class S> inline bool BB_BB_intersect_sat(const V3<S>& e_a, const V3<S>& e_b, const V3<S>& t, const M3<S>& R, const M3<S>& abs_R) {
    auto any_lt = [](const V3<S>& a, const V3<S>& b) { return (value(a[0]) < value(b[0])) || (value(a[1]) < value(b[1])) || (value(a[2]) < value(b[2])); };
    auto abs3 = [](const V3<S>& a) { return mk3<S>(abs_(a[0]), abs_(a[1]), abs_(a[2])); };
    auto s221 = [](const V3<S>& r) { return mk3<S>(r[2], r[2], r[1]); };
    auto s100 = [](const V3<S>& r) { return mk3<S>(r[1], r[0], r[0]); };
    auto had = [](const V3<S>& a, const V3<S>& b) { return mk3<S>(a[0] * b[0], a[1] * b[1], a[2] * b[2]); };
    V3<S> R0 = mk3<S>(R.m[0], R.m[3], R.m[6]);
    V3<S> R1 = mk3<S>(R.m[1], R.m[4], R.m[7]);
    V3<S> R2 = mk3<S>(R.m[2], R.m[5], R.m[8]);
    V3<S> aR0 = mk3<S>(abs_R.m[0], abs_R.m[3], abs_R.m[6]);
    V3<S> aR1 = mk3<S>(abs_R.m[1], abs_R.m[4], abs_R.m[7]);
    V3<S> aR2 = mk3<S>(abs_R.m[2], abs_R.m[5], abs_R.m[8]);
    // face test 1/2
    V3<S> T_dot_L = abs3(t);
    V3<S> r_a = e_a;
    V3<S> r_b = mul3v(abs_R, e_b);
    if (any_lt(r_a + r_b, T_dot_L)) return false;
    // face test 2/2
    T_dot_L = abs3(mul3tv(R, t));
    r_a = mul3tv(abs_R, e_a);
    r_b = e_b;
    if (any_lt(r_a + r_b, T_dot_L)) return false;
    V3<S> eb_100 = s100(e_b), eb_221 = s221(e_b);
    S t0 = t[0], t1 = t[1], t2 = t[2], ea0 = e_a[0], ea1 = e_a[1], ea2 = e_a[2];
    // cross test 1/3
    T_dot_L = abs3(scale(R1, t2) - scale(R2, t1));
    r_a = scale(aR2, ea1) + scale(aR1, ea2);
    r_b = had(eb_100, s221(aR0)) + had(eb_221, s100(aR0));
    if (any_lt(r_a + r_b, T_dot_L)) return false;
    // cross test 2/3
    T_dot_L = abs3(scale(R2, t0) - scale(R0, t2));
    r_a = scale(aR2, ea0) + scale(aR0, ea2);
    r_b = had(eb_100, s221(aR1)) + had(eb_221, s100(aR1));
    if (any_lt(r_a + r_b, T_dot_L)) return false;
    // cross test 3/3
    T_dot_L = abs3(scale(R0, t1) - scale(R1, t0));
    r_a = scale(aR1, ea0) + scale(aR0, ea1);
    r_b = had(eb_100, s221(aR2)) + had(eb_221, s100(aR2));
    if (any_lt(r_a + r_b, T_dot_L)) return false;
    return true;
}

// basic_dh(R, t): src/math_kernel/basic_dh.jl:52-59
template <class S> inline M4<S> basic_dh(const M3<S>& R, const V3<S>& t) {
    M4<S> m;
    for (int c = 0; c < 3; ++c) { for (int r = 0; r < 3; ++r) m(r, c) = R(r, c); m(3, c) = S(0.0); }
    m(0, 3) = t[0]; m(1, 3) = t[1]; m(2, 3) = t[2]; m(3, 3) = S(1.0);
    return m;
}

// src/obb/bb_intersection.jl:2-12
template <class S> inline bool BB_BB_intersect(const M3<double>& R_a_b_d, const V3<double>& t_a_b_d, const OBB& a_d, const OBB& b_d) {
    struct { V3<S> c, e; M3<S> R; } a, b;
    M3<S> R_a_b; V3<S> t_a_b;
    for (int k = 0; k < 3; ++k) { a.c[k] = S(a_d.c[k]); a.e[k] = S(a_d.e[k]); b.c[k] = S(b_d.c[k]); b.e[k] = S(b_d.e[k]); t_a_b[k] = S(t_a_b_d[k]); }
    for (int k = 0; k < 9; ++k) { a.R.m[k] = S(a_d.R.m[k]); b.R.m[k] = S(b_d.R.m[k]); R_a_b.m[k] = S(R_a_b_d.m[k]); }
    M3<S> aRt;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) aRt(r, c) = a.R(c, r);
    M3<S> naRt;  // -a.R' (unary minus binds first in `-a.R' * a.c`)
    for (int k = 0; k < 9; ++k) naRt.m[k] = -aRt.m[k];
    M4<S> i_dh_a = basic_dh<S>(aRt, mul3v(naRt, a.c));
    M4<S> dh_a_b = basic_dh<S>(R_a_b, t_a_b);
    M4<S> dh_b = basic_dh<S>(b.R, b.c);
    M4<S> dh_final = mul44<S, S, S>(mul44<S, S, S>(i_dh_a, dh_a_b), dh_b);
    M3<S> R_tot, abs_R_tot;
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) R_tot(r, c) = dh_final(r, c);
    V3<S> t = mk3<S>(dh_final(0, 3), dh_final(1, 3), dh_final(2, 3));
    for (int k = 0; k < 9; ++k) abs_R_tot.m[k] = abs_(R_tot.m[k]) + 1.0e-14;
    return BB_BB_intersect_sat<S>(a.e, b.e, t, R_tot, abs_R_tot);
}

// src/obb/util.jl:17-51, src/obb/box_types.jl:11-15
inline void calc_min_max(const OBB& a, V3<double>& lo, V3<double>& hi) {
    M3<double> aR; for (int k = 0; k < 9; ++k) aR.m[k] = std::fabs(a.R.m[k]);
    V3<double> d = mul3v(aR, a.e);
    lo = a.c - d; hi = a.c + d;
}
inline OBB calc_obb(const V3<double>& a, const V3<double>& b) {
    V3<double> lo, hi;
    for (int k = 0; k < 3; ++k) { lo[k] = std::fmin(a[k], b[k]); hi[k] = std::fmax(a[k], b[k]); }
    OBB o;
    o.c = scale(hi + lo, 0.5);
    o.e = scale(hi - lo, 0.5);
    for (int k = 0; k < 9; ++k) o.R.m[k] = (k % 4 == 0) ? 1.0 : 0.0;
    return o;
}
inline OBB calc_obb_points(const V3<double>* v, int n) {
    V3<double> lo = v[0], hi = v[0];
    for (int i = 1; i < n; ++i) for (int k = 0; k < 3; ++k) { lo[k] = std::fmin(lo[k], v[i][k]); hi[k] = std::fmax(hi[k], v[i][k]); }
    return calc_obb(lo, hi);
}
inline OBB merge_obb(const OBB& a, const OBB& b) {  // (::Type{BB_Type})(a, b)
    V3<double> min_1, max_1, min_2, max_2;
    calc_min_max(a, min_1, max_1);
    calc_min_max(b, min_2, max_2);
    V3<double> v[4] = {min_1, max_1, min_2, max_2};
    return calc_obb_points(v, 4);
}
// src/obb/extensions.jl:2-3
inline double obb_area(const OBB& a) { return 8 * (a.e[0] * a.e[1] + a.e[1] * a.e[2] + a.e[2] * a.e[0]); }
inline double obb_volume(const OBB& a) { return 8 * (a.e[0] * a.e[1] * a.e[2]); }

// src/obb/obb_construction.jl:13-26  (p has n = 3 or 4 points, i_start is 1-based)
inline OBB make_obb(const V3<double>* p, int n, int i_start) {
    int i_next = (i_start % 3) + 1;  // mod1(i_start + 1, 3)
    V3<double> e1 = normalize(p[i_next - 1] - p[i_start - 1]);
    V3<double> e3 = triangleNormal(p[0], p[1], p[2]);
    V3<double> e2 = cross(e3, e1);
    V3<double> pmin, pmax;
    const V3<double>* ax[3] = {&e1, &e2, &e3};
    for (int a = 0; a < 3; ++a) {
        double lo = dot(p[0], *ax[a]), hi = lo;
        for (int k = 1; k < n; ++k) { double d = dot(p[k], *ax[a]); lo = std::fmin(lo, d); hi = std::fmax(hi, d); }
        pmin[a] = lo; pmax[a] = hi;
    }
    OBB o;
    V3<double> c = scale(pmax + pmin, 0.5);
    o.e = scale(pmax - pmin, 0.5);
    for (int r = 0; r < 3; ++r) { o.R(r, 0) = e1[r]; o.R(r, 1) = e2[r]; o.R(r, 2) = e3[r]; }
    o.c = mul3v(o.R, c);
    return o;
}
inline OBB fit_tri_obb(const V3<double>* p) { return make_obb(p, 3, 1); }
// src/obb/util.jl:54-66
inline void tet_perm_by_num(int n, int perm[4]) {
    static const int P[4][4] = {{2, 4, 3, 1}, {4, 1, 3, 2}, {1, 4, 2, 3}, {1, 2, 3, 4}};
    for (int k = 0; k < 4; ++k) perm[k] = P[n - 1][k];
}
inline int findmax_abs4(const double* e) {  // findmax returns the first maximal element (1-based)
    int i = 0;
    for (int k = 1; k < 4; ++k) if (std::fabs(e[k]) > std::fabs(e[i])) i = k;
    return i + 1;
}
// src/obb/obb_construction.jl:28-41 ; returns false for error("inverted tet")
inline bool fit_tet_obb(const V3<double>* p_in, const double* eps, OBB& out) {
    if (!(0.0 < tet_volume(p_in[0], p_in[1], p_in[2], p_in[3]))) return false;
    int perm[4];
    tet_perm_by_num(findmax_abs4(eps), perm);
    V3<double> p[4];
    for (int k = 0; k < 4; ++k) p[k] = p_in[perm[k] - 1];
    OBB o1 = make_obb(p, 4, 1), o2 = make_obb(p, 4, 2), o3 = make_obb(p, 4, 3);
    double a1 = obb_area(o1), a2 = obb_area(o2), a3 = obb_area(o3);
    if (std::fmax(a2, a3) <= a1) { out = o1; return true; }
    if (std::fmax(a1, a3) <= a2) { out = o2; return true; }
    out = o3;
    return true;
}

// Flattened bin_BB_Tree (src/obb/tree_types.jl:1-16): node k has box[k], children left/right
// (-1 for a leaf) and leaf_id (0-based primitive index, -1 for internal ~ id == -9999).
struct Tree {
    std::vector<OBB> box;
    std::vector<int32_t> left, right, leaf_id;
    int root = 0;
};

// src/obb/tree_types.jl:88-111
template <class S = double> inline void tree_tree_intersect(std::vector<std::pair<int32_t, int32_t>>& vc, int64_t& n_visited, const M3<double>& R_a_b,
                                const V3<double>& t_a_b, const Tree& t1, int n1, const Tree& t2, int n2) {
    ++n_visited;
    if (!BB_BB_intersect<S>(R_a_b, t_a_b, t1.box[n1], t2.box[n2])) return;
    bool is_leaf_1 = t1.leaf_id[n1] >= 0;
    bool is_leaf_2 = t2.leaf_id[n2] >= 0;
    if (is_leaf_1) {
        if (is_leaf_2) {
            vc.emplace_back(t1.leaf_id[n1], t2.leaf_id[n2]);
        } else {
            tree_tree_intersect<S>(vc, n_visited, R_a_b, t_a_b, t1, n1, t2, t2.left[n2]);
            tree_tree_intersect<S>(vc, n_visited, R_a_b, t_a_b, t1, n1, t2, t2.right[n2]);
        }
    } else {
        if (is_leaf_2) {
            tree_tree_intersect<S>(vc, n_visited, R_a_b, t_a_b, t1, t1.left[n1], t2, n2);
            tree_tree_intersect<S>(vc, n_visited, R_a_b, t_a_b, t1, t1.right[n1], t2, n2);
        } else {
            tree_tree_intersect<S>(vc, n_visited, R_a_b, t_a_b, t1, t1.left[n1], t2, t2.left[n2]);
            tree_tree_intersect<S>(vc, n_visited, R_a_b, t_a_b, t1, t1.right[n1], t2, t2.left[n2]);
            tree_tree_intersect<S>(vc, n_visited, R_a_b, t_a_b, t1, t1.left[n1], t2, t2.right[n2]);
            tree_tree_intersect<S>(vc, n_visited, R_a_b, t_a_b, t1, t1.right[n1], t2, t2.right[n2]);
        }
    }
}

// ----------------------------------------------------------------------------------------------
// Friction laws  (src/contact_algorithms_friction.jl:2-48, src/mechanism_scenario.jl:5-34)
// ----------------------------------------------------------------------------------------------
template <class T> inline T calc_clamped_piecewise(const T& x, double x_1, double x_2, double y_1, double y_2) {
    double k = (y_2 - y_1) / (x_2 - x_1);
    T y = y_1 + (x - x_1) * k;
    return clamp_(y, y_2, y_1);
}
struct Regularized { double mu_s, mu_d, v_c, v_mu_s, v_mu_d; };
struct Bristle { int id; double tau, k_bar, mu_s, mu_d, Ts_mu_s, Ts_mu_d, magic; };
inline Regularized make_regularized(double v_c, double mu_s, double mu_d) { return {mu_s, mu_d, v_c, 2 * v_c, 3 * v_c}; }
inline Bristle make_bristle(int id, double tau, double k_bar, double mu_s, double mu_d, double magic) {
    return {id, tau, k_bar, mu_s, mu_d, 2 * mu_s, 3 * mu_s, magic};
}
template <class T> inline V3<T> traction(const Regularized& ins, const V3<T>& vel_t, const T& p_dA) {
    T mag2 = dot(vel_t, vel_t);
    V3<T> Tc;
    if (value(mag2) < ins.v_c * ins.v_c) {
        Tc = divide(scale(vel_t, -ins.mu_s), ins.v_c);
    } else {
        T mag = sqrt_(mag2);
        T mu = calc_clamped_piecewise(mag, ins.v_mu_s, ins.v_mu_d, ins.mu_s, ins.mu_d);
        Tc = divide(mk3<T>((-mu) * vel_t[0], (-mu) * vel_t[1], (-mu) * vel_t[2]), mag);
    }
    return scale(Tc, p_dA);
}
template <class T> inline V3<T> traction(const Bristle& ins, const V3<T>& Ts, const T& p_dA) {
    T mag2 = dot(Ts, Ts);
    V3<T> Tc;
    if (value(mag2) < ins.mu_s * ins.mu_s) {
        Tc = Ts;
    } else {
        T mag = sqrt_(mag2);
        T mu = calc_clamped_piecewise(mag, ins.Ts_mu_s, ins.Ts_mu_d, ins.mu_s, ins.mu_d);
        Tc = divide(mk3<T>(mu * Ts[0], mu * Ts[1], mu * Ts[2]), mag);
    }
    return scale(Tc, p_dA);
}

// ----------------------------------------------------------------------------------------------
// Symmetric 6x6 eigen-decomposition.  The reference calls LAPACK (Float64) /
// GenericLinearAlgebra 0.1.0 (Dual) -- neither is under /root/reference.  Only
// V * f(lambda) * V' is consumed (friction.jl:85-96), which is independent of eigenvector sign
// and ordering, so a cyclic Jacobi iteration run to convergence is an equivalent restatement.
// In Dual mode the partials of V f(L) V' are the exact first-order perturbation of that matrix
// function (Daleckii-Krein), which is what differentiating through any converged generic
// eigen-solver yields wherever the result is differentiable.
// ----------------------------------------------------------------------------------------------
inline void jacobi_eigen6(double A[6][6], double V[6][6], double lam[6]) {
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) V[i][j] = (i == j ? 1.0 : 0.0);
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < 6; ++i) { diag += A[i][i] * A[i][i]; for (int j = i + 1; j < 6; ++j) off += A[i][j] * A[i][j]; }
        if (off == 0.0 || off <= 1.0e-44 * diag) break;
        for (int p = 0; p < 5; ++p)
            for (int q = p + 1; q < 6; ++q) {
                if (A[p][q] == 0.0) continue;
                double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                double t = (theta >= 0.0) ? 1.0 / (theta + std::sqrt(1.0 + theta * theta)) : -1.0 / (-theta + std::sqrt(1.0 + theta * theta));
                double c = 1.0 / std::sqrt(1.0 + t * t);
                double s = t * c;
                for (int k = 0; k < 6; ++k) { double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
                for (int k = 0; k < 6; ++k) { double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
                for (int k = 0; k < 6; ++k) { double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
            }
    }
    for (int i = 0; i < 6; ++i) lam[i] = A[i][i];
}

// src/mechanism_scenario.jl:60-76
template <class T> struct SpatialStiffness { T K[6][6]; T Kbar[6][6]; T Kbar_sqrt_inv[6][6]; T Sinv[6]; };

// src/contact_algorithms_friction.jl:85-96
inline void calc_Kbar_sqrt_inv(SpatialStiffness<double>& s) {
    double A[6][6], V[6][6], lam[6];
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) A[i][j] = (i <= j) ? s.Kbar[i][j] : s.Kbar[j][i];
    jacobi_eigen6(A, V, lam);
    double max_sig = lam[0];
    for (int k = 1; k < 6; ++k) max_sig = std::fmax(max_sig, lam[k]);
    double sig[6];
    for (int k = 0; k < 6; ++k) sig[k] = 1.0 / std::sqrt(max_(lam[k], max_sig * 1.0e-16));
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 6; ++k) acc += (V[i][k] * sig[k]) * V[j][k];
            s.Kbar_sqrt_inv[i][j] = acc;
        }
}
inline void calc_Kbar_sqrt_inv(SpatialStiffness<Counted>& s) {
    // flop accounting only: LAPACK's dsyev on a 6x6 is ~9 n^3 = 1944 FLOPs (estimate), the two
    // mul! of friction.jl:94-95 are 36 + 396
    SpatialStiffness<double> d;
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) d.Kbar[i][j] = s.Kbar[i][j].v;
    calc_Kbar_sqrt_inv(d);
    count_flops(1944 + 36 + 396 + 18);
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) s.Kbar_sqrt_inv[i][j] = Counted(d.Kbar_sqrt_inv[i][j]);
}
template <int N> inline void calc_Kbar_sqrt_inv(SpatialStiffness<Dual<N>>& s) {
    double A[6][6], V[6][6], lam[6];
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) A[i][j] = (i <= j) ? s.Kbar[i][j].v : s.Kbar[j][i].v;
    jacobi_eigen6(A, V, lam);
    int m = 0;
    for (int k = 1; k < 6; ++k) if (lam[k] > lam[m]) m = k;
    double floor_ = lam[m] * 1.0e-16;
    bool clamped[6]; double g[6], f[6], fp[6];
    for (int k = 0; k < 6; ++k) {
        clamped[k] = !(floor_ < lam[k]);
        g[k] = clamped[k] ? floor_ : lam[k];
        f[k] = 1.0 / std::sqrt(g[k]);
        fp[k] = -0.5 * f[k] / g[k];
    }
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 6; ++k) acc += (V[i][k] * f[k]) * V[j][k];
            s.Kbar_sqrt_inv[i][j] = Dual<N>(acc);
        }
    for (int d = 0; d < N; ++d) {
        double dA[6][6], tmp[6][6], B[6][6], G[6][6];
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) dA[i][j] = (i <= j) ? s.Kbar[i][j].p[d] : s.Kbar[j][i].p[d];
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double a = 0; for (int k = 0; k < 6; ++k) a += dA[i][k] * V[k][j]; tmp[i][j] = a; }
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double a = 0; for (int k = 0; k < 6; ++k) a += V[k][i] * tmp[k][j]; B[i][j] = a; }
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) {
                if (i == j) { G[i][j] = fp[i] * (clamped[i] ? 1.0e-16 * B[m][m] : B[i][i]); continue; }
                double F;
                if (clamped[i] && clamped[j]) F = 0.0;
                else if (!clamped[i] && !clamped[j]) { double si = std::sqrt(lam[i]), sj = std::sqrt(lam[j]); F = -1.0 / (si * sj * (si + sj)); }
                else F = (lam[i] == lam[j]) ? 0.0 : (f[i] - f[j]) / (lam[i] - lam[j]);
                G[i][j] = F * B[i][j];
            }
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double a = 0; for (int k = 0; k < 6; ++k) a += V[i][k] * G[k][j]; tmp[i][j] = a; }
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double a = 0; for (int k = 0; k < 6; ++k) a += tmp[i][k] * V[j][k]; s.Kbar_sqrt_inv[i][j].p[d] = a; }
    }
}

// src/contact_algorithms_friction.jl:98-117
template <class T> inline void decompose_K(SpatialStiffness<T>& s, double magic) {
    // Hermitian wrapper reads the upper triangle
    T Kf[6][6];
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) Kf[i][j] = (i <= j) ? s.K[i][j] : s.K[j][i];
    T t_1 = Kf[0][0] + Kf[1][1] + Kf[2][2];
    T t_2 = Kf[3][3] + Kf[4][4] + Kf[5][5];
    T s_1 = 1.0 / sqrt_(t_1);
    for (int k = 0; k < 3; ++k) s.Sinv[k] = s_1 * magic;
    T s_2 = 1.0 / sqrt_(t_2);
    for (int k = 3; k < 6; ++k) s.Sinv[k] = s_2;
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) s.Kbar[i][j] = s.Sinv[i] * Kf[i][j] * s.Sinv[j];
    calc_Kbar_sqrt_inv(s);
}

// ----------------------------------------------------------------------------------------------
// Scene description  (src/structs.jl:33-54, src/mechanism_scenario.jl:36-58)
// ----------------------------------------------------------------------------------------------
struct Mesh {
    int kind = 0;  // 0 = Tri, 1 = Tet
    std::vector<V3<double>> point;
    std::vector<int32_t> idx;  // 3 or 4 per primitive, 0-based
    std::vector<double> eps;
    double Ebar = 0.0;
    Tree tree;
    int n_prim() const { return int(idx.size()) / (kind == 0 ? 3 : 4); }
};
struct Instruction {
    int id_1, id_2;
    double chi;
    int model;  // 0 regularized, 1 bristle
    Regularized reg;
    Bristle bri;
    TriQuadRule quad;
};
template <class T> struct TractionCache { V3<T> n; V3<T> r_cart; T dA; T p; };
template <class T> inline T calc_p_dA(const TractionCache<T>& t) { return t.p * t.dA; }

// src/mechanism_scenario.jl:78-97
template <class T> struct BodyBodyCache {
    SpatialStiffness<T> stiff;
    std::vector<TractionCache<T>> traction;
    TriQuadRule quad;
    const Mesh* mesh_1 = nullptr;
    const Mesh* mesh_2 = nullptr;
    M4<T> x_r1_r2, x_r2_r1;
    V6<T> twist_r2_r1_r2;  // angular first, then linear (src/utility.jl:13-14)
    double chi = 0.0, Ebar = 0.0;
    ClipStatus clip_status;
};

// src/contact_algorithms_non_friction.jl:249-265
template <class T> inline void fillTractionCacheInnerLoop(int k, const TriQuadRule& quad, const BodyBodyCache<T>& b, const T& area_quad_k,
                                                           const V3<T> A[3], const V4<double>& eps_r2, V3<T>& r2, T& dA, T& p_hc) {
    const double* rphi = quad.zeta[k];
    for (int i = 0; i < 3; ++i) r2[i] = A[0][i] * rphi[0] + A[1][i] * rphi[1] + A[2][i] * rphi[2];
    T eps_quad = a_dot_one_pad_b(eps_r2, r2);
    V3<T> ang = mk3<T>(b.twist_r2_r1_r2[0], b.twist_r2_r1_r2[1], b.twist_r2_r1_r2[2]);
    V3<T> lin = mk3<T>(b.twist_r2_r1_r2[3], b.twist_r2_r1_r2[4], b.twist_r2_r1_r2[5]);
    V3<T> rdot = lin + cross(ang, r2);
    T ee = -(eps_r2[0] * rdot[0] + eps_r2[1] * rdot[1] + eps_r2[2] * rdot[2]);
    T damp_term = max_(T(0.0), 1.0 + b.chi * ee);
    p_hc = eps_quad * b.Ebar * damp_term;
    dA = quad.w[k] * area_quad_k;
}

// src/contact_algorithms_non_friction.jl:217-247
template <class T> inline void integrate_over_polygon_patch(BodyBodyCache<T>& b, const V3<T>& n2, const Poly<4, T>& poly_z2,
                                                             const M4<double>& x_r2_z2, const V4<double>& eps_r2) {
    const TriQuadRule& quad = b.quad;
    Poly<3, T> poly_r2 = mul_then_un_pad(x_r2_z2, poly_z2);
    V3<T> centroid_r2 = poly_centroid(poly_r2, n2).second;
    int N = poly_z2.n;
    V3<T> vert_2 = poly_r2.v[N - 1];
    for (int k = 0; k < N; ++k) {
        V3<T> vert_1 = vert_2;
        vert_2 = poly_r2.v[k];
        T area_quad_k = triangle_area(vert_1, vert_2, centroid_r2, n2);
        V3<T> A[3] = {vert_1, vert_2, centroid_r2};
        if (0.0 < value(area_quad_k)) {
            for (int q = 0; q < quad.n; ++q) {
                V3<T> r; T dA, p;
                fillTractionCacheInnerLoop(q, quad, b, area_quad_k, A, eps_r2, r, dA, p);
                if (0.0 < value(p)) b.traction.push_back(TractionCache<T>{n2, r, dA, p});
            }
        }
    }
}

inline M4<double> asMatOnePad4(const V3<double>* v) {
    M4<double> A;
    for (int c = 0; c < 4; ++c) { A(0, c) = v[c][0]; A(1, c) = v[c][1]; A(2, c) = v[c][2]; A(3, c) = 1.0; }
    return A;
}

template <class T> inline M3<T> rot_of(const M4<T>& x) { M3<T> R; for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) R(r, c) = x(r, c); return R; }

// src/contact_algorithms_non_friction.jl:196-215 (tri - tet)
template <class T> inline void integrate_over_tri_tet(int i_1, int i_2, BodyBodyCache<T>& b) {
    const Mesh& m1 = *b.mesh_1; const Mesh& m2 = *b.mesh_2;
    V3<double> vert_1[3], vert_2[4]; V4<double> eps2;
    for (int k = 0; k < 3; ++k) vert_1[k] = m1.point[m1.idx[3 * i_1 + k]];
    for (int k = 0; k < 4; ++k) { vert_2[k] = m2.point[m2.idx[4 * i_2 + k]]; eps2[k] = m2.eps[m2.idx[4 * i_2 + k]]; }
    M4<double> x_r2_z2 = asMatOnePad4(vert_2);
    M4<double> x_z2_r2 = inv44(x_r2_z2);
    V4<double> eps_r2 = mulrow4<double, double, double>(eps2, x_z2_r2);
    count_flops(28);  // the Float64-only row product above
    M4<T> x_z2_r1 = mul44<double, T, T>(x_z2_r2, b.x_r2_r1);
    V4<T> z[3];
    for (int k = 0; k < 3; ++k) {
        V4<T> o; o[0] = T(vert_1[k][0]); o[1] = T(vert_1[k][1]); o[2] = T(vert_1[k][2]); o[3] = T(1.0);
        z[k] = mul4v<T, T, T>(x_z2_r1, o);
    }
    Poly<4, T> poly_z2 = clip_in_tet_coordinates(z[0], z[1], z[2], b.clip_status);
    if (3 <= poly_z2.n) {
        V3<double> n_r1 = triangleNormal(vert_1[0], vert_1[1], vert_1[2]);
        count_flops(27);  // Float64-only: 2 differences, cross, scaling, dot, sqrt, 3 divisions
        V3<T> n2 = mul3v(rot_of(b.x_r2_r1), lift3<T>(n_r1));  // transform(::FreeVector3D, x) = R * v
        integrate_over_polygon_patch(b, n2, poly_z2, x_r2_z2, eps_r2);
    }
}

// src/contact_algorithms_non_friction.jl:164-194 (tet - tet)
template <class T> inline void integrate_over_tet_tet(int i_1, int i_2, BodyBodyCache<T>& b) {
    const Mesh& m1 = *b.mesh_1; const Mesh& m2 = *b.mesh_2;
    V3<double> vert_1[4], vert_2[4]; V4<double> eps1, eps2;
    for (int k = 0; k < 4; ++k) { vert_1[k] = m1.point[m1.idx[4 * i_1 + k]]; eps1[k] = m1.eps[m1.idx[4 * i_1 + k]]; }
    for (int k = 0; k < 4; ++k) { vert_2[k] = m2.point[m2.idx[4 * i_2 + k]]; eps2[k] = m2.eps[m2.idx[4 * i_2 + k]]; }
    M4<double> x_r1_z1 = asMatOnePad4(vert_1), x_z1_r1 = inv44(x_r1_z1);
    M4<double> x_r2_z2 = asMatOnePad4(vert_2), x_z2_r2 = inv44(x_r2_z2);
    // find_plane_tet(E, eps, X) = (E * eps) * X   (non_friction.jl:164)
    V4<double> E1e, E2e;
    for (int k = 0; k < 4; ++k) { E1e[k] = m1.Ebar * eps1[k]; E2e[k] = m2.Ebar * eps2[k]; }
    M4<T> x_z1_r2 = mul44<double, T, T>(x_z1_r1, b.x_r1_r2);
    V4<T> eps_plane_1_r2 = mulrow4<double, T, T>(E1e, x_z1_r2);
    V4<double> eps_r2 = mulrow4<double, double, double>(eps2, x_z2_r2);
    V4<double> eps_plane_2_r2 = mulrow4<double, double, double>(E2e, x_z2_r2);
    count_flops(8 + 28 + 28);  // Float64-only: E * eps (x2), two row products
    V4<T> eps_plane_r2;
    for (int k = 0; k < 4; ++k) eps_plane_r2[k] = eps_plane_2_r2[k] - eps_plane_1_r2[k];
    M4<T> x_r2_z1 = mul44<T, double, T>(b.x_r2_r1, x_r1_z1);
    Poly<3, T> poly_r2 = clip_plane_tet(eps_plane_r2, x_r2_z1);
    if (3 <= poly_r2.n) {
        Poly<4, T> poly_z2 = one_pad_then_mul(x_z2_r2, poly_r2);
        poly_z2 = zero_small_coordinates(poly_z2);
        poly_z2 = clip_in_tet_coordinates(poly_z2, b.clip_status);
        if (3 <= poly_z2.n) {
            V3<T> n2 = normalize(mk3<T>(eps_plane_r2[0], eps_plane_r2[1], eps_plane_r2[2]));
            integrate_over_polygon_patch(b, n2, poly_z2, x_r2_z2, eps_r2);
        }
    }
}

// src/contact_algorithms_normal.jl:2-34
template <class T> inline void normal_wrench(const BodyBodyCache<T>& b, V3<T>& ang, V3<T>& lin) {
    lin = zero3<T>(); ang = zero3<T>();
    for (const auto& trac : b.traction) {
        T p_dA = calc_p_dA(trac);
        V3<T> lam = scale(trac.n, p_dA);
        lin = lin + lam;
        ang = ang + cross(trac.r_cart, lam);
    }
}
template <class T> inline V3<T> normal_wrench_cop(const BodyBodyCache<T>& b, V3<T>& ang, V3<T>& lin) {
    lin = zero3<T>(); ang = zero3<T>();
    V3<T> int_p_dA_cop = zero3<T>();
    T int_p_dA = T(0.0);
    for (const auto& trac : b.traction) {
        T p_dA = calc_p_dA(trac);
        V3<T> lam = scale(trac.n, p_dA);
        lin = lin + lam;
        ang = ang + cross(trac.r_cart, lam);
        int_p_dA = int_p_dA + p_dA;
        int_p_dA_cop = int_p_dA_cop + scale(trac.r_cart, p_dA);
    }
    return divide(int_p_dA_cop, int_p_dA);
}

template <class T> inline V3<T> spatial_vel_formula(const V6<T>& v, const V3<T>& r) {
    return mk3<T>(v[3], v[4], v[5]) + cross(mk3<T>(v[0], v[1], v[2]), r);
}

// src/contact_algorithms_friction.jl:50-72
template <class T> inline V6<T> yes_contact_regularized(const Regularized& fric, const BodyBodyCache<T>& b) {
    V3<T> lin = zero3<T>(), ang = zero3<T>();
    for (const auto& trac : b.traction) {
        V3<T> cart_vel = spatial_vel_formula(b.twist_r2_r1_r2, trac.r_cart);
        V3<T> cart_vel_t = vec_sub_vec_proj(cart_vel, trac.n);
        T p_dA = calc_p_dA(trac);
        V3<T> T_c = traction(fric, cart_vel_t, p_dA);
        V3<T> traction_k = scale(trac.n, p_dA) + T_c;
        lin = lin + traction_k;
        ang = ang + cross(trac.r_cart, traction_k);
    }
    V6<T> w; for (int k = 0; k < 3; ++k) { w[k] = ang[k]; w[3 + k] = lin[k]; }
    return w;
}

// src/contact_algorithms_friction.jl:147-169
template <class T> inline void calc_patch_spatial_stiffness(BodyBodyCache<T>& b, const Bristle& BF, const V3<T>& cop) {
    T K11[3][3], K12[3][3], K22[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) K11[i][j] = K12[i][j] = K22[i][j] = T(0.0);
    for (const auto& trac : b.traction) {
        const V3<T>& n = trac.n;
        T p_dA = calc_p_dA(trac);
        V3<T> r = trac.r_cart - cop;
        V3<T> rxn = cross(r, n);
        // vector_to_skew_symmetric / _squared: RigidBodyDynamics 1.4.0 Spatial (not vendored)
        T skew[3][3] = {{T(0.0), -r[2], r[1]}, {r[2], T(0.0), -r[0]}, {-r[1], r[0], T(0.0)}};
        T a0 = r[0] * r[0], a1 = r[1] * r[1], a2 = r[2] * r[2];
        T b12 = r[0] * r[1], b13 = r[0] * r[2], b23 = r[1] * r[2];
        T sk2[3][3] = {{-a1 - a2, b12, b13}, {b12, -a0 - a2, b23}, {b13, b23, -a0 - a1}};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                T eye = T(i == j ? 1.0 : 0.0);
                K22[i][j] = K22[i][j] + p_dA * (eye - n[i] * n[j]);
                K12[i][j] = K12[i][j] + p_dA * (skew[i][j] - rxn[i] * n[j]);
                K11[i][j] = K11[i][j] - p_dA * (sk2[i][j] + rxn[i] * rxn[j]);
            }
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            b.stiff.K[i][j] = K11[i][j] * BF.k_bar;
            b.stiff.K[3 + i][j] = K12[j][i] * BF.k_bar;
            b.stiff.K[i][3 + j] = K12[i][j] * BF.k_bar;
            b.stiff.K[3 + i][3 + j] = K22[i][j] * BF.k_bar;
        }
}

// src/contact_algorithms_friction.jl:171-201
template <class T> inline void calc_spatial_bristle_force(const BodyBodyCache<T>& b, const Bristle& BF, const V6<T>& Delta2, const V3<T>& cop,
                                                          V6<T>& wrench_cop_fric, V6<T>& wrench2_fric) {
    V3<T> lin = zero3<T>(), ang = zero3<T>();
    for (const auto& trac : b.traction) {
        const V3<T>& n = trac.n;
        const V3<T>& r = trac.r_cart;
        V3<T> x2 = r - cop;
        V3<T> delta2 = spatial_vel_formula(Delta2, x2);
        V3<T> rdot_perp = spatial_vel_formula(b.twist_r2_r1_r2, r);
        T p_dA = calc_p_dA(trac);
        V3<T> Ts = scale(delta2 + scale(rdot_perp, BF.tau), -BF.k_bar);
        Ts = vec_sub_vec_proj(Ts, n);
        V3<T> T_c = traction(BF, Ts, p_dA);
        lin = lin + T_c;
        ang = ang + cross(x2, T_c);
    }
    V3<T> ang2 = ang + cross(cop, lin);
    for (int k = 0; k < 3; ++k) { wrench_cop_fric[k] = ang[k]; wrench_cop_fric[3 + k] = lin[k]; wrench2_fric[k] = ang2[k]; wrench2_fric[3 + k] = lin[k]; }
}

// src/contact_algorithms_friction.jl:119-143
template <class T> inline V6<T> yes_contact_bristle(const Bristle& BF, BodyBodyCache<T>& b, const V6<T>& s, V6<T>& sdot) {
    double tau_inv = 1 / BF.tau;
    V3<T> n_ang, n_lin;
    V3<T> cop = normal_wrench_cop(b, n_ang, n_lin);
    calc_patch_spatial_stiffness(b, BF, cop);
    decompose_K(b.stiff, BF.magic);
    V6<T> tmp, Delta2;
    for (int i = 0; i < 6; ++i) { T acc = T(0.0); for (int j = 0; j < 6; ++j) acc = acc + b.stiff.Kbar_sqrt_inv[i][j] * s[j]; tmp[i] = acc; }
    for (int i = 0; i < 6; ++i) Delta2[i] = b.stiff.Sinv[i] * tmp[i];
    V6<T> w_cop, w2;
    calc_spatial_bristle_force(b, BF, Delta2, cop, w_cop, w2);
    V6<T> sw;
    for (int i = 0; i < 6; ++i) sw[i] = b.stiff.Sinv[i] * w_cop[i];
    for (int i = 0; i < 6; ++i) {
        T acc = T(0.0);
        for (int j = 0; j < 6; ++j) acc = acc + b.stiff.Kbar_sqrt_inv[i][j] * sw[j];
        sdot[i] = (-tau_inv) * (acc + s[i]);
    }
    V6<T> w;
    for (int k = 0; k < 3; ++k) { w[k] = n_ang[k] + w2[k]; w[3 + k] = n_lin[k] + w2[3 + k]; }
    return w;
}

// RigidBodyDynamics 1.4.0 inv(::Transform3D): R' and -(R' * t)  (not vendored; call sites
// src/contact_algorithms_non_friction.jl:111,113)
template <class T> inline M4<T> inv_transform(const M4<T>& x) {
    M4<T> r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = x(j, i);
    for (int i = 0; i < 3; ++i) r(i, 3) = -(r(i, 0) * x(0, 3) + r(i, 1) * x(1, 3) + r(i, 2) * x(2, 3));
    r(3, 0) = T(0.0); r(3, 1) = T(0.0); r(3, 2) = T(0.0); r(3, 3) = T(1.0);
    return r;
}

struct Scene {
    std::vector<Mesh> mesh;
    std::vector<Instruction> ins;
    int n_bristle = 0;
    // last-evaluated diagnostics (per env*ins), kept for pfc_get_pairs-style parity checks
    std::vector<std::vector<std::pair<int32_t, int32_t>>> last_pairs;
    std::vector<std::vector<double>> last_traction;  // 8 doubles per point (values): n(3) r(3) dA p
    int64_t n_visited = 0;
    std::string err;
};

// One instruction, one environment:  src/contact_algorithms_non_friction.jl:70-134 (+ :136-143)
//   X_bp   : 4x4 col-major Float64 x_r2_r1 used for the broad phase (the Float64 state, R8/H6)
//   X      : x_r2_r1 in mode T;  twist: [ang; lin] of r2 w.r.t. r1 expressed in r2
//   s/sdot : bristle state (6) when model == 1
// Returns flags: bit0 contact, bit1 non-finite vertex, bit2 bad arity.
template <class T>
inline int force_single_elastic_intersection(const Scene& sc, const Instruction& ci, const double* X_bp, const T* X, const T* twist, const T* s_in,
                                             T* wrench_out, T* sdot_out, std::vector<std::pair<int32_t, int32_t>>& pairs,
                                             std::vector<double>* traction_dump, int64_t& n_visited, int64_t* broad_flops = nullptr, int64_t* n_points = nullptr) {
    const Mesh& m1 = sc.mesh[ci.id_1];
    const Mesh& m2 = sc.mesh[ci.id_2];
    // calcTriTetIntersections! : broad phase always on the Float64 transform
    M4<double> xf; for (int k = 0; k < 16; ++k) xf.m[k] = X_bp[k];
    M4<double> x_r1_r2_f = inv_transform(xf);
    M3<double> R_a_b = rot_of(x_r1_r2_f);
    V3<double> t_a_b = mk3<double>(x_r1_r2_f(0, 3), x_r1_r2_f(1, 3), x_r1_r2_f(2, 3));
    pairs.clear();
    {
        const int64_t before = flop_counter();
        if (std::is_same<T, Counted>::value) tree_tree_intersect<Counted>(pairs, n_visited, R_a_b, t_a_b, m1.tree, m1.tree.root, m2.tree, m2.tree.root);
        else tree_tree_intersect<double>(pairs, n_visited, R_a_b, t_a_b, m1.tree, m1.tree.root, m2.tree, m2.tree.root);
        if (broad_flops) *broad_flops += flop_counter() - before;
    }
    for (int k = 0; k < 6; ++k) wrench_out[k] = T(0.0);
    int flags = 0;
    bool contact = false;
    if (!pairs.empty()) {
        BodyBodyCache<T> b;
        b.mesh_1 = &m1; b.mesh_2 = &m2;
        for (int k = 0; k < 16; ++k) b.x_r2_r1.m[k] = X[k];
        b.x_r1_r2 = inv_transform(b.x_r2_r1);
        for (int k = 0; k < 6; ++k) b.twist_r2_r1_r2[k] = twist[k];
        b.chi = ci.chi; b.Ebar = m2.Ebar; b.quad = ci.quad;
        for (const auto& pr : pairs) {
            if (m1.kind == 0) integrate_over_tri_tet(pr.first, pr.second, b);
            else integrate_over_tet_tet(pr.first, pr.second, b);
        }
        if (b.clip_status.non_finite) flags |= 2;
        if (b.clip_status.bad_arity) flags |= 4;
        if (traction_dump) {
            traction_dump->clear();
            for (const auto& tc : b.traction) {
                for (int k = 0; k < 3; ++k) traction_dump->push_back(value(tc.n[k]));
                for (int k = 0; k < 3; ++k) traction_dump->push_back(value(tc.r_cart[k]));
                traction_dump->push_back(value(tc.dA));
                traction_dump->push_back(value(tc.p));
            }
        }
        if (n_points) *n_points += int64_t(b.traction.size());
        if (!b.traction.empty()) {
            contact = true;
            V6<T> w;
            if (ci.model == 0) {
                w = yes_contact_regularized(ci.reg, b);
            } else {
                V6<T> s, sd;
                for (int k = 0; k < 6; ++k) s[k] = s_in[6 * ci.bri.id + k];
                w = yes_contact_bristle(ci.bri, b, s, sd);
                for (int k = 0; k < 6; ++k) sdot_out[6 * ci.bri.id + k] = sd[k];
            }
            for (int k = 0; k < 6; ++k) wrench_out[k] = w[k];
        }
    } else if (traction_dump) {
        traction_dump->clear();
    }
    if (!contact && ci.model == 1) {  // no_contact!(::Bristle): friction.jl:76-81
        for (int k = 0; k < 6; ++k) sdot_out[6 * ci.bri.id + k] = (-(1 / ci.bri.tau)) * s_in[6 * ci.bri.id + k];
    }
    if (contact) flags |= 1;
    return flags;
}

}  // namespace orc
