// pfc_math.cuh -- scalar and small-vector device helpers for the contact-wrench kernels.
//
// The kernels are templated on the scalar type T in {double, Dual<6>} (forward-mode value + 6
// partials; the reference's Jacobian mode runs the same source on ForwardDiff.Dual{Nothing,
// Float64,6}: /root/reference/src/radau/radau_functions.jl:2-26).  Mesh constants
// are always double.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pfc {

#define PFC_HD __host__ __device__ __forceinline__
#define PFC_D __device__ __forceinline__

// ---- non-contracted arithmetic for the broad phase --------------------------------------------------
// Julia never fuses a*b+c unless the source says muladd; the SAT outcome must be bit-exact, so the
// broad phase uses these (ptxas cannot contract the _rn intrinsics).
PFC_D double mul_(double a, double b) { return __dmul_rn(a, b); }
PFC_D double add_(double a, double b) { return __dadd_rn(a, b); }
PFC_D double sub_(double a, double b) { return __dadd_rn(a, -b); }

// ---- wide gathers -----------------------------------------------------------------------------------------
// Every lane gathers its own primitive / node record, so one load instruction touches up to 32 different lines and the L1 data
// pipe pays per request, not per byte (ncu: l1tex__data_pipe_lsu_wavefronts is the busiest unit of the SAT and clip kernels).
// sm_100 has 256-bit global loads (ld.global.v4.b64): 4x fewer requests than 64-bit loads for the same record.
// p must be 32 B aligned; N = number of doubles, a multiple of 4.
template <int N> PFC_D void load_wide(const double* __restrict__ p, double* out) {
    static_assert(N % 4 == 0, "load_wide moves groups of 4 doubles");
#ifdef PFC_HOST_CHECK   // tests/test_device_*_on_host.py compile the device headers with g++ (no PTX there)
    for (int j = 0; j < N; ++j) out[j] = p[j];
#else
#pragma unroll
    for (int j = 0; j < N / 4; ++j) {
        unsigned long long a, b, c, d;
        asm volatile("ld.global.nc.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p + 4 * j));
        out[4 * j] = __longlong_as_double((long long)a); out[4 * j + 1] = __longlong_as_double((long long)b);
        out[4 * j + 2] = __longlong_as_double((long long)c); out[4 * j + 3] = __longlong_as_double((long long)d);
    }
#endif
}

// ---- Dual<N> ---------------------------------------------------------------------------------------------
template <int N>
struct Dual {
    double v;
    double p[N];
    PFC_HD Dual() {}
    PFC_HD Dual(double x) : v(x) {
#pragma unroll
        for (int i = 0; i < N; ++i) p[i] = 0.0;
    }
};

PFC_HD double val(double x) { return x; }
template <int N> PFC_HD double val(const Dual<N>& x) { return x.v; }

#define PFC_DUAL_LOOP _Pragma("unroll") for (int i = 0; i < N; ++i)
template <int N> PFC_HD Dual<N> operator-(const Dual<N>& a) { Dual<N> r; r.v = -a.v; PFC_DUAL_LOOP r.p[i] = -a.p[i]; return r; }
template <int N> PFC_HD Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v + b.v; PFC_DUAL_LOOP r.p[i] = a.p[i] + b.p[i]; return r; }
template <int N> PFC_HD Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v - b.v; PFC_DUAL_LOOP r.p[i] = a.p[i] - b.p[i]; return r; }
template <int N> PFC_HD Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v * b.v; PFC_DUAL_LOOP r.p[i] = a.p[i] * b.v + a.v * b.p[i]; return r; }
template <int N> PFC_HD Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; double ib = 1.0 / b.v; r.v = a.v * ib;
    PFC_DUAL_LOOP r.p[i] = (a.p[i] - r.v * b.p[i]) * ib;
    return r;
}
template <int N> PFC_HD Dual<N> operator+(const Dual<N>& a, double b) { Dual<N> r = a; r.v = a.v + b; return r; }
template <int N> PFC_HD Dual<N> operator+(double a, const Dual<N>& b) { Dual<N> r = b; r.v = a + b.v; return r; }
template <int N> PFC_HD Dual<N> operator-(const Dual<N>& a, double b) { Dual<N> r = a; r.v = a.v - b; return r; }
template <int N> PFC_HD Dual<N> operator-(double a, const Dual<N>& b) { Dual<N> r; r.v = a - b.v; PFC_DUAL_LOOP r.p[i] = -b.p[i]; return r; }
template <int N> PFC_HD Dual<N> operator*(const Dual<N>& a, double b) { Dual<N> r; r.v = a.v * b; PFC_DUAL_LOOP r.p[i] = a.p[i] * b; return r; }
template <int N> PFC_HD Dual<N> operator*(double a, const Dual<N>& b) { return b * a; }
template <int N> PFC_HD Dual<N> operator/(const Dual<N>& a, double b) { double ib = 1.0 / b; return a * ib; }
template <int N> PFC_HD Dual<N> operator/(double a, const Dual<N>& b) {
    Dual<N> r; double ib = 1.0 / b.v; r.v = a * ib; double s = -r.v * ib;
    PFC_DUAL_LOOP r.p[i] = s * b.p[i];
    return r;
}
template <int N> PFC_HD Dual<N>& operator+=(Dual<N>& a, const Dual<N>& b) { a.v += b.v; PFC_DUAL_LOOP a.p[i] += b.p[i]; return a; }
template <int N> PFC_HD Dual<N>& operator-=(Dual<N>& a, const Dual<N>& b) { a.v -= b.v; PFC_DUAL_LOOP a.p[i] -= b.p[i]; return a; }

PFC_HD double sqrt_(double x) { return sqrt(x); }
template <int N> PFC_HD Dual<N> sqrt_(const Dual<N>& a) {
    Dual<N> r; r.v = sqrt(a.v); double s = 0.5 / r.v;
    PFC_DUAL_LOOP r.p[i] = a.p[i] * s;
    return r;
}
PFC_HD double fma_(double a, double b, double c) { return fma(a, b, c); }
template <int N> PFC_HD Dual<N> fma_(const Dual<N>& a, const Dual<N>& b, const Dual<N>& c) { return a * b + c; }
template <int N> PFC_HD Dual<N> fma_(double a, const Dual<N>& b, const Dual<N>& c) { return a * b + c; }
template <int N> PFC_HD Dual<N> fma_(double a, const Dual<N>& b, double c) { return a * b + c; }

// ---- 3-vectors ---------------------------------------------------------------------------------------------
template <class T> struct Vec3 { T x, y, z; };
template <class T> PFC_HD Vec3<T> mk(const T& x, const T& y, const T& z) { Vec3<T> r; r.x = x; r.y = y; r.z = z; return r; }
template <class T> PFC_HD Vec3<T> operator+(const Vec3<T>& a, const Vec3<T>& b) { return mk<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class T> PFC_HD Vec3<T> operator-(const Vec3<T>& a, const Vec3<T>& b) { return mk<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class T> PFC_HD Vec3<T> operator*(const Vec3<T>& a, const T& s) { return mk<T>(a.x * s, a.y * s, a.z * s); }
template <int N> PFC_HD Vec3<Dual<N>> operator*(const Vec3<Dual<N>>& a, double s) { return mk<Dual<N>>(a.x * s, a.y * s, a.z * s); }
template <class T> PFC_HD T dot(const Vec3<T>& a, const Vec3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> PFC_HD Vec3<T> cross(const Vec3<T>& a, const Vec3<T>& b) {
    return mk<T>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
template <class T> PFC_HD Vec3<T> lift(const Vec3<double>& a) { return mk<T>(T(a.x), T(a.y), T(a.z)); }

// Rigid transform in mode T: p -> R p + t, R row-major r[3*i+j]
template <class T> struct Xform { T r[9]; T t[3]; };
template <class T> PFC_HD Vec3<T> rot(const Xform<T>& X, const Vec3<T>& p) {
    return mk<T>(X.r[0] * p.x + X.r[1] * p.y + X.r[2] * p.z, X.r[3] * p.x + X.r[4] * p.y + X.r[5] * p.z, X.r[6] * p.x + X.r[7] * p.y + X.r[8] * p.z);
}
template <class T> PFC_HD Vec3<T> rot_d(const Xform<T>& X, const Vec3<double>& p) {  // constant (double) point
    return mk<T>(X.r[0] * p.x + X.r[1] * p.y + X.r[2] * p.z, X.r[3] * p.x + X.r[4] * p.y + X.r[5] * p.z, X.r[6] * p.x + X.r[7] * p.y + X.r[8] * p.z);
}
template <class T> PFC_HD Vec3<T> apply_d(const Xform<T>& X, const Vec3<double>& p) {
    Vec3<T> q = rot_d(X, p);
    return mk<T>(q.x + X.t[0], q.y + X.t[1], q.z + X.t[2]);
}
// inverse of a rigid transform (RigidBodyDynamics inv(::Transform3D): R', -(R' t))
template <class T> PFC_HD Xform<T> inverse(const Xform<T>& X) {
    Xform<T> Y;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Y.r[3 * i + j] = X.r[3 * j + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) Y.t[i] = -(Y.r[3 * i] * X.t[0] + Y.r[3 * i + 1] * X.t[1] + Y.r[3 * i + 2] * X.t[2]);
    return Y;
}

// ---- warp reductions in a fixed (xor-butterfly) order: every lane ends with the same bits -------------------
PFC_D double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
template <int N> PFC_D Dual<N> warp_sum(Dual<N> x) {
    x.v = warp_sum(x.v);
    PFC_DUAL_LOOP x.p[i] = warp_sum(x.p[i]);
    return x;
}

}  // namespace pfc
