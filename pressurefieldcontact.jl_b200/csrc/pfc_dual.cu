// pfc_dual.cu -- Jacobian mode: the narrow phase, friction and reduction on Dual<6> scalars.
//
// Radau obtains the Jacobian of the ODE right-hand side by forward-mode AD in chunks of 6
// (/root/reference/src/radau/radau_functions.jl:2-26, N_chunk = 6 at src/mechanism_scenario.jl:181):
// forceAllElasticIntersections! then runs on ForwardDiff.Dual{Nothing,Float64,6}, i.e. every scalar
// that depends on the state carries its value and 6 partials, while the candidate-pair lists come
// from the Float64 state (calcTriTetIntersections! always uses m.float, non_friction.jl:94-101).
// The device code is the same templated source as the Float64 path (pfc_clip.cuh, pfc_patch.cuh)
// instantiated with T = Dual<6>; one warp owns one (environment, instruction), lanes stride over the
// pair list, and the per-lane partial sums (value and partials) are combined by the fixed-order
// butterfly.  The 6x6 matrix function K̄^(-1/2) of the bristle model is differentiated analytically
// (Daleckii-Krein), which is what differentiating through any converged eigen-solver yields.
#include "pfc_bristle.cuh"
#include "pfc_dual.cuh"

namespace pfc {

typedef Dual<6> D6;

namespace {

PFC_D D6 ld7(const double* p) { D6 r; r.v = p[0];
#pragma unroll
    for (int i = 0; i < 6; ++i) r.p[i] = p[1 + i];
    return r; }
PFC_D void st7(double* p, const D6& x) { p[0] = x.v;
#pragma unroll
    for (int i = 0; i < 6; ++i) p[1 + i] = x.p[i]; }
PFC_D D6 max0(const D6& a, const D6& b) { return (b.v < a.v) ? a : b; }  // Julia max(x, y) = ifelse(y < x, x, y)

// decompose_K! in Dual mode.  a21: K11 upper (0..5), K12 row-major (6..14), K22 upper (15..20), times k_bar.
__device__ __noinline__ void decompose_K_dual(const D6* a21, double magic, D6* Sinv, D6* Kh) {
    D6 K[36];
    K[0] = a21[0]; K[1] = a21[1]; K[2] = a21[2]; K[7] = a21[3]; K[8] = a21[4]; K[14] = a21[5];
    K[6] = K[1]; K[12] = K[2]; K[13] = K[8];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { K[6 * i + 3 + j] = a21[6 + 3 * i + j]; K[6 * (3 + j) + i] = a21[6 + 3 * i + j]; }
    K[21] = a21[15]; K[22] = a21[16]; K[23] = a21[17]; K[28] = a21[18]; K[29] = a21[19]; K[35] = a21[20];
    K[27] = K[22]; K[33] = K[23]; K[34] = K[29];
    const D6 t1 = K[0] + K[7] + K[14], t2 = K[21] + K[28] + K[35];
    const D6 s1 = 1.0 / sqrt_(t1), s2 = 1.0 / sqrt_(t2);
    for (int k = 0; k < 3; ++k) { Sinv[k] = s1 * magic; Sinv[3 + k] = s2; }
    D6 Kb[36];
    double A[36], V[36], lam[6];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) { Kb[6 * i + j] = Sinv[i] * K[6 * i + j] * Sinv[j]; A[6 * i + j] = Kb[6 * i + j].v; }
    jacobi6(A, V, lam);
    int m = 0;
    for (int k = 1; k < 6; ++k) if (lam[k] > lam[m]) m = k;
    const double floor_ = lam[m] * 1.0e-16;
    bool clamped[6];
    double f[6], fp[6];
    for (int k = 0; k < 6; ++k) {
        clamped[k] = !(floor_ < lam[k]);
        const double g = clamped[k] ? floor_ : lam[k];
        f[k] = 1.0 / sqrt(g);
        fp[k] = -0.5 * f[k] / g;
    }
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 6; ++k) acc += (V[6 * i + k] * f[k]) * V[6 * j + k];
            Kh[6 * i + j] = D6(acc);
        }
    for (int d = 0; d < 6; ++d) {
        double T1[36], B[36];
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) { double a = 0; for (int k = 0; k < 6; ++k) a += Kb[6 * i + k].p[d] * V[6 * k + j]; T1[6 * i + j] = a; }
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) { double a = 0; for (int k = 0; k < 6; ++k) a += V[6 * k + i] * T1[6 * k + j]; B[6 * i + j] = a; }
        const double Bmm = B[7 * m];
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) {
                double G;
                if (i == j) G = fp[i] * (clamped[i] ? 1.0e-16 * Bmm : B[7 * i]);
                else {
                    double F;
                    if (clamped[i] && clamped[j]) F = 0.0;
                    else if (!clamped[i] && !clamped[j]) { const double si = sqrt(lam[i]), sj = sqrt(lam[j]); F = -1.0 / (si * sj * (si + sj)); }
                    else F = (lam[i] == lam[j]) ? 0.0 : (f[i] - f[j]) / (lam[i] - lam[j]);
                    G = F * B[6 * i + j];
                }
                T1[6 * i + j] = G;
            }
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) { double a = 0; for (int k = 0; k < 6; ++k) a += V[6 * i + k] * T1[6 * k + j]; B[6 * i + j] = a; }
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) { double a = 0; for (int k = 0; k < 6; ++k) a += B[6 * i + k] * V[6 * j + k]; Kh[6 * i + j].p[d] = a; }
    }
}

PFC_D void value_ctx(const PatchCtx<D6>& cx, PatchCtx<double>& v) {
#pragma unroll
    for (int i = 0; i < 9; ++i) { v.x21.r[i] = cx.x21.r[i].v; v.x12.r[i] = cx.x12.r[i].v; }
#pragma unroll
    for (int i = 0; i < 3; ++i) { v.x21.t[i] = cx.x21.t[i].v; v.x12.t[i] = cx.x12.t[i].v; }
    v.w_ang = mk<double>(cx.w_ang.x.v, cx.w_ang.y.v, cx.w_ang.z.v);
    v.w_lin = mk<double>(cx.w_lin.x.v, cx.w_lin.y.v, cx.w_lin.z.v);
    v.chi = cx.chi; v.Ebar1 = cx.Ebar1; v.Ebar2 = cx.Ebar2; v.n_quad = cx.n_quad;
}

// not inlined: the bristle branch makes four passes over the pair list, and four inlined copies of the Dual<6> pipeline cost ptxas minutes
template <int NA>
__device__ __noinline__ void run_pairs_dual(const SceneDev& sc, const InsDev& ins, const PairSource& ps, long long env, int k, long long ei, int n, int lane,
                          const PatchCtx<D6>& cx, const PatchCtx<double>& cxv, Accum<D6, NA>& acc, int& flags) {
    if (ins.small) {
        const unsigned* pl = ps.small_pairs + (size_t)ps.small_cap * ei;
        for (int i = lane; i < n; i += 32) {
            const unsigned e = pl[i];
            const int a = int((e >> 15) & 0x7fffu), b = int(e & 0x7fffu);
            if (survives_f64(sc, ins, a, b, cxv)) integrate_pair(sc, ins, a, b, cx, acc, flags);
        }
    } else {
        const int3* pl = ps.large_sorted + ps.seg_start[env * ps.n_large + ps.large_index[k]];
        for (int i = lane; i < n; i += 32) {
            const int3 e = pl[i];
            if (survives_f64(sc, ins, e.y, e.z, cxv)) integrate_pair(sc, ins, e.y, e.z, cx, acc, flags);
        }
    }
}

// HB: the scene has bristle instructions (21 Dual accumulators, the three passes and the Dual matrix function are compiled in);
// regularized-only scenes get a kernel with 6 accumulators and a fraction of the stack.
template <bool HB>
__global__ void __launch_bounds__(128) eval_dual6_kernel(SceneDev sc, DualIO io, PairSource ps) {
    constexpr int NA = HB ? 21 : 6;
    const int lane = threadIdx.x & 31;
    const long long n_prob = io.n_env * sc.n_ins;
    for (long long ei = (long long)blockIdx.x * 4 + (threadIdx.x >> 5); ei < n_prob; ei += (long long)gridDim.x * 4) {
        const long long env = ei / sc.n_ins;
        const int k = int(ei - env * sc.n_ins);
        const InsDev& ins = sc.ins[k];
        const int n = (int)io.n_pairs[ei];
        int flags = 0;
        D6 w[6];
        for (int j = 0; j < 6; ++j) w[j] = D6(0.0);
        bool contact = false;
        const bool bristle = HB && ins.model == PFC_MODEL_BRISTLE;
        const double* sv = bristle ? io.s7 + 42 * ((long long)sc.n_bristle * env + ins.bristle_id) : nullptr;
        double* sd = bristle ? io.sdot7 + 42 * ((long long)sc.n_bristle * env + ins.bristle_id) : nullptr;
        if (n > 0) {
            PatchCtx<D6> cx;
            const double* Xp = io.X7 + 112 * ei;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
#pragma unroll
                for (int j = 0; j < 3; ++j) cx.x21.r[3 * i + j] = ld7(Xp + 7 * (4 * j + i));
                cx.x21.t[i] = ld7(Xp + 7 * (12 + i));
            }
            cx.x12 = inverse(cx.x21);
            const double* tw = io.twist7 + 42 * ei;
            cx.w_ang = mk<D6>(ld7(tw), ld7(tw + 7), ld7(tw + 14));
            cx.w_lin = mk<D6>(ld7(tw + 21), ld7(tw + 28), ld7(tw + 35));
            cx.chi = ins.chi; cx.Ebar1 = ins.Ebar1; cx.Ebar2 = ins.Ebar2; cx.n_quad = ins.n_quad;
            PatchCtx<double> cxv;
            value_ctx(cx, cxv);
            // An instruction none of whose inputs depends on the seeded state entries (every partial of x_r2_r1, the twist and, for
            // bristle friction, s is zero -- e.g. the seeds sit on another body) has zero wrench partials: its Dual evaluation is the
            // Float64 evaluation, 7x cheaper.  The reference evaluates such instructions on Duals all the same; the values agree to rounding.
            bool seeded = false;
#pragma unroll
            for (int i = 0; i < 9; ++i)
#pragma unroll
                for (int q = 0; q < 6; ++q) seeded |= (cx.x21.r[i].p[q] != 0.0);
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int q = 0; q < 6; ++q) seeded |= (cx.x21.t[i].p[q] != 0.0) | ((&cx.w_ang.x)[i].p[q] != 0.0) | ((&cx.w_lin.x)[i].p[q] != 0.0);
            if (!seeded && !bristle) {
                Accum<double, 6> av;
                av.fp = ins.p; av.w_ang = cxv.w_ang; av.w_lin = cxv.w_lin; av.dump = nullptr; av.dump_cap = 0;
                av.reset(ACC_REGULARIZED);
                if (ins.small) {
                    const unsigned* pl = ps.small_pairs + (size_t)ps.small_cap * ei;
                    for (int i = lane; i < n; i += 32) {
                        const unsigned e = pl[i];
                        const int a = int((e >> 15) & 0x7fffu), b = int(e & 0x7fffu);
                        if (prefilter_pair(sc, ins, a, b, cxv)) integrate_pair(sc, ins, a, b, cxv, av, flags);
                    }
                } else {
                    const int3* pl = ps.large_sorted + ps.seg_start[env * ps.n_large + ps.large_index[k]];
                    for (int i = lane; i < n; i += 32) {
                        const int3 e = pl[i];
                        if (prefilter_pair(sc, ins, e.y, e.z, cxv)) integrate_pair(sc, ins, e.y, e.z, cxv, av, flags);
                    }
                }
                int pts = av.n_points;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) pts += __shfl_xor_sync(0xffffffffu, pts, o);
                contact = pts > 0;
                for (int j = 0; j < 6; ++j) w[j] = D6(warp_sum(av.a[j]));
            } else {
            Accum<D6, NA> acc;
            acc.fp = ins.p; acc.w_ang = cx.w_ang; acc.w_lin = cx.w_lin; acc.dump = nullptr; acc.dump_cap = 0;
            if (!bristle) {
                acc.reset(ACC_REGULARIZED);
                run_pairs_dual(sc, ins, ps, env, k, ei, n, lane, cx, cxv, acc, flags);
                int pts = acc.n_points;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) pts += __shfl_xor_sync(0xffffffffu, pts, o);
                contact = pts > 0;
                for (int j = 0; j < 6; ++j) w[j] = warp_sum(acc.a[j]);
            } else if constexpr (HB) {
                acc.reset(ACC_COP);
                run_pairs_dual(sc, ins, ps, env, k, ei, n, lane, cx, cxv, acc, flags);
                int pts = acc.n_points;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) pts += __shfl_xor_sync(0xffffffffu, pts, o);
                contact = pts > 0;
                if (contact) {
                    D6 c[10];
                    for (int j = 0; j < 10; ++j) c[j] = warp_sum(acc.a[j]);
                    const Vec3<D6> cop = mk<D6>(c[7] / c[6], c[8] / c[6], c[9] / c[6]);
                    acc.cop = cop;
                    acc.reset(ACC_STIFFNESS);
                    run_pairs_dual(sc, ins, ps, env, k, ei, n, lane, cx, cxv, acc, flags);
                    D6 K21[21], Sinv[6], Kh[36], s[6];
                    for (int j = 0; j < 21; ++j) K21[j] = warp_sum(acc.a[j]) * ins.p[1];
                    decompose_K_dual(K21, ins.p[6], Sinv, Kh);  // every lane redundantly: identical inputs, no broadcast needed
                    for (int j = 0; j < 6; ++j) s[j] = ld7(sv + 7 * j);
                    for (int i = 0; i < 6; ++i) {
                        D6 t = D6(0.0);
                        for (int j = 0; j < 6; ++j) t += Kh[6 * i + j] * s[j];
                        acc.delta[i] = Sinv[i] * t;
                    }
                    acc.reset(ACC_BRISTLE);
                    run_pairs_dual(sc, ins, ps, env, k, ei, n, lane, cx, cxv, acc, flags);
                    D6 f[6];
                    for (int j = 0; j < 6; ++j) f[j] = warp_sum(acc.a[j]);
                    const Vec3<D6> shift = cross(cop, mk<D6>(f[3], f[4], f[5]));
                    w[0] = c[0] + (f[0] + shift.x); w[1] = c[1] + (f[1] + shift.y); w[2] = c[2] + (f[2] + shift.z);
                    w[3] = c[3] + f[3]; w[4] = c[4] + f[4]; w[5] = c[5] + f[5];
                    if (lane == 0) {
                        const double ti = -(1.0 / ins.p[0]);
                        D6 sw[6];
                        for (int i = 0; i < 6; ++i) sw[i] = Sinv[i] * f[i];
                        for (int i = 0; i < 6; ++i) {
                            D6 t = D6(0.0);
                            for (int j = 0; j < 6; ++j) t += Kh[6 * i + j] * sw[j];
                            st7(sd + 7 * i, (t + s[i]) * ti);
                        }
                    }
                }
            }
            }   // seeded
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) flags |= __shfl_xor_sync(0xffffffffu, flags, o);
        if (lane == 0) {
            if (!contact) {
                for (int j = 0; j < 6; ++j) w[j] = D6(0.0);
                if (bristle) {  // no_contact!(::Bristle)
                    const double ti = -(1.0 / ins.p[0]);
                    for (int j = 0; j < 6; ++j) st7(sd + 7 * j, ld7(sv + 7 * j) * ti);
                }
            }
            double* wo = io.wrench7 + 42 * ei;
            for (int j = 0; j < 6; ++j) st7(wo + 7 * j, w[j]);
            io.flags[ei] = (io.flags[ei] & ~kFlagContact) | flags | (contact ? kFlagContact : 0);
        }
    }
}

}  // namespace

const unsigned* large_seg_start_ptr(const LargeBuffers* b);
const int3* large_sorted_ptr(const LargeBuffers* b);

cudaError_t launch_eval_dual6(const SceneDev& sc, long long n_env, const double* X7, const double* twist7, const double* s7, double* wrench7, double* sdot7,
                              const long long* n_pairs, int* flags, const unsigned* small_pairs, int small_cap, const LargeBuffers* lb,
                              const int32_t* large_index, int n_large, cudaStream_t stream) {
    DualIO io{n_env, X7, twist7, s7, wrench7, sdot7, n_pairs, flags};
    PairSource ps{small_pairs, small_cap, lb ? large_sorted_ptr(lb) : nullptr, lb ? large_seg_start_ptr(lb) : nullptr, large_index, n_large};
    const long long n_prob = n_env * sc.n_ins;
    if (n_prob == 0) return cudaSuccess;
    int n_sm = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
    const unsigned blocks = (unsigned)std::min<long long>((n_prob + 3) / 4, (long long)n_sm * 8);
    if (sc.n_bristle > 0) eval_dual6_kernel<true><<<blocks, 128, 0, stream>>>(sc, io, ps);
    else return launch_eval_dual6_chunked(sc, io, ps, n_prob, n_sm, stream);
    return cudaGetLastError();
}

}  // namespace pfc
