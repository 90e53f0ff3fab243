// pfc_small.cu -- fused contact-wrench evaluation for small instructions: ONE WARP PER
// (environment, contact instruction), everything on chip, one launch per batch.
//
// This is the batched-environments path (config C3: thousands of copies of test/boxes.jl).  What
// the reference does per instruction in force_single_elastic_intersection!
// (/root/reference/src/contact_algorithms_non_friction.jl:70-84) --
//   calcTriTetIntersections! (dual-tree traversal, src/obb/tree_types.jl:88-111)
//   integrate_over!          (loop over candidate pairs, :136-143)
//   yes_contact!/no_contact! (src/contact_algorithms_friction.jl:50-81, 119-143)
// -- is done by one warp of a persistent grid without leaving the SM:
//   1. Broad phase: warp-cooperative in-place expansion of the node-pair frontier held in shared
//      memory.  Each round every lane tests one node pair (15-axis SAT, bit-exact) and the
//      survivors' children replace it *in order* through a warp prefix sum, so when the frontier
//      holds only leaf pairs it is exactly the reference's recursion (DFS) order -- no atomics, no
//      sort, deterministic.
//   2. Narrow phase in chunks of 32 candidate pairs: (A) each lane clips one pair and, if a polygon
//      survives, leaves it in its shared-memory slot; (B) the chunk's (polygon, edge) sub-triangles
//      are dealt out one per lane for quadrature + friction.  Most pairs die in (A); (B) keeps the
//      lanes dense where the FLOPs are.
//   3. Fixed-order xor-butterfly warp reduction of the per-lane partial sums (bitwise
//      reproducible; no floating-point atomics).  Bristle friction runs the three passes of
//      bristle_wrench_in_world with a warp reduction between passes.
// HBM traffic per instruction is the boundary data only: 22 doubles in, 6 doubles + 2 words out.
#include "pfc_bristle.cuh"
#include "pfc_launch.h"
#include "pfc_patch.cuh"
#include "pfc_sat.cuh"

namespace pfc {

namespace {

constexpr unsigned kDone = 0x80000000u;
PFC_D unsigned enc(int a, int b) { return (unsigned(a) << 15) | unsigned(b); }
PFC_D int dec_a(unsigned e) { return int((e >> 15) & 0x7fffu); }
PFC_D int dec_b(unsigned e) { return int(e & 0x7fffu); }

PFC_D int warp_sum_int(int x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
PFC_D int warp_incl_scan(int x, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += v; }
    return x;
}

// Loads x_r2_r1 (col-major 4x4) into a row-major rotation + translation.
PFC_D void load_xform(const double* __restrict__ X, Xform<double>& x) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) x.r[3 * i + j] = X[4 * j + i];
        x.t[i] = X[12 + i];
    }
}

// x_r1_r2 for the broad phase with the reference's rounding: R' and -(R' * t), products summed left to right.
PFC_D void broad_phase_xform(const Xform<double>& x21, double* Rab, double* tab) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) Rab[3 * i + j] = x21.r[3 * j + i];
        tab[i] = -add_(add_(mul_(x21.r[i], x21.t[0]), mul_(x21.r[3 + i], x21.t[1])), mul_(x21.r[6 + i], x21.t[2]));
    }
}

// per-warp shared memory
template <int CAP> struct WarpSmem {
    PolyRec<double> poly[32];   // stage A output, one slot per lane
    unsigned frontier[2][CAP];
    unsigned char items[256];   // stage B work list: (slot << 3) | edge
    double bris[42];            // bristle: Sinv (6) + K̄^(-1/2) (36), kept across the friction pass
};

// Narrow phase over the warp's pair list for one accumulator mode.
template <int CAP>
PFC_D void run_pairs(const SceneDev& sc, const InsDev& ins, WarpSmem<CAP>& sm, const unsigned* pairs, int n, int lane, const PatchCtx<double>& cx,
                     Accum<double>& acc, int& flags) {
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        int nv = 0;
        if (i < n) {
            const unsigned e = pairs[i];
            if (clip_pair(sc, ins, dec_a(e), dec_b(e), cx, sm.poly[lane], flags)) nv = sm.poly[lane].n;
        }
        const int incl = warp_incl_scan(nv, lane);
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        for (int k = 0; k < nv; ++k) sm.items[incl - nv + k] = (unsigned char)((lane << 3) | k);
        __syncwarp();
        for (int it = lane; it < total; it += 32) {
            const int code = sm.items[it];
            const PolyRec<double>& pr = sm.poly[code >> 3];
            const int k = code & 7;
            const int kp = (k == 0) ? pr.n - 1 : k - 1;
            integrate_subtri(pr.v[kp], pr.v[k], pr.cen, pr.nrm, pr.eps_r, cx, acc);
        }
        __syncwarp();
    }
}

template <int WARPS, int CAP>
__global__ void __launch_bounds__(WARPS * 32) eval_small_f64_kernel(SceneDev sc, EvalIO io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    WarpSmem<CAP>& sm = reinterpret_cast<WarpSmem<CAP>*>(smem_raw)[wib];
    const long long n_prob = io.n_env * sc.n_small;
    for (long long prob = (long long)blockIdx.x * WARPS + wib; prob < n_prob; prob += (long long)gridDim.x * WARPS) {
        const long long env = prob / sc.n_small;
        const int k = sc.small_ins[prob - env * sc.n_small];
        const InsDev& ins = sc.ins[k];
        const long long ei = env * sc.n_ins + k;

        PatchCtx<double> cx;
        load_xform(io.X + 16 * ei, cx.x21);
        double Rab[9], tab[3];
        broad_phase_xform(cx.x21, Rab, tab);

        // ---- 1. broad phase ---------------------------------------------------------------------------
        unsigned* cur = sm.frontier[0];
        unsigned* nxt = sm.frontier[1];
        int n = 1;
        int flags = 0;
        if (lane == 0) cur[0] = enc(0, 0);
        __syncwarp();
        for (;;) {
            int n_out = 0;
            bool open = false;
            for (int base = 0; base < n; base += 32) {
                const int i = base + lane;
                int cnt = 0;
                unsigned c0 = 0, c1 = 0, c2 = 0, c3 = 0;
                if (i < n) {
                    const unsigned e = cur[i];
                    if (e & kDone) { cnt = 1; c0 = e; }
                    else {
                        const int ia = dec_a(e), ib = dec_b(e);
                        const NodeRec& a = sc.nodes[ins.node_base1 + ia];
                        const NodeRec& b = sc.nodes[ins.node_base2 + ib];
                        SatA A;
                        sat_prepare_a(a, Rab, tab, A);
                        if (sat_test(A, b)) {
                            const int al = a.left, ar = a.right, bl = b.left, br = b.right;
                            if (al < 0) {
                                if (bl < 0) { cnt = 1; c0 = kDone | enc(ar, br); }
                                else { cnt = 2; c0 = enc(ia, bl); c1 = enc(ia, br); open = true; }
                            } else if (bl < 0) { cnt = 2; c0 = enc(al, ib); c1 = enc(ar, ib); open = true; }
                            else { cnt = 4; c0 = enc(al, bl); c1 = enc(ar, bl); c2 = enc(al, br); c3 = enc(ar, br); open = true; }
                        }
                    }
                }
                const int incl = warp_incl_scan(cnt, lane);
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                const int at = n_out + incl - cnt;
                if (at + cnt <= CAP) {
                    if (cnt > 0) nxt[at] = c0;
                    if (cnt > 1) nxt[at + 1] = c1;
                    if (cnt > 2) { nxt[at + 2] = c2; nxt[at + 3] = c3; }
                } else if (cnt > 0) flags |= kFlagOverflow;
                n_out += total;
            }
            __syncwarp();
            unsigned* t = cur; cur = nxt; nxt = t;
            n = n_out < CAP ? n_out : CAP;
            if (!__any_sync(0xffffffffu, open)) break;
        }
        if (io.dbg_pairs) {
            int* out = io.dbg_pairs + 2 * (long long)io.dbg_cap * ei;
            for (int i = lane; i < n && i < io.dbg_cap; i += 32) { out[2 * i] = dec_a(cur[i]); out[2 * i + 1] = dec_b(cur[i]); }
        }

        // ---- 2 + 3. narrow phase, friction, fixed-order reduction ------------------------------------------
        double w[6] = {0, 0, 0, 0, 0, 0};
        bool contact = false;
        const double* sv = (ins.model == PFC_MODEL_BRISTLE) ? io.s + 6 * ((long long)sc.n_bristle * env + ins.bristle_id) : nullptr;
        double* sd = (ins.model == PFC_MODEL_BRISTLE) ? io.sdot + 6 * ((long long)sc.n_bristle * env + ins.bristle_id) : nullptr;
        if (n > 0) {
            cx.x12 = inverse(cx.x21);
            const double* tw = io.twist + 6 * ei;
            cx.w_ang = mk<double>(tw[0], tw[1], tw[2]);
            cx.w_lin = mk<double>(tw[3], tw[4], tw[5]);
            cx.chi = ins.chi; cx.Ebar1 = ins.Ebar1; cx.Ebar2 = ins.Ebar2; cx.n_quad = ins.n_quad;
            Accum<double> acc;
            acc.fp = ins.p; acc.w_ang = cx.w_ang; acc.w_lin = cx.w_lin; acc.dump = nullptr; acc.dump_cap = 0;
            if (ins.model == PFC_MODEL_REGULARIZED) {
                acc.reset(ACC_REGULARIZED);
                run_pairs(sc, ins, sm, cur, n, lane, cx, acc, flags);
                contact = warp_sum_int(acc.n_points) > 0;
#pragma unroll
                for (int j = 0; j < 6; ++j) w[j] = warp_sum(acc.a[j]);
            } else {
                acc.reset(ACC_COP);
                run_pairs(sc, ins, sm, cur, n, lane, cx, acc, flags);
                contact = warp_sum_int(acc.n_points) > 0;
                if (contact) {
                    double c[10];
#pragma unroll
                    for (int j = 0; j < 10; ++j) c[j] = warp_sum(acc.a[j]);
                    const Vec3<double> cop = mk<double>(c[7] / c[6], c[8] / c[6], c[9] / c[6]);
                    acc.cop = cop;
                    acc.reset(ACC_STIFFNESS);
                    run_pairs(sc, ins, sm, cur, n, lane, cx, acc, flags);
                    // lane 0 factors the 6x6 stiffness using the (now idle) polygon slots as scratch
                    double* scr = reinterpret_cast<double*>(sm.poly);
                    double* K21 = scr + 108;
                    double* Sinv = sm.bris;
                    double* Kh = sm.bris + 6;
#pragma unroll
                    for (int j = 0; j < 21; ++j) { const double v = warp_sum(acc.a[j]) * ins.p[1]; if (lane == 0) K21[j] = v; }
                    if (lane == 0) decompose_K(K21, ins.p[6], Sinv, Kh, scr);
                    __syncwarp();
                    double s[6];
                    for (int j = 0; j < 6; ++j) s[j] = sv[j];
                    for (int i = 0; i < 6; ++i) {
                        double t = 0.0;
                        for (int j = 0; j < 6; ++j) t += Kh[6 * i + j] * s[j];
                        acc.delta[i] = Sinv[i] * t;
                    }
                    acc.reset(ACC_BRISTLE);
                    run_pairs(sc, ins, sm, cur, n, lane, cx, acc, flags);
                    double f[6];
#pragma unroll
                    for (int j = 0; j < 6; ++j) f[j] = warp_sum(acc.a[j]);
                    const Vec3<double> lin = mk<double>(f[3], f[4], f[5]);
                    const Vec3<double> shift = cross(cop, lin);
                    w[0] = c[0] + (f[0] + shift.x); w[1] = c[1] + (f[1] + shift.y); w[2] = c[2] + (f[2] + shift.z);
                    w[3] = c[3] + f[3]; w[4] = c[4] + f[4]; w[5] = c[5] + f[5];
                    if (lane == 0) {
                        const double ti = -(1.0 / ins.p[0]);
                        double sw[6];
                        for (int i = 0; i < 6; ++i) sw[i] = Sinv[i] * f[i];
                        for (int i = 0; i < 6; ++i) {
                            double t = 0.0;
                            for (int j = 0; j < 6; ++j) t += Kh[6 * i + j] * sw[j];
                            sd[i] = ti * (t + s[i]);
                        }
                    }
                }
            }
        }
        flags = (int)__reduce_or_sync(0xffffffffu, (unsigned)flags);
        if (lane == 0) {
            if (!contact) {
#pragma unroll
                for (int j = 0; j < 6; ++j) w[j] = 0.0;
                if (ins.model == PFC_MODEL_BRISTLE) {  // no_contact!(::Bristle)
                    const double ti = -(1.0 / ins.p[0]);
                    for (int j = 0; j < 6; ++j) sd[j] = ti * sv[j];
                }
            }
            double* wo = io.wrench + 6 * ei;
#pragma unroll
            for (int j = 0; j < 6; ++j) wo[j] = w[j];
            io.n_pairs[ei] = n;
            io.flags[ei] = flags | (contact ? kFlagContact : 0);
        }
        __syncwarp();
    }
}

// One thread walks a given pair list in order and records every traction point (TractionCache).
__global__ void dump_traction_kernel(SceneDev sc, EvalIO io, long long env, int k, const int* pairs, long long n_pairs, double* out, int cap, int* n_points) {
    const InsDev& ins = sc.ins[k];
    const long long ei = env * sc.n_ins + k;
    PatchCtx<double> cx;
    load_xform(io.X + 16 * ei, cx.x21);
    cx.x12 = inverse(cx.x21);
    const double* tw = io.twist + 6 * ei;
    cx.w_ang = mk<double>(tw[0], tw[1], tw[2]);
    cx.w_lin = mk<double>(tw[3], tw[4], tw[5]);
    cx.chi = ins.chi; cx.Ebar1 = ins.Ebar1; cx.Ebar2 = ins.Ebar2; cx.n_quad = ins.n_quad;
    Accum<double> acc;
    acc.fp = ins.p; acc.w_ang = cx.w_ang; acc.w_lin = cx.w_lin; acc.dump = out; acc.dump_cap = cap;
    acc.reset(ACC_DUMP);
    int flags = 0;
    for (long long i = 0; i < n_pairs; ++i) integrate_pair(sc, ins, pairs[2 * i], pairs[2 * i + 1], cx, acc, flags);
    *n_points = acc.n_points;
}

int g_small_blocks = 0;

}  // namespace

cudaError_t launch_eval_small_f64(const SceneDev& sc, const EvalIO& io, cudaStream_t stream, int* n_launches) {
    const long long n_prob = io.n_env * sc.n_small;
    if (n_prob == 0) return cudaSuccess;
    auto kern = eval_small_f64_kernel<kSmallWarps, kSmallCap>;
    const size_t smem = sizeof(WarpSmem<kSmallCap>) * kSmallWarps;
    if (g_small_blocks == 0) {  // persistent grid: as many CTAs as can be resident on the device
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm = 0, dev = 0, n_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSmallWarps * 32, smem);
        if (e != cudaSuccess) return e;
        g_small_blocks = n_sm * (per_sm > 0 ? per_sm : 1);
    }
    long long blocks = (n_prob + kSmallWarps - 1) / kSmallWarps;
    if (blocks > g_small_blocks) blocks = g_small_blocks;
    kern<<<(unsigned)blocks, kSmallWarps * 32, smem, stream>>>(sc, io);
    if (n_launches) ++*n_launches;
    return cudaGetLastError();
}

cudaError_t launch_dump_traction(const SceneDev& sc, const EvalIO& io, long long env, int ins, const int* pairs, long long n_pairs, double* out,
                                 int cap_points, int* n_points, cudaStream_t stream) {
    dump_traction_kernel<<<1, 1, 0, stream>>>(sc, io, env, ins, pairs, n_pairs, out, cap_points, n_points);
    return cudaGetLastError();
}

}  // namespace pfc
