// pfc_dual_chunked.cu -- Jacobian mode of the regularized-friction instructions: the 6 partials travel in three chunks of 2.
//
// Radau obtains the Jacobian of the ODE right-hand side by forward-mode AD in chunks of 6
// (/root/reference/src/radau/radau_functions.jl:2-26, N_chunk = 6 at src/mechanism_scenario.jl:181):
// forceAllElasticIntersections! then runs on ForwardDiff.Dual{Nothing,Float64,6}, i.e. every scalar that depends on the state carries
// its value and 6 partials, while the candidate-pair lists come from the Float64 state (calcTriTetIntersections! always uses m.float,
// non_friction.jl:94-101).  The device code is the templated source of the Float64 path (pfc_clip.cuh, pfc_patch.cuh) on Dual<2>.
// Bristle instructions are skipped here: pfc_exact.cu evaluates them on Dual<6> in the reference's operation order.
#include <algorithm>

#include "pfc_dual.cuh"

namespace pfc {

namespace {

// ---- regularized-only scenes: the 6 partials in three chunks of 2 ---------------------------------------------------------------
// One sub-triangle on Dual<6> keeps ~240 doubles alive (its polygon, the twist, the accumulators, temporaries): twice the register
// file per thread, so the kernel above spills ~9 KB per thread and runs out of DRAM bandwidth (ncu: 6 GB of local-memory traffic per
// launch, 8 lanes of 32 busy).  A Dual<2> needs 3 doubles per scalar instead of 7: the same pipeline fits in registers, and carrying
// the 6 partials as 3 independent (pair, chunk) items triples the number of busy lanes.  The value part is recomputed by every chunk
// (9 instead of 7 units of work per pair); the three chunks of a pair produce bit-identical value parts, chunk 0's is the one stored.
typedef Dual<2> D2;
constexpr int kChunkSlots = 10;   // surviving pairs per round: 3 chunks x 10 pairs = 30 lanes

struct Chunk3Smem {
    PatchCtx<D2> cx[3];
    double acc[30][19];   // per lane: 6 sums x (value, 2 partials), padded to an odd stride
    int pts[30];
    double wout[42];      // the instruction's wrench: 6 components x (value, 6 partials)
};

__global__ void __launch_bounds__(128) eval_dual6_chunked_kernel(SceneDev sc, DualIO io, PairSource ps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Chunk3Smem& sm = reinterpret_cast<Chunk3Smem*>(smem_raw)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int chunk = lane / kChunkSlots, slot = lane - chunk * kChunkSlots;   // lanes 30, 31: chunk 3 = idle
    const long long n_prob = io.n_env * sc.n_ins;
    for (long long ei = (long long)blockIdx.x * 4 + (threadIdx.x >> 5); ei < n_prob; ei += (long long)gridDim.x * 4) {
        const long long env_v = ei / sc.n_ins;
        const int k = int(ei - env_v * sc.n_ins);
        const long long env = env_v % io.n_real, er = env * sc.n_ins + k;   // the real environment whose pair list this problem reads
        const InsDev& ins = sc.ins[k];
        if (ins.model != PFC_MODEL_REGULARIZED) continue;   // bristle: pfc_exact.cu
        const int n = (int)io.n_pairs[er];
        int flags = 0;
        bool contact = false;
        for (int j = lane; j < 42; j += 32) sm.wout[j] = 0.0;
        if (n > 0) {
            const double* Xp = io.X7 + 112 * ei;
            const double* tw = io.twist7 + 42 * ei;
            // is any input of this instruction seeded?  (lanes share the 18 x 6 partials)
            bool mine = false;
            for (int e = lane; e < 18 * 6; e += 32) {
                const int sc_i = e / 6, q = e - sc_i * 6;
                const double* base = sc_i < 12 ? Xp + 7 * (sc_i < 9 ? (4 * (sc_i % 3) + sc_i / 3) : (12 + sc_i - 9)) : tw + 7 * (sc_i - 12);
                mine |= (base[1 + q] != 0.0);
            }
            const bool seeded = __any_sync(0xffffffffu, mine);
            __syncwarp();
            if (lane < 3) {   // chunk `lane`'s context: partials 2 lane and 2 lane + 1
                PatchCtx<D2>& cx = sm.cx[lane];
                auto ld = [&](const double* p7) { D2 r; r.v = p7[0]; r.p[0] = p7[1 + 2 * lane]; r.p[1] = p7[2 + 2 * lane]; return r; };
#pragma unroll
                for (int i = 0; i < 3; ++i) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) cx.x21.r[3 * i + j] = ld(Xp + 7 * (4 * j + i));
                    cx.x21.t[i] = ld(Xp + 7 * (12 + i));
                }
                cx.x12 = inverse(cx.x21);
                cx.w_ang = mk<D2>(ld(tw), ld(tw + 7), ld(tw + 14));
                cx.w_lin = mk<D2>(ld(tw + 21), ld(tw + 28), ld(tw + 35));
                cx.chi = ins.chi; cx.Ebar1 = ins.Ebar1; cx.Ebar2 = ins.Ebar2; cx.n_quad = ins.n_quad;
            }
            __syncwarp();
            PatchCtx<double> cxv;
            {
                const PatchCtx<D2>& c0 = sm.cx[0];
#pragma unroll
                for (int i = 0; i < 9; ++i) { cxv.x21.r[i] = c0.x21.r[i].v; cxv.x12.r[i] = c0.x12.r[i].v; }
#pragma unroll
                for (int i = 0; i < 3; ++i) { cxv.x21.t[i] = c0.x21.t[i].v; cxv.x12.t[i] = c0.x12.t[i].v; }
                cxv.w_ang = mk<double>(c0.w_ang.x.v, c0.w_ang.y.v, c0.w_ang.z.v);
                cxv.w_lin = mk<double>(c0.w_lin.x.v, c0.w_lin.y.v, c0.w_lin.z.v);
                cxv.chi = c0.chi; cxv.Ebar1 = c0.Ebar1; cxv.Ebar2 = c0.Ebar2; cxv.n_quad = c0.n_quad;
            }
            const unsigned* pl_s = ins.small ? ps.small_pairs + (size_t)ps.small_cap * er : nullptr;
            const int3* pl_l = ins.small ? nullptr : ps.large_sorted + ps.seg_start[env * ps.n_large + ps.large_index[k]];
            // An instruction none of whose inputs depends on the seeded state entries (every partial of x_r2_r1 and the twist is zero -- e.g. the
            // seeds sit on another body) has zero wrench partials: its Dual evaluation is the Float64 evaluation, 7x cheaper.  The reference
            // evaluates such instructions on Duals all the same; the values agree to rounding.
            if (!seeded) {
                Accum<double, 6> av;
                av.fp = ins.p; av.w_ang = cxv.w_ang; av.w_lin = cxv.w_lin; av.dump = nullptr; av.dump_cap = 0;
                av.reset(ACC_REGULARIZED);
                for (int i = lane; i < n; i += 32) {
                    int a, b;
                    if (ins.small) { const unsigned e = pl_s[i]; a = int((e >> 15) & 0x7fffu); b = int(e & 0x7fffu); }
                    else { const int3 e = pl_l[i]; a = e.y; b = e.z; }
                    if (prefilter_pair(sc, ins, a, b, cxv)) integrate_pair(sc, ins, a, b, cxv, av, flags);
                }
                int pts = av.n_points;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) pts += __shfl_xor_sync(0xffffffffu, pts, o);
                contact = pts > 0;
                double wv[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) wv[j] = warp_sum(av.a[j]);
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < 6; ++j) sm.wout[7 * j] = wv[j];   // value slots; the partial slots stay 0
                }
            } else {
                Accum<D2, 6> acc;
                const PatchCtx<D2>& cx = sm.cx[chunk < 3 ? chunk : 0];
                acc.fp = ins.p; acc.w_ang = cx.w_ang; acc.w_lin = cx.w_lin; acc.dump = nullptr; acc.dump_cap = 0;
                acc.reset(ACC_REGULARIZED);
                for (int base = 0; base < n; base += 32) {
                    const int i = base + lane;
                    int a = 0, b = 0;
                    bool keep = false;
                    if (i < n) {
                        if (ins.small) { const unsigned e = pl_s[i]; a = int((e >> 15) & 0x7fffu); b = int(e & 0x7fffu); }
                        else { const int3 e = pl_l[i]; a = e.y; b = e.z; }
                        keep = survives_f64(sc, ins, a, b, cxv);
                    }
                    unsigned m = __ballot_sync(0xffffffffu, keep);
                    while (m) {   // rounds of kChunkSlots survivors, each handled by three lanes (one per chunk of partials)
                        const unsigned src = __fns(m, 0, slot + 1);            // the lane that holds survivor `slot` of this round
                        const int pa = __shfl_sync(0xffffffffu, a, src & 31u), pb = __shfl_sync(0xffffffffu, b, src & 31u);
                        if (chunk < 3 && src <= 31u) integrate_pair(sc, ins, pa, pb, cx, acc, flags);
#pragma unroll 1
                        for (int r = 0; r < kChunkSlots && m; ++r) m &= m - 1;
                    }
                }
                // per-lane sums -> shared memory -> one lane per output scalar adds its ten contributions in slot order
                __syncwarp();
                if (lane < 30) {
#pragma unroll
                    for (int j = 0; j < 6; ++j) { sm.acc[lane][3 * j] = acc.a[j].v; sm.acc[lane][3 * j + 1] = acc.a[j].p[0]; sm.acc[lane][3 * j + 2] = acc.a[j].p[1]; }
                    sm.pts[lane] = acc.n_points;
                }
                __syncwarp();
                int pts = 0;
                for (int r = 0; r < kChunkSlots; ++r) pts += sm.pts[r];
                contact = pts > 0;
                for (int j = lane; j < 42; j += 32) {
                    const int comp = j / 7, which = j - 7 * comp;
                    const int c = which == 0 ? 0 : (which - 1) / 2, col = which == 0 ? 0 : 1 + (which - 1) % 2;
                    double sum = 0.0;
                    for (int r = 0; r < kChunkSlots; ++r) sum += sm.acc[c * kChunkSlots + r][3 * comp + col];
                    sm.wout[j] = sum;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) flags |= __shfl_xor_sync(0xffffffffu, flags, o);
        __syncwarp();
        double* wo = io.wrench7 + 42 * ei;
        for (int j = lane; j < 42; j += 32) wo[j] = contact ? sm.wout[j] : 0.0;
        if (lane == 0) {
            if (io.n_real == io.n_env) io.flags[er] = (io.flags[er] & ~kFlagContact) | flags | (contact ? kFlagContact : 0);
            else if (flags | (contact ? kFlagContact : 0)) atomicOr(&io.flags[er], flags | (contact ? kFlagContact : 0));   // several chunks share the word
        }
        __syncwarp();
    }
}

}  // namespace

const unsigned* large_seg_start_ptr(const LargeBuffers* b);
const int3* large_sorted_ptr(const LargeBuffers* b);

cudaError_t launch_eval_dual6(const SceneDev& sc, long long n_env, const double* X7, const double* twist7, const double* s7, double* wrench7, double* sdot7,
                              const long long* n_pairs, int* flags, const unsigned* small_pairs, int small_cap, const LargeBuffers* lb,
                              const int32_t* large_index, int n_large, cudaStream_t stream, long long n_real) {
    DualIO io{n_env, X7, twist7, s7, wrench7, sdot7, n_pairs, flags, n_real > 0 ? n_real : n_env};
    PairSource ps{small_pairs, small_cap, lb ? large_sorted_ptr(lb) : nullptr, lb ? large_seg_start_ptr(lb) : nullptr, large_index, n_large};
    const long long n_prob = n_env * sc.n_ins;
    if (n_prob == 0) return cudaSuccess;
    int n_sm = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
    const unsigned blocks = (unsigned)std::min<long long>((n_prob + 3) / 4, (long long)n_sm * 4);
    eval_dual6_chunked_kernel<<<blocks, 128, sizeof(Chunk3Smem) * 4, stream>>>(sc, io, ps);
    return cudaGetLastError();
}

}  // namespace pfc
