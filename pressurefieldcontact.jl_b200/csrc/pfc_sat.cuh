// pfc_sat.cuh -- oriented-box overlap test of the broad phase.
//
// Follows BB_BB_intersect (/root/reference/src/obb/bb_intersection.jl:2-74): compose
// inv(OBB_a) * X_a_b * OBB_b, add 1e-14 to |R|, then the 15 separating-axis tests of Ericson's
// table 4.1 with strict '<'.  The boolean must be bit-exact against the reference, so every
// product/sum below is an explicitly rounded, never-contracted operation in the reference's
// evaluation order (StaticArrays row-by-column sums, left to right).  Terms that multiply the
// constant 0/1 bottom row of a homogeneous matrix are dropped: adding +-0.0 cannot change a value
// that is later compared or passed through abs().
#pragma once
#include "pfc_math.cuh"
#include "pfc_types.cuh"

namespace pfc {

// M = inv(OBB_a) * X_a_b : the part of the composition that depends on node a only.
struct SatA { double R[9]; double t[3]; double e[3]; };

// Axis-aligned nodes (kind == kNodeInternalAabb: R == I exactly).  The reference multiplies by the identity all the same; with R = I every
// product is x * 1 or x * (+-0) and every sum is x + (+-0), so the composed entries are the operands themselves -- bit for bit for finite
// inputs, up to the sign of a zero, which neither fabs() nor '<' can see.  The identity products are therefore skipped (AABB = false keeps
// the general path: tests/test_device_sat_on_host.py checks that both give the same boolean everywhere).
template <bool AABB = true> PFC_D void sat_prepare_a(const NodeRec& a, const double* __restrict__ Rab, const double* __restrict__ tab, SatA& out) {
    if (AABB && a.kind == kNodeInternalAabb) {   // inv(OBB_a) = [I, -c]
#pragma unroll
        for (int k = 0; k < 9; ++k) out.R[k] = Rab[k];
#pragma unroll
        for (int i = 0; i < 3; ++i) out.t[i] = add_(tab[i], -a.c[i]);
        out.e[0] = a.e[0]; out.e[1] = a.e[1]; out.e[2] = a.e[2];
        return;
    }
    // i_dh_a = [a.R' , (-a.R') * a.c]
    double mt[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)  // row i of a.R' is column i of a.R
        mt[i] = add_(add_(mul_(-a.R[0 + i], a.c[0]), mul_(-a.R[3 + i], a.c[1])), mul_(-a.R[6 + i], a.c[2]));
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double r0 = a.R[0 + i], r1 = a.R[3 + i], r2 = a.R[6 + i];
#pragma unroll
        for (int j = 0; j < 3; ++j) out.R[3 * i + j] = add_(add_(mul_(r0, Rab[0 + j]), mul_(r1, Rab[3 + j])), mul_(r2, Rab[6 + j]));
        out.t[i] = add_(add_(add_(mul_(r0, tab[0]), mul_(r1, tab[1])), mul_(r2, tab[2])), mt[i]);
    }
    out.e[0] = a.e[0]; out.e[1] = a.e[1]; out.e[2] = a.e[2];
}

// Rab/tab: row-major rotation and translation of x_r1_r2 (frame of tree 2 expressed in tree 1).
template <bool AABB = true> PFC_D bool sat_test(const SatA& A, const NodeRec& b) {
    double R[9], aR[9], t[3];
    const bool b_aabb = AABB && b.kind == kNodeInternalAabb;   // OBB_b = [I, c]: the rotation product is the left operand itself
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            R[3 * i + j] = b_aabb ? A.R[3 * i + j]
                                  : add_(add_(mul_(A.R[3 * i], b.R[0 + j]), mul_(A.R[3 * i + 1], b.R[3 + j])), mul_(A.R[3 * i + 2], b.R[6 + j]));
            aR[3 * i + j] = add_(fabs(R[3 * i + j]), 1.0e-14);
        }
        t[i] = add_(add_(add_(mul_(A.R[3 * i], b.c[0]), mul_(A.R[3 * i + 1], b.c[1])), mul_(A.R[3 * i + 2], b.c[2])), A.t[i]);
    }
    const double ea0 = A.e[0], ea1 = A.e[1], ea2 = A.e[2], eb0 = b.e[0], eb1 = b.e[1], eb2 = b.e[2];
    // All 15 axes are evaluated and OR-ed: the outcome equals the reference's early-exit chain, but
    // the 15 short dependency chains are independent, which is what hides the FP64 latency.
    bool sep = false;
    // face test 1/2: r_a = e_a, r_b = abs_R * e_b
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double rb = add_(add_(mul_(aR[3 * i], eb0), mul_(aR[3 * i + 1], eb1)), mul_(aR[3 * i + 2], eb2));
        sep |= add_(A.e[i], rb) < fabs(t[i]);
    }
    // face test 2/2: T = |R' t|, r_a = abs_R' * e_a, r_b = e_b
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double T = fabs(add_(add_(mul_(R[j], t[0]), mul_(R[3 + j], t[1])), mul_(R[6 + j], t[2])));
        const double ra = add_(add_(mul_(aR[j], ea0), mul_(aR[3 + j], ea1)), mul_(aR[6 + j], ea2));
        sep |= add_(ra, b.e[j]) < T;
    }
    // edge-edge tests.  R0/R1/R2 are the rows of R; s100(v) = (v1, v0, v0), s221(v) = (v2, v2, v1).
    const double eb100[3] = {eb1, eb0, eb0}, eb221[3] = {eb2, eb2, eb1};
    const int i100[3] = {1, 0, 0}, i221[3] = {2, 2, 1};
    // cross 1/3: a0 x b_j
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double T = fabs(sub_(mul_(t[2], R[3 + j]), mul_(t[1], R[6 + j])));
        const double ra = add_(mul_(ea1, aR[6 + j]), mul_(ea2, aR[3 + j]));
        const double rb = add_(mul_(eb100[j], aR[i221[j]]), mul_(eb221[j], aR[i100[j]]));
        sep |= add_(ra, rb) < T;
    }
    // cross 2/3: a1 x b_j
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double T = fabs(sub_(mul_(t[0], R[6 + j]), mul_(t[2], R[j])));
        const double ra = add_(mul_(ea0, aR[6 + j]), mul_(ea2, aR[j]));
        const double rb = add_(mul_(eb100[j], aR[3 + i221[j]]), mul_(eb221[j], aR[3 + i100[j]]));
        sep |= add_(ra, rb) < T;
    }
    // cross 3/3: a2 x b_j
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double T = fabs(sub_(mul_(t[1], R[j]), mul_(t[0], R[3 + j])));
        const double ra = add_(mul_(ea0, aR[3 + j]), mul_(ea1, aR[j]));
        const double rb = add_(mul_(eb100[j], aR[6 + i221[j]]), mul_(eb221[j], aR[6 + i100[j]]));
        sep |= add_(ra, rb) < T;
    }
    return !sep;
}

}  // namespace pfc
