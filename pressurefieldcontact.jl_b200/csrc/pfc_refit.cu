// pfc_refit.cu -- device-side refit of one mesh after its vertices moved (the refit half).
//
// With the connectivity and the tree TOPOLOGY kept, new vertex positions change three things the kernels read:
//   * the primitive records: TriRec (vertices, unit normal) / TetRec (vertices, inv([V; 1]), pressure gradient), which the
//     reference recomputes for every candidate pair (src/contact_algorithms_non_friction.jl:145-164) and pfc_finalize precomputes;
//   * the leaf boxes: fit_tri_obb / fit_tet_obb (src/obb/obb_construction.jl:1-53; tight_fit_leaves!, src/geometry/blob_types.jl:170-190);
//   * the internal boxes: OBB(a, b) of the two children's axis-aligned boxes (src/obb/box_types.jl:11-15), bottom-up; the leaves
//     enter with calc_obb of their vertices, as in eMesh_to_tree (src/geometry/blob_types.jl:136-168, recursive_top_down).
// One thread per primitive, one thread per leaf, and one launch per tree level from the deepest one up.
#include <cuda_runtime.h>

#include "pfc_launch.h"
#include "pfc_math.cuh"

namespace pfc {

namespace {

PFC_D void sub3(const double* a, const double* b, double* o) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
PFC_D void cross3r(const double* a, const double* b, double* o) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; }
PFC_D void normalize3(double* a) { const double l = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); a[0] /= l; a[1] /= l; a[2] /= l; }

PFC_D double tet_volume6_d(const double* v) {
    const double* a = v; const double* b = v + 3; const double* c = v + 6; const double* d = v + 9;
    double V = (b[0] - a[0]) * (c[1] * d[2] - c[2] * d[1]);
    V += (b[1] - a[1]) * (c[2] * d[0] - c[0] * d[2]);
    V += (b[2] - a[2]) * (c[0] * d[1] - c[1] * d[0]);
    V += (c[0] - d[0]) * (a[2] * b[1] - a[1] * b[2]);
    V += (c[1] - d[1]) * (a[0] * b[2] - a[2] * b[0]);
    V += (c[2] - d[2]) * (a[1] * b[0] - a[0] * b[1]);
    return V;
}

// inverse of [v0 v1 v2 v3; 1 1 1 1] by the adjugate (the arithmetic of invert_tet_matrix in pfc_api.cu)
PFC_D bool invert_tet_matrix_d(const double* v, double* inv) {
    double a[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) { a[0 + j] = v[3 * j]; a[4 + j] = v[3 * j + 1]; a[8 + j] = v[3 * j + 2]; a[12 + j] = 1.0; }
    const double s0 = a[0] * a[5] - a[4] * a[1], s1 = a[0] * a[6] - a[4] * a[2], s2 = a[0] * a[7] - a[4] * a[3];
    const double s3 = a[1] * a[6] - a[5] * a[2], s4 = a[1] * a[7] - a[5] * a[3], s5 = a[2] * a[7] - a[6] * a[3];
    const double c5 = a[10] * a[15] - a[14] * a[11], c4 = a[9] * a[15] - a[13] * a[11], c3 = a[9] * a[14] - a[13] * a[10];
    const double c2 = a[8] * a[15] - a[12] * a[11], c1 = a[8] * a[14] - a[12] * a[10], c0 = a[8] * a[13] - a[12] * a[9];
    const double det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
    if (!(det != 0.0) || !isfinite(det)) return false;
    const double id = 1.0 / det;
    inv[0] = (a[5] * c5 - a[6] * c4 + a[7] * c3) * id;    inv[1] = (-a[1] * c5 + a[2] * c4 - a[3] * c3) * id;
    inv[2] = (a[13] * s5 - a[14] * s4 + a[15] * s3) * id; inv[3] = (-a[9] * s5 + a[10] * s4 - a[11] * s3) * id;
    inv[4] = (-a[4] * c5 + a[6] * c2 - a[7] * c1) * id;   inv[5] = (a[0] * c5 - a[2] * c2 + a[3] * c1) * id;
    inv[6] = (-a[12] * s5 + a[14] * s2 - a[15] * s1) * id; inv[7] = (a[8] * s5 - a[10] * s2 + a[11] * s1) * id;
    inv[8] = (a[4] * c4 - a[5] * c2 + a[7] * c0) * id;    inv[9] = (-a[0] * c4 + a[1] * c2 - a[3] * c0) * id;
    inv[10] = (a[12] * s4 - a[13] * s2 + a[15] * s0) * id; inv[11] = (-a[8] * s4 + a[9] * s2 - a[11] * s0) * id;
    inv[12] = (-a[4] * c3 + a[5] * c1 - a[6] * c0) * id;  inv[13] = (a[0] * c3 - a[1] * c1 + a[2] * c0) * id;
    inv[14] = (-a[12] * s3 + a[13] * s1 - a[14] * s0) * id; inv[15] = (a[8] * s3 - a[9] * s1 + a[10] * s0) * id;
    return true;
}

__global__ void refit_prims_kernel(int kind, long long n_prim, const int* __restrict__ idx, const double* __restrict__ eps, const double* __restrict__ xyz,
                                   TriRec* tris, TetRec* tets, int* err) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n_prim; k += (long long)gridDim.x * blockDim.x) {
        if (kind == 0) {
            TriRec t;
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int i = 0; i < 3; ++i) t.v[3 * j + i] = xyz[3 * idx[3 * k + j] + i];
            double a[3], b[3], n[3];
            sub3(t.v + 3, t.v, a); sub3(t.v + 6, t.v + 3, b); cross3r(a, b, n);
            n[0] *= 0.5; n[1] *= 0.5; n[2] *= 0.5;
            normalize3(n);
            t.n[0] = n[0]; t.n[1] = n[1]; t.n[2] = n[2];
            tris[k] = t;
        } else {
            TetRec t;
            double e4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int vi = idx[4 * k + j];
#pragma unroll
                for (int i = 0; i < 3; ++i) t.v[3 * j + i] = xyz[3 * vi + i];
                e4[j] = eps[vi];
            }
            if (!(0.0 < tet_volume6_d(t.v)) || !invert_tet_matrix_d(t.v, t.inv)) { atomicOr(err, 1); continue; }   // inverted / degenerate tetrahedron
#pragma unroll
            for (int j = 0; j < 4; ++j) t.eps_r[j] = e4[0] * t.inv[j] + e4[1] * t.inv[4 + j] + e4[2] * t.inv[8 + j] + e4[3] * t.inv[12 + j];
            tets[k] = t;
        }
    }
}

// make_obb (src/obb/obb_construction.jl): axes e1 = edge i_start -> i_next, e3 = normal of the first three points, e2 = e3 x e1; extents
// from the projections of all n points.  Returns the box in (c, e, R) with R's columns the axes; area = 8 (e0 e1 + e1 e2 + e2 e0).
PFC_D double make_obb_d(const double (*p)[3], int n, int i_start, double* c, double* e, double* R) {
    const int i_next = (i_start % 3) + 1;
    double e1[3], e2[3], e3[3], a[3], b[3];
    sub3(p[i_next - 1], p[i_start - 1], e1); normalize3(e1);
    sub3(p[1], p[0], a); sub3(p[2], p[1], b); cross3r(a, b, e3);
    e3[0] *= 0.5; e3[1] *= 0.5; e3[2] *= 0.5;
    normalize3(e3);
    cross3r(e3, e1, e2);
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int k = 0; k < n; ++k) {
        const double pr[3] = {p[k][0] * e1[0] + p[k][1] * e1[1] + p[k][2] * e1[2], p[k][0] * e2[0] + p[k][1] * e2[1] + p[k][2] * e2[2],
                              p[k][0] * e3[0] + p[k][1] * e3[1] + p[k][2] * e3[2]};
#pragma unroll
        for (int i = 0; i < 3; ++i) { lo[i] = fmin(lo[i], pr[i]); hi[i] = fmax(hi[i], pr[i]); }
    }
    double cc[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { cc[i] = (hi[i] + lo[i]) * 0.5; e[i] = (hi[i] - lo[i]) * 0.5; }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        c[i] = e1[i] * cc[0] + e2[i] * cc[1] + e3[i] * cc[2];
        R[3 * i] = e1[i]; R[3 * i + 1] = e2[i]; R[3 * i + 2] = e3[i];    // row-major, columns = axes (NodeRec layout)
    }
    return 8.0 * (e[0] * e[1] + e[1] * e[2] + e[2] * e[0]);
}

// one thread per node: leaves get their tight box and the axis-aligned box of their vertices (what the parent merges)
__global__ void refit_leaves_kernel(int kind, long long n_node, NodeRec* nodes, const int* __restrict__ idx, const double* __restrict__ eps,
                                    const double* __restrict__ xyz, double* aabb) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n_node; k += (long long)gridDim.x * blockDim.x) {
        NodeRec& nd = nodes[k];
        if (nd.kind >= 0) continue;
        const int prim = nd.right;
        const int w = kind == 0 ? 3 : 4;
        double p[4][3], ev[4] = {0, 0, 0, 0};
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int j = 0; j < w; ++j) {
            const int vi = idx[w * prim + j];
            for (int i = 0; i < 3; ++i) { p[j][i] = xyz[3 * vi + i]; lo[i] = fmin(lo[i], p[j][i]); hi[i] = fmax(hi[i], p[j][i]); }
            if (kind == 1) ev[j] = eps[vi];
        }
        for (int i = 0; i < 3; ++i) { aabb[6 * k + i] = lo[i]; aabb[6 * k + 3 + i] = hi[i]; }
        if (kind == 0) { make_obb_d(p, 3, 1, nd.c, nd.e, nd.R); continue; }
        // fit_tet_obb: put the vertex with the largest |eps| last (the permutations of obb_construction.jl), try the three base edges
        int im = 0;
        for (int j = 1; j < 4; ++j) if (fabs(ev[j]) > fabs(ev[im])) im = j;
        const int PERM[4][4] = {{1, 3, 2, 0}, {3, 0, 2, 1}, {0, 3, 1, 2}, {0, 1, 2, 3}};
        double q[4][3];
        for (int j = 0; j < 4; ++j) for (int i = 0; i < 3; ++i) q[j][i] = p[PERM[im][j]][i];
        double c[3][3], e[3][3], R[3][9], area[3];
        for (int s = 0; s < 3; ++s) area[s] = make_obb_d(q, 4, s + 1, c[s], e[s], R[s]);
        const int pick = (fmax(area[1], area[2]) <= area[0]) ? 0 : ((fmax(area[0], area[2]) <= area[1]) ? 1 : 2);
        for (int i = 0; i < 3; ++i) { nd.c[i] = c[pick][i]; nd.e[i] = e[pick][i]; }
        for (int i = 0; i < 9; ++i) nd.R[i] = R[pick][i];
    }
}

// one thread per internal node of one level: OBB(a, b) of the children's axis-aligned boxes, through (centre, extent) like the reference
__global__ void refit_level_kernel(const int* __restrict__ level_nodes, int n, NodeRec* nodes, double* aabb) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const int k = level_nodes[t];
        NodeRec& nd = nodes[k];
        const double* A = aabb + 6 * (long long)(k + 1);   // pre-order: child 1 is the next record
        const double* B = aabb + 6 * (long long)nd.right;
        for (int i = 0; i < 3; ++i) {
            const double c1 = (A[3 + i] + A[i]) * 0.5, e1 = (A[3 + i] - A[i]) * 0.5, c2 = (B[3 + i] + B[i]) * 0.5, e2 = (B[3 + i] - B[i]) * 0.5;
            const double lo = fmin(fmin(c1 - e1, c1 + e1), fmin(c2 - e2, c2 + e2)), hi = fmax(fmax(c1 - e1, c1 + e1), fmax(c2 - e2, c2 + e2));
            aabb[6 * (long long)k + i] = lo; aabb[6 * (long long)k + 3 + i] = hi;
            nd.c[i] = (hi + lo) * 0.5; nd.e[i] = (hi - lo) * 0.5;
            nd.R[3 * i] = i == 0 ? 1.0 : 0.0; nd.R[3 * i + 1] = i == 1 ? 1.0 : 0.0; nd.R[3 * i + 2] = i == 2 ? 1.0 : 0.0;
        }
        nd.kind = kNodeInternalAabb;
    }
}

}  // namespace

cudaError_t launch_refit(int kind, long long n_prim, long long n_node, const int* idx, const double* eps, const double* xyz, TriRec* tris, TetRec* tets,
                         NodeRec* nodes, double* aabb, const int* level_nodes, const int* level_ptr, int n_level, int* err, cudaStream_t stream,
                         int* n_launches) {
    const unsigned gp = (unsigned)((n_prim + 127) / 128 < 148 * 16 ? (n_prim + 127) / 128 : 148 * 16);
    const unsigned gn = (unsigned)((n_node + 127) / 128 < 148 * 16 ? (n_node + 127) / 128 : 148 * 16);
    refit_prims_kernel<<<gp, 128, 0, stream>>>(kind, n_prim, idx, eps, xyz, tris, tets, err);
    refit_leaves_kernel<<<gn, 128, 0, stream>>>(kind, n_node, nodes, idx, eps, xyz, aabb);
    for (int l = n_level - 1; l >= 0; --l) {   // deepest internal level first
        const int n = level_ptr[l + 1] - level_ptr[l];
        if (n == 0) continue;
        refit_level_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(level_nodes + level_ptr[l], n, nodes, aabb);
        if (n_launches) *n_launches += 1;
    }
    if (n_launches) *n_launches += 2;
    return cudaGetLastError();
}

}  // namespace pfc
