// pfc_radau.cu -- the dense linear algebra of the reference's Radau step, batched over environments.
//
// updateInvC! (/root/reference/src/radau/radau_functions.jl:88-99) forms, for every stage i, the complex matrix
//     C_i = (h^-1 lambda_i) I - J      (the reference keeps -J in neg_J and adds the shift on the diagonal)
// and inverts it with LAPACK getrf! + getri!.  For a batch of environments this is thousands of independent NX x NX complex
// inversions (NX = 48 for test/boxes.jl), far too small for a library call per matrix and slow through batched cuSOLVER
// (17.6 ms for 4096 matrices measured on B200).  Here one CTA inverts one matrix in shared memory: in-place Gauss-Jordan with
// partial (row) pivoting, the pivot search by one warp, the rank-1 update by all threads; the column permutation is undone at
// the end.  Input is the REAL matrix -J shared by the stages of an environment plus the complex shift; output is interleaved
// (re, im) row-major, which is the memory layout of a torch.complex128 / Julia ComplexF64 array.
#include <cuda_runtime.h>

#include "pfc_launch.h"

namespace pfc {

namespace {

struct cplx { double re, im; };
__device__ __forceinline__ cplx cmul(const cplx& a, const cplx& b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cplx cinv(const cplx& a) {   // Smith's algorithm (what LAPACK's zladiv guards against: overflow in |a|^2)
    if (fabs(a.re) >= fabs(a.im)) { const double r = a.im / a.re, d = a.re + a.im * r; return {1.0 / d, -r / d}; }
    const double r = a.re / a.im, d = a.re * r + a.im;
    return {r / d, -1.0 / d};
}

// grid: one CTA per matrix; dynamic shared memory: cplx A[n][n + 1] (odd-ish pitch) | int piv[n]
__global__ void __launch_bounds__(256) radau_inv_c_kernel(int n, const double* __restrict__ neg_J, const double* __restrict__ shift, const int* __restrict__ index,
                                                          double* __restrict__ inv_c, int* __restrict__ info) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int pitch = n + 1;
    cplx* A = reinterpret_cast<cplx*>(smem_raw);
    int* piv = reinterpret_cast<int*>(A + (size_t)n * pitch);
    __shared__ int s_p;
    __shared__ cplx s_pivinv;
    const long long mat = blockIdx.x;
    const long long src = index ? index[mat] : mat;             // environment whose -J this matrix is built from
    const double* J = neg_J + src * (long long)n * n;
    const cplx sh = {shift[2 * mat], shift[2 * mat + 1]};
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        const int i = e / n, j = e - i * n;
        A[i * pitch + j] = {J[e] + (i == j ? sh.re : 0.0), (i == j ? sh.im : 0.0)};
    }
    __syncthreads();
    bool singular = false;
    for (int k = 0; k < n; ++k) {
        if (threadIdx.x < 32) {   // pivot: the row i >= k with the largest |re| + |im| in column k (izamax's measure)
            double best = -1.0; int bi = k;
            for (int i = k + (int)threadIdx.x; i < n; i += 32) {
                const double v = fabs(A[i * pitch + k].re) + fabs(A[i * pitch + k].im);
                if (v > best) { best = v; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (threadIdx.x == 0) { s_p = bi; piv[k] = bi; s_pivinv = (best > 0.0) ? cinv(A[bi * pitch + k]) : cplx{0.0, 0.0}; if (!(best > 0.0)) singular = true; }
        }
        __syncthreads();
        const int p = s_p;
        if (p != k) for (int j = threadIdx.x; j < n; j += blockDim.x) { const cplx t = A[k * pitch + j]; A[k * pitch + j] = A[p * pitch + j]; A[p * pitch + j] = t; }
        __syncthreads();
        // scale the pivot row (its pivot entry becomes the inverse of the pivot), then eliminate column k from every other row
        const cplx pinv = s_pivinv;
        for (int j = threadIdx.x; j < n; j += blockDim.x) A[k * pitch + j] = (j == k) ? pinv : cmul(A[k * pitch + j], pinv);
        __syncthreads();
        for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
            const int i = e / n, j = e - i * n;
            if (i == k || j == k) continue;
            const cplx f = A[i * pitch + k], r = A[k * pitch + j];
            A[i * pitch + j].re -= f.re * r.re - f.im * r.im;
            A[i * pitch + j].im -= f.re * r.im + f.im * r.re;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x)
            if (i != k) { const cplx f = A[i * pitch + k]; const cplx t = cmul(f, pinv); A[i * pitch + k] = {-t.re, -t.im}; }
        __syncthreads();
    }
    for (int k = n - 1; k >= 0; --k) {   // undo the row interchanges as column interchanges, last first
        const int p = piv[k];
        if (p != k) for (int i = threadIdx.x; i < n; i += blockDim.x) { const cplx t = A[i * pitch + k]; A[i * pitch + k] = A[i * pitch + p]; A[i * pitch + p] = t; }
        __syncthreads();
    }
    double* out = inv_c + mat * 2ll * n * n;
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        const int i = e / n, j = e - i * n;
        out[2 * e] = A[i * pitch + j].re; out[2 * e + 1] = A[i * pitch + j].im;
    }
    if (threadIdx.x == 0 && singular && info) atomicOr(info, 1);
}

}  // namespace

cudaError_t launch_radau_inv_c(long long n_mat, int n, const double* neg_J, const double* shift, const int* index, double* inv_c, int* info, cudaStream_t stream) {
    if (n_mat == 0) return cudaSuccess;
    const size_t smem = sizeof(cplx) * (size_t)n * (n + 1) + sizeof(int) * n;
    {
        struct Tag {};
        std::lock_guard<std::mutex> g(launch_mutex());
        LaunchSlot& sl = launch_slot<Tag>();
        if (smem > sl.smem) {
            cudaError_t e = cudaFuncSetAttribute(radau_inv_c_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            sl.smem = smem;
        }
    }
    radau_inv_c_kernel<<<(unsigned)n_mat, 256, smem, stream>>>(n, neg_J, shift, index, inv_c, info);
    return cudaGetLastError();
}

}  // namespace pfc
