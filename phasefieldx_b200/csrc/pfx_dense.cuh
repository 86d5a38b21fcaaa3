// pfx_dense.cuh — inverse of a dense symmetric positive definite matrix on the device (fp64, column-major).
//
// The coarse matrices Z^T A Z of the two-level PCG (pfx_coarse.cuh; at most kCoarseMax = 6144 rows per region) are
// inverted once per refresh.  Round 1 called cuSOLVER potrf / potri for that; this is the in-tree replacement, so that
// no library kernel runs on the path: blocked right-looking Cholesky A = L L^T (64-wide panels), Y = L^-1 by blocked
// forward substitution, A^-1 = Y^T Y (lower triangle).  Everything except the 64 x 64 diagonal blocks is one tiled
// DGEMM kernel (64 x 64 output tiles, 16-deep k-tiles in shared memory, 4 x 4 register blocking: ~n^3 * 1.5 flops at
// a few TFLOP/s, ~0.1 s for n = 6144, three or four times per run).  Fixed summation order: bit-reproducible.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pfx {

constexpr int kDenseNB = 64;

// C[m x n] = beta C + alpha op(A)[m x k] op(B)[k x n]   (column-major; TA / TB: operand stored transposed).
// tri = 1: only tiles on or below the diagonal of C (tile row >= tile column) are computed;
// tri = 2: as 1, and the k-loop starts at the first row of the later of the two tiles (Y^T Y with lower-triangular Y).
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) dense_gemm_kernel(int m, int n, int k, double alpha, const double* __restrict__ A, int64_t lda,
                                                         const double* __restrict__ B, int64_t ldb, double beta,
                                                         double* __restrict__ C, int64_t ldc, int tri) {
    __shared__ double As[16][kDenseNB + 1];
    __shared__ double Bs[16][kDenseNB + 1];
    const int tr = blockIdx.x, tc = blockIdx.y;
    if (tri && tc > tr) return;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int row0 = tr * kDenseNB, col0 = tc * kDenseNB;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    const int kbeg = (tri == 2) ? max(row0, col0) / 16 * 16 : 0;
    for (int k0 = kbeg; k0 < k; k0 += 16) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int e = tid + t * 256;                 // 1024 elements per tile
            {   // A tile: rows row0 + (e % 64), k index k0 + (e / 64)
                const int r = e & 63, kk = e >> 6;
                const int gr = row0 + r, gk = k0 + kk;
                double v = 0.0;
                if (gr < m && gk < k) v = TA ? A[(int64_t)gk + (int64_t)gr * lda] : A[(int64_t)gr + (int64_t)gk * lda];
                As[kk][r] = v;
            }
            {   // B tile: k index k0 + (e % 16), cols col0 + (e / 16)
                const int kk = e & 15, cc = e >> 4;
                const int gk = k0 + kk, gc = col0 + cc;
                double v = 0.0;
                if (gk < k && gc < n) v = TB ? B[(int64_t)gc + (int64_t)gk * ldb] : B[(int64_t)gk + (int64_t)gc * ldb];
                Bs[kk][cc] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][tx * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][ty * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gr = row0 + tx * 4 + i, gc = col0 + ty * 4 + j;
            if (gr < m && gc < n) {
                double* p = C + (int64_t)gr + (int64_t)gc * ldc;
                *p = (beta == 0.0 ? 0.0 : beta * (*p)) + alpha * acc[i][j];
            }
        }
}

// Diagonal block (nb <= 64 rows, at A + k + k lda): Cholesky in shared memory, L written back to the lower triangle,
// L^-1 (lower triangular, dense 64 x 64 with zeros above the diagonal) to Linv[64 * 64] (column-major, ld 64).
// A non-positive pivot raises *info (the caller drops the coarse space: Jacobi only).
__global__ void __launch_bounds__(256) dense_diag_kernel(double* A, int64_t lda, int nb, double* Linv, int* info) {
    __shared__ double L[kDenseNB][kDenseNB + 1];
    __shared__ int bad;
    const int tid = threadIdx.x;
    if (tid == 0) bad = 0;
    for (int e = tid; e < kDenseNB * kDenseNB; e += 256) {
        const int r = e & 63, c = e >> 6;
        L[r][c] = (r < nb && c < nb && r >= c) ? A[(int64_t)r + (int64_t)c * lda] : 0.0;
    }
    __syncthreads();
    for (int j = 0; j < nb; ++j) {
        if (tid == 0) {
            const double d = L[j][j];
            if (!(d > 0.0)) { bad = 1; L[j][j] = 1.0; } else L[j][j] = sqrt(d);
        }
        __syncthreads();
        const double djj = L[j][j];
        for (int r = j + 1 + tid; r < nb; r += 256) L[r][j] /= djj;
        __syncthreads();
        // trailing update of the lower triangle: L[r][c] -= L[r][j] L[c][j], j < c <= r
        const int w = nb - j - 1;
        for (int e = tid; e < w * w; e += 256) {
            const int r = j + 1 + e % w, c = j + 1 + e / w;
            if (r >= c) L[r][c] -= L[r][j] * L[c][j];
        }
        __syncthreads();
    }
    for (int e = tid; e < kDenseNB * kDenseNB; e += 256) {
        const int r = e & 63, c = e >> 6;
        if (r < nb && c < nb && r >= c) A[(int64_t)r + (int64_t)c * lda] = L[r][c];
    }
    // column c of L^-1 by forward substitution (one thread per column, fixed order), straight to global memory
    if (tid < kDenseNB) {
        const int c = tid;
        double x[kDenseNB];
#pragma unroll 1
        for (int r = 0; r < kDenseNB; ++r) {
            double s = 0.0;
            if (c < nb && r < nb && r >= c) {
                s = (r == c) ? 1.0 : 0.0;
                for (int q = c; q < r; ++q) s -= L[r][q] * x[q];
                s /= L[r][r];
            }
            x[r] = s;
            Linv[r + c * kDenseNB] = s;
        }
    }
    if (tid == 0 && bad) *info = 1;
}

__global__ void dense_copy_kernel(int m, int n, const double* src, int64_t lds, double* dst, int64_t ldd) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)m * n) return;
    const int r = (int)(i % m), c = (int)(i / m);
    dst[(int64_t)r + (int64_t)c * ldd] = src[(int64_t)r + (int64_t)c * lds];
}

// A (n x n, column-major, ld = n, symmetric positive definite; only the lower triangle is read) -> A^-1 in the lower
// triangle of A.  Y: n x n work matrix; T: n x 64 panel (also used as 64 x n); Linv: ceil(n / 64) blocks of 64 x 64; info: device flag
// (set to 1 on a non-positive pivot, never cleared here).  Returns the number of kernels launched.
inline int dense_spd_inverse(cudaStream_t st, int n, double* A, double* Y, double* T, double* Linv, int* info) {
    const int NB = kDenseNB, nblk = (n + NB - 1) / NB;
    const int64_t ld = n;
    int launches = 0;
    auto tiles = [](int x) { return (x + kDenseNB - 1) / kDenseNB; };
    // ---- Cholesky, right looking
    for (int b = 0; b < nblk; ++b) {
        const int k = b * NB, nb = (n - k < NB) ? n - k : NB, r = n - k - nb;
        double* Akk = A + k + (int64_t)k * ld;
        double* Li = Linv + (size_t)b * NB * NB;
        dense_diag_kernel<<<1, 256, 0, st>>>(Akk, ld, nb, Li, info);
        ++launches;
        if (r <= 0) break;
        double* P = A + (k + nb) + (int64_t)k * ld;                       // r x nb panel below the diagonal block
        // T = P L_kk^-T = P Linv^T
        dense_gemm_kernel<false, true><<<dim3(tiles(r), 1), 256, 0, st>>>(r, nb, nb, 1.0, P, ld, Li, NB, 0.0, T, ld, 0);
        dense_copy_kernel<<<(int)(((int64_t)r * nb + 255) / 256), 256, 0, st>>>(r, nb, T, ld, P, ld);
        // trailing lower triangle: A22 -= P P^T
        double* A22 = A + (k + nb) + (int64_t)(k + nb) * ld;
        dense_gemm_kernel<false, true><<<dim3(tiles(r), tiles(r)), 256, 0, st>>>(r, r, nb, -1.0, P, ld, P, ld, 1.0, A22, ld, 1);
        launches += 3;
    }
    // ---- Y = L^-1 (lower triangular): block row i:  Y_i,0:i = -Linv_ii (L_i,0:i Y_0:i,0:i),  Y_ii = Linv_ii
    cudaMemsetAsync(Y, 0, (size_t)n * n * sizeof(double), st);
    for (int b = 0; b < nblk; ++b) {
        const int k = b * NB, nb = (n - k < NB) ? n - k : NB;
        double* Li = Linv + (size_t)b * NB * NB;
        if (k > 0) {
            // T(nb x k) = L[k:k+nb, 0:k] * Y[0:k, 0:k]
            dense_gemm_kernel<false, false><<<dim3(1, tiles(k)), 256, 0, st>>>(nb, k, k, 1.0, A + k, ld, Y, ld, 0.0, T, NB, 0);
            // Y[k:k+nb, 0:k] = -Linv_ii * T
            dense_gemm_kernel<false, false><<<dim3(1, tiles(k)), 256, 0, st>>>(nb, k, nb, -1.0, Li, NB, T, NB, 0.0, Y + k, ld, 0);
            launches += 2;
        }
        dense_copy_kernel<<<(nb * nb + 255) / 256, 256, 0, st>>>(nb, nb, Li, NB, Y + k + (int64_t)k * ld, ld);
        ++launches;
    }
    // ---- A^-1 = Y^T Y, lower triangle, k-range cut to the non-zero rows of Y
    dense_gemm_kernel<true, false><<<dim3(tiles(n), tiles(n)), 256, 0, st>>>(n, n, n, 1.0, Y, ld, Y, ld, 0.0, A, ld, 2);
    return launches + 1;
}

}  // namespace pfx
