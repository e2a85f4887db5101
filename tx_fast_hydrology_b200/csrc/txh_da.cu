// txh_da.cu -- data-assimilation kernels for sm_100a.
//
// Device form of KalmanFilter.filter (tx_fast_hydrology/da.py:91-136).  The dense contractions
// of the update -- innovation covariance, ensemble transform, gain application, and the
// covariance products of the dense filter -- run on the FP64 tensor cores (mma.sync m8n8k4,
// DMMA in SASS); everything else (means, gathers, the small SPD solve) is plain FP64.
// Matrices are row-major.  State matrices are the routing layout: one row per reach (schedule
// order), members contiguous, `ld` doubles per row.
#include <cuda_runtime.h>

#include <cstdint>

#include "txh_kernels.cuh"

namespace txh {

void count_launch();

namespace {

// D(8x8) += A(8x4) * B(4x8), FP64.  Fragment ownership (lane = 4*g + t):
//   a = A[g][t], b = B[t][g], c0 = C[g][2t], c1 = C[g][2t+1]
__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---- generic C = alpha * op(A) op(B) + beta * C ------------------------------------------------
// CTA tile 64x64, K tile 16; 8 warps, each a 16x32 sub-tile = 2x4 DMMA tiles.
constexpr int GB_M = 64, GB_N = 64, GB_K = 16;

// Split-K: blockIdx.z owns the k range [z*kchunk, (z+1)*kchunk) and writes its partial product to
// C + z*cstride (kchunk == K, gridDim.z == 1: the plain product).
__global__ void __launch_bounds__(256)
dgemm_kernel(int transA, int transB, int M, int N, int K, double alpha, const double* __restrict__ A, int lda,
             const double* __restrict__ B, int ldb, double beta, double* __restrict__ C, int ldc, int kchunk,
             long long cstride)
{
    __shared__ double sA[GB_M][GB_K + 1];      // sA[m][k]
    __shared__ double sB[GB_K][GB_N + 1];      // sB[k][n]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int m0 = blockIdx.y * GB_M, n0 = blockIdx.x * GB_N;
    const int wm = (warp >> 1) * 16, wn = (warp & 1) * 32;
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int kbeg = blockIdx.z * kchunk;
    K = min(K, kbeg + kchunk);
    C += (size_t)blockIdx.z * cstride;
    for (int k0 = kbeg; k0 < K; k0 += GB_K) {
        for (int e = tid; e < GB_M * GB_K; e += 256) {
            int m, k;
            if (transA) { m = e % GB_M; k = e / GB_M; } else { k = e % GB_K; m = e / GB_K; }
            const int gm = m0 + m, gk = k0 + k;
            double v = 0.0;
            if (gm < M && gk < K) v = transA ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk];
            sA[m][k] = v;
        }
        for (int e = tid; e < GB_K * GB_N; e += 256) {
            int k, n;
            if (transB) { k = e % GB_K; n = e / GB_K; } else { n = e % GB_N; k = e / GB_N; }
            const int gk = k0 + k, gn = n0 + n;
            double v = 0.0;
            if (gk < K && gn < N) v = transB ? B[(size_t)gn * ldb + gk] : B[(size_t)gk * ldb + gn];
            sB[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GB_K; kk += 4) {
            double a[2], b[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) a[i] = sA[wm + 8 * i + g][kk + t];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sB[kk + t][wn + 8 * j + g];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int gm = m0 + wm + 8 * i + g, gn = n0 + wn + 8 * j + 2 * t + c;
                if (gm < M && gn < N) {
                    double* p = C + (size_t)gm * ldc + gn;
                    *p = alpha * acc[i][j][c] + (beta == 0.0 ? 0.0 : beta * *p);
                }
            }
}

// ---- ensemble statistics ---------------------------------------------------------------------
// rowsum[k] = sum over this shard's members of X[k][:]   (one warp per row)
__global__ void __launch_bounds__(256)
rowsum_kernel(const double* __restrict__ X, int ld, int M, long long n, double* __restrict__ rowsum)
{
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const double* p = X + (size_t)row * ld;
    double s = 0.0;
    for (int m = lane; m < M; m += 32) s += p[m];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) rowsum[row] = s;
}

// HA[k][m] = HX[k][m] - mean_k ; dz[k][m] = Zp[k][m] - HX[k][m]     (da.py:112)
__global__ void __launch_bounds__(256)
innovation_kernel(const double* __restrict__ HX, const double* __restrict__ Zp, const double* __restrict__ mean_obs,
                  int m, int M, double* __restrict__ HA, double* __restrict__ dz)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= m * M) return;
    const int k = gid / M;
    const double hx = HX[gid];
    HA[gid] = hx - mean_obs[k];
    dz[gid] = Zp[gid] - hx;
}

// S = S * scale + diag(qs) + R        (P[s][:, s] + R_cov with P = sample covariance + Q, da.py:117-119)
__global__ void __launch_bounds__(256)
innov_cov_finish_kernel(double* __restrict__ S, const double* __restrict__ qs, const double* __restrict__ R, int m,
                        double scale)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= m * m) return;
    const int i = gid / m, j = gid - i * m;
    S[gid] = S[gid] * scale + R[gid] + (i == j ? qs[i] : 0.0);
}

// In-place Cholesky factorisation + solve S X = B for SPD S [m][m], B [m][k].  One CTA of 1024
// threads; S stays in global memory (L2 resident).  Left-looking by column, the dot products
// of a column spread over the whole CTA (warp per row, lanes over the inner index).
__global__ void __launch_bounds__(1024)
spd_solve_kernel(double* __restrict__ S, double* __restrict__ B, int m, int k, int* __restrict__ info)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    __shared__ double diag;
    // factor: S = L L^T, L stored in the lower triangle
    for (int j = 0; j < m; ++j) {
        // rows i >= j: S[i][j] -= sum_{p<j} L[i][p] L[j][p]
        for (int i = j + warp; i < m; i += nwarp) {
            const double* Li = S + (size_t)i * m;
            const double* Lj = S + (size_t)j * m;
            double s = 0.0;
            for (int p = lane; p < j; p += 32) s += Li[p] * Lj[p];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) S[(size_t)i * m + j] -= s;
        }
        __syncthreads();
        if (tid == 0) {
            const double d = S[(size_t)j * m + j];
            if (!(d > 0.0)) { *info = j + 1; diag = 1.0; } else diag = sqrt(d);
            S[(size_t)j * m + j] = diag;
        }
        __syncthreads();
        const double inv = 1.0 / diag;
        for (int i = j + 1 + tid; i < m; i += blockDim.x) S[(size_t)i * m + j] *= inv;
        __syncthreads();
    }
    // forward substitution L Y = B, then backward L^T X = Y.  16 threads share a right-hand-side
    // column (inner index strided over them, combined with shuffles); 64 columns per pass.
    const int q = tid & 15, cl = tid >> 4;
    for (int c0 = 0; c0 < k; c0 += 64) {
        const int c = c0 + cl;
        const bool on = c < k;
        for (int i = 0; i < m; ++i) {
            const double* Li = S + (size_t)i * m;
            double s = 0.0;
            if (on) for (int p = q; p < i; p += 16) s += Li[p] * B[(size_t)p * k + c];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, 16);
            if (on && q == 0) B[(size_t)i * k + c] = (B[(size_t)i * k + c] - s) / Li[i];
            __syncwarp();
        }
        for (int i = m - 1; i >= 0; --i) {
            double s = 0.0;
            if (on) for (int p = i + 1 + q; p < m; p += 16) s += S[(size_t)p * m + i] * B[(size_t)p * k + c];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, 16);
            if (on && q == 0) B[(size_t)i * k + c] = (B[(size_t)i * k + c] - s) / S[(size_t)i * m + i];
            __syncwarp();
        }
    }
}

// Gauss-Jordan inverse with partial pivoting (np.linalg.inv of da.py:119), one CTA, in place:
// A <- inv(A); `work` [m][m] receives the running inverse.
__global__ void __launch_bounds__(1024)
inverse_kernel(double* __restrict__ A, double* __restrict__ W, int m, int* __restrict__ info)
{
    const int tid = threadIdx.x;
    __shared__ int piv;
    __shared__ double pval;
    for (int e = tid; e < m * m; e += blockDim.x) W[e] = (e / m == e % m) ? 1.0 : 0.0;
    __syncthreads();
    for (int j = 0; j < m; ++j) {
        if (tid == 0) {
            int best = j; double bv = fabs(A[(size_t)j * m + j]);
            for (int i = j + 1; i < m; ++i) { const double v = fabs(A[(size_t)i * m + j]); if (v > bv) { bv = v; best = i; } }
            piv = best; pval = A[(size_t)best * m + j];
            if (bv == 0.0) { *info = j + 1; pval = 1.0; }
        }
        __syncthreads();
        const int pr = piv;
        if (pr != j)
            for (int c = tid; c < m; c += blockDim.x) {
                double x = A[(size_t)j * m + c]; A[(size_t)j * m + c] = A[(size_t)pr * m + c]; A[(size_t)pr * m + c] = x;
                x = W[(size_t)j * m + c]; W[(size_t)j * m + c] = W[(size_t)pr * m + c]; W[(size_t)pr * m + c] = x;
            }
        __syncthreads();
        const double inv = 1.0 / pval;
        for (int c = tid; c < m; c += blockDim.x) { A[(size_t)j * m + c] *= inv; W[(size_t)j * m + c] *= inv; }
        __syncthreads();
        // eliminate column j from every other row; each thread owns (row, column-chunk) pairs
        for (int e = tid; e < m * m; e += blockDim.x) {
            const int i = e / m, c = e - i * m;
            if (i == j) continue;
            const double f = A[(size_t)i * m + j];
            if (f != 0.0 && c != j) A[e] -= f * A[(size_t)j * m + c];
            W[e] -= f * W[(size_t)j * m + c];
        }
        __syncthreads();
        for (int i = tid; i < m; i += blockDim.x) if (i != j) A[(size_t)i * m + j] = 0.0;
        __syncthreads();
    }
    for (int e = tid; e < m * m; e += blockDim.x) A[e] = W[e];
}


// ---- ensemble-space (Woodbury) form of the innovation solve ----------------------------------------
// With D = R + diag(Q[s,s]) constant between updates and its inverse precomputed, the m x m system
// S W = dz,  S = HA HA^T/(Mt-1) + D,  collapses to an Mt x Mt one:
//     Y  = D^-1 [HA | dz]                      C = HA^T Y = [C0 | C1]
//     (C0 + (Mt-1) I) Z = C1                   T = HA^T W/(Mt-1) = Z          W = Y_dz - Y_HA Z
// (HA^T S^-1 = (Mt-1) (C0 + (Mt-1) I)^-1 HA^T D^-1, so the ensemble transform IS the small solve.)

// Bc[k] = [HA_k | dz_k] (2*Mt doubles per gauge); Y = dinv_k * Bc when D is diagonal.
__global__ void __launch_bounds__(256)
innovation_cat_kernel(const double* __restrict__ HX, const double* __restrict__ Zp, const double* __restrict__ mean,
                      const int32_t* __restrict__ obs_pos, const double* __restrict__ dinv_diag, int m, int Mt,
                      double* __restrict__ Bc, double* __restrict__ Y)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= m * Mt) return;
    const int k = gid / Mt, c = gid - k * Mt;
    const double hx = HX[gid];
    const double ha = hx - mean[obs_pos[k]], dz = Zp[gid] - hx;
    const size_t o = (size_t)k * 2 * Mt + c;
    Bc[o] = ha; Bc[o + Mt] = dz;
    if (dinv_diag) { const double d = dinv_diag[k]; Y[o] = d * ha; Y[o + Mt] = d * dz; }
}

// (sum of the split-K partials of C) -> A = C0 + shift*I, B = C1; Cholesky A = L L^T in shared
// memory (right-looking, one CTA), then L Y = B, L^T Z = Y for the Mt right-hand sides; Z -> global.
// Mt <= 128.
__global__ void __launch_bounds__(512)
chol_solve_small_kernel(const double* __restrict__ Cpart, int nsplit, long long pstride, int Mt, double shift,
                        double* __restrict__ Z, int* __restrict__ info)
{
    extern __shared__ double sm[];
    const int ldA = Mt + 1;
    double* A = sm;                       // [Mt][Mt+1]
    double* B = sm + (size_t)Mt * ldA;    // [Mt][Mt+1]
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < Mt * 2 * Mt; e += nt) {
        const int i = e / (2 * Mt), j = e - i * 2 * Mt;
        double v = 0.0;
        for (int s = 0; s < nsplit; ++s) v += Cpart[(size_t)s * pstride + e];
        if (j < Mt) A[i * ldA + j] = v + (i == j ? shift : 0.0);
        else B[i * ldA + (j - Mt)] = v;
    }
    __syncthreads();
    for (int j = 0; j < Mt; ++j) {
        const double d = A[j * ldA + j];                      // every thread reads the pivot
        if (tid == 0 && !(d > 0.0)) *info = j + 1;
        const double r = 1.0 / sqrt(d > 0.0 ? d : 1.0);
        __syncthreads();
        for (int i = j + tid; i < Mt; i += nt) A[i * ldA + j] *= r;          // column j of L (the pivot too)
        __syncthreads();
        // trailing update of the lower triangle: A[i][c] -= L[i][j] L[c][j], j < c <= i
        const int rem = Mt - j - 1;
        for (int e = tid; e < rem * rem; e += nt) {
            const int ii = e / rem, cc = e - ii * rem;
            if (cc <= ii) {
                const int i = j + 1 + ii, c = j + 1 + cc;
                A[i * ldA + c] -= A[i * ldA + j] * A[c * ldA + j];
            }
        }
        __syncthreads();
    }
    // triangular solves: 8 threads per right-hand-side column, 64 columns per pass
    const int q = tid & 7, cl = tid >> 3, cper = nt >> 3;
    for (int c0 = 0; c0 < Mt; c0 += cper) {
        const int c = c0 + cl;
        const bool on = c < Mt;
        for (int i = 0; i < Mt; ++i) {
            double s = 0.0;
            if (on) for (int p = q; p < i; p += 8) s += A[i * ldA + p] * B[p * ldA + c];
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, 8);
            if (on && q == 0) B[i * ldA + c] = (B[i * ldA + c] - s) / A[i * ldA + i];
            __syncwarp();
        }
        for (int i = Mt - 1; i >= 0; --i) {
            double s = 0.0;
            if (on) for (int p = i + 1 + q; p < Mt; p += 8) s += A[p * ldA + i] * B[p * ldA + c];
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, 8);
            if (on && q == 0) B[i * ldA + c] = (B[i * ldA + c] - s) / A[i * ldA + i];
            __syncwarp();
        }
    }
    __syncthreads();
    for (int e = tid; e < Mt * Mt; e += nt) Z[e] = B[(e / Mt) * ldA + (e % Mt)];
}

// W[k][c] = Y_dz[k][c] - sum_j Y_HA[k][j] Z[j][c]      (Y rows are [Y_HA | Y_dz], 2*Mt doubles)
__global__ void __launch_bounds__(256)
woodbury_w_kernel(const double* __restrict__ Y, const double* __restrict__ Z, int m, int Mt, double* __restrict__ W)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= m * Mt) return;
    const int k = gid / Mt, c = gid - k * Mt;
    const double* y = Y + (size_t)k * 2 * Mt;
    double s = y[Mt + c];
    for (int j = 0; j < Mt; ++j) s -= y[j] * Z[(size_t)j * Mt + c];
    W[gid] = s;
}

// G[k][c] = sum_m (X[k][m] - mean_k) T[m][c]  for the shard's columns c: the ensemble transform
// applied to every reach.  CTA = 4 warps, 32 rows per CTA; T staged in shared memory once; each
// warp owns 8 rows and walks the 8x8 output tiles with FP64 tensor-core MMAs.
// Xall: [n][ldx] rows with Mtot members (the gathered ensemble, or the state itself when Mtot == Mloc)
constexpr int EG_ROWS = 32;

__global__ void __launch_bounds__(128)
enkf_gain_kernel(const double* __restrict__ Xall, int ldx, int Mtot, const double* __restrict__ mean,
                 const double* __restrict__ T, int ldt, int Mloc, double* __restrict__ G, int ldg, long long n)
{
    extern __shared__ double smem[];
    const int Kp = (Mtot + 3) & ~3;            // K padded to the MMA depth
    const int Np = (Mloc + 7) & ~7;            // N padded to the MMA width
    double* sT = smem;                          // [Kp][Np + 1]
    double* sX = smem + (size_t)Kp * (Np + 1);  // [EG_ROWS][Kp + 1]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    for (int e = tid; e < Kp * Np; e += 128) {
        const int k = e / Np, c = e - k * Np;
        sT[k * (Np + 1) + c] = (k < Mtot && c < Mloc) ? T[(size_t)k * ldt + c] : 0.0;
    }
    for (long long r0 = (long long)blockIdx.x * EG_ROWS; r0 < n; r0 += (long long)gridDim.x * EG_ROWS) {
        __syncthreads();
        for (int e = tid; e < EG_ROWS * Kp; e += 128) {
            const int r = e / Kp, k = e - r * Kp;
            const long long row = r0 + r;
            sX[r * (Kp + 1) + k] = (row < n && k < Mtot) ? Xall[(size_t)row * ldx + k] - mean[row] : 0.0;
        }
        __syncthreads();
        const int rw = warp * 8;
        for (int c0 = 0; c0 < Np; c0 += 8) {
            double c0v = 0.0, c1v = 0.0;
            for (int k0 = 0; k0 < Kp; k0 += 4)
                dmma8x8x4(c0v, c1v, sX[(rw + g) * (Kp + 1) + k0 + t], sT[(k0 + t) * (Np + 1) + c0 + g]);
            const long long row = r0 + rw + g;
            const int col = c0 + 2 * t;
            if (row < n) {
                if (col < Mloc) G[(size_t)row * ldg + col] = c0v;
                if (col + 1 < Mloc) G[(size_t)row * ldg + col + 1] = c1v;
            }
        }
    }
}

// gauge rows: G[pos_k][c] += qs[k] * W[k][col0 + c]      (the Q[:, s] term of the gain, da.py:117-121)
__global__ void __launch_bounds__(256)
enkf_gauge_term_kernel(const int32_t* __restrict__ obs_pos, const double* __restrict__ qs, const double* __restrict__ W,
                       int m, int Mtot, int col0, int Mloc, double* __restrict__ G, int ldg)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= m * Mloc) return;
    const int k = gid / Mloc, c = gid - k * Mloc;
    G[(size_t)obs_pos[k] * ldg + c] += qs[k] * W[(size_t)k * Mtot + col0 + c];
}

__global__ void __launch_bounds__(256)
scale_kernel(double* __restrict__ X, long long count, double s)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < count) X[gid] *= s;
}

// dense filter helpers: P += Q ;  Ps = P[:, s] (columns) ; Pss = P[s][:, s]
__global__ void __launch_bounds__(256)
gather_cols_kernel(const double* __restrict__ P, int n, const int32_t* __restrict__ idx, int m, double* __restrict__ out)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n * m) return;
    const int i = (int)(gid / m), k = (int)(gid - (long long)i * m);
    out[gid] = P[(size_t)i * n + idx[k]];
}

__global__ void __launch_bounds__(256)
gather_rows_dense_kernel(const double* __restrict__ P, int ncols, const int32_t* __restrict__ idx, int m,
                         double* __restrict__ out)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)m * ncols) return;
    const int k = (int)(gid / ncols), c = (int)(gid - (long long)k * ncols);
    out[gid] = P[(size_t)idx[k] * ncols + c];
}

inline unsigned nblk(long long work, int threads) { return (unsigned)((work + threads - 1) / threads); }

}  // namespace

cudaError_t launch_dgemm(int transA, int transB, int M, int N, int K, double alpha, const double* A, int lda,
                         const double* B, int ldb, double beta, double* C, int ldc, cudaStream_t st)
{
    if (M <= 0 || N <= 0) return cudaSuccess;
    dim3 grid((N + GB_N - 1) / GB_N, (M + GB_M - 1) / GB_M);
    dgemm_kernel<<<grid, 256, 0, st>>>(transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, K > 0 ? K : 1, 0);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_dgemm_splitk(int transA, int transB, int M, int N, int K, const double* A, int lda, const double* B,
                                int ldb, double* Cpart, int ldc, int nsplit, long long pstride, cudaStream_t st)
{
    if (M <= 0 || N <= 0 || nsplit < 1) return cudaSuccess;
    int kchunk = (K + nsplit - 1) / nsplit;
    kchunk = (kchunk + GB_K - 1) / GB_K * GB_K;
    dim3 grid((N + GB_N - 1) / GB_N, (M + GB_M - 1) / GB_M, nsplit);
    dgemm_kernel<<<grid, 256, 0, st>>>(transA, transB, M, N, K, 1.0, A, lda, B, ldb, 0.0, Cpart, ldc, kchunk, pstride);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_innovation_cat(const double* HX, const double* Zp, const double* mean, const int32_t* obs_pos,
                                  const double* dinv_diag, int m, int Mt, double* Bc, double* Y, cudaStream_t st)
{
    innovation_cat_kernel<<<nblk((long long)m * Mt, 256), 256, 0, st>>>(HX, Zp, mean, obs_pos, dinv_diag, m, Mt, Bc, Y);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_chol_solve_small(const double* Cpart, int nsplit, long long pstride, int Mt, double shift, double* Z,
                                    int* info, cudaStream_t st)
{
    const size_t smem = 2 * (size_t)Mt * (Mt + 1) * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(chol_solve_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    chol_solve_small_kernel<<<1, 512, smem, st>>>(Cpart, nsplit, pstride, Mt, shift, Z, info);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_woodbury_w(const double* Y, const double* Z, int m, int Mt, double* W, cudaStream_t st)
{
    woodbury_w_kernel<<<nblk((long long)m * Mt, 256), 256, 0, st>>>(Y, Z, m, Mt, W);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_rowsum(const double* X, int ld, int M, int64_t n, double* rowsum, cudaStream_t st)
{
    rowsum_kernel<<<nblk(n, 8), 256, 0, st>>>(X, ld, M, n, rowsum);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_innovation(const double* HX, const double* Zp, const double* mean_obs, int m, int M, double* HA,
                              double* dz, cudaStream_t st)
{
    innovation_kernel<<<nblk((long long)m * M, 256), 256, 0, st>>>(HX, Zp, mean_obs, m, M, HA, dz);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_innov_cov_finish(double* S, const double* qs, const double* R, int m, double scale, cudaStream_t st)
{
    innov_cov_finish_kernel<<<nblk((long long)m * m, 256), 256, 0, st>>>(S, qs, R, m, scale);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_spd_solve(double* S, double* B, int m, int k, int* info, cudaStream_t st)
{
    spd_solve_kernel<<<1, 1024, 0, st>>>(S, B, m, k, info);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_inverse(double* A, double* W, int m, int* info, cudaStream_t st)
{
    inverse_kernel<<<1, 1024, 0, st>>>(A, W, m, info);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_enkf_gain(const double* Xall, int ldx, int Mtot, const double* mean, const double* T, int ldt, int Mloc,
                             double* G, int ldg, int64_t n, int num_sms, cudaStream_t st)
{
    const int Kp = (Mtot + 3) & ~3, Np = (Mloc + 7) & ~7;
    const size_t smem = ((size_t)Kp * (Np + 1) + (size_t)EG_ROWS * (Kp + 1)) * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(enkf_gain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long tiles = (n + EG_ROWS - 1) / EG_ROWS;
    long long grid = tiles < (long long)num_sms * 4 ? tiles : (long long)num_sms * 4;
    enkf_gain_kernel<<<(unsigned)grid, 128, smem, st>>>(Xall, ldx, Mtot, mean, T, ldt, Mloc, G, ldg, n);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_enkf_gauge_term(const int32_t* obs_pos, const double* qs, const double* W, int m, int Mtot, int col0,
                                   int Mloc, double* G, int ldg, cudaStream_t st)
{
    enkf_gauge_term_kernel<<<nblk((long long)m * Mloc, 256), 256, 0, st>>>(obs_pos, qs, W, m, Mtot, col0, Mloc, G, ldg);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_scale(double* X, int64_t count, double s, cudaStream_t st)
{
    scale_kernel<<<nblk(count, 256), 256, 0, st>>>(X, count, s);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_gather_cols(const double* P, int n, const int32_t* idx, int m, double* out, cudaStream_t st)
{
    gather_cols_kernel<<<nblk((long long)n * m, 256), 256, 0, st>>>(P, n, idx, m, out);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_gather_rows_dense(const double* P, int ncols, const int32_t* idx, int m, double* out, cudaStream_t st)
{
    gather_rows_dense_kernel<<<nblk((long long)m * ncols, 256), 256, 0, st>>>(P, ncols, idx, m, out);
    count_launch();
    return cudaGetLastError();
}

}  // namespace txh
