// hpfw_b200/csrc/learn.cu — index-time filter learning.
//
// Replaces HashprintHandle::calc_cov and calc_filters (/root/reference/include/hpfw/core/hashprint_handle.h:96-112) and the
// mutex-protected accumulate of ParallelCollector::preprocess (/root/reference/include/hpfw/core/parallel_collector.h:92-97,111).
//
// calc_cov(frames^T) is a 2420 x 2420 x frames SYRK in the reference (170 GFLOP per 3-minute track). The context frames are
// a sliding window over the spectrogram, X[t,(b,c)] = S[b,t+c], so with P_d[u] = S[b,u] S[b',u+d]
//     G[(b,c),(b',c+d)] = sum_{u=c}^{c+nf-1} P_d[u] = T0_d[b,b'] - sum_{u<c} P_d[u] + sum_{u=nf}^{nf+c-1} P_d[u],
//     T0_d[b,b'] = sum_{u<nf} P_d[u]:
// twenty 121 x 121 x nf GEMMs (4.2 GFLOP) plus <= 19-term edge corrections give the same matrix as the SYRK, 40x cheaper.
// The per-band mean is removed first (it cancels in the covariance) so that G - nf mu mu' does not cancel catastrophically
// in fp32. cov = (G - nf mu_i mu_j) / (nf - 1) is added to the device-resident accumulator of the context.
//
// calc_filters: only the top 64 eigenvectors are needed, so instead of a dense eigen-solve the GPU runs block subspace
// iteration V <- orth(A V) on a 2420 x 128 block, device-resident: the product and the CholQR2 orthonormalisation (Gram
// matrix, Cholesky and triangular inverse in one CTA, block rotation) accumulate in double and store float; the host only
// does the 128 x 128 Rayleigh-Ritz eigenproblem (cyclic Jacobi, double) at checkpoints 4, 8, 16, ... Eigenvector signs are
// arbitrary in any solver (MKL ssyev vs Eigen's QL differ too); they are normalised so that the largest-magnitude component is
// positive. Hamming distances do not depend on the sign as long as DB and queries use the same filters.
#include "common.cuh"
#include "eigh_host.h"

#include <chrono>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace hpfw_b200 {

int cov_tc_run(hpfw_ctx *ctx, const float *d_hi, const float *d_lo, int cols, int nf, int splits, float *d_part,
               int *rows_done, cudaStream_t stream);


constexpr int LN_BINS = HPFW_BINS;          // 121
constexpr int LN_CTX = HPFW_CONTEXT;        // 20
constexpr int LN_FS = HPFW_FRAME_SIZE;      // 2420
constexpr int LN_KS = 29;                   // time splits of the T0 GEMMs: 29 x 5 lag groups = 145 CTAs of cov_tc_kernel
constexpr int LN_P = 128;                   // subspace block size (64 wanted + 64 guard vectors)

// ---- per-band sums in one coalesced pass: partial[cta][0][b] = sum_{t < nf} S[t][b], partial[cta][1][b] = sum_{t >= nf} S[t][b] --
constexpr int LN_SUM_CTAS = 592;
__global__ void __launch_bounds__(128) band_sums_kernel(const float *__restrict__ S, int cols, int nf, double *__restrict__ partial) {
    const int b = threadIdx.x;
    double a0 = 0.0, a1 = 0.0;
    if (b < LN_BINS) {
        int t = blockIdx.x;
        for (; t + 3 * (int)gridDim.x < cols; t += 4 * gridDim.x) {      // four independent loads in flight
            const float v0 = S[(size_t)t * LN_BINS + b], v1 = S[(size_t)(t + gridDim.x) * LN_BINS + b];
            const float v2 = S[(size_t)(t + 2 * gridDim.x) * LN_BINS + b], v3 = S[(size_t)(t + 3 * gridDim.x) * LN_BINS + b];
            if (t < nf) a0 += (double)v0; else a1 += (double)v0;
            if (t + (int)gridDim.x < nf) a0 += (double)v1; else a1 += (double)v1;
            if (t + 2 * (int)gridDim.x < nf) a0 += (double)v2; else a1 += (double)v2;
            if (t + 3 * (int)gridDim.x < nf) a0 += (double)v3; else a1 += (double)v3;
        }
        for (; t < cols; t += gridDim.x) {
            const double v = (double)S[(size_t)t * LN_BINS + b];
            if (t < nf) a0 += v; else a1 += v;
        }
    }
    partial[((size_t)blockIdx.x * 2 + 0) * 128 + b] = a0;
    partial[((size_t)blockIdx.x * 2 + 1) * 128 + b] = a1;
}

// mean[b] over all columns and sum0[b] = sum_{t < nf} (S[t][b] - mean_b), from the partial sums in a fixed order
__global__ void __launch_bounds__(128)
band_mean_kernel(const double *__restrict__ partial, int nparts, int cols, int nf, float *__restrict__ mean, float *__restrict__ sum0) {
    const int b = threadIdx.x;
    double a0 = 0.0, a1 = 0.0, c0 = 0.0, c1 = 0.0;
    int p = 0;
    for (; p + 1 < nparts; p += 2) {
        a0 += partial[((size_t)p * 2 + 0) * 128 + b];
        a1 += partial[((size_t)p * 2 + 1) * 128 + b];
        c0 += partial[((size_t)(p + 1) * 2 + 0) * 128 + b];
        c1 += partial[((size_t)(p + 1) * 2 + 1) * 128 + b];
    }
    for (; p < nparts; ++p) {
        a0 += partial[((size_t)p * 2 + 0) * 128 + b];
        a1 += partial[((size_t)p * 2 + 1) * 128 + b];
    }
    a0 += c0;
    a1 += c1;
    const float m = b < LN_BINS ? (float)((a0 + a1) / (double)cols) : 0.f;
    mean[b] = m;
    sum0[b] = b < LN_BINS ? (float)(a0 - (double)nf * (double)m) : 0.f;
}

// ---- centred copy Sc[t][b] = S[t][b] - mean_b (+ its tf32 hi / lo parts, bands padded to 128, for the tensor-core kernel) ------
__global__ void __launch_bounds__(256)
center_kernel(const float *__restrict__ S, int cols, const float *__restrict__ mean, float *__restrict__ Sc,
              float *__restrict__ hi, float *__restrict__ lo) {
    __shared__ float mean_s[128];
    if (threadIdx.x < 128) mean_s[threadIdx.x] = mean[threadIdx.x];
    __syncthreads();
    const size_t total = (size_t)cols * 128;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t t = i >> 7;
        const int b = (int)(i & 127);
        float v = 0.f;
        if (b < LN_BINS) {
            v = S[t * LN_BINS + b] - mean_s[b];
            Sc[t * LN_BINS + b] = v;
        }
        if (hi) {
            uint32_t h, l;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
            const float r = v - __uint_as_float(h);
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(r));
            hi[i] = __uint_as_float(h);
            lo[i] = __uint_as_float(l);
        }
    }
}

// ---- T0 partials: part[ks][d][b][b'] = sum_{u in chunk ks} Sc[u][b] * Sc[u+d][b'] ------------------------------------------
// grid (2, 2, 20 * LN_KS), 256 threads, 64 x 64 output tile, 4 x 4 per thread, K-chunks of 16 rows.
__global__ void __launch_bounds__(256)
t0_gemm_kernel(const float *__restrict__ Sc, int nf, float *__restrict__ part) {
    __shared__ float As[16][64 + 4], Bs[16][64 + 4];
    const int d = blockIdx.z / LN_KS, ks = blockIdx.z % LN_KS;
    const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    const int per = (nf + LN_KS - 1) / LN_KS;
    const int u_begin = ks * per, u_end = min(nf, u_begin + per);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int u0 = u_begin; u0 < u_end; u0 += 16) {
        for (int e = threadIdx.x; e < 16 * 64; e += 256) {
            const int kk = e >> 6, col = e & 63;
            const int u = u0 + kk;
            const bool ok = u < u_end;
            As[kk][col] = (ok && m0 + col < LN_BINS) ? Sc[(size_t)u * LN_BINS + m0 + col] : 0.f;
            Bs[kk][col] = (ok && n0 + col < LN_BINS) ? Sc[(size_t)(u + d) * LN_BINS + n0 + col] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    float *dst = part + ((size_t)ks * LN_CTX + d) * (128 * 128);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[(m0 + ty * 4 + i) * 128 + n0 + tx * 4 + j] = acc[i][j];
}

// ---- T0[d][b][b'] = sum over the time splits (+ the rows the tensor-core kernel left out of its whole 8-row K steps) -----------
// Lag 0 is symmetric: (b, b') and (b', b) both take the upper-triangle entry, so that the covariance stays exactly symmetric
// whatever the order in which the GEMM kernel accumulated its split products.
__global__ void __launch_bounds__(256)
t0_reduce_kernel(const float *__restrict__ part, const float *__restrict__ Sc, int nf, int tail_begin, float *__restrict__ T0) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= LN_CTX * LN_BINS * LN_BINS) return;
    const int bp = idx % LN_BINS, b = (idx / LN_BINS) % LN_BINS, d = idx / (LN_BINS * LN_BINS);
    const int rb = (d == 0 && b > bp) ? bp : b, rbp = (d == 0 && b > bp) ? b : bp;
    float t0 = 0.f;
    for (int ks = 0; ks < LN_KS; ++ks) t0 += part[((size_t)ks * LN_CTX + d) * (128 * 128) + rb * 128 + rbp];
    for (int u = tail_begin; u < nf; ++u) t0 = fmaf(Sc[(size_t)u * LN_BINS + rb], Sc[(size_t)(u + d) * LN_BINS + rbp], t0);
    T0[idx] = t0;
}

// ---- rs[b][c] = sum_{u=c}^{c+nf-1} Sc[u][b]: the frame-row sums behind the means, by the running update from sum0 ----------------
__global__ void __launch_bounds__(128) rowsum_kernel(const float *__restrict__ Sc, int nf, const float *__restrict__ sum0,
                                                     float *__restrict__ rs) {
    const int b = threadIdx.x;
    if (b >= LN_BINS) return;
    float s = sum0[b];
    for (int c = 0; c < LN_CTX; ++c) {
        rs[b * LN_CTX + c] = s;
        if (c + 1 < LN_CTX) s += Sc[(size_t)(nf + c) * LN_BINS + b] - Sc[(size_t)c * LN_BINS + b];      // nf + 18 = cols - 1
    }
}

// ---- edge corrections, means, normalisation, accumulate: one warp per ordered band pair (b, b') ------------------------------
// Lane d < 20 walks diagonal d of the 20 x 20 block: entries (i = b*20 + c, j = b'*20 + c + d), c = 0 .. 19 - d, with
//   g_{c+1} = g_c + Sc[nf+c][b] Sc[nf+c+d][b'] - Sc[c][b] Sc[c+d][b']      (window slides by one frame),   g_0 = T0_d[b, b'].
// The block is staged in shared memory and added to the accumulator in runs that are contiguous in memory: directly
// (column j, rows i) and mirrored for d > 0 (column i, rows j) - the same value goes to both, so the matrix stays symmetric.
__global__ void __launch_bounds__(128)
cov_block_kernel(const float *__restrict__ Sc, int nf, const float *__restrict__ T0, const float *__restrict__ rs,
                 float *__restrict__ accum) {
    __shared__ float tile[4][LN_CTX][LN_CTX + 1];
    __shared__ float edge[4][4][LN_CTX];              // [head b, head b', tail b, tail b'][row offset 0 .. 18]
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = blockIdx.x * 4 + w;
    const bool live = pair < LN_BINS * LN_BINS;
    const int b = live ? pair / LN_BINS : 0, bp = live ? pair % LN_BINS : 0;
    if (lane < LN_CTX - 1) {       // the recurrence touches rows 0 .. 18 and nf .. nf + 18 (= cols - 1) of both bands
        edge[w][0][lane] = Sc[(size_t)lane * LN_BINS + b];
        edge[w][1][lane] = Sc[(size_t)lane * LN_BINS + bp];
        edge[w][2][lane] = Sc[(size_t)(nf + lane) * LN_BINS + b];
        edge[w][3][lane] = Sc[(size_t)(nf + lane) * LN_BINS + bp];
    }
    __syncwarp();
    const float inv_nf = 1.0f / (float)nf, inv_nm1 = 1.0f / (float)(nf - 1);
    if (live && lane < LN_CTX) {
        const int d = lane;
        float g = T0[((size_t)d * LN_BINS + b) * LN_BINS + bp];
        for (int c = 0; c + d < LN_CTX; ++c) {
            const float mi = rs[b * LN_CTX + c] * inv_nf, mj = rs[bp * LN_CTX + c + d] * inv_nf;
            tile[w][c][c + d] = (g - (float)nf * (mi * mj)) * inv_nm1;   // (mi * mj) first: bit-identical for (i,j) and (j,i)
            if (c + d + 1 >= LN_CTX) break;
            g += edge[w][2][c] * edge[w][3][c + d] - edge[w][0][c] * edge[w][1][c + d];
        }
    }
    __syncwarp();
    if (!live) return;
    // direct: column j = b'*20 + cp, rows i = b*20 + c for c <= cp
    for (int e = lane; e < LN_CTX * LN_CTX; e += 32) {
        const int cp = e / LN_CTX, c = e % LN_CTX;
        if (c <= cp) accum[(size_t)(b * LN_CTX + c) + (size_t)LN_FS * (bp * LN_CTX + cp)] += tile[w][c][cp];
    }
    // mirror (d > 0): column i = b*20 + c, rows j = b'*20 + cp for cp > c
    for (int e = lane; e < LN_CTX * LN_CTX; e += 32) {
        const int c = e / LN_CTX, cp = e % LN_CTX;
        if (cp > c) accum[(size_t)(bp * LN_CTX + cp) + (size_t)LN_FS * (b * LN_CTX + c)] += tile[w][c][cp];
    }
}

// ---- subspace iteration on the GPU: every product accumulates in double, blocks are stored in float ----------------------------
// Z = A V: A symmetric n x n, V and Z column-major n x p. 32 x 64 output tile per CTA (152 CTAs at n = 2420, p = 128).
__global__ void __launch_bounds__(256)
symm_block_mul_kernel(const float *__restrict__ A, const float *__restrict__ V, float *__restrict__ Z, int n, int p) {
    __shared__ float As[16][32 + 1], Vs[16][64 + 4];
    const int m0 = blockIdx.x * 32, c0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;      // thread: rows ty*2 .. +1, columns tx*4 .. +3
    double acc[2][4] = {};
    for (int k0 = 0; k0 < n; k0 += 16) {
        for (int e = threadIdx.x; e < 16 * 32; e += 256) {
            const int kk = e >> 5, col = e & 31;
            const int k = k0 + kk;
            // A(m, k) = A(k, m): read row k of the column-major matrix as a contiguous run over m
            As[kk][col] = (k < n && m0 + col < n) ? A[(size_t)k * n + m0 + col] : 0.f;
        }
        for (int e = threadIdx.x; e < 16 * 64; e += 256) {
            const int col = e >> 4, kk = e & 15;
            const int k = k0 + kk;
            Vs[kk][col] = (k < n && c0 + col < p) ? V[(size_t)(c0 + col) * n + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            double a[2], b[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) a[i] = (double)As[kk][ty * 2 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = (double)Vs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty * 2 + i, c = c0 + tx * 4 + j;
            if (m < n && c < p) Z[(size_t)c * n + m] = (float)acc[i][j];
        }
}

// C[i*p + j] += sum_k X[i*n + k] * Y[j*n + k] over this CTA's K-chunk (C double, zeroed by the caller). grid (p/32, p/32, KS).
__global__ void __launch_bounds__(256)
gram_tn_kernel(const float *__restrict__ X, const float *__restrict__ Y, double *__restrict__ C, int n, int p) {
    __shared__ float Xs[32][32 + 1], Ys[32][32 + 1];
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    const int per = (n + gridDim.z - 1) / gridDim.z, kb = blockIdx.z * per, ke = min(n, kb + per);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;      // thread: rows ty*2 .. +1, columns tx*2 .. +1
    double acc[2][2] = {};
    for (int k0 = kb; k0 < ke; k0 += 32) {
        for (int e = threadIdx.x; e < 32 * 32; e += 256) {
            const int col = e >> 5, kk = e & 31;
            const int k = k0 + kk;
            Xs[kk][col] = (k < ke && i0 + col < p) ? X[(size_t)(i0 + col) * n + k] : 0.f;
            Ys[kk][col] = (k < ke && j0 + col < p) ? Y[(size_t)(j0 + col) * n + k] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < 32; ++kk) {
            const double a0 = Xs[kk][ty * 2], a1 = Xs[kk][ty * 2 + 1], b0 = Ys[kk][tx * 2], b1 = Ys[kk][tx * 2 + 1];
            acc[0][0] = fma(a0, b0, acc[0][0]);
            acc[0][1] = fma(a0, b1, acc[0][1]);
            acc[1][0] = fma(a1, b0, acc[1][0]);
            acc[1][1] = fma(a1, b1, acc[1][1]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int r = i0 + ty * 2 + i, c = j0 + tx * 2 + j;
            if (r < p && c < p) atomicAdd(&C[(size_t)r * p + c], acc[i][j]);
        }
}

// One CTA: G = Z^T Z (p x p, double, p <= 128) -> Cholesky G = L L^T in shared memory -> W = L^{-T}, so that Q = Z W has
// orthonormal columns. W[k*p + j] = Linv[j][k] (zero for k > j). A pivot that has lost all its digits (dependent column)
// is clamped; the second CholQR pass and the Rayleigh-Ritz step clean that direction up.
__global__ void __launch_bounds__(256) chol_inv_kernel(const double *__restrict__ G, double *__restrict__ W, int p) {
    extern __shared__ double Ls[];          // p x p, row-major
    for (int e = threadIdx.x; e < p * p; e += blockDim.x) Ls[e] = G[e];
    __syncthreads();
    double trace = 0.0;
    for (int i = 0; i < p; ++i) trace += Ls[(size_t)i * p + i];
    const double tiny = trace * 1e-28 + 1e-300;
    for (int j = 0; j < p; ++j) {
        const double d = sqrt(fmax(Ls[(size_t)j * p + j], tiny));
        __syncthreads();
        if (threadIdx.x == 0) Ls[(size_t)j * p + j] = d;
        for (int i = j + 1 + threadIdx.x; i < p; i += blockDim.x) Ls[(size_t)i * p + j] /= d;
        __syncthreads();
        const int m = p - j - 1;            // trailing (i, k), j < k <= i < p
        for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
            const int i = j + 1 + e / m, k = j + 1 + e % m;
            if (k <= i) Ls[(size_t)i * p + k] -= Ls[(size_t)i * p + j] * Ls[(size_t)k * p + j];
        }
        __syncthreads();
    }
    // column c of Linv by forward substitution (one thread per column); x lives in the unused upper triangle: Ls[c][i], i > c
    for (int c = threadIdx.x; c < p; c += blockDim.x) {
        const double xc = 1.0 / Ls[(size_t)c * p + c];
        W[(size_t)c * p + c] = xc;                      // Linv[c][c]
        for (int k = 0; k < c; ++k) W[(size_t)c * p + k] = 0.0;      // W[k'=c][j=k<c] = Linv[k][c] = 0
        for (int i = c + 1; i < p; ++i) {
            double sum = Ls[(size_t)i * p + c] * xc;
            for (int k = c + 1; k < i; ++k) sum = fma(Ls[(size_t)i * p + k], Ls[(size_t)c * p + k], sum);   // x[k] stored at Ls[c][k]
            const double xi = -sum / Ls[(size_t)i * p + i];
            Ls[(size_t)c * p + i] = xi;                 // upper triangle, row c: never read by other threads
            W[(size_t)c * p + i] = xi;                  // W[k=c][j=i] = Linv[i][c]
        }
    }
}

// V = Z W: Z column-major n x p (float), W row-major p x p (double), V column-major n x p (float). One thread per (row, 4 columns).
__global__ void __launch_bounds__(256)
gemm_nn_kernel(const float *__restrict__ Z, const double *__restrict__ W, float *__restrict__ V, int n, int p) {
    const int row = blockIdx.x * 64 + (threadIdx.x & 63);
    const int cg = blockIdx.y * 4 + (threadIdx.x >> 6);       // column group of 8
    if (row >= n) return;
    double acc[8] = {};
    for (int k = 0; k < p; ++k) {
        const double z = (double)Z[(size_t)k * n + row];
        const double *w = W + (size_t)k * p + cg * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fma(z, w[j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (cg * 8 + j < p) V[(size_t)(cg * 8 + j) * n + row] = (float)acc[j];
}

// r[c] = || Zr[:, c] - theta[c] Vr[:, c] ||_2  (one CTA per column)
__global__ void __launch_bounds__(256)
resid_kernel(const float *__restrict__ Zr, const float *__restrict__ Vr, const double *__restrict__ theta, int n,
             double *__restrict__ r) {
    __shared__ double red[256];
    const int c = blockIdx.x;
    double acc = 0.0;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const double d = (double)Zr[(size_t)c * n + k] - theta[c] * (double)Vr[(size_t)c * n + k];
        acc = fma(d, d, acc);
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) r[c] = sqrt(red[0]);
}

// ================================================================================================ host numerics (double)
// cyclic Jacobi eigen-decomposition of the symmetric p x p matrix H (row-major); eigenvectors in the columns of Q
static void jacobi_eigh(std::vector<double> &H, int p, std::vector<double> &w, std::vector<double> &Q) {
    Q.assign((size_t)p * p, 0.0);
    for (int i = 0; i < p; ++i) Q[(size_t)i * p + i] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < p; ++i)
            for (int j = 0; j < p; ++j) (i == j ? diag : off) += H[(size_t)i * p + j] * H[(size_t)i * p + j];
        if (off <= 1e-30 * diag) break;
        for (int a = 0; a < p - 1; ++a)
            for (int b = a + 1; b < p; ++b) {
                const double hab = H[(size_t)a * p + b];
                if (std::fabs(hab) < 1e-300) continue;
                const double haa = H[(size_t)a * p + a], hbb = H[(size_t)b * p + b];
                const double theta = (hbb - haa) / (2.0 * hab);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < p; ++k) {
                    const double hka = H[(size_t)k * p + a], hkb = H[(size_t)k * p + b];
                    H[(size_t)k * p + a] = c * hka - s * hkb;
                    H[(size_t)k * p + b] = s * hka + c * hkb;
                }
                for (int k = 0; k < p; ++k) {
                    const double hak = H[(size_t)a * p + k], hbk = H[(size_t)b * p + k];
                    H[(size_t)a * p + k] = c * hak - s * hbk;
                    H[(size_t)b * p + k] = s * hak + c * hbk;
                }
                for (int k = 0; k < p; ++k) {
                    const double qka = Q[(size_t)k * p + a], qkb = Q[(size_t)k * p + b];
                    Q[(size_t)k * p + a] = c * qka - s * qkb;
                    Q[(size_t)k * p + b] = s * qka + c * qkb;
                }
            }
    }
    w.resize(p);
    for (int i = 0; i < p; ++i) w[i] = H[(size_t)i * p + i];
}

}  // namespace hpfw_b200

using namespace hpfw_b200;

// HPFW_COV_IMPL: 1 = tcgen05 correlation kernel (default), 0 = CUDA-core kernel (measurement baseline)
static int cov_impl() {
    const char *v = getenv("HPFW_COV_IMPL");
    return (v && *v) ? atoi(v) : 1;
}

// accum += other; other = 0 (the side accumulator of the extraction stream's second covariance slot)
__global__ void __launch_bounds__(256) cov_fold_kernel(float *__restrict__ accum, float *__restrict__ other, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        accum[i] += other[i];
        other[i] = 0.f;
    }
}

int hpfw_b200::cov_fold_side(hpfw_ctx *ctx, cudaStream_t s) {
    if (!ctx->cov_accum_side.ptr || !ctx->cov_side_dirty) return HPFW_OK;
    if (!ctx->cov_accum.ptr) {
        HPFW_TRY(ctx->cov_accum.reserve(sizeof(float) * (size_t)LN_FS * LN_FS));
        HPFW_CUDA_TRY(cudaMemsetAsync(ctx->cov_accum.ptr, 0, sizeof(float) * (size_t)LN_FS * LN_FS, s));
    }
    KernelScope ks(ctx, HPFW_K_OTHER, s);
    cov_fold_kernel<<<ctx->sm_count * 4, 256, 0, s>>>(ctx->cov_accum.as<float>(), ctx->cov_accum_side.as<float>(),
                                                      (size_t)LN_FS * LN_FS);
    HPFW_CUDA_TRY(cudaGetLastError());
    ctx->cov_side_dirty = false;
    return HPFW_OK;
}

// slot 0 = the context's accumulator and scratch (every public entry point); slot 1 = a second scratch set and a side
// accumulator, so that the extraction stream can run the (latency-bound, 9-kernel) covariance of two tracks concurrently on two
// streams; cov_fold_side adds the side accumulator into the main one when the stream joins.
int hpfw_b200::cov_add_device(hpfw_ctx *ctx, const float *d_spec, int cols, cudaStream_t s, int slot) {
    const int nf = cols - (LN_CTX - 1);
    if (nf < 2) HPFW_FAIL(HPFW_ERR_SHORT, "covariance needs at least 21 spectrogram columns (got %d)", cols);
    DeviceBuffer &accum = slot ? ctx->cov_accum_side : ctx->cov_accum;
    DeviceBuffer &scratch = slot ? ctx->cov_scratch_side : ctx->cov_scratch;
    if (!accum.ptr) {
        HPFW_TRY(accum.reserve(sizeof(float) * (size_t)LN_FS * LN_FS));
        HPFW_CUDA_TRY(cudaMemsetAsync(accum.ptr, 0, sizeof(float) * (size_t)LN_FS * LN_FS, s));
    }
    if (slot) ctx->cov_side_dirty = true;
    // scratch: [part | hi | lo | Sc | T0 | rs | sum0 | partial sums (double)]; hi / lo start on 1 KB boundaries for the TMA maps
    const size_t n_part = (size_t)LN_KS * LN_CTX * 128 * 128, n_pad = ((size_t)std::max(cols, 128) * 128 + 255) & ~size_t(255);
    const size_t n_sc = ((size_t)cols * LN_BINS + 255) & ~size_t(255), n_t0 = ((size_t)LN_CTX * LN_BINS * LN_BINS + 255) & ~size_t(255);
    const size_t n_rs = 128 * LN_CTX, n_floats = n_part + 2 * n_pad + n_sc + n_t0 + n_rs + 256;
    HPFW_TRY(scratch.reserve(sizeof(float) * n_floats + sizeof(double) * LN_SUM_CTAS * 2 * 128));
    float *part = scratch.as<float>();
    float *hi = part + n_part, *lo = hi + n_pad;
    float *Sc = lo + n_pad;
    float *T0 = Sc + n_sc, *rs = T0 + n_t0, *sum0 = rs + n_rs, *mean = sum0 + 128;
    double *partial = reinterpret_cast<double *>(mean + 128);
    const int impl = cov_impl();
    {
        KernelScope ks(ctx, HPFW_K_OTHER, s);
        band_sums_kernel<<<LN_SUM_CTAS, 128, 0, s>>>(d_spec, cols, nf, partial);
    }
    {
        KernelScope ks(ctx, HPFW_K_OTHER, s);
        band_mean_kernel<<<1, 128, 0, s>>>(partial, LN_SUM_CTAS, cols, nf, mean, sum0);
    }
    if (impl == 1 && cols < 128) HPFW_CUDA_TRY(cudaMemsetAsync(hi, 0, sizeof(float) * 2 * n_pad, s));   // rows a TMA box may touch
    {
        KernelScope ks(ctx, HPFW_K_OTHER, s);
        center_kernel<<<ctx->sm_count * 4, 256, 0, s>>>(d_spec, cols, mean, Sc, impl == 1 ? hi : nullptr,
                                                        impl == 1 ? lo : nullptr);
    }
    int tail_begin = nf;
    if (impl == 1) {      // tcgen05: 3 x tf32 split products, cov_tc.cu
        HPFW_TRY(cov_tc_run(ctx, hi, lo, std::max(cols, 128), nf, LN_KS, part, &tail_begin, s));
    } else {
        KernelScope ks(ctx, HPFW_K_OTHER, s);
        t0_gemm_kernel<<<dim3(2, 2, LN_CTX * LN_KS), 256, 0, s>>>(Sc, nf, part);
    }
    {
        KernelScope ks(ctx, HPFW_K_OTHER, s);
        const int total = LN_CTX * LN_BINS * LN_BINS;
        t0_reduce_kernel<<<(total + 255) / 256, 256, 0, s>>>(part, Sc, nf, tail_begin, T0);
    }
    {
        KernelScope ks(ctx, HPFW_K_OTHER, s);
        rowsum_kernel<<<1, 128, 0, s>>>(Sc, nf, sum0, rs);
    }
    {
        KernelScope ks(ctx, HPFW_K_OTHER, s);
        cov_block_kernel<<<(LN_BINS * LN_BINS + 3) / 4, 128, 0, s>>>(Sc, nf, T0, rs, accum.as<float>());
    }
    HPFW_CUDA_TRY(cudaGetLastError());
    ctx->cov_tracks++;
    return HPFW_OK;
}

extern "C" {

int hpfw_cov_reset(hpfw_ctx *ctx) {
    if (!ctx) HPFW_FAIL(HPFW_ERR_ARG, "ctx is NULL");
    DeviceGuard g(ctx->device);
    ctx->order_on(ctx->stream);
    HPFW_TRY(ctx->cov_accum.reserve(sizeof(float) * (size_t)LN_FS * LN_FS));
    HPFW_CUDA_TRY(cudaMemsetAsync(ctx->cov_accum.ptr, 0, sizeof(float) * (size_t)LN_FS * LN_FS, ctx->stream));
    if (ctx->cov_accum_side.ptr)
        HPFW_CUDA_TRY(cudaMemsetAsync(ctx->cov_accum_side.ptr, 0, sizeof(float) * (size_t)LN_FS * LN_FS, ctx->stream));
    ctx->cov_side_dirty = false;
    ctx->cov_tracks = 0;
    return HPFW_OK;
}

int hpfw_cov_set(hpfw_ctx *ctx, const float *accum_2420x2420) {
    if (!ctx || !accum_2420x2420) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_cov_set: NULL argument");
    DeviceGuard g(ctx->device);
    ctx->order_on(ctx->stream);
    HPFW_TRY(ctx->cov_accum.reserve(sizeof(float) * (size_t)LN_FS * LN_FS));
    HPFW_CUDA_TRY(cudaMemcpyAsync(ctx->cov_accum.ptr, accum_2420x2420, sizeof(float) * (size_t)LN_FS * LN_FS,
                                  cudaMemcpyHostToDevice, ctx->stream));
    HPFW_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return HPFW_OK;
}

int hpfw_cov_get(hpfw_ctx *ctx, float *accum_2420x2420) {
    if (!ctx || !accum_2420x2420) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_cov_get: NULL argument");
    DeviceGuard g(ctx->device);
    ctx->order_on(ctx->stream);
    if (!ctx->cov_accum.ptr) HPFW_TRY(hpfw_cov_reset(ctx));
    HPFW_CUDA_TRY(cudaMemcpyAsync(accum_2420x2420, ctx->cov_accum.ptr, sizeof(float) * (size_t)LN_FS * LN_FS,
                                  cudaMemcpyDeviceToHost, ctx->stream));
    HPFW_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return HPFW_OK;
}

// Device-side access to the accumulator for the multi-GPU index: every rank accumulates its own tracks, the accumulators are
// summed by ONE all-reduce over NVLink (torch.distributed / NCCL on the caller's tensor) and written back, and
// hpfw_calc_filters then sees the covariance of all tracks (hpfw_b200/sharded.py: allreduce_covariance).
int hpfw_cov_get_device(hpfw_ctx *ctx, float *d_accum_out, void *stream) {
    if (!ctx || !d_accum_out) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_cov_get_device: NULL argument");
    DeviceGuard g(ctx->device);
    if (!ctx->cov_accum.ptr) HPFW_TRY(hpfw_cov_reset(ctx));
    HPFW_CUDA_TRY(cudaStreamSynchronize(ctx->stream));     // a reset / set on the context's stream comes first
    HPFW_CUDA_TRY(cudaMemcpyAsync(d_accum_out, ctx->cov_accum.ptr, sizeof(float) * (size_t)LN_FS * LN_FS,
                                  cudaMemcpyDeviceToDevice, ctx->pick(stream)));
    return HPFW_OK;
}

int hpfw_cov_set_device(hpfw_ctx *ctx, const float *d_accum, void *stream) {
    if (!ctx || !d_accum) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_cov_set_device: NULL argument");
    DeviceGuard g(ctx->device);
    HPFW_TRY(ctx->cov_accum.reserve(sizeof(float) * (size_t)LN_FS * LN_FS));
    HPFW_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    HPFW_CUDA_TRY(cudaMemcpyAsync(ctx->cov_accum.ptr, d_accum, sizeof(float) * (size_t)LN_FS * LN_FS,
                                  cudaMemcpyDeviceToDevice, ctx->pick(stream)));
    HPFW_CUDA_TRY(cudaStreamSynchronize(ctx->pick(stream)));   // hpfw_calc_filters reads it on the context's stream
    return HPFW_OK;
}

int hpfw_cov_add_spectrogram_device(hpfw_ctx *ctx, const float *d_spectrogram, int cols, void *stream) {
    if (!ctx || !d_spectrogram) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_cov_add_spectrogram_device: NULL argument");
    DeviceGuard g(ctx->device);
    return cov_add_device(ctx, d_spectrogram, cols, ctx->pick(stream));
}

int hpfw_cov_add_spectrogram(hpfw_ctx *ctx, const float *spectrogram, int cols) {
    if (!ctx || !spectrogram) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_cov_add_spectrogram: NULL argument");
    if (cols < LN_CTX + 1) HPFW_FAIL(HPFW_ERR_SHORT, "covariance needs at least 21 spectrogram columns (got %d)", cols);
    DeviceGuard g(ctx->device);
    ctx->order_on(ctx->stream);
    const size_t sb = sizeof(float) * (size_t)cols * LN_BINS;
    HPFW_TRY(ctx->spectro.reserve(sb));
    HPFW_CUDA_TRY(cudaMemcpyAsync(ctx->spectro.ptr, spectrogram, sb, cudaMemcpyHostToDevice, ctx->stream));
    HPFW_TRY(cov_add_device(ctx, ctx->spectro.as<float>(), cols, ctx->stream));
    HPFW_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return HPFW_OK;
}

int hpfw_calc_filters(hpfw_ctx *ctx, const float *cov, float *filters_out, float *eigenvalues_out) {
    if (!ctx || !filters_out) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_calc_filters: NULL argument");
    DeviceGuard g(ctx->device);
    ctx->order_on(ctx->stream);
    const int n = LN_FS, p = LN_P, want = HPFW_NFILTERS;
    cudaStream_t s = ctx->stream;
    // scratch lives in the context: cudaMalloc / cudaFree per call synchronise the whole device, which stalls behind the
    // background device->host copies of the index pipeline's cache writers (100-200 ms per call measured)
    DeviceBuffer dA;
    DeviceBuffer &dV = ctx->eig_scratch[0], &dZ = ctx->eig_scratch[1], &dT = ctx->eig_scratch[2], &dG = ctx->eig_scratch[3],
                 &dW = ctx->eig_scratch[4], &dR = ctx->eig_scratch[5];
    const float *A = nullptr;
    if (cov) {
        HPFW_TRY(dA.reserve(sizeof(float) * (size_t)n * n));
        HPFW_CUDA_TRY(cudaMemcpy(dA.ptr, cov, sizeof(float) * (size_t)n * n, cudaMemcpyHostToDevice));
        A = dA.as<float>();
    } else {
        if (!ctx->cov_accum.ptr) HPFW_FAIL(HPFW_ERR_STATE, "hpfw_calc_filters: no covariance accumulated");
        A = ctx->cov_accum.as<float>();
    }
    const size_t blk = sizeof(float) * (size_t)n * p, pp = sizeof(double) * (size_t)p * p;
    int status = HPFW_OK;
    auto fail = [&](int st) { status = st; };
    if (dV.reserve(blk) || dZ.reserve(blk) || dT.reserve(blk) || dG.reserve(pp) || dW.reserve(pp) ||
        dR.reserve(sizeof(double) * 2 * p))
        fail(HPFW_ERR_CUDA);
    float *V = dV.as<float>(), *Z = dZ.as<float>(), *T = dT.as<float>();
    double *G = dG.as<double>(), *W = dW.as<double>(), *R = dR.as<double>();
    if (status == HPFW_OK &&
        cudaFuncSetAttribute(chol_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pp) != cudaSuccess)
        fail(HPFW_ERR_CUDA);

    const dim3 g_mul((n + 31) / 32, (p + 63) / 64), g_gram((p + 31) / 32, (p + 31) / 32, 8), g_nn((n + 63) / 64, (p + 31) / 32);
    auto gram = [&](const float *X, const float *Y) {
        cudaMemsetAsync(G, 0, pp, s);
        KernelScope ks(ctx, HPFW_K_OTHER, s);
        gram_tn_kernel<<<g_gram, 256, 0, s>>>(X, Y, G, n, p);
    };
    auto rotate = [&](const float *X, const double *Wm, float *Y) {       // Y = X Wm
        KernelScope ks(ctx, HPFW_K_OTHER, s);
        gemm_nn_kernel<<<g_nn, 256, 0, s>>>(X, Wm, Y, n, p);
    };
    // dst = orthonormal basis of span(src) by CholQR2 (Gram matrices and triangular factors in double); src is overwritten
    // (twice = false: one pass, dst orthonormal to ~cond(src) * 1e-7 — enough between two multiplications by A, where the
    // only job is to keep the block from collapsing onto the dominant vectors; the block a Rayleigh-Ritz checkpoint reads is
    // always made by the two-pass form)
    auto orth = [&](float *src, float *tmp, float *dst, bool twice = true) {
        gram(src, src);
        { KernelScope ks(ctx, HPFW_K_OTHER, s); chol_inv_kernel<<<1, 256, pp, s>>>(G, W, p); }
        if (!twice) {
            rotate(src, W, dst);
            return;
        }
        rotate(src, W, tmp);
        gram(tmp, tmp);
        { KernelScope ks(ctx, HPFW_K_OTHER, s); chol_inv_kernel<<<1, 256, pp, s>>>(G, W, p); }
        rotate(tmp, W, dst);
    };

    std::vector<float> Vf((size_t)n * p);
    std::vector<double> H, w, Q, Wq((size_t)p * p), theta(p), prev(want, 0.0), res(p);
    int n_checks = 0, n_iters = 0;
    const auto t_begin = std::chrono::steady_clock::now();
    // Start block: random, except that the first 64 columns are the filters this context already holds (when it holds any
    // and HPFW_FILTERS_WARM_START is not 0): re-indexing a collection that grew by a few tracks moves the dominant subspace
    // only slightly, and the iteration then converges at its first checkpoint. Any start in general position converges to
    // the same subspace; the stopping rule below is unchanged.
    bool warm = ctx->have_filters && ctx->filters_host.size() == (size_t)want * n;
    if (const char *env = getenv("HPFW_FILTERS_WARM_START")) warm = warm && atoi(env) != 0;
    if (status == HPFW_OK) {
        uint64_t lcg = 0x9E3779B97F4A7C15ull;
        for (auto &v : Vf) {
            lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
            v = (float)((double)((lcg >> 11) & 0xFFFFFF) / 16777216.0 - 0.5);
        }
        if (warm)
            for (int f = 0; f < want; ++f)
                for (int k = 0; k < n; ++k) Vf[(size_t)f * n + k] = ctx->filters_host[(size_t)f + (size_t)want * k];
        if (cudaMemcpyAsync(Z, Vf.data(), blk, cudaMemcpyHostToDevice, s) != cudaSuccess) fail(HPFW_ERR_CUDA);
        orth(Z, T, V);
    }
    // Block subspace iteration V <- orth(A V), entirely on the device; a Rayleigh-Ritz checkpoint (p x p eigenproblem on the
    // host, cyclic Jacobi in double) at iterations 4, 8, 16, ... rotates the block to Ritz vectors and tests the residuals.
    const int max_it = 4096;
    // a Rayleigh-Ritz checkpoint costs as much as ~7 iterations (host eigen-solve + two synchronisations): a warm start is
    // tested at once, a cold start first after 16 iterations and then every 8
    int next_check = warm ? 4 : 16;
    bool done = false;
    for (int it = 1; status == HPFW_OK && !done; ++it) {
        n_iters = it;
        { KernelScope ks(ctx, HPFW_K_OTHER, s); symm_block_mul_kernel<<<g_mul, 256, 0, s>>>(A, V, Z, n, p); }
        if (it < next_check && it < max_it) {
            orth(Z, T, V, it + 1 >= next_check || it + 1 >= max_it);
            continue;
        }
        gram(V, Z);                                            // H = V^T A V
        H.resize((size_t)p * p);
        cudaError_t e = cudaMemcpyAsync(H.data(), G, pp, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { set_error("hpfw_calc_filters: %s", cudaGetErrorString(e)); fail(HPFW_ERR_CUDA); break; }
        for (int i = 0; i < p; ++i)
            for (int j = i + 1; j < p; ++j) H[(size_t)i * p + j] = H[(size_t)j * p + i] = 0.5 * (H[(size_t)i * p + j] + H[(size_t)j * p + i]);
        {
            // Rayleigh-Ritz: tridiagonalisation + implicit QL (eigh_host.h, ~9 ms at p = 128); the cyclic Jacobi is the fallback
            std::vector<double> Hc = H;
            if (!sym_eigh_ql(Hc, p, w, Q)) jacobi_eigh(H, p, w, Q);
        }
        ++n_checks;
        std::vector<int> order(p);
        for (int i = 0; i < p; ++i) order[i] = i;
        std::sort(order.begin(), order.end(), [&](int a, int b) { return w[a] > w[b]; });
        for (int k = 0; k < p; ++k)
            for (int j = 0; j < p; ++j) Wq[(size_t)k * p + j] = Q[(size_t)k * p + order[j]];
        for (int j = 0; j < p; ++j) theta[j] = w[order[j]];
        e = cudaMemcpyAsync(W, Wq.data(), pp, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(R, theta.data(), sizeof(double) * p, cudaMemcpyHostToDevice, s);
        rotate(V, W, T);                                       // T = Ritz vectors
        rotate(Z, W, V);                                       // V = A * Ritz vectors
        { KernelScope ks(ctx, HPFW_K_OTHER, s); resid_kernel<<<p, 256, 0, s>>>(V, T, R, n, R + p); }
        if (e == cudaSuccess) e = cudaMemcpyAsync(res.data(), R + p, sizeof(double) * p, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { set_error("hpfw_calc_filters: %s", cudaGetErrorString(e)); fail(HPFW_ERR_CUDA); break; }
        const double top = std::fabs(theta[0]) + 1e-300;
        double rmax = 0.0, change = 0.0;
        for (int i = 0; i < want; ++i) {
            rmax = std::max(rmax, res[i]);
            change = std::max(change, std::fabs(theta[i] - prev[i]) / top);
            prev[i] = theta[i];
        }
        // converged: residuals at the float noise floor of the block, or the Ritz values have stopped moving
        if (rmax <= 2e-6 * top || (it > 4 && change < 1e-10) || it >= max_it) {
            done = true;                                       // T holds the Ritz vectors
        } else {
            orth(V, Z, T);                                     // next block = orth(A * Ritz vectors) -> T
            std::swap(V, T);
            next_check = it < 64 ? it + 8 : std::min(it * 2, it + 512);
        }
    }
    if (status == HPFW_OK) {
        // final polish of the wanted vectors' orthonormality (they are Ritz vectors of an orthonormal block already)
        cudaError_t e = cudaMemcpyAsync(Vf.data(), T, blk, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { set_error("hpfw_calc_filters: %s", cudaGetErrorString(e)); fail(HPFW_ERR_CUDA); }
    }
    if (status == HPFW_OK) {
        // filters(f, i) = eigenvector_f[i], column-major 64 x 2420; sign: largest-|component| positive.
        for (int f = 0; f < want; ++f) {
            const float *v = &Vf[(size_t)f * n];
            int arg = 0;
            for (int k = 1; k < n; ++k)
                if (std::fabs(v[k]) > std::fabs(v[arg])) arg = k;
            const float sgn = v[arg] < 0 ? -1.f : 1.f;
            for (int k = 0; k < n; ++k) filters_out[(size_t)f + (size_t)want * k] = sgn * v[k];
            if (eigenvalues_out) eigenvalues_out[f] = (float)prev[f];
        }
    }
    cudaStreamSynchronize(s);
    dA.release();
    if (getenv("HPFW_TRACE"))
        fprintf(stderr, "[hpfw trace] calc_filters: %d iterations, %d Rayleigh-Ritz checkpoints, %s start, %.1f ms\n", n_iters,
                n_checks, warm ? "warm" : "cold",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    return status;
}

}  // extern "C"
