// hpfw_b200/csrc/cqt.cu — stage 1: the constant-Q front end.
//
// Replaces spectrum::CQT::spectrogram after decoding (/root/reference/include/hpfw/spectrum/cqt.h:54-84: essentia
// NSGConstantQ over the whole track, abs, keep every 3rd coefficient) and amplitude_to_db
// (/root/reference/include/hpfw/spectrum/convert.h:7-25). The arithmetic of NSGConstantQ lives in essentia (absent from the
// reference tree); the algorithm restated here is the one in oracle/nsgcq.py's header ("parity unpinned").
//
//   X = FFT_N(x);  per band j: buf[k] = X[pos_j + k] * hann_j[k], k in [-Lg_j/2, Lg_j/2);  c_j = IFFT_M(buf);
//   S[j, i] = 20 log10 |c_j[3 i]|  relative to the track maximum, floored at -80 dB.
//
// B200 formulation (no library FFT; every transform below is a kernel in this file):
//   (1) N real samples are packed to H = N/2 complex points and transformed by a two-pass ("four-step") FFT, H = n1 * n2
//       with n1, n2 <= 8192 products of 2,3,5,7: pass A = length-n1 column FFTs + inter-pass twiddle, pass B = length-n2
//       row FFTs whose transposed store keeps only the two bin ranges the 121 bands need (~19 % of the half spectrum).
//       Each small FFT is a Stockham autosort over radices 2..16 (composite radices 6,9,10,12,14,15,16 are evaluated in
//       registers) whose first stage reads global memory and whose last stage writes it: fft_pass_kernel.
//   (2) per band, only every 3rd of the M IFFT outputs is wanted and M is in general not smooth (43,528 = 8 * 5441), so
//       c_j[3 i] is evaluated as a chirp-z transform: |c_j[3i]| = |sum_k a[k] e^{2 pi i 3 i k / M}| =
//       |IFFT_L( FFT_L(a .* chirp) .* FFT_L(chirp_filter) )[i]|, L = 16 * L2 >= Lg_j + F - 1 with L2 = 256 r0 the
//       smallest of r0 in {2..10,12,14,15,16} that fits (a power of two would waste 38 % of the work at 3 minutes).
//       FFT_L is again four-step: a radix-16 register FFT down the columns (fused with the half-spectrum untangle, the
//       Hann window and the chirp), czt_rows3_kernel doing FFT_L2 -> multiply -> IFFT_L2 per row with the data in
//       registers between stages, and a radix-16 register IFFT fused with |.|^2, the 1/(L M) scale and the per-track
//       maximum. Rows shorter than 512 or longer than 4096 points use the generic shared-memory kernel (czt_rows_kernel).
//       Tracks longer than 6.7 minutes need L > 16 * 8192: their plans use 32-point columns (dft32 in registers,
//       up to ~13.5 minutes) or 64-point columns (dft64, up to ~27 minutes); everything else is unchanged.
//   (3) dB conversion + transpose to the reference's column-major [121 x cols] layout.
// Twiddles: one table W_n^m per transform length; a butterfly loads one entry and forms its powers by a depth-4 product
// tree on the FMA pipe (error <= 5 ulp of a unit phasor) instead of R-1 table loads.
#include "common.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>

namespace hpfw_b200 {

constexpr int CQ_BINS = HPFW_BINS;
constexpr double CQ_SR = 44100.0;     // essentia NSGConstantQ default sampleRate: cqt.h:54-61 never forwards SampleRate
constexpr double CQ_FMIN = 130.81;    // cqt.h:59
constexpr int CQ_BPO = 24;            // cqt.h:20,56
constexpr int CQ_MINWIN = 96;         // cqt.h:19,57
constexpr int CQ_DOWN = 3;            // cqt.h:22
constexpr int CQ_L1 = 16;             // CZT column radix
constexpr int CQ_MAX_ROW = 8192;      // longest shared-memory FFT (2 x 64 KB ping-pong)
constexpr int CQ_MAXRAD = 14;      // stages per shared-memory FFT (8192 = 16*16*16*2 needs 4; 2^13 in radix 2 would need 13)
constexpr int CQ_TW_S = 1024;         // two-level twiddle: W^m = hi[m / S] * lo[m % S]
constexpr int CQ_THREADS = 256;
constexpr int CQ_LANES = HPFW_CTX_LANES;   // concurrent tracks of a batch (streams + scratch sets)
#ifndef CQ_FFT_THREADS_N
#define CQ_FFT_THREADS_N 256
#endif
constexpr int CQ_FFT_THREADS = CQ_FFT_THREADS_N;   // FFT kernels: 3 CTAs per SM (80 registers, <= 74 KB shared memory each)
#ifndef CQ_ROWS3_CTAS
#define CQ_ROWS3_CTAS 3
#endif
#ifndef CQ_PASS_CTAS
#define CQ_PASS_CTAS 3
#endif
constexpr int CQ_ROW_POINTS = 4096;   // CZT row pass: a CTA transforms G rows with G * L2 <= 4096 points

struct FftDesc {
    int n;
    int nrad;
    int rad[CQ_MAXRAD];
    int tw_off[CQ_MAXRAD];   // offset of stage s in the stage-twiddle table: entry [(t-1)*Ns + k] = e^{-2 pi i k t/(Ns R)}
    unsigned long long mg_m[CQ_MAXRAD];    // ceil(2^40 / (n / R_s)): x / m  = (x * mg_m) >> 40 for x < 2^20
    unsigned long long mg_ns[CQ_MAXRAD];   // ceil(2^40 / Ns_s)
};

// exact x / d for x < 2^20, d < 2^20 with mg = ceil(2^40 / d): one 64-bit multiply instead of an integer division
__device__ __forceinline__ int fastdiv(int x, unsigned long long mg) { return (int)(((unsigned long long)(unsigned)x * mg) >> 40); }

// Shared-memory index padding: one extra element per 8. A radix-8 Stockham stage with Ns = 1 writes element 8j+t from
// thread j (64-byte stride = 16-way bank conflict); padded, the stride is 72 bytes and a half-warp hits 16 distinct bank
// pairs. The same map keeps the later stages and the contiguous reads conflict-free.
#define CQ_PAD(i) ((i) + ((i) >> 3))

struct BandMeta {
    int first_bin;        // pos_j - floor(Lg_j / 2)
    int lg;               // Lg_j
    int half;             // floor(Lg_j / 2)
    int L;                // CZT length
    int L2;               // L / 16
    int btab;             // index of the chirp-filter spectrum table for this L
    int r0;               // > 0: L2 = 256 * r0 and the row pass runs czt_rows3_kernel; 0: generic shared-memory FFT
    long long work_off;   // offset (complex elements) of this band's L-point work area
};

struct RowTile {          // CZT row pass: rows [c0, c0 + g) of one band (contiguous in the work area)
    int band, c0, g;
};

// ------------------------------------------------------------------------------------------------ complex helpers
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// multiply by -i (sign < 0) or +i (sign > 0)
__device__ __forceinline__ float2 mul_i(float2 a, int sign) {
    return sign < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}
__device__ __forceinline__ float2 twiddle2(const float2 *__restrict__ hi, const float2 *__restrict__ lo, long long m,
                                           int sign) {
    float2 w = cmul(hi[m / CQ_TW_S], lo[m % CQ_TW_S]);
    return sign < 0 ? w : cconj(w);    // tables hold e^{-2 pi i m / P}
}

// ------------------------------------------------------------------------------------------------ small DFTs
// cos / sin of 2 pi i / R for the odd radices (compile-time constants once the butterfly loops are unrolled)
template <int R> __host__ __device__ constexpr float odd_cos(int i) {
    if (R == 3) return i == 0 ? 1.f : -0.5f;
    if (R == 5) return i == 0 ? 1.f : ((i == 1 || i == 4) ? 0.30901699437494745f : -0.8090169943749473f);
    return i == 0 ? 1.f
                  : ((i == 1 || i == 6) ? 0.6234898018587336f
                                        : ((i == 2 || i == 5) ? -0.2225209339563144f : -0.9009688679024191f));
}
template <int R> __host__ __device__ constexpr float odd_sin(int i) {
    if (R == 3) return i == 0 ? 0.f : (i == 1 ? 0.8660254037844386f : -0.8660254037844386f);
    if (R == 5)
        return i == 0 ? 0.f
                      : (i == 1 ? 0.9510565162951535f
                                : (i == 2 ? 0.5877852522924732f : (i == 3 ? -0.5877852522924732f : -0.9510565162951535f)));
    return i == 0 ? 0.f
                  : (i == 1 ? 0.7818314824680298f
                            : (i == 2 ? 0.9749279121818236f
                                      : (i == 3 ? 0.4338837391175581f
                                                : (i == 4 ? -0.4338837391175581f
                                                          : (i == 5 ? -0.9749279121818236f : -0.7818314824680298f)))));
}

// X_k = sum_m v_m e^{sign 2 pi i k m / R}, in place. Odd R through the symmetric pairs (v_m +- v_{R-m}).
template <int R> __device__ __forceinline__ void dft_odd(float2 (&v)[R], int sign) {
    constexpr int Hh = (R - 1) / 2;
    float2 sm[Hh], df[Hh];
#pragma unroll
    for (int m = 1; m <= Hh; ++m) {
        sm[m - 1] = cadd(v[m], v[R - m]);
        df[m - 1] = csub(v[m], v[R - m]);
    }
    float2 x0 = v[0];
    float2 out0 = x0;
#pragma unroll
    for (int m = 0; m < Hh; ++m) out0 = cadd(out0, sm[m]);
#pragma unroll
    for (int k = 1; k <= Hh; ++k) {
        float2 A = x0, B = make_float2(0.f, 0.f);
#pragma unroll
        for (int m = 1; m <= Hh; ++m) {
            const float c = odd_cos<R>((k * m) % R), s = odd_sin<R>((k * m) % R);
            A.x = fmaf(sm[m - 1].x, c, A.x);
            A.y = fmaf(sm[m - 1].y, c, A.y);
            B.x = fmaf(df[m - 1].x, s, B.x);
            B.y = fmaf(df[m - 1].y, s, B.y);
        }
        const float2 iB = mul_i(B, sign);    // sign<0: -i B
        v[k] = cadd(A, iB);
        v[R - k] = csub(A, iB);
    }
    v[0] = out0;
}

__device__ __forceinline__ void dft2(float2 (&v)[2]) {
    const float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
}
__device__ __forceinline__ void dft4(float2 &v0, float2 &v1, float2 &v2, float2 &v3, int sign) {
    const float2 t0 = cadd(v0, v2), t1 = csub(v0, v2), t2 = cadd(v1, v3), t3 = mul_i(csub(v1, v3), sign);
    v0 = cadd(t0, t2);
    v1 = cadd(t1, t3);
    v2 = csub(t0, t2);
    v3 = csub(t1, t3);
}
__device__ __forceinline__ void dft8(float2 (&v)[8], int sign) {
    float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    float2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    dft4(e0, e1, e2, e3, sign);
    dft4(o0, o1, o2, o3, sign);
    const float r = 0.7071067811865476f;
    // w8^k = e^{sign 2 pi i k / 8}
    const float2 w1 = make_float2(r, sign < 0 ? -r : r), w3 = make_float2(-r, sign < 0 ? -r : r);
    o1 = cmul(o1, w1);
    o2 = mul_i(o2, sign);
    o3 = cmul(o3, w3);
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
    v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
    v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}
// 16-point DFT in registers: a = 4 a1 + a0 -> c = c0 + 4 c1 (see file header of the derivation in DESIGN.md)
__device__ __forceinline__ void dft16(float2 (&v)[16], int sign) {
    constexpr float C[4] = {1.f, 0.9238795325112867f, 0.7071067811865476f, 0.3826834323650898f};   // cos(2 pi m/16)
    constexpr float S[4] = {0.f, 0.3826834323650898f, 0.7071067811865476f, 0.9238795325112867f};   // sin(2 pi m/16)
    float2 y[4][4];   // y[a0][c0]
#pragma unroll
    for (int a0 = 0; a0 < 4; ++a0) {
        float2 t0 = v[a0], t1 = v[4 + a0], t2 = v[8 + a0], t3 = v[12 + a0];
        dft4(t0, t1, t2, t3, sign);
        y[a0][0] = t0; y[a0][1] = t1; y[a0][2] = t2; y[a0][3] = t3;
    }
#pragma unroll
    for (int a0 = 1; a0 < 4; ++a0)
#pragma unroll
        for (int c0 = 1; c0 < 4; ++c0) {
            const int m = a0 * c0;   // 1,2,3,4,6,9
            float c, s;
            if (m <= 3) { c = C[m]; s = S[m]; }
            else if (m == 4) { c = 0.f; s = 1.f; }
            else if (m == 6) { c = -C[2]; s = S[2]; }
            else { c = -C[1]; s = -S[1]; }     // m == 9: cos(9 pi/8) = -cos(pi/8), sin(9 pi/8) = -sin(pi/8)
            const float2 w = make_float2(c, sign < 0 ? -s : s);
            y[a0][c0] = cmul(y[a0][c0], w);
        }
#pragma unroll
    for (int c0 = 0; c0 < 4; ++c0) {
        float2 t0 = y[0][c0], t1 = y[1][c0], t2 = y[2][c0], t3 = y[3][c0];
        dft4(t0, t1, t2, t3, sign);
        v[c0] = t0; v[c0 + 4] = t1; v[c0 + 8] = t2; v[c0 + 12] = t3;
    }
}

// 32-point DFT in registers (natural order in and out): two 16-point DFTs of the even / odd inputs and one radix-2 stage.
// Used by the chirp-z column passes of tracks longer than 6.7 minutes (column length 32 instead of 16).
__device__ __forceinline__ void dft32(float2 (&v)[32], int sign) {
    constexpr float C[16] = {1.f, 0.9807852804032304f, 0.9238795325112867f, 0.8314696123025452f, 0.7071067811865476f,
                             0.5555702330196022f, 0.3826834323650898f, 0.19509032201612825f, 0.f, -0.19509032201612825f,
                             -0.3826834323650898f, -0.5555702330196022f, -0.7071067811865476f, -0.8314696123025452f,
                             -0.9238795325112867f, -0.9807852804032304f};                       // cos(2 pi k / 32)
    constexpr float S[16] = {0.f, 0.19509032201612825f, 0.3826834323650898f, 0.5555702330196022f, 0.7071067811865476f,
                             0.8314696123025452f, 0.9238795325112867f, 0.9807852804032304f, 1.f, 0.9807852804032304f,
                             0.9238795325112867f, 0.8314696123025452f, 0.7071067811865476f, 0.5555702330196022f,
                             0.3826834323650898f, 0.19509032201612825f};                        // sin(2 pi k / 32)
    float2 e[16], o[16];
#pragma unroll
    for (int a = 0; a < 16; ++a) {
        e[a] = v[2 * a];
        o[a] = v[2 * a + 1];
    }
    dft16(e, sign);
    dft16(o, sign);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const float2 t = cmul(o[k], make_float2(C[k], sign < 0 ? -S[k] : S[k]));
        v[k] = cadd(e[k], t);
        v[k + 16] = csub(e[k], t);
    }
}
// 64-point DFT in registers (natural order in and out): a = 4 a1 + a0 -> four 16-point DFTs over a1, twiddle W_64^(a0 c0),
// radix-4 over a0: X[c0 + 16 c1]. Column pass of tracks between 13.5 and 27 minutes.
__device__ __forceinline__ void dft64(float2 (&v)[64], int sign) {
    constexpr float C[46] = {1.0000000000f, 0.9951847267f, 0.9807852804f, 0.9569403357f, 0.9238795325f,
        0.8819212643f, 0.8314696123f, 0.7730104534f, 0.7071067812f, 0.6343932842f, 0.5555702330f, 0.4713967368f,
        0.3826834324f, 0.2902846773f, 0.1950903220f, 0.0980171403f, 0.0000000000f, -0.0980171403f, -0.1950903220f,
        -0.2902846773f, -0.3826834324f, -0.4713967368f, -0.5555702330f, -0.6343932842f, -0.7071067812f,
        -0.7730104534f, -0.8314696123f, -0.8819212643f, -0.9238795325f, -0.9569403357f, -0.9807852804f,
        -0.9951847267f, -1.0000000000f, -0.9951847267f, -0.9807852804f, -0.9569403357f, -0.9238795325f,
        -0.8819212643f, -0.8314696123f, -0.7730104534f, -0.7071067812f, -0.6343932842f, -0.5555702330f,
        -0.4713967368f, -0.3826834324f, -0.2902846773f};   // cos(2 pi m / 64), m = a0 c0 <= 45
    constexpr float S[46] = {0.0000000000f, 0.0980171403f, 0.1950903220f, 0.2902846773f, 0.3826834324f,
        0.4713967368f, 0.5555702330f, 0.6343932842f, 0.7071067812f, 0.7730104534f, 0.8314696123f, 0.8819212643f,
        0.9238795325f, 0.9569403357f, 0.9807852804f, 0.9951847267f, 1.0000000000f, 0.9951847267f, 0.9807852804f,
        0.9569403357f, 0.9238795325f, 0.8819212643f, 0.8314696123f, 0.7730104534f, 0.7071067812f, 0.6343932842f,
        0.5555702330f, 0.4713967368f, 0.3826834324f, 0.2902846773f, 0.1950903220f, 0.0980171403f, 0.0000000000f,
        -0.0980171403f, -0.1950903220f, -0.2902846773f, -0.3826834324f, -0.4713967368f, -0.5555702330f,
        -0.6343932842f, -0.7071067812f, -0.7730104534f, -0.8314696123f, -0.8819212643f, -0.9238795325f,
        -0.9569403357f};   // sin(2 pi m / 64)
    float2 y[4][16];
#pragma unroll
    for (int a0 = 0; a0 < 4; ++a0) {
#pragma unroll
        for (int a1 = 0; a1 < 16; ++a1) y[a0][a1] = v[4 * a1 + a0];
        dft16(y[a0], sign);
    }
#pragma unroll
    for (int c0 = 0; c0 < 16; ++c0) {
        float2 t0 = y[0][c0], t1 = y[1][c0], t2 = y[2][c0], t3 = y[3][c0];
        if (c0 > 0) {
            t1 = cmul(t1, make_float2(C[c0], sign < 0 ? -S[c0] : S[c0]));
            t2 = cmul(t2, make_float2(C[2 * c0], sign < 0 ? -S[2 * c0] : S[2 * c0]));
            t3 = cmul(t3, make_float2(C[3 * c0], sign < 0 ? -S[3 * c0] : S[3 * c0]));
        }
        dft4(t0, t1, t2, t3, sign);
        v[c0] = t0; v[c0 + 16] = t1; v[c0 + 32] = t2; v[c0 + 48] = t3;
    }
}
template <int L1> __device__ __forceinline__ void dft_col(float2 (&v)[L1], int sign);
template <> __device__ __forceinline__ void dft_col<16>(float2 (&v)[16], int sign) { dft16(v, sign); }
template <> __device__ __forceinline__ void dft_col<32>(float2 (&v)[32], int sign) { dft32(v, sign); }
template <> __device__ __forceinline__ void dft_col<64>(float2 (&v)[64], int sign) { dft64(v, sign); }

template <int R> __device__ __forceinline__ void dft_r(float2 (&v)[R], int sign);
template <> __device__ __forceinline__ void dft_r<2>(float2 (&v)[2], int) { dft2(v); }
template <> __device__ __forceinline__ void dft_r<3>(float2 (&v)[3], int sign) { dft_odd<3>(v, sign); }
template <> __device__ __forceinline__ void dft_r<4>(float2 (&v)[4], int sign) { dft4(v[0], v[1], v[2], v[3], sign); }
template <> __device__ __forceinline__ void dft_r<5>(float2 (&v)[5], int sign) { dft_odd<5>(v, sign); }
template <> __device__ __forceinline__ void dft_r<7>(float2 (&v)[7], int sign) { dft_odd<7>(v, sign); }
template <> __device__ __forceinline__ void dft_r<8>(float2 (&v)[8], int sign) { dft8(v, sign); }
template <> __device__ __forceinline__ void dft_r<16>(float2 (&v)[16], int sign) { dft16(v, sign); }

// cos / sin of 2 pi m / R for the composite radices (immediates once the butterfly loops are unrolled)
template <int R> struct TwTab;
template <> struct TwTab<6> {
    static __device__ __forceinline__ float c(int m) {
        constexpr float t[6] = {1.0f, 0.5f, -0.5f, -1.0f, -0.5f, 0.5f};
        return t[m];
    }
    static __device__ __forceinline__ float s(int m) {
        constexpr float t[6] = {0.0f, 0.8660254037844386f, 0.86602540378443871f, 0.0f, -0.86602540378443837f, -0.8660254037844386f};
        return t[m];
    }
};
template <> struct TwTab<9> {
    static __device__ __forceinline__ float c(int m) {
        constexpr float t[9] = {1.0f, 0.76604444311897801f, 0.17364817766693041f, -0.5f, -0.93969262078590832f, -0.93969262078590843f, -0.5f, 0.17364817766692997f, 0.76604444311897779f};
        return t[m];
    }
    static __device__ __forceinline__ float s(int m) {
        constexpr float t[9] = {0.0f, 0.64278760968653925f, 0.98480775301220802f, 0.86602540378443871f, 0.34202014332566888f, -0.34202014332566866f, -0.86602540378443837f, -0.98480775301220813f, -0.64278760968653958f};
        return t[m];
    }
};
template <> struct TwTab<10> {
    static __device__ __forceinline__ float c(int m) {
        constexpr float t[10] = {1.0f, 0.80901699437494745f, 0.30901699437494745f, -0.30901699437494734f, -0.80901699437494734f, -1.0f, -0.80901699437494756f, -0.30901699437494756f, 0.30901699437494723f, 0.80901699437494734f};
        return t[m];
    }
    static __device__ __forceinline__ float s(int m) {
        constexpr float t[10] = {0.0f, 0.58778525229247314f, 0.95105651629515353f, 0.95105651629515364f, 0.58778525229247325f, 0.0f, -0.58778525229247303f, -0.95105651629515353f, -0.95105651629515364f, -0.58778525229247336f};
        return t[m];
    }
};
template <> struct TwTab<12> {
    static __device__ __forceinline__ float c(int m) {
        constexpr float t[12] = {1.0f, 0.86602540378443871f, 0.5f, 0.0f, -0.5f, -0.86602540378443871f, -1.0f, -0.86602540378443882f, -0.5f, 0.0f, 0.5f, 0.86602540378443837f};
        return t[m];
    }
    static __device__ __forceinline__ float s(int m) {
        constexpr float t[12] = {0.0f, 0.5f, 0.8660254037844386f, 1.0f, 0.86602540378443871f, 0.5f, 0.0f, -0.5f, -0.86602540378443837f, -1.0f, -0.8660254037844386f, -0.5f};
        return t[m];
    }
};
template <> struct TwTab<14> {
    static __device__ __forceinline__ float c(int m) {
        constexpr float t[14] = {1.0f, 0.90096886790241915f, 0.62348980185873359f, 0.22252093395631445f, -0.22252093395631434f, -0.62348980185873348f, -0.90096886790241903f, -1.0f, -0.90096886790241915f, -0.62348980185873371f, -0.22252093395631459f, 0.22252093395631334f, 0.62348980185873337f, 0.90096886790241937f};
        return t[m];
    }
    static __device__ __forceinline__ float s(int m) {
        constexpr float t[14] = {0.0f, 0.43388373911755812f, 0.7818314824680298f, 0.97492791218182362f, 0.97492791218182362f, 0.78183148246802991f, 0.43388373911755823f, 0.0f, -0.43388373911755801f, -0.78183148246802969f, -0.97492791218182362f, -0.97492791218182384f, -0.78183148246802991f, -0.43388373911755751f};
        return t[m];
    }
};
template <> struct TwTab<15> {
    static __device__ __forceinline__ float c(int m) {
        constexpr float t[15] = {1.0f, 0.91354545764260087f, 0.66913060635885824f, 0.30901699437494745f, -0.10452846326765333f, -0.5f, -0.80901699437494734f, -0.97814760073380569f, -0.97814760073380569f, -0.80901699437494756f, -0.5f, -0.10452846326765423f, 0.30901699437494723f, 0.66913060635885846f, 0.91354545764260098f};
        return t[m];
    }
    static __device__ __forceinline__ float s(int m) {
        constexpr float t[15] = {0.0f, 0.40673664307580015f, 0.74314482547739413f, 0.95105651629515353f, 0.9945218953682734f, 0.86602540378443871f, 0.58778525229247325f, 0.20791169081775931f, -0.20791169081775907f, -0.58778525229247303f, -0.86602540378443837f, -0.99452189536827329f, -0.95105651629515364f, -0.74314482547739402f, -0.40673664307580015f};
        return t[m];
    }
};

// Composite radix R = R1 * R2 in registers (Cooley-Tukey inside the butterfly): input a = R2 a1 + a0, output
// c = c0 + R1 c1:  X[c] = sum_a0 W_R2^{a0 c1} W_R^{a0 c0} sum_a1 v[R2 a1 + a0] W_R1^{a1 c0}. One Stockham stage of radix
// 14 / 15 / 16 replaces two stages (and their shared-memory round trip and barrier) of radix 2x7 / 3x5 / 4x4.
template <int R1, int R2> __device__ __forceinline__ void dft_comp(float2 (&v)[R1 * R2], int sign) {
    constexpr int R = R1 * R2;
    float2 y[R2][R1];
#pragma unroll
    for (int a0 = 0; a0 < R2; ++a0) {
        float2 t[R1];
#pragma unroll
        for (int a1 = 0; a1 < R1; ++a1) t[a1] = v[R2 * a1 + a0];
        dft_r<R1>(t, sign);
#pragma unroll
        for (int c0 = 0; c0 < R1; ++c0) y[a0][c0] = t[c0];
    }
#pragma unroll
    for (int a0 = 1; a0 < R2; ++a0)
#pragma unroll
        for (int c0 = 1; c0 < R1; ++c0) {
            const int m = (a0 * c0) % R;
            const float c = TwTab<R>::c(m), sn = TwTab<R>::s(m);
            y[a0][c0] = cmul(y[a0][c0], make_float2(c, sign < 0 ? -sn : sn));
        }
#pragma unroll
    for (int c0 = 0; c0 < R1; ++c0) {
        float2 t[R2];
#pragma unroll
        for (int a0 = 0; a0 < R2; ++a0) t[a0] = y[a0][c0];
        dft_r<R2>(t, sign);
#pragma unroll
        for (int c1 = 0; c1 < R2; ++c1) v[c0 + R1 * c1] = t[c1];
    }
}
template <> __device__ __forceinline__ void dft_r<6>(float2 (&v)[6], int sign) { dft_comp<3, 2>(v, sign); }
template <> __device__ __forceinline__ void dft_r<9>(float2 (&v)[9], int sign) { dft_comp<3, 3>(v, sign); }
template <> __device__ __forceinline__ void dft_r<10>(float2 (&v)[10], int sign) { dft_comp<5, 2>(v, sign); }
template <> __device__ __forceinline__ void dft_r<12>(float2 (&v)[12], int sign) { dft_comp<4, 3>(v, sign); }
template <> __device__ __forceinline__ void dft_r<14>(float2 (&v)[14], int sign) { dft_comp<7, 2>(v, sign); }
template <> __device__ __forceinline__ void dft_r<15>(float2 (&v)[15], int sign) { dft_comp<5, 3>(v, sign); }

// ------------------------------------------------------------------------------------------------ shared-memory FFT
// One Stockham autosort stage of radix R over G sequences of length n (sequence g at logical index g*n; every shared-memory
// access goes through CQ_PAD). tws = this stage's twiddles, laid out [(t-1)*Ns + k] so that consecutive threads
// (consecutive k) read consecutive entries.
template <int R>
__device__ __forceinline__ void stockham_stage(const float2 *in, float2 *out, int n, int G, int Ns,
                                               const float2 *__restrict__ tws, int sign, unsigned long long mg_m,
                                               unsigned long long mg_ns) {
    const int m = n / R;
    for (int idx = threadIdx.x; idx < G * m; idx += blockDim.x) {
        const int g = (G == 1) ? 0 : fastdiv(idx, mg_m), j = idx - g * m;
        const int blk = fastdiv(j, mg_ns), k = j - blk * Ns;
        float2 v[R];
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = in[CQ_PAD(g * n + j + t * m)];
        if (Ns > 1) {
#pragma unroll
            for (int t = 1; t < R; ++t) {
                float2 w = tws[(t - 1) * Ns + k];          // e^{-2 pi i k t / (Ns R)}
                if (sign > 0) w = cconj(w);
                v[t] = cmul(v[t], w);
            }
        }
        dft_r<R>(v, sign);
        const int base = g * n + blk * Ns * R + k;
#pragma unroll
        for (int t = 0; t < R; ++t) out[CQ_PAD(base + t * Ns)] = v[t];
    }
}

// FFT of G sequences of length d.n held (CQ_PAD-indexed) in `a`; `b` is scratch of the same size; `tw` = the stage-twiddle
// table of d (< d.n entries). Returns the buffer holding the result (natural order, CQ_PAD-indexed). All threads of the CTA must call it; the data in `a` must be visible (caller syncs before).
__device__ float2 *smem_fft(float2 *a, float2 *b, const FftDesc &d, int G, const float2 *__restrict__ tw, int sign) {
    int Ns = 1;
    for (int s = 0; s < d.nrad; ++s) {
        const int r = d.rad[s];
        switch (r) {
            case 2: stockham_stage<2>(a, b, d.n, G, Ns, tw + d.tw_off[s], sign, d.mg_m[s], d.mg_ns[s]); break;
            case 3: stockham_stage<3>(a, b, d.n, G, Ns, tw + d.tw_off[s], sign, d.mg_m[s], d.mg_ns[s]); break;
            case 4: stockham_stage<4>(a, b, d.n, G, Ns, tw + d.tw_off[s], sign, d.mg_m[s], d.mg_ns[s]); break;
            case 5: stockham_stage<5>(a, b, d.n, G, Ns, tw + d.tw_off[s], sign, d.mg_m[s], d.mg_ns[s]); break;
            case 7: stockham_stage<7>(a, b, d.n, G, Ns, tw + d.tw_off[s], sign, d.mg_m[s], d.mg_ns[s]); break;
            case 6: stockham_stage<6>(a, b, d.n, G, Ns, tw + d.tw_off[s], sign, d.mg_m[s], d.mg_ns[s]); break;
            case 9: stockham_stage<9>(a, b, d.n, G, Ns, tw + d.tw_off[s], sign, d.mg_m[s], d.mg_ns[s]); break;
            case 10: stockham_stage<10>(a, b, d.n, G, Ns, tw + d.tw_off[s], sign, d.mg_m[s], d.mg_ns[s]); break;
            case 12: stockham_stage<12>(a, b, d.n, G, Ns, tw + d.tw_off[s], sign, d.mg_m[s], d.mg_ns[s]); break;
            case 14: stockham_stage<14>(a, b, d.n, G, Ns, tw + d.tw_off[s], sign, d.mg_m[s], d.mg_ns[s]); break;
            case 15: stockham_stage<15>(a, b, d.n, G, Ns, tw + d.tw_off[s], sign, d.mg_m[s], d.mg_ns[s]); break;
            case 16: stockham_stage<16>(a, b, d.n, G, Ns, tw + d.tw_off[s], sign, d.mg_m[s], d.mg_ns[s]); break;
            default: stockham_stage<8>(a, b, d.n, G, Ns, tw + d.tw_off[s], sign, d.mg_m[s], d.mg_ns[s]); break;
        }
        __syncthreads();
        float2 *t = a; a = b; b = t;
        Ns *= r;
    }
    return a;
}

// ------------------------------------------------------------------------------------------------ register-FFT helpers
// shared-memory index padding of the register kernels: one extra element per 16 (thread strides of 2..16 elements stay
// conflict-free or 2-way)
__device__ __forceinline__ int cr_pad(int i) { return i + (i >> 4); }
#ifndef HPFW_CQT_UNROLL
#define HPFW_CQT_UNROLL 1
#endif
#define HPFW_CQT_PRAGMA_(x) _Pragma(#x)
#define HPFW_CQT_PRAGMA(x) HPFW_CQT_PRAGMA_(x)
#define HPFW_CQT_IDX_UNROLL HPFW_CQT_PRAGMA(unroll HPFW_CQT_UNROLL)
#ifndef HPFW_CQT_CHAIN_TW
#define HPFW_CQT_CHAIN_TW 0
#endif
#ifndef HPFW_CQT_NOPAD0
#define HPFW_CQT_NOPAD0 0
#endif
// pass kernels: MODE 0 (columns interleaved, every access linear in the thread index) can run without padding
template <int MODE> __device__ __forceinline__ int ps_pad(int i) { return (MODE == 0 && HPFW_CQT_NOPAD0) ? i : cr_pad(i); }
__host__ __device__ constexpr int hibit(int u) { int h = 1; while (h * 2 <= u) h *= 2; return h; }

// v[u] *= w^u, u = 1 .. R-1
template <int R> __device__ __forceinline__ void apply_powers(float2 (&v)[R], float2 w) {
    float2 p[R];
    p[0] = make_float2(1.f, 0.f);
    if (R > 1) p[1] = w;
#pragma unroll
    for (int u = 2; u < R; ++u) {
        const int hi = hibit(u), lo = u - hi;
        p[u] = lo == 0 ? cmul(p[hi / 2], p[hi / 2]) : cmul(p[hi], p[lo]);
    }
#pragma unroll
    for (int u = 1; u < R; ++u) v[u] = cmul(v[u], p[u]);
}

// ------------------------------------------------------------------------------------------------ main FFT passes
// Two-pass ("four-step") FFT of H = n1 * n2 points, element (a, b) at a*n2 + b:
//   MODE 0 (pass A): length-n1 FFTs down G adjacent columns per CTA, out[c*n2 + b] = W_H^{b c} * sum_a in[a*n2 + b] W_n1^{a c};
//   MODE 1 (pass B): length-n2 FFTs along G rows per CTA, Z[c + n1*d] = sum_b in[c*n2 + b] W_n2^{b d}, keeping only
//           k in [klo, khi] (-> out_lo[k - klo]) and k in [H - khi, H - klo] (-> out_hi[k - (H - khi)]); keep_all != 0
//           stores the whole spectrum to out_lo[k].
// Each length-n FFT is a Stockham autosort over the plan's radices (2..16, composite radices in registers). The first
// stage reads global memory straight into its butterflies and the last stage applies the pass's output twiddle / bin
// selection from registers, so an element makes (stages - 1) shared-memory round trips instead of (stages + 2). A
// butterfly loads ONE twiddle W_n^{step k} (L1-resident table of n entries) and forms its powers by a product tree.
__device__ __forceinline__ float2 tw_dir(float2 w, int sign) { return sign < 0 ? w : cconj(w); }   // tables hold e^{-2 pi i ./n}

template <> __device__ __forceinline__ void dft_r<1>(float2 (&)[1], int) {}

template <int R, int MODE>
__device__ __forceinline__ void pass_first(const float2 *__restrict__ in, float2 *S, int n, int G, int g_here, int pitch,
                                           int sign, unsigned long long mg_g, unsigned long long mg_m) {
    const int m = n / R, tot = G * m;
    HPFW_CQT_IDX_UNROLL
    for (int idx = threadIdx.x; idx < tot; idx += blockDim.x) {
        int g, j;
        if (MODE == 0) { j = fastdiv(idx, mg_g); g = idx - j * G; }     // adjacent columns are adjacent in memory
        else { g = fastdiv(idx, mg_m); j = idx - g * m; }
        float2 v[R];
        if (g < g_here) {
#pragma unroll
            for (int u = 0; u < R; ++u) v[u] = MODE == 0 ? in[(j + u * m) * pitch + g] : in[g * pitch + j + u * m];
        } else {
#pragma unroll
            for (int u = 0; u < R; ++u) v[u] = make_float2(0.f, 0.f);
        }
        dft_r<R>(v, sign);
#pragma unroll
        for (int u = 0; u < R; ++u) S[ps_pad<MODE>(MODE == 0 ? (j * R + u) * G + g : g * n + j * R + u)] = v[u];
    }
}

// Shared-memory layout of a pass: MODE 0 interleaves the G columns (position p of column g at p*G + g) and maps threads
// with g fastest, so that a warp's accesses are linear in the thread index in every stage and the final column-adjacent
// store reads shared memory linearly; MODE 1 keeps each row contiguous (g*n + p) with the butterfly index fastest.
template <int R, int MODE>
__device__ __forceinline__ void pass_mid(const float2 *Sin, float2 *Sout, int n, int G, int Ns, const float2 *__restrict__ T,
                                         int sign, unsigned long long mg_g, unsigned long long mg_m,
                                         unsigned long long mg_ns) {
    const int m = n / R, tot = G * m, step = m / Ns;      // W_{Ns R}^k = W_n^{step k}
    HPFW_CQT_IDX_UNROLL
    for (int idx = threadIdx.x; idx < tot; idx += blockDim.x) {
        int g, j;
        if (MODE == 0) { j = fastdiv(idx, mg_g); g = idx - j * G; }
        else { g = fastdiv(idx, mg_m); j = idx - g * m; }
        const int blk = fastdiv(j, mg_ns), k = j - blk * Ns;
        float2 v[R];
#pragma unroll
        for (int u = 0; u < R; ++u) v[u] = Sin[ps_pad<MODE>(MODE == 0 ? idx + u * tot : g * n + j + u * m)];
        apply_powers<R>(v, tw_dir(T[step * k], sign));
        dft_r<R>(v, sign);
        const int base = blk * Ns * R + k;
#pragma unroll
        for (int u = 0; u < R; ++u) Sout[ps_pad<MODE>(MODE == 0 ? (base + u * Ns) * G + g : g * n + base + u * Ns)] = v[u];
    }
}

struct PassOut {
    float2 *lo, *hi;          // MODE 0: lo = out (+ b0); MODE 1: the two kept ranges
    const float2 *tw_hi, *tw_lo;
    int other;                // MODE 0: n2 (row pitch); MODE 1: n1
    int first;                // MODE 0: b0; MODE 1: c0
    int klo, khi, H, keep_all;
};

template <int R, int MODE>
__device__ __forceinline__ void pass_last(const float2 *Sin, float2 *Sout, int n, int G, int g_here,
                                          const float2 *__restrict__ T, int sign, unsigned long long mg_g,
                                          unsigned long long mg_m, const PassOut &po) {
    const int m = n / R, tot = G * m;       // Ns = m: blk = 0, k = j; outputs d = j + u m
    HPFW_CQT_IDX_UNROLL
    for (int idx = threadIdx.x; idx < tot; idx += blockDim.x) {
        int g, j;
        if (MODE == 0) { j = fastdiv(idx, mg_g); g = idx - j * G; }
        else { g = fastdiv(idx, mg_m); j = idx - g * m; }
        float2 v[R];
#pragma unroll
        for (int u = 0; u < R; ++u) v[u] = Sin[ps_pad<MODE>(MODE == 0 ? idx + u * tot : g * n + j + u * m)];
        if (R > 1) apply_powers<R>(v, tw_dir(T[j], sign));
        dft_r<R>(v, sign);
        if (MODE == 0) {
            // inter-pass twiddle W_H^{b d} = W_H^{b j} (W_H^{b m})^u, then staged for the column-adjacent store
            const int bcol = po.first + g;
            const float2 wa = twiddle2(po.tw_hi, po.tw_lo, (long long)bcol * j, sign);
#if HPFW_CQT_CHAIN_TW
            // q_u = wa * wm^u by a running product: R - 1 complex multiplies instead of a power tree plus a second multiply
            float2 q = wa;
            const float2 wm = twiddle2(po.tw_hi, po.tw_lo, (long long)bcol * m, sign);
#pragma unroll
            for (int u = 0; u < R; ++u) {
                Sout[ps_pad<MODE>(idx + u * tot)] = cmul(v[u], q);
                if (u + 1 < R) q = cmul(q, wm);
            }
#else
            if (R > 1) apply_powers<R>(v, twiddle2(po.tw_hi, po.tw_lo, (long long)bcol * m, sign));
#pragma unroll
            for (int u = 0; u < R; ++u) Sout[ps_pad<MODE>(idx + u * tot)] = cmul(v[u], wa);
#endif
        } else if (g < g_here) {
            const int c = po.first + g, mlo = po.H - po.khi, mhi = po.H - po.klo;
#pragma unroll
            for (int u = 0; u < R; ++u) {
                const int k = c + po.other * (j + u * m);
                if (po.keep_all) {
                    po.lo[k] = v[u];
                } else {
                    if (k >= po.klo && k <= po.khi) po.lo[k - po.klo] = v[u];
                    if (po.hi != nullptr && k >= mlo && k <= mhi) po.hi[k - mlo] = v[u];
                }
            }
        }
    }
}

#define HPFW_RADIX_SWITCH(r, CALL)                                                                                      \
    switch (r) {                                                                                                        \
        case 2: { constexpr int R = 2; CALL; } break;                                                                   \
        case 3: { constexpr int R = 3; CALL; } break;                                                                   \
        case 4: { constexpr int R = 4; CALL; } break;                                                                   \
        case 5: { constexpr int R = 5; CALL; } break;                                                                   \
        case 6: { constexpr int R = 6; CALL; } break;                                                                   \
        case 7: { constexpr int R = 7; CALL; } break;                                                                   \
        case 8: { constexpr int R = 8; CALL; } break;                                                                   \
        case 9: { constexpr int R = 9; CALL; } break;                                                                   \
        case 10: { constexpr int R = 10; CALL; } break;                                                                 \
        case 12: { constexpr int R = 12; CALL; } break;                                                                 \
        case 14: { constexpr int R = 14; CALL; } break;                                                                 \
        case 15: { constexpr int R = 15; CALL; } break;                                                                 \
        case 16: { constexpr int R = 16; CALL; } break;                                                                 \
        default: { constexpr int R = 1; CALL; } break;                                                                  \
    }

#ifdef HPFW_CQT_PHASECLK
// tuning aid (never in the shipped build): clock cycles thread 0 of every CTA spends in each phase of a pass, summed
__device__ unsigned long long g_phase_clk[2][6];
#define HPFW_PHASE_MARK(i)                                                                   \
    if (threadIdx.x == 0) {                                                                  \
        const long long now_ = clock64();                                                    \
        atomicAdd(&g_phase_clk[MODE][i], (unsigned long long)(now_ - phase_t_));             \
        phase_t_ = now_;                                                                     \
    }
#else
#define HPFW_PHASE_MARK(i)
#endif

template <int MODE>
__global__ void __launch_bounds__(CQ_FFT_THREADS, CQ_PASS_CTAS)
fft_pass_kernel(const float2 *__restrict__ in, float2 *__restrict__ out_lo, float2 *__restrict__ out_hi, FftDesc d,
                int other, int G, unsigned long long mg_g, const float2 *__restrict__ T,
                const float2 *__restrict__ twH_hi, const float2 *__restrict__ twH_lo, int klo, int khi, int H,
                int keep_all, int sign) {
    extern __shared__ __align__(16) float2 fsm[];
    const int n = d.n;
    const int first = blockIdx.x * G;
    const int pitch = MODE == 0 ? other : n;                   // row pitch n2 of the n1 x n2 matrix
    const int g_here = min(G, (MODE == 0 ? other : other) - first);   // MODE 0: columns left (n2); MODE 1: rows left (n1)
    float2 *S0 = fsm, *S1 = fsm + ps_pad<MODE>(G * n) + 16;
    const float2 *src = MODE == 0 ? in + first : in + (long long)first * pitch;
    PassOut po{MODE == 0 ? out_lo + first : out_lo, out_hi, twH_hi, twH_lo, other, first, klo, khi, H, keep_all};

#ifdef HPFW_CQT_PHASECLK
    long long phase_t_ = clock64();
#endif
    HPFW_RADIX_SWITCH(d.rad[0], (pass_first<R, MODE>(src, S0, n, G, g_here, pitch, sign, mg_g, d.mg_m[0])));
    HPFW_PHASE_MARK(0)        // thread 0's own first-stage work (global loads + butterflies)
    __syncthreads();
    HPFW_PHASE_MARK(1)        // its wait at the barrier
    int Ns = d.rad[0];
    float2 *cur = S0, *nxt = S1;
    for (int s = 1; s + 1 < d.nrad; ++s) {
        HPFW_RADIX_SWITCH(d.rad[s], (pass_mid<R, MODE>(cur, nxt, n, G, Ns, T, sign, mg_g, d.mg_m[s], d.mg_ns[s])));
        __syncthreads();
        float2 *t = cur; cur = nxt; nxt = t;
        Ns *= d.rad[s];
    }
    HPFW_PHASE_MARK(2)        // middle stages (with their barriers)
    const int sl = d.nrad - 1;
    HPFW_RADIX_SWITCH(d.rad[sl], (pass_last<R, MODE>(cur, nxt, n, G, g_here, T, sign, mg_g, d.mg_m[sl], po)));
    HPFW_PHASE_MARK(3)        // last stage
    if (MODE == 0) {
        __syncthreads();
        HPFW_PHASE_MARK(4)
        // nxt holds out-values at [g][c]; store with the G adjacent columns contiguous (32-byte runs for G = 4)
        const int tot = G * n;
#pragma unroll 4
        for (int idx = threadIdx.x; idx < tot; idx += blockDim.x) {
            const int c = fastdiv(idx, mg_g), g = idx - c * G;
            if (g < g_here) po.lo[c * pitch + g] = nxt[ps_pad<MODE>(idx)];
        }
        HPFW_PHASE_MARK(5)    // column-adjacent store
    }
}

// ------------------------------------------------------------------------------------------------ Bluestein (any N)
// Audio lengths whose half is not a product of 2,3,5,7 (or that are odd) cannot use the packed two-pass FFT. For those the
// wanted bins X[klo..khi] of the length-N DFT are a chirp convolution, n k = (n^2 + k^2 - (k-n)^2) / 2:
//   X[k] = c[k] * sum_n (x[n] c[n]) b[k - n],  c[m] = e^{-i pi m^2 / N},  b = conj(c),
// evaluated by the same pass kernels on a smooth length P >= N + (khi - klo): FFT_P(a) .* FFT_P(b') -> IFFT_P, where b' is
// b shifted by klo so that output index 0 is bin klo. Phases are reduced mod 2N in 64-bit integers before sincospi.
__device__ __forceinline__ float2 bl_chirp(long long m, long long N, int sign) {     // e^{sign i pi m^2 / N}
    const unsigned long long r = ((unsigned long long)(m * m)) % (2ull * (unsigned long long)N);
    float s, c;
    sincospif((float)((double)r / (double)N), &s, &c);
    return make_float2(c, sign < 0 ? -s : s);
}
__global__ void __launch_bounds__(CQ_THREADS)
bl_prep_kernel(const float *__restrict__ x, long long N, int P, float2 *__restrict__ a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float2 v = make_float2(0.f, 0.f);
    if (i < N) {
        const float2 c = bl_chirp(i, N, -1);
        const float xv = x[i];
        v = make_float2(xv * c.x, xv * c.y);
    }
    a[i] = v;
}
__global__ void __launch_bounds__(CQ_THREADS)
bl_filter_kernel(long long N, int P, int klo, int K, float2 *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float2 v = make_float2(0.f, 0.f);
    if (i < K) v = bl_chirp((long long)i + klo, N, +1);
    else if ((long long)i > (long long)P - N) v = bl_chirp((long long)i - P + klo, N, +1);     // m' = i - P in (-N, 0)
    out[i] = v;
}
__global__ void __launch_bounds__(CQ_THREADS)
bl_mul_kernel(float2 *__restrict__ a, const float2 *__restrict__ b, int P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) a[i] = cmul(a[i], b[i]);
}
__global__ void __launch_bounds__(CQ_THREADS)
bl_post_kernel(float2 *__restrict__ x, long long N, int P, int klo, int K) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    const float2 c = bl_chirp((long long)i + klo, N, -1);
    const float inv = 1.0f / (float)P;
    const float2 v = cmul(x[i], c);
    x[i] = make_float2(v.x * inv, v.y * inv);
}

// The band window at index kp = k + floor(Lg/2), k = -floor(Lg/2) .. ceil(Lg/2)-1 (the reference passes "window","hann" to
// NSGConstantQ, /root/reference/include/hpfw/spectrum/cqt.h:58; what essentia builds from that name cannot be checked here):
//   0 (default) periodic Hann centred on k = 0:  0.5 + 0.5 cos(2 pi k / Lg)          (the NSG toolbox's winfuns('hann'))
//   1           symmetric Hann over the Lg taps: 0.5 - 0.5 cos(2 pi kp / (Lg - 1))    (a generic "hann" window of size Lg,
//               rotated so that tap floor(Lg/2) sits on the band centre)
// Selected per context: HPFW_CQT_WINDOW={periodic,symmetric} or hpfw_set_cqt_window(); oracle/nsgcq.py has the same switch.
__device__ __forceinline__ float band_window(int window, int kp, int half, int lg) {
    if (window == 1) return 0.5f - 0.5f * cospif(2.0f * (float)kp / (float)(lg - 1));
    return 0.5f + 0.5f * cospif(2.0f * (float)(kp - half) / (float)lg);
}

// ------------------------------------------------------------------------------------------------ CZT column pass
// MODE 0: band input a[k'] = X[first_bin + k'] * hann * chirp from the packed half spectrum (untangled on the fly).
// MODE 1: chirp filter b[n] = e^{-i pi 3 n^2 / M}, n = k' (k' < F) or k' - L (k' >= F), for plan creation.
// MODE 2: as MODE 0 but X[k] is read as is (the Bluestein path already produced the full-spectrum bins).
// Then 16-point FFT down the columns (k' = L2*a + b), twiddle W_L^{b c}, store work[c*L2 + b].
template <int MODE, int L1>
__global__ void __launch_bounds__(CQ_THREADS)
czt_cols_kernel(const BandMeta *__restrict__ bands, const float2 *__restrict__ z_lo, const float2 *__restrict__ z_hi,
                int klo, int khi, const float2 *__restrict__ twN_hi, const float2 *__restrict__ twN_lo,
                const float2 *__restrict__ chirp, int M, int F, float2 *__restrict__ work, int window,
                unsigned int *__restrict__ pmax) {
    // the track maximum czt_out_kernel accumulates (stream order: after this kernel) starts at 0: saves a memset launch
    if (pmax != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *pmax = 0u;
    const BandMeta bm = bands[blockIdx.y];
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= bm.L2) return;
    const unsigned long long twoM = 2ull * (unsigned long long)M;
    float2 v[L1];
#pragma unroll
    for (int a = 0; a < L1; ++a) {
        const int kp = bm.L2 * a + b;
        float2 val = make_float2(0.f, 0.f);
        if (MODE == 2) {          // Bluestein plans: z_lo holds X[klo ..] itself
            if (kp < bm.lg) {
                const float2 X = z_lo[bm.first_bin + kp - klo];
                const float hann = band_window(window, kp, bm.half, bm.lg);
                val = cmul(make_float2(X.x * hann, X.y * hann), chirp[kp]);
            }
        } else if (MODE == 0) {
            if (kp < bm.lg) {
                const int k = bm.first_bin + kp;
                const float2 zk = z_lo[k - klo];
                const float2 zm = cconj(z_hi[khi - k]);             // conj(Z[H - k])
                const float2 w = twiddle2(twN_hi, twN_lo, k, -1);   // e^{-2 pi i k / N}
                const float2 se = cadd(zk, zm), so = cmul(w, csub(zk, zm));
                // X[k] = (Zk + conj Zm)/2 - (i/2) W (Zk - conj Zm)
                const float2 X = make_float2(0.5f * (se.x + so.y), 0.5f * (se.y - so.x));
                const float hann = band_window(window, kp, bm.half, bm.lg);
                val = cmul(make_float2(X.x * hann, X.y * hann), chirp[kp]);     // chirp[kp] = e^{+i pi 3 kp^2 / M}
            }
        } else {
            const long long n = kp < F ? kp : (long long)kp - bm.L;
            const unsigned long long ph = (3ull * (unsigned long long)(n * n)) % twoM;
            float s, c;
            sincospif((float)((double)ph / (double)M), &s, &c);
            val = make_float2(c, -s);
        }
        v[a] = val;
    }
    dft_col<L1>(v, -1);
    float2 *dst = work + bm.work_off;
    {   // inter-pass twiddle W_L^{b c}, c = 0..L1-1: one sincospif, powers by product tree
        float s, co;
        sincospif(-2.0f * (float)b / (float)bm.L, &s, &co);
        apply_powers<L1>(v, make_float2(co, s));
    }
#pragma unroll
    for (int c = 0; c < L1; ++c) dst[c * bm.L2 + b] = v[c];
}

// ------------------------------------------------------------------------------------------------ CZT row pass
// One CTA per (row c, band): FFT_L2 along the row (spectrum index k = c + 16 d at position d), multiply by the chirp-filter
// spectrum, inverse FFT_L2 (-> b'), inverse inter-pass twiddle e^{+2 pi i b' c / L}. MODE 1 (plan creation) stops after
// the forward FFT and stores the spectrum.
template <int MODE>
__global__ void __launch_bounds__(CQ_FFT_THREADS, 2)
czt_rows_kernel(const BandMeta *__restrict__ bands, const RowTile *__restrict__ tiles, float2 *__restrict__ work,
                const float2 *const *__restrict__ btabs, const FftDesc *__restrict__ descs,
                const float2 *const *__restrict__ tws) {
    extern __shared__ __align__(16) float2 fsm[];
    const RowTile tl = tiles[blockIdx.x];
    const BandMeta bm = bands[tl.band];
    __shared__ FftDesc d;
    if (threadIdx.x == 0) d = descs[bm.btab];
    const float2 *tw_g = tws[bm.btab];
    const int L2 = bm.L2, G = tl.g, tot = G * L2;
    float2 *rows = work + bm.work_off + (long long)tl.c0 * L2;
    float2 *tw = fsm, *A = fsm + L2, *B = A + CQ_PAD(tot) + 1;
#pragma unroll 4
    for (int i = threadIdx.x; i < L2; i += blockDim.x) tw[i] = tw_g[i];
#pragma unroll 4
    for (int i = threadIdx.x; i < tot; i += blockDim.x) A[CQ_PAD(i)] = rows[i];
    __syncthreads();
    float2 *R = smem_fft(A, B, d, G, tw, -1);
    if (MODE == 1) {
        for (int i = threadIdx.x; i < tot; i += blockDim.x) rows[i] = R[CQ_PAD(i)];
        return;
    }
    const float2 *bt = btabs[bm.btab] + (long long)tl.c0 * L2;
#pragma unroll 4
    for (int i = threadIdx.x; i < tot; i += blockDim.x) R[CQ_PAD(i)] = cmul(R[CQ_PAD(i)], bt[i]);
    __syncthreads();
    float2 *O = (R == A) ? B : A;
    float2 *R2 = smem_fft(R, O, d, G, tw, +1);
    const float invL = 1.0f / (float)bm.L;
    for (int g = 0; g < G; ++g) {
        const int c = tl.c0 + g;
#pragma unroll 4
        for (int i = threadIdx.x; i < L2; i += blockDim.x) {
            float s, co;
            sincospif(2.0f * (float)(i * c) * invL, &s, &co);      // i * c < L <= 2^17: exact in float
            rows[g * L2 + i] = cmul(R2[CQ_PAD(g * L2 + i)], make_float2(co, s));
        }
    }
}

// ------------------------------------------------------------------------------------------------ CZT row pass, L2 = 256 r0
// The row transform for L2 = r0 * 16 * 16 (r0 in {2..10,12,14,15,16}) with every stage in registers:
//   forward  Stockham (r0, 16, 16): stage 1 reads global memory, stage 3 leaves X[k + 16 r0 u] in the thread that, in the
//   inverse  Stockham (16, 16, r0), owns exactly those inputs — the chirp-filter multiply happens in registers and the
//   last stage stores to global memory with the inter-pass twiddle applied. Four shared-memory exchanges per element in
//   all (ping-pong buffers, one barrier each) against eleven in the generic kernel. Per butterfly ONE twiddle W_N^m is
//   loaded (L1-resident table) and its powers are formed by a depth-4 product tree on the otherwise idle FMA pipe.
template <int R>
__device__ __forceinline__ void rows3_first(const float2 *__restrict__ rows, float2 *S, int G, int N) {
    const int j = threadIdx.x;
    for (int g = 0; g < G; ++g) {
        float2 v[R];
#pragma unroll
        for (int u = 0; u < R; ++u) v[u] = rows[g * N + j + 256 * u];
        dft_r<R>(v, -1);
#pragma unroll
        for (int u = 0; u < R; ++u) S[cr_pad(g * N + j * R + u)] = v[u];
    }
}

template <int R>
__device__ __forceinline__ void rows3_last(float2 *__restrict__ rows, const float2 *S, const float2 *__restrict__ T,
                                           int G, int N, int c0, float invL, float inv16r) {
    const int j = threadIdx.x;
    const float2 wt = cconj(T[j]);
    for (int g = 0; g < G; ++g) {
        float2 v[R];
#pragma unroll
        for (int u = 0; u < R; ++u) v[u] = S[cr_pad(g * N + j + 256 * u)];
        apply_powers<R>(v, wt);
        dft_r<R>(v, +1);
        // inter-pass twiddle e^{+2 pi i b' c / L}, b' = j + 256 u:  wa * wb^u
        const int c = c0 + g;
        float sa, ca, sb, cb;
        sincospif(2.0f * (float)(j * c) * invL, &sa, &ca);
        sincospif(2.0f * (float)c * inv16r, &sb, &cb);
        apply_powers<R>(v, make_float2(cb, sb));
        const float2 wa = make_float2(ca, sa);
#pragma unroll
        for (int u = 0; u < R; ++u) rows[g * N + j + 256 * u] = cmul(v[u], wa);
    }
}

__global__ void __launch_bounds__(256, CQ_ROWS3_CTAS)
czt_rows3_kernel(const BandMeta *__restrict__ bands, const RowTile *__restrict__ tiles, float2 *__restrict__ work,
                 const float2 *const *__restrict__ btabs, const float2 *const *__restrict__ ttabs) {
    extern __shared__ __align__(16) float2 fsm[];
    const RowTile tl = tiles[blockIdx.x];
    const BandMeta bm = bands[tl.band];
    const int r0 = bm.r0, N = bm.L2, G = tl.g, m16 = 16 * r0;     // N = 256 r0; m16 = butterflies per row of a radix-16 stage
    float2 *rows = work + bm.work_off + (long long)tl.c0 * N;
    const float2 *__restrict__ T = ttabs[bm.btab];                  // T[m] = e^{-2 pi i m / N}
    float2 *S0 = fsm, *S1 = fsm + cr_pad(CQ_ROW_POINTS) + 16;

    switch (r0) {
        case 2: rows3_first<2>(rows, S0, G, N); break;
        case 3: rows3_first<3>(rows, S0, G, N); break;
        case 4: rows3_first<4>(rows, S0, G, N); break;
        case 5: rows3_first<5>(rows, S0, G, N); break;
        case 6: rows3_first<6>(rows, S0, G, N); break;
        case 7: rows3_first<7>(rows, S0, G, N); break;
        case 8: rows3_first<8>(rows, S0, G, N); break;
        case 9: rows3_first<9>(rows, S0, G, N); break;
        case 10: rows3_first<10>(rows, S0, G, N); break;
        case 12: rows3_first<12>(rows, S0, G, N); break;
        case 14: rows3_first<14>(rows, S0, G, N); break;
        case 15: rows3_first<15>(rows, S0, G, N); break;
        default: rows3_first<16>(rows, S0, G, N); break;
    }
    __syncthreads();

    // radix-16 stages: thread -> (row g, butterfly j2 < 16 r0); G * 16 r0 <= 256
    const int idx = threadIdx.x;
    const bool active = idx < G * m16;
    const int g = __float2int_rz(((float)idx + 0.5f) / (float)m16);
    const int j2 = idx - g * m16;
    const int base = g * N;
    float2 v[16];
    if (active) {   // forward stage 2: Ns = r0
        const int blk = __float2int_rz(((float)j2 + 0.5f) / (float)r0), k = j2 - blk * r0;
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = S0[cr_pad(base + j2 + u * m16)];
        apply_powers<16>(v, T[16 * k]);
        dft16(v, -1);
#pragma unroll
        for (int u = 0; u < 16; ++u) S1[cr_pad(base + blk * m16 + k + u * r0)] = v[u];
    }
    __syncthreads();
    if (active) {   // forward stage 3: Ns = 16 r0; then the chirp-filter multiply and inverse stage 1 (Ns = 1), all in registers
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = S1[cr_pad(base + j2 + u * m16)];
        apply_powers<16>(v, T[j2]);
        dft16(v, -1);
        const float2 *bt = btabs[bm.btab] + (long long)(tl.c0 + g) * N + j2;
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = cmul(v[u], bt[u * m16]);
        dft16(v, +1);
#pragma unroll
        for (int u = 0; u < 16; ++u) S0[cr_pad(base + j2 * 16 + u)] = v[u];
    }
    __syncthreads();
    if (active) {   // inverse stage 2: Ns = 16
        const int blk = j2 >> 4, k = j2 & 15;
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = S0[cr_pad(base + j2 + u * m16)];
        apply_powers<16>(v, cconj(T[r0 * k]));
        dft16(v, +1);
#pragma unroll
        for (int u = 0; u < 16; ++u) S1[cr_pad(base + blk * 256 + k + 16 * u)] = v[u];
    }
    __syncthreads();
    const float invL = 1.0f / (float)bm.L, inv16r = 256.0f / (float)bm.L;    // 256 / L = 1 / (L1 r0), L1 = column length
    switch (r0) {
        case 2: rows3_last<2>(rows, S1, T, G, N, tl.c0, invL, inv16r); break;
        case 3: rows3_last<3>(rows, S1, T, G, N, tl.c0, invL, inv16r); break;
        case 4: rows3_last<4>(rows, S1, T, G, N, tl.c0, invL, inv16r); break;
        case 5: rows3_last<5>(rows, S1, T, G, N, tl.c0, invL, inv16r); break;
        case 6: rows3_last<6>(rows, S1, T, G, N, tl.c0, invL, inv16r); break;
        case 7: rows3_last<7>(rows, S1, T, G, N, tl.c0, invL, inv16r); break;
        case 8: rows3_last<8>(rows, S1, T, G, N, tl.c0, invL, inv16r); break;
        case 9: rows3_last<9>(rows, S1, T, G, N, tl.c0, invL, inv16r); break;
        case 10: rows3_last<10>(rows, S1, T, G, N, tl.c0, invL, inv16r); break;
        case 12: rows3_last<12>(rows, S1, T, G, N, tl.c0, invL, inv16r); break;
        case 14: rows3_last<14>(rows, S1, T, G, N, tl.c0, invL, inv16r); break;
        case 15: rows3_last<15>(rows, S1, T, G, N, tl.c0, invL, inv16r); break;
        default: rows3_last<16>(rows, S1, T, G, N, tl.c0, invL, inv16r); break;
    }
}

// ------------------------------------------------------------------------------------------------ CZT output pass
// 16-point inverse FFT up the columns (-> a'), output index i = L2*a' + b'; power[band][i] = |c|^2 / (L M)^2 for i < F,
// plus the per-track maximum (atomicMax on the bit pattern of a non-negative float).
template <int L1>
__global__ void __launch_bounds__(CQ_THREADS)
czt_out_kernel(const BandMeta *__restrict__ bands, const float2 *__restrict__ work, int M, int F, int fpitch,
               float *__restrict__ power, unsigned int *__restrict__ pmax) {
    const BandMeta bm = bands[blockIdx.y];
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    float best = 0.f;
    if (b < bm.L2 && b < F) {
        const float2 *src = work + bm.work_off;
        float2 v[L1];
#pragma unroll
        for (int c = 0; c < L1; ++c) v[c] = src[c * bm.L2 + b];
        dft_col<L1>(v, +1);
        const float scale = 1.0f / ((float)bm.L * (float)M);
        float *dst = power + (long long)blockIdx.y * fpitch;
#pragma unroll
        for (int a = 0; a < L1; ++a) {
            const int i = bm.L2 * a + b;
            if (i < F) {
                const float re = v[a].x * scale, im = v[a].y * scale;
                const float p = fmaf(re, re, im * im);
                dst[i] = p;
                best = fmaxf(best, p);
            }
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) best = fmaxf(best, __shfl_xor_sync(0xFFFFFFFFu, best, s));
    if ((threadIdx.x & 31) == 0 && best > 0.f) atomicMax(pmax, __float_as_uint(best));
}

// ------------------------------------------------------------------------------------------------ dB + transpose
// power[band][i] (band-major, pitch fpitch) -> out[col][band] (the reference's column-major 121 x cols).
// MODE 0: amplitude_to_db (convert.h:7-25): 10 log10(max(p,1e-10)) - 10 log10(max(1e-10, pmax)), floored at -80 below the
// maximum (which is 0 dB by construction). MODE 1: linear magnitude sqrt(p). Columns >= F (the one the reference leaves
// unwritten when M % 3 == 0, cqt.h:73-81) are amplitude 0.
template <int MODE>
__global__ void __launch_bounds__(CQ_THREADS)
db_kernel(const float *__restrict__ power, const unsigned int *__restrict__ pmax, int F, int cols, int fpitch,
          float *__restrict__ out) {
    __shared__ float tile[32][CQ_BINS + 2];
    const int col0 = blockIdx.x * 32;
    for (int idx = threadIdx.x; idx < 32 * CQ_BINS; idx += blockDim.x) {
        const int band = idx / 32, cc = idx - band * 32;
        const int col = col0 + cc;
        tile[cc][band] = (col < F) ? power[(long long)band * fpitch + col] : 0.f;
    }
    __syncthreads();
    const float mx = fmaxf(1e-10f, __uint_as_float(*pmax));
    const float ref = __fmul_rn(10.0f, log10f(mx));   // _rn: no FMA contraction, so the maximum maps to exactly 0 dB
    for (int idx = threadIdx.x; idx < 32 * CQ_BINS; idx += blockDim.x) {
        const int cc = idx / CQ_BINS, band = idx - cc * CQ_BINS;
        const int col = col0 + cc;
        if (col < cols) {
            const float p = tile[cc][band];
            float r;
            if (MODE == 0) r = fmaxf(__fsub_rn(__fmul_rn(10.0f, log10f(fmaxf(p, 1e-10f))), ref), -80.0f);
            else r = sqrtf(p);
            out[(long long)col * CQ_BINS + band] = r;
        }
    }
}

// ------------------------------------------------------------------------------------------------ plan tables
// Fills every twiddle / chirp table of a plan in one launch (blockIdx.y = job): double-precision sincospi on the device,
// rounded once to float.
enum { TAB_UNIT = 0, TAB_HI = 1, TAB_LO = 2, TAB_CHIRP = 3, TAB_STAGE = 4 };
struct TabJob {
    int kind, count;
    long long off;      // float2 offset from the arena base
    long long p0;       // period (UNIT/HI/LO), M (CHIRP), index into the descriptor array (STAGE)
};
__device__ __forceinline__ float2 unit_phasor(double frac2) {      // e^{i pi frac2}
    double s, c;
    sincospi(frac2, &s, &c);
    return make_float2((float)c, (float)s);
}
__global__ void __launch_bounds__(CQ_THREADS)
table_fill_kernel(const TabJob *__restrict__ jobs, const FftDesc *__restrict__ descs, float2 *__restrict__ arena) {
    const TabJob jb = jobs[blockIdx.y];
    float2 *dst = arena + jb.off;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < jb.count; i += gridDim.x * blockDim.x) {
        float2 v = make_float2(1.f, 0.f);
        switch (jb.kind) {
            case TAB_UNIT: v = unit_phasor(-2.0 * (double)i / (double)jb.p0); break;
            case TAB_HI: v = unit_phasor(-2.0 * (double)(((long long)i * CQ_TW_S) % jb.p0) / (double)jb.p0); break;
            case TAB_LO: v = unit_phasor(-2.0 * (double)i / (double)jb.p0); break;
            case TAB_CHIRP: {   // e^{+i pi 3 k^2 / M}, phase reduced mod 2M in integers
                const unsigned long long ph = (3ull * (unsigned long long)i * (unsigned long long)i) % (2ull * (unsigned long long)jb.p0);
                v = unit_phasor((double)ph / (double)jb.p0);
            } break;
            default: {          // TAB_STAGE: stage twiddles of the generic shared-memory FFT, [(t-1)*Ns + k] = e^{-2 pi i k t/(Ns R)}
                const FftDesc &d = descs[jb.p0];
                int Ns = 1;
                for (int s = 0; s < d.nrad; ++s) {
                    const int R = d.rad[s];
                    if (Ns > 1 && i >= d.tw_off[s] && i < d.tw_off[s] + (R - 1) * Ns) {
                        const int e = i - d.tw_off[s], t = e / Ns + 1, k = e - (t - 1) * Ns;
                        v = unit_phasor(-2.0 * (double)((long long)k * t) / (double)((long long)Ns * R));
                    }
                    Ns *= R;
                }
            } break;
        }
        dst[i] = v;
    }
}

// ================================================================================================ host side: design + plan
struct CqtDesign {
    int pos[CQ_BINS], lg[CQ_BINS];
    int M = 0, F = 0, cols = 0;
};

// oracle/nsgcq.py nsg_design, same double arithmetic (libm pow / floor)
static bool cqt_design(int64_t n_samples, CqtDesign &d) {
    if (n_samples < 2) return false;
    const double q = std::pow(2.0, 1.0 / CQ_BPO) - std::pow(2.0, -1.0 / CQ_BPO);
    const double fftres = CQ_SR / (double)n_samples;
    for (int j = 0; j < CQ_BINS; ++j) {
        const double f = CQ_FMIN * std::pow(2.0, (double)j / CQ_BPO);
        const double bw = q * f;
        d.pos[j] = (int)std::floor(f / fftres);
        const long long r = (long long)std::floor(bw / fftres + 0.5);
        d.lg[j] = (int)std::max<long long>(r, CQ_MINWIN);
    }
    d.M = d.lg[CQ_BINS - 1];
    d.F = (d.M + CQ_DOWN - 1) / CQ_DOWN;
    d.cols = d.M / CQ_DOWN + 1;
    return true;
}

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// Radix sequence of a {2,3,5,7}-smooth n: the fewest Stockham stages over the butterflies this file has (every stage is
// one shared-memory round trip + barrier), ties broken towards the smaller radix sum; stages run in ascending radix order
// (the first stage, Ns = 1, is twiddle-free and its strided stores conflict least for a small radix).
// The best radix list depends only on the exponents of 2, 3, 5, 7: memoised on those (a few thousand states at most), so
// that planning a new audio length costs map look-ups, not a search.
struct RadixPlan {
    int cost = 1 << 30;      // stages * 1000 + radix sum
    int nrad = 0;
    int rad[CQ_MAXRAD] = {};
};
static const RadixPlan &radix_plan(int e2, int e3, int e5, int e7, int maxrad) {
    static std::mutex mu;
    static std::map<long long, RadixPlan> memo;
    static const int all[13] = {16, 15, 14, 12, 10, 9, 8, 7, 6, 5, 4, 3, 2};
    static const int ex[13][4] = {{4, 0, 0, 0}, {0, 1, 1, 0}, {1, 0, 0, 1}, {2, 1, 0, 0}, {1, 0, 1, 0}, {0, 2, 0, 0}, {3, 0, 0, 0},
                                  {0, 0, 0, 1}, {1, 1, 0, 0}, {0, 0, 1, 0}, {2, 0, 0, 0}, {0, 1, 0, 0}, {1, 0, 0, 0}};
    std::lock_guard<std::mutex> lk(mu);
    std::function<const RadixPlan &(int, int, int, int)> go = [&](int a, int b, int c, int d) -> const RadixPlan & {
        const long long key = ((((long long)maxrad * 64 + a) * 64 + b) * 64 + c) * 64 + d;
        auto it = memo.find(key);
        if (it != memo.end()) return it->second;
        RadixPlan best;
        if (a == 0 && b == 0 && c == 0 && d == 0) {
            best.cost = 0;
        } else {
            for (int i = 0; i < 13; ++i) {
                if (all[i] > maxrad || ex[i][0] > a || ex[i][1] > b || ex[i][2] > c || ex[i][3] > d) continue;
                const RadixPlan &sub = go(a - ex[i][0], b - ex[i][1], c - ex[i][2], d - ex[i][3]);
                if (sub.cost >= (1 << 29) || sub.nrad >= CQ_MAXRAD) continue;
                const int cst = sub.cost + 1000 + all[i];
                if (cst < best.cost) {
                    best = sub;
                    best.cost = cst;
                    best.rad[best.nrad++] = all[i];
                }
            }
        }
        return memo.emplace(key, best).first->second;
    };
    return go(e2, e3, e5, e7);
}
static bool smooth_exponents(int n, int (&e)[4]) {
    const int p[4] = {2, 3, 5, 7};
    for (int i = 0; i < 4; ++i) {
        e[i] = 0;
        while (n > 1 && n % p[i] == 0) { n /= p[i]; e[i]++; }
    }
    return n == 1;
}
static bool factor_smooth(int n, FftDesc &d) {
    d.n = n;
    d.nrad = 0;
    int e[4];
    if (n < 1 || !smooth_exponents(n, e)) return false;
    const RadixPlan &rp = radix_plan(e[0], e[1], e[2], e[3], env_int("HPFW_CQT_MAXRADIX", 16));
    if (rp.cost >= (1 << 29)) return false;
    for (int i = 0; i < rp.nrad; ++i) {     // ascending radix order (insertion sort of <= 14 entries)
        int j = d.nrad++;
        for (; j > 0 && d.rad[j - 1] > rp.rad[i]; --j) d.rad[j] = d.rad[j - 1];
        d.rad[j] = rp.rad[i];
    }
    int off = 0, Ns = 1;
    auto magic = [](unsigned long long dv) { return ((1ull << 40) + dv - 1) / dv; };
    for (int s = 0; s < d.nrad; ++s) {
        d.tw_off[s] = off;
        d.mg_m[s] = magic((unsigned long long)(d.n / d.rad[s]));
        d.mg_ns[s] = magic((unsigned long long)Ns);
        if (Ns > 1) off += (d.rad[s] - 1) * Ns;
        Ns *= d.rad[s];
    }
    return true;
}

struct TwoLevel {
    DeviceBuffer hi, lo;
    int upload(long long P) {
        const long long nhi = P / CQ_TW_S + 2;
        std::vector<float2> h((size_t)nhi), l((size_t)CQ_TW_S);
        for (long long i = 0; i < nhi; ++i) {
            const double a = -2.0 * M_PI * (double)((i * CQ_TW_S) % P) / (double)P;
            h[(size_t)i] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
        for (int i = 0; i < CQ_TW_S; ++i) {
            const double a = -2.0 * M_PI * (double)i / (double)P;
            l[(size_t)i] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
        HPFW_TRY(hi.reserve(sizeof(float2) * h.size()));
        HPFW_TRY(lo.reserve(sizeof(float2) * l.size()));
        HPFW_CUDA_TRY(cudaMemcpy(hi.ptr, h.data(), sizeof(float2) * h.size(), cudaMemcpyHostToDevice));
        HPFW_CUDA_TRY(cudaMemcpy(lo.ptr, l.data(), sizeof(float2) * l.size(), cudaMemcpyHostToDevice));
        return HPFW_OK;
    }
    void release() { hi.release(); lo.release(); }
};

// ---- plan memory ---------------------------------------------------------------------------------------------------
// A plan = every table and index list one audio length needs, in ONE device allocation (the arena) filled by two kernels
// (table_fill_kernel + the chirp-filter transforms) from ONE pinned metadata block; no host trigonometry, no synchronous
// copy, no cudaMalloc in steady state (arenas and pinned blocks are recycled through free lists when a plan is evicted).
// Music libraries have a different sample count for every track, so planning has to cost about as much as a kernel launch.
struct PlanMem {
    DeviceBuffer dev;
    PinnedBuffer pin;
};

struct CqtPlan {
    int64_t N = 0;
    int H = 0;
    bool bluestein = false;     // N odd or N/2 not {2,3,5,7}-smooth: chirp convolution on length P (d1, d2, tw1, tw2, twH are P's)
    int P = 0;
    CqtDesign des;
    int klo = 0, khi = 0;
    FftDesc d1{}, d2{};
    int G1 = 1, G2 = 1;
    int T1 = CQ_FFT_THREADS, T2 = CQ_FFT_THREADS;   // threads per CTA of pass A / pass B (<= CQ_FFT_THREADS)
    size_t smem1 = 0, smem2 = 0;
    PlanMem mem;
    DeviceBuffer bhat;          // Bluestein: FFT_P of the shifted chirp filter
    // pointers into the arena
    const float2 *tw1 = nullptr, *tw2 = nullptr, *twH_hi = nullptr, *twH_lo = nullptr, *twN_hi = nullptr, *twN_lo = nullptr,
                 *chirp = nullptr;
    const BandMeta *d_bands = nullptr;
    const RowTile *d_tiles = nullptr, *d_tiles3 = nullptr;
    const FftDesc *d_descs = nullptr;
    const float2 *const *d_tw_ptrs = nullptr, *const *d_btab_ptrs = nullptr, *const *d_tt_ptrs = nullptr;
    int n_tiles = 0, n_tiles3 = 0;
    size_t smem_rows = 0;
    int max_L2 = 0;
    int L1 = 16;             // chirp-z column length: 16, 32 beyond 6.7 minutes, 64 beyond 13.5 minutes (L = L1 * L2)
    long long work_elems = 0;
    int fpitch = 0;
    cudaEvent_t ready = nullptr;     // recorded on the creating stream after the tables are filled
    cudaEvent_t last_done[CQ_LANES] = {};   // per lane: recorded after the latest transform that used this plan
    cudaStream_t created_on = nullptr;
    bool settled = false;            // `ready` has been observed complete: no stream needs to wait for it any more
    uint64_t last_use = 0;
    struct CqtGang *gangs = nullptr;     // CQ_GANGS captured launch graphs of this plan (batch entry), created on first use
};

// per-lane scratch, shared by all plans (grow-only): tracks of a batch run concurrently on CQ_LANES streams
struct CqtLane {
    DeviceBuffer zbuf, zlo, zhi, work, power, pmax, bl_a;
    uint64_t gen = 0;           // bumped whenever a buffer above moves: captured graphs that point into it are stale
};

// A gang = CQ_GANG_LANES tracks of one length transformed by ONE cudaGraphLaunch: the six launch stages of every lane are
// captured once per (plan, gang) as parallel branches; a replay only re-points the first node of each branch at the lane's
// audio and the last at its output. The host cost of a track drops from ~10 runtime calls (75-80 us on the bench host, close
// to the 117 us the GPU needs) to a quarter of one graph launch plus two node updates.
constexpr int CQ_GANG_LANES = 4;
constexpr int CQ_GANGS = CQ_LANES / CQ_GANG_LANES;
constexpr int CQ_PASS_ARGS = 15, CQ_DB_ARGS = 6;
struct CqtGang {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaGraphNode_t in_node[CQ_GANG_LANES] = {}, out_node[CQ_GANG_LANES] = {};
    cudaKernelNodeParams in_params[CQ_GANG_LANES] = {}, out_params[CQ_GANG_LANES] = {};
    void *in_args[CQ_GANG_LANES][CQ_PASS_ARGS] = {}, *out_args[CQ_GANG_LANES][CQ_DB_ARGS] = {};
    const float2 *in_ptr[CQ_GANG_LANES] = {};
    float *out_ptr[CQ_GANG_LANES] = {};
    uint64_t lane_gen[CQ_GANG_LANES] = {};
    int window = -1;
    int kernels = 0;            // kernel nodes per replay (for the context's launch counter)
    void destroy() {
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        exec = nullptr;
        graph = nullptr;
    }
};

struct CqtPlanCache {
    std::map<int64_t, std::unique_ptr<CqtPlan>> plans;
    std::vector<PlanMem> free_mem;
    std::vector<DeviceBuffer> free_big;
    std::vector<cudaEvent_t> free_events;
    CqtLane lanes[CQ_LANES];
    cudaEvent_t cap_fork = nullptr, cap_join[CQ_LANES] = {};   // fork / join events used only inside gang captures
    bool smem_set = false;
    uint64_t tick = 0;
};

static void plan_release(CqtPlanCache *c, CqtPlan &p, bool recycle) {
    if (p.gangs) {
        for (int g = 0; g < CQ_GANGS; ++g) p.gangs[g].destroy();
        delete[] p.gangs;
        p.gangs = nullptr;
    }
    if (recycle) {
        if (p.mem.dev.ptr) c->free_mem.push_back(p.mem);
        if (p.bhat.ptr) c->free_big.push_back(p.bhat);
        if (p.ready) c->free_events.push_back(p.ready);
        for (cudaEvent_t e : p.last_done)
            if (e) c->free_events.push_back(e);
    } else {
        p.mem.dev.release();
        p.mem.pin.release();
        p.bhat.release();
        if (p.ready) cudaEventDestroy(p.ready);
        for (cudaEvent_t e : p.last_done)
            if (e) cudaEventDestroy(e);
    }
    p.mem = PlanMem();
    p.bhat = DeviceBuffer();
    p.ready = nullptr;
    for (cudaEvent_t &e : p.last_done) e = nullptr;
}

void cqt_cache_destroy(CqtPlanCache *c) {
    if (!c) return;
    for (auto &kv : c->plans) plan_release(c, *kv.second, false);
    for (auto &m : c->free_mem) { m.dev.release(); m.pin.release(); }
    for (auto &b : c->free_big) b.release();
    for (auto e : c->free_events) cudaEventDestroy(e);
    if (c->cap_fork) cudaEventDestroy(c->cap_fork);
    for (auto e : c->cap_join)
        if (e) cudaEventDestroy(e);
    for (auto &l : c->lanes) {
        l.zbuf.release(); l.zlo.release(); l.zhi.release(); l.work.release(); l.power.release(); l.pmax.release();
        l.bl_a.release();
    }
    delete c;
}

// A recycled buffer for a new plan: the smallest that is large enough, else the largest (reserve() regrows it), else none.
static void take_mem(std::vector<PlanMem> &pool, size_t bytes, PlanMem &out) {
    int best = -1;
    for (size_t i = 0; i < pool.size(); ++i) {
        if (best < 0) { best = (int)i; continue; }
        const size_t cb = pool[(size_t)best].dev.cap, ci = pool[i].dev.cap;
        if (cb >= bytes ? (ci >= bytes && ci < cb) : ci > cb) best = (int)i;
    }
    if (best < 0) return;
    out = pool[(size_t)best];
    pool.erase(pool.begin() + best);
}
static void take_big(std::vector<DeviceBuffer> &pool, size_t bytes, DeviceBuffer &out) {
    int best = -1;
    for (size_t i = 0; i < pool.size(); ++i) {
        if (best < 0) { best = (int)i; continue; }
        const size_t cb = pool[(size_t)best].cap, ci = pool[i].cap;
        if (cb >= bytes ? (ci >= bytes && ci < cb) : ci > cb) best = (int)i;
    }
    if (best < 0) return;
    out = pool[(size_t)best];
    pool.erase(pool.begin() + best);
}

// Grow-only; a buffer that has to grow takes 50 % headroom so that a library of mixed track lengths stops reallocating
// after its first few tracks (cudaFree / cudaMalloc of tens of MB stall the device for milliseconds).
static int reserve_roomy(DeviceBuffer &b, size_t bytes) { return bytes <= b.cap ? HPFW_OK : b.reserve(bytes + bytes / 2); }
static int lane_reserve(CqtPlanCache *c, const CqtPlan &pl, int lane) {
    CqtLane &sc = c->lanes[lane];
    const size_t nkeep = (size_t)(pl.khi - pl.klo + 1);
    const void *before[7] = {sc.zbuf.ptr, sc.zlo.ptr, sc.zhi.ptr, sc.work.ptr, sc.power.ptr, sc.pmax.ptr, sc.bl_a.ptr};
    struct Bump {
        CqtLane &l;
        const void **b;
        ~Bump() {
            const void *after[7] = {l.zbuf.ptr, l.zlo.ptr, l.zhi.ptr, l.work.ptr, l.power.ptr, l.pmax.ptr, l.bl_a.ptr};
            for (int i = 0; i < 7; ++i)
                if (after[i] != b[i]) { ++l.gen; break; }
        }
    } bump{sc, before};
    HPFW_TRY(reserve_roomy(sc.zbuf, sizeof(float2) * (size_t)(pl.bluestein ? pl.P : pl.H)));
    HPFW_TRY(reserve_roomy(sc.zlo, sizeof(float2) * nkeep));
    if (pl.bluestein) HPFW_TRY(reserve_roomy(sc.bl_a, sizeof(float2) * (size_t)pl.P));
    else HPFW_TRY(reserve_roomy(sc.zhi, sizeof(float2) * nkeep));
    HPFW_TRY(reserve_roomy(sc.work, sizeof(float2) * (size_t)pl.work_elems));
    HPFW_TRY(reserve_roomy(sc.power, sizeof(float) * (size_t)CQ_BINS * pl.fpitch));
    HPFW_TRY(sc.pmax.reserve(sizeof(unsigned int)));
    return HPFW_OK;
}

static size_t czt_row_smem(int L2, int G) { return 8 * ((size_t)L2 + 2 * ((size_t)CQ_PAD(G * L2) + 1)); }
static size_t rows3_smem() { return 2 * 8 * ((size_t)CQ_ROW_POINTS + CQ_ROW_POINTS / 16 + 16); }
static size_t fft_pass_smem(int n, int G) { const size_t t = (size_t)G * n; return 2 * 8 * (t + t / 16 + 16); }
static unsigned long long magic40(unsigned long long d) { return ((1ull << 40) + d - 1) / d; }
// the pass kernels need a first and a last stage: a single-radix transform gets a pass-through radix-1 last stage
static void pad_single_stage(FftDesc &d) {
    if (d.nrad != 1) return;
    d.rad[1] = 1;
    d.tw_off[1] = 0;
    d.mg_m[1] = magic40((unsigned long long)d.n);
    d.mg_ns[1] = magic40((unsigned long long)d.n);
    d.nrad = 2;
}
// T[m] = e^{-2 pi i m / n}, m < n: the one twiddle table of a register-FFT of length n
static std::vector<float2> unit_table(int n) {
    std::vector<float2> t((size_t)n);
    for (int m = 0; m < n; ++m) {
        const double a = -2.0 * M_PI * (double)m / (double)n;
        t[(size_t)m] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    return t;
}

// CZT row length L2 (the chirp-z length is L = 16 * L2 >= need): the {2,3,5,7}-smooth size in [need2, 2 need2) with the
// least row-pass work L2 * (stages + 1.5). Sizes are kept on a coarse grid (128 / 64 / 16) so that a plan has few distinct
// chirp-filter tables and every row starts on a 128-byte boundary. HPFW_CQT_POW2L=1 restores power-of-two lengths.
static int pick_row_len(long long need2) {
    if (need2 < 64) need2 = 64;
    if (env_int("HPFW_CQT_POW2L", 0)) {
        long long p = 64;
        while (p < need2) p <<= 1;
        return (int)p;
    }
    if (need2 > 256 && need2 <= 4096 && !env_int("HPFW_CQT_ROWS_GENERIC", 0)) {   // L2 = 256 r0: czt_rows3_kernel
        const int r0s[13] = {2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 15, 16};
        for (int r0 : r0s)
            if (256 * r0 >= need2) return 256 * r0;
    }
    double best = 1e300;
    int best_n = 0;
    for (long long c = need2; c < 2 * need2 && c <= CQ_MAX_ROW; ++c) {
        const int grid = c >= 1024 ? 128 : (c >= 512 ? 64 : 16);
        if (c % grid) continue;
        FftDesc d{};
        if (!factor_smooth((int)c, d)) continue;
        const double cost = (double)c * (d.nrad + 1.5);
        if (cost < best) { best = cost; best_n = (int)c; }
    }
    return best_n;   // 0: nothing fits under CQ_MAX_ROW
}

// split H = n1 * n2 (n1 <= n2 <= CQ_MAX_ROW), both {2,3,5,7}-smooth: the fewest Stockham stages in total, then a row pitch
// that keeps pass A's column groups sector-aligned, then the most balanced; returns false if there is none
static bool split_smooth_capped(int H, int n1max, int &n1, int &n2) {
    int e[4];
    if (H < 4 || !smooth_exponents(H, e)) return false;
    const int maxrad = env_int("HPFW_CQT_MAXRADIX", 16);
    long long best = -1;
    int best_stages = 1 << 30, best_al = -1;
    long long va = 1;
    for (int a = 0; a <= e[0]; ++a, va *= 2) {
        long long vb = va;
        for (int b = 0; b <= e[1]; ++b, vb *= 3) {
            long long vc = vb;
            for (int c = 0; c <= e[2]; ++c, vc *= 5) {
                long long v = vc;
                for (int d = 0; d <= e[3]; ++d, v *= 7) {
                    const long long w = H / v;
                    if (!(v <= w && w <= CQ_MAX_ROW && v >= 2 && v <= n1max)) continue;
                    const RadixPlan &pa = radix_plan(a, b, c, d, maxrad);
                    const RadixPlan &pb = radix_plan(e[0] - a, e[1] - b, e[2] - c, e[3] - d, maxrad);
                    if (pa.cost >= (1 << 29) || pb.cost >= (1 << 29)) continue;
                    const int st = pa.nrad + pb.nrad;
                    // pass A touches 4 adjacent columns = one 32-byte sector only if the row pitch n2 is a multiple of 4
                    const int al = (w % 4 == 0) ? 1 : 0;
                    if (st < best_stages || (st == best_stages && (al > best_al || (al == best_al && v > best)))) {
                        best_stages = st; best_al = al; best = v;
                    }
                }
            }
        }
    }
    if (best < 2) return false;
    n1 = (int)best;
    n2 = (int)(H / best);
    return true;
}

// Column FFTs longer than ~1000 points push pass A to one CTA per SM (4 adjacent columns = one 32-byte sector is the least
// it should load): prefer a split with n1 <= 1024 when there is one. HPFW_CQT_N1MAX overrides the cap (tuning).
static bool split_smooth(int H, int &n1, int &n2) {
    const int cap = env_int("HPFW_CQT_N1MAX", 0);
    if (cap > 0) return split_smooth_capped(H, cap, n1, n2);
    return split_smooth_capped(H, 1024, n1, n2) || split_smooth_capped(H, 1 << 30, n1, n2);
}

// Shared memory of the two-pass FFT kernels: ping-pong buffers of G * n padded complex values. G (columns / rows per CTA)
// is sized for three resident CTAs per SM (<= 74 KB each) when the transform allows it, one CTA otherwise.
static void fft_group_sizes(hpfw_ctx *ctx, int n1, int n2, int &G1, int &G2, size_t &smem1, size_t &smem2) {
    const size_t multi_cta = (size_t)env_int("HPFW_CQT_SMEM_KB", 74) * 1024, one_cta = (size_t)ctx->max_smem_optin - 2048;
    auto pick = [&](int n, int gmax) {
        int g = gmax;
        while (g > 1 && fft_pass_smem(n, g) > multi_cta) --g;
        if (fft_pass_smem(n, g) > multi_cta) {
            g = gmax;
            while (g > 1 && fft_pass_smem(n, g) > one_cta) --g;
        }
        return g;
    };
    G1 = pick(n1, 8);
    G1 = G1 >= 8 ? 8 : (G1 >= 4 ? 4 : G1);   // whole 32-byte sectors (4 adjacent columns) wherever shared memory allows
    if (G1 < 4 && fft_pass_smem(n1, 4) <= one_cta) G1 = 4;
    G2 = pick(n2, 4);
    // tuning overrides (experiments only)
    const int g1 = env_int("HPFW_CQT_G1", 0), g2 = env_int("HPFW_CQT_G2", 0);
    if (g1 > 0 && fft_pass_smem(n1, g1) <= one_cta) G1 = g1;
    if (g2 > 0 && fft_pass_smem(n2, g2) <= one_cta) G2 = g2;
    smem1 = fft_pass_smem(n1, G1);
    smem2 = fft_pass_smem(n2, G2);
}

// Threads per CTA of a pass kernel. A stage of radix R has G * n / R butterflies, one per thread and loop trip: with 256
// threads a stage of 300 butterflies keeps 59 % of the lanes busy over its two trips. `forced` (HPFW_CQT_T1 / T2) overrides;
// 0 = CQ_FFT_THREADS (the automatic choice below is only taken when HPFW_CQT_TAUTO is set: see profiles/r01y follow-up).
static int pass_threads(const FftDesc &d, int G, int forced) {
    if (forced >= 32 && forced <= CQ_FFT_THREADS) return forced / 32 * 32;
    if (!env_int("HPFW_CQT_TAUTO", 0)) return CQ_FFT_THREADS;
    int best_t = CQ_FFT_THREADS;
    double best_u = 0.0;
    for (int t = CQ_FFT_THREADS; t >= 128; t -= 32) {
        double busy = 0.0, slots = 0.0;
        for (int s = 0; s < d.nrad; ++s) {
            const int tot = G * (d.n / std::max(d.rad[s], 1));
            busy += tot;
            slots += (double)((tot + t - 1) / t) * t;
        }
        const double u = busy / slots * (0.75 + 0.25 * t / CQ_FFT_THREADS);   // fewer warps hide less latency
        if (u > best_u + 1e-9) {
            best_u = u;
            best_t = t;
        }
    }
    return best_t;
}

// every FFT kernel may use up to the device's opt-in shared memory (plans of different sizes share the kernels)
static int set_smem_limits(hpfw_ctx *ctx) {
    const int lim = ctx->max_smem_optin - 2048;   // dynamic + static shared memory must stay within the opt-in limit
    HPFW_CUDA_TRY(cudaFuncSetAttribute(czt_rows_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    HPFW_CUDA_TRY(cudaFuncSetAttribute(czt_rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    HPFW_CUDA_TRY(cudaFuncSetAttribute(fft_pass_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    HPFW_CUDA_TRY(cudaFuncSetAttribute(czt_rows3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rows3_smem()));
    HPFW_CUDA_TRY(cudaFuncSetAttribute(fft_pass_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    return HPFW_OK;
}

// the plan's two-pass FFT (H points for the packed path, P for Bluestein): in -> tmp -> out_lo / out_hi
static int fft_pass_a(hpfw_ctx *ctx, const CqtPlan &pl, const float2 *in, float2 *tmp, int sign, cudaStream_t stream) {
    const int n2 = pl.d2.n, len = pl.bluestein ? pl.P : pl.H;
    KernelScope ks(ctx, HPFW_K_CQT, stream);
    fft_pass_kernel<0><<<(n2 + pl.G1 - 1) / pl.G1, pl.T1, pl.smem1, stream>>>(
        in, tmp, nullptr, pl.d1, n2, pl.G1, magic40((unsigned long long)pl.G1), pl.tw1, pl.twH_hi, pl.twH_lo, 0, 0, len, 1,
        sign);
    return HPFW_OK;
}
static int fft_pass_b(hpfw_ctx *ctx, const CqtPlan &pl, const float2 *tmp, float2 *out_lo, float2 *out_hi, int klo, int khi,
                      int keep_all, int sign, cudaStream_t stream) {
    const int n1 = pl.d1.n, len = pl.bluestein ? pl.P : pl.H;
    KernelScope ks(ctx, HPFW_K_CQT, stream);
    fft_pass_kernel<1><<<(n1 + pl.G2 - 1) / pl.G2, pl.T2, pl.smem2, stream>>>(
        tmp, out_lo, out_hi, pl.d2, n1, pl.G2, magic40((unsigned long long)pl.G2), pl.tw2, nullptr, nullptr, klo, khi, len,
        keep_all, sign);
    return HPFW_OK;
}
static int fft_two_pass(hpfw_ctx *ctx, const CqtPlan &pl, const float2 *in, float2 *tmp, float2 *out_lo, float2 *out_hi,
                        int klo, int khi, int keep_all, int sign, cudaStream_t stream) {
    HPFW_TRY(fft_pass_a(ctx, pl, in, tmp, sign, stream));
    return fft_pass_b(ctx, pl, tmp, out_lo, out_hi, klo, khi, keep_all, sign, stream);
}

// {2,3,5,7}-smooth numbers in [lo, hi], ascending
static std::vector<long long> smooth_in_range(long long lo, long long hi) {
    std::vector<long long> out;
    for (long long a = 1; a <= hi; a *= 2)
        for (long long b = a; b <= hi; b *= 3)
            for (long long c = b; c <= hi; c *= 5)
                for (long long d = c; d <= hi; d *= 7)
                    if (d >= lo) out.push_back(d);
    std::sort(out.begin(), out.end());
    return out;
}

static int event_get(CqtPlanCache *c, cudaEvent_t *e) {
    if (!c->free_events.empty()) {
        *e = c->free_events.back();
        c->free_events.pop_back();
        return HPFW_OK;
    }
    HPFW_CUDA_TRY(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    return HPFW_OK;
}

static int plan_create(hpfw_ctx *ctx, int64_t N, CqtPlan &pl, cudaStream_t stream, int lane) {
    CqtPlanCache *cache = ctx->cqt;
    pl.N = N;
    if (N < 2) HPFW_FAIL(HPFW_ERR_ARG, "CQT: empty audio");
    if (N > (int64_t(1) << 27)) HPFW_FAIL(HPFW_ERR_LIMIT, "CQT: audio length %lld exceeds 2^27 samples", (long long)N);
    if (!cqt_design(N, pl.des)) HPFW_FAIL(HPFW_ERR_ARG, "CQT: bad length");
    const CqtDesign &d = pl.des;
    pl.H = (int)(N / 2);
    pl.klo = 1 << 30;
    pl.khi = 0;
    for (int j = 0; j < CQ_BINS; ++j) {
        const int fb = d.pos[j] - d.lg[j] / 2;
        pl.klo = std::min(pl.klo, fb);
        pl.khi = std::max(pl.khi, fb + d.lg[j] - 1);
    }
    if (pl.klo < 1 || pl.khi >= pl.H)
        HPFW_FAIL(HPFW_ERR_SHORT, "CQT: audio of %lld samples is too short for the 121-band design", (long long)N);
    int n1 = 0, n2 = 0;
    pl.bluestein = (N & 1) || !split_smooth(pl.H, n1, n2) || env_int("HPFW_CQT_BLUESTEIN", 0);
    if (pl.bluestein) {
        // smooth P >= N + K - 1 that the two-pass engine can split: among the 8 smallest, the fewest stages
        const long long need = N + (long long)(pl.khi - pl.klo);
        int found = 0, best_st = 1 << 30;
        for (long long c : smooth_in_range(need, std::min<long long>(1ll << 26, need + need / 4))) {
            int a1 = 0, a2 = 0;
            if (!split_smooth((int)c, a1, a2)) continue;
            FftDesc da{}, db{};
            if (!factor_smooth(a1, da) || !factor_smooth(a2, db)) continue;
            if (da.nrad + db.nrad < best_st) { best_st = da.nrad + db.nrad; pl.P = (int)c; n1 = a1; n2 = a2; }
            if (++found >= 8) break;
        }
        if (!pl.P)
            HPFW_FAIL(HPFW_ERR_LIMIT, "CQT: audio of %lld samples needs a %lld-point chirp convolution; limit 2^26",
                      (long long)N, need);
    }
    if (!factor_smooth(n1, pl.d1) || !factor_smooth(n2, pl.d2)) HPFW_FAIL(HPFW_ERR_LIMIT, "CQT: too many FFT stages");
    fft_group_sizes(ctx, n1, n2, pl.G1, pl.G2, pl.smem1, pl.smem2);
    pad_single_stage(pl.d1);
    pad_single_stage(pl.d2);
    pl.T1 = pass_threads(pl.d1, pl.G1, env_int("HPFW_CQT_T1", 0));
    pl.T2 = pass_threads(pl.d2, pl.G2, env_int("HPFW_CQT_T2", 0));
    if (env_int("HPFW_CQT_DEBUG", 0)) {
        fprintf(stderr, "[hpfw cqt] N=%lld H=%d n1=%d (G=%d, T=%d, radices", (long long)N, pl.H, n1, pl.G1, pl.T1);
        for (int s2 = 0; s2 < pl.d1.nrad; ++s2) fprintf(stderr, " %d", pl.d1.rad[s2]);
        fprintf(stderr, ") n2=%d (G=%d, T=%d, radices", n2, pl.G2, pl.T2);
        for (int s2 = 0; s2 < pl.d2.nrad; ++s2) fprintf(stderr, " %d", pl.d2.rad[s2]);
        fprintf(stderr, ")\n");
    }

    // CZT layout
    std::vector<BandMeta> bands(CQ_BINS);
    std::vector<int> Ls;
    long long off = 0;
    long long need_max = 0;
    for (int j = 0; j < CQ_BINS; ++j) need_max = std::max(need_max, (long long)d.lg[j] + d.F - 1);
    pl.L1 = need_max <= (long long)CQ_L1 * CQ_MAX_ROW ? CQ_L1 : (need_max <= 2ll * CQ_L1 * CQ_MAX_ROW ? 2 * CQ_L1 : 4 * CQ_L1);
    const int L1 = pl.L1;
    for (int j = 0; j < CQ_BINS; ++j) {
        BandMeta &b = bands[j];
        b.lg = d.lg[j];
        b.half = d.lg[j] / 2;
        b.first_bin = d.pos[j] - b.half;
        const long long need = (long long)d.lg[j] + d.F - 1;
        b.L2 = pick_row_len((need + L1 - 1) / L1);
        b.L = b.L2 * L1;
        b.r0 = 0;
        if (b.L2 > 256 && b.L2 <= 4096 && b.L2 % 256 == 0 && !env_int("HPFW_CQT_ROWS_GENERIC", 0)) {
            const int r0 = b.L2 / 256;
            if ((r0 >= 2 && r0 <= 10) || r0 == 12 || r0 == 14 || r0 == 15 || r0 == 16) b.r0 = r0;
        }
        if (b.L2 <= 0)
            HPFW_FAIL(HPFW_ERR_LIMIT, "CQT: audio of %lld samples needs a %lld-point chirp-z transform; limit %d "
                      "(about 27 minutes at 44.1 kHz)", (long long)N, need, 4 * CQ_L1 * CQ_MAX_ROW);
        auto it = std::find(Ls.begin(), Ls.end(), b.L);
        if (it == Ls.end()) { Ls.push_back(b.L); b.btab = (int)Ls.size() - 1; }
        else b.btab = (int)(it - Ls.begin());
        b.work_off = off;
        off += b.L;
        pl.max_L2 = std::max(pl.max_L2, b.L2);
    }
    pl.work_elems = off;
    pl.fpitch = (d.F + 31) & ~31;
    const int nL = (int)Ls.size();

    // row-pass tiles: G consecutive rows of a band per CTA (G * L2 <= CQ_ROW_POINTS; for the register kernel also
    // G * 16 r0 <= 256), so that every CTA of a launch has about the same work and shared-memory footprint whatever the
    // band's chirp length. One tile list per kernel.
    std::vector<RowTile> tiles, tiles3, ftiles;
    pl.smem_rows = 0;
    for (int j = 0; j < CQ_BINS; ++j) {
        const int L2 = bands[j].L2, r0 = bands[j].r0;
        const int G = std::max(1, std::min(L1, r0 ? 16 / r0 : CQ_ROW_POINTS / L2));
        for (int c0 = 0; c0 < L1; c0 += G) {
            const int g = std::min(G, L1 - c0);
            if (r0) {
                tiles3.push_back({j, c0, g});
            } else {
                tiles.push_back({j, c0, g});
                pl.smem_rows = std::max(pl.smem_rows, czt_row_smem(L2, g));
            }
        }
    }
    pl.n_tiles = (int)tiles.size();
    pl.n_tiles3 = (int)tiles3.size();

    // ---- arena layout: tables (float2 units) first, then the metadata block (bytes, 16-byte aligned pieces)
    const long long lenP = pl.bluestein ? pl.P : pl.H;
    long long fo = 0;
    auto take = [&](long long n) { const long long o = fo; fo += (n + 1) & ~1ll; return o; };   // keep 16-byte alignment
    std::vector<TabJob> jobs;
    const long long o_tw1 = take(n1), o_tw2 = take(n2);
    jobs.push_back({TAB_UNIT, n1, o_tw1, n1});
    jobs.push_back({TAB_UNIT, n2, o_tw2, n2});
    const long long nhiP = lenP / CQ_TW_S + 2, o_hhi = take(nhiP), o_hlo = take(CQ_TW_S);
    jobs.push_back({TAB_HI, (int)nhiP, o_hhi, lenP});
    jobs.push_back({TAB_LO, CQ_TW_S, o_hlo, lenP});
    long long o_nhi = 0, o_nlo = 0;
    if (!pl.bluestein) {
        const long long nhiN = N / CQ_TW_S + 2;
        o_nhi = take(nhiN);
        o_nlo = take(CQ_TW_S);
        jobs.push_back({TAB_HI, (int)nhiN, o_nhi, (long long)N});
        jobs.push_back({TAB_LO, CQ_TW_S, o_nlo, (long long)N});
    }
    const long long o_chirp = take(d.M);
    jobs.push_back({TAB_CHIRP, d.M, o_chirp, d.M});
    std::vector<FftDesc> descs((size_t)nL);
    std::vector<long long> o_rowtw((size_t)nL), o_tt((size_t)nL), o_bt((size_t)nL);
    std::vector<BandMeta> fb((size_t)nL);
    int max_fL2 = 0;
    for (int i = 0; i < nL; ++i) {
        const int L2 = Ls[(size_t)i] / L1;
        if (!factor_smooth(L2, descs[(size_t)i])) HPFW_FAIL(HPFW_ERR_LIMIT, "CQT: too many FFT stages");
        o_rowtw[(size_t)i] = take(L2);
        o_tt[(size_t)i] = take(L2);
        o_bt[(size_t)i] = take(Ls[(size_t)i]);
        jobs.push_back({TAB_STAGE, L2, o_rowtw[(size_t)i], i});
        jobs.push_back({TAB_UNIT, L2, o_tt[(size_t)i], L2});
        // chirp-filter spectrum: a pseudo-band whose work area IS the table
        fb[(size_t)i] = BandMeta{0, 0, 0, Ls[(size_t)i], L2, i, 0, o_bt[(size_t)i]};
        for (int c = 0; c < L1; ++c) ftiles.push_back({i, c, 1});
        max_fL2 = std::max(max_fL2, L2);
    }
    const size_t table_bytes = sizeof(float2) * (size_t)fo;
    size_t mo = 0;
    auto mtake = [&](size_t bytes) { const size_t o = mo; mo += (bytes + 15) & ~size_t(15); return o; };
    const size_t m_bands = mtake(sizeof(BandMeta) * CQ_BINS), m_fb = mtake(sizeof(BandMeta) * (size_t)nL),
                 m_tiles = mtake(sizeof(RowTile) * std::max<size_t>(1, tiles.size())),
                 m_tiles3 = mtake(sizeof(RowTile) * std::max<size_t>(1, tiles3.size())),
                 m_ftiles = mtake(sizeof(RowTile) * ftiles.size()), m_descs = mtake(sizeof(FftDesc) * (size_t)nL),
                 m_twp = mtake(sizeof(void *) * (size_t)nL), m_btp = mtake(sizeof(void *) * (size_t)nL),
                 m_ttp = mtake(sizeof(void *) * (size_t)nL), m_jobs = mtake(sizeof(TabJob) * jobs.size());
    const size_t meta_bytes = mo, arena_bytes = table_bytes + meta_bytes;

    take_mem(cache->free_mem, arena_bytes, pl.mem);
    HPFW_TRY(reserve_roomy(pl.mem.dev, arena_bytes));
    HPFW_TRY(pl.mem.pin.reserve(meta_bytes));
    char *arena = pl.mem.dev.as<char>();
    float2 *tab = pl.mem.dev.as<float2>();
    char *dmeta = arena + table_bytes, *hmeta = pl.mem.pin.as<char>();
    std::vector<const float2 *> twp((size_t)nL), btp((size_t)nL), ttp((size_t)nL);
    for (int i = 0; i < nL; ++i) {
        twp[(size_t)i] = tab + o_rowtw[(size_t)i];
        btp[(size_t)i] = tab + o_bt[(size_t)i];
        ttp[(size_t)i] = tab + o_tt[(size_t)i];
    }
    memcpy(hmeta + m_bands, bands.data(), sizeof(BandMeta) * CQ_BINS);
    memcpy(hmeta + m_fb, fb.data(), sizeof(BandMeta) * (size_t)nL);
    if (!tiles.empty()) memcpy(hmeta + m_tiles, tiles.data(), sizeof(RowTile) * tiles.size());
    if (!tiles3.empty()) memcpy(hmeta + m_tiles3, tiles3.data(), sizeof(RowTile) * tiles3.size());
    memcpy(hmeta + m_ftiles, ftiles.data(), sizeof(RowTile) * ftiles.size());
    memcpy(hmeta + m_descs, descs.data(), sizeof(FftDesc) * (size_t)nL);
    memcpy(hmeta + m_twp, twp.data(), sizeof(void *) * (size_t)nL);
    memcpy(hmeta + m_btp, btp.data(), sizeof(void *) * (size_t)nL);
    memcpy(hmeta + m_ttp, ttp.data(), sizeof(void *) * (size_t)nL);
    memcpy(hmeta + m_jobs, jobs.data(), sizeof(TabJob) * jobs.size());
    HPFW_CUDA_TRY(cudaMemcpyAsync(dmeta, hmeta, meta_bytes, cudaMemcpyHostToDevice, stream));

    pl.tw1 = tab + o_tw1;
    pl.tw2 = tab + o_tw2;
    pl.twH_hi = tab + o_hhi;
    pl.twH_lo = tab + o_hlo;
    pl.twN_hi = pl.bluestein ? nullptr : tab + o_nhi;
    pl.twN_lo = pl.bluestein ? nullptr : tab + o_nlo;
    pl.chirp = tab + o_chirp;
    pl.d_bands = reinterpret_cast<const BandMeta *>(dmeta + m_bands);
    pl.d_tiles = reinterpret_cast<const RowTile *>(dmeta + m_tiles);
    pl.d_tiles3 = reinterpret_cast<const RowTile *>(dmeta + m_tiles3);
    pl.d_descs = reinterpret_cast<const FftDesc *>(dmeta + m_descs);
    pl.d_tw_ptrs = reinterpret_cast<const float2 *const *>(dmeta + m_twp);
    pl.d_btab_ptrs = reinterpret_cast<const float2 *const *>(dmeta + m_btp);
    pl.d_tt_ptrs = reinterpret_cast<const float2 *const *>(dmeta + m_ttp);

    if (!cache->smem_set) {
        HPFW_TRY(set_smem_limits(ctx));
        cache->smem_set = true;
    }
    {   // every twiddle / chirp table in one launch
        int maxc = 0;
        for (const TabJob &j : jobs) maxc = std::max(maxc, j.count);
        KernelScope ks(ctx, HPFW_K_CQT, stream);
        table_fill_kernel<<<dim3((unsigned)std::min(64, (maxc + CQ_THREADS - 1) / CQ_THREADS), (unsigned)jobs.size()),
                            CQ_THREADS, 0, stream>>>(reinterpret_cast<const TabJob *>(dmeta + m_jobs), pl.d_descs, tab);
    }
    {   // chirp-filter spectra of all distinct lengths: b -> radix-16 columns -> row FFTs, in place in the arena
        const BandMeta *dfb = reinterpret_cast<const BandMeta *>(dmeta + m_fb);
        {
            KernelScope ks(ctx, HPFW_K_CQT, stream);
            const dim3 gf((max_fL2 + CQ_THREADS - 1) / CQ_THREADS, nL);
            if (L1 == 16)
                czt_cols_kernel<1, 16><<<gf, CQ_THREADS, 0, stream>>>(dfb, nullptr, nullptr, 0, 0, nullptr, nullptr, nullptr,
                                                                      d.M, d.F, tab, 0, nullptr);
            else if (L1 == 32)
                czt_cols_kernel<1, 32><<<gf, CQ_THREADS, 0, stream>>>(dfb, nullptr, nullptr, 0, 0, nullptr, nullptr, nullptr,
                                                                      d.M, d.F, tab, 0, nullptr);
            else
                czt_cols_kernel<1, 64><<<gf, CQ_THREADS, 0, stream>>>(dfb, nullptr, nullptr, 0, 0, nullptr, nullptr, nullptr,
                                                                      d.M, d.F, tab, 0, nullptr);
        }
        {
            KernelScope ks(ctx, HPFW_K_CQT, stream);
            czt_rows_kernel<1><<<(unsigned)ftiles.size(), CQ_FFT_THREADS, czt_row_smem(max_fL2, 1), stream>>>(
                dfb, reinterpret_cast<const RowTile *>(dmeta + m_ftiles), tab, pl.d_btab_ptrs, pl.d_descs, pl.d_tw_ptrs);
        }
    }
    if (pl.bluestein) {     // FFT_P of the shifted chirp filter, through the calling lane's scratch (same stream)
        HPFW_TRY(lane_reserve(cache, pl, lane));
        CqtLane &sc = cache->lanes[lane];
        const int P = pl.P, K = pl.khi - pl.klo + 1, gb = (P + CQ_THREADS - 1) / CQ_THREADS;
        take_big(cache->free_big, sizeof(float2) * (size_t)P, pl.bhat);
        HPFW_TRY(reserve_roomy(pl.bhat, sizeof(float2) * (size_t)P));
        {
            KernelScope ks(ctx, HPFW_K_CQT, stream);
            bl_filter_kernel<<<gb, CQ_THREADS, 0, stream>>>((long long)N, P, pl.klo, K, sc.bl_a.as<float2>());
        }
        HPFW_TRY(fft_two_pass(ctx, pl, sc.bl_a.as<float2>(), sc.zbuf.as<float2>(), pl.bhat.as<float2>(), nullptr, 0, 0, 1,
                              -1, stream));
    }
    HPFW_CUDA_TRY(cudaGetLastError());
    // other streams wait on `ready` before they use this plan; eviction waits on the per-lane `last_done` events
    HPFW_TRY(event_get(cache, &pl.ready));
    HPFW_CUDA_TRY(cudaEventRecord(pl.ready, stream));
    pl.created_on = stream;
    return HPFW_OK;
}

constexpr size_t CQ_PLAN_CACHE = 48;        // plans are a few MB of tables each ...
constexpr size_t CQ_PLAN_CACHE_BLUESTEIN = 6;   // ... except Bluestein plans, which also hold a P-point filter spectrum (70 MB at 3 min)

static int plan_get(hpfw_ctx *ctx, int64_t N, CqtPlan **out, cudaStream_t stream, int lane) {
    if (!ctx->cqt) ctx->cqt = new CqtPlanCache();
    CqtPlanCache *c = ctx->cqt;
    auto it = c->plans.find(N);
    if (it == c->plans.end()) {
        // evict the least recently used plan when the cache is full; its memory goes to the free lists. Whether the new length
        // needs the Bluestein path is a property of N alone (odd, or N/2 not smooth).
        int e4[4];
        const bool will_bluestein = (N & 1) || !smooth_exponents((int)(N / 2), e4) || env_int("HPFW_CQT_BLUESTEIN", 0);
        size_t n_bl = 0;
        for (auto &kv : c->plans) n_bl += kv.second->bluestein ? 1 : 0;
        const bool evict_bl = will_bluestein && n_bl >= CQ_PLAN_CACHE_BLUESTEIN;
        if (c->plans.size() >= CQ_PLAN_CACHE || evict_bl) {
            auto victim = c->plans.end();
            for (auto k = c->plans.begin(); k != c->plans.end(); ++k) {
                if (evict_bl && !k->second->bluestein) continue;
                if (victim == c->plans.end() || k->second->last_use < victim->second->last_use) victim = k;
            }
            for (cudaEvent_t e : victim->second->last_done)
                if (e) HPFW_CUDA_TRY(cudaEventSynchronize(e));
            plan_release(c, *victim->second, true);
            c->plans.erase(victim);
        }
        std::unique_ptr<CqtPlan> p(new CqtPlan());
        int st = plan_create(ctx, N, *p, stream, lane);
        if (st != HPFW_OK) {
            cudaStreamSynchronize(stream);
            plan_release(c, *p, false);
            return st;
        }
        it = c->plans.emplace(N, std::move(p)).first;
    }
    it->second->last_use = ++c->tick;
    *out = it->second.get();
    return HPFW_OK;
}

// One transform = six launch stages (main FFT pass A, pass B, chirp-z columns, rows, output, dB). cqt_run issues them back to
// back on one stream; the batch entry can issue stage by stage across its lanes instead (HPFW_CQT_WAVE), so that the lanes
// run the same kernel at the same time.
struct CqtJob {
    CqtPlan *pl = nullptr;
    CqtLane *sc = nullptr;
    const float *d_audio = nullptr;
    int64_t N = 0;
    float *d_out = nullptr;
    int mode = 0;
    cudaStream_t stream = nullptr;
    int lane = 0;
    bool captured = false;      // inside a stream capture: no event bookkeeping (the gang records it after the graph launch)
};
constexpr int CQ_STAGES = 6;

static int cqt_prepare(hpfw_ctx *ctx, CqtJob &jb) {
    HPFW_TRY(plan_get(ctx, jb.N, &jb.pl, jb.stream, jb.lane));
    CqtPlanCache *cache = ctx->cqt;
    HPFW_TRY(lane_reserve(cache, *jb.pl, jb.lane));
    jb.sc = &cache->lanes[jb.lane];
    if (jb.stream != jb.pl->created_on && !jb.pl->settled) {
        // the tables are filled once; as soon as that has been seen complete no later use needs to wait for it
        if (cudaEventQuery(jb.pl->ready) == cudaSuccess) jb.pl->settled = true;
        else {
            cudaGetLastError();     // cudaErrorNotReady is not sticky, but clear it for the checks below
            HPFW_CUDA_TRY(cudaStreamWaitEvent(jb.stream, jb.pl->ready, 0));
        }
    }
    if (!jb.pl->bluestein && (reinterpret_cast<uintptr_t>(jb.d_audio) & 7) != 0)
        HPFW_FAIL(HPFW_ERR_ARG, "CQT: the audio buffer must be 8-byte aligned");
    return HPFW_OK;
}

static int cqt_stage(hpfw_ctx *ctx, const CqtJob &jb, int stage) {
    CqtPlan *pl = jb.pl;
    CqtLane *sc = jb.sc;
    cudaStream_t stream = jb.stream;
    const CqtDesign &d = pl->des;
    const dim3 gcol((pl->max_L2 + CQ_THREADS - 1) / CQ_THREADS, CQ_BINS);
#ifdef HPFW_CQT_ABLATE
    // timing ablation (never in the shipped build): HPFW_CQT_SKIP is a bit mask of stages whose launches are dropped
    if (stage < CQ_STAGES - 1 && ((env_int("HPFW_CQT_SKIP", 0) >> stage) & 1)) return HPFW_OK;
#endif
    switch (stage) {
    case 0:
        if (!pl->bluestein) {
            HPFW_TRY(fft_pass_a(ctx, *pl, reinterpret_cast<const float2 *>(jb.d_audio), sc->zbuf.as<float2>(), -1, stream));
        } else {
            const int P = pl->P, K = pl->khi - pl->klo + 1, gb = (P + CQ_THREADS - 1) / CQ_THREADS;
            {
                KernelScope ks(ctx, HPFW_K_CQT, stream);
                bl_prep_kernel<<<gb, CQ_THREADS, 0, stream>>>(jb.d_audio, (long long)jb.N, P, sc->bl_a.as<float2>());
            }
            HPFW_TRY(fft_two_pass(ctx, *pl, sc->bl_a.as<float2>(), sc->zbuf.as<float2>(), sc->bl_a.as<float2>(), nullptr, 0, 0,
                                  1, -1, stream));
            {
                KernelScope ks(ctx, HPFW_K_CQT, stream);
                bl_mul_kernel<<<gb, CQ_THREADS, 0, stream>>>(sc->bl_a.as<float2>(), pl->bhat.as<float2>(), P);
            }
            HPFW_TRY(fft_two_pass(ctx, *pl, sc->bl_a.as<float2>(), sc->zbuf.as<float2>(), sc->zlo.as<float2>(), nullptr, 0,
                                  K - 1, 0, +1, stream));
            {
                KernelScope ks(ctx, HPFW_K_CQT, stream);
                bl_post_kernel<<<(K + CQ_THREADS - 1) / CQ_THREADS, CQ_THREADS, 0, stream>>>(sc->zlo.as<float2>(),
                                                                                             (long long)jb.N, P, pl->klo, K);
            }
        }
        break;
    case 1:
        if (!pl->bluestein)
            HPFW_TRY(fft_pass_b(ctx, *pl, sc->zbuf.as<float2>(), sc->zlo.as<float2>(), sc->zhi.as<float2>(), pl->klo, pl->khi,
                                0, -1, stream));
        break;
    case 2: {
        KernelScope ks(ctx, HPFW_K_CQT, stream);
        float2 *zlo = sc->zlo.as<float2>(), *zhi = sc->zhi.as<float2>(), *wk = sc->work.as<float2>();
        if (pl->bluestein) {
            if (pl->L1 == 16)
                czt_cols_kernel<2, 16><<<gcol, CQ_THREADS, 0, stream>>>(pl->d_bands, zlo, nullptr, pl->klo, pl->khi, nullptr,
                                                                        nullptr, pl->chirp, d.M, d.F, wk, ctx->cqt_window, sc->pmax.as<unsigned int>());
            else if (pl->L1 == 32)
                czt_cols_kernel<2, 32><<<gcol, CQ_THREADS, 0, stream>>>(pl->d_bands, zlo, nullptr, pl->klo, pl->khi, nullptr,
                                                                        nullptr, pl->chirp, d.M, d.F, wk, ctx->cqt_window, sc->pmax.as<unsigned int>());
            else
                czt_cols_kernel<2, 64><<<gcol, CQ_THREADS, 0, stream>>>(pl->d_bands, zlo, nullptr, pl->klo, pl->khi, nullptr,
                                                                        nullptr, pl->chirp, d.M, d.F, wk, ctx->cqt_window, sc->pmax.as<unsigned int>());
        } else {
            if (pl->L1 == 16)
                czt_cols_kernel<0, 16><<<gcol, CQ_THREADS, 0, stream>>>(pl->d_bands, zlo, zhi, pl->klo, pl->khi, pl->twN_hi,
                                                                        pl->twN_lo, pl->chirp, d.M, d.F, wk, ctx->cqt_window, sc->pmax.as<unsigned int>());
            else if (pl->L1 == 32)
                czt_cols_kernel<0, 32><<<gcol, CQ_THREADS, 0, stream>>>(pl->d_bands, zlo, zhi, pl->klo, pl->khi, pl->twN_hi,
                                                                        pl->twN_lo, pl->chirp, d.M, d.F, wk, ctx->cqt_window, sc->pmax.as<unsigned int>());
            else
                czt_cols_kernel<0, 64><<<gcol, CQ_THREADS, 0, stream>>>(pl->d_bands, zlo, zhi, pl->klo, pl->khi, pl->twN_hi,
                                                                        pl->twN_lo, pl->chirp, d.M, d.F, wk, ctx->cqt_window, sc->pmax.as<unsigned int>());
        }
        break;
    }
    case 3:
        if (pl->n_tiles3 > 0) {
            KernelScope ks(ctx, HPFW_K_CQT, stream);
            czt_rows3_kernel<<<pl->n_tiles3, 256, rows3_smem(), stream>>>(pl->d_bands, pl->d_tiles3, sc->work.as<float2>(),
                                                                          pl->d_btab_ptrs, pl->d_tt_ptrs);
        }
        if (pl->n_tiles > 0) {
            KernelScope ks(ctx, HPFW_K_CQT, stream);
            czt_rows_kernel<0><<<pl->n_tiles, CQ_FFT_THREADS, pl->smem_rows, stream>>>(
                pl->d_bands, pl->d_tiles, sc->work.as<float2>(), pl->d_btab_ptrs, pl->d_descs, pl->d_tw_ptrs);
        }
        break;
    case 4: {
        KernelScope ks(ctx, HPFW_K_CQT, stream);
        if (pl->L1 == 16)
            czt_out_kernel<16><<<gcol, CQ_THREADS, 0, stream>>>(pl->d_bands, sc->work.as<float2>(), d.M, d.F, pl->fpitch,
                                                                sc->power.as<float>(), sc->pmax.as<unsigned int>());
        else if (pl->L1 == 32)
            czt_out_kernel<32><<<gcol, CQ_THREADS, 0, stream>>>(pl->d_bands, sc->work.as<float2>(), d.M, d.F, pl->fpitch,
                                                                sc->power.as<float>(), sc->pmax.as<unsigned int>());
        else
            czt_out_kernel<64><<<gcol, CQ_THREADS, 0, stream>>>(pl->d_bands, sc->work.as<float2>(), d.M, d.F, pl->fpitch,
                                                                sc->power.as<float>(), sc->pmax.as<unsigned int>());
        break;
    }
    default: {
        {
            KernelScope ks(ctx, HPFW_K_CQT, stream);
            const int gb = (d.cols + 31) / 32;
            if (jb.mode == 0)
                db_kernel<0><<<gb, CQ_THREADS, 0, stream>>>(sc->power.as<float>(), sc->pmax.as<unsigned int>(), d.F, d.cols,
                                                             pl->fpitch, jb.d_out);
            else
                db_kernel<1><<<gb, CQ_THREADS, 0, stream>>>(sc->power.as<float>(), sc->pmax.as<unsigned int>(), d.F, d.cols,
                                                             pl->fpitch, jb.d_out);
        }
        if (jb.captured) break;
        HPFW_CUDA_TRY(cudaGetLastError());
        if (!pl->last_done[jb.lane]) HPFW_TRY(event_get(ctx->cqt, &pl->last_done[jb.lane]));
        HPFW_CUDA_TRY(cudaEventRecord(pl->last_done[jb.lane], stream));
        break;
    }
    }
    return HPFW_OK;
}

// The gang's launches as a graph: lane 0's stream is the origin, the other lanes are forked from it and joined back.
static int gang_capture_body(hpfw_ctx *ctx, CqtJob *jobs, int gang) {
    const int l0 = gang * CQ_GANG_LANES;
    cudaStream_t gs = ctx->lane_stream[l0];
    CqtPlanCache *c = ctx->cqt;
    HPFW_CUDA_TRY(cudaEventRecord(c->cap_fork, gs));
    for (int l = 1; l < CQ_GANG_LANES; ++l) HPFW_CUDA_TRY(cudaStreamWaitEvent(ctx->lane_stream[l0 + l], c->cap_fork, 0));
    for (int st = 0; st < CQ_STAGES; ++st)
        for (int l = 0; l < CQ_GANG_LANES; ++l) HPFW_TRY(cqt_stage(ctx, jobs[l], st));
    for (int l = 1; l < CQ_GANG_LANES; ++l) {
        HPFW_CUDA_TRY(cudaEventRecord(c->cap_join[l0 + l], ctx->lane_stream[l0 + l]));
        HPFW_CUDA_TRY(cudaStreamWaitEvent(gs, c->cap_join[l0 + l], 0));
    }
    return HPFW_OK;
}

static int gang_build(hpfw_ctx *ctx, CqtJob *jobs, int gang, CqtGang &gg) {
    const int l0 = gang * CQ_GANG_LANES;
    cudaStream_t gs = ctx->lane_stream[l0];
    gg.destroy();
    CqtPlanCache *c = ctx->cqt;
    if (!c->cap_fork) HPFW_CUDA_TRY(cudaEventCreateWithFlags(&c->cap_fork, cudaEventDisableTiming));
    for (int l = 0; l < CQ_LANES; ++l)
        if (!c->cap_join[l]) HPFW_CUDA_TRY(cudaEventCreateWithFlags(&c->cap_join[l], cudaEventDisableTiming));
    const uint64_t launches_before = ctx->launches;
    HPFW_CUDA_TRY(cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal));
    for (int l = 0; l < CQ_GANG_LANES; ++l) jobs[l].captured = true;
    const int st = gang_capture_body(ctx, jobs, gang);
    for (int l = 0; l < CQ_GANG_LANES; ++l) jobs[l].captured = false;
    const cudaError_t ce = cudaStreamEndCapture(gs, &gg.graph);
    gg.kernels = (int)(ctx->launches - launches_before);
    ctx->launches = launches_before;          // nothing ran yet: the replay counts them
    if (st != HPFW_OK || ce != cudaSuccess) {
        if (gg.graph) cudaGraphDestroy(gg.graph);
        gg.graph = nullptr;
        cudaGetLastError();
        if (st != HPFW_OK) return st;
        HPFW_FAIL(HPFW_ERR_CUDA, "CQT: stream capture failed: %s", cudaGetErrorString(ce));
    }
    size_t nn = 0;
    HPFW_CUDA_TRY(cudaGraphGetNodes(gg.graph, nullptr, &nn));
    std::vector<cudaGraphNode_t> nodes(nn);
    HPFW_CUDA_TRY(cudaGraphGetNodes(gg.graph, nodes.data(), &nn));
    int found_in = 0, found_out = 0;
    for (cudaGraphNode_t nd : nodes) {
        cudaGraphNodeType ty;
        HPFW_CUDA_TRY(cudaGraphNodeGetType(nd, &ty));
        if (ty != cudaGraphNodeTypeKernel) continue;
        cudaKernelNodeParams kp{};
        HPFW_CUDA_TRY(cudaGraphKernelNodeGetParams(nd, &kp));
        if (!kp.kernelParams) continue;
        if (kp.func == (void *)fft_pass_kernel<0>) {
            const float2 *in = *reinterpret_cast<const float2 *const *>(kp.kernelParams[0]);
            for (int l = 0; l < CQ_GANG_LANES; ++l)
                if (in == reinterpret_cast<const float2 *>(jobs[l].d_audio)) {
                    gg.in_node[l] = nd;
                    gg.in_params[l] = kp;
                    for (int a = 0; a < CQ_PASS_ARGS; ++a) gg.in_args[l][a] = kp.kernelParams[a];
                    gg.in_ptr[l] = in;
                    gg.in_args[l][0] = &gg.in_ptr[l];
                    gg.in_params[l].kernelParams = gg.in_args[l];
                    ++found_in;
                }
        } else if (kp.func == (void *)db_kernel<0>) {
            float *out = *reinterpret_cast<float *const *>(kp.kernelParams[CQ_DB_ARGS - 1]);
            for (int l = 0; l < CQ_GANG_LANES; ++l)
                if (out == jobs[l].d_out) {
                    gg.out_node[l] = nd;
                    gg.out_params[l] = kp;
                    for (int a = 0; a < CQ_DB_ARGS; ++a) gg.out_args[l][a] = kp.kernelParams[a];
                    gg.out_ptr[l] = out;
                    gg.out_args[l][CQ_DB_ARGS - 1] = &gg.out_ptr[l];
                    gg.out_params[l].kernelParams = gg.out_args[l];
                    ++found_out;
                }
        }
    }
    if (found_in != CQ_GANG_LANES || found_out != CQ_GANG_LANES) {
        gg.destroy();
        HPFW_FAIL(HPFW_ERR_STATE, "CQT: captured graph has %d input and %d output nodes, expected %d each", found_in,
                  found_out, CQ_GANG_LANES);
    }
    HPFW_CUDA_TRY(cudaGraphInstantiate(&gg.exec, gg.graph, 0));
    for (int l = 0; l < CQ_GANG_LANES; ++l) gg.lane_gen[l] = jobs[l].sc->gen;
    gg.window = ctx->cqt_window;
    return HPFW_OK;
}

// CQ_GANG_LANES prepared jobs of ONE non-Bluestein plan (mode 0) on the lanes of `gang`: one graph launch.
static int cqt_gang_run(hpfw_ctx *ctx, CqtJob *jobs, int gang) {
    CqtPlan *pl = jobs[0].pl;
    const int l0 = gang * CQ_GANG_LANES;
    cudaStream_t gs = ctx->lane_stream[l0];
    if (!pl->gangs) pl->gangs = new CqtGang[CQ_GANGS];
    CqtGang &gg = pl->gangs[gang];
    bool stale = !gg.exec || gg.window != ctx->cqt_window;
    for (int l = 0; l < CQ_GANG_LANES; ++l) stale = stale || gg.lane_gen[l] != jobs[l].sc->gen;
    if (stale) {
        HPFW_TRY(gang_build(ctx, jobs, gang, gg));
    } else {
        for (int l = 0; l < CQ_GANG_LANES; ++l) {
            if (gg.in_ptr[l] != reinterpret_cast<const float2 *>(jobs[l].d_audio)) {
                gg.in_ptr[l] = reinterpret_cast<const float2 *>(jobs[l].d_audio);
                HPFW_CUDA_TRY(cudaGraphExecKernelNodeSetParams(gg.exec, gg.in_node[l], &gg.in_params[l]));
            }
            if (gg.out_ptr[l] != jobs[l].d_out) {
                gg.out_ptr[l] = jobs[l].d_out;
                HPFW_CUDA_TRY(cudaGraphExecKernelNodeSetParams(gg.exec, gg.out_node[l], &gg.out_params[l]));
            }
        }
    }
    HPFW_CUDA_TRY(cudaGraphLaunch(gg.exec, gs));
    ctx->launches += (uint64_t)gg.kernels;
    if (!pl->last_done[l0]) HPFW_TRY(event_get(ctx->cqt, &pl->last_done[l0]));
    HPFW_CUDA_TRY(cudaEventRecord(pl->last_done[l0], gs));
    return HPFW_OK;
}

// mode 0: dB spectrogram; mode 1: linear magnitudes. d_audio must be 8-byte aligned.
static int cqt_run(hpfw_ctx *ctx, const float *d_audio, int64_t N, float *d_out, int mode, cudaStream_t stream,
                   int lane = 0) {
    CqtJob jb;
    jb.d_audio = d_audio; jb.N = N; jb.d_out = d_out; jb.mode = mode; jb.stream = stream; jb.lane = lane;
    HPFW_TRY(cqt_prepare(ctx, jb));
    for (int st = 0; st < CQ_STAGES; ++st) HPFW_TRY(cqt_stage(ctx, jb, st));
    return HPFW_OK;
}

// Diagnostic / unit-test entry: forward or inverse complex FFT of a smooth length through the same two-pass kernels.
static int fft_c2c_device(hpfw_ctx *ctx, const float2 *d_in, float2 *d_out, int n, int inverse, cudaStream_t stream) {
    int n1 = 0, n2 = 0;
    FftDesc d1{}, d2{};
    if (!split_smooth(n, n1, n2) || !factor_smooth(n1, d1) || !factor_smooth(n2, d2))
        HPFW_FAIL(HPFW_ERR_LIMIT, "hpfw_fft_c2c: %d is not a product n1*n2 of {2,3,5,7}-smooth factors <= %d", n,
                  CQ_MAX_ROW);
    int G1 = 1, G2 = 1;
    size_t smem1 = 0, smem2 = 0;
    fft_group_sizes(ctx, n1, n2, G1, G2, smem1, smem2);
    DeviceBuffer tw1, tw2, tmp;
    TwoLevel twP;
    pad_single_stage(d1);
    pad_single_stage(d2);
    auto t1 = unit_table(n1), t2 = unit_table(n2);
    HPFW_TRY(tw1.reserve(sizeof(float2) * t1.size()));
    HPFW_TRY(tw2.reserve(sizeof(float2) * t2.size()));
    HPFW_TRY(tmp.reserve(sizeof(float2) * (size_t)n));
    HPFW_CUDA_TRY(cudaMemcpy(tw1.ptr, t1.data(), sizeof(float2) * t1.size(), cudaMemcpyHostToDevice));
    HPFW_CUDA_TRY(cudaMemcpy(tw2.ptr, t2.data(), sizeof(float2) * t2.size(), cudaMemcpyHostToDevice));
    HPFW_TRY(twP.upload(n));
    HPFW_TRY(set_smem_limits(ctx));
    const int sign = inverse ? +1 : -1;
    {
        KernelScope ks(ctx, HPFW_K_CQT, stream);
        fft_pass_kernel<0><<<(n2 + G1 - 1) / G1, CQ_FFT_THREADS, smem1, stream>>>(
            d_in, tmp.as<float2>(), nullptr, d1, n2, G1, magic40((unsigned long long)G1), tw1.as<float2>(),
            twP.hi.as<float2>(), twP.lo.as<float2>(), 0, 0, n, 1, sign);
    }
    {
        KernelScope ks(ctx, HPFW_K_CQT, stream);
        fft_pass_kernel<1><<<(n1 + G2 - 1) / G2, CQ_FFT_THREADS, smem2, stream>>>(
            tmp.as<float2>(), d_out, nullptr, d2, n1, G2, magic40((unsigned long long)G2), tw2.as<float2>(), nullptr,
            nullptr, 0, 0, n, 1, sign);
    }
    HPFW_CUDA_TRY(cudaGetLastError());
    HPFW_CUDA_TRY(cudaStreamSynchronize(stream));
    tw1.release(); tw2.release(); tmp.release(); twP.release();
    return HPFW_OK;
}

}  // namespace hpfw_b200

using namespace hpfw_b200;

namespace hpfw_b200 {
// internal entry points for the extraction stream (xstream.cu): no stream-ordering hooks, the caller forks / joins the lanes
int cqt_run_lane(hpfw_ctx *ctx, const float *d_audio, int64_t n_samples, float *d_out, cudaStream_t stream, int lane) {
    return cqt_run(ctx, d_audio, n_samples, d_out, 0, stream, lane);
}
}  // namespace hpfw_b200

int hpfw_b200::ctx_lanes_init(hpfw_ctx *ctx) {
    if (ctx->lane_fork) return HPFW_OK;
    for (int l = 0; l < HPFW_CTX_LANES; ++l) {
        HPFW_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->lane_stream[l], cudaStreamNonBlocking));
        HPFW_CUDA_TRY(cudaEventCreateWithFlags(&ctx->lane_join[l], cudaEventDisableTiming));
    }
    HPFW_CUDA_TRY(cudaEventCreateWithFlags(&ctx->lane_fork, cudaEventDisableTiming));
    return HPFW_OK;
}

extern "C" {

#ifdef HPFW_CQT_PHASECLK
int hpfw_cqt_debug_phases(unsigned long long *out12, int reset) {
    if (out12 && cudaMemcpyFromSymbol(out12, g_phase_clk, sizeof(unsigned long long) * 12) != cudaSuccess) return HPFW_ERR_CUDA;
    if (reset) {
        unsigned long long z[12] = {};
        if (cudaMemcpyToSymbol(g_phase_clk, z, sizeof(z)) != cudaSuccess) return HPFW_ERR_CUDA;
    }
    return HPFW_OK;
}
#endif

int hpfw_cqt_design(int64_t n_samples, int *pos_out, int *lg_out, int *m_out) {
    CqtDesign d;
    if (!cqt_design(n_samples, d)) return HPFW_ERR_ARG;
    if (pos_out) memcpy(pos_out, d.pos, sizeof(d.pos));
    if (lg_out) memcpy(lg_out, d.lg, sizeof(d.lg));
    if (m_out) *m_out = d.M;
    return HPFW_OK;
}

int hpfw_set_cqt_window(hpfw_ctx *ctx, int window) {
    if (!ctx || window < 0 || window > 1) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_set_cqt_window: window must be 0 (periodic) or 1 (symmetric)");
    ctx->cqt_window = window;
    return HPFW_OK;
}

int hpfw_cqt_cols(int64_t n_samples) {
    CqtDesign d;
    if (!cqt_design(n_samples, d)) return 0;
    return d.cols;
}

int hpfw_hashprint_words_for_samples(int64_t n_samples) {
    const int cols = hpfw_cqt_cols(n_samples);
    return std::max(0, hpfw_hashprint_words_for_cols(cols));
}

int hpfw_cqt_spectrogram_device(hpfw_ctx *ctx, const float *d_audio, int64_t n_samples, float *d_out, void *stream) {
    if (!ctx || !d_audio || !d_out) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_cqt_spectrogram_device: NULL argument");
    DeviceGuard g(ctx->device);
    return cqt_run(ctx, d_audio, n_samples, d_out, 0, ctx->pick(stream));
}

static int cqt_host(hpfw_ctx *ctx, const float *audio, int64_t n_samples, float *out, int *cols_out, int mode) {
    if (!ctx || !audio || !out) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_cqt: NULL argument");
    DeviceGuard g(ctx->device);
    ctx->order_on(ctx->stream);
    const int cols = hpfw_cqt_cols(n_samples);
    if (cols <= 0) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_cqt: bad length");
    HPFW_TRY(ctx->audio.reserve(sizeof(float) * (size_t)n_samples));
    HPFW_TRY(ctx->spectro.reserve(sizeof(float) * (size_t)cols * CQ_BINS));
    HPFW_CUDA_TRY(cudaMemcpyAsync(ctx->audio.ptr, audio, sizeof(float) * (size_t)n_samples, cudaMemcpyHostToDevice,
                                  ctx->stream));
    HPFW_TRY(cqt_run(ctx, ctx->audio.as<float>(), n_samples, ctx->spectro.as<float>(), mode, ctx->stream));
    HPFW_CUDA_TRY(cudaMemcpyAsync(out, ctx->spectro.ptr, sizeof(float) * (size_t)cols * CQ_BINS, cudaMemcpyDeviceToHost,
                                  ctx->stream));
    HPFW_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (cols_out) *cols_out = cols;
    return HPFW_OK;
}

int hpfw_cqt_spectrogram(hpfw_ctx *ctx, const float *audio, int64_t n_samples, float *out, int *cols_out) {
    return cqt_host(ctx, audio, n_samples, out, cols_out, 0);
}

int hpfw_cqt_magnitude(hpfw_ctx *ctx, const float *audio, int64_t n_samples, float *out, int *cols_out) {
    return cqt_host(ctx, audio, n_samples, out, cols_out, 1);
}

int hpfw_calc_hashprint_audio_device(hpfw_ctx *ctx, const float *d_audio, int64_t n_samples, uint64_t *d_hp_out,
                                     void *stream) {
    if (!ctx || !d_audio || !d_hp_out) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_calc_hashprint_audio_device: NULL argument");
    DeviceGuard g(ctx->device);
    const int cols = hpfw_cqt_cols(n_samples);
    if (hpfw_hashprint_words_for_cols(cols) <= 0)
        HPFW_FAIL(HPFW_ERR_SHORT, "audio of %lld samples gives %d spectrogram columns; at least 100 are needed",
                  (long long)n_samples, cols);
    HPFW_TRY(ctx->spectro.reserve(sizeof(float) * (size_t)cols * CQ_BINS));
    cudaStream_t s = ctx->pick(stream);
    HPFW_TRY(cqt_run(ctx, d_audio, n_samples, ctx->spectro.as<float>(), 0, s));
    const int64_t co[2] = {0, cols};
    return hpfw_hashprint_from_spectrogram_device(ctx, ctx->spectro.as<float>(), co, 1, d_hp_out, s);
}

int hpfw_calc_hashprint_audio(hpfw_ctx *ctx, const float *audio, int64_t n_samples, uint64_t *hp_out, int *n_out) {
    if (!ctx || !audio || !hp_out || !n_out) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_calc_hashprint_audio: NULL argument");
    *n_out = 0;
    DeviceGuard g(ctx->device);
    ctx->order_on(ctx->stream);
    const int n = hpfw_hashprint_words_for_samples(n_samples);
    if (n <= 0)
        HPFW_FAIL(HPFW_ERR_SHORT, "audio of %lld samples is too short for one hashprint word", (long long)n_samples);
    HPFW_TRY(ctx->audio.reserve(sizeof(float) * (size_t)n_samples));
    HPFW_TRY(ctx->hp.reserve(sizeof(uint64_t) * (size_t)n));
    HPFW_CUDA_TRY(cudaMemcpyAsync(ctx->audio.ptr, audio, sizeof(float) * (size_t)n_samples, cudaMemcpyHostToDevice,
                                  ctx->stream));
    HPFW_TRY(hpfw_calc_hashprint_audio_device(ctx, ctx->audio.as<float>(), n_samples, ctx->hp.as<uint64_t>(), ctx->stream));
    HPFW_CUDA_TRY(cudaMemcpyAsync(hp_out, ctx->hp.ptr, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    HPFW_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *n_out = n;
    return HPFW_OK;
}

int hpfw_calc_hashprint_audio_batch_device(hpfw_ctx *ctx, const float *d_audio, const int64_t *sample_offsets, int n,
                                           uint64_t *d_hp_out, void *stream) {
    if (!ctx || !sample_offsets || n < 0) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_calc_hashprint_audio_batch_device: bad argument");
    if (n == 0) return HPFW_OK;
    if (!d_audio || !d_hp_out) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_calc_hashprint_audio_batch_device: NULL buffer");
    DeviceGuard g(ctx->device);
    cudaStream_t s = ctx->pick(stream);
    // chunks of tracks: CQT per track into one spectrogram buffer, then ONE projection launch for the chunk
    const size_t chunk_bytes = size_t(512) << 20;
    int i = 0;
    int64_t hp_off = 0;
    while (i < n) {
        std::vector<int64_t> co{0};
        int j = i;
        while (j < n) {
            const int64_t ns = sample_offsets[j + 1] - sample_offsets[j];
            if (ns < 0) HPFW_FAIL(HPFW_ERR_ARG, "track %d: offsets must be monotone", j);
            const int cols = hpfw_cqt_cols(ns);
            if (hpfw_hashprint_words_for_cols(cols) <= 0)
                HPFW_FAIL(HPFW_ERR_SHORT, "track %d (%lld samples) is too short for one hashprint word", j, (long long)ns);
            if (j > i && sizeof(float) * size_t(co.back() + cols) * CQ_BINS > chunk_bytes) break;
            co.push_back(co.back() + cols);
            ++j;
        }
        HPFW_TRY(ctx->spectro.reserve(sizeof(float) * size_t(co.back()) * CQ_BINS));
        // fork: the tracks of the chunk run round-robin on CQ_LANES streams (own scratch each) so that the small tail waves
        // of one track's kernels overlap another track's; join before the chunk's single projection launch
        HPFW_TRY(ctx_lanes_init(ctx));
        const int nl = std::max(1, std::min(CQ_LANES, env_int("HPFW_CQT_LANES", CQ_LANES)));
        HPFW_CUDA_TRY(cudaEventRecord(ctx->lane_fork, s));
        for (int l = 0; l < nl; ++l) HPFW_CUDA_TRY(cudaStreamWaitEvent(ctx->lane_stream[l], ctx->lane_fork, 0));
        // Tracks of one length whose transform takes the packed-FFT path go CQ_GANG_LANES at a time through a captured graph
        // (cqt_gang_run), the gangs alternating so that two graphs overlap; the rest (odd lengths out, Bluestein lengths,
        // timing mode) is launched kernel by kernel, one track per lane, stage by stage across the lanes.
        std::vector<int> plain;
        const bool trace = env_int("HPFW_TRACE", 0) != 0;
        const auto t_enq = std::chrono::steady_clock::now();
        const bool use_graph = env_int("HPFW_CQT_GRAPH", 1) != 0 && !ctx->timing && nl == CQ_LANES;
        if (use_graph) {
            std::map<int64_t, std::vector<int>> by_len;
            for (int t = i; t < j; ++t) by_len[sample_offsets[t + 1] - sample_offsets[t]].push_back(t);
            int gang = 0;
            for (auto &kv : by_len) {
                std::vector<int> &ts = kv.second;
                size_t k = 0;
                // capturing a graph costs a few plain tracks' worth of host time: only for lengths that repeat
                auto it = ctx->cqt ? ctx->cqt->plans.find(kv.first) : std::map<int64_t, std::unique_ptr<CqtPlan>>::iterator();
                const bool have = ctx->cqt && it != ctx->cqt->plans.end() && it->second->gangs != nullptr;
                if (have || ts.size() >= 2 * (size_t)CQ_GANG_LANES) {
                    for (; k + CQ_GANG_LANES <= ts.size(); k += CQ_GANG_LANES) {
                        CqtJob jobs[CQ_GANG_LANES];
                        for (int l = 0; l < CQ_GANG_LANES; ++l) {
                            const int t = ts[k + (size_t)l];
                            jobs[l].d_audio = d_audio + sample_offsets[t];
                            jobs[l].N = kv.first;
                            jobs[l].d_out = ctx->spectro.as<float>() + size_t(co[t - i]) * CQ_BINS;
                            jobs[l].lane = gang * CQ_GANG_LANES + l;
                            jobs[l].stream = ctx->lane_stream[jobs[l].lane];
                            HPFW_TRY(cqt_prepare(ctx, jobs[l]));
                        }
                        if (jobs[0].pl->bluestein) break;
                        HPFW_TRY(cqt_gang_run(ctx, jobs, gang));
                        gang = (gang + 1) % CQ_GANGS;
                    }
                }
                for (; k < ts.size(); ++k) plain.push_back(ts[k]);
            }
            std::sort(plain.begin(), plain.end());
            // the lanes' scratch was last used by the graphs, which ran on each gang's first lane stream
            if (!plain.empty())
                for (int g = 0; g < CQ_GANGS; ++g) {
                    const int l0 = g * CQ_GANG_LANES;
                    HPFW_CUDA_TRY(cudaEventRecord(ctx->lane_join[l0], ctx->lane_stream[l0]));
                    for (int l = 1; l < CQ_GANG_LANES; ++l)
                        HPFW_CUDA_TRY(cudaStreamWaitEvent(ctx->lane_stream[l0 + l], ctx->lane_join[l0], 0));
                }
        } else {
            for (int t = i; t < j; ++t) plain.push_back(t);
        }
        for (size_t t0 = 0; t0 < plain.size(); t0 += (size_t)nl) {
            CqtJob jobs[CQ_LANES];
            const int nj = (int)std::min((size_t)nl, plain.size() - t0);
            for (int l = 0; l < nj; ++l) {
                const int t = plain[t0 + (size_t)l];
                jobs[l].d_audio = d_audio + sample_offsets[t];
                jobs[l].N = sample_offsets[t + 1] - sample_offsets[t];
                jobs[l].d_out = ctx->spectro.as<float>() + size_t(co[t - i]) * CQ_BINS;
                jobs[l].stream = ctx->lane_stream[l];
                jobs[l].lane = l;
                HPFW_TRY(cqt_prepare(ctx, jobs[l]));
                // Bluestein plans are few (CQ_PLAN_CACHE_BLUESTEIN) and preparing the next one may evict this one: launch
                // such a track at once; the packed-path plans of one wave are the most recently used of CQ_PLAN_CACHE
                if (jobs[l].pl->bluestein) {
                    for (int st = 0; st < CQ_STAGES; ++st) HPFW_TRY(cqt_stage(ctx, jobs[l], st));
                    jobs[l].pl = nullptr;
                }
            }
            for (int st = 0; st < CQ_STAGES; ++st)
                for (int l = 0; l < nj; ++l)
                    if (jobs[l].pl) HPFW_TRY(cqt_stage(ctx, jobs[l], st));
        }
        if (trace)
            fprintf(stderr, "[hpfw] batch chunk: %d tracks, CQT launches enqueued in %.1f us/track (%zu kernel by kernel)\n",
                    j - i, std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_enq).count() / (j - i),
                    plain.size());
        for (int l = 0; l < nl; ++l) {
            HPFW_CUDA_TRY(cudaEventRecord(ctx->lane_join[l], ctx->lane_stream[l]));
            HPFW_CUDA_TRY(cudaStreamWaitEvent(s, ctx->lane_join[l], 0));
        }
        HPFW_TRY(hpfw_hashprint_from_spectrogram_device(ctx, ctx->spectro.as<float>(), co.data(), j - i, d_hp_out + hp_off, s));
        for (int t = i; t < j; ++t) hp_off += hpfw_hashprint_words_for_cols(int(co[t - i + 1] - co[t - i]));
        i = j;
    }
    return HPFW_OK;
}

int hpfw_fft_c2c(hpfw_ctx *ctx, const float *in_interleaved, float *out_interleaved, int n, int inverse) {
    if (!ctx || !in_interleaved || !out_interleaved || n < 4) HPFW_FAIL(HPFW_ERR_ARG, "hpfw_fft_c2c: bad argument");
    DeviceGuard g(ctx->device);
    ctx->order_on(ctx->stream);
    DeviceBuffer a, b;
    HPFW_TRY(a.reserve(sizeof(float2) * (size_t)n));
    HPFW_TRY(b.reserve(sizeof(float2) * (size_t)n));
    HPFW_CUDA_TRY(cudaMemcpy(a.ptr, in_interleaved, sizeof(float2) * (size_t)n, cudaMemcpyHostToDevice));
    int st = fft_c2c_device(ctx, a.as<float2>(), b.as<float2>(), n, inverse, ctx->stream);
    if (st == HPFW_OK) {
        cudaError_t e = cudaMemcpy(out_interleaved, b.ptr, sizeof(float2) * (size_t)n, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
            set_error("hpfw_fft_c2c: D2H failed: %s", cudaGetErrorString(e));
            st = HPFW_ERR_CUDA;
        }
    }
    a.release();
    b.release();
    return st;
}

}  // extern "C"
