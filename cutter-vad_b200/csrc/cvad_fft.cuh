// cvad_fft.cuh -- complex-256 FFT in double precision, the building block of the exact resampler and of v4's STFT.
//
// Why FP64: both users feed ill-conditioned consumers.  v4 takes log(1 + 2^20 |STFT|), so on band-limited input the
// FP32 rounding noise of a 256-term dot product (or of an FP32 FFT resampler in front of it) decides bins that hold
// ~1e-7 of the frame's energy and moves the probability by up to 3e-3 -- two FP32 CPU executors of the reference's own
// graph differ by 1.3e-3 there (tests/test_oracle_pinning.py).  Evaluated in double the same stages agree with the
// float64 interpretation of the graph to 1e-6.  B200 issues 64 FP64 operations per clock and SM (measured,
// tools/dev/fp64_rate.cu: 17.7 T DFMA/s), and an FFT needs ~30x fewer operations than the dense operator it replaces.
//
// Everything reduces to ONE transform: a 256-point complex FFT as 16 x 16 (four-step): 16 threads, each a 16-point FFT
// in registers per pass, two passes through a shared-memory buffer of 16 rows x 17 double2 (the odd row stride makes
// the column accesses of pass 1 and the row accesses of pass 2 both bank-conflict-free).
//
//   resampler (AudioUtils.resample_audio = scipy.signal.resample, /root/reference/src/real_time_vad/utils/audio.py:19-55):
//     n_x = 256 R source samples (R = 1, 3, 6 for 8 / 24 / 48 kHz) -> 512 samples.  scipy: X = rfft(x); keep bins
//     0..m/2 (m = min(512, n_x)); the unpaired bin m/2 is doubled (down-sampling) or halved (up-sampling);
//     y = irfft(X * 512 / n_x, 512).  Here: X[k] = sum_r W_nx^(r k) F_r[k mod 256] with F_r = DFT_256 of x[R n + r]
//     (decimation in time; only bins 0..256 are ever formed), two real sub-sequences per complex FFT; the inverse
//     real FFT of 512 points is one more complex-256 transform (even / odd output samples = real / imaginary part).
//   v4 STFT (silero_vad.onnx: reflect-pad 96|96, conv k256 s64 with basis = Hann x DFT-256): 8 windows per frame =
//     4 complex FFTs of two windows each.  The file's basis is the float32 image of Hann x DFT (|delta| <= 7.7e-8),
//     and that rounding matters as much as the arithmetic does; the difference is a dense product with a tiny matrix,
//     which a single BF16 pass on the tensor cores evaluates to 0.4 % (v4tc_stft_kernel<true>, cvad_v4tc.cuh).
//
// The per-thread routines below are __host__ __device__: tests/test_fft_host.py compiles this header with g++
// (tools/fft_host_check.cpp), runs the same code thread by thread on the CPU and compares with scipy in float64.
#pragma once

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#include <math.h>
#define CVAD_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define CVAD_HD inline
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
#endif

namespace cvad {
namespace fft {

constexpr int kRow = 17;               // row stride of a 16 x 16 buffer, in double2 elements
constexpr int kBuf = 16 * kRow;        // elements per complex-256 buffer (4,352 bytes)
constexpr int kMaster = 1536;          // twiddle table T[j] = exp(-2 pi i j / 1536): W_256 = T[6 .], W_512 = T[3 .], W_768 = T[2 .]

CVAD_HD double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
CVAD_HD double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
CVAD_HD double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
CVAD_HD double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
CVAD_HD double2 cscale(double2 a, double s) { return make_double2(a.x * s, a.y * s); }

// position of element n of a transform's INPUT (natural order) and of element k of its OUTPUT
CVAD_HD int pos_in(int n) { return (n >> 4) * kRow + (n & 15); }
CVAD_HD int pos_out(int k) { return (k & 15) * kRow + (k >> 4); }

// z * W16^M (forward) or z * conj(W16^M) (inverse), M a compile-time constant
template <int M, bool INV>
CVAD_HD double2 mul_w16(double2 z) {
    constexpr double c1 = 0.92387953251128673848, s1 = 0.38268343236508977173, h = 0.70710678118654752440;
    // W16^M = (cr, ci) with ci = -sin; the inverse uses the conjugate
    if (M == 0) return z;
    if (M == 4) return INV ? make_double2(-z.y, z.x) : make_double2(z.y, -z.x);                 // -+ i
    if (M == 2) return INV ? make_double2((z.x - z.y) * h, (z.x + z.y) * h) : make_double2((z.x + z.y) * h, (z.y - z.x) * h);
    if (M == 6) return INV ? make_double2(-(z.x + z.y) * h, (z.x - z.y) * h) : make_double2((z.y - z.x) * h, -(z.x + z.y) * h);
    double cr = 1.0, ci = 0.0;
    if (M == 1) { cr = c1; ci = -s1; }
    if (M == 3) { cr = s1; ci = -c1; }
    if (M == 9) { cr = -c1; ci = s1; }
    if (INV) ci = -ci;
    return make_double2(z.x * cr - z.y * ci, z.x * ci + z.y * cr);
}

template <bool INV>
CVAD_HD void bfly4(double2 &a, double2 &b, double2 &c, double2 &d) {
    const double2 t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = csub(b, d);
    const double2 it3 = INV ? make_double2(-t3.y, t3.x) : make_double2(t3.y, -t3.x);   // -+ i t3
    a = cadd(t0, t2);
    c = csub(t0, t2);
    b = cadd(t1, it3);
    d = csub(t1, it3);
}

// 16-point FFT in registers.  In: v[n].  Out: v[4 k1 + k2] = X[k1 + 4 k2].
template <bool INV>
CVAD_HD void fft16(double2 (&v)[16]) {
#define CVAD_B4(I0, I1, I2, I3) bfly4<INV>(v[I0], v[I1], v[I2], v[I3])
    CVAD_B4(0, 4, 8, 12); CVAD_B4(1, 5, 9, 13); CVAD_B4(2, 6, 10, 14); CVAD_B4(3, 7, 11, 15);
    // v[4 k1 + n2] *= W16^(n2 k1)
    v[5] = mul_w16<1, INV>(v[5]);   v[6] = mul_w16<2, INV>(v[6]);   v[7] = mul_w16<3, INV>(v[7]);
    v[9] = mul_w16<2, INV>(v[9]);   v[10] = mul_w16<4, INV>(v[10]); v[11] = mul_w16<6, INV>(v[11]);
    v[13] = mul_w16<3, INV>(v[13]); v[14] = mul_w16<6, INV>(v[14]); v[15] = mul_w16<9, INV>(v[15]);
    CVAD_B4(0, 1, 2, 3); CVAD_B4(4, 5, 6, 7); CVAD_B4(8, 9, 10, 11); CVAD_B4(12, 13, 14, 15);
#undef CVAD_B4
}
CVAD_HD int fft16_index(int j) { return (j >> 2) + 4 * (j & 3); }   // v[j] holds X[fft16_index(j)]

// Pass 1 of the 256-point transform, thread t = n2 (0..15): v[n1] = z[16 n1 + t] on entry.
// tw[m * tws] = exp(-2 pi i m / 256).  The 16 twiddles W^(t k1) are powers of ONE table entry W^t: they are formed by
// repeated multiplication (two interleaved chains, even and odd powers; 8 products deep = 1e-15 in double) instead of 16
// dependent table loads per pass -- the loads, not the FP64 pipe, were what the first version of the kernels waited for.
template <bool INV>
CVAD_HD void pass1_regs(double2 (&v)[16], double2 *buf, int t, const double2 *tw, int tws) {
    fft16<INV>(v);
    double2 w1 = tw[t * tws];
    if (INV) w1.y = -w1.y;
    const double2 w2 = cmul(w1, w1);
    double2 we = make_double2(1.0, 0.0), wo = w1;          // W^(t k1) for even / odd k1
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int k1 = 0; k1 < 16; k1 += 2) {
        // v[j] holds Y[fft16_index(j)]; fft16_index is an involution (it swaps the two base-4 digits)
        buf[k1 * kRow + t] = k1 == 0 ? v[0] : cmul(v[fft16_index(k1)], we);
        buf[(k1 + 1) * kRow + t] = cmul(v[fft16_index(k1 + 1)], wo);
        we = cmul(we, w2);
        wo = cmul(wo, w2);
    }
}
template <bool INV>
CVAD_HD void pass1(double2 *buf, int t, const double2 *tw, int tws) {
    double2 v[16];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int n1 = 0; n1 < 16; ++n1) v[n1] = buf[n1 * kRow + t];
    pass1_regs<INV>(v, buf, t, tw, tws);
}
// Pass 2, thread t = k1: Z[k1 + 16 k2] -> buf[k1 * 17 + k2] = buf[pos_out(k)]
template <bool INV>
CVAD_HD void pass2(double2 *buf, int t) {
    double2 v[16];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int n2 = 0; n2 < 16; ++n2) v[n2] = buf[t * kRow + n2];
    fft16<INV>(v);
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int j = 0; j < 16; ++j) buf[t * kRow + fft16_index(j)] = v[j];
}

// Two real sequences per transform: Zf = FFT(a + i b)  ->  A[k], B[k]  (0 <= k <= 128)
CVAD_HD void unpack2(const double2 *buf, int k, double2 &A, double2 &B) {
    const double2 zk = buf[pos_out(k)], zm = buf[pos_out((256 - k) & 255)];
    A = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y));
    B = make_double2(0.5 * (zk.y + zm.y), -0.5 * (zk.x - zm.x));
}

// ---- resampler, spectral stage.  bufs = the frame's (R + 1) / 2 forward transforms (sub-sequences 2p, 2p + 1 in
// transform p).  Returns the inverse transform's inputs Z[k] and Z[256 - k] for one k in 0..128.
//   X[k]       = sum_r W^(r k) F_r[k]                                    (W = exp(-2 pi i / n_x), F_r = DFT_256 of x[R n + r])
//   X[256 - k] = sum_r W^(r (256 - k)) conj(F_r[k]) = sum_r c_r conj(W^(r k) F_r[k]),   c_r = exp(-2 pi i r / R)
// so both come from the same R products, and W^(r k) = (W^k)^r needs ONE table entry per bin.
template <int R>
CVAD_HD void rs_spectrum(const double2 *bufs, int k, const double2 *T, double2 &Zk, double2 &Zm) {
    constexpr int NX = 256 * R;
    constexpr int TS = kMaster / NX;                     // T[TS m] = W_nx^m
    constexpr double H3 = 0.86602540378443864676;        // sqrt(3) / 2
    double2 P = make_double2(0.0, 0.0), Q = make_double2(0.0, 0.0);
    const double2 wk = R == 1 ? make_double2(1.0, 0.0) : T[k * TS];
    double2 wr = make_double2(1.0, 0.0);                 // W^(r k)
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int p = 0; p < (R + 1) / 2; ++p) {
        double2 F[2];
        unpack2(bufs + p * kBuf, k, F[0], F[1]);
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int c = 0; c < 2; ++c) {
            const int r = 2 * p + c;
            if (r >= R) break;
            const double2 g = r == 0 ? F[c] : cmul(F[c], wr);       // W^(r k) F_r[k]
            P = cadd(P, g);
            if (R > 1) {
                // c_r conj(g): c_r = exp(-2 pi i r / R), an R-th root of unity known at compile time
                const int q = (r * 6) / R;                           // c_r = exp(-2 pi i q / 6), q = 0..5
                const double cr = (q == 0) ? 1.0 : (q == 3) ? -1.0 : (q == 1 || q == 5) ? 0.5 : -0.5;
                const double ci = (q == 0 || q == 3) ? 0.0 : (q < 3 ? -H3 : H3);
                Q = cadd(Q, cmul(make_double2(cr, ci), cconj(g)));
            }
            wr = cmul(wr, wk);
        }
    }
    const double inv = 1.0 / (double)NX;
    P = cscale(P, inv);
    Q = cscale(Q, inv);
    if (R == 1) {
        // up-sampling: bins above 128 are empty, the source's Nyquist bin is halved (scipy: X[m/2] *= 0.5)
        if (k == 128) { P = cscale(P, 0.5); Q = P; }
    } else if (k == 0) {
        // down-sampling: the folded Nyquist bin is doubled and irfft keeps its real part (scipy: X[m/2] *= 2)
        Q = make_double2(2.0 * Q.x, 0.0);
    }
    const double2 w = cconj(T[3 * k]);                                           // exp(+2 pi i k / 512)
    const double2 G = cadd(P, cconj(Q));
    const double2 H = cmul(csub(P, cconj(Q)), w);
    Zk = make_double2(G.x - H.y, G.y + H.x);                                     // G + i H
    Zm = make_double2(G.x + H.y, H.x - G.y);                                     // conj(G) + i conj(H) = Z[256 - k]
}

// periodic Hann window of 256 points from the master table: 0.5 - 0.5 cos(2 pi n / 256)
CVAD_HD double hann256(const double2 *T, int n) { return 0.5 - 0.5 * T[6 * n].x; }

// host: the master twiddle table, exp(-2 pi i j / 1536) in double (octant-reduced so that the table is exactly symmetric)
inline void build_master(double2 *T) {
    const long double PI = 3.14159265358979323846264338327950288L;
    for (int j = 0; j < kMaster; ++j) {
        const long double a = 2.0L * PI * (long double)j / (long double)kMaster;
        T[j].x = (double)cosl(a);
        T[j].y = (double)(-sinl(a));
    }
    T[0].x = 1.0; T[0].y = 0.0;
    T[kMaster / 4].x = 0.0; T[kMaster / 4].y = -1.0;
    T[kMaster / 2].x = -1.0; T[kMaster / 2].y = 0.0;
    T[3 * kMaster / 4].x = 0.0; T[3 * kMaster / 4].y = 1.0;
}

}  // namespace fft
}  // namespace cvad
