// Host build of csrc/cvad_fft.cuh: the per-thread routines the CUDA kernels run, executed thread by thread on the CPU so
// that tests/test_fft_host.py can compare them with scipy / numpy in float64 without a GPU.  Test infrastructure only.
#include <string.h>
#include <vector>

#include "../cutter-vad_b200/csrc/cvad_fft.cuh"

using namespace cvad::fft;

static std::vector<double2> &master() {
    static std::vector<double2> T;
    if (T.empty()) { T.resize(kMaster); build_master(T.data()); }
    return T;
}

template <bool INV>
static void run256(double2 *buf) {
    const double2 *T = master().data();
    for (int t = 0; t < 16; ++t) pass1<INV>(buf, t, T, 6);     // "16 threads", then the barrier
    for (int t = 0; t < 16; ++t) pass2<INV>(buf, t);
}

template <int R>
static void resample_frame(const float *x, float *y) {
    constexpr int NP = (R + 1) / 2;
    const double2 *T = master().data();
    std::vector<double2> bufs(NP * kBuf, make_double2(0.0, 0.0));
    for (int m = 0; m < 256 * R; ++m) {            // the kernel's placement: sample m -> transform (m % R) / 2, element m / R
        const int r = m % R, n = m / R;
        double2 &z = bufs[(r >> 1) * kBuf + pos_in(n)];
        if (r & 1) z.y = (double)x[m]; else z.x = (double)x[m];
    }
    for (int p = 0; p < NP; ++p) run256<false>(bufs.data() + p * kBuf);
    double2 Zk[129], Zm[129];
    for (int k = 0; k <= 128; ++k) rs_spectrum<R>(bufs.data(), k, T, Zk[k], Zm[k]);
    double2 *b0 = bufs.data();
    for (int k = 0; k <= 128; ++k) {
        b0[pos_in(k)] = Zk[k];
        if (k > 0 && k < 128) b0[pos_in(256 - k)] = Zm[k];
    }
    run256<true>(b0);
    for (int n = 0; n < 256; ++n) {
        const double2 z = b0[pos_out(n)];
        y[2 * n] = (float)z.x;
        y[2 * n + 1] = (float)z.y;
    }
}

extern "C" {

// in/out: 256 complex values as interleaved doubles
void fftchk_fft256(const double *in, double *out, int inverse) {
    std::vector<double2> buf(kBuf);
    for (int n = 0; n < 256; ++n) buf[pos_in(n)] = make_double2(in[2 * n], in[2 * n + 1]);
    if (inverse) run256<true>(buf.data()); else run256<false>(buf.data());
    for (int k = 0; k < 256; ++k) { out[2 * k] = buf[pos_out(k)].x; out[2 * k + 1] = buf[pos_out(k)].y; }
}

// one chunk of 256 R source samples -> 512 samples (float32, as AudioUtils.resample_audio returns them)
int fftchk_resample(const float *x, int R, float *y) {
    if (R == 1) resample_frame<1>(x, y);
    else if (R == 3) resample_frame<3>(x, y);
    else if (R == 6) resample_frame<6>(x, y);
    else return -1;
    return 0;
}

// reflect-padded frame xp[704] -> STFT with the exact Hann x DFT-256 basis: re / im [8 columns][129 bins], double
void fftchk_stft(const float *xp, double *re, double *im) {
    const double2 *T = master().data();
    std::vector<double2> buf(kBuf);
    for (int j = 0; j < 4; ++j) {
        for (int t = 0; t < 16; ++t) {             // what thread t of the kernel does: its 16 inputs straight into pass 1
            double2 v[16];
            for (int i = 0; i < 16; ++i) {
                const int n = 16 * i + t;
                const double w = hann256(T, n);
                v[i] = make_double2(w * (double)xp[128 * j + n], w * (double)xp[128 * j + 64 + n]);
            }
            pass1_regs<false>(v, buf.data(), t, T, 6);
        }
        for (int t = 0; t < 16; ++t) pass2<false>(buf.data(), t);
        for (int k = 0; k <= 128; ++k) {
            double2 A, B;
            unpack2(buf.data(), k, A, B);
            re[(2 * j) * 129 + k] = A.x; im[(2 * j) * 129 + k] = A.y;
            re[(2 * j + 1) * 129 + k] = B.x; im[(2 * j + 1) * 129 + k] = B.y;
        }
    }
}

}  // extern "C"
