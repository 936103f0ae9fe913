/*
 * ORACLE (test infrastructure, never shipped): plain-C restatement of the Silero
 * VAD v5 / v4 16 kHz graphs as the reference runs them, plus the start/end state
 * machine.  It follows, stage by stage, what `session.run` executes at
 *   /root/reference/src/real_time_vad/core/silero_model.py:433
 * on the model files under /root/reference/src/real_time_vad/models/ (the
 * arithmetic itself lives in the unpinned third-party wheel `onnxruntime>=1.10.0`,
 * pyproject.toml:32, absent from this image), and the state machine at
 *   silero_model.py:790-923.
 * The restatement is validated against the op-by-op ONNX interpreter
 * (oracle/onnx_interp.py) in tests/test_oracle_pinning.py, and that interpreter against
 * two third-party executors present in this image: OpenCV's DNN module running the
 * reference's own graphs, and PyTorch's conv1d / LSTMCell.  Not pinned: onnxruntime's
 * own rounding -- the wheel cannot be installed here and the reference's tests hold no
 * golden probabilities (tests/test_silero_model.py:278-292 mocks every run).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product never links it.
 *
 * Canonical weight blob (float32, this order) -- the same order the product's
 * C ABI takes in cvad_create():
 *   v5: basis[258][256] enc0_w[128][129][3] enc0_b[128] enc1_w[64][128][3] enc1_b[64]
 *       enc2_w[64][64][3] enc2_b[64] enc3_w[128][64][3] enc3_b[128]
 *       w_ih[512][128] w_hh[512][128] b_ih[512] b_hh[512] dec_w[128] dec_b[1]
 *       (PyTorch gate order i,f,g,o)                      = 309,633 floats
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define V5_BLOB_FLOATS 309633

typedef struct {
    /* transposed so the inner loop runs over contiguous outputs and every
       output accumulates its terms in ascending-k order */
    float *basis_t; /* [256 k][258 f] */
    float *e0_t;    /* [129 c][3 tap][128 o] */
    float *e1_t;    /* [128 c][3][64] */
    float *e2_t;    /* [64 c][3][64] */
    float *e3_t;    /* [64 c][3][128] */
    float *wih_t;   /* [128 k][512] */
    float *whh_t;   /* [128 k][512] */
    float e0_b[128], e1_b[64], e2_b[64], e3_b[128];
    float b_ih[512], b_hh[512];
    float dec_w[128];
    float dec_b;
} sref_v5;

static float *transpose_conv(const float *w, int O, int C, int K) {
    /* w[o][c][k] -> t[c][k][o] */
    float *t = (float *)malloc(sizeof(float) * (size_t)O * C * K);
    for (int o = 0; o < O; ++o)
        for (int c = 0; c < C; ++c)
            for (int k = 0; k < K; ++k)
                t[((size_t)c * K + k) * O + o] = w[((size_t)o * C + c) * K + k];
    return t;
}

sref_v5 *sref_v5_create(const float *blob) {
    sref_v5 *m = (sref_v5 *)calloc(1, sizeof(sref_v5));
    const float *p = blob;
    m->basis_t = transpose_conv(p, 258, 256, 1); /* [f][k] -> [k][f] */
    p += 258 * 256;
    m->e0_t = transpose_conv(p, 128, 129, 3); p += 128 * 129 * 3;
    memcpy(m->e0_b, p, sizeof(m->e0_b)); p += 128;
    m->e1_t = transpose_conv(p, 64, 128, 3); p += 64 * 128 * 3;
    memcpy(m->e1_b, p, sizeof(m->e1_b)); p += 64;
    m->e2_t = transpose_conv(p, 64, 64, 3); p += 64 * 64 * 3;
    memcpy(m->e2_b, p, sizeof(m->e2_b)); p += 64;
    m->e3_t = transpose_conv(p, 128, 64, 3); p += 128 * 64 * 3;
    memcpy(m->e3_b, p, sizeof(m->e3_b)); p += 128;
    m->wih_t = transpose_conv(p, 512, 128, 1); p += 512 * 128;
    m->whh_t = transpose_conv(p, 512, 128, 1); p += 512 * 128;
    memcpy(m->b_ih, p, sizeof(m->b_ih)); p += 512;
    memcpy(m->b_hh, p, sizeof(m->b_hh)); p += 512;
    memcpy(m->dec_w, p, sizeof(m->dec_w)); p += 128;
    m->dec_b = *p;
    return m;
}

void sref_v5_free(sref_v5 *m) {
    if (!m) return;
    free(m->basis_t); free(m->e0_t); free(m->e1_t); free(m->e2_t); free(m->e3_t);
    free(m->wih_t); free(m->whh_t); free(m);
}

static inline float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

/* conv1d over a [C][T] input with kernel 3, zero padding 1, given stride.
   wt is [c][k][O]; out is [O][Tout].  Zero-pad taps are skipped (they add 0). */
static void conv3(const float *in, int C, int T, const float *wt, const float *b,
                  int O, int stride, int Tout, float *out) {
    float acc[128];
    for (int to = 0; to < Tout; ++to) {
        for (int o = 0; o < O; ++o) acc[o] = 0.0f;
        for (int c = 0; c < C; ++c)
            for (int k = 0; k < 3; ++k) {
                int ti = to * stride + k - 1;
                if (ti < 0 || ti >= T) continue;
                float v = in[c * T + ti];
                const float *w = wt + ((size_t)c * 3 + k) * O;
                for (int o = 0; o < O; ++o) acc[o] += w[o] * v;
            }
        for (int o = 0; o < O; ++o) {
            float r = acc[o] + b[o];
            out[o * Tout + to] = r > 0.0f ? r : 0.0f;
        }
    }
}

/* dbg (optional, 1603 floats): mag[129][3] e0[128][3] e1[64][2] e2[64] e3[128] gates[512] (i,f,g,o pre-activation) */
#define V5_DBG_FLOATS 1603

/* One 512-sample frame (already padded / gated by the caller), one stream.
   v5.onnx then-branch (sr == 16000), SURVEY.md section 8a stage table:
   reflect-pad right 64 (never read) -> conv k256 s128 -> magnitude -> 4x conv+relu
   -> LSTMCell(128) -> relu -> 1x1 conv -> sigmoid. */
void sref_v5_frame(const sref_v5 *m, const float *x, float *h, float *c, float *prob, float *dbg) {
    float spec[3][258];
    float mag[129 * 3], e0[128 * 3], e1[64 * 2], e2[64], e3[128], gates[512];
    for (int t = 0; t < 3; ++t) {
        float *s = spec[t];
        for (int f = 0; f < 258; ++f) s[f] = 0.0f;
        const float *xs = x + 128 * t;
        for (int k = 0; k < 256; ++k) {
            float v = xs[k];
            const float *w = m->basis_t + (size_t)k * 258;
            for (int f = 0; f < 258; ++f) s[f] += w[f] * v;
        }
        for (int f = 0; f < 129; ++f) mag[f * 3 + t] = sqrtf(s[f] * s[f] + s[129 + f] * s[129 + f]);
    }
    conv3(mag, 129, 3, m->e0_t, m->e0_b, 128, 1, 3, e0);
    conv3(e0, 128, 3, m->e1_t, m->e1_b, 64, 2, 2, e1);
    conv3(e1, 64, 2, m->e2_t, m->e2_b, 64, 2, 1, e2);
    conv3(e2, 64, 1, m->e3_t, m->e3_b, 128, 1, 1, e3);
    for (int n = 0; n < 512; ++n) gates[n] = 0.0f;
    for (int k = 0; k < 128; ++k) {
        float v = e3[k];
        const float *w = m->wih_t + (size_t)k * 512;
        for (int n = 0; n < 512; ++n) gates[n] += w[n] * v;
    }
    for (int k = 0; k < 128; ++k) {
        float v = h[k];
        const float *w = m->whh_t + (size_t)k * 512;
        for (int n = 0; n < 512; ++n) gates[n] += w[n] * v;
    }
    for (int n = 0; n < 512; ++n) gates[n] += m->b_ih[n] + m->b_hh[n];
    float acc = 0.0f;
    for (int j = 0; j < 128; ++j) {
        float ig = sigmoidf_(gates[j]);
        float fg = sigmoidf_(gates[128 + j]);
        float gg = tanhf(gates[256 + j]);
        float og = sigmoidf_(gates[384 + j]);
        float cn = fg * c[j] + ig * gg;
        float hn = og * tanhf(cn);
        c[j] = cn;
        h[j] = hn;
        acc += m->dec_w[j] * (hn > 0.0f ? hn : 0.0f);
    }
    *prob = sigmoidf_(acc + m->dec_b);
    if (dbg) {
        float *d = dbg;
        memcpy(d, mag, sizeof(mag)); d += 387;
        memcpy(d, e0, sizeof(e0)); d += 384;
        memcpy(d, e1, sizeof(e1)); d += 128;
        memcpy(d, e2, sizeof(e2)); d += 64;
        memcpy(d, e3, sizeof(e3)); d += 128;
        memcpy(d, gates, sizeof(gates));
    }
}

/* Frame loader as the reference's Python does it:
 *   split_into_frames  audio.py:164-190 (frame j = samples [j*hop, j*hop+frame_len))
 *   denoise_audio      audio.py:104-121 (|x| > float32(0.01) ? x : 0)
 *   _prepare_audio_input silero_model.py:449-474 (zero-pad right / truncate to 512) */
static void load_frame(const float *src, int frame_len, int denoise, float *x512) {
    int n = frame_len < 512 ? frame_len : 512;
    for (int i = 0; i < n; ++i) {
        float v = src[i];
        if (denoise && !(fabsf(v) > 0.01f)) v = 0.0f;
        x512[i] = v;
    }
    for (int i = n; i < 512; ++i) x512[i] = 0.0f;
}

/* n_streams independent streams, n_frames frames each; audio[s*stride + j*hop + i].
   h,c: [n_streams][128] in/out.  probs: [n_streams][n_frames]. */
void sref_v5_run(const sref_v5 *m, const float *audio, long stride, int n_streams, int n_frames,
                 int hop, int frame_len, int denoise, float *h, float *c, float *probs, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int s = 0; s < n_streams; ++s) {
        float x[512];
        for (int j = 0; j < n_frames; ++j) {
            load_frame(audio + (size_t)s * stride + (size_t)j * hop, frame_len, denoise, x);
            sref_v5_frame(m, x, h + (size_t)s * 128, c + (size_t)s * 128, probs + (size_t)s * n_frames + j, 0);
        }
    }
}


/* ========================================================================
 * Silero VAD v4, 16 kHz branch (silero_vad.onnx, If(sr==16000) then-branch),
 * SURVEY.md section 8a table "S4".  Canonical blob (155,908 floats):
 *   basis[258][256] norm_filter[7]
 *   first: dw_w[258][5] dw_b[258] pw_w[16][258] pw_b[16] proj_w[16][258] proj_b[16]
 *   c1_w[16][16] c1_b[16]                                   (initializers 1110,1111; stride 2)
 *   enc3: dw_w[16][5] dw_b[16] pw_w[32][16] pw_b[32] proj_w[32][16] proj_b[32]
 *   c2_w[32][32] c2_b[32]                                   (1113,1114; stride 2)
 *   enc7: dw_w[32][5] dw_b[32] pw_w[32][32] pw_b[32]        (identity residual)
 *   c3_w[32][32] c3_b[32]                                   (1116,1117; stride 2)
 *   enc11: dw_w[32][5] dw_b[32] pw_w[64][32] pw_b[64] proj_w[64][32] proj_b[64]
 *   c4_w[64][64] c4_b[64]                                   (1119,1120)
 *   lstm1: W[256][64] R[256][64] B[512]   lstm2: W[256][64] R[256][64] B[512]   (ONNX gate order i,o,f,c)
 *   dec_w[64] dec_b[1]
 */
#define V4_BLOB_FLOATS 155908

typedef struct {
    const float *basis, *nfilt;
    const float *f_dw, *f_dwb, *f_pw, *f_pwb, *f_pj, *f_pjb;
    const float *c1, *c1b;
    const float *e3_dw, *e3_dwb, *e3_pw, *e3_pwb, *e3_pj, *e3_pjb;
    const float *c2, *c2b;
    const float *e7_dw, *e7_dwb, *e7_pw, *e7_pwb;
    const float *c3, *c3b;
    const float *e11_dw, *e11_dwb, *e11_pw, *e11_pwb, *e11_pj, *e11_pjb;
    const float *c4, *c4b;
    const float *l1w, *l1r, *l1b, *l2w, *l2r, *l2b;
    const float *decw, *decb;
    float *basis_t; /* [256 k][258 f] */
    float *blob;
} sref_v4;

sref_v4 *sref_v4_create(const float *blob_in) {
    sref_v4 *m = (sref_v4 *)calloc(1, sizeof(sref_v4));
    m->blob = (float *)malloc(sizeof(float) * V4_BLOB_FLOATS);
    memcpy(m->blob, blob_in, sizeof(float) * V4_BLOB_FLOATS);
    const float *p = m->blob;
#define TAKE(field, n) m->field = p; p += (n)
    TAKE(basis, 258 * 256); TAKE(nfilt, 7);
    TAKE(f_dw, 258 * 5); TAKE(f_dwb, 258); TAKE(f_pw, 16 * 258); TAKE(f_pwb, 16); TAKE(f_pj, 16 * 258); TAKE(f_pjb, 16);
    TAKE(c1, 256); TAKE(c1b, 16);
    TAKE(e3_dw, 80); TAKE(e3_dwb, 16); TAKE(e3_pw, 512); TAKE(e3_pwb, 32); TAKE(e3_pj, 512); TAKE(e3_pjb, 32);
    TAKE(c2, 1024); TAKE(c2b, 32);
    TAKE(e7_dw, 160); TAKE(e7_dwb, 32); TAKE(e7_pw, 1024); TAKE(e7_pwb, 32);
    TAKE(c3, 1024); TAKE(c3b, 32);
    TAKE(e11_dw, 160); TAKE(e11_dwb, 32); TAKE(e11_pw, 2048); TAKE(e11_pwb, 64); TAKE(e11_pj, 2048); TAKE(e11_pjb, 64);
    TAKE(c4, 4096); TAKE(c4b, 64);
    TAKE(l1w, 16384); TAKE(l1r, 16384); TAKE(l1b, 512); TAKE(l2w, 16384); TAKE(l2r, 16384); TAKE(l2b, 512);
    TAKE(decw, 64); TAKE(decb, 1);
#undef TAKE
    m->basis_t = transpose_conv(m->basis, 258, 256, 1);
    return m;
}

void sref_v4_free(sref_v4 *m) {
    if (!m) return;
    free(m->basis_t); free(m->blob); free(m);
}

/* depthwise conv k5 pad 2 + bias + relu over [C][T] */
static void dw5_relu(const float *in, int C, int T, const float *w, const float *b, float *out) {
    for (int c = 0; c < C; ++c)
        for (int t = 0; t < T; ++t) {
            float a = 0.0f;
            for (int d = 0; d < 5; ++d) {
                int ti = t + d - 2;
                if (ti >= 0 && ti < T) a += w[c * 5 + d] * in[c * T + ti];
            }
            a += b[c];
            out[c * T + t] = a > 0.0f ? a : 0.0f;
        }
}

/* 1x1 conv with stride over [Cin][T] -> [Cout][Tout], no activation */
static void pw(const float *in, int Cin, int T, const float *w, const float *b, int Cout, int stride, int Tout,
               float *out) {
    for (int o = 0; o < Cout; ++o)
        for (int t = 0; t < Tout; ++t) {
            float a = 0.0f;
            for (int c = 0; c < Cin; ++c) a += w[o * Cin + c] * in[c * T + t * stride];
            out[o * Tout + t] = a + b[o];
        }
}

static void relu_n(float *x, int n) { for (int i = 0; i < n; ++i) x[i] = x[i] > 0.0f ? x[i] : 0.0f; }

static void lstm_iofc(const float *W, const float *R, const float *B, const float *x, float *h, float *c) {
    float g[256];
    for (int n = 0; n < 256; ++n) {
        float a = 0.0f;
        for (int k = 0; k < 64; ++k) a += W[n * 64 + k] * x[k];
        float r = 0.0f;
        for (int k = 0; k < 64; ++k) r += R[n * 64 + k] * h[k];
        g[n] = a + r + B[n] + B[256 + n];
    }
    for (int j = 0; j < 64; ++j) {
        float ig = sigmoidf_(g[j]), og = sigmoidf_(g[64 + j]), fg = sigmoidf_(g[128 + j]), cg = tanhf(g[192 + j]);
        float cn = fg * c[j] + ig * cg;
        c[j] = cn;
        h[j] = og * tanhf(cn);
    }
}

/* dbg (optional): mag[129][8] norm[129][8] r3[16][8] r7[16][4] r15[32][4] r19[32][2] r27[32][2] r31[32] r39[64] r43[64] */
#define V4_DBG_FLOATS (1032 + 1032 + 128 + 64 + 128 + 64 + 64 + 32 + 64 + 64)

/* h, c: [2][64] (layer-major), as the reference keeps them (silero_model.py:397-401). */
void sref_v4_frame(const sref_v4 *m, const float *x, float *h, float *c, float *prob, float *dbg) {
    float xp[704];
    for (int i = 0; i < 96; ++i) xp[i] = x[96 - i];            /* reflect, no edge repeat */
    for (int i = 0; i < 512; ++i) xp[96 + i] = x[i];
    for (int i = 0; i < 96; ++i) xp[608 + i] = x[510 - i];
    float mag[129 * 8], sp[129 * 8], x1[258 * 8];
    for (int t = 0; t < 8; ++t) {
        float s[258];
        for (int f = 0; f < 258; ++f) s[f] = 0.0f;
        for (int k = 0; k < 256; ++k) {
            float v = xp[64 * t + k];
            const float *w = m->basis_t + (size_t)k * 258;
            for (int f = 0; f < 258; ++f) s[f] += w[f] * v;
        }
        for (int f = 0; f < 129; ++f) {
            float mg = sqrtf(s[f] * s[f] + s[129 + f] * s[129 + f]);
            mag[f * 8 + t] = mg;
            sp[f * 8 + t] = logf(1.0f + mg * 1048576.0f);
        }
    }
    /* adaptive normalisation: mean over bins, reflect-pad 3, 7-tap filter, mean over time */
    float mean[8], mp[14];
    for (int t = 0; t < 8; ++t) {
        float a = 0.0f;
        for (int f = 0; f < 129; ++f) a += sp[f * 8 + t];
        mean[t] = a / 129.0f;
    }
    for (int i = 0; i < 3; ++i) mp[i] = mean[3 - i];
    for (int i = 0; i < 8; ++i) mp[3 + i] = mean[i];
    for (int i = 0; i < 3; ++i) mp[11 + i] = mean[6 - i];
    float mm = 0.0f;
    for (int t = 0; t < 8; ++t) {
        float a = 0.0f;
        for (int d = 0; d < 7; ++d) a += m->nfilt[d] * mp[t + d];
        mm += a;
    }
    mm /= 8.0f;
    for (int i = 0; i < 1032; ++i) { x1[i] = mag[i]; x1[1032 + i] = sp[i] - mm; }

    float d1[258 * 8], a16[16 * 8], b16[16 * 8], r7[16 * 4];
    dw5_relu(x1, 258, 8, m->f_dw, m->f_dwb, d1);
    pw(d1, 258, 8, m->f_pw, m->f_pwb, 16, 1, 8, a16);
    pw(x1, 258, 8, m->f_pj, m->f_pjb, 16, 1, 8, b16);
    for (int i = 0; i < 128; ++i) a16[i] += b16[i];
    relu_n(a16, 128);                                           /* r3 */
    pw(a16, 16, 8, m->c1, m->c1b, 16, 2, 4, r7); relu_n(r7, 64);
    float d3[16 * 4], a32[32 * 4], b32[32 * 4], r19[32 * 2];
    dw5_relu(r7, 16, 4, m->e3_dw, m->e3_dwb, d3);
    pw(d3, 16, 4, m->e3_pw, m->e3_pwb, 32, 1, 4, a32);
    pw(r7, 16, 4, m->e3_pj, m->e3_pjb, 32, 1, 4, b32);
    for (int i = 0; i < 128; ++i) a32[i] += b32[i];
    relu_n(a32, 128);                                           /* r15 */
    pw(a32, 32, 4, m->c2, m->c2b, 32, 2, 2, r19); relu_n(r19, 64);
    float d7[32 * 2], r27[32 * 2], r31[32];
    dw5_relu(r19, 32, 2, m->e7_dw, m->e7_dwb, d7);
    pw(d7, 32, 2, m->e7_pw, m->e7_pwb, 32, 1, 2, r27);
    for (int i = 0; i < 64; ++i) r27[i] += r19[i];
    relu_n(r27, 64);
    pw(r27, 32, 2, m->c3, m->c3b, 32, 2, 1, r31); relu_n(r31, 32);
    float d11[32], r39[64], q39[64], r43[64];
    dw5_relu(r31, 32, 1, m->e11_dw, m->e11_dwb, d11);
    pw(d11, 32, 1, m->e11_pw, m->e11_pwb, 64, 1, 1, r39);
    pw(r31, 32, 1, m->e11_pj, m->e11_pjb, 64, 1, 1, q39);
    for (int i = 0; i < 64; ++i) r39[i] += q39[i];
    relu_n(r39, 64);
    pw(r39, 64, 1, m->c4, m->c4b, 64, 1, 1, r43); relu_n(r43, 64);

    lstm_iofc(m->l1w, m->l1r, m->l1b, r43, h, c);
    lstm_iofc(m->l2w, m->l2r, m->l2b, h, h + 64, c + 64);
    float acc = 0.0f;
    for (int j = 0; j < 64; ++j) acc += m->decw[j] * (h[64 + j] > 0.0f ? h[64 + j] : 0.0f);
    *prob = sigmoidf_(acc + m->decb[0]);
    if (dbg) {
        float *d = dbg;
        memcpy(d, mag, 4 * 1032); d += 1032;
        memcpy(d, x1 + 1032, 4 * 1032); d += 1032;
        memcpy(d, a16, 4 * 128); d += 128;
        memcpy(d, r7, 4 * 64); d += 64;
        memcpy(d, a32, 4 * 128); d += 128;
        memcpy(d, r19, 4 * 64); d += 64;
        memcpy(d, r27, 4 * 64); d += 64;
        memcpy(d, r31, 4 * 32); d += 32;
        memcpy(d, r39, 4 * 64); d += 64;
        memcpy(d, r43, 4 * 64);
    }
}

/* h,c: [n_streams][2][64] */
void sref_v4_run(const sref_v4 *m, const float *audio, long stride, int n_streams, int n_frames, int hop,
                 int frame_len, int denoise, float *h, float *c, float *probs, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int s = 0; s < n_streams; ++s) {
        float x[512];
        for (int j = 0; j < n_frames; ++j) {
            load_frame(audio + (size_t)s * stride + (size_t)j * hop, frame_len, denoise, x);
            sref_v4_frame(m, x, h + (size_t)s * 128, c + (size_t)s * 128, probs + (size_t)s * n_frames + j, 0);
        }
    }
}
int sref_v4_blob_floats(void) { return V4_BLOB_FLOATS; }
int sref_v4_dbg_floats(void) { return V4_DBG_FLOATS; }

/* ------------------------------------------------------------------------
 * Start / end state machine, silero_model.py:790-923 restated with the deques
 * kept literally (recent_start_frames maxlen 20 :620-623, recent_end_frames
 * maxlen 100 :625-628) so that the ratio tests are evaluated, not assumed.
 * state[]: [0] is_voice_active [1] voice_start_frame_count [2] voice_end_frame_count
 *          [3] len(recent_start) [4] len(recent_end) [5..24] start ring [25..124] end ring
 * flags out per frame: bit0 voice_started, bit1 voice_ended, bit2 voice_continuing.
 * Thresholds are compared in double, as Python compares float(np.float32 p)
 * with the config's double (silero_model.py:832, :898).
 */
#define SM_STATE_INTS 125

static void ring_push(int *ring, int cap, int *len, int v) {
    if (*len < cap) { ring[(*len)++] = v; return; }
    memmove(ring, ring + 1, sizeof(int) * (cap - 1));
    ring[cap - 1] = v;
}

void sref_sm_run(const float *probs, int n, double start_p, double end_p, double start_ratio,
                 double end_ratio, int n_start, int n_end, int *state, unsigned char *flags) {
    int *rs = state + 5, *re = state + 25;
    for (int i = 0; i < n; ++i) {
        double p = (double)probs[i];
        unsigned char fl = 0;
        if (!state[0]) {
            int above = p >= start_p;
            ring_push(rs, 20, &state[3], above);
            if (above) {
                state[1] += 1;
                if (state[1] >= n_start && state[3] >= n_start) {
                    int cnt = 0;
                    for (int k = state[3] - n_start; k < state[3]; ++k) cnt += rs[k];
                    if ((double)cnt / (double)n_start >= start_ratio) {
                        state[0] = 1; state[1] = 0; state[2] = 0;
                        fl |= 1;
                    }
                }
            } else {
                state[1] = 0;
            }
        } else {
            fl |= 4;
            int below = p < end_p;
            ring_push(re, 100, &state[4], below);
            if (below) {
                state[2] += 1;
                if (state[2] >= n_end && state[4] >= n_end) {
                    int cnt = 0;
                    for (int k = state[4] - n_end; k < state[4]; ++k) cnt += re[k];
                    if ((double)cnt / (double)n_end >= end_ratio) {
                        state[0] = 0; state[2] = 0;
                        fl |= 2;
                    }
                }
            } else {
                state[2] = 0;
            }
        }
        flags[i] = fl;
    }
}

int sref_v5_blob_floats(void) { return V5_BLOB_FLOATS; }
int sref_v5_dbg_floats(void) { return V5_DBG_FLOATS; }
int sref_sm_state_ints(void) { return SM_STATE_INTS; }
int sref_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
