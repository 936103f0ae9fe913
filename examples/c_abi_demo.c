/*
 * c_abi_demo.c -- the engine from plain C, through include/cutter_vad_b200.h only (no Python, no torch).
 *
 *   python tools/export_weights.py v5 /tmp/silero_v5.f32          # canonical float32 blob from the reference's .onnx
 *   gcc -O2 -Iinclude examples/c_abi_demo.c -Lcutter-vad_b200 -lcvad_b200 -Wl,-rpath,$PWD/cutter-vad_b200 -lm -o /tmp/c_abi_demo
 *   /tmp/c_abi_demo /tmp/silero_v5.f32
 *
 * Opens 64 streams on a stream feeder (the service-mode host path), pushes 30 ms int16 messages of a gated tone
 * for two seconds of stream time, steps once per message interval and prints the voice start / end events --
 * what the reference's websocket server does per client with one VADWrapper each
 * (websocket_service/server/vad_websocket_server.py:326-369, thresholds :565-572).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "cutter_vad_b200.h"

#define N_STREAMS 64
#define MSG 480 /* 30 ms at 16 kHz */

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s weights.f32\n", argv[0]); return 2; }
    FILE *fh = fopen(argv[1], "rb");
    if (!fh) { perror(argv[1]); return 2; }
    float *w = (float *)malloc(sizeof(float) * CVAD_V5_WEIGHT_FLOATS);
    if (fread(w, sizeof(float), CVAD_V5_WEIGHT_FLOATS, fh) != CVAD_V5_WEIGHT_FLOATS) { fprintf(stderr, "short weight file\n"); return 2; }
    fclose(fh);

    if (cvad_device_count() < 1) { fprintf(stderr, "no sm_100 device: %s\n", cvad_last_error(NULL)); return 3; }
    cvad_engine *eng = NULL;
    if (cvad_create(w, CVAD_V5_WEIGHT_FLOATS, CVAD_MODEL_V5, N_STREAMS, 0, &eng)) { fprintf(stderr, "create: %s\n", cvad_last_error(NULL)); return 1; }
    free(w);
    /* websocket defaults: 0.4 / 0.3 / 6 / 12, denoising on */
    if (cvad_configure(eng, 0, NULL, 0.4, 0.3, 6, 12, 1)) { fprintf(stderr, "configure: %s\n", cvad_last_error(eng)); return 1; }

    cvad_feeder *f = NULL;
    if (cvad_feeder_create(eng, 0, CVAD_PCM_S16_32767, MSG, MSG, 16000, 8, &f)) { fprintf(stderr, "feeder: %s\n", cvad_last_error(eng)); return 1; }
    for (int s = 0; s < N_STREAMS; ++s)
        if (cvad_feeder_open(f, s, 0, CVAD_PAYLOAD_SEGMENTS, 0.4, 1)) { fprintf(stderr, "open: %s\n", cvad_feeder_last_error(f)); return 1; }

    int16_t msg[MSG];
    long starts = 0, ends = 0, frames = 0;
    double seg_samples = 0;
    for (int m = 0; m < 67; ++m) {                       /* ~2 s */
        for (int s = 0; s < N_STREAMS; ++s) {
            /* a harmonic "voice" between 0.3 s and 1.2 s (stream-dependent pitch), faint noise elsewhere */
            for (int k = 0; k < MSG; ++k) {
                const double t = (m * MSG + k) / 16000.0, f0 = 120.0 + 2.0 * s;
                const int voiced = t > 0.3 && t < 1.2;
                double x = 0.002 * ((rand() % 2001) / 1000.0 - 1.0);
                if (voiced) x += 0.7 * (0.4 * sin(6.2831853 * f0 * t) + 0.3 * sin(6.2831853 * 2 * f0 * t) + 0.2 * sin(6.2831853 * 4 * f0 * t)) *
                                 (1.0 + 0.5 * sin(6.2831853 * 4 * t));
                msg[k] = (int16_t)lrint(fmax(-1.0, fmin(1.0, x)) * 32767.0);
            }
            if (cvad_feeder_push(f, s, msg, MSG)) { fprintf(stderr, "push: %s\n", cvad_feeder_last_error(f)); return 1; }
        }
        cvad_feeder_result r;
        if (cvad_feeder_step(f, &r)) { fprintf(stderr, "step: %s\n", cvad_feeder_last_error(f)); return 1; }
        frames += r.n_frames_total;
        for (int i = 0; i < r.n_deliveries; ++i) {
            const cvad_delivery *d = &r.deliveries[i];
            if (d->flags & CVAD_FLAG_STARTED) ++starts;
            if (d->flags & CVAD_FLAG_ENDED) { ++ends; seg_samples += (double)d->segment_len; }
            if (d->slot == 0 && (d->flags & (CVAD_FLAG_STARTED | CVAD_FLAG_ENDED)))
                printf("stream 0: %s at message %d (p = %.3f)\n", (d->flags & CVAD_FLAG_STARTED) ? "VOICE_START" : "VOICE_END", m, d->prob);
        }
    }
    printf("frames %ld, voice starts %ld, ends %ld, mean segment %.0f ms, kernel launches %lld\n", frames, starts, ends,
           ends ? seg_samples / ends / 16.0 : 0.0, (long long)cvad_launch_count(eng));
    cvad_feeder_destroy(f);
    cvad_destroy(eng);
    return (starts == N_STREAMS && ends == N_STREAMS) ? 0 : 4;
}
