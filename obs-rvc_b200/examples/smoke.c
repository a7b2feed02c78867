/* Plain-C smoke test of the drop-in boundary: links librvc_b200.so, loads the three models of a data directory and
 * pushes one 160 ms window through rvc_infer (the call rvc-rpc/src/main.rs:93 makes), then three blocks through the
 * device-resident streaming loop.  Build: make -C obs-rvc_b200 smoke ; run: ./smoke <data_dir> <model.rvcw>        */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/rvc_b200.h"

#define CHECK(call) do { int rc_ = (call); if (rc_ != RVC_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, ctx ? rvc_last_error(ctx) : rvc_last_create_error()); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s <data_dir> <model.rvcw>\n", argv[0]); return 2; }
    rvc_ctx* ctx = NULL;
    CHECK(rvc_create(argv[1], NULL, &ctx));
    CHECK(rvc_load_contentvec(ctx, RVC_MODEL_V2));
    CHECK(rvc_load_f0(ctx, RVC_PITCH_RMVPE));
    CHECK(rvc_load_model(ctx, argv[2]));
    /* the reference's 160 ms geometry: 35840 samples, advance 2560, skip_head 200, return_length 21 */
    enum { N = 35840, SF = 2560, SKIP = 200, RET = 21 };
    float* pcm = (float*)malloc(sizeof(float) * N);
    float* out = (float*)malloc(sizeof(float) * RET * 480);
    for (int i = 0; i < N; ++i) pcm[i] = 0.3f * sinf(6.2831853f * 220.0f * (float)i / 16000.0f) + 0.1f * sinf(6.2831853f * 3300.0f * (float)i / 16000.0f);
    size_t n = 0;
    CHECK(rvc_infer(ctx, pcm, N, SF, 12, SKIP, RET, out, RET * 480, &n));
    double e = 0.0;
    for (size_t i = 0; i < n; ++i) e += (double)out[i] * out[i];
    printf("rvc_infer: %zu samples, rms %.4f\n", n, sqrt(e / (double)n));
    if (n != RET * 400 || !(e > 0.0) || isnan(e)) { fprintf(stderr, "unexpected output\n"); return 1; }
    rvc_stream_config sc;
    rvc_stream_config_default(&sc);
    sc.sample_length = 0.16; sc.crossfade_length = 0.04;
    uint32_t frame = 0;
    CHECK(rvc_stream_open(ctx, &sc, &frame));
    float* blk = (float*)malloc(sizeof(float) * frame);
    float* res = (float*)malloc(sizeof(float) * frame);
    for (int f = 0; f < 3; ++f) {
        for (uint32_t i = 0; i < frame; ++i) blk[i] = 0.3f * sinf(6.2831853f * 220.0f * (float)(f * frame + i) / 48000.0f);
        uint32_t off = 0;
        CHECK(rvc_process_frame(ctx, blk, res, &off));
        printf("rvc_process_frame %d: %u samples, sola offset %u\n", f, frame, off);
    }
    CHECK(rvc_stream_close(ctx));
    uint64_t launches = 0;
    CHECK(rvc_kernel_launches(ctx, &launches));
    printf("kernel launches: %llu\nSMOKE_C ok\n", (unsigned long long)launches);
    rvc_destroy(ctx);
    free(pcm); free(out); free(blk); free(res);
    return 0;
}
