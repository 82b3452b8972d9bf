/* A host program in plain C99 (no Python, no PyTorch) driving the C ABI end to end -- what a cgo / JNI / Rust-FFI
 * binding would do: plan tiles on the host, own every device buffer, call spev_logmel + the pYIN stages on a stream,
 * read the results back.  tests/test_gpu_host_demo.py builds it with gcc, runs it on the GPU and compares what it
 * writes with the oracle on the same deterministic signal.
 *
 *   abi_host_demo <out.bin>     writes: int64 n_frames, float logmel[n_frames*80], int32 states[n_frames]
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "spev_b200.h"

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 2; } } while (0)
#define SP(call) do { int rc_ = (call); if (rc_ != SPEV_OK) { fprintf(stderr, "%s: %d %s\n", #call, rc_, spev_last_error()); return 3; } } while (0)

/* two utterances: a 150 Hz harmonic stack with a little LCG noise, and LCG noise alone */
static void synth(float* y, int64_t n, int voiced, uint32_t seed) {
    const double two_pi = 6.283185307179586;
    double ph = 0.0;
    uint32_t s = seed;
    for (int64_t i = 0; i < n; ++i) {
        s = s * 1664525u + 1013904223u;
        const double noise = ((double)(s >> 8) / 16777216.0 - 0.5);
        double v = 0.02 * noise;
        if (voiced) {
            ph += two_pi * 150.0 / 22050.0;
            if (ph > two_pi) ph -= two_pi;
            v = 0.002 * noise + 0.25 * sin(ph) + 0.12 * sin(2.0 * ph) + 0.06 * sin(3.0 * ph);
        }
        y[i] = (float)v;
    }
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s out.bin\n", argv[0]); return 1; }
    enum { N_ITEMS = 2 };
    const int64_t lens[N_ITEMS] = {22050, 9000};
    int64_t starts[N_ITEMS], frames[N_ITEMS], frame_off[N_ITEMS + 1];
    int64_t total = 0, n_frames = 0;
    for (int i = 0; i < N_ITEMS; ++i) {
        starts[i] = total;
        total += (lens[i] + 3) / 4 * 4;                    /* item starts on 16-byte boundaries */
        frames[i] = 1 + lens[i] / 256;
        frame_off[i] = n_frames;
        n_frames += frames[i];
    }
    frame_off[N_ITEMS] = n_frames;
    float* h_y = (float*)calloc((size_t)total + 4, sizeof(float));
    synth(h_y + starts[0], lens[0], 1, 12345u);
    synth(h_y + starts[1], lens[1], 0, 777u);

    /* host-side tile plan */
    const int64_t n_tiles = spev_plan_frame_tiles(frames, starts, lens, N_ITEMS, NULL);
    if (n_tiles <= 0) { fprintf(stderr, "plan failed\n"); return 4; }
    spev_tile* h_tiles = (spev_tile*)malloc((size_t)n_tiles * sizeof(spev_tile));
    spev_plan_frame_tiles(frames, starts, lens, N_ITEMS, h_tiles);

    CK(cudaSetDevice(0));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    float *d_y, *d_mel, *d_yin, *d_logobs, *d_lunv, *d_vp, *d_f0;
    spev_tile* d_tiles;
    int64_t* d_fo;
    int32_t* d_states;
    uint8_t* d_flag;
    CK(cudaMalloc((void**)&d_y, ((size_t)total + 4) * sizeof(float)));
    CK(cudaMalloc((void**)&d_tiles, (size_t)n_tiles * sizeof(spev_tile)));
    CK(cudaMalloc((void**)&d_fo, (N_ITEMS + 1) * sizeof(int64_t)));
    CK(cudaMalloc((void**)&d_mel, (size_t)n_frames * 80 * sizeof(float)));
    CK(cudaMemcpyAsync(d_y, h_y, ((size_t)total + 4) * sizeof(float), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_tiles, h_tiles, (size_t)n_tiles * sizeof(spev_tile), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_fo, frame_off, (N_ITEMS + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));

    spev_batch b;
    memset(&b, 0, sizeof b);
    b.n_items = N_ITEMS; b.n_ftiles = (int32_t)n_tiles; b.n_ctiles = 0; b.n_frames = n_frames;
    b.frame_off = d_fo; b.ftiles = d_tiles; b.ctiles = NULL;

    spev_ctx* ctx = NULL;
    SP(spev_create(&ctx, 0, 22050, 1024, 256, 1024, 80, 0.0f, 0.0f));
    SP(spev_logmel(ctx, &b, d_y, d_mel, 1, 1e-5f, -10.0f, 2.0f, st));

    spev_pyin* py = NULL;
    int n_bins = 0, n_lags = 0;
    SP(spev_pyin_create(&py, 0, 22050, 256, 60.0f, 500.0f, NULL));
    SP(spev_pyin_info(py, &n_bins, NULL, NULL, &n_lags));
    const size_t ws_bytes = spev_pyin_decode_workspace_bytes(py, n_frames);
    void* d_ws;
    CK(cudaMalloc((void**)&d_yin, (size_t)n_frames * n_lags * sizeof(float)));
    CK(cudaMalloc((void**)&d_logobs, (size_t)n_frames * n_bins * sizeof(float)));
    CK(cudaMalloc((void**)&d_lunv, (size_t)n_frames * sizeof(float)));
    CK(cudaMalloc((void**)&d_vp, (size_t)n_frames * sizeof(float)));
    CK(cudaMalloc((void**)&d_f0, (size_t)n_frames * sizeof(float)));
    CK(cudaMalloc((void**)&d_states, (size_t)n_frames * sizeof(int32_t)));
    CK(cudaMalloc((void**)&d_flag, (size_t)n_frames));
    CK(cudaMalloc(&d_ws, ws_bytes));
    SP(spev_pyin_cmnd(py, &b, d_y, d_yin, st));
    SP(spev_pyin_observe(py, d_yin, n_frames, d_logobs, d_lunv, d_vp, st));
    SP(spev_pyin_decode(py, d_logobs, d_lunv, d_fo, N_ITEMS, n_frames, d_states, d_f0, d_flag, d_ws, ws_bytes, st));

    float* h_mel = (float*)malloc((size_t)n_frames * 80 * sizeof(float));
    int32_t* h_states = (int32_t*)malloc((size_t)n_frames * sizeof(int32_t));
    CK(cudaMemcpyAsync(h_mel, d_mel, (size_t)n_frames * 80 * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_states, d_states, (size_t)n_frames * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));

    FILE* f = fopen(argv[1], "wb");
    if (!f) { perror("fopen"); return 5; }
    fwrite(&n_frames, sizeof n_frames, 1, f);
    fwrite(h_mel, sizeof(float), (size_t)n_frames * 80, f);
    fwrite(h_states, sizeof(int32_t), (size_t)n_frames, f);
    fwrite(h_y, sizeof(float), (size_t)total, f);              /* the signal, so the checker needs no libm twin */
    fclose(f);
    spev_pyin_destroy(py);
    spev_destroy(ctx);
    printf("ok: %lld frames, %lld tiles, abi %d\n", (long long)n_frames, (long long)n_tiles, spev_abi_version());
    return 0;
}
