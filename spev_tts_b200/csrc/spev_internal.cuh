// spev_internal.cuh -- ctx layout, error plumbing and launch helpers shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "../../include/spev_b200.h"
#include "fft_core.cuh"

namespace spev {

constexpr int kTileFrames = SPEV_TILE_FRAMES;
constexpr int kTileChunks = SPEV_TILE_CHUNKS;
constexpr int kSpecLd = SPEV_SPEC_LD;
constexpr int kWarps = 16;
constexpr int kThreads = kWarps * 32;
constexpr int kStageSamples = (kTileFrames - 1) * kHop + kNfft;   // 8960

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define SPEV_CUDA(call)                                                        \
    do {                                                                       \
        cudaError_t _e = (call);                                               \
        if (_e != cudaSuccess) return ::spev::cuda_fail(_e, #call);            \
    } while (0)

#define SPEV_REQUIRE(cond, code, ...)                                          \
    do {                                                                       \
        if (!(cond)) {                                                         \
            ::spev::set_error(__VA_ARGS__);                                    \
            return (code);                                                     \
        }                                                                      \
    } while (0)

// Banded (CSR-by-band) form of the mel basis: band m covers bins [start, start+len) with
// weights w[woff .. woff+len).
struct MelBands {
    int n_mels;
    int nnz;
    const int* start;   // dev [n_mels]
    const int* len;     // dev [n_mels]
    const int* woff;    // dev [n_mels]
    const float* w;     // dev [nnz]
};

}  // namespace spev

struct spev_ctx {
    int device, sr, n_fft, hop, win, n_mels, num_sms;
    float fmin, fmax;
    // host copies
    std::vector<float> h_basis;      // [n_mels*513]
    std::vector<float> h_pinv;       // [513*n_mels]
    std::vector<float> h_window;     // [1024]
    // device constants
    float2* d_tw;        // [32*32] exp(-2 pi i k1 l / 1024)
    float* d_window;     // [1024] periodic Hann
    float* d_basis_pad;  // [n_mels, 520] zero padded (K-major B operand of the mel GEMM)
    float* d_basis_hi;   // tf32-truncated part, same shape
    float* d_basis_lo;   // residual, same shape
    float* d_pinv_t;     // [n_mels, 520]: pinv transposed (pinv_t[m][k] = pinv[k][m])
    float* d_pinv_hi;    // [528, n_mels_pad]: pinv rows (K-major B operand), tf32 hi
    float* d_pinv_lo;    //   residual
    int* d_band_start;
    int* d_band_len;
    int* d_band_woff;
    float* d_band_w;
    int band_nnz;
    int band_max_len;
    void* tma;           // opaque: tensor-map cache (gemm_tc.cu)
    int use_tc;          // mel->magnitude on frame-major input uses the tcgen05 GEMM
};
