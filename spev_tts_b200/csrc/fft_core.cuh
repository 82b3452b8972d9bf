// fft_core.cuh -- warp-level 1024-point complex FFT for sm_100a.
//
// One warp transforms 1024 complex points as 32 x 32 (Cooley-Tukey, decimation in
// frequency over the strided index): every lane runs a fully unrolled 32-point DFT in
// registers (4 x 8 split, compile-time twiddles -> immediate-form FFMA), multiplies by the
// inter-stage twiddles w1024^(lane*k1) read conflict-free from a [32][32] shared-memory
// table, transposes through a warp-private padded shared-memory tile, and runs the second
// 32-point DFT.  There is exactly ONE shared-memory exchange per 1024 complex points.
//
// Two real 1024-sample frames are packed into one complex transform (frame a -> real part,
// frame b -> imaginary part).  The Hermitian split needs Z[k] and Z[1024-k]; with the
// output layout "lane l, register r holds Z[l + 32 r]" the partner of (l, r) is
// ((32-l)&31, 31-r) (lane 0: (0, (32-r)&31)), i.e. one warp shuffle per value and no
// second exchange.  The inverse path packs two Hermitian spectra the same way.
//
// Replaces: numpy/scipy pocketfft under librosa.stft / librosa.istft, reached from
// /root/reference/spev_real_metrics.py:363 (melspectrogram) and :730-733 (mel_to_audio ->
// griffinlim -> istft/stft).
#pragma once
#include <cuda_runtime.h>
#include <cmath>
#include <type_traits>

#if defined(__CUDACC__)
#define SPEV_HD __host__ __device__ __forceinline__
#define SPEV_D __device__ __forceinline__
#else
#define SPEV_HD inline
#define SPEV_D inline
#endif

namespace spev {

constexpr int kNfft = 1024;
constexpr int kHop = 256;
constexpr int kBins = 513;
constexpr int kXPitch = 33;                       // float2 per row of the transpose tile
constexpr int kXWords = 32 * kXPitch * 2;         // 2112 32-bit words
// Warp-private region stride and second-frame offset of the |X|^2 slots (K1): in 16-byte units they
// are == 2 and == 1 (mod 8), so that lane f's float4 reads land in 16-byte bank group f % 8 --
// conflict-free for every quarter-warp (LDS.128) and for scalar accesses alike.
constexpr int kWarpRegionWords = kXWords + 8;     // 2120
constexpr int kPSlot = 516;                       // frame b's |X|^2 slot starts here (513 bins + 3 zero words)

// cos/sin(2*pi*e/32), e = 0..31 (float-rounded from float64)
constexpr float kCos32[32] = {
    1.0f, 0.980785280403230449f, 0.923879532511286756f, 0.831469612302545237f,
    0.707106781186547524f, 0.555570233019602225f, 0.382683432365089772f, 0.195090322016128268f,
    0.0f, -0.195090322016128268f, -0.382683432365089772f, -0.555570233019602225f,
    -0.707106781186547524f, -0.831469612302545237f, -0.923879532511286756f, -0.980785280403230449f,
    -1.0f, -0.980785280403230449f, -0.923879532511286756f, -0.831469612302545237f,
    -0.707106781186547524f, -0.555570233019602225f, -0.382683432365089772f, -0.195090322016128268f,
    0.0f, 0.195090322016128268f, 0.382683432365089772f, 0.555570233019602225f,
    0.707106781186547524f, 0.831469612302545237f, 0.923879532511286756f, 0.980785280403230449f};
constexpr float kSin32[32] = {
    0.0f, 0.195090322016128268f, 0.382683432365089772f, 0.555570233019602225f,
    0.707106781186547524f, 0.831469612302545237f, 0.923879532511286756f, 0.980785280403230449f,
    1.0f, 0.980785280403230449f, 0.923879532511286756f, 0.831469612302545237f,
    0.707106781186547524f, 0.555570233019602225f, 0.382683432365089772f, 0.195090322016128268f,
    0.0f, -0.195090322016128268f, -0.382683432365089772f, -0.555570233019602225f,
    -0.707106781186547524f, -0.831469612302545237f, -0.923879532511286756f, -0.980785280403230449f,
    -1.0f, -0.980785280403230449f, -0.923879532511286756f, -0.831469612302545237f,
    -0.707106781186547524f, -0.555570233019602225f, -0.382683432365089772f, -0.195090322016128268f};

template <int I, int N, class F>
SPEV_HD void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

SPEV_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
SPEV_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
SPEV_HD float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// multiply by DIR * i   (DIR = -1: forward transform e^{-i...}; DIR = +1: inverse)
template <int DIR>
SPEV_HD float2 mul_i(float2 a) {
    if constexpr (DIR > 0) return make_float2(-a.y, a.x);
    else return make_float2(a.y, -a.x);
}

// multiply by exp(DIR * 2*pi*i * E / 32), E a compile-time constant
template <int DIR, int E>
SPEV_HD float2 tw32(float2 a) {
    constexpr int e = E & 31;
    if constexpr (e == 0) return a;
    else if constexpr (e == 8) return mul_i<DIR>(a);
    else if constexpr (e == 16) return make_float2(-a.x, -a.y);
    else if constexpr (e == 24) return mul_i<-DIR>(a);
    else {
        constexpr float c = kCos32[e];
        constexpr float s = (DIR > 0 ? 1.0f : -1.0f) * kSin32[e];
        return make_float2(fmaf(a.x, c, -a.y * s), fmaf(a.x, s, a.y * c));
    }
}

template <int DIR>
SPEV_HD void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2);
    const float2 t2 = cadd(a1, a3), t3 = mul_i<DIR>(csub(a1, a3));
    a0 = cadd(t0, t2); a2 = csub(t0, t2);
    a1 = cadd(t1, t3); a3 = csub(t1, t3);
}

// 8-point DFT, natural order in and out, on a[0..7]
template <int DIR>
SPEV_HD void dft8(float2 (&a)[8]) {
    dft4<DIR>(a[0], a[2], a[4], a[6]);   // E0..E3 in a0,a2,a4,a6
    dft4<DIR>(a[1], a[3], a[5], a[7]);   // O0..O3 in a1,a3,a5,a7
    const float2 o0 = a[1];
    const float2 o1 = tw32<DIR, 4>(a[3]);
    const float2 o2 = tw32<DIR, 8>(a[5]);
    const float2 o3 = tw32<DIR, 12>(a[7]);
    const float2 e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6];
    a[0] = cadd(e0, o0); a[4] = csub(e0, o0);
    a[1] = cadd(e1, o1); a[5] = csub(e1, o1);
    a[2] = cadd(e2, o2); a[6] = csub(e2, o2);
    a[3] = cadd(e3, o3); a[7] = csub(e3, o3);
}

// 32-point DFT in registers, natural order in and out.
// n = 8*n1 + n2, k = k1 + 4*k2:  X[k1+4k2] = sum_n2 w8^(n2 k2) w32^(n2 k1) sum_n1 w4^(n1 k1) x[8n1+n2]
template <int DIR>
SPEV_HD void dft32(float2 (&v)[32]) {
    static_for<0, 8>([&](auto n2c) {
        constexpr int n2 = decltype(n2c)::value;
        dft4<DIR>(v[n2], v[8 + n2], v[16 + n2], v[24 + n2]);   // v[8*k1 + n2] = y[k1][n2]
    });
    float2 o[32];
    static_for<0, 4>([&](auto k1c) {
        constexpr int k1 = decltype(k1c)::value;
        float2 a[8];
        static_for<0, 8>([&](auto n2c) {
            constexpr int n2 = decltype(n2c)::value;
            a[n2] = tw32<DIR, k1 * n2>(v[8 * k1 + n2]);
        });
        dft8<DIR>(a);
        static_for<0, 8>([&](auto k2c) {
            constexpr int k2 = decltype(k2c)::value;
            o[k1 + 4 * k2] = a[k2];
        });
    });
    static_for<0, 32>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        v[i] = o[i];
    });
}

#if defined(__CUDACC__)
// Warp-level 1024-point complex FFT.
//   in : v[j]  = x[32*j + lane]          (j = 0..31)
//   out: v[k2] = X[lane + 32*k2]         (k2 = 0..31), unnormalised, sign = DIR
//   xb : warp-private shared tile of 32*33 float2
//   tw : shared table tw[k1*32 + l] = exp(-2*pi*i * k1*l / 1024)   (forward sign)
template <int DIR>
SPEV_D void warp_fft1024(float2 (&v)[32], float2* xb, const float2* tw, int lane) {
    dft32<DIR>(v);
    static_for<1, 32>([&](auto k1c) {
        constexpr int k1 = decltype(k1c)::value;
        float2 w = tw[k1 * 32 + lane];
        if constexpr (DIR > 0) w.y = -w.y;
        v[k1] = cmul(v[k1], w);
    });
    __syncwarp();
    static_for<0, 32>([&](auto k1c) {
        constexpr int k1 = decltype(k1c)::value;
        xb[k1 * kXPitch + lane] = v[k1];
    });
    __syncwarp();
    static_for<0, 32>([&](auto n2c) {
        constexpr int n2 = decltype(n2c)::value;
        v[n2] = xb[lane * kXPitch + n2];
    });
    dft32<DIR>(v);
}

// After a forward transform of z = a + i b (a, b real frames): fetch Z[1024-k] for the 16
// bins k = lane + 32*k2 (k2 = 0..15) this lane owns.
SPEV_D void fetch_mirror(const float2 (&v)[32], float2 (&p)[16], int lane) {
    const int pl = (32 - lane) & 31;
    static_for<0, 16>([&](auto kc) {
        constexpr int k2 = decltype(kc)::value;
        const float2 s = v[31 - k2];
        p[k2].x = __shfl_sync(0xffffffffu, s.x, pl);
        p[k2].y = __shfl_sync(0xffffffffu, s.y, pl);
    });
    if (lane == 0) {
        p[0] = v[0];
        static_for<1, 16>([&](auto kc) {
            constexpr int k2 = decltype(kc)::value;
            p[k2] = v[32 - k2];
        });
    }
}

#endif  // __CUDACC__

// Hermitian split when the 1/2 has been folded into the window (exact: power of two)
SPEV_HD void split_pair_prescaled(float2 z, float2 zm, float2& xa, float2& xb) {
    xa = make_float2(z.x + zm.x, z.y - zm.y);
    xb = make_float2(z.y + zm.y, zm.x - z.x);
}

// Hermitian split: Z[k] = (a,b), Z[N-k] = (c,d)  ->  Xa[k], Xb[k]
SPEV_HD void split_pair(float2 z, float2 zm, float2& xa, float2& xb) {
    xa = make_float2(0.5f * (z.x + zm.x), 0.5f * (z.y - zm.y));
    xb = make_float2(0.5f * (z.y + zm.y), 0.5f * (zm.x - z.x));
}

}  // namespace spev
