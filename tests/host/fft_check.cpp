// Host-side check of the register-level DFT building blocks in fft_core.cuh (compiled with
// g++; the same templates are compiled by nvcc for the device).  Prints max abs errors vs an
// O(N^2) double DFT; exit code 0 iff all below tolerance.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <complex>
#include <vector>
#include "fft_core.cuh"

using namespace spev;
template <int DIR> static double check32(unsigned seed) {
    srand(seed);
    float2 v[32]; std::complex<double> x[32];
    for (int i = 0; i < 32; ++i) { v[i].x = rand() / (float)RAND_MAX - 0.5f; v[i].y = rand() / (float)RAND_MAX - 0.5f; x[i] = {v[i].x, v[i].y}; }
    dft32<DIR>(v);
    double err = 0;
    for (int k = 0; k < 32; ++k) {
        std::complex<double> s = 0;
        for (int n = 0; n < 32; ++n) s += x[n] * std::polar(1.0, DIR * 2 * M_PI * n * k / 32.0);
        err = fmax(err, std::abs(s - std::complex<double>(v[k].x, v[k].y)));
    }
    return err;
}
template <int DIR> static double check8(unsigned seed) {
    srand(seed);
    float2 v[8]; std::complex<double> x[8];
    for (int i = 0; i < 8; ++i) { v[i].x = rand() / (float)RAND_MAX - 0.5f; v[i].y = rand() / (float)RAND_MAX - 0.5f; x[i] = {v[i].x, v[i].y}; }
    dft8<DIR>(v);
    double err = 0;
    for (int k = 0; k < 8; ++k) {
        std::complex<double> s = 0;
        for (int n = 0; n < 8; ++n) s += x[n] * std::polar(1.0, DIR * 2 * M_PI * n * k / 8.0);
        err = fmax(err, std::abs(s - std::complex<double>(v[k].x, v[k].y)));
    }
    return err;
}
int main() {
    double e = 0;
    for (unsigned s = 1; s < 20; ++s) { e = fmax(e, check32<-1>(s)); e = fmax(e, check32<1>(s)); e = fmax(e, check8<-1>(s)); e = fmax(e, check8<1>(s)); }
    printf("max_abs_err %.3e\n", e);
    return e < 5e-6 ? 0 : 1;
}
