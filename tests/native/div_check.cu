// Stand-alone check (built by tests/test_gpu_native.py with nvcc): SharedDivisor::div(a) must give the same
// bits as the IEEE '/' for every (a, b), including zeros, subnormals, infinities, NaNs and huge/tiny exponents.
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <random>
#include <cuda_runtime.h>
#include "cplb_device.cuh"

__global__ void check(const double* a, const double* b, int n, unsigned long long* bad, double* first_bad)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const cplb::SharedDivisor d(b[i]);
    double q = d.div(a[i]);
    double r = a[i] / b[i];
    bool same = (q != q && r != r) || (__double_as_longlong(q) == __double_as_longlong(r));
    if (!same) {
        if (atomicAdd(bad, 1ull) == 0) { first_bad[0] = a[i]; first_bad[1] = b[i]; first_bad[2] = q; first_bad[3] = r; }
    }
}

int main()
{
    const int n = 1 << 24;
    std::vector<double> a(n), b(n);
    std::mt19937_64 rng(12345);
    const double specials[] = {0.0, -0.0, 1.0, -1.0, 4.9e-324, 2.2250738585072014e-308, 1.7976931348623157e308,
                               INFINITY, -INFINITY, NAN, 1e-200, 1e200, 3.0, 1.0 / 3.0, 1e-310, 5e-160, 2e160};
    const int ns = sizeof(specials) / sizeof(double);
    for (int i = 0; i < n; i++) {
        uint64_t u = rng(), v = rng();
        double x, y;
        if (i < ns * ns) { x = specials[i / ns]; y = specials[i % ns]; }
        else if (i % 4 == 0) { memcpy(&x, &u, 8); memcpy(&y, &v, 8); }            // arbitrary bit patterns
        else if (i % 4 == 1) { x = std::ldexp((double)(u >> 11) / 9007199254740992.0 + 0.5, (int)(v % 600) - 300); y = std::ldexp((double)(v >> 11) / 9007199254740992.0 + 0.5, (int)(u % 600) - 300); }
        else { x = ((double)(int64_t)u) * 1e-12; y = std::fabs((double)(int64_t)v) * 1e-15 + 1e-3; }  // force-like magnitudes
        a[i] = x; b[i] = y;
    }
    double *da, *db, *dfirst; unsigned long long* dbad;
    cudaMalloc(&da, n * 8); cudaMalloc(&db, n * 8); cudaMalloc(&dbad, 8); cudaMalloc(&dfirst, 32);
    cudaMemcpy(da, a.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), n * 8, cudaMemcpyHostToDevice);
    cudaMemset(dbad, 0, 8);
    check<<<(n + 255) / 256, 256>>>(da, db, n, dbad, dfirst);
    unsigned long long bad = 0; double fb[4];
    if (cudaMemcpy(&bad, dbad, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("CUDA error\n"); return 2; }
    cudaMemcpy(fb, dfirst, 32, cudaMemcpyDeviceToHost);
    printf("checked %d pairs, mismatches %llu\n", n, bad);
    if (bad) printf("first: a=%a b=%a fast=%a ieee=%a\n", fb[0], fb[1], fb[2], fb[3]);
    return bad ? 1 : 0;
}
