// pointer-chase latency probe (development aid): one thread walks a random cycle
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>
__global__ void chase(const unsigned* p, int steps, unsigned* out, long long* cyc) {
    unsigned v = 0; long long t0 = clock64();
    for (int i = 0; i < steps; ++i) v = p[v];
    long long t1 = clock64(); *out = v; *cyc = t1 - t0;
}
int main() {
    for (size_t mb : {1, 8, 32, 96, 400, 2000}) {
        size_t n = mb * 1024 * 1024 / 128;      // one entry per 128-byte line
        std::vector<unsigned> perm(n); for (size_t i = 0; i < n; ++i) perm[i] = i;
        std::mt19937 r(1); std::shuffle(perm.begin() + 1, perm.end(), r);
        std::vector<unsigned> h(n * 32, 0);
        for (size_t i = 0; i < n; ++i) h[(size_t)perm[i] * 32] = perm[(i + 1) % n] * 32;
        unsigned* d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
        unsigned* o; long long* c; cudaMalloc(&o, 4); cudaMalloc(&c, 8);
        int steps = 20000;
        chase<<<1, 1>>>(d, steps, o, c); cudaDeviceSynchronize();
        chase<<<1, 1>>>(d, steps, o, c); cudaDeviceSynchronize();
        long long hc; cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
        printf("footprint %zu MB: %.1f cycles/hop\n", mb, (double)hc / steps);
        cudaFree(d); cudaFree(o); cudaFree(c);
    }
}
