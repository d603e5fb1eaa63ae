// does a line fetched by another thread of the CTA serve later loads of thread 0 from L1?
#include <cstdio>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>
__global__ void probe(const unsigned* p, unsigned* wr, int hops, int mode, unsigned* out, long long* cyc) {
    __shared__ unsigned sink;
    // the chain's nodes are p[0], p[p[0]], ... ; helper threads walk it too (mode 1) so the lines are in L1
    if (mode == 1 && threadIdx.x == 32) { unsigned v = 0; for (int i = 0; i < hops; ++i) v = p[v]; sink = v; }
    if (mode == 2 && threadIdx.x == 0) { unsigned v = 0; for (int i = 0; i < hops; ++i) v = p[v]; sink = v; }
    if (mode == 3 && threadIdx.x == 32) { unsigned v = 0; for (int i = 0; i < hops; ++i) { wr[v + 1] = i; v = p[v]; } sink = v; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned v = 0; long long t0 = clock64();
        for (int i = 0; i < hops; ++i) v = p[v];
        long long t1 = clock64(); *out = v + sink; *cyc = t1 - t0;
    }
}
int main() {
    size_t n = 64 * 1024 * 1024 / 128;
    std::vector<unsigned> perm(n); for (size_t i = 0; i < n; ++i) perm[i] = i;
    std::mt19937 r(1); std::shuffle(perm.begin() + 1, perm.end(), r);
    std::vector<unsigned> h(n * 32, 0);
    for (size_t i = 0; i < n; ++i) h[(size_t)perm[i] * 32] = perm[(i + 1) % n] * 32;
    unsigned* d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    unsigned* o; long long* c; cudaMalloc(&o, 4); cudaMalloc(&c, 8);
    const char* names[] = {"cold (nobody touched)", "another thread walked it first", "thread 0 walked it first",
                           "another thread walked it and STORED into each line"};
    for (int mode = 0; mode < 4; ++mode) {
        int hops = 300;
        probe<<<1, 64>>>(d, d, hops, mode, o, c); cudaDeviceSynchronize();
        long long hc; cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
        printf("%-55s %.1f cycles/hop\n", names[mode], (double)hc / hops);
        // evict: walk somewhere else between runs is not needed, each launch starts with a cold L1
    }
}
