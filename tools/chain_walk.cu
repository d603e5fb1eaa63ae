// Micro-benchmark behind DESIGN.md's "cost-propagation" notes: one thread walks a linked chain of 32-byte
// records (the shape of tree_t::updateCost on a single-child chain): plain pointer chase against a walk that
// requests the record three hops ahead through a per-record hint and keeps four records in flight.
//   nvcc -arch=sm_100a -O3 -o chain_walk tools/chain_walk.cu && ./chain_walk
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>

struct __align__(32) Rec
{
    uint32_t next, hint;   // successor, successor three hops ahead
    double cost;
    uint32_t pad[4];
};

__global__ void walk_plain(Rec* r, uint32_t start, int steps, long long* cycles, double* sink, uint32_t* out)
{
    uint32_t v = start;
    double acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < steps; ++i)
    {
        Rec* p = r + v;
        double c = p->cost;
        uint32_t nx = p->next;
        p->cost = 1.5 + c - 0.25;     // the propagation's read-modify-write
        out[i] = nx;                  // the updated_nodes list
        acc += c;
        v = nx;
    }
    *cycles = clock64() - t0;
    *sink = acc + v;
}

struct Slot
{
    uint32_t id;
    uint32_t next, hint;
    double cost;
};

__device__ __forceinline__ void fetch(Rec* r, uint32_t id, Slot& s)
{
    s.id = id;
    Rec* p = r + id;
    s.cost = p->cost;
    s.next = p->next;
    s.hint = p->hint;
}

__global__ void walk_ahead(Rec* r, uint32_t start, int steps, long long* cycles, double* sink, uint32_t* out)
{
    Slot a, b, c, d;
    fetch(r, start, a);
    b.id = c.id = d.id = 0xFFFFFFFFu;
    double acc = 0;
    int i = 0;
    long long t0 = clock64();
    auto level = [&](Slot& cur, Slot& n1, Slot& fill) {
        r[cur.id].cost = 1.5 + cur.cost - 0.25;
        out[i] = cur.next;
        acc += cur.cost;
        fetch(r, cur.hint, fill);                        // three hops ahead
        if (n1.id != cur.next) fetch(r, cur.next, n1);   // not requested (start of the walk)
        ++i;
    };
    while (i + 4 <= steps)
    {
        level(a, b, d);
        level(b, c, a);
        level(c, d, b);
        level(d, a, c);
    }
    *cycles = clock64() - t0;
    *sink = acc + a.id;
}

int main()
{
    for (size_t n : {size_t(1) << 15, size_t(1) << 22, size_t(1) << 25})   // 1 MB, 128 MB, 1 GB of records
    {
        std::vector<uint32_t> perm(n);
        for (size_t i = 0; i < n; ++i) perm[i] = (uint32_t)i;
        std::mt19937_64 rng(1);
        std::shuffle(perm.begin(), perm.end(), rng);
        std::vector<Rec> h(n);
        for (size_t i = 0; i < n; ++i)
        {
            Rec& x = h[perm[i]];
            x.next = perm[(i + 1) % n];
            x.hint = perm[(i + 3) % n];
            x.cost = 1.0;
        }
        Rec* d;
        long long* cyc;
        double* sink;
        uint32_t* out;
        const int steps = 20000;
        cudaMalloc(&d, n * sizeof(Rec));
        cudaMalloc(&cyc, 8);
        cudaMalloc(&sink, 8);
        cudaMalloc(&out, steps * 4);
        cudaMemcpy(d, h.data(), n * sizeof(Rec), cudaMemcpyHostToDevice);
        long long c1 = 0, c2 = 0;
        for (int rep = 0; rep < 2; ++rep)
        {
            walk_plain<<<1, 1>>>(d, perm[0], steps, cyc, sink, out);
            cudaMemcpy(&c1, cyc, 8, cudaMemcpyDeviceToHost);
            walk_ahead<<<1, 1>>>(d, perm[steps + 7], steps, cyc, sink, out);
            cudaMemcpy(&c2, cyc, 8, cudaMemcpyDeviceToHost);
        }
        printf("%8.0f MB of records: plain chase %6.0f cycles/hop, three hops ahead %6.0f cycles/hop\n",
               n * 32.0 / 1e6, (double)c1 / steps, (double)c2 / steps);
        cudaFree(d); cudaFree(cyc); cudaFree(sink); cudaFree(out);
    }
    return 0;
}
