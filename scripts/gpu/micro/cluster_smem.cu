// Micro-benchmark: latency of a CTA's OWN shared memory inside a thread-block cluster, per cluster rank.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o cluster_smem cluster_smem.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

struct Sh {
    int next[1024];
    double v[1024];
    int peer;
};

__global__ void k(long long* out, int n, int use_map, int fence_first)
{
    extern __shared__ __align__(16) unsigned char raw[];
    Sh& sm = *reinterpret_cast<Sh*>(raw);
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm.next[i] = (i * 37 + 11) & 1023, sm.v[i] = i;
    if (use_map && threadIdx.x == 0) {  // make the kernel take the address of its shared memory for the cluster (as map_shared_rank does)
        Sh* o = cluster.map_shared_rank(&sm, (rank + 1) % cluster.num_blocks());
        o->peer = rank;
    }
    cluster.sync();
    if (threadIdx.x == 0) {
        if (fence_first == 1) asm volatile("fence.acq_rel.cluster;" ::: "memory");
        if (fence_first == 2) {  // a volatile poll of a flag in own shared memory written by the neighbour, then the fence
            while (*reinterpret_cast<volatile int*>(&sm.peer) < 0) {
            }
            asm volatile("fence.acq_rel.cluster;" ::: "memory");
        }
        // (a) compiler-addressed accesses
        int p = 0;
        double acc = 0;
        long long t0 = clock64();
        for (int i = 0; i < n; i++) {
            p = sm.next[p];
            acc += sm.v[p];
        }
        long long t1 = clock64();
        // (b) explicit ld.shared with the CTA-local 32-bit address
        unsigned base = (unsigned)__cvta_generic_to_shared(raw);
        unsigned q = 0;
        long long t2 = clock64();
        for (int i = 0; i < n; i++) {
            unsigned a = base + q * 4;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(q) : "r"(a));
        }
        long long t3 = clock64();
        out[rank * 4 + 0] = t1 - t0;
        out[rank * 4 + 1] = t3 - t2;
        out[rank * 4 + 2] = (long long)base;
        out[rank * 4 + 3] = (long long)(p + q + (int)acc);
    }
    cluster.sync();
}

int main()
{
    long long* d;
    cudaMalloc(&d, 8 * 4 * 8);
    const int n = 20000;
    for (int fence_first = 0; fence_first < 3; fence_first++)
    for (int use_map = 1; use_map < 2; use_map++)
        for (int X : {1, 8}) {
            cudaMemset(d, 0, 8 * 4 * 8);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(X);
            cfg.blockDim = dim3(128);
            cfg.dynamicSmemBytes = sizeof(Sh);
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = X, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, k, d, n, use_map, fence_first);
            cudaDeviceSynchronize();
            long long h[32];
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            for (int r = 0; r < X; r++)
                printf("fence %d map %d cluster %d rank %d: compiler-addressed %.1f cycles/iteration, explicit ld.shared %.1f cycles/load, cvta base 0x%llx (%s)\n",
                       fence_first, use_map, X, r, (double)h[r * 4] / n, (double)h[r * 4 + 1] / n, h[r * 4 + 2], cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
