// libarrow_probes.so -- measurement-only microbenchmarks (never loaded by the product path).
//
// Question (VERDICT r1 item 5 / north_star "TMA-stages the k-column dense panel into shared memory"): can an X panel
// that is k-sliced into the distributed shared memory of a thread-block cluster feed the SpMM's row gathers faster
// than L2 does?  The gather of the arrow SpMM reads, per non-zero, one contiguous run of the X row (512 B at k = 128
// from L2; 128 B per k-slice of 32 from shared memory).  Three gather loops with the same lane layout as
// k_spmm_tiles (8 lanes x float4 per 128 B run, UNROLL independent gathers in flight) differ only in where the run
// comes from:
//   mode 0  global memory, panel resident in L2           (what k_spmm_tiles does today)
//   mode 1  the CTA's own shared memory                   (upper bound: a panel slice that fits one SM)
//   mode 2  distributed shared memory of an 8-CTA cluster (the k-sliced cluster panel: 7/8 of the runs are remote)
// Each returns the time of `iters` gathers per lane group; the caller converts to bytes/s and B/clk/SM.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include <stdio.h>

namespace cg = cooperative_groups;

namespace {

constexpr int THREADS = 256;
constexpr int UNROLL = 8;

__device__ __forceinline__ uint32_t lcg(uint32_t &s) {
    s = s * 1664525u + 1013904223u;
    return s >> 8;
}

// mode 0: rows of `row_f4` float4 (128 B -> 8, 512 B -> 32) out of a `panel_rows`-row panel in global memory
template <int G, int VPL>
__global__ void __launch_bounds__(THREADS, 4) k_gather_global(const float4 *__restrict__ panel, int panel_rows, int iters,
                                                              float4 *__restrict__ sink) {
    const int lane = threadIdx.x & 31;
    const int gl = lane % G, gi = lane / G;
    uint32_t seed = (blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5)) * 97u + gi * 7919u + 12345u;
    float4 acc[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int row_f4 = G * VPL;
    for (int it = 0; it < iters; it += UNROLL) {
        float4 x[UNROLL][VPL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int r = lcg(seed) % panel_rows;
            const float4 *p = panel + (long long)r * row_f4 + gl;
#pragma unroll
            for (int i = 0; i < VPL; ++i) x[u][i] = __ldg(p + i * G);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                acc[i].x += x[u][i].x; acc[i].y += x[u][i].y; acc[i].z += x[u][i].z; acc[i].w += x[u][i].w;
            }
    }
    float4 s = acc[0];
#pragma unroll
    for (int i = 1; i < VPL; ++i) { s.x += acc[i].x; s.y += acc[i].y; s.z += acc[i].z; s.w += acc[i].w; }
    if (s.x == 123.456f) sink[blockIdx.x * THREADS + threadIdx.x] = s;      // keep the loads alive
}

// modes 1 / 2: 128-byte runs (8 lanes x float4) out of shared memory; CLUSTER > 1 spreads the panel over the cluster
template <int CLUSTER>
__global__ void __launch_bounds__(THREADS, 1) k_gather_smem(int rows_per_cta, int iters, float4 *__restrict__ sink) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *panel = reinterpret_cast<float4 *>(smem_raw);
    const int lane = threadIdx.x & 31;
    constexpr int G = 8;
    const int gl = lane % G, gi = lane / G;
    for (int i = threadIdx.x; i < rows_per_cta * G; i += THREADS) panel[i] = make_float4((float)i, 1.f, 2.f, 3.f);
    uint32_t my_rank = 0;
    if constexpr (CLUSTER > 1) {
        cg::cluster_group cl = cg::this_cluster();
        my_rank = cl.block_rank();
        cl.sync();
    } else {
        __syncthreads();
    }
    uint32_t seed = (blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5)) * 97u + gi * 7919u + 12345u;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(panel);
    for (int it = 0; it < iters; it += UNROLL) {
        float4 x[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint32_t q = lcg(seed);
            const int r = q % rows_per_cta;
            const uint32_t addr = base + (uint32_t)(r * G + gl) * 16u;
            if constexpr (CLUSTER > 1) {
                const uint32_t target = (q / (uint32_t)rows_per_cta) % CLUSTER;      // uniformly any CTA of the cluster
                uint32_t remote;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(addr), "r"(target));
                asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];"
                             : "=f"(x[u].x), "=f"(x[u].y), "=f"(x[u].z), "=f"(x[u].w)
                             : "r"(remote));
            } else {
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                             : "=f"(x[u].x), "=f"(x[u].y), "=f"(x[u].z), "=f"(x[u].w)
                             : "r"(addr));
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) { acc.x += x[u].x; acc.y += x[u].y; acc.z += x[u].z; acc.w += x[u].w; }
    }
    if constexpr (CLUSTER > 1) cg::this_cluster().sync();      // nobody leaves while a peer still reads its memory
    if (acc.x == 123.456f) sink[blockIdx.x * THREADS + threadIdx.x] = acc;
    (void)my_rank;
}

}  // namespace

extern "C" {

// returns 0 on success; *ms = time of one launch, *bytes = bytes gathered by it, *ctas = grid size
int arrow_probe_gather(int mode, int row_bytes, int panel_rows, int iters, int ctas_per_sm, float *ms, double *bytes, int *ctas) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) return -1;
    const int sms = prop.multiProcessorCount;
    float4 *sink = nullptr;
    cudaMalloc(&sink, (size_t)sms * 8 * THREADS * sizeof(float4));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaError_t e = cudaSuccess;
    int grid = 0;
    double total = 0;
    if (mode == 0) {
        float4 *panel = nullptr;
        const size_t pbytes = (size_t)panel_rows * row_bytes;
        cudaMalloc(&panel, pbytes);
        cudaMemset(panel, 0, pbytes);
        grid = sms * (ctas_per_sm > 0 ? ctas_per_sm : 4);
        auto launch = [&]() {
            if (row_bytes == 512) k_gather_global<8, 4><<<grid, THREADS>>>(panel, panel_rows, iters, sink);
            else if (row_bytes == 128) k_gather_global<8, 1><<<grid, THREADS>>>(panel, panel_rows, iters, sink);
            else k_gather_global<4, 1><<<grid, THREADS>>>(panel, panel_rows, iters, sink);      // 64 B
        };
        launch();                                    // warm-up: panel into L2
        cudaEventRecord(a);
        launch();
        cudaEventRecord(b);
        e = cudaEventSynchronize(b);
        const int groups_per_cta = (THREADS / 32) * (32 / (row_bytes == 64 ? 4 : 8));
        total = (double)grid * groups_per_cta * (double)iters * row_bytes;
        cudaFree(panel);
    } else {
        const int cluster = (mode == 2) ? 8 : 1;
        const int rows_per_cta = panel_rows / cluster;                 // 128-byte runs per CTA
        const size_t smem = (size_t)rows_per_cta * 128;
        grid = (sms / cluster) * cluster;
        if (mode == 1) {
            cudaFuncSetAttribute(k_gather_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k_gather_smem<1><<<grid, THREADS, smem>>>(rows_per_cta, iters, sink);
            cudaEventRecord(a);
            k_gather_smem<1><<<grid, THREADS, smem>>>(rows_per_cta, iters, sink);
            cudaEventRecord(b);
        } else {
            cudaFuncSetAttribute(k_gather_smem<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid);
            cfg.blockDim = dim3(THREADS);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 8;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            float4 *sk = sink;
            e = cudaLaunchKernelEx(&cfg, k_gather_smem<8>, rows_per_cta, iters, sk);
            cudaEventRecord(a);
            if (e == cudaSuccess) e = cudaLaunchKernelEx(&cfg, k_gather_smem<8>, rows_per_cta, iters, sk);
            cudaEventRecord(b);
        }
        if (e == cudaSuccess) e = cudaEventSynchronize(b);
        total = (double)grid * (THREADS / 32) * 4 * (double)iters * 128;
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        fprintf(stderr, "arrow_probe_gather(mode %d): %s\n", mode, cudaGetErrorString(e));
        cudaFree(sink);
        return -2;
    }
    cudaEventElapsedTime(ms, a, b);
    *bytes = total;
    *ctas = grid;
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(sink);
    return 0;
}

}  // extern "C"
