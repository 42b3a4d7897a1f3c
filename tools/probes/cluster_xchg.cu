// Probe (diagnostic, not product): per-step latency floor of an all-to-all exchange inside one thread-block cluster of 16
// CTAs, the pattern of a cluster-resident LSTM recurrence.  Each step every CTA pushes `bytes_per_pair` bytes into every
// peer's shared memory (st.async ... mbarrier::complete_tx, no cluster barrier) and waits until its own buffer has received
// 16 x bytes_per_pair.  Variant 1 does the same with plain st.shared::cluster + barrier.cluster.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_xchg cluster_xchg.cu && ./cluster_xchg
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

constexpr int CL = 16;
constexpr int NT = 512;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    long long spins = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (++spins > (1ll << 24)) { printf("probe: mbarrier timeout block %d\n", blockIdx.x); __trap(); }
    }
}

// mode 0: st.async push + mbarrier tx-count;  mode 1: st.shared::cluster push + barrier.cluster
template <int MODE>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NT, 1) xchg_kernel(int steps, int vec_per_pair, float* out) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);                  // [2]
    uint4* buf = reinterpret_cast<uint4*>(smem + 64);                    // [2][CL][vec_per_pair]
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const int total_vec = CL * vec_per_pair;
    float acc = 0.f;
    for (int s = 0; s < steps; ++s) {
        const int par = s & 1;
        if (MODE == 0) {
            if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[par])), "r"(total_vec * 16) : "memory");
            // every thread pushes vectors: item i -> (peer = i / vec_per_pair, slot = i % vec_per_pair)
            for (int i = tid; i < total_vec; i += NT) {
                const int peer = i / vec_per_pair, slot = i % vec_per_pair;
                const uint32_t dst = mapa(smem_u32(&buf[(par * CL + rank) * vec_per_pair + slot]), peer);
                const uint32_t rbar = mapa(smem_u32(&bars[par]), peer);
                const uint32_t v = s * 1000 + rank;
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                             ::"r"(dst), "r"(v), "r"(v), "r"(v), "r"(v), "r"(rbar) : "memory");
            }
            mbar_wait(&bars[par], (s >> 1) & 1);
        } else if (MODE == 1) {
            for (int i = tid; i < total_vec; i += NT) {
                const int peer = i / vec_per_pair, slot = i % vec_per_pair;
                const uint32_t dst = mapa(smem_u32(&buf[(par * CL + rank) * vec_per_pair + slot]), peer);
                const uint32_t v = s * 1000 + rank;
                asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v), "r"(v), "r"(v), "r"(v) : "memory");
            }
            asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        } else {
            // mode 2: the slice is first written to a local staging tile, then ONE bulk copy per peer (the copy engine moves
            // vec_per_pair * 16 bytes and completes them on the peer's mbarrier)
            uint4* stage = buf + 2 * total_vec;                          // [vec_per_pair]
            for (int i = tid; i < vec_per_pair; i += NT) stage[i] = make_uint4(s, rank, s, rank);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[par])), "r"(total_vec * 16) : "memory");
            if (tid < CL) {
                const uint32_t dst = mapa(smem_u32(&buf[(par * CL + rank) * vec_per_pair]), tid);
                const uint32_t rbar = mapa(smem_u32(&bars[par]), tid);
                asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "r"(smem_u32(stage)), "r"(vec_per_pair * 16), "r"(rbar) : "memory");
            }
            mbar_wait(&bars[par], (s >> 1) & 1);
        }
        // consume: every thread reads one received vector (as the mat-vec would) -- keeps the loads from being elided
        const uint4 r = buf[par * total_vec + (tid % total_vec)];
        acc += (float)(r.x & 0xff);
        __syncthreads();
    }
    if (out) out[blockIdx.x * NT + tid] = acc;
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int MODE>
static void run(int n_clusters, int steps, int vec_per_pair, float* out) {
    auto kern = xchg_kernel<MODE>;
    const size_t smem = 64 + 2 * CL * (size_t)vec_per_pair * 16 + (size_t)vec_per_pair * 16;
    cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        kern<<<n_clusters * CL, NT, smem>>>(steps, vec_per_pair, out);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("mode %d: launch failed: %s\n", MODE, cudaGetErrorString(err)); return; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep == 2)
            printf("mode %d (%s) clusters=%d bytes/pair=%5d: %8.1f us total, %6.3f us/step\n", MODE,
                   MODE == 0 ? "st.async+mbarrier" : MODE == 1 ? "st.cluster+barrier.cluster" : "bulk copy per peer+mbarrier", n_clusters, vec_per_pair * 16, ms * 1e3, ms * 1e3 / steps);
    }
}

int main() {
    float* out;
    cudaMalloc(&out, sizeof(float) * 64 * NT);
    int max_clusters = 0;
    {
        cudaLaunchConfig_t q = {};
        q.gridDim = dim3(2 * CL); q.blockDim = dim3(NT); q.dynamicSmemBytes = 64 + 2 * CL * 132 * 16 + 132 * 16;
        cudaLaunchAttribute qa[1];
        qa[0].id = cudaLaunchAttributeClusterDimension;
        qa[0].val.clusterDim.x = CL; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
        q.attrs = qa; q.numAttrs = 1;
        cudaFuncSetAttribute(xchg_kernel<0>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaFuncSetAttribute(xchg_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q.dynamicSmemBytes);
        cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, xchg_kernel<0>, &q);
        printf("max active clusters of 16 (512 threads, %zu B smem): %d (%s)\n", (size_t)q.dynamicSmemBytes, max_clusters, cudaGetErrorString(e));
    }
    const int steps = 400;
    for (int vec : {33, 66, 132, 264}) {           // 528 B (h hi, B=8), 1056 B (hi+lo / fp32 partials), 2112, 4224
        run<0>(2, steps, vec, out);
        run<1>(2, steps, vec, out);
        run<2>(2, steps, vec, out);
    }
    run<0>(1, steps, 33, out);
    cudaFree(out);
    return 0;
}
