// microbench.cu — measured FP64 denominators for the rooflines (MEASURED_PEAKS.json
// has no FP64 figure) and a few latency probes that size the sweep kernel's ILP.
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

// register-resident DFMA: ILP independent chains per thread
template <int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = (double)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

// register-resident DMMA m8n8k4 (mma.sync f64): ILP independent accumulator fragments per warp
template <int ILP>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a, double b) {
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = 0.0; c1[i] = 0.0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    if (s == 123.456) out[0] = s;
}

// single warp, one dependent DFMA chain: cycles per DFMA = latency
__global__ void dfma_latency_kernel(double* out, int iters, double a, double b) {
    double acc = (double)threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc = fma(acc, a, b);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (double)(t1 - t0) / ((double)iters * 16.0);
    if (acc == 123.456) out[1] = acc;
}

// single warp, 8 independent chains: cycles per DFMA = issue interval of one warp
__global__ void dfma_issue_kernel(double* out, int iters, double a, double b) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = (double)(threadIdx.x + i);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
    }
    long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (threadIdx.x == 0) out[0] = (double)(t1 - t0) / ((double)iters * 16.0);
    if (s == 123.456) out[1] = s;
}

// single warp DMMA: ILP independent accumulator chains; cycles per DMMA
template <int ILP>
__global__ void dmma_chain_kernel(double* out, int iters, double a, double b) {
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = 0.0; c1[i] = 0.0; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    if (threadIdx.x == 0) out[0] = (double)(t1 - t0) / ((double)iters * 4.0 * ILP);
    if (s == 123.456) out[1] = s;
}

// cluster barrier round trip (arrive.release + wait.acquire), cycles
__global__ void __launch_bounds__(256) cluster_barrier_kernel(double* out, int iters) {
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        asm volatile("barrier.cluster.arrive.release;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0) / (double)iters;
}

static int time_kernel(cudaStream_t st, float* ms, void (*launch)(cudaStream_t)) {
    cudaEvent_t e0, e1;
    PGAS_CUDA(cudaEventCreate(&e0));
    PGAS_CUDA(cudaEventCreate(&e1));
    launch(st);                                   // warm-up
    PGAS_CUDA(cudaStreamSynchronize(st));
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        PGAS_CUDA(cudaEventRecord(e0, st));
        launch(st);
        PGAS_CUDA(cudaEventRecord(e1, st));
        PGAS_CUDA(cudaEventSynchronize(e1));
        float t;
        PGAS_CUDA(cudaEventElapsedTime(&t, e0, e1));
        best = t < best ? t : best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms = best;
    PGAS_KERNEL_CHECK();
    return 0;
}

static double* g_scratch = nullptr;
static int g_sms = 148;
constexpr int MB_ITERS = 20000;

extern "C" int pgas_measure_fp64_peaks(double* dfma_tflops, double* dmma_tflops, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0;
    PGAS_CUDA(cudaGetDevice(&dev));
    PGAS_CUDA(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
    if (!g_scratch) PGAS_CUDA(cudaMalloc((void**)&g_scratch, 64 * sizeof(double)));
    float ms;
    // 8 CTAs of 256 threads per SM, 8 chains per thread
    if (int rc = time_kernel(st, &ms, [](cudaStream_t s) { dfma_kernel<8><<<g_sms * 8, 256, 0, s>>>(g_scratch, MB_ITERS, 1.0000001, 1e-9); }))
        return rc;
    if (dfma_tflops) *dfma_tflops = 2.0 * 8.0 * MB_ITERS * 256.0 * g_sms * 8.0 / (ms * 1e-3) / 1e12;
    if (int rc = time_kernel(st, &ms, [](cudaStream_t s) { dmma_kernel<8><<<g_sms * 8, 256, 0, s>>>(g_scratch, MB_ITERS, 1.0000001, 1e-9); }))
        return rc;
    // one m8n8k4 = 8*8*4 FMA = 512 flop per warp
    if (dmma_tflops) *dmma_tflops = 512.0 * 8.0 * MB_ITERS * 8.0 * g_sms * 8.0 / (ms * 1e-3) / 1e12;
    return 0;
}

// out[5] = dependent DMMA latency (cycles), out[6] = cycles per DMMA of one warp with 5 independent chains,
// out[7] = cycles per DMMA with 4 warps on one SM sub-partition each running 5 chains (16 warps/SM)
// out[0] = dependent DFMA latency (cycles), out[1] = single-warp DFMA issue interval (cycles),
// out[2] = cluster barrier round trip at cluster size 16 (cycles; -1 if not launchable),
// out[3] = same at cluster size 8, out[4] = DFMA TFLOP/s with 2 warps/SMSP x ILP 4 (the sweep's shape)
extern "C" int pgas_microbench_f64(double* out5, void* stream) {
    {
        cudaStream_t st0 = (cudaStream_t)stream;
        if (!g_scratch) PGAS_CUDA(cudaMalloc((void**)&g_scratch, 64 * sizeof(double)));
        double hh[2];
        dmma_chain_kernel<1><<<1, 32, 0, st0>>>(g_scratch, 2000, 1.0000001, 1e-9);
        PGAS_CUDA(cudaMemcpyAsync(hh, g_scratch, sizeof(hh), cudaMemcpyDeviceToHost, st0));
        PGAS_CUDA(cudaStreamSynchronize(st0));
        out5[5] = hh[0];
        dmma_chain_kernel<5><<<1, 32, 0, st0>>>(g_scratch, 2000, 1.0000001, 1e-9);
        PGAS_CUDA(cudaMemcpyAsync(hh, g_scratch, sizeof(hh), cudaMemcpyDeviceToHost, st0));
        PGAS_CUDA(cudaStreamSynchronize(st0));
        out5[6] = hh[0];
        dmma_chain_kernel<5><<<1, 512, 0, st0>>>(g_scratch, 2000, 1.0000001, 1e-9);
        PGAS_CUDA(cudaMemcpyAsync(hh, g_scratch, sizeof(hh), cudaMemcpyDeviceToHost, st0));
        PGAS_CUDA(cudaStreamSynchronize(st0));
        out5[7] = hh[0];
    }
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0;
    PGAS_CUDA(cudaGetDevice(&dev));
    PGAS_CUDA(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
    if (!g_scratch) PGAS_CUDA(cudaMalloc((void**)&g_scratch, 64 * sizeof(double)));
    double h[2];
    dfma_latency_kernel<<<1, 32, 0, st>>>(g_scratch, 2000, 1.0000001, 1e-9);
    PGAS_CUDA(cudaMemcpyAsync(h, g_scratch, sizeof(h), cudaMemcpyDeviceToHost, st));
    PGAS_CUDA(cudaStreamSynchronize(st));
    out5[0] = h[0];
    dfma_issue_kernel<<<1, 32, 0, st>>>(g_scratch, 2000, 1.0000001, 1e-9);
    PGAS_CUDA(cudaMemcpyAsync(h, g_scratch, sizeof(h), cudaMemcpyDeviceToHost, st));
    PGAS_CUDA(cudaStreamSynchronize(st));
    out5[1] = h[0];
    for (int ci = 0; ci < 2; ++ci) {
        const int C = ci == 0 ? 16 : 8;
        out5[2 + ci] = -1.0;
        if (C > 8 && cudaFuncSetAttribute(cluster_barrier_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(C, 1, 1);
        cfg.blockDim = dim3(256, 1, 1);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        if (cudaLaunchKernelEx(&cfg, cluster_barrier_kernel, g_scratch, 2000) != cudaSuccess) { cudaGetLastError(); continue; }
        PGAS_CUDA(cudaMemcpyAsync(h, g_scratch, sizeof(double), cudaMemcpyDeviceToHost, st));
        PGAS_CUDA(cudaStreamSynchronize(st));
        out5[2 + ci] = h[0];
    }
    float ms;
    if (int rc = time_kernel(st, &ms, [](cudaStream_t s) { dfma_kernel<4><<<g_sms, 256, 0, s>>>(g_scratch, MB_ITERS, 1.0000001, 1e-9); }))
        return rc;
    out5[4] = 2.0 * 4.0 * MB_ITERS * 256.0 * g_sms / (ms * 1e-3) / 1e12;
    return 0;
}

// FP64 pipe concurrency probe: every warp interleaves NF independent DFMA chains with NM independent DMMA fragments.  If the FP64
// FMA pipe and the FP64 tensor pipe were separate units, the combined rate would exceed either peak.
template <int NF, int NM>
__global__ void __launch_bounds__(256) dmix_kernel(double* out, int iters, double a, double b) {
    double acc[NF > 0 ? NF : 1], c0[NM > 0 ? NM : 1], c1[NM > 0 ? NM : 1];
#pragma unroll
    for (int i = 0; i < NF; ++i) acc[i] = (double)(threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < NM; ++i) { c0[i] = 0.0; c1[i] = 0.0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < (NF > NM ? NF : NM); ++i) {
            if (i < NM)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
            if (i < NF) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(a), "d"(b));
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NF; ++i) s += acc[i];
#pragma unroll
    for (int i = 0; i < NM; ++i) s += c0[i] + c1[i];
    if (s == 123.456) out[0] = s;
}

// out[0..2] = TFLOP/s of (8 DFMA chains), (8 DMMA fragments), (8 + 8 interleaved); out[3] = (8 DFMA + 2 DMMA)
extern "C" int pgas_microbench_mix_f64(double* out4, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0;
    PGAS_CUDA(cudaGetDevice(&dev));
    PGAS_CUDA(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
    if (!g_scratch) PGAS_CUDA(cudaMalloc((void**)&g_scratch, 64 * sizeof(double)));
    float ms;
    const double warps = 8.0 * g_sms * 8.0, it = MB_ITERS;
    if (int rc = time_kernel(st, &ms, [](cudaStream_t s) { dmix_kernel<8, 0><<<g_sms * 8, 256, 0, s>>>(g_scratch, MB_ITERS, 1.0000001, 1e-9); })) return rc;
    out4[0] = 64.0 * 8 * it * warps / (ms * 1e-3) / 1e12;
    if (int rc = time_kernel(st, &ms, [](cudaStream_t s) { dmix_kernel<0, 8><<<g_sms * 8, 256, 0, s>>>(g_scratch, MB_ITERS, 1.0000001, 1e-9); })) return rc;
    out4[1] = 512.0 * 8 * it * warps / (ms * 1e-3) / 1e12;
    if (int rc = time_kernel(st, &ms, [](cudaStream_t s) { dmix_kernel<8, 8><<<g_sms * 8, 256, 0, s>>>(g_scratch, MB_ITERS, 1.0000001, 1e-9); })) return rc;
    out4[2] = (512.0 + 64.0) * 8 * it * warps / (ms * 1e-3) / 1e12;
    if (int rc = time_kernel(st, &ms, [](cudaStream_t s) { dmix_kernel<8, 2><<<g_sms * 8, 256, 0, s>>>(g_scratch, MB_ITERS, 1.0000001, 1e-9); })) return rc;
    out4[3] = (512.0 * 2 + 64.0 * 8) * it * warps / (ms * 1e-3) / 1e12;
    return 0;
}
