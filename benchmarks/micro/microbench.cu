// microbench.cu -- instruction-throughput probes that decide kernel design choices on B200 (sm_100a):
// warp shuffles vs shared-memory accesses vs packed FP32, alone and mixed, at 8 / 12 / 16 warps per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu ; run: ./microbench
// Output: one JSON line per probe with warp-instructions per clock per SM (clock64-based, per-SM average).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

enum { P_SHFL = 0, P_LDS64, P_STS64, P_XCHG64, P_FFMA2, P_SHFL_FFMA2, P_LDS_FFMA2, P_LDS128, P_SHFL_LDS, P_FFMA, P_COUNT };
static const char* NAMES[] = {"shfl_idx_b32", "lds64", "sts64", "sts64+lds64 exchange", "ffma2", "shfl + ffma2 (1:4)", "lds64 + ffma2 (1:4)",
                              "lds128", "shfl + lds64 (1:1)", "ffma"};

template <int P> __global__ void __launch_bounds__(1024) probe(float* out, long long* cycles, int iters)
{
    extern __shared__ float2 sm[];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    float2* mine = sm + w * 1024;                    // 8 KB per warp
    for (int i = lane; i < 1024; i += 32) mine[i] = make_float2(i * 1e-3f, 1.f - i * 1e-3f);
    __syncthreads();
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(t * 1e-3f + i, 1.0f - i * 1e-2f);
    const float2 m = make_float2(0.999f, 1.001f), c = make_float2(1e-3f, -1e-3f);
    const int src = (32 - lane) & 31;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (P == P_SHFL) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i].x = __shfl_sync(0xffffffffu, a[i].x, src); a[i].y = __shfl_sync(0xffffffffu, a[i].y, src); }
        } else if (P == P_LDS64) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { float2 v = mine[lane + 32 * ((i + it) & 31)]; a[i].x += v.x; a[i].y += v.y; }
        } else if (P == P_LDS128) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { float4 v = reinterpret_cast<float4*>(mine)[lane + 32 * ((i + it) & 15)]; a[i].x += v.x + v.z; a[i].y += v.y + v.w; }
        } else if (P == P_STS64) {
#pragma unroll
            for (int i = 0; i < 8; ++i) mine[lane + 32 * ((i + it) & 31)] = a[i];
        } else if (P == P_XCHG64) {
#pragma unroll
            for (int i = 0; i < 8; ++i) mine[lane + 32 * i] = a[i];
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = mine[src + 32 * i];
            __syncwarp();
        } else if (P == P_FFMA2) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(a[i], m, c);
        } else if (P == P_FFMA) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
        } else if (P == P_SHFL_FFMA2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a[i].x = __shfl_sync(0xffffffffu, a[i].x, src);
#pragma unroll
                for (int u = 0; u < 4; ++u) a[(i + 1 + u) & 7] = __ffma2_rn(a[(i + 1 + u) & 7], m, c);
            }
        } else if (P == P_LDS_FFMA2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float2 v = mine[lane + 32 * ((i + it) & 31)];
#pragma unroll
                for (int u = 0; u < 4; ++u) a[(i + 1 + u) & 7] = __ffma2_rn(a[(i + 1 + u) & 7], m, c);
                a[i].x += v.x; a[i].y += v.y;
            }
        } else if (P == P_SHFL_LDS) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float2 v = mine[lane + 32 * ((i + it) & 31)];
                a[i].x = __shfl_sync(0xffffffffu, a[i].x, src);
                a[i].y += v.x + v.y;
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    if (s == 123.456f) out[0] = s;
    if (t == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int P> static void run(int warps, float* d_out, long long* d_cyc, int sms)
{
    const int iters = 4096, threads = warps * 32;
    const size_t smem = (size_t)warps * 8192;
    CK(cudaFuncSetAttribute(probe<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe<P><<<sms, threads, smem>>>(d_out, d_cyc, 16);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    probe<P><<<sms, threads, smem>>>(d_out, d_cyc, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    long long* h = (long long*)malloc(sizeof(long long) * sms);
    CK(cudaMemcpy(h, d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double cyc = 0; for (int i = 0; i < sms; ++i) cyc += (double)h[i]; cyc /= sms;
    free(h);
    // warp-level instructions of the probed kind per iteration per warp
    double per_iter;
    switch (P) {
        case P_SHFL: per_iter = 16; break;
        case P_LDS64: case P_STS64: case P_LDS128: per_iter = 8; break;
        case P_XCHG64: per_iter = 16; break;
        case P_FFMA2: per_iter = 32; break;
        case P_FFMA: per_iter = 64; break;
        case P_SHFL_FFMA2: case P_LDS_FFMA2: per_iter = 8 + 32; break;
        default: per_iter = 16; break;
    }
    const double total = per_iter * iters * warps;
    printf("{\"probe\": \"%s\", \"warps_per_sm\": %d, \"warp_instr_per_clk_per_sm\": %.3f, \"cycles\": %.0f, \"ms\": %.4f, \"implied_mhz\": %.0f}\n",
           NAMES[P], warps, total / cyc, cyc, ms, cyc / (ms * 1e-3) / 1e6);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

int main()
{
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, 4)); CK(cudaMalloc(&d_cyc, sizeof(long long) * sms));
    const int ws[] = {4, 8, 12, 16, 24};
    for (int wi = 0; wi < 5; ++wi) {
        const int w = ws[wi];
        run<P_SHFL>(w, d_out, d_cyc, sms);
        run<P_LDS64>(w, d_out, d_cyc, sms);
        run<P_LDS128>(w, d_out, d_cyc, sms);
        run<P_STS64>(w, d_out, d_cyc, sms);
        run<P_XCHG64>(w, d_out, d_cyc, sms);
        run<P_FFMA2>(w, d_out, d_cyc, sms);
        run<P_FFMA>(w, d_out, d_cyc, sms);
        run<P_SHFL_FFMA2>(w, d_out, d_cyc, sms);
        run<P_LDS_FFMA2>(w, d_out, d_cyc, sms);
        run<P_SHFL_LDS>(w, d_out, d_cyc, sms);
    }
    return 0;
}
