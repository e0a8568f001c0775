// fp64_peak.cu — microbenchmark of the FP64 pipes on this GPU: dependent-chain latency and saturated throughput of
// DADD / DFMA / DSETP, and of the FP64 tensor-core instruction mma.sync.m8n8k4.f64 (DMMA).  Gives the denominators the
// wLOD roofline is quoted against; MEASURED_PEAKS.json has no FP64 entry.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak tools/fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void lat_dadd(double* out, long long* cyc, int n)
{
    double x = out[0], y = out[1];
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 32; ++k) x = __dadd_rn(x, y);
    }
    long long t1 = clock64();
    out[threadIdx.x + 2] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP>
__global__ void thr(double* out, int n)
{
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = out[k] + threadIdx.x;
    const double y = out[9], z = out[10];
    unsigned f = 0;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (OP == 0) a[k] = __dadd_rn(a[k], y);
                if (OP == 1) a[k] = __fma_rn(a[k], y, z);
                if (OP == 2) { f += (a[k] >= y + (double)(r + i)); }
            }
    }
    double s = f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x + 16] = s;
}

__global__ void thr_dmma(double* out, int n)
{
    double c[8][2];
#pragma unroll
    for (int k = 0; k < 8; ++k) { c[k][0] = 0; c[k][1] = 0; }
    const double a = out[0] + threadIdx.x, b = out[1];
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
    out[blockIdx.x * blockDim.x + threadIdx.x + 16] = s;
}

__global__ void lat_dmma(double* out, long long* cyc, int n)
{
    double c0 = 0, c1 = 0;
    const double a = out[0] + threadIdx.x, b = out[1];
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    }
    long long t1 = clock64();
    out[threadIdx.x + 16] = c0 + c1;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main()
{
    double* d; long long* c;
    cudaMalloc(&d, (1 << 22) * sizeof(double)); cudaMemset(d, 0, (1 << 22) * sizeof(double));
    cudaMalloc(&c, 64);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    long long h;
    lat_dadd<<<1, 32>>>(d, c, 1000); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"dadd_dependent_latency_cycles\": %.2f", p.name, p.multiProcessorCount, h / 32000.0);
    lat_dmma<<<1, 32>>>(d, c, 1000); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf(", \"dmma_m8n8k4_dependent_latency_cycles\": %.2f", h / 16000.0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = p.multiProcessorCount * 8, threads = 256, n = 4000;
    const char* names[3] = {"dadd", "dfma", "dsetp_plus_iadd"};
    for (int op = 0; op < 3; ++op) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (op == 0) thr<0><<<blocks, threads>>>(d, n);
            if (op == 1) thr<1><<<blocks, threads>>>(d, n);
            if (op == 2) thr<2><<<blocks, threads>>>(d, n);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double ops = (double)blocks * threads * n * 32.0;
        printf(", \"%s_Gops\": %.1f", names[op], ops / ms / 1e6);
    }
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        thr_dmma<<<blocks, threads>>>(d, n);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fl = (double)blocks * (threads / 32) * n * 8.0 * 512.0;   // m8n8k4 = 256 FMA = 512 flop per warp instruction
    printf(", \"dmma_m8n8k4_TFLOPs\": %.2f, \"clock_khz_attr\": %d}\n", fl / ms / 1e9, clk);
    return 0;
}
