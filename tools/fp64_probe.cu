// FP64 issue latency / throughput probe for sm_100a (tuning aid; nvcc -arch=sm_100a -fmad=false -o /tmp/fp64_probe tools/fp64_probe.cu).
// Dependent chains of DADD / DMUL / division per thread, ILP independent chains, W warps per SM: cycles per instruction
// per warp tell how many warps (or how much ILP) the float64 chains of the rasteriser need to keep the pipe busy.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int OP>
__global__ void k_chain(double *out, double seed, int iters, long long *cycles) {
    double v[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) v[k] = seed + k + threadIdx.x * 1e-3;
    const double c = seed * 0.5 + 1.0000001;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) {
            if (OP == 0) v[k] = v[k] + c;
            else if (OP == 1) v[k] = v[k] * c;
            else v[k] = c / v[k] + 1.0;
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s += v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int ILP, int OP>
void run(const char *name, int warps_per_sm, int sms) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, sizeof(double) * 2048 * sms);
    cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k_chain<ILP, OP><<<sms, warps_per_sm * 32>>>(out, 1.0, iters, cyc);
    k_chain<ILP, OP><<<sms, warps_per_sm * 32>>>(out, 1.0, iters, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_inst = (double)h / iters / ILP;  // cycles per instruction of one warp's chain set
    printf("%-5s ILP %d  warps/SM %2d : %7.2f cycles per op per warp;  SM throughput %6.2f warp-ops/cycle\n", name, ILP, warps_per_sm,
           per_inst * ILP / ILP, warps_per_sm / per_inst);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    int sms = 148;
    for (int w : {1, 4, 8, 16, 24, 32}) { run<1, 0>("DADD", w, sms); }
    for (int w : {1, 4, 8, 16, 24, 32}) { run<2, 0>("DADD", w, sms); }
    for (int w : {1, 4, 8, 16, 24, 32}) { run<4, 0>("DADD", w, sms); }
    for (int w : {1, 24}) { run<1, 1>("DMUL", w, sms); }
    for (int w : {1, 4, 8, 16, 24, 32}) { run<1, 2>("DDIV", w, sms); }
    for (int w : {1, 24}) { run<3, 2>("DDIV", w, sms); }
    return 0;
}
