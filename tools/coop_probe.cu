// Probe (tuning aid): can a kernel with grid-wide barriers be launched with the cooperative attribute AND programmatic
// stream serialization, captured into a CUDA graph, and run from several streams at once without deadlock?
//   nvcc -arch=sm_100a -o /tmp/coop_probe tools/coop_probe.cu && timeout 60 /tmp/coop_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void grid_barrier(unsigned int *ctr, unsigned int &epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += gridDim.x;
        __threadfence();
        atomicAdd(ctr, 1u);
        while (*(volatile unsigned int *)ctr < epoch) { }
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(1024) k_coop(unsigned int *ctr, unsigned int *data, int rounds, unsigned int base) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    unsigned int epoch = base;
    for (int r = 0; r < rounds; r++) {
        if (threadIdx.x == 0) data[blockIdx.x] = r;
        grid_barrier(ctr, epoch);
        // every block must see round r of every other block
        if (threadIdx.x < gridDim.x && data[threadIdx.x] != (unsigned)r) atomicAdd(&data[1024], 1u);
        grid_barrier(ctr, epoch);
    }
}
__global__ void k_dummy(unsigned int *x) { if (threadIdx.x == 0 && blockIdx.x == 0) x[2000] += 1; }

static cudaError_t launch(cudaStream_t st, bool coop, bool pdl, unsigned int *ctr, unsigned int *data, int rounds, unsigned base) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = 148; cfg.blockDim = 1024; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (coop) { attr[n].id = cudaLaunchAttributeCooperative; attr[n].val.cooperative = 1; n++; }
    if (pdl) { attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[n].val.programmaticStreamSerializationAllowed = 1; n++; }
    cfg.attrs = attr; cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, k_coop, ctr, data, rounds, base);
}

int main() {
    const int NS = 4;
    cudaStream_t st[NS];
    unsigned int *ctr[NS], *data[NS];
    for (int i = 0; i < NS; i++) {
        cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
        cudaMalloc(&ctr[i], 4); cudaMemset(ctr[i], 0, 4);
        cudaMalloc(&data[i], 4 * 4096); cudaMemset(data[i], 0, 4 * 4096);
    }
    for (int coop = 1; coop >= 0; coop--) for (int pdl = 0; pdl <= 1; pdl++) {
        cudaMemset(ctr[0], 0, 4);
        k_dummy<<<1, 32, 0, st[0]>>>(data[0]);
        cudaError_t e = launch(st[0], coop, pdl, ctr[0], data[0], 8, 0);
        cudaError_t e2 = cudaStreamSynchronize(st[0]);
        unsigned bad = 0; cudaMemcpy(&bad, data[0] + 1024, 4, cudaMemcpyDeviceToHost);
        printf("coop %d pdl %d: launch %s, sync %s, mismatches %u\n", coop, pdl, cudaGetErrorString(e), cudaGetErrorString(e2), bad);
        cudaGetLastError();
    }
    // graph capture with both attributes
    {
        cudaMemset(ctr[0], 0, 4);
        cudaGraph_t g; cudaGraphExec_t ge;
        cudaError_t e = cudaStreamBeginCapture(st[0], cudaStreamCaptureModeThreadLocal);
        k_dummy<<<1, 32, 0, st[0]>>>(data[0]);
        cudaError_t el = launch(st[0], true, true, ctr[0], data[0], 4, 0);
        cudaError_t ec = cudaStreamEndCapture(st[0], &g);
        printf("capture: begin %s launch %s end %s\n", cudaGetErrorString(e), cudaGetErrorString(el), cudaGetErrorString(ec));
        if (ec == cudaSuccess) {
            cudaError_t ei = cudaGraphInstantiate(&ge, g, 0);
            printf("instantiate %s\n", cudaGetErrorString(ei));
            if (ei == cudaSuccess) {
                cudaError_t er = cudaGraphLaunch(ge, st[0]);
                cudaError_t es = cudaStreamSynchronize(st[0]);
                printf("graph launch %s sync %s\n", cudaGetErrorString(er), cudaGetErrorString(es));
            }
        }
        cudaGetLastError();
    }
    // four streams at once, 20 launches each (cooperative): gang scheduling must keep them from dead-locking
    for (int coop = 1; coop >= 1; coop--) {
        for (int i = 0; i < NS; i++) cudaMemset(ctr[i], 0, 4);
        cudaDeviceSynchronize();
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, st[0]);
        for (int k = 0; k < 20; k++)
            for (int i = 0; i < NS; i++) launch(st[i], coop, true, ctr[i], data[i], 4, (unsigned)k * 148u * 8u);
        for (int i = 0; i < NS; i++) cudaStreamSynchronize(st[i]);
        cudaEventRecord(b, st[0]); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        unsigned bad = 0, tot = 0;
        for (int i = 0; i < NS; i++) { cudaMemcpy(&bad, data[i] + 1024, 4, cudaMemcpyDeviceToHost); tot += bad; }
        printf("4 streams x 20 launches (coop %d): %.3f ms, mismatches %u, last error %s\n", coop, ms, tot, cudaGetErrorString(cudaGetLastError()));
    }
    // cost of one launch with 8 barriers
    {
        cudaMemset(ctr[0], 0, 4);
        cudaDeviceSynchronize();
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        launch(st[0], true, true, ctr[0], data[0], 4, 0);
        cudaEventRecord(a, st[0]);
        for (int k = 1; k <= 10; k++) launch(st[0], true, true, ctr[0], data[0], 4, (unsigned)k * 148u * 8u);
        cudaEventRecord(b, st[0]); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("10 launches x 8 grid barriers: %.3f ms -> %.2f us per launch\n", ms, ms * 100.0);
        cudaEventRecord(a, st[0]);
        for (int k = 11; k <= 20; k++) launch(st[0], true, true, ctr[0], data[0], 16, (unsigned)(10 * 148 * 8 + (k - 11) * 148 * 32 + 148 * 8));
        cudaEventRecord(b, st[0]); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        printf("10 launches x 32 grid barriers: %.3f ms -> %.2f us per launch\n", ms, ms * 100.0);
    }
    return 0;
}
