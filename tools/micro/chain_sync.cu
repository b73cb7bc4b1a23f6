// Micro-benchmark (development): period of a dependent chain of small kernels launched with programmatic stream
// serialization inside a CUDA graph, when a kernel waits for its predecessor with
//   mode 0: griddepcontrol.wait (completion + memory flush of the predecessor grid)
//   mode 1: a counter the predecessor's CTAs add to after their last store (release), polled by one thread per CTA (acquire)
// Every kernel reads the rows its predecessor wrote (ROWS x 1024 floats) and writes them back.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain_sync chain_sync.cu && ./chain_sync [ctas] [threads]
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void step_kernel(const float* __restrict__ in, float* __restrict__ out, int n4, unsigned int* wait_ctr, unsigned int expected,
                            unsigned int* arrive_ctr, int mode) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (mode == 0 || wait_ctr == nullptr) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
    } else {
        if (threadIdx.x == 0) {
            unsigned int v;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(wait_ctr) : "memory");
            } while (v < expected);
        }
        __syncthreads();
    }
    const float4* i4 = reinterpret_cast<const float4*>(in);
    float4* o4 = reinterpret_cast<float4*>(out);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        float4 v = __ldcg(i4 + i);
        v.x += 1.f; v.y += 1.f; v.z += 1.f; v.w += 1.f;
        o4[i] = v;
    }
    if (mode == 1 && arrive_ctr) {
        __syncthreads();
        if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(arrive_ctr) : "memory");
    }
}

int main(int argc, char** argv) {
    const int ctas = argc > 1 ? atoi(argv[1]) : 256, threads = argc > 2 ? atoi(argv[2]) : 128, N = 200, rows = 256;
    const int n4 = rows * 1024 / 4;
    float *a, *b;
    unsigned int* ctr;
    cudaMalloc(&a, n4 * 16); cudaMalloc(&b, n4 * 16); cudaMalloc(&ctr, (N + 1) * 4);
    cudaMemset(a, 0, n4 * 16);
    cudaStream_t st;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    for (int mode = 0; mode < 2; ++mode) {
        cudaGraph_t graph;
        cudaGraphExec_t exec;
        cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed);
        cudaMemsetAsync(ctr, 0, (N + 1) * 4, st);
        for (int k = 0; k < N; ++k) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(threads); cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            const float* in = (k & 1) ? b : a;
            float* out = (k & 1) ? a : b;
            unsigned int* w = k > 0 ? ctr + k - 1 : nullptr;
            cudaLaunchKernelEx(&cfg, step_kernel, in, out, n4, w, (unsigned int)ctas, ctr + k, mode);
        }
        cudaStreamEndCapture(st, &graph);
        cudaGraphInstantiate(&exec, graph, 0);
        for (int i = 0; i < 3; ++i) cudaGraphLaunch(exec, st);
        cudaStreamSynchronize(st);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
        for (int i = 0; i < 20; ++i) cudaGraphLaunch(exec, st);
        cudaEventRecord(e1, st);
        cudaStreamSynchronize(st);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        float h;
        cudaMemcpy(&h, a, 4, cudaMemcpyDeviceToHost);
        printf("mode %d (%s): %.3f us per kernel (%d CTAs x %d threads, 1 MB in / out each), check %.0f, err %s\n", mode,
               mode ? "counter arrive / poll" : "griddepcontrol.wait", ms * 1e3 / (20 * N), ctas, threads, h, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
