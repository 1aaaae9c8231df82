// Micro-benchmark: fp64 FMA latency / throughput on sm_100a as a function of ILP and warps per SM.
#include <cstdio>
#include <cuda_runtime.h>

template<int ILP>
__global__ void chain(double* out, int iters, double a, double b)
{
    double x[ILP];
    for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x * 1e-9 + k;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
    {
        #pragma unroll
        for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = double(t1 - t0);
}

template<int ILP>
void run(int threads)
{
    double* out;
    cudaMalloc(&out, 148 * 1024 * sizeof(double));
    int iters = 4096;
    chain<ILP><<<148, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    chain<ILP><<<148, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    double cyc;
    cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
    double per_fma_per_warp = cyc / (double(iters) * ILP);
    double warps = threads / 32.0;
    printf("ILP %d warps/SM %4.0f: %.2f cycles per dependent step, %.3f DFMA warp-instr/cycle/SM\n", ILP, warps, cyc / iters, warps * ILP * iters / cyc);
    cudaFree(out);
}

int main()
{
    for (int threads : {32, 128, 256, 512, 1024})
    {
        run<1>(threads); run<2>(threads); run<4>(threads); run<8>(threads);
    }
    return 0;
}
