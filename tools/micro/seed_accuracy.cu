// Accuracy of the fp64 reciprocal / reciprocal-square-root hardware seeds on sm_100a and of the refinements built on them
// (iso2d_device.cuh: fast_rcp, fast_rsqrt).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o seed_accuracy seed_accuracy.cu
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>

__global__ void probe(int n, double lo, double ratio, double* out)
{
    // out[0..5]: max relative error of rcp seed, rsqrt seed, rcp quadratic, rcp cubic, rsqrt quadratic, rsqrt cubic
    double e[6] = {0, 0, 0, 0, 0, 0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        double x = lo * pow(ratio, double(i) / n) * (1.0 + 1e-9 * (i % 977));
        double y, z;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(z) : "d"(x));
        double ry = 1.0 / x, rz = 1.0 / sqrt(x);
        e[0] = fmax(e[0], fabs(y - ry) / ry);
        e[1] = fmax(e[1], fabs(z - rz) / rz);
        { double d = fma(-x, y, 1.0); double q = fma(y, d, y); e[2] = fmax(e[2], fabs(q - ry) / ry);
          double t = fma(d, d, d); double c = fma(y, t, y); e[3] = fmax(e[3], fabs(c - ry) / ry); }
        { double t = x * z; double d = fma(-t, z, 1.0); double q = fma(0.5 * z, d, z); e[4] = fmax(e[4], fabs(q - rz) / rz);
          double p = fma(0.375, d, 0.5); double c = fma(z, p * d, z); e[5] = fmax(e[5], fabs(c - rz) / rz); }
    }
    for (int k = 0; k < 6; ++k)
    {
        for (int o = 16; o > 0; o >>= 1) e[k] = fmax(e[k], __shfl_xor_sync(0xffffffffu, e[k], o));
        if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned long long*>(out + k), __double_as_longlong(e[k]));
    }
}

int main()
{
    double* d; cudaMalloc(&d, 6 * sizeof(double)); cudaMemset(d, 0, 6 * sizeof(double));
    probe<<<592, 256>>>(1 << 28, 1e-9, 1e12, d);
    double h[6]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[6] = {"rcp.approx.ftz.f64 seed", "rsqrt.approx.ftz.f64 seed", "rcp + 1 quadratic step", "rcp + 1 cubic step (fast_rcp)", "rsqrt + 1 quadratic step", "rsqrt + 1 cubic step (fast_rsqrt)"};
    for (int k = 0; k < 6; ++k) printf("%-36s max rel err %.3e  (2^%.1f)\n", names[k], h[k], log2(h[k]));
    return cudaGetLastError() != cudaSuccess;
}
