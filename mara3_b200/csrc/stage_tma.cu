/**
 * stage_tma.cu -- translation unit of the persistent, TMA-staged stage kernel (stage_tma.cuh) and its launcher.
 * Kept apart from kernels.cu so that the hot kernel compiles in seconds.
 */
#include <algorithm>
#include <stdexcept>
#include <string>
#include "kernel_common.cuh"
#include "stage_tma.cuh"

namespace m3b { namespace dev {

size_t stage_tma_shared_bytes() { return sizeof(tma_smem_t<2>); }

template<typename K> static void set_smem(K kernel, size_t bytes)
{
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
    if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e) + " (stage_tma shared memory)");
}

/** the variants: 3 CTAs per SM with two tile buffers (168 registers), or 4 CTAs per SM with one (128 registers) */
template<int CTAS> static void configure()
{
    constexpr size_t bytes = sizeof(tma_smem_t<CTAS >= 4 ? 1 : 2>);
    set_smem(stage_tma<CTAS, 0, false, 0>, bytes);
    set_smem(stage_tma<CTAS, 0, true, 0>, bytes);
    set_smem(stage_tma<CTAS, 64, false, 0>, bytes);
    set_smem(stage_tma<CTAS, 64, true, 0>, bytes);
    set_smem(stage_tma<CTAS, 64, true, 1>, bytes);
    set_smem(stage_tma<CTAS, 64, true, 2>, bytes);
}

void stage_tma_configure()
{
    configure<3>();
    configure<4>();
}

template<int CTAS> static void launch(const stage_tma_launch_t& a, cudaStream_t stream)
{
    constexpr size_t bytes = sizeof(tma_smem_t<CTAS >= 4 ? 1 : 2>);
    auto kernel = a.N == 64 ? (a.fast ? stage_tma<CTAS, 64, true, 0> : stage_tma<CTAS, 64, false, 0>)
                            : (a.fast ? stage_tma<CTAS, 0, true, 0> : stage_tma<CTAS, 0, false, 0>);
    if (a.N == 64 && a.fast && a.stage_mode == 1) kernel = stage_tma<CTAS, 64, true, 1>;
    if (a.N == 64 && a.fast && a.stage_mode == 2) kernel = stage_tma<CTAS, 64, true, 2>;
    // every CTA of the grid must be resident at once (the kernel is persistent, and with the fused exchange its CTAs wait for
    // each other's unpacked strips): never more CTAs than the occupancy calculator grants
    int per_sm = 0, device = 0, sms = 0;
    cudaGetDevice(&device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, STRIP_THREADS, bytes) != cudaSuccess || per_sm < 1)
        throw std::runtime_error("stage_tma: the kernel does not fit on this device");
    const int grid = std::max(1, std::min(a.grid, std::min(per_sm, a.ctas_per_sm) * sms));
    kernel<<<grid, STRIP_THREADS, bytes, stream>>>(a.mesh, a.model, a.stage, a.tile_info, a.num_tiles, a.Uin, a.Un, a.Uout, a.partials, a.jump_partials, a.fail, a.exchange);
}

void stage_tma_launch(const stage_tma_launch_t& a, cudaStream_t stream)
{
    if (a.ctas_per_sm >= 4) launch<4>(a, stream); else launch<3>(a, stream);
}

}} // namespace m3b::dev
