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

size_t stage_tma_shared_bytes() { return sizeof(tma_smem_t); }

template<typename K> static void set_smem(K kernel)
{
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(tma_smem_t)));
    if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e) + " (stage_tma shared memory)");
}

void stage_tma_configure()
{
    set_smem(stage_tma<3, 0, false, 0>);
    set_smem(stage_tma<3, 0, true, 0>);
    set_smem(stage_tma<3, 64, false, 0>);
    set_smem(stage_tma<3, 64, true, 0>);
    set_smem(stage_tma<3, 64, true, 1>);
    set_smem(stage_tma<3, 64, true, 2>);
}

void stage_tma_launch(const stage_tma_launch_t& a, cudaStream_t stream)
{
    auto kernel = a.N == 64 ? (a.fast ? stage_tma<3, 64, true, 0> : stage_tma<3, 64, false, 0>)
                            : (a.fast ? stage_tma<3, 0, true, 0> : stage_tma<3, 0, false, 0>);
    if (a.N == 64 && a.fast && a.stage_mode == 1) kernel = stage_tma<3, 64, true, 1>;
    if (a.N == 64 && a.fast && a.stage_mode == 2) kernel = stage_tma<3, 64, true, 2>;
    // every CTA of the grid must be resident at once (the kernel is persistent, and with the fused exchange its CTAs wait for
    // each other's unpacked strips): never more CTAs than the occupancy calculator grants
    int per_sm = 0, device = 0, sms = 0;
    cudaGetDevice(&device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, STRIP_THREADS, sizeof(tma_smem_t)) != cudaSuccess || per_sm < 1)
        throw std::runtime_error("stage_tma: the kernel does not fit on this device");
    const int grid = std::max(1, std::min(a.grid, per_sm * sms));
    kernel<<<grid, STRIP_THREADS, sizeof(tma_smem_t), stream>>>(a.mesh, a.model, a.stage, a.tile_info, a.num_tiles, a.Uin, a.Un, a.Uout, a.partials, a.fail, a.exchange);
}

}} // namespace m3b::dev
