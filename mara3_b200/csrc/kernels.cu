/**
 * kernels.cu -- device_solver_t: the launch logic of the iso2d `binary` update (what runs on which stream, in which order, with
 * which tile lists) and the device-side state it owns.  One translation unit with the kernels it launches directly:
 *
 *   any_tree_kernels.cuh   stage_fused<TX, TY> (regular blocks, block sizes that are not multiples of 32), general_gradients*,
 *                          general_update* (blocks at refinement jumps outside the strip kernel's JUMP variant)
 *   stage_strip.cuh        the one-shot strip kernel: blocks at jumps (JUMP), conserved_q (QMODE), M3B_STAGE=strip
 *   finish_kernels.cuh     max_timestep_kernel, prepare_positions, finish_stage / finish_stage_cluster / peer_prepare /
 *                          prepare_next, the product and state-layout kernels
 *
 * The persistent stage kernel (stage_tma.cuh) and the guard-zone transport of the non-fused paths have their own translation
 * units (stage_tma.cu, transport.cu).  Reference: phases P1-P8 and P11 of binary::advance_u (Mara3
 * src/subprog_binary_scheme.cpp:790-904), binary::maximum_timestep (:1107-1126), next_solution (subprog_binary.cpp:258-293).
 */
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include "comm.hpp"
#include "device_solver.hpp"
#include "iso2d_device.cuh"
#include "kernel_common.cuh"
#include "device_solver_impl.cuh"

using namespace m3b;
using namespace m3b::dev;


#include "any_tree_kernels.cuh"     // reductions, stage_fused, general_gradients*, general_update*
#include "stage_strip.cuh"      // after the any-tree helpers: its JUMP variant fetches guard cells through them
#include "finish_kernels.cuh"       // max_timestep_kernel, prepare_positions, finish_stage*, prepare_next, product kernels
namespace
{
    template<typename T>
    T* device_upload(const std::vector<T>& v)
    {
        T* p = nullptr;
        M3B_CUDA(cudaMalloc(&p, std::max<size_t>(1, v.size()) * sizeof(T)));
        if (! v.empty()) M3B_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
        return p;
    }
}




// ===========================================================================
// device_field_t
// ===========================================================================
device_field_t::device_field_t(std::size_t num_doubles, int device) : count(num_doubles), device(device)
{
    M3B_CUDA(cudaSetDevice(device));
    M3B_CUDA(cudaMalloc(&data, std::max<size_t>(1, count) * sizeof(double)));
}

device_field_t::~device_field_t()
{
    if (data) cudaFree(data);
}




// ===========================================================================
// device_solver_t
// ===========================================================================
device_solver_t::device_solver_t(const solver_data_t& sd, int device, bool general_only, bool tiled_kernel) : impl(new impl_t), device_id(device), force_general(general_only)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        throw std::runtime_error("mara3_b200: no CUDA device is available (this library has no CPU fallback)");
    M3B_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    M3B_CUDA(cudaGetDeviceProperties(&prop, device));
    impl->sm_count = prop.multiProcessorCount;

    N = sd.block_size;
    B = sd.num_local;               // blocks stored on this rank: owned, then ghosts
    BO = sd.num_owned;
    cells = sd.num_local_cells();
    const auto& tree = *sd.tree;
    const auto& part = sd.partition;
    impl->num_global_blocks = sd.num_blocks;
    auto local = [&part] (int global) { return global < 0 ? -1 : part.global_to_local[global]; };

    cudaStream_t s;
    M3B_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    stream_ = own_stream = s;

    // ---- static mesh data
    auto spacing = std::vector<double>(B);
    auto nbr = std::vector<face_nbr_dev_t>(size_t(B) * 4);
    auto nbr9 = std::vector<int>(size_t(B) * 9, -1);
    auto in_gradient_set = std::vector<char>(B, 0);

    for (int b = 0; b < B; ++b)
    {
        const int g = sd.global_block(b);
        spacing[b] = sd.spacing(tree.index(g).level);

        // face tables also for ghost blocks: the any-tree kernels compute gradients of layer-1 ghosts from
        // layer-2 ghosts and walk from a fine ghost back to the coarse block it borders (ids of blocks this
        // rank does not store are -1 and never followed)
        for (int side = 0; side < 4; ++side)
        {
            auto fn = tree.face_neighbor(g, side);
            auto& d = nbr[size_t(b) * 4 + side];
            d.kind = int(fn.kind);
            for (int q = 0; q < 4; ++q) d.leaf[q] = local(fn.leaf[q]);
            d.bx = fn.bx; d.by = fn.by; d.pad = 0;
            for (int q = 0; q < 4; ++q) d.gs[q] = -1;
        }
        if (b >= BO) continue;          // ghost blocks are only read from
        bool regular = true;
        for (int di = -1; di <= 1; ++di)
            for (int dj = -1; dj <= 1; ++dj)
            {
                int l = local(tree.same_level_neighbor(g, di, dj));
                nbr9[size_t(b) * 9 + (di + 1) * 3 + (dj + 1)] = l;
                if (l < 0) regular = false;
            }
        (regular ? impl->regular : impl->irregular).push_back(b);
    }
    // tile shape of the fused kernel; blocks whose size no tile divides all take the general path
    if      (N % 32 == 0) { impl->tile_x = 16; impl->tile_y = 32; impl->strip = ! tiled_kernel; }
    else if (N % 24 == 0) { impl->tile_x = 12; impl->tile_y = 24; }
    else if (N % 16 == 0) { impl->tile_x = 16; impl->tile_y = 16; }
    else if (N % 8 == 0)  { impl->tile_x = 8;  impl->tile_y = 8; }
    else
    {
        impl->irregular.insert(impl->irregular.end(), impl->regular.begin(), impl->regular.end());
        std::sort(impl->irregular.begin(), impl->irregular.end());
        impl->regular.clear();
    }
    // interior blocks (every neighbour owned by this rank) first: they can be updated while the guard zones
    // of the boundary blocks are still in flight
    std::stable_partition(impl->regular.begin(), impl->regular.end(), [&] (int b)
    {
        for (int k = 0; k < 9; ++k) if (nbr9[size_t(b) * 9 + k] >= BO) return false;
        return true;
    });
    impl->num_interior = 0;
    for (int b : impl->regular)
    {
        bool interior = true;
        for (int k = 0; k < 9; ++k) if (nbr9[size_t(b) * 9 + k] >= BO) interior = false;
        impl->num_interior += interior;
    }
    num_regular = int(impl->regular.size());

    auto gslot = std::vector<int>(B, -1);
    auto mark = [&] (int l) { if (l >= 0) in_gradient_set[l] = 1; };

    // the general path reads gradients of its own blocks and of every face neighbour they fetch from
    for (int b : impl->irregular)
    {
        mark(b);
        for (int side = 0; side < 4; ++side) for (int q = 0; q < 4; ++q) mark(nbr[size_t(b) * 4 + side].leaf[q]);
    }
    if (force_general)
        for (int b = 0; b < BO; ++b)        // every owned block takes the any-tree kernels: its face neighbours' gradients too (ghosts included)
        {
            mark(b);
            for (int side = 0; side < 4; ++side) for (int q = 0; q < 4; ++q) mark(nbr[size_t(b) * 4 + side].leaf[q]);
        }
    for (int b = 0; b < B; ++b) if (in_gradient_set[b]) { gslot[b] = int(impl->gradient_blocks.size()); impl->gradient_blocks.push_back(b); }
    for (auto& d : nbr) for (int q = 0; q < 4; ++q) d.gs[q] = d.leaf[q] >= 0 ? gslot[d.leaf[q]] : -1;

    impl->mesh.B = B;
    impl->mesh.N = N;
    impl->mesh.FS = cells;
    impl->mesh.GS = impl->gradient_blocks.size() * size_t(N) * N;
    impl->mesh.xv = device_upload(sd.xv);
    impl->mesh.yv = device_upload(sd.yv);
    impl->mesh.spacing = device_upload(spacing);
    {
        std::vector<double> inv(spacing.size());
        for (size_t k = 0; k < inv.size(); ++k) inv[k] = 1.0 / spacing[k];
        impl->mesh.inv_spacing = device_upload(inv);
    }
    impl->mesh.nbr = device_upload(nbr);
    impl->mesh.nbr9 = device_upload(nbr9);
    impl->mesh.gslot = device_upload(gslot);
    impl->mesh.U0 = device_upload(sd.initial_conserved_u);
    impl->mesh.br = device_upload(sd.buffer_rate_field);
    std::vector<unsigned char> tile_flags_host;
    if (impl->tile_x)
    {
        // static per-tile flags: bit 0 = the buffer-zone rate is non-zero somewhere in the tile
        const int tx = impl->tile_x, ty = impl->tile_y, tiles_y = N / ty, tpb = (N / tx) * tiles_y;
        auto flags = std::vector<unsigned char>(size_t(B) * tpb, 0);
        for (int b = 0; b < B; ++b)
            for (int t = 0; t < tpb; ++t)
            {
                bool any = false;
                for (int i = (t / tiles_y) * tx; i < (t / tiles_y + 1) * tx && ! any; ++i)
                    for (int j = (t % tiles_y) * ty; j < (t % tiles_y + 1) * ty; ++j)
                        if (sd.buffer_rate_field[(size_t(b) * N + i) * N + j] != 0.0) { any = true; break; }
                flags[size_t(b) * tpb + t] = any;
            }
        impl->d_tile_flags = device_upload(flags);
        tile_flags_host = flags;
    }
    impl->d_regular = device_upload(impl->regular);
    auto regular_tile_info = std::vector<tile_info_t>();
    if (impl->tile_x)
    {
        const int tpb = (N / impl->tile_x) * (N / impl->tile_y);
        auto info = std::vector<tile_info_t>(impl->regular.size() * tpb);
        for (size_t r = 0; r < impl->regular.size(); ++r)
            for (int t = 0; t < tpb; ++t)
            {
                auto& ti = info[r * tpb + t];
                ti.b = impl->regular[r];
                for (int k = 0; k < 9; ++k) ti.n9[k] = nbr9[size_t(ti.b) * 9 + k];
                ti.flags = tile_flags_host[size_t(ti.b) * tpb + t] | (t << TILE_POS_SHIFT);
                ti.row = int(r) * tpb + t;
            }
        regular_tile_info = info;       // (uploaded below, with the jump blocks' tiles that touch same-level leaves only behind them)
    }
    // (conserved_q with general_only keeps the any-tree kernels: the reference implementation of that variable set on the device)
    impl->jump_strip = N % 32 == 0 && ! tiled_kernel && (sd.conserve_linear_p || ! general_only);
    if (const char* e = std::getenv("M3B_JUMP_STRIP")) impl->jump_strip = impl->jump_strip && std::atoi(e) != 0;
    if (const char* e = std::getenv("M3B_SERIAL_JUMP")) impl->serial_jump = std::atoi(e) != 0;
    if (const char* e = std::getenv("M3B_JUMP_AFTER")) impl->jump_after_regular = std::atoi(e) != 0;
    if (const char* e = std::getenv("M3B_JUMP_MODE0")) impl->jump_mode0 = std::atoi(e) != 0;
    if (impl->jump_strip)
    {
        // the any-tree list (blocks at refinement jumps, or every owned block with general_only) in 16 x 32 strip tiles
        const int tiles_y = N / 32, tpb = (N / 16) * tiles_y;
        auto list = std::vector<int>();
        if (general_only) for (int b = 0; b < BO; ++b) list.push_back(b);
        else list = impl->irregular;
        // A tile of such a block whose 20 x 36 region (tile + two guard layers) lies in same-level leaves only -- its own block
        // and the <= 3 neighbours it touches -- sees exactly what a tile of a regular block sees: guard cells are copies, no
        // prolongation / restriction, no corrected face (mesh_tree_operators.hpp:223-252, scheme.cpp:614-720 touch the block's
        // sides at the jump only).  With a block of 64^2 in 4 x 2 tiles that is half to three quarters of the tiles of a block
        // with one side at a jump: they are appended to the persistent regular kernel's list (stage_tma) and leave
        // stage_strip<.., JUMP>'s, keeping their rows among the block's.  M3B_SPLIT_JUMP=0: every tile of these blocks through JUMP.
        bool split = ! general_only && sd.conserve_linear_p;        // (stage_tma: linear-momentum variables)
        if (const char* e = std::getenv("M3B_STAGE")) split = split && std::string(e) != "strip";
        if (const char* e = std::getenv("M3B_SPLIT_JUMP")) split = split && std::atoi(e) != 0;
        auto info = std::vector<tile_info_t>();
        for (size_t r = 0; r < list.size(); ++r)
            for (int t = 0; t < tpb; ++t)
            {
                tile_info_t ti;
                const int i0 = (t / tiles_y) * 16, j0 = (t % tiles_y) * 32;
                ti.b = list[r];
                ti.row = int(r) * tpb + t;
                ti.flags = tile_flags_host[size_t(ti.b) * tpb + t] | (t << TILE_POS_SHIFT);
                bool same_level_only = split;
                for (int di = -1; di <= 1 && same_level_only; ++di)
                    for (int dj = -1; dj <= 1; ++dj)
                    {
                        const bool touched = (di == 0 || (di < 0 ? i0 == 0 : i0 + 16 == N)) && (dj == 0 || (dj < 0 ? j0 == 0 : j0 + 32 == N));
                        if (touched && (di | dj) && nbr9[size_t(ti.b) * 9 + (di + 1) * 3 + (dj + 1)] < 0) { same_level_only = false; break; }
                    }
                if (same_level_only)
                {
                    for (int k = 0; k < 9; ++k) ti.n9[k] = nbr9[size_t(ti.b) * 9 + k];     // (-1 where there is no such leaf: never touched)
                    ti.flags |= TILE_JUMP_ROWS;
                    regular_tile_info.push_back(ti);
                    ++impl->num_extra_tiles;
                    continue;
                }
                for (int k = 0; k < 9; ++k) ti.n9[k] = ti.b;        // cells beyond the block come through resolve_cell
                if (i0 == 0       && nbr[size_t(ti.b) * 4 + 0].kind == 2) ti.flags |= 2;
                if (i0 + 16 == N  && nbr[size_t(ti.b) * 4 + 1].kind == 2) ti.flags |= 4;
                if (j0 == 0       && nbr[size_t(ti.b) * 4 + 2].kind == 2) ti.flags |= 8;
                if (j0 + 32 == N  && nbr[size_t(ti.b) * 4 + 3].kind == 2) ti.flags |= 16;
                info.push_back(ti);
            }
        impl->num_jump_tiles = int(info.size());
        if (info.empty()) info.push_back(tile_info_t());
        impl->d_jump_tile_info = device_upload(info);
    }
    if (! regular_tile_info.empty()) impl->d_tile_info = device_upload(regular_tile_info);
    impl->d_irregular = device_upload(impl->irregular);
    impl->d_gradient_blocks = device_upload(impl->gradient_blocks);

    auto all_blocks = std::vector<int>(BO);
    for (int b = 0; b < BO; ++b) all_blocks[b] = b;
    impl->owned.push_back(device_upload(all_blocks));

    impl->model.softening_radius2   = sd.softening_radius * sd.softening_radius;
    impl->model.sink_rate           = sd.sink_rate;
    impl->model.sink_inv_2s2        = 1.0 / (sd.sink_radius * sd.sink_radius) / 2.0;
    impl->model.inv_mach2           = 1.0 / sd.mach_number / sd.mach_number;
    impl->model.inv_mach            = 1.0 / sd.mach_number;
    impl->model.alpha               = sd.alpha;
    impl->model.nu                  = sd.nu;
    impl->model.alpha_cutoff_radius = sd.alpha_cutoff_radius;
    impl->model.density_floor       = sd.density_floor;
    impl->model.axisymmetric_cs2    = sd.axisymmetric_cs2;
    impl->model.qmode               = ! sd.conserve_linear_p;
    impl->model.domain_radius       = sd.domain_radius;
    impl->model.gst_suppr_radius2   = sd.gst_suppr_radius * sd.gst_suppr_radius;
    impl->mesh.qmode                = ! sd.conserve_linear_p;

    // ---- scratch
    size_t max_rows = size_t(BO) * std::max(1, (N / std::max(1, impl->tile_x)) * (N / std::max(1, impl->tile_y))) + B;
    M3B_CUDA(cudaMalloc(&impl->d_partials, max_rows * ROW * sizeof(double)));
    M3B_CUDA(cudaMalloc(&impl->d_partials2, max_rows * ROW * sizeof(double)));
    for (auto& p : impl->d_block_rows) M3B_CUDA(cudaMalloc(&p, size_t(BO + 1) * ROW * sizeof(double)));
    // tile of the tiled any-tree kernels: 16 x 16 where it divides the block, else 12 x 12 (the reference's default block size 24),
    // else 8 x 8; block sizes none of them divides keep one CTA per block
    impl->gtile = N % 16 == 0 ? 16 : (N % 12 == 0 ? 12 : (N % 8 == 0 ? 8 : 0));
    if (impl->gtile)
    {
        const int g = impl->gtile;
        for (auto& p : impl->d_general_tile_rows) M3B_CUDA(cudaMalloc(&p, std::max<size_t>(1, size_t(BO) * (N / g) * (N / g)) * ROW * sizeof(double)));
        M3B_CUDA(cudaFuncSetAttribute(general_update_tiled<16, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(tile_t<16, 16>))));
        M3B_CUDA(cudaFuncSetAttribute(general_update_tiled<12, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(tile_t<12, 12>))));
        M3B_CUDA(cudaFuncSetAttribute(general_update_tiled<8, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(tile_t<8, 8>))));
    }
    if (const char* e = std::getenv("M3B_UNTILED_GENERAL")) impl->untiled_general = std::atoi(e) != 0;
    if (const char* e = std::getenv("M3B_MULTI_CTA_FINISH")) impl->multi_cta_finish = std::atoi(e) != 0;
    impl->cta_rows_stride = size_t(BO / FINISH_ROWS_PER_CTA + 1) * ROW;
    M3B_CUDA(cudaMalloc(&impl->d_cta_rows, 2 * impl->cta_rows_stride * sizeof(double)));
    // (default priority: with the highest one the first stage's finish_stage displaces CTAs of the second stage kernel, +2 us)
    M3B_CUDA(cudaStreamCreateWithFlags(&impl->finish_stream, cudaStreamNonBlocking));
    M3B_CUDA(cudaEventCreateWithFlags(&impl->stage_done, cudaEventDisableTiming));
    {
        // above the compute stream's priority: ring gradients -> jump blocks is the critical path of a stage on a nested tree,
        // the regular blocks' CTAs fill the slots it leaves (M3B_JUMP_PRIORITY=0: same priority)
        int least = 0, greatest = 0;
        M3B_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        const char* e = std::getenv("M3B_JUMP_PRIORITY");
        M3B_CUDA(cudaStreamCreateWithPriority(&impl->jump_stream, cudaStreamNonBlocking, e && std::atoi(e) == 0 ? least : greatest));
    }
    M3B_CUDA(cudaEventCreateWithFlags(&impl->gradients_done, cudaEventDisableTiming));
    M3B_CUDA(cudaEventCreateWithFlags(&impl->jump_done, cudaEventDisableTiming));
    M3B_CUDA(cudaEventCreateWithFlags(&impl->side_finish_done, cudaEventDisableTiming));
    M3B_CUDA(cudaEventCreateWithFlags(&impl->fast_prepare_done, cudaEventDisableTiming));
    for (auto& e : impl->positions_done) M3B_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    M3B_CUDA(cudaMalloc(&impl->d_counters, 2 * sizeof(int)));
    M3B_CUDA(cudaMemset(impl->d_counters, 0, 2 * sizeof(int)));
    impl->partial_rows = max_rows;
    M3B_CUDA(cudaMalloc(&impl->d_stage, num_slots * sizeof(stage_t)));
    M3B_CUDA(cudaMallocHost(&impl->h_stage_ring, stage_ring_size * sizeof(stage_t)));
    M3B_CUDA(cudaMalloc(&impl->d_results_local, num_slots * sizeof(stage_result_t)));
    M3B_CUDA(cudaMemset(impl->d_results_local, 0, num_slots * sizeof(stage_result_t)));
    for (auto& e : impl->step_done) M3B_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    {
        // highest priority: the exchange's small kernels must not queue behind the interior update's CTAs
        int least = 0, greatest = 0;
        M3B_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        M3B_CUDA(cudaStreamCreateWithPriority(&impl->comm_stream, cudaStreamNonBlocking, greatest));
    }
    // Measured on 4096^2 (profiles/): +4 % at 4 GPUs but -27 % at 8 GPUs, where the NCCL kernel spins on SMs until
    // the slowest peer arrives and the boundary launch adds a partial wave; off unless asked for.
    if (const char* e = std::getenv("M3B_TRACE")) impl->trace = std::atoi(e) != 0;
    if (const char* e = std::getenv("M3B_OVERLAP_EXCHANGE")) impl->overlap_exchange = std::atoi(e) != 0;
    M3B_CUDA(cudaEventCreateWithFlags(&impl->input_ready, cudaEventDisableTiming));
    M3B_CUDA(cudaEventCreateWithFlags(&impl->halo_ready, cudaEventDisableTiming));
    M3B_CUDA(cudaMalloc(&impl->d_staging, 3 * sd.num_owned_cells() * sizeof(double)));
    M3B_CUDA(cudaMalloc(&impl->d_fail, num_slots * sizeof(fail_dev_t)));
    // stage results live in mapped pinned host memory: finish_stage writes them straight to the host
    M3B_CUDA(cudaHostAlloc(&host_results, num_slots * sizeof(stage_result_t), cudaHostAllocMapped));
    M3B_CUDA(cudaHostGetDevicePointer(&impl->d_results, host_results, 0));
    M3B_CUDA(cudaMemset(impl->d_fail, 0, num_slots * sizeof(fail_dev_t)));
    M3B_CUDA(cudaMalloc(&impl->d_gradients, std::max<size_t>(1, 6 * impl->mesh.GS) * sizeof(double)));
    M3B_CUDA(cudaMalloc(&impl->d_exchange_clock, 6 * sizeof(unsigned long long)));
    M3B_CUDA(cudaMemset(impl->d_exchange_clock, 0, 6 * sizeof(unsigned long long)));
    M3B_CUDA(cudaMalloc(&impl->d_fused_counters, 8 * sizeof(int)));
    M3B_CUDA(cudaMemset(impl->d_fused_counters, 0, 8 * sizeof(int)));
    if (const char* e = std::getenv("M3B_FUSED_EXCHANGE")) impl->fused_exchange = std::atoi(e) != 0;

    // ---- multi-GPU exchange plan (partition.hpp): per peer, the strips in the order both sides agree on
    if (part.is_distributed())
    {
        auto build = [&] (const std::vector<std::vector<halo_region_t>>& lists, std::vector<size_t>& counts, std::vector<size_t>& starts)
        {
            auto entries = std::vector<halo_entry_dev_t>();
            size_t offset = 0;
            counts.assign(part.nranks, 0);
            starts.assign(part.nranks, 0);
            for (int p = 0; p < part.nranks; ++p)
            {
                starts[p] = offset;
                for (const auto& r : lists[p])
                {
                    halo_entry_dev_t e;
                    e.block = r.block;
                    e.i0 = r.di < 0 ? N - 2 : 0; e.ni = r.di ? 2 : N;
                    e.j0 = r.dj < 0 ? N - 2 : 0; e.nj = r.dj ? 2 : N;
                    e.pad = p;              // the other side's rank
                    e.offset = offset;
                    offset += size_t(3) * e.ni * e.nj;
                    entries.push_back(e);
                }
                counts[p] = offset - starts[p];
            }
            return std::make_pair(entries, offset);
        };
        auto send_starts = std::vector<size_t>(), recv_starts = std::vector<size_t>();
        auto [send_entries, send_total] = build(part.send, impl->send_count, send_starts);
        auto [recv_entries, recv_total] = build(part.recv, impl->recv_count, recv_starts);
        impl->num_send_entries = int(send_entries.size());
        impl->num_recv_entries = int(recv_entries.size());
        impl->d_send_entries = device_upload(send_entries);
        impl->d_recv_entries = device_upload(recv_entries);
        impl->send_entries_host = send_entries;
        impl->send_starts_host = send_starts;
        impl->recv_starts_host = recv_starts;
        impl->recv_total = recv_total;
        M3B_CUDA(cudaMalloc(&impl->d_send_buffer, std::max<size_t>(1, send_total) * sizeof(double)));
        M3B_CUDA(cudaMalloc(&impl->d_recv_buffer, std::max<size_t>(1, recv_total) * sizeof(double)));
        for (int p = 0; p < part.nranks; ++p)
        {
            impl->send_ptr.push_back(impl->d_send_buffer + send_starts[p]);
            impl->recv_ptr.push_back(impl->d_recv_buffer + recv_starts[p]);
        }
        impl->halo_bytes_per_exchange = send_total * sizeof(double);
        M3B_CUDA(cudaMalloc(&impl->d_results_all, size_t(part.nranks) * num_slots * sizeof(stage_result_t)));
        M3B_CUDA(cudaMallocHost(&impl->h_results_all, size_t(part.nranks) * num_slots * sizeof(stage_result_t)));
        std::memset(impl->h_results_all, 0, size_t(part.nranks) * num_slots * sizeof(stage_result_t));
    }
    num_ranks = part.nranks;
    rank_ = part.rank;

    auto set_smem = [&] (auto kernel, size_t bytes)
    {
        impl->fused_smem = bytes;
        M3B_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)));
    };
    if (impl->tile_x == 16 && impl->tile_y == 32) set_smem(stage_fused<16, 32>, sizeof(tile_t<16, 32>));
    if (impl->strip)
    {
        // (3 CTAs x 168 registers and 2 x 220 were measured slower than 4 x 128: DESIGN.md section 3)
        const char* pa = std::getenv("M3B_PREFETCH_AHEAD");
        impl->mesh.prefetch_ahead = pa ? std::atoi(pa) : impl->sm_count * 4;
        set_smem(stage_strip<4, 0, false, 0>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 0, true, 0>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 64, false, 0>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 64, true, 0>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 64, true, 1>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 64, true, 2>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 0, false, 0, false, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 0, true, 0, false, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 64, false, 0, false, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 64, true, 0, false, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 64, true, 1, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 64, true, 2, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 0, false, 0, true, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 0, true, 0, true, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 64, false, 0, true, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 64, true, 0, true, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 0, false, 0, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 0, true, 0, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 64, false, 0, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 64, true, 0, true>, sizeof(strip_smem_t));
        impl->fast_eos = ! sd.axisymmetric_cs2 && sd.nu == 0.0 && sd.alpha_cutoff_radius == 0.0 && sd.density_floor == 0.0;
        // stage_tma: linear-momentum variables only (conserved_q keeps stage_strip<.., QMODE>)
        impl->tma = sd.conserve_linear_p;
        if (const char* e = std::getenv("M3B_STAGE")) impl->tma = impl->tma && std::string(e) != "strip";
        if (const char* e = std::getenv("M3B_TMA_CTAS")) impl->tma_ctas_per_sm = std::max(1, std::min(4, std::atoi(e)));
        impl->tma_fast = impl->fast_eos && sd.alpha > 0.0;
        stage_tma_configure();      // stage_tma.cu
    }
    if (impl->tile_x == 12 && impl->tile_y == 24) set_smem(stage_fused<12, 24>, sizeof(tile_t<12, 24>));
    if (impl->tile_x == 16 && impl->tile_y == 16) set_smem(stage_fused<16, 16>, sizeof(tile_t<16, 16>));
    if (impl->tile_x == 8  && impl->tile_y == 8)  set_smem(stage_fused<8, 8>, sizeof(tile_t<8, 8>));
}

device_solver_t::~device_solver_t()
{
    cudaSetDevice(device_id);
    cudaStreamSynchronize(cudaStream_t(stream_));
    for (auto p : {(void*) impl->mesh.xv, (void*) impl->mesh.yv, (void*) impl->mesh.spacing, (void*) impl->mesh.inv_spacing, (void*) impl->mesh.nbr,
                   (void*) impl->mesh.nbr9, (void*) impl->mesh.gslot, (void*) impl->mesh.U0, (void*) impl->mesh.br,
                   (void*) impl->d_regular, (void*) impl->d_irregular, (void*) impl->d_gradient_blocks, (void*) impl->d_gradients,
                   (void*) impl->d_partials, (void*) impl->d_staging, (void*) impl->d_tile_flags, (void*) impl->d_tile_info, (void*) impl->d_jump_tile_info, (void*) impl->d_fail})
        if (p) cudaFree(p);
    for (auto p : impl->owned) cudaFree(p);
    for (auto p : {(void*) impl->d_send_entries, (void*) impl->d_recv_entries, (void*) impl->d_send_buffer, (void*) impl->d_recv_buffer,
                   (void*) impl->d_results_local, (void*) impl->d_results_all})
        if (p) cudaFree(p);
    if (impl->trace && impl->trace_steps.size() > 20)
    {
        const size_t n = impl->trace_names.size(), first = 10;
        auto acc = std::vector<double>(n, 0.0);
        size_t used = 0;
        for (size_t k = first; k < impl->trace_steps.size(); ++k)
        {
            auto& ev = impl->trace_steps[k];
            if (ev.size() != n) continue;
            ++used;
            for (size_t m = 1; m < n; ++m) { float ms = 0; cudaEventElapsedTime(&ms, ev[m - 1], ev[m]); acc[m] += ms; }
            if (k + 1 < impl->trace_steps.size() && impl->trace_steps[k + 1].size() == n)
            { float ms = 0; cudaEventElapsedTime(&ms, ev[n - 1], impl->trace_steps[k + 1][0]); acc[0] += ms; }
        }
        std::fprintf(stderr, "[m3b trace rank %d] %zu steps, us per step:\n  %-28s %8.1f\n", rank_, used, "(gap to next step)", acc[0] / used * 1e3);
        for (size_t m = 1; m < n; ++m) std::fprintf(stderr, "  -> %-25s %8.1f\n", impl->trace_names[m], acc[m] / used * 1e3);
    }
    for (auto p : impl->peer_mailbox) if (p) cudaIpcCloseMemHandle(p);
    for (auto p : {(void*) impl->mailbox, (void*) impl->d_push_entries, (void*) impl->d_push_ticket, (void*) impl->d_ready, (void*) impl->d_exchange_clock, (void*) impl->d_fused_counters}) if (p) cudaFree(p);
    if (impl->h_results_all) cudaFreeHost(impl->h_results_all);
    if (impl->h_stage_ring) cudaFreeHost(impl->h_stage_ring);
    for (auto p : {(void*) impl->d_stage, (void*) impl->d_partials2, (void*) impl->d_block_rows[0], (void*) impl->d_block_rows[1],
                   (void*) impl->d_cta_rows, (void*) impl->d_counters, (void*) impl->d_general_tile_rows[0], (void*) impl->d_general_tile_rows[1]}) if (p) cudaFree(p);
    for (auto e : impl->step_done) if (e) cudaEventDestroy(e);
    if (impl->stage_done) cudaEventDestroy(impl->stage_done);
    if (impl->side_finish_done) cudaEventDestroy(impl->side_finish_done);
    if (impl->fast_prepare_done) cudaEventDestroy(impl->fast_prepare_done);
    for (auto e : impl->positions_done) if (e) cudaEventDestroy(e);
    if (impl->finish_stream) cudaStreamDestroy(impl->finish_stream);
    if (impl->input_ready) cudaEventDestroy(impl->input_ready);
    if (impl->halo_ready) cudaEventDestroy(impl->halo_ready);
    if (impl->comm_stream) cudaStreamDestroy(impl->comm_stream);
    for (auto& ev : impl->timing_events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    for (auto e : impl->event_pool) cudaEventDestroy(e);
    if (host_results) cudaFreeHost(host_results);
    cudaStreamDestroy(cudaStream_t(own_stream));
}

void device_solver_t::set_stream(void* cuda_stream)
{
    sync();
    stream_ = cuda_stream ? cuda_stream : own_stream;
}

void device_solver_t::upload(const double* host, device_field_t& dst)
{
    auto s = cudaStream_t(stream_);
    M3B_CUDA(cudaSetDevice(device_id));
    M3B_CUDA(cudaMemcpyAsync(impl->d_staging, host, 3 * size_t(BO) * N * N * sizeof(double), cudaMemcpyHostToDevice, s));
    permute_state<<<impl->sm_count * 4, 256, 0, s>>>(impl->d_staging, dst.data, BO, N * N, cells, 1);
    ++launches;
    M3B_CUDA(cudaGetLastError());
}

void device_solver_t::download(const device_field_t& src, double* host)
{
    auto s = cudaStream_t(stream_);
    M3B_CUDA(cudaSetDevice(device_id));
    permute_state<<<impl->sm_count * 4, 256, 0, s>>>(src.data, impl->d_staging, BO, N * N, cells, 0);
    ++launches;
    M3B_CUDA(cudaGetLastError());
    M3B_CUDA(cudaMemcpyAsync(host, impl->d_staging, 3 * size_t(BO) * N * N * sizeof(double), cudaMemcpyDeviceToHost, s));
    M3B_CUDA(cudaStreamSynchronize(s));
}




/** [block][3][N][N] of every rank's owned blocks on rank 0 (the layout of download()). */
void device_solver_t::gather_state(const device_field_t& src, double* host_all)
{
    M3B_CUDA(cudaSetDevice(device_id));
    permute_state<<<impl->sm_count * 4, 256, 0, cudaStream_t(stream_)>>>(src.data, impl->d_staging, BO, N * N, cells, 0);
    ++launches;
    gather_blocks(impl->d_staging, size_t(3) * N * N, host_all);
}

void device_solver_t::gather_diagnostic_fields(const device_field_t& src, double* host_all)
{
    M3B_CUDA(cudaSetDevice(device_id));
    diagnostic_fields_kernel<<<impl->sm_count * 4, 256, 0, cudaStream_t(stream_)>>>(impl->mesh, src.data, impl->d_staging, BO);
    ++launches;
    gather_blocks(impl->d_staging, size_t(3) * N * N, host_all);
}

void device_solver_t::disk_totals(const device_field_t& src, double out[2])
{
    M3B_CUDA(cudaSetDevice(device_id));
    auto s = cudaStream_t(stream_);
    double* d = impl->d_staging;                    // [BO][2], idle between transfers
    disk_totals_kernel<<<BO, THREADS, 0, s>>>(impl->mesh, src.data, d);
    ++launches;
    auto h = std::vector<double>(size_t(BO) * 2);
    M3B_CUDA(cudaMemcpyAsync(h.data(), d, h.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    M3B_CUDA(cudaStreamSynchronize(s));
    out[0] = out[1] = 0.0;
    for (int b = 0; b < BO; ++b) { out[0] += h[2 * b]; out[1] += h[2 * b + 1]; }    // tree order, as the reference's .sum()
    if (num_ranks > 1)
    {
        // per-rank sums folded in rank order on every rank
        for (int k = 0; k < 2; ++k)
        {
            auto parts = all_gather_scalar(out[k]);
            out[k] = 0.0;
            for (double v : parts) out[k] += v;
        }
    }
}

void device_solver_t::diagnostic_fields(const device_field_t& src, double* host)
{
    M3B_CUDA(cudaSetDevice(device_id));
    auto s = cudaStream_t(stream_);
    diagnostic_fields_kernel<<<impl->sm_count * 4, 256, 0, s>>>(impl->mesh, src.data, impl->d_staging, BO);
    ++launches;
    M3B_CUDA(cudaMemcpyAsync(host, impl->d_staging, size_t(BO) * 3 * N * N * sizeof(double), cudaMemcpyDeviceToHost, s));
    M3B_CUDA(cudaStreamSynchronize(s));
}

void device_solver_t::copy(const device_field_t& src, device_field_t& dst)
{
    M3B_CUDA(cudaSetDevice(device_id));
    M3B_CUDA(cudaMemcpyAsync(dst.data, src.data, 3 * cells * sizeof(double), cudaMemcpyDeviceToDevice, cudaStream_t(stream_)));
}

void device_solver_t::load_initial(device_field_t& dst)
{
    M3B_CUDA(cudaSetDevice(device_id));
    M3B_CUDA(cudaMemcpyAsync(dst.data, impl->mesh.U0, 3 * cells * sizeof(double), cudaMemcpyDeviceToDevice, cudaStream_t(stream_)));
}

void device_solver_t::combine(const device_field_t& a, double wa, const device_field_t& b, double wb, device_field_t& dst)
{
    M3B_CUDA(cudaSetDevice(device_id));
    axpby_kernel<<<impl->sm_count * 4, 256, 0, cudaStream_t(stream_)>>>(a.data, wa, b.data, wb, dst.data, 3 * cells);
    ++launches;
    M3B_CUDA(cudaGetLastError());
}

void device_solver_t::set_stage_timing(bool on)
{
    if (on && impl->event_pool.size() < 512)
    {
        for (int k = 0; k < 512; ++k) { cudaEvent_t e; M3B_CUDA(cudaEventCreate(&e)); impl->event_pool.push_back(e); }
    }
    stage_timing = on;
}

void device_solver_t::collect_stage_timing()
{
    sync();
    for (auto& ev : impl->timing_events)
    {
        float ms = 0.f;
        M3B_CUDA(cudaEventElapsedTime(&ms, ev.first, ev.second));
        stage_ms_total += ms;
        ++stage_timed_launches;
        impl->event_pool.push_back(ev.first);
        impl->event_pool.push_back(ev.second);
    }
    impl->timing_events.clear();
    auto fold = [&] (std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& list, double& total_us, std::uint64_t& count)
    {
        for (auto& ev : list)
        {
            float ms = 0.f;
            M3B_CUDA(cudaEventElapsedTime(&ms, ev.first, ev.second));
            total_us += ms * 1e3;
            ++count;
            impl->event_pool.push_back(ev.first);
            impl->event_pool.push_back(ev.second);
        }
        list.clear();
    };
    fold(impl->exchange_events, exchange_us_total, exchanges_timed);
    std::uint64_t gaps = 0;
    fold(impl->gap_events, exposed_wait_us_total, gaps);
    unsigned long long words[6] = {0, 0, 0, 0, 0, 0};
    M3B_CUDA(cudaMemcpy(words, impl->d_exchange_clock, sizeof(words), cudaMemcpyDeviceToHost));
    M3B_CUDA(cudaMemset(impl->d_exchange_clock, 0, sizeof(words)));
    result_wait_us_total += words[0] * 1e-3;
    result_waits_timed += words[1];
    unpack_cta_wait_us_total += words[2] * 1e-3;
    unpack_cta_waits += words[3];
    // fused exchange: kernel start -> all strips stored in the neighbours' landing buffers and the flags raised
    exchange_us_total += words[4] * 1e-3;
    exchanges_timed += words[5];
}

void device_solver_t::upload_stage(const stage_inputs_t& inputs, int slot)
{
    auto s = cudaStream_t(stream_);
    stage_t& st = impl->h_stage_ring[impl->ring_next];      // the ring is far longer than the launches in flight
    impl->ring_next = (impl->ring_next + 1) % stage_ring_size;
    st.time = inputs.time;
    st.dt = inputs.dt;
    st.theta = inputs.theta;
    st.x1 = inputs.bodies.body1.x; st.y1 = inputs.bodies.body1.y; st.m1 = inputs.bodies.body1.mass;
    st.x2 = inputs.bodies.body2.x; st.y2 = inputs.bodies.body2.y; st.m2 = inputs.bodies.body2.mass;
    st.vx1 = inputs.bodies.body1.vx; st.vy1 = inputs.bodies.body1.vy;
    st.vx2 = inputs.bodies.body2.vx; st.vy2 = inputs.bodies.body2.vy;
    st.rk_b0 = inputs.rk_b0;
    st.combine = inputs.combine;
    st.compute_dt = inputs.compute_dt;
    M3B_CUDA(cudaMemcpyAsync(impl->d_stage + slot, &st, sizeof(stage_t), cudaMemcpyHostToDevice, s));
}

/** The stage kernels + finish_stage for the inputs already in d_stage[slot].  With `exchange` the guard
 *  zones of `in` are refreshed from the other ranks first, overlapped with the update of the interior blocks. */
void device_solver_t::launch_stage_kernels(const device_field_t& in, const device_field_t* un, device_field_t& out, int slot, bool exchange, int finish_mode, int stage_mode)
{
    auto s = cudaStream_t(stream_);
    if (in.data == out.data) throw std::invalid_argument("launch_stage: in-place stages are not supported");

    const stage_t* st = impl->d_stage + slot;
    double* partials = (slot & 1) ? impl->d_partials2 : impl->d_partials;
    double* block_rows = impl->d_block_rows[slot & 1];
    const double* un_data = un ? un->data : nullptr;
    int num_fused = force_general ? 0 : int(impl->regular.size());
    int num_general = force_general ? BO : int(impl->irregular.size());
    const int* d_general = force_general ? static_cast<const int*>(impl->owned[0]) : impl->d_irregular;
    const int tpb = impl->tile_x ? (N / impl->tile_x) * (N / impl->tile_y) : 0;
    int fused_ctas = num_fused * tpb;
    // the any-tree path in 16 x 16 tiles where the block size allows (one row per tile), else one CTA and one row per block
    const bool jump_strip = impl->jump_strip && impl->strip && ! impl->untiled_general;
    const int gt = impl->untiled_general ? 0 : impl->gtile;
    const int ggtpb = gt ? (N / gt) * (N / gt) : 1;     // tiles of general_gradients_tiled
    const int gtpb = jump_strip ? tpb : ggtpb;
    double* general_rows = gtpb > 1 ? impl->d_general_tile_rows[slot & 1] : block_rows;
    exchange = exchange && num_ranks > 1;
    impl->mark(s, "stage begin");
    bool waiting_tiles = false;
    // stage_tma is persistent: its CTAs hold every register of the SM until the tile list is done, so a tile that waited inside
    // the kernel for the guard-zone unpack could keep the unpack kernel from ever becoming resident.  Its launch is split
    // instead: interior blocks beside the exchange (which runs on its own, higher-priority stream), then the blocks with
    // ghost neighbours once the unpack has finished.  No kernel of this path waits for another kernel of the same GPU.
    const bool tma_exchange = exchange && impl->strip && impl->tma && impl->peer_transport && num_general == 0 && num_fused > 0;
    // (default) the stage kernel does the exchange itself: push first, interior blocks, unpack by the CTAs that reach the blocks
    // with ghost neighbours first -- one launch, nothing on another stream
    const bool fused = tma_exchange && impl->fused_exchange;
    fused_exchange_t X = fused_exchange_t();
    if (fused)
    {
        const unsigned long long counter = ++impl->exchange_counter;
        X.enabled = 1;
        X.push = impl->d_push_entries; X.n_push = impl->num_send_entries;
        X.recv = impl->d_recv_entries; X.n_recv = impl->num_recv_entries;
        X.peers = impl->peers; X.parity = int(counter & 1); X.me = rank_; X.dest_mask = impl->dest_mask; X.counter = counter;
        X.counters = impl->d_fused_counters; X.cset = int(++impl->fused_launches & 1); X.U = const_cast<double*>(in.data);
        X.clock_words = stage_timing ? impl->d_exchange_clock : nullptr;
        exchange = false;
    }
    const bool split_launch = tma_exchange && ! fused && impl->num_recv_entries > 0 && impl->num_interior > 0;
    if (exchange && ! impl->overlap_exchange && ! split_launch)
    {
        waiting_tiles = impl->peer_transport && impl->in_kernel_wait && impl->strip && ! impl->tma && num_general == 0 && impl->num_recv_entries > 0;
        impl->defer_unpack = waiting_tiles;
        exchange_on(stream_, const_cast<device_field_t&>(in));     // plain ordering: exchange, then every block
        impl->defer_unpack = false;
        exchange = false;
        impl->mark(s, "exchange done");
    }

    // blocks [first, first + count) of the regular list: tile rows, block rows and tickets are indexed by list position;
    // extra_tiles: the jump blocks' tiles that touch same-level leaves only (they follow the regular blocks' tiles in d_tile_info,
    // so they can only ride with a launch that ends with the last regular block)
    const int num_extra_tiles = (impl->strip && impl->tma && jump_strip && num_general > 0) ? impl->num_extra_tiles : 0;
    auto launch_fused = [&] (int first, int count, int extra_tiles = 0)
    {
        if (count <= 0 && extra_tiles <= 0) return;
        if (extra_tiles > 0 && first + count != num_fused) throw std::logic_error("launch_fused: the extra tiles follow the last regular block");
        const int ctas = count * tpb + extra_tiles;
        const int* list = impl->d_regular + first;
        double* tiles = partials + size_t(first) * tpb * ROW;
        #define M3B_LAUNCH_FUSED(TX, TY) stage_fused<TX, TY><<<ctas, THREADS, sizeof(tile_t<TX, TY>), s>>>( \
            impl->mesh, impl->model, st, list, in.data, un_data, out.data, tiles, impl->d_fail + slot)
        if (impl->strip && impl->tma)
        {
            stage_tma_launch_t a;
            a.mesh = impl->mesh;
            a.mesh.first_wait_cta = (waiting_tiles || fused) ? std::max(0, impl->num_interior - first) * tpb : 0x7fffffff;
            a.exchange = X;
            a.mesh.ready_flag = impl->d_ready;
            a.mesh.ready_value = impl->exchange_counter;
            a.model = impl->model; a.stage = st; a.tile_info = impl->d_tile_info + size_t(first) * tpb; a.num_tiles = ctas;
            a.Uin = in.data; a.Un = un_data; a.Uout = out.data; a.fail = impl->d_fail + slot;
            a.partials = partials; a.jump_partials = general_rows;      // (rows are addressed through tile_info_t::row)
            a.N = N; a.fast = impl->tma_fast; a.stage_mode = stage_mode;
            // 3 CTAs per SM with two tile buffers unless M3B_TMA_CTAS=4 asks for one buffer and 4: measured on 4096^2, a tile costs a
            // 4-per-SM CTA 1.36 x what it costs a 3-per-SM CTA (10.6 against 7.8 us, its load is exposed), so 589 against 578 us
            // on one GPU; on 2 GPUs 0.680 against 0.634 ms per step (the CTAs also spend 19 instead of 5 us in the fused unpack).
            const int per_sm = impl->tma_ctas_per_sm ? impl->tma_ctas_per_sm : 3;
            a.ctas_per_sm = per_sm;
            a.grid = std::min(ctas, impl->sm_count * per_sm);       // persistent: every CTA walks the tile list with stride `grid`
            stage_tma_launch(a, s);
        }
        else if (impl->strip)
        {
            auto kernel = N == 64 ? (impl->fast_eos ? stage_strip<4, 64, true, 0> : stage_strip<4, 64, false, 0>)
                                  : (impl->fast_eos ? stage_strip<4, 0, true, 0> : stage_strip<4, 0, false, 0>);
            if (N == 64 && impl->fast_eos && stage_mode == 1) kernel = stage_strip<4, 64, true, 1>;
            if (N == 64 && impl->fast_eos && stage_mode == 2) kernel = stage_strip<4, 64, true, 2>;
            if (impl->mesh.qmode)       // conserved_q = (sigma, Sr, Lz): advance_q (scheme.cpp:906-1020) through the same kernel
                kernel = N == 64 ? (impl->fast_eos ? stage_strip<4, 64, true, 0, false, true> : stage_strip<4, 64, false, 0, false, true>)
                                 : (impl->fast_eos ? stage_strip<4, 0, true, 0, false, true> : stage_strip<4, 0, false, 0, false, true>);
            mesh_dev_t mesh = impl->mesh;
            mesh.first_wait_cta = waiting_tiles ? std::max(0, impl->num_interior - first) * tpb : 0x7fffffff;
            mesh.ready_flag = impl->d_ready;
            mesh.ready_value = impl->exchange_counter;
            kernel<<<ctas, STRIP_THREADS, sizeof(strip_smem_t), s>>>(mesh, impl->model, st, impl->d_tile_info + size_t(first) * tpb,
                in.data, un_data, out.data, partials, impl->d_fail + slot, nullptr);
        }
        else if (impl->tile_x == 16 && impl->tile_y == 32) M3B_LAUNCH_FUSED(16, 32);
        else if (impl->tile_x == 12 && impl->tile_y == 24) M3B_LAUNCH_FUSED(12, 24);
        else if (impl->tile_x == 16 && impl->tile_y == 16) M3B_LAUNCH_FUSED(16, 16);
        else                                               M3B_LAUNCH_FUSED(8, 8);
        #undef M3B_LAUNCH_FUSED
        ++launches;
        M3B_CUDA(cudaGetLastError());
    };

    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (stage_timing && (fused_ctas > 0 || num_general > 0))
    {
        // events come from a pool: creating them here would delay the launches behind this one
        if (impl->event_pool.size() < 2)
        {
            for (int k = 0; k < 64; ++k) { cudaEvent_t e; M3B_CUDA(cudaEventCreate(&e)); impl->event_pool.push_back(e); }
        }
        e0 = impl->event_pool.back(); impl->event_pool.pop_back();
        e1 = impl->event_pool.back(); impl->event_pool.pop_back();
    }
    // Blocks at refinement jumps through the strip kernel: their launch runs on its own stream beside the regular blocks'
    // (two partial waves of CTAs fill each other's tails); the gradients both read come first.
    auto launch_jump_strip = [&] (cudaStream_t js)
    {
        auto kernel = N == 64 ? (impl->fast_eos ? stage_strip<4, 64, true, 0, true> : stage_strip<4, 64, false, 0, true>)
                              : (impl->fast_eos ? stage_strip<4, 0, true, 0, true> : stage_strip<4, 0, false, 0, true>);
        if (N == 64 && impl->fast_eos && stage_mode == 1 && ! impl->jump_mode0) kernel = stage_strip<4, 64, true, 1, true>;
        if (N == 64 && impl->fast_eos && stage_mode == 2 && ! impl->jump_mode0) kernel = stage_strip<4, 64, true, 2, true>;
        if (impl->mesh.qmode)
            kernel = N == 64 ? (impl->fast_eos ? stage_strip<4, 64, true, 0, true, true> : stage_strip<4, 64, false, 0, true, true>)
                             : (impl->fast_eos ? stage_strip<4, 0, true, 0, true, true> : stage_strip<4, 0, false, 0, true, true>);
        mesh_dev_t mesh = impl->mesh;
        mesh.first_wait_cta = 0x7fffffff;
        const int jump_tiles = force_general ? num_general * tpb : impl->num_jump_tiles;
        if (jump_tiles == 0) return;
        kernel<<<jump_tiles, STRIP_THREADS, sizeof(strip_smem_t), js>>>(mesh, impl->model, st, impl->d_jump_tile_info,
            in.data, un_data, out.data, general_rows, impl->d_fail + slot, impl->d_gradients);
        ++launches;
        M3B_CUDA(cudaGetLastError());
    };
    const int ng = int(impl->gradient_blocks.size());
    bool jump_forked = false, jump_pending = false, e0_recorded = false;
    if (jump_strip && num_general > 0 && ! exchange)
    {
        // fork: gradients and the jump blocks' update on the side stream, the regular blocks' update on the compute stream
        const bool fork = (num_fused > 0 || num_extra_tiles > 0) && ! impl->serial_jump;
        cudaStream_t js = fork ? impl->jump_stream : s;
        if (e0) { M3B_CUDA(cudaEventRecord(e0, s)); e0_recorded = true; }      // the timed region covers the jump blocks' kernels
        if (fork)
        {
            M3B_CUDA(cudaEventRecord(impl->gradients_done, s));
            M3B_CUDA(cudaStreamWaitEvent(js, impl->gradients_done, 0));
        }
        general_gradients_ring<<<ng * 4, 128, size_t(12) * (N + 2) * sizeof(double), js>>>(impl->mesh, st, impl->d_gradient_blocks, in.data, impl->d_gradients);
        ++launches;
        if (fork)
        {
            // jump_after_regular: only the ring gradients run beside the persistent regular kernel (whose CTAs hold their SMs
            // until its tile list is done); the tiles at the jumps follow on the compute stream once both are through
            if (impl->jump_after_regular) jump_pending = true; else launch_jump_strip(js);
            M3B_CUDA(cudaEventRecord(impl->jump_done, js));
            jump_forked = true;
        }
    }
    if (exchange)
    {
        // guard zones travel on their own stream while the interior blocks are updated
        auto& field = const_cast<device_field_t&>(in);
        auto pooled = [&] { cudaEvent_t e = impl->event_pool.back(); impl->event_pool.pop_back(); return e; };
        const bool timed = stage_timing && impl->event_pool.size() >= 8;
        M3B_CUDA(cudaEventRecord(impl->input_ready, s));
        M3B_CUDA(cudaStreamWaitEvent(impl->comm_stream, impl->input_ready, 0));
        cudaEvent_t x0 = nullptr, x1 = nullptr, g0 = nullptr, g1 = nullptr;
        if (timed) { x0 = pooled(); x1 = pooled(); g0 = pooled(); g1 = pooled(); M3B_CUDA(cudaEventRecord(x0, impl->comm_stream)); }
        exchange_on(impl->comm_stream, field);
        M3B_CUDA(cudaEventRecord(impl->halo_ready, impl->comm_stream));
        if (timed) M3B_CUDA(cudaEventRecord(x1, impl->comm_stream));
        if (e0 && ! e0_recorded) M3B_CUDA(cudaEventRecord(e0, s));
        launch_fused(0, impl->num_interior);
        if (timed) M3B_CUDA(cudaEventRecord(g0, s));
        M3B_CUDA(cudaStreamWaitEvent(s, impl->halo_ready, 0));
        if (timed) M3B_CUDA(cudaEventRecord(g1, s));
        launch_fused(impl->num_interior, num_fused - impl->num_interior, num_extra_tiles);
        if (timed) { impl->exchange_events.emplace_back(x0, x1); impl->gap_events.emplace_back(g0, g1); }
        if (num_fused == 0) {}      // (general blocks below run after the wait as well)
    }
    else
    {
        if (e0 && ! e0_recorded) M3B_CUDA(cudaEventRecord(e0, s));
        launch_fused(0, num_fused, num_extra_tiles);
    }
    if (jump_forked)
    {
        M3B_CUDA(cudaStreamWaitEvent(s, impl->jump_done, 0));
        if (jump_pending) launch_jump_strip(s);
    }
    else if (num_general > 0 && jump_strip)
    {
        if (exchange)       // (overlapped exchange: the guard zones have only just arrived)
        {
            general_gradients_ring<<<ng * 4, 128, size_t(12) * (N + 2) * sizeof(double), s>>>(impl->mesh, st, impl->d_gradient_blocks, in.data, impl->d_gradients);
            ++launches;
        }
        launch_jump_strip(s);
    }
    else if (num_general > 0)
    {
        // gradients are needed for the general blocks and every block they can fetch from
        #define M3B_GENERAL_TILED(G_) do { \
            general_gradients_tiled<G_, G_><<<ng * ggtpb, THREADS, 0, s>>>(impl->mesh, st, impl->d_gradient_blocks, in.data, impl->d_gradients); \
            general_update_tiled<G_, G_><<<num_general * gtpb, THREADS, sizeof(tile_t<G_, G_>), s>>>(impl->mesh, impl->model, st, d_general, \
                in.data, impl->d_gradients, un_data, out.data, general_rows, impl->d_fail + slot); } while (0)
        if (ggtpb > 1 && gtpb > 1)
        {
            if (gt == 16) M3B_GENERAL_TILED(16); else if (gt == 12) M3B_GENERAL_TILED(12); else M3B_GENERAL_TILED(8);
        }
        else
        {
            general_gradients<<<ng, THREADS, 0, s>>>(impl->mesh, st, impl->d_gradient_blocks, in.data, impl->d_gradients);
            general_update<<<num_general, THREADS, 0, s>>>(impl->mesh, impl->model, st, d_general,
                in.data, impl->d_gradients, un_data, out.data, general_rows, impl->d_fail + slot);
        }
        #undef M3B_GENERAL_TILED
        launches += 2;
        M3B_CUDA(cudaGetLastError());
    }
    if (e0)         // after the join: regular blocks, jump blocks and their gradients
    {
        M3B_CUDA(cudaEventRecord(e1, s));
        impl->timing_events.emplace_back(e0, e1);
    }
    if (waiting_tiles) M3B_CUDA(cudaStreamWaitEvent(s, impl->halo_ready, 0));     // join the exchange stream
    impl->mark(s, "stage kernels done");
    launch_finish(partials, num_fused, tpb, general_rows, gtpb, num_fused + num_general, slot, finish_mode);
    impl->mark(s, "finish done (or forked)");
    M3B_CUDA(cudaGetLastError());
}

/** finish_mode 0: on the compute stream.  1: on the side stream, beside the next stage (its result is only read at the end of
 *  the step).  2: on the compute stream after the side stream's finish, with the next step's stage inputs written by the last CTA. */
void device_solver_t::launch_finish(const double* tile_rows, int num_fused, int tpb, const double* general_rows, int gtpb, int num_rows, int slot, int finish_mode)
{
    auto s = cudaStream_t(stream_);
    int ctas = std::max(1, (num_rows + FINISH_ROWS_PER_CTA - 1) / FINISH_ROWS_PER_CTA);
    double* cta_rows = impl->d_cta_rows + size_t(slot & 1) * impl->cta_rows_stride;
    prepare_args_t prep = prepare_args_t();

    if (finish_mode == 1)
    {
        M3B_CUDA(cudaEventRecord(impl->stage_done, s));
        M3B_CUDA(cudaStreamWaitEvent(impl->finish_stream, impl->stage_done, 0));
        s = impl->finish_stream;
    }
    if (finish_mode == 2)
    {
        M3B_CUDA(cudaStreamWaitEvent(s, impl->side_finish_done, 0));
        prep = impl->pending_prepare;
    }
    // on the critical path (not beside a stage kernel, finish_mode 1, where eight 1024-thread CTAs would wait for SMs to drain)
    if (num_rows <= FINISH_CLUSTER_MAX_ROWS && ! impl->multi_cta_finish && finish_mode != 1)
        finish_stage_cluster<<<FINISH_CLUSTER, FINISH_CLUSTER_THREADS, 0, s>>>(tile_rows, num_fused, tpb, general_rows, gtpb,
            num_rows, impl->d_stage + slot, impl->d_fail + slot, result_target(slot), prep);
    else
        finish_stage<<<ctas, FINISH_THREADS, 0, s>>>(tile_rows, num_fused, tpb, general_rows, gtpb, num_rows, cta_rows,
            impl->d_counters + (slot & 1), impl->d_stage + slot, impl->d_fail + slot, result_target(slot), prep);
    ++launches;
    M3B_CUDA(cudaGetLastError());
    if (finish_mode == 1) M3B_CUDA(cudaEventRecord(impl->side_finish_done, impl->finish_stream));
}

/** Where finish_stage writes: host-mapped memory for synchronous single-rank use, device memory when the
 *  result still has to be folded over ranks or consumed by prepare_next. */
stage_result_t* device_solver_t::result_target(int slot)
{
    bool on_device = num_ranks > 1;     // a single rank publishes straight to the host (finish_stage also prepares the next step)
    return (on_device ? impl->d_results_local : impl->d_results) + slot;
}

void device_solver_t::launch_stage(const device_field_t& in, const device_field_t* un, device_field_t& out, const stage_inputs_t& inputs, int slot)
{
    M3B_CUDA(cudaSetDevice(device_id));
    if (inputs.combine && ! un) throw std::invalid_argument("launch_stage: combine requires the step-start state");
    upload_stage(inputs, slot);
    launch_stage_kernels(in, un, out, slot, /*exchange*/ true);
}

void device_solver_t::launch_step_async(device_field_t& in, device_field_t& scratch, device_field_t& out, int parity,
                                        const elements_t& elements, double cfl_number, double recommended_time_step, double theta, bool fixed_dt)
{
    auto s = cudaStream_t(stream_);
    M3B_CUDA(cudaSetDevice(device_id));
    const int a = first_async_slot + 2 * parity, b = a + 1;
    const int na = first_async_slot + 2 * (1 - parity), nb = na + 1;

    step_config_t cfg;
    cfg.elements = elements;
    cfg.cfl_number = cfl_number;
    cfg.recommended_time_step = recommended_time_step;
    cfg.theta = theta;
    cfg.fixed_dt = fixed_dt;

    const bool split_prepare = num_ranks == 1 || impl->peer_transport;
    if (split_prepare)
    {
        if (impl->fresh_pipeline)
        {
            // the next step's first stage needs its body positions: normally prepared a step ahead, here for the first time
            // (after the host's upload of this step's inputs, which is queued on the compute stream)
            M3B_CUDA(cudaEventRecord(impl->fast_prepare_done, s));
            M3B_CUDA(cudaStreamWaitEvent(impl->finish_stream, impl->fast_prepare_done, 0));
            prepare_positions<<<1, 32, 0, impl->finish_stream>>>(cfg, impl->d_stage + a, nullptr, impl->d_stage + na);
            ++launches;
            M3B_CUDA(cudaEventRecord(impl->positions_done[1 - parity], impl->finish_stream));
        }
        M3B_CUDA(cudaStreamWaitEvent(s, impl->positions_done[parity], 0));         // positions of this step's first stage
    }
    impl->fresh_pipeline = false;

    if (num_ranks == 1)
    {
        // the first stage's rows are folded on the side stream while the second stage runs; the second
        // stage's finish_stage also writes time and dt of the next step's stages (no separate prepare_next)
        impl->pending_prepare.enabled = 1;
        impl->pending_prepare.cfg = cfg;
        impl->pending_prepare.current_a = impl->d_stage + a;
        impl->pending_prepare.next_a = impl->d_stage + na;
        impl->pending_prepare.next_b = impl->d_stage + nb;
        launch_stage_kernels(in, nullptr, scratch, a, /*exchange*/ true, 1, 1);
        M3B_CUDA(cudaStreamWaitEvent(s, impl->positions_done[1 - parity], 0));     // positions of the second stage
        launch_stage_kernels(scratch, &in, out, b, /*exchange*/ true, 2, fixed_dt ? 0 : 2);
    }
    else if (impl->peer_transport)
    {
        // as on one rank, the first stage's rows are folded beside the second stage; the per-rank results then
        // travel through the mailboxes and prepare_next_peer folds them (no NCCL call in the step)
        auto& pp = impl->pending_prepare;
        pp.enabled = 2;
        pp.cfg = cfg;
        pp.current_a = impl->d_stage + a;
        pp.next_a = impl->d_stage + na;
        pp.next_b = impl->d_stage + nb;
        pp.peers = impl->peers;
        pp.local = impl->d_results_local;
        pp.host_results = impl->d_results;
        pp.me = rank_; pp.nranks = num_ranks; pp.slot_stride = num_slots; pp.slot_a = a; pp.slot_b = b;
        pp.counter = ++impl->step_counter;
        pp.clock_words = stage_timing ? impl->d_exchange_clock : nullptr;
        launch_stage_kernels(in, nullptr, scratch, a, /*exchange*/ true, 1, 1);
        M3B_CUDA(cudaStreamWaitEvent(s, impl->positions_done[1 - parity], 0));
        launch_stage_kernels(scratch, &in, out, b, /*exchange*/ true, 2, fixed_dt ? 0 : 2);
    }
    else
    {
        launch_stage_kernels(in, nullptr, scratch, a, /*exchange*/ true, 0, 1);
        launch_stage_kernels(scratch, &in, out, b, /*exchange*/ true, 0, fixed_dt ? 0 : 2);
        const size_t doubles = num_slots * sizeof(stage_result_t) / sizeof(double);
        impl->comm->all_gather(reinterpret_cast<const double*>(impl->d_results_local), reinterpret_cast<double*>(impl->d_results_all), doubles, stream_);
        prepare_next<<<1, 32, 0, s>>>(impl->d_results_all, num_ranks, num_slots, a, b, cfg, impl->d_stage, impl->d_stage + na, impl->d_stage + nb, impl->d_results);
        ++launches;
        M3B_CUDA(cudaGetLastError());
    }
    if (split_prepare)
    {
        // the slow half: positions of the next step's second stage and of the first stage of the step after it
        M3B_CUDA(cudaEventRecord(impl->fast_prepare_done, s));
        M3B_CUDA(cudaStreamWaitEvent(impl->finish_stream, impl->fast_prepare_done, 0));
        prepare_positions<<<1, 32, 0, impl->finish_stream>>>(cfg, impl->d_stage + na, impl->d_stage + nb, impl->d_stage + a);
        ++launches;
        M3B_CUDA(cudaEventRecord(impl->positions_done[parity], impl->finish_stream));
        M3B_CUDA(cudaGetLastError());
    }
    impl->mark(s, "prepare done");
    impl->end_step();
    M3B_CUDA(cudaEventRecord(impl->step_done[parity], s));
}

void device_solver_t::upload_step_inputs(int parity, const stage_inputs_t& first, const stage_inputs_t& second)
{
    M3B_CUDA(cudaSetDevice(device_id));
    upload_stage(first, first_async_slot + 2 * parity);
    upload_stage(second, first_async_slot + 2 * parity + 1);
    impl->fresh_pipeline = true;
}

void device_solver_t::wait_step(int parity)
{
    M3B_CUDA(cudaEventSynchronize(impl->step_done[parity]));
}

stage_result_t device_solver_t::async_result(int parity, int stage) const
{
    const auto& r = host_results[first_async_slot + 2 * parity + stage];     // already folded over ranks by prepare_next
    if (num_ranks > 1 && impl->peer_transport) throw_if_called_off(r.pad);
    return r;
}

int device_solver_t::async_slot(int parity, int stage) const
{
    return first_async_slot + 2 * parity + stage;
}

void device_solver_t::launch_max_timestep(const device_field_t& in, double time, const two_body_t& bodies, int slot)
{
    auto s = cudaStream_t(stream_);
    M3B_CUDA(cudaSetDevice(device_id));
    auto inputs = stage_inputs_t();
    inputs.time = time;
    inputs.dt = 0.0;
    inputs.theta = 0.0;
    inputs.bodies = bodies;
    upload_stage(inputs, slot);
    double* block_rows = impl->d_block_rows[slot & 1];
    max_timestep_kernel<<<BO, THREADS, 0, s>>>(impl->mesh, impl->model, impl->d_stage + slot, in.data, block_rows);
    ++launches;
    launch_finish(nullptr, 0, 1, block_rows, 1, BO, slot, 0);
}








/** Some rank gave up waiting for a peer (bounded_wait_sys): name both in an exception, which the C ABI turns into M3B_ERROR. */
void m3b::throw_if_called_off(unsigned int word)
{
    if (word == 0) return;
    const unsigned waiting = (word - 1) & 255u, awaited = (word - 1) >> 8;
    throw std::runtime_error("mara3_b200: rank " + std::to_string(waiting) + " waited longer than the deadline (M3B_SPIN_DEADLINE_MS) for rank "
        + std::to_string(awaited) + " to deliver its guard zones or stage results; the run was called off on every rank");
}

stage_result_t device_solver_t::stage_result(int slot) const
{
    if (num_ranks == 1) return host_results[slot];
    if (impl->peer_transport) throw_if_called_off(host_results[slot].pad);

    // fold the ranks in rank order: every rank computes the same bits
    auto r = stage_result_t();
    r.dt_min = 1e300;
    for (int p = 0; p < num_ranks; ++p)
    {
        const auto& q = impl->h_results_all[size_t(p) * num_slots + slot];
        for (int k = 0; k < 16; ++k) r.sums[k] += q.sums[k];
        r.work[0] += q.work[0];
        r.work[1] += q.work[1];
        r.dt_min = std::min(r.dt_min, q.dt_min);
        r.num_negative += q.num_negative;
    }
    return r;
}

unsigned int device_solver_t::local_num_negative(int slot) const
{
    return num_ranks == 1 ? host_results[slot].num_negative : impl->h_results_all[size_t(rank_) * num_slots + slot].num_negative;
}

void device_solver_t::sync()
{
    M3B_CUDA(cudaSetDevice(device_id));
    M3B_CUDA(cudaStreamSynchronize(cudaStream_t(stream_)));
    if (impl->finish_stream) M3B_CUDA(cudaStreamSynchronize(impl->finish_stream));
    if (impl->comm_stream) M3B_CUDA(cudaStreamSynchronize(impl->comm_stream));
}

std::vector<offender_t> device_solver_t::offenders(int slot)
{
    fail_dev_t f;
    M3B_CUDA(cudaSetDevice(device_id));
    M3B_CUDA(cudaStreamSynchronize(cudaStream_t(stream_)));
    M3B_CUDA(cudaMemcpy(&f, impl->d_fail + slot, sizeof(fail_dev_t), cudaMemcpyDeviceToHost));
    auto n = std::min<unsigned>(f.pad, max_offenders);
    auto v = std::vector<offender_t>(f.list, f.list + n);
    std::sort(v.begin(), v.end(), [] (const offender_t& a, const offender_t& b) { return a.block != b.block ? a.block < b.block : a.cell < b.cell; });
    return v;
}
