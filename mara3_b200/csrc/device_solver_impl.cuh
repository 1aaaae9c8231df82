/**
 * device_solver_impl.cuh -- the private state of device_solver_t, shared by the translation units that implement it:
 * kernels.cu (stage kernels of every tree, finish kernels, the step pipeline) and transport.cu (guard-zone and result
 * transport between GPUs: peer-memory mailboxes over NVLink, NCCL as set-up channel and fallback).
 */
#pragma once
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "comm.hpp"
#include "device_solver.hpp"
#include "kernel_common.cuh"

#ifndef M3B_CUDA
#define M3B_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #call); } while (0)
#endif

namespace m3b { namespace dev
{
    /** What prepare_next needs to set up the following step without the host (constant while the binary is not live). */
    struct step_config_t
    {
        elements_t elements;
        double cfl_number, recommended_time_step, theta;
        int fixed_dt;
    };


    /** Optional epilogue of finish_stage (single rank): what prepare_next does, in the last CTA of the step's last stage. */
    struct prepare_args_t
    {
        int enabled;
        step_config_t cfg;
        const stage_t* current_a;       // first stage of the step that is ending (its time and dt)
        stage_t* next_a;
        stage_t* next_b;
        // enabled == 2: several ranks, results exchanged through the peer mailboxes (peer_prepare)
        peer_table_t peers;
        const stage_result_t* local;
        stage_result_t* host_results;
        int me, nranks, slot_stride, slot_a, slot_b;
        unsigned long long counter;
        unsigned long long* clock_words;       // stage timing: ns waited for the other ranks' results, calls
    };

}} // namespace m3b::dev

namespace m3b
{
using namespace m3b::dev;

void throw_if_called_off(unsigned int word);

struct device_solver_t::impl_t
{
    mesh_dev_t mesh {};
    model_t model {};
    int tile_x = 0, tile_y = 0;
    int num_global_blocks = 0;
    int num_interior = 0;                   // leading entries of `regular` that touch no ghost block
    bool overlap_exchange = false;          // M3B_OVERLAP_EXCHANGE=1: exchange on its own stream beside the interior update
    cudaStream_t comm_stream = nullptr;     // guard-zone exchange runs here, beside the interior update
    cudaEvent_t input_ready = nullptr, halo_ready = nullptr;
    bool fast_eos = false;                  // default equation of state / viscosity: branch-free kernel variant
    bool strip = false;                     // stage_strip (16 x 32 tiles, warp-organised) instead of stage_fused
    bool tma = false;                       // regular blocks through stage_tma (persistent, cp.async.bulk staging); M3B_STAGE=strip: stage_strip
    bool tma_fast = false;                  // stage_tma's branch-free equation of state (fast_eos and alpha > 0)
    int tma_ctas_per_sm = 0;                // M3B_TMA_CTAS: 3 (two tile buffers, 168 registers) or 4 (one buffer, 128 registers); 0: 3
    unsigned char* d_tile_flags = nullptr;
    tile_info_t* d_tile_info = nullptr;     // [regular list position][tile]
    tile_info_t* d_jump_tile_info = nullptr;    // stage_strip<.., JUMP>: the tiles of blocks at refinement jumps that touch a jump
    bool jump_after_regular = false;            // M3B_JUMP_AFTER=1: the JUMP tiles' launch on the compute stream behind the regular kernel instead of beside it (measured: c4 equal, c4x 2 % slower)
    int num_jump_tiles = 0;                     // entries of d_jump_tile_info
    int num_extra_tiles = 0;                    // tiles of blocks at jumps that touch same-level leaves only: behind the regular blocks' in d_tile_info
    cudaStream_t jump_stream = nullptr;     // stage_strip<.., JUMP> runs here, beside the regular blocks' launch
    cudaEvent_t gradients_done = nullptr, jump_done = nullptr;   // fork (stage input ready on the compute stream) and join
    bool jump_mode0 = false;                // M3B_JUMP_MODE0=1: run-time stage flags in the JUMP variant (experiment)
    bool serial_jump = false;               // M3B_SERIAL_JUMP=1: the jump blocks' launch follows the regular blocks' on the compute stream
    bool jump_strip = false;                // blocks at jumps take stage_strip<.., JUMP> (M3B_JUMP_STRIP=0: the 16 x 16 any-tree kernels)
    std::vector<int> regular, irregular, gradient_blocks;
    int* d_regular = nullptr;
    int* d_irregular = nullptr;
    int* d_gradient_blocks = nullptr;
    double* d_gradients = nullptr;
    double* d_partials = nullptr;
    double* d_staging = nullptr;            // [B][3][NN] for layout changes
    fail_dev_t* d_fail = nullptr;           // [num_slots]
    stage_result_t* d_results = nullptr;    // [num_slots]
    std::vector<void*> owned;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing_events;
    // stage timing on several ranks: [input ready -> ghosts unpacked] on the exchange stream, and what of it the compute stream
    // sees: [interior blocks done -> blocks with ghost neighbours may start]
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> exchange_events, gap_events;
    unsigned long long* d_exchange_clock = nullptr;     // [0] ns spent in peer_prepare's wait for the other ranks, [1] calls, [2] ns CTAs of stage_tma spent in exchange_unpack, [3] calls, [4] ns from the start of a fused stage kernel to its flags raised on the neighbours, [5] pushes
    int* d_fused_counters = nullptr;                    // fused_exchange_t::counters
    unsigned long long fused_launches = 0;
    bool fused_exchange = true;                         // M3B_FUSED_EXCHANGE=0: halo_push / halo_wait_unpack kernels beside a split stage launch
    std::vector<cudaEvent_t> event_pool;
    size_t fused_smem = 0;
    int sm_count = 148;

    // stage inputs live in device memory (one per result slot): uploaded by the host, or written by prepare_next
    stage_t* d_stage = nullptr;                 // [num_slots]
    stage_t* h_stage_ring = nullptr;            // pinned staging ring for the uploads
    int ring_next = 0;
    double* d_partials2 = nullptr;              // second row buffer: the two stages of a step stay separate
    double* d_block_rows[2] = {nullptr, nullptr};   // one row per block, folded by the last tile CTA of the block
    double* d_general_tile_rows[2] = {nullptr, nullptr};    // general_update_tiled: one row per tile of a block at a refinement jump
    bool multi_cta_finish = false;              // M3B_MULTI_CTA_FINISH=1: never use finish_stage_cluster
    int gtile = 0;                              // tile of general_gradients_tiled / general_update_tiled (16, 12, 8; 0: none divides the block)
    bool untiled_general = false;               // M3B_UNTILED_GENERAL=1: the one-CTA-per-block any-tree update (reference for the tiled one)
    double* d_cta_rows = nullptr;               // finish_stage's per-CTA rows, one set per slot parity
    size_t cta_rows_stride = 0;
    cudaStream_t finish_stream = nullptr;       // finish_stage of a step's first stage runs here, beside the second stage
    cudaEvent_t stage_done = nullptr, side_finish_done = nullptr;
    prepare_args_t pending_prepare = prepare_args_t();
    cudaEvent_t fast_prepare_done = nullptr, positions_done[2] = {nullptr, nullptr};
    bool fresh_pipeline = false;                // the host uploaded this step's inputs: nobody has prepared the next step's positions
    int* d_counters = nullptr;                  // per-block tile tickets, then the finish ticket
    size_t partial_rows = 0;
    cudaEvent_t step_done[2] = {nullptr, nullptr};

    // multi-GPU: guard-zone exchange plan and cross-rank reduction of the stage results
    communicator_t* comm = nullptr;
    int num_send_entries = 0, num_recv_entries = 0;
    halo_entry_dev_t* d_send_entries = nullptr;
    halo_entry_dev_t* d_recv_entries = nullptr;
    double* d_send_buffer = nullptr;
    double* d_recv_buffer = nullptr;
    // M3B_TRACE=1: CUDA events at the marks of launch_step_async, averaged and printed by the destructor
    bool trace = false;
    std::vector<std::vector<cudaEvent_t>> trace_steps;
    std::vector<cudaEvent_t> trace_current;
    std::vector<const char*> trace_names;
    void mark(cudaStream_t s, const char* name)
    {
        if (! trace || trace_steps.size() >= 400) return;
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s);
        if (trace_steps.empty()) trace_names.push_back(name);
        trace_current.push_back(e);
    }
    void end_step() { if (trace && ! trace_current.empty()) { trace_steps.push_back(trace_current); trace_current.clear(); } }
    // peer-memory transport (set up in set_communicator; falls back to NCCL send / recv if CUDA IPC is unavailable)
    bool peer_transport = false;
    void* mailbox = nullptr;                        // this rank's mailbox: flags, results of all ranks, two landing buffers
    std::vector<void*> peer_mailbox;                // cudaIpcOpenMemHandle mappings (null for this rank)
    peer_table_t peers = peer_table_t();
    halo_entry_dev_t* d_push_entries = nullptr;     // send entries addressed into the destination's landing buffer
    std::vector<halo_entry_dev_t> send_entries_host;
    std::vector<size_t> send_starts_host, recv_starts_host;
    size_t recv_total = 0;
    unsigned int dest_mask = 0;
    unsigned long long exchange_counter = 0, step_counter = 0;
    int* d_push_ticket = nullptr;                   // [0] halo_push, [1] halo_wait_unpack
    unsigned long long* d_ready = nullptr;          // number of the last exchange whose ghost blocks are complete
    bool defer_unpack = false;                      // set around the exchange of a stage launched with in-kernel waiting
    bool in_kernel_wait = false;                    // boundary tiles wait inside the stage kernel (enough interior work to hide the exchange)
    std::vector<const double*> send_ptr;
    std::vector<double*> recv_ptr;
    std::vector<size_t> send_count, recv_count;
    stage_result_t* d_results_local = nullptr;      // [num_slots] device copy that NCCL can read
    stage_result_t* d_results_all = nullptr;        // [nranks][num_slots]
    stage_result_t* h_results_all = nullptr;        // pinned
    std::uint64_t halo_bytes_per_exchange = 0;
};


} // namespace m3b
