/**
 * kernel_common.cuh -- device-side records shared by the translation units that hold kernels (kernels.cu, stage_tma.cu):
 * the static mesh tables, the per-tile record of the strip kernels, the negative-density report.
 */
#pragma once
#include <cuda_runtime.h>
#include "device_solver.hpp"
#include "iso2d_device.cuh"

namespace m3b { namespace dev
{
    constexpr int THREADS = 256;
    constexpr int ROW = 20;             // doubles per partial row: 16 sums, dt_min, pad
    constexpr int FINISH_THREADS = 256;
    constexpr int FINISH_ROWS_PER_CTA = 32;     // block rows folded by one finish_stage CTA
    constexpr int stage_ring_size = 64;

    struct __align__(16) face_nbr_dev_t
    {
        int kind;                       // 0 same, 1 coarser, 2 finer
        int leaf[4];
        int bx, by;
        int pad;
        int gs[4];                      // gslot of leaf[q]: where its gradients live in the scratch (stage_strip<.., JUMP> reads the record as three int4)
    };
    static_assert(sizeof(face_nbr_dev_t) == 48, "stage_strip reads face_nbr_dev_t as three int4");

    struct mesh_dev_t
    {
        int B, N;
        size_t FS;                      // field stride = B * N * N
        const double* xv;               // [B][N+1]
        const double* yv;               // [B][N+1]
        const double* spacing;          // [B]
        const double* inv_spacing;      // [B] 1 / spacing
        const face_nbr_dev_t* nbr;      // [B][4]
        const int* nbr9;                // [B][9] same-level neighbour leaf ids (regular blocks only)
        const int* gslot;               // [B] slot of the block in the gradient scratch, or -1
        const double* U0;               // [3][FS]
        const double* br;               // [FS]
        size_t GS;                      // gradient scratch stride
        int qmode;                      // the state is conserved_q = (sigma, Sr, Lz): primitives need the cell position
        int prefetch_ahead;             // stage_strip: CTAs resident at once (L2 prefetch distance), 0 = off
        // multi-GPU: CTAs from first_wait_cta on update blocks with ghost neighbours and wait until the guard-zone
        // unpack (running beside this kernel on the exchange stream) has published ready_value
        int first_wait_cta;
        const unsigned long long* ready_flag;
        unsigned long long ready_value;
    };

    /** Everything a stage_strip CTA needs to find its data, in one 48-byte record per tile (one load instead of
     *  the dependent chain regular list -> neighbour table / tile flags). */
    struct __align__(16) tile_info_t
    {
        int b;                          // block
        int n9[9];                      // same-level neighbour ids, (di + 1) * 3 + (dj + 1)
        // bit 0: the buffer-zone rate is non-zero somewhere in the tile; bits 1-4 (stage_strip<.., JUMP>): the block side at the
        // tile's low-x / high-x / low-y / high-y edge has a finer neighbour; bit 5 (stage_tma): the tile belongs to a block at a
        // refinement jump but touches same-level leaves only, and its partial row goes to the any-tree rows;
        // bits 8 and up: position of the tile in its block (row-major over the block's tiles)
        int flags;
        int row;                        // the tile's partial row (index into the launch's rows array)
    };
    constexpr int TILE_POS_SHIFT = 8, TILE_JUMP_ROWS = 32;

    struct fail_dev_t
    {
        unsigned int count;
        unsigned int pad;
        m3b::offender_t list[m3b::device_solver_t::max_offenders];
    };

    __device__ __forceinline__ void report_negative(fail_dev_t* fail, int block, int cell, double sigma)
    {
        unsigned int n = atomicAdd(&fail->count, 1u);
        if (n < m3b::device_solver_t::max_offenders) fail->list[n] = {block, cell, sigma};
    }

    // geometry of the strip kernels (stage_strip, stage_tma): CTA = 4 warps on a 16 x 32 tile, warp <-> strip of 4 rows
    constexpr int SX = 16, SY = 32, STRIP = 4, STRIP_THREADS = 128;

    __device__ __forceinline__ double shfl_down1(double v)
    {
        return __shfl_down_sync(0xffffffffu, v, 1);
    }

    // ---- multi-GPU transport records (peer memory over NVLink: kernels.cu sets the mailboxes up, see set_communicator)
    constexpr int MAX_PEERS = 16;

    /** Mapped (CUDA IPC) pointers into every rank's mailbox; index = rank.  [me] points at the local mailbox. */
    struct peer_table_t
    {
        double* recv[MAX_PEERS][2];                     // guard-zone landing buffers, one per exchange parity
        unsigned long long* halo_flag[MAX_PEERS];       // [src rank]: number of the last exchange `src` has delivered
        stage_result_t* results[MAX_PEERS];             // [src rank][num_slots]
        unsigned long long* result_flag[MAX_PEERS];     // [src rank]: number of the last step whose results `src` has delivered
        // every wait on another rank is bounded: a rank that has waited `deadline_cycles` for a peer writes
        // 1 + (waiting rank) + 256 (awaited rank) into the abort word of EVERY mailbox, all spins end, and the host turns the
        // word into M3B_ERROR naming the stalled peer (device_solver_t::stage_result)
        unsigned long long* abort_word[MAX_PEERS];
        long long deadline_cycles;
    };

    __device__ __forceinline__ void store_release_sys(unsigned long long* p, unsigned long long v)
    {
        asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
    }

    __device__ __forceinline__ unsigned long long load_acquire_sys(const unsigned long long* p)
    {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        return v;
    }



    /** Spin until *flag >= want, the deadline passes, or any rank has called the run off.  false: give up (the abort word is set). */
    __device__ __forceinline__ bool bounded_wait_sys(const unsigned long long* flag, unsigned long long want, const peer_table_t& peers, int me, int awaited)
    {
        if (load_acquire_sys(flag) >= want) return true;
        const long long t0 = clock64();
        for (unsigned n = 1; ; ++n)
        {
            if (load_acquire_sys(flag) >= want) return true;
            if ((n & 255u) == 0)
            {
                if (load_acquire_sys(peers.abort_word[me]) != 0) return false;
                if (clock64() - t0 > peers.deadline_cycles)
                {
                    const unsigned long long code = 1ull + unsigned(me) + 256ull * unsigned(awaited);
                    for (int p = 0; p < MAX_PEERS; ++p) if (peers.abort_word[p]) store_release_sys(peers.abort_word[p], code);
                    return false;
                }
            }
        }
    }

    struct halo_entry_dev_t
    {
        int block;              // local block id
        int i0, ni, j0, nj;     // cells [i0, i0 + ni) x [j0, j0 + nj)
        int pad;
        size_t offset;          // first double of this entry in the packed buffer (3 fields x ni x nj)
    };


    /**
     * The guard-zone exchange of one stage, done by the persistent stage kernel itself (stage_tma.cuh): at its start every CTA
     * stores its share of this rank's strips into the neighbours' landing buffers over NVLink and the last one raises this rank's
     * flag there; the interior blocks are then updated; before the first block with ghost neighbours is fetched, the CTAs that
     * get there first wait for the neighbours' flags and scatter the landed strips into the ghost blocks.  enabled = 0: no
     * exchange in this launch.  counters: [parity][0] strips pushed, [1] next strip to unpack, [2] strips unpacked, [3] next strip to push; a launch
     * zeroes the other set, which the previous fused launch used.
     */
    struct fused_exchange_t
    {
        int enabled;
        const halo_entry_dev_t* push; int n_push;
        const halo_entry_dev_t* recv; int n_recv;
        peer_table_t peers;
        int parity, me;                 // landing buffer of this exchange (exchange number & 1), this rank
        int cset;                       // counter set of this launch (fused launch number & 1)
        unsigned int dest_mask;
        unsigned long long counter;
        int* counters;                  // [2][4]
        double* U;                      // the stage input, writable: the ghost blocks are filled in
        unsigned long long* clock_words;    // stage timing: [2] ns CTAs spent waiting for neighbours' flags or for the unpack, [3] waits
    };

    /** stage_tma.cu: one launch of the persistent TMA-staged stage kernel over `num_tiles` entries of `tile_info`. */
    struct stage_tma_launch_t
    {
        mesh_dev_t mesh;
        model_t model;
        const stage_t* stage;
        const tile_info_t* tile_info;
        int num_tiles;
        const double* Uin;
        const double* Un;
        double* Uout;
        double* partials;
        double* jump_partials;      // rows of the tiles flagged TILE_JUMP_ROWS
        fail_dev_t* fail;
        int N;              // block size
        bool fast;          // branch-free equation of state
        int stage_mode;     // 0: flags from stage_t, 1 / 2: first / last stage of an adaptive RK2 step
        int grid;           // CTAs asked for (clamped to what is resident at once)
        int ctas_per_sm;    // 3: two tile buffers, 168 registers; 4: one buffer, 128 registers
        fused_exchange_t exchange;
    };
    void stage_tma_configure();
    void stage_tma_launch(const stage_tma_launch_t& a, cudaStream_t stream);
    size_t stage_tma_shared_bytes();
}} // namespace m3b::dev
