/**
 * device_solver.hpp -- device-resident state and kernel launches for the iso2d update.
 *
 * Host-visible interface of the CUDA side (kernels.cu).  One device_solver_t owns
 * the static mesh data on one GPU (block geometry, neighbour tables, buffer-zone
 * rate, initial state) and the scratch needed by a stage; device_field_t is a
 * conserved state U in the device layout [field][block][i][j] (structure of
 * arrays, y fastest -- the reference's row-major (N,N) block per tuple component).
 */
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>
#include "solver_data.hpp"

namespace m3b
{
    /** Reduced outputs of one stage launch (host copy). */
    struct stage_result_t
    {
        double sums[16];        // dev::ACC_* / GRV_* / BUF_* running sums, already multiplied by cell area (not by dt)
        double work[2];         // sum over blocks of the per-block work integral on each body (scheme.cpp:407-408)
        double dt_min;          // min over cells of spacing / max wavespeed of the output state (if requested)
        unsigned int num_negative;   // cells with sigma < 0 in the un-combined update (validate_u, scheme.cpp:726-752)
        unsigned int pad;
    };

    struct offender_t { int block, cell; double sigma; };     // block: local id on the reporting rank
    class communicator_t;

    struct stage_inputs_t
    {
        double time, dt, theta;
        two_body_t bodies;
        double rk_b0 = 0.0;
        bool combine = false;       // out = Un * b0 + updated * (1 - b0)
        bool compute_dt = false;
    };

    class device_field_t
    {
    public:
        device_field_t(std::size_t num_doubles, int device);
        ~device_field_t();
        device_field_t(const device_field_t&) = delete;
        device_field_t& operator=(const device_field_t&) = delete;
        double* data = nullptr;
        std::size_t count = 0;
        int device = 0;
    };

    class device_solver_t
    {
    public:
        /** general_only: route every block through the any-tree kernels; tiled_kernel: use stage_fused
         *  even where stage_strip applies (both used by tests to cross-check the kernels). */
        device_solver_t(const solver_data_t& solver_data, int device, bool general_only = false, bool tiled_kernel = false);
        ~device_solver_t();

        int device() const { return device_id; }
        std::size_t field_stride() const { return cells; }      // doubles per conserved component
        std::size_t state_doubles() const { return 3 * cells; }
        void* stream() const { return stream_; }
        /** Launch on a caller-owned CUDA stream (e.g. torch's current stream) from now on. */
        void set_stream(void* cuda_stream);

        /** Copy a state between the host layout [B][3][N][N] and the device layout [3][B][N][N]. */
        void upload(const double* host_block_major, device_field_t& dst);
        void download(const device_field_t& src, double* host_block_major);
        void copy(const device_field_t& src, device_field_t& dst);
        /** Collective: every rank's owned blocks on rank 0, in global block order (one rank: a download). */
        void gather_blocks(const double* d_local, std::size_t doubles_per_block, double* host_all);
        void gather_state(const device_field_t& src, double* host_all);              // [global block][3][N][N] on rank 0
        void gather_diagnostic_fields(const device_field_t& src, double* host_all);  // sigma, v_r, v_phi, same layout
        int rank() const { return rank_; }
        int num_ranks_() const { return num_ranks; }
        int ranks() const { return num_ranks; }
        /** disk_mass, disk_angular_momentum (subprog_binary_diagnostics.cpp:19-41); summed over the ranks (collective). */
        void disk_totals(const device_field_t& src, double out[2]);
        /** sigma, radial velocity, azimuthal velocity, [owned block][3][N][N] on the host (subprog_binary_diagnostics.cpp:48-82). */
        void diagnostic_fields(const device_field_t& src, double* host);
        void load_initial(device_field_t& dst);
        /** dst = a * wa + b * wb (solution_t::operator+ / *, scheme.cpp:1033-1069) */
        void combine(const device_field_t& a, double wa, const device_field_t& b, double wb, device_field_t& dst);

        /**
         * One RK stage, binary::advance_u phases P1-P8 + P11 (scheme.cpp:790-904):
         * out = update(in) [optionally RK-combined with un].  On several ranks the guard zones of `in`
         * are exchanged first, overlapped with the update of the interior blocks.  Asynchronous; results
         * land in the slot returned by stage_result() after sync().
         */
        void launch_stage(const device_field_t& in, const device_field_t* un, device_field_t& out, const stage_inputs_t& inputs, int slot);

        /**
         * A whole RK2 step queued without host involvement: exchange, stage 1 (in -> scratch), exchange,
         * stage 2 (scratch, in -> out, RK combination and CFL estimate fused), result folding, and the
         * stage inputs of the FOLLOWING step computed on the device (prepare_next).  The inputs of THIS
         * step must already be in the device slots of `parity` (upload_step_inputs, or the previous
         * step's prepare_next).  Only valid while the binary's orbital elements are constant.
         */
        void launch_step_async(device_field_t& in, device_field_t& scratch, device_field_t& out, int parity,
                               const elements_t& elements, double cfl_number, double recommended_time_step, double theta, bool fixed_dt);
        void upload_step_inputs(int parity, const stage_inputs_t& first, const stage_inputs_t& second);
        void wait_step(int parity);
        stage_result_t async_result(int parity, int stage) const;
        int async_slot(int parity, int stage) const;

        /** binary::maximum_timestep (scheme.cpp:1107-1126) of a state; result in slot.dt_min. */
        void launch_max_timestep(const device_field_t& in, double time, const two_body_t& bodies, int slot);

        // ---- multi-GPU (no-ops / identities on one rank)
        void set_communicator(communicator_t* comm);
        /** Fill the ghost blocks of `field` with the neighbours' edge strips (NCCL send / recv). */
        void exchange_halos(device_field_t& field);
        /** All-gather every slot's stage result; after sync(), stage_result() folds the ranks. */
        void gather_results();
        /** One double from every rank, in rank order (blocking). */
        std::vector<double> all_gather_scalar(double value);
        std::uint64_t halo_bytes_per_exchange() const;
        int exchange_transport() const;     // 0 none, 1 NCCL send / recv, 2 peer memory (CUDA IPC)
        unsigned int local_num_negative(int slot) const;

        void sync();
        /** Result of the stage launched with this slot, summed / minimised over all ranks. */
        stage_result_t stage_result(int slot) const;
        std::vector<offender_t> offenders(int slot);

        /** Number of kernel launches issued so far (bench.py reports it as gpu_launches). */
        std::uint64_t launch_count() const { return launches; }

        int num_regular_blocks() const { return num_regular; }

        /** Per-kernel CUDA-event timing of the fused stage kernel (bench.py roofline). */
        void set_stage_timing(bool on);
        double stage_kernel_ms_total() const { return stage_ms_total; }
        std::uint64_t stage_kernel_launches() const { return stage_timed_launches; }
        void collect_stage_timing();
        /** Multi-GPU side of the stage timing (bench.py `exchange`): microseconds between "stage input ready" and "ghost blocks
         *  unpacked" on the exchange stream (push over NVLink + wait for the neighbours' pushes + unpack); what the compute stream
         *  saw of it (interior blocks done -> blocks with ghost neighbours may start); and the time the last finish_stage CTA of a
         *  step waited for the other ranks' stage results. */
        double exchange_us_total = 0.0, exposed_wait_us_total = 0.0, result_wait_us_total = 0.0;
        std::uint64_t exchanges_timed = 0, result_waits_timed = 0;
        double unpack_cta_wait_us_total = 0.0;      // fused exchange: time CTAs of the stage kernel spent waiting for flags / unpacking
        std::uint64_t unpack_cta_waits = 0;

        static constexpr int num_slots = 8;
        static constexpr int first_async_slot = 2;      // slots 2..5: two steps in flight x two stages
        static constexpr int max_offenders = 64;

    private:
        void upload_stage(const stage_inputs_t& inputs, int slot);
        void launch_stage_kernels(const device_field_t& in, const device_field_t* un, device_field_t& out, int slot, bool exchange, int finish_mode = 0, int stage_mode = 0);
        void exchange_on(void* cuda_stream, device_field_t& field);
        stage_result_t* result_target(int slot);
        void launch_finish(const double* tile_rows, int num_fused, int tpb, const double* general_rows, int gtpb, int num_rows, int slot, int finish_mode);
        struct impl_t;
        std::unique_ptr<impl_t> impl;
        int device_id = 0;
        int N = 0, B = 0, BO = 0;       // block size, blocks stored here, blocks owned (updated) here
        int num_ranks = 1, rank_ = 0;
        std::size_t cells = 0;
        void* stream_ = nullptr;
        void* own_stream = nullptr;
        stage_result_t* host_results = nullptr;     // pinned
        std::uint64_t launches = 0;
        bool force_general = false;
        int num_regular = 0;
        bool stage_timing = false;
        double stage_ms_total = 0.0;
        std::uint64_t stage_timed_launches = 0;
    };
}
