/**
 * scheme.hpp -- host orchestration of the `binary` time step on one GPU.
 *
 * Mirrors the operator API of the reference for this path (Mara3
 * src/subprog_binary.hpp:180-208): create_solver_data + set_scheme_globals ->
 * binary_solver_t, create_solution, maximum_timestep, advance, and the step rule
 * binary::next_solution (src/subprog_binary.cpp:258-293).  The heavy lifting is
 * done by device_solver_t (kernels.cu); everything here is O(1) per stage.
 */
#pragma once
#include <memory>
#include <string>
#include <vector>
#include "comm.hpp"
#include "config.hpp"
#include "device_solver.hpp"
#include "solver_data.hpp"
#include "two_body.hpp"

namespace m3b
{
    /** solution_t (subprog_binary.hpp:108-126): conserved field on the device + scalars on the host. */
    struct solution_t
    {
        double time = 0.0;
        int iteration_num = 0, iteration_den = 1;      // rational iteration counter
        std::shared_ptr<device_field_t> conserved_u;
        double mass_accreted_on[2] = {0, 0};
        double angular_momentum_accreted_on[2] = {0, 0};
        double integrated_torque_on[2] = {0, 0};
        double work_done_on[2] = {0, 0};
        double mass_ejected = 0.0;
        double angular_momentum_ejected = 0.0;
        elements_t orbital_elements_acc = elements_zero();
        elements_t orbital_elements_grav = elements_zero();
        elements_t orbital_elements;

        static constexpr int num_scalars = 43;
        void get_scalars(double out[num_scalars]) const;
        void set_scalars(const double in[num_scalars]);
    };

    enum status_t : int
    {
        status_ok = 0,
        status_negative_density = 1,    // validate_u would throw (scheme.cpp:747-750)
        status_unbound_orbit = 2,       // compute_orbital_elements would throw (model_two_body.hpp:385-386)
        status_unsupported = 3,
        status_error = -1,
    };

    class binary_solver_t
    {
    public:
        /** rank / nranks / nccl_unique_id: one process per GPU; the leaf blocks are cut into Morton-contiguous
         *  ranges (partition.hpp).  nccl_unique_id may be null for host-only use (device < 0). */
        binary_solver_t(const config_t& run_config, int device, bool general_only = false, bool tiled_kernel = false,
                        int rank = 0, int nranks = 1, const unsigned char* nccl_unique_id = nullptr);

        const config_t& run_config() const { return config; }
        const solver_data_t& solver_data() const { return data; }
        device_solver_t& device();
        bool has_device() const { return bool(gpu); }

        /** create_solution (subprog_binary.cpp:196-227): the initial disk. */
        solution_t create_solution();
        solution_t clone(const solution_t& s);
        std::shared_ptr<device_field_t> new_field();

        /** binary::maximum_timestep (scheme.cpp:1107-1126). */
        double maximum_timestep(const solution_t& s);

        /** binary::advance (scheme.cpp:1022-1027): out = one RK stage applied to in. */
        status_t advance(const solution_t& in, double dt, bool safe_mode, solution_t& out);

        /** s0 * b0 + s2 * (1 - b0) over every field (scheme.cpp:1033-1069). */
        void combine(const solution_t& s0, const solution_t& s2, double b0, solution_t& out);

        /** binary::next_solution (subprog_binary.cpp:258-293), in place. */
        status_t next_solution(solution_t& s, double* dt_used, bool* fell_back, bool speculate = true);
        /** Forget a step queued ahead of the caller (after the caller changed a state in place). */
        void invalidate() { drop_speculation(); dt_cache_field = nullptr; }

        /** Messages the reference would have printed for the last negative-density failure. */
        const std::vector<std::string>& last_messages() const { return messages; }
        const std::string& last_error() const { return error; }
        void set_quiet(bool q) { quiet = q; }
        /** Queue the following step before waiting for the current one (default on). */
        void set_pipelining(bool on) { drop_speculation(); pipelining = on; }

    private:
        status_t try_step(solution_t& s, double dt, bool safe_mode);
        status_t bookkeeping(const solution_t& in, const stage_result_t& r, const two_body_t& bodies, double dt, solution_t& out);
        void record_offenders(int slot);
        stage_inputs_t stage_inputs(const solution_t& in, double dt, bool safe_mode) const;

        config_t config;
        solver_data_t data;
        std::unique_ptr<device_solver_t> gpu;
        std::unique_ptr<communicator_t> comm;
        std::shared_ptr<device_field_t> scratch1, scratch2;
        struct field_pool_t { std::vector<std::unique_ptr<device_field_t>> free; };
        std::shared_ptr<field_pool_t> pool = std::make_shared<field_pool_t>();     // recycled state buffers (cudaMalloc is slow)
        std::vector<std::string> messages;
        std::string error;
        bool quiet = false;

        /**
         * A step that has been queued on the GPU for a state the caller has not asked to advance yet:
         * next_solution() queues step n + 1 (from the output of step n, with dt and body positions
         * computed on the device) before it waits for step n, so the GPU never idles on the host.
         */
        struct speculation_t
        {
            bool valid = false;
            std::shared_ptr<device_field_t> in, out;    // the state it starts from / produces
            double time = 0.0;                          // ... and that state's time
            double dt = 0.0;                            // known once the step before has been consumed
            elements_t elements;                        // orbital elements its body positions were computed from
            int parity = 0;
        };
        speculation_t speculation;
        bool pipelining = true;
        bool can_pipeline(const solution_t& s, double dt) const;
        void launch_pipelined(const std::shared_ptr<device_field_t>& in, const std::shared_ptr<device_field_t>& out, int parity, const elements_t& elements);
        status_t finish_pipelined(solution_t& s, const speculation_t& step);
        void drop_speculation();

        // CFL estimate produced by the last fused step, valid for exactly one state
        const device_field_t* dt_cache_field = nullptr;
        double dt_cache_time = 0.0;
        double dt_cache_value = 0.0;
    };
}
