/**
 * binary_io.hpp -- run state of the `binary` subprogram and its HDF5 products.
 *
 * Counterpart of the reference's state_t / schedule_t / time_series_sample_t / diagnostic_fields_t
 * (Mara3 src/subprog_binary.hpp:108-175, src/app_schedule.hpp:57-196) and of their serialisers
 * (src/subprog_binary_io.cpp:9-215, src/app_serialize.hpp:69-188, src/app_serialize_tree.hpp:72-180):
 * same group / dataset names, shapes and datatypes, written through h5lite.
 */
#pragma once
#include <map>
#include <string>
#include <vector>
#include "scheme.hpp"

namespace m3b
{
    /** mara::schedule_t (app_schedule.hpp:57-147). */
    struct schedule_t
    {
        struct task_t { std::string name; int num_times_performed = 0; double last_performed = 0.0; bool is_due = false; };
        std::map<std::string, task_t> tasks;

        void create_and_mark_as_due(const std::string& name) { tasks[name] = {name, 0, 0.0, true}; }
        task_t& at(const std::string& name);
        const task_t& at(const std::string& name) const;
        void mark_as_due(const std::string& name, double increase_last_performed_by = 0.0) { auto& t = at(name); t.is_due = true; t.last_performed += increase_last_performed_by; }
        void mark_as_completed(const std::string& name) { auto& t = at(name); t.is_due = false; t.num_times_performed += 1; }
    };

    /** binary::time_series_sample_t with the reference's memory layout (subprog_binary.hpp:144-161): the HDF5
     *  compound type records these offsets (subprog_binary_io.cpp:96-125). */
    struct time_series_sample_t
    {
        double time = 0.0;
        double disk_mass = 0.0;
        double disk_angular_momentum = 0.0;
        double mass_ejected = 0.0;
        double angular_momentum_ejected = 0.0;
        double mass_accreted_on[2] = {0, 0};
        double angular_momentum_accreted_on[2] = {0, 0};
        double integrated_torque_on[2] = {0, 0};
        double work_done_on[2] = {0, 0};
        double orbital_elements_acc[10] = {0};      // full_orbital_elements_t: pomega, tau, cm x/y, cm vx/vy, {separation, total_mass, mass_ratio, eccentricity}
        double orbital_elements_grav[10] = {0};
        double orbital_elements[10] = {0};
        double position_of_mass1[2] = {0, 0};
        double position_of_mass2[2] = {0, 0};
    };
    static_assert(sizeof(time_series_sample_t) == 376, "time_series_sample_t must match the reference's layout");

    /** binary::state_t (subprog_binary.hpp:165-175); the time series is kept oldest first (the order it is written in). */
    struct state_t
    {
        solution_t solution;
        schedule_t schedule;
        std::vector<time_series_sample_t> time_series;
    };

    /** format_tree_index (app_serialize_tree.hpp:72-87): "level:ii-jj", zero padded to the width of 2^level. */
    std::string format_tree_index(int level, int i, int j);

    /** record_time_series' sample (subprog_binary.cpp:358-379). */
    time_series_sample_t make_time_series_sample(binary_solver_t& solver, const solution_t& solution);

    /** mara::write<state_t> into chkpt.NNNN.h5 (subprog_binary_io.cpp:131-158). */
    void write_checkpoint(const std::string& filename, binary_solver_t& solver, const state_t& state);

    /** mara::write<diagnostic_fields_t> into diagnostics.NNNN.h5 (subprog_binary_io.cpp:160-172). */
    void write_diagnostics(const std::string& filename, binary_solver_t& solver, const solution_t& solution);

    /** run_config stored in a checkpoint, as key -> text (read_config, app_serialize.hpp:96-113). */
    std::map<std::string, std::string> read_checkpoint_config(const std::string& filename);

    /** mara::read<state_t> (subprog_binary_io.cpp:174-200): solution, time series and schedule; is_due flags are not stored. */
    state_t read_checkpoint(const std::string& filename, binary_solver_t& solver);

    /**
     * The subprogram: `binary key=value ...` (subprog_binary.cpp:414-436) -- config, initial or restarted state,
     * run loop with the scheduled tasks, the same stdout lines and output files.  Returns the exit code.
     * With nranks > 1 (one process per GPU) every rank runs the same loop; rank 0 prints and writes the gathered products.
     */
    int binary_main(int argc, const char* const argv[], int device, int rank = 0, int nranks = 1, const unsigned char* nccl_unique_id = nullptr);
}
