/**
 * capi.cpp -- the extern "C" boundary declared in include/mara3_b200.h.
 * Thin: argument checks, exception -> status code translation, layout copies.
 */
#include <cstring>
#include <sstream>
#include <string>
#include <cuda_runtime.h>
#include "../../include/mara3_b200.h"
#include "scheme.hpp"
#include "h5lite.hpp"
#include "binary_io.hpp"
#include <chrono>
#include <iostream>

using namespace m3b;

struct m3b_solver
{
    std::unique_ptr<binary_solver_t> solver;
    std::string error;
};

struct m3b_solution
{
    solution_t solution;
};

namespace
{
    thread_local std::string global_error;

    template<typename Function>
    int guarded(m3b_solver* s, Function&& fn)
    {
        try {
            return fn();
        }
        catch (const std::exception& e)
        {
            if (s) s->error = e.what();
            return M3B_ERROR;
        }
    }

    template<typename T>
    void copy_out(const std::vector<T>& v, T* out)
    {
        std::memcpy(out, v.data(), v.size() * sizeof(T));
    }
}

extern "C" {

const char* m3b_version(void) { return "mara3_b200 0.1 (sm_100a, fp64)"; }

int m3b_device_count(void)
{
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

const char* m3b_global_error(void) { return global_error.c_str(); }

m3b_solver_t* m3b_solver_create(int argc, const char* const* argv, int device, int flags)
{
    return m3b_solver_create_distributed(argc, argv, device, flags, 0, 1, nullptr);
}

int m3b_nccl_unique_id(unsigned char* out128)
{
    try { communicator_t::make_unique_id(out128); return M3B_OK; }
    catch (const std::exception& e) { global_error = e.what(); return M3B_ERROR; }
}

m3b_solver_t* m3b_solver_create_distributed(int argc, const char* const* argv, int device, int flags, int rank, int nranks, const unsigned char* nccl_unique_id)
{
    try {
        auto s = std::make_unique<m3b_solver>();
        auto config = config_t::from_argv(argc, argv);
        if (! config.get_string("restart").empty())
            throw std::invalid_argument("restart= is handled by the checkpoint reader, not by m3b_solver_create");
        s->solver = std::make_unique<binary_solver_t>(config, (flags & 2) ? -1 : device, (flags & 1) != 0, (flags & 4) != 0, rank, nranks, nccl_unique_id);
        return s.release();
    }
    catch (const std::exception& e)
    {
        global_error = e.what();
        return nullptr;
    }
}

void m3b_solver_destroy(m3b_solver_t* s) { delete s; }

const char* m3b_last_error(const m3b_solver_t* s)
{
    if (! s) return global_error.c_str();
    return s->error.empty() ? s->solver->last_error().c_str() : s->error.c_str();
}

int m3b_num_blocks(const m3b_solver_t* s) { return s->solver->solver_data().num_owned; }
int m3b_num_global_blocks(const m3b_solver_t* s) { return s->solver->solver_data().num_blocks; }
int m3b_first_block(const m3b_solver_t* s) { return s->solver->solver_data().partition.first_owned; }
int m3b_num_local_blocks(const m3b_solver_t* s) { return s->solver->solver_data().num_local; }
int64_t m3b_num_owned_cells(const m3b_solver_t* s) { return int64_t(s->solver->solver_data().num_owned_cells()); }

void m3b_local_to_global(const m3b_solver_t* s, int* out)
{
    const auto& v = s->solver->solver_data().partition.local_to_global;
    std::memcpy(out, v.data(), v.size() * sizeof(int));
}

int m3b_halo_plan_size(const m3b_solver_t* s, int peer, int send)
{
    const auto& p = s->solver->solver_data().partition;
    if (peer < 0 || peer >= p.nranks) return 0;
    return int((send ? p.send : p.recv)[peer].size());
}

void m3b_halo_plan(const m3b_solver_t* s, int peer, int send, int* out)
{
    const auto& p = s->solver->solver_data().partition;
    const auto& list = (send ? p.send : p.recv)[peer];
    for (std::size_t k = 0; k < list.size(); ++k)
    {
        out[3 * k + 0] = list[k].block;
        out[3 * k + 1] = list[k].di;
        out[3 * k + 2] = list[k].dj;
    }
}

void m3b_neighbor_table(const m3b_solver_t* s, int* out)
{
    const auto& d = s->solver->solver_data();
    for (int b = 0; b < d.num_owned; ++b)
        for (int di = -1; di <= 1; ++di)
            for (int dj = -1; dj <= 1; ++dj)
            {
                int g = d.tree->same_level_neighbor(d.global_block(b), di, dj);
                out[b * 9 + (di + 1) * 3 + (dj + 1)] = g < 0 ? -1 : d.partition.global_to_local[g];
            }
}

void m3b_face_neighbor_table(const m3b_solver_t* s, int* out)
{
    const auto& d = s->solver->solver_data();
    for (int b = 0; b < d.num_local; ++b)
        for (int side = 0; side < 4; ++side)
        {
            auto fn = d.tree->face_neighbor(d.global_block(b), side);
            int* o = out + (b * 4 + side) * 5;
            o[0] = int(fn.kind);
            for (int q = 0; q < 4; ++q) o[1 + q] = fn.leaf[q] < 0 ? -1 : d.partition.global_to_local[fn.leaf[q]];
        }
}

void m3b_format_tree_index(int level, int i, int j, char* out, int out_len)
{
    std::snprintf(out, out_len, "%s", m3b::format_tree_index(level, i, j).c_str());
}

int m3b_write_checkpoint(m3b_solver_t* s, const m3b_solution_t* u, const char* filename)
{
    return guarded(s, [&]
    {
        // a state with the initial schedule (all three tasks performed zero times, none due) and no time series:
        // what the run loop would store before its first task
        auto state = m3b::state_t();
        state.solution = u->solution;
        for (const char* task : {"write_checkpoint", "write_diagnostics", "record_time_series"}) state.schedule.tasks[task] = {task, 0, 0.0, false};
        m3b::write_checkpoint(filename, *s->solver, state);
        return M3B_OK;
    });
}

int m3b_write_diagnostics(m3b_solver_t* s, const m3b_solution_t* u, const char* filename)
{
    return guarded(s, [&] { m3b::write_diagnostics(filename, *s->solver, u->solution); return M3B_OK; });
}

int m3b_read_checkpoint(m3b_solver_t* s, m3b_solution_t* u, const char* filename)
{
    return guarded(s, [&] { s->solver->invalidate(); u->solution = m3b::read_checkpoint(filename, *s->solver).solution; return M3B_OK; });
}

int m3b_time_series_sample(m3b_solver_t* s, const m3b_solution_t* u, double* out47)
{
    return guarded(s, [&]
    {
        auto sample = m3b::make_time_series_sample(*s->solver, u->solution);
        std::memcpy(out47, &sample, sizeof(sample));
        return M3B_OK;
    });
}

int m3b_binary_main_distributed(int argc, const char* const* argv, int device, int rank, int nranks, const unsigned char* nccl_unique_id)
{
    try
    {
        auto t0 = std::chrono::high_resolution_clock::now();
        int code = m3b::binary_main(argc, argv, device, rank, nranks, nccl_unique_id);
        double s = 1e-9 * double(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::high_resolution_clock::now() - t0).count());
        if (rank == 0) std::cout << "total execution time: " << s << " seconds" << std::endl;
        return code;
    }
    catch (const std::exception& e)
    {
        std::cout << e.what() << std::endl;
        return 1;
    }
}

int m3b_binary_main(int argc, const char* const* argv, int device)
{
    // app_main.cpp:75-79: run the subprogram, then report the wall time; an exception ends the run with its message
    try
    {
        auto t0 = std::chrono::high_resolution_clock::now();
        int code = m3b::binary_main(argc, argv, device);
        double s = 1e-9 * double(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::high_resolution_clock::now() - t0).count());
        std::cout << "total execution time: " << s << " seconds" << std::endl;
        return code;
    }
    catch (const std::exception& e)
    {
        std::cout << e.what() << std::endl;
        return 1;
    }
}

int m3b_h5_selftest(const char* write_path, const char* read_path, char* report, int report_len)
{
    // exercises every structure h5lite emits (test infrastructure for tests/test_h5lite.py) and, if read_path is given,
    // lists that file's root group through the reader
    try
    {
        using m3b::h5::type_t;
        std::string out;
        if (write_path && *write_path)
        {
            m3b::h5::writer_t w(write_path);
            struct inner_t { double a, b; };
            struct outer_t { double t; double v[2]; inner_t in; int n; int pad; };
            auto inner = type_t::compound(sizeof(inner_t), {type_t::member("a", offsetof(inner_t, a), type_t::f64()), type_t::member("b", offsetof(inner_t, b), type_t::f64())});
            auto outer = type_t::compound(sizeof(outer_t), {type_t::member("t", offsetof(outer_t, t), type_t::f64()),
                type_t::member("v", offsetof(outer_t, v), type_t::array(type_t::f64(), 2)), type_t::member("in", offsetof(outer_t, in), inner),
                type_t::member("n", offsetof(outer_t, n), type_t::i32())});
            std::vector<outer_t> series(5);
            for (int k = 0; k < 5; ++k) series[k] = {0.5 * k, {1.0 + k, 2.0 + k}, {10.0 * k, -1.0 * k}, k, 0};
            w.write("/series", outer, {5}, series.data(), true);
            w.write_double("/group/time", 3.25);
            w.write_int("/group/count", -7);
            w.write_string("/group/name", "write_checkpoint");
            w.write_string("/group/empty", "");
            int rational[2] = {22, 7};
            w.write("/group/iteration", type_t::array(type_t::i32(), 2), {}, rational, true);
            std::vector<double> field(6 * 4 * 3);
            for (std::size_t k = 0; k < field.size(); ++k) field[k] = 0.125 * double(k);
            w.write("/group/nested/field", type_t::array(type_t::f64(), 3), {6, 4}, field.data(), true);
            w.write("/group/nested/nothing", type_t::array(type_t::f64(), 3), {0, 0}, field.data(), true);
            w.require_group("/hollow");
            std::vector<double> values(300);
            for (int k = 0; k < 300; ++k)
            {
                char name[32];
                std::snprintf(name, sizeof(name), "8:%03d-%03d", k, 299 - k);
                values[k] = k;
                w.write(std::string("/many/") + name, type_t::f64(), {1}, &values[k]);
            }
            w.close();
        }
        if (read_path && *read_path)
        {
            m3b::h5::reader_t r(read_path);
            for (auto& k : r.keys("/")) out += k + (r.is_group("/" + k) ? "/ " : " ");
        }
        if (report && report_len > 0) std::snprintf(report, report_len, "%s", out.c_str());
        return 0;
    }
    catch (const std::exception& e)
    {
        if (report && report_len > 0) std::snprintf(report, report_len, "error: %s", e.what());
        return -1;
    }
}

int m3b_h5_selftest_shared(const char* whole_path, const char* shared_path, int parts)
{
    // One file written by `parts` processes (h5lite writer roles root / part, as the multi-GPU product writers use them) must
    // be byte-identical to the same content written by one process.  The roles run one after another here; tests/test_h5lite.py.
    try
    {
        using m3b::h5::type_t;
        using role_t = m3b::h5::writer_t::role_t;
        const int blocks = 37, cells = 5 * 5 * 3;
        std::vector<double> field(std::size_t(blocks) * cells);
        for (std::size_t k = 0; k < field.size(); ++k) field[k] = 1.0 / double(k + 1);
        auto describe = [&] (m3b::h5::writer_t& w, int me, int n)
        {
            // me < 0: everything; otherwise blocks [blocks me / n, blocks (me + 1) / n)
            w.write_double("/solution/time", 1.5);
            w.write_string("/run_config/outdir", "data");
            for (int b = 0; b < blocks; ++b)
            {
                char name[32];
                std::snprintf(name, sizeof(name), "6:%02d-%02d", b, blocks - b);
                const bool mine = me < 0 || (b >= blocks * me / n && b < blocks * (me + 1) / n);
                w.write(std::string("/solution/conserved_u/") + name, type_t::array(type_t::f64(), 3), {5, 5}, mine ? field.data() + std::size_t(b) * cells : nullptr, false, mine);
            }
            w.write("/solution/conserved_q/0:0-0", type_t::array(type_t::f64(), 3), {0, 0}, field.data());
            w.write_int("/schedule/write_checkpoint/num_times_performed", 3);
        };
        { m3b::h5::writer_t w(whole_path); describe(w, -1, 1); w.close(); }
        { m3b::h5::writer_t w(shared_path, role_t::root); describe(w, 0, parts); w.close(); }
        for (int p = 1; p < parts; ++p) { m3b::h5::writer_t w(shared_path, role_t::part); describe(w, p, parts); w.close(); }
        return 0;
    }
    catch (const std::exception&)
    {
        return -1;
    }
}

int m3b_exchange_transport(const m3b_solver_t* s) { return s->solver->has_device() ? s->solver->device().exchange_transport() : 0; }
uint64_t m3b_halo_bytes_per_exchange(const m3b_solver_t* s) { return s->solver->has_device() ? s->solver->device().halo_bytes_per_exchange() : 0; }
int m3b_block_size(const m3b_solver_t* s) { return s->solver->solver_data().block_size; }
int64_t m3b_num_cells(const m3b_solver_t* s) { return int64_t(s->solver->solver_data().num_cells()); }
int m3b_num_regular_blocks(const m3b_solver_t* s) { return s->solver->has_device() ? s->solver->device().num_regular_blocks() : -1; }

void m3b_tree_index(const m3b_solver_t* s, int64_t* out)
{
    const auto& d = s->solver->solver_data();
    for (int b = 0; b < d.num_owned; ++b)
    {
        const auto& idx = d.tree->index(d.global_block(b));
        out[3 * b + 0] = idx.level;
        out[3 * b + 1] = idx.i;
        out[3 * b + 2] = idx.j;
    }
}

void m3b_vertices(const m3b_solver_t* s, double* out) { copy_out(s->solver->solver_data().vertices(), out); }
void m3b_cell_centers(const m3b_solver_t* s, double* out) { copy_out(s->solver->solver_data().cell_centers(), out); }
void m3b_cell_areas(const m3b_solver_t* s, double* out) { copy_out(s->solver->solver_data().cell_areas(), out); }
void m3b_buffer_rate_field(const m3b_solver_t* s, double* out)
{
    const auto& d = s->solver->solver_data();
    std::memcpy(out, d.buffer_rate_field.data(), d.num_owned_cells() * sizeof(double));     // owned blocks come first
}

void m3b_initial_conserved_u(const m3b_solver_t* s, double* out)
{
    // stored field-major [3][B][NN]; the ABI layout is block-major [B][3][NN]
    const auto& d = s->solver->solver_data();
    const std::size_t NN = d.cells_per_block(), FS = d.num_local_cells();
    for (int b = 0; b < d.num_owned; ++b)
        for (int q = 0; q < 3; ++q)
            std::memcpy(out + (std::size_t(b) * 3 + q) * NN, &d.initial_conserved_u[q * FS + b * NN], NN * sizeof(double));
}

double m3b_recommended_time_step(const m3b_solver_t* s) { return s->solver->solver_data().recommended_time_step; }
double m3b_gst_suppr_radius(const m3b_solver_t* s) { return s->solver->solver_data().gst_suppr_radius; }
double m3b_density_floor(const m3b_solver_t* s) { return s->solver->solver_data().density_floor; }

int m3b_config_get(const m3b_solver_t* s, const char* key, char* out, int out_len)
{
    const auto& all = s->solver->run_config().all();
    auto it = all.find(key);
    if (it == all.end()) return 1;
    auto ss = std::ostringstream();
    std::visit([&ss] (const auto& v) { ss << v; }, it->second);
    std::strncpy(out, ss.str().c_str(), out_len > 0 ? out_len - 1 : 0);
    if (out_len > 0) out[out_len - 1] = 0;
    return 0;
}

m3b_solution_t* m3b_solution_create(m3b_solver_t* s)
{
    try {
        auto u = std::make_unique<m3b_solution>();
        u->solution = s->solver->create_solution();
        return u.release();
    }
    catch (const std::exception& e) { s->error = e.what(); return nullptr; }
}

m3b_solution_t* m3b_solution_clone(m3b_solver_t* s, const m3b_solution_t* u)
{
    try {
        auto r = std::make_unique<m3b_solution>();
        r->solution = s->solver->clone(u->solution);
        return r.release();
    }
    catch (const std::exception& e) { s->error = e.what(); return nullptr; }
}

void m3b_solution_destroy(m3b_solution_t* u) { delete u; }

int m3b_solution_set_conserved(m3b_solver_t* s, m3b_solution_t* u, const double* host)
{
    return guarded(s, [&] { s->solver->invalidate(); s->solver->device().upload(host, *u->solution.conserved_u); s->solver->device().sync(); return M3B_OK; });
}

int m3b_solution_get_conserved(m3b_solver_t* s, const m3b_solution_t* u, double* host)
{
    return guarded(s, [&] { s->solver->device().download(*u->solution.conserved_u, host); return M3B_OK; });
}

void m3b_solution_set_scalars(m3b_solution_t* u, const double* in43) { u->solution.set_scalars(in43); }
void m3b_solution_get_scalars(const m3b_solution_t* u, double* out43) { u->solution.get_scalars(out43); }

int m3b_maximum_timestep(m3b_solver_t* s, const m3b_solution_t* u, double* dt_out)
{
    return guarded(s, [&] { *dt_out = s->solver->maximum_timestep(u->solution); return M3B_OK; });
}

int m3b_advance(m3b_solver_t* s, const m3b_solution_t* in, double dt, int safe_mode, m3b_solution_t* out)
{
    return guarded(s, [&] { s->error.clear(); return int(s->solver->advance(in->solution, dt, safe_mode != 0, out->solution)); });
}

int m3b_solution_combine(m3b_solver_t* s, const m3b_solution_t* a, const m3b_solution_t* b, double b0, m3b_solution_t* out)
{
    return guarded(s, [&] { s->solver->combine(a->solution, b->solution, b0, out->solution); return M3B_OK; });
}

int m3b_next_solution(m3b_solver_t* s, m3b_solution_t* u, double* dt_used, int* fell_back)
{
    return guarded(s, [&]
    {
        bool fb = false;
        s->error.clear();
        auto st = s->solver->next_solution(u->solution, dt_used, &fb);
        if (fell_back) *fell_back = fb;
        return int(st);
    });
}

int m3b_run_steps(m3b_solver_t* s, m3b_solution_t* u, int count, int* num_fallbacks)
{
    return guarded(s, [&]
    {
        int fallbacks = 0;
        s->error.clear();
        for (int n = 0; n < count; ++n)
        {
            bool fb = false;
            auto st = s->solver->next_solution(u->solution, nullptr, &fb);
            fallbacks += fb;
            if (st != status_ok) { if (num_fallbacks) *num_fallbacks = fallbacks; return int(st); }
        }
        if (num_fallbacks) *num_fallbacks = fallbacks;
        return M3B_OK;
    });
}

int m3b_advance_host(m3b_solver_t* s, const double* u_in, const double* scalars_in, double dt, int safe_mode, double* u_out, double* scalars_out)
{
    return guarded(s, [&]
    {
        auto& solver = *s->solver;
        auto in = solution_t();
        auto out = solution_t();
        s->error.clear();
        in.conserved_u = solver.new_field();
        in.set_scalars(scalars_in);
        solver.device().upload(u_in, *in.conserved_u);
        auto st = solver.advance(in, dt, safe_mode != 0, out);
        if (st == status_ok || st == status_negative_density)
        {
            solver.device().download(*out.conserved_u, u_out);
            out.get_scalars(scalars_out);
        }
        return int(st);
    });
}

int m3b_next_solution_host(m3b_solver_t* s, const double* u_in, const double* scalars_in, double* u_out, double* scalars_out, double* dt_used, int* fell_back)
{
    return guarded(s, [&]
    {
        auto& solver = *s->solver;
        auto u = solution_t();
        bool fb = false;
        s->error.clear();
        u.conserved_u = solver.new_field();
        u.set_scalars(scalars_in);
        solver.device().upload(u_in, *u.conserved_u);
        auto st = solver.next_solution(u, dt_used, &fb, /*speculate*/ false);   // a one-shot call: nothing to queue ahead
        if (fell_back) *fell_back = fb;
        if (st == status_ok)
        {
            solver.device().download(*u.conserved_u, u_out);
            u.get_scalars(scalars_out);
        }
        return int(st);
    });
}

int m3b_two_body_state(const double* e, double t, double* out)
{
    auto el = elements_t{e[0], e[1], e[2], e[3], e[4], e[5], e[6], e[7], e[8], e[9]};
    auto s = two_body_state(el, t);
    const point_mass_t* b[2] = {&s.body1, &s.body2};
    for (int k = 0; k < 2; ++k)
    {
        out[5 * k + 0] = b[k]->mass; out[5 * k + 1] = b[k]->x; out[5 * k + 2] = b[k]->y;
        out[5 * k + 3] = b[k]->vx;   out[5 * k + 4] = b[k]->vy;
    }
    return M3B_OK;
}

int m3b_orbital_elements(const double* b, double t, double* out)
{
    auto s = two_body_t{{b[0], b[1], b[2], b[3], b[4]}, {b[5], b[6], b[7], b[8], b[9]}};
    auto e = elements_t();
    if (! orbital_elements(s, t, e)) return M3B_UNBOUND_ORBIT;
    const double v[10] = {e.pomega, e.tau, e.cm_position_x, e.cm_position_y, e.cm_velocity_x, e.cm_velocity_y,
                          e.separation, e.total_mass, e.mass_ratio, e.eccentricity};
    for (int k = 0; k < 10; ++k) out[k] = v[k];
    return M3B_OK;
}

int m3b_num_messages(const m3b_solver_t* s) { return int(s->solver->last_messages().size()); }
const char* m3b_message(const m3b_solver_t* s, int n) { return s->solver->last_messages().at(n).c_str(); }
void m3b_set_quiet(m3b_solver_t* s, int quiet) { s->solver->set_quiet(quiet != 0); }
void m3b_set_pipelining(m3b_solver_t* s, int on) { try { s->solver->set_pipelining(on != 0); } catch (...) {} }
uint64_t m3b_kernel_launches(const m3b_solver_t* s) { return s->solver->has_device() ? s->solver->device().launch_count() : 0; }
void m3b_stage_timing(m3b_solver_t* s, int enable) { if (s->solver->has_device()) s->solver->device().set_stage_timing(enable != 0); }

int m3b_stage_timing_read(m3b_solver_t* s, double* total_ms, uint64_t* launches)
{
    return guarded(s, [&]
    {
        s->solver->device().collect_stage_timing();
        *total_ms = s->solver->device().stage_kernel_ms_total();
        *launches = s->solver->device().stage_kernel_launches();
        return M3B_OK;
    });
}

int m3b_exchange_timing(m3b_solver_t* s, double* out8)
{
    return guarded(s, [&]
    {
        auto& d = s->solver->device();
        if (d.exchange_transport() == 0) return M3B_ERROR;      // one rank: nothing to report
        d.collect_stage_timing();
        // [0] steps, [1] exposed wait us, [2] result wait us, [3] exchange us, [4] exchanges, [5] bytes pushed, [6..7] reserved
        out8[0] = double(d.result_waits_timed);
        out8[1] = d.exposed_wait_us_total;
        out8[2] = d.result_wait_us_total;
        out8[3] = d.exchange_us_total;
        out8[4] = double(d.exchanges_timed);
        out8[5] = double(d.halo_bytes_per_exchange()) * double(d.exchanges_timed);
        out8[6] = d.unpack_cta_wait_us_total;
        out8[7] = double(d.unpack_cta_waits);
        d.unpack_cta_wait_us_total = 0.0; d.unpack_cta_waits = 0;
        d.exposed_wait_us_total = d.result_wait_us_total = d.exchange_us_total = 0.0;
        d.exchanges_timed = d.result_waits_timed = 0;
        return M3B_OK;
    });
}

int m3b_set_stream(m3b_solver_t* s, void* cuda_stream)
{
    return guarded(s, [&] { s->solver->device().set_stream(cuda_stream); return M3B_OK; });
}

void m3b_synchronize(m3b_solver_t* s) { try { s->solver->device().sync(); } catch (...) {} }

} // extern "C"
