/*
 * oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, not product code.
 *
 * A small driver (own code) that links the reference's OWN, unmodified
 * translation units, compiled in place from /root/reference/src by
 * oracle/Makefile, and calls only their public API:
 *
 *   binary::create_run_config / create_solver_data / create_solution /
 *   set_scheme_globals / maximum_timestep / advance, solution_t::operator+ / *
 *   (reference: src/subprog_binary.hpp:180-208)
 *
 * binary::next_solution is TU-local in the reference (`auto` return type,
 * src/subprog_binary.cpp:46,258-293), so its ~15 lines of logic (dt rule, RK1 /
 * RK2, safe-mode retry) are re-stated in next_solution() below.
 *
 * Outputs go to an "M3BD" dump file (see write_record) that tests and bench.py
 * read with numpy.  Fields are always extracted with mara::get<I>() -- never
 * through raw data() -- because the std::tuple memory image is reversed under
 * libstdc++ (SURVEY.md section 8c caveat 1).
 *
 * usage: mara_ref [--steps K] [--dump FILE] [--dump-steps a,b,c] [--mesh-only]
 *                 [--stages] [--timing] [--warmup W] [--max-seconds S] key=value ...
 * --max-seconds: stop the timed steps early once S seconds have passed (at least one step is taken; bench.py --impl reference)
 * Tokens without '=' are ignored by the reference's argv parser
 * (src/app_config.hpp:223-245), so harness options never clash with config keys.
 */
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <set>
#include <sstream>
#include <string>
#include <vector>
#include "subprog_binary.hpp"




//=============================================================================
struct dump_t
{
    std::FILE* f = nullptr;

    explicit dump_t(const std::string& fname)
    {
        if (! fname.empty())
        {
            f = std::fopen(fname.c_str(), "wb");
            if (! f) throw std::runtime_error("cannot open dump file " + fname);
            std::fwrite("M3BD0001", 1, 8, f);
        }
    }
    ~dump_t() { if (f) std::fclose(f); }

    void write_record(const std::string& name, std::uint8_t dtype, const std::vector<std::uint64_t>& dims, const void* data)
    {
        if (! f) return;
        std::uint32_t nl = name.size();
        std::uint32_t nd = dims.size();
        std::uint64_t count = 1;
        for (auto d : dims) count *= d;
        std::fwrite(&nl, 4, 1, f);
        std::fwrite(name.data(), 1, nl, f);
        std::fwrite(&dtype, 1, 1, f);
        std::fwrite(&nd, 4, 1, f);
        std::fwrite(dims.data(), 8, nd, f);
        std::fwrite(data, 8, count, f);
    }
    void f64(const std::string& name, const std::vector<std::uint64_t>& dims, const std::vector<double>& v) { write_record(name, 0, dims, v.data()); }
    void i64(const std::string& name, const std::vector<std::uint64_t>& dims, const std::vector<std::int64_t>& v) { write_record(name, 1, dims, v.data()); }
    void scalar(const std::string& name, double x) { f64(name, {1}, {x}); }
};




//=============================================================================
static std::vector<double> elements_vector(const mara::full_orbital_elements_t& E)
{
    return {E.pomega, E.tau, E.cm_position_x, E.cm_position_y, E.cm_velocity_x, E.cm_velocity_y,
            E.elements.separation, E.elements.total_mass, E.elements.mass_ratio, E.elements.eccentricity};
}

// conserve_linear_p = 0: the active variable set is conserved_q = (sigma, Sr, Lz); it is dumped under the same
// names ("conserved_u", "initial_conserved_u") so that every consumer of the dump format stays as it is
static bool g_q_mode = false;

static void dump_solution(dump_t& dump, const std::string& prefix, const binary::solution_t& s, std::size_t N)
{
    auto B = g_q_mode ? s.conserved_q.size() : s.conserved_u.size();
    auto U = std::vector<double>();
    U.reserve(B * 3 * N * N);

    // layout: [block][field][i][j], blocks in the reference's traversal order
    auto flatten = [&] (auto block)
    {
        for (std::size_t i = 0; i < N; ++i) for (std::size_t j = 0; j < N; ++j) U.push_back(mara::get<0>(block(i, j)).value);
        for (std::size_t i = 0; i < N; ++i) for (std::size_t j = 0; j < N; ++j) U.push_back(mara::get<1>(block(i, j)).value);
        for (std::size_t i = 0; i < N; ++i) for (std::size_t j = 0; j < N; ++j) U.push_back(mara::get<2>(block(i, j)).value);
    };
    if (g_q_mode) s.conserved_q.sink(flatten); else s.conserved_u.sink(flatten);
    dump.f64(prefix + "conserved_u", {B, 3, N, N}, U);
    dump.scalar(prefix + "time", s.time.value);
    dump.i64(prefix + "iteration", {2}, {s.iteration.get_numerator(), s.iteration.get_denominator()});
    dump.f64(prefix + "mass_accreted_on", {2}, {s.mass_accreted_on[0].value, s.mass_accreted_on[1].value});
    dump.f64(prefix + "angular_momentum_accreted_on", {2}, {s.angular_momentum_accreted_on[0].value, s.angular_momentum_accreted_on[1].value});
    dump.f64(prefix + "integrated_torque_on", {2}, {s.integrated_torque_on[0].value, s.integrated_torque_on[1].value});
    dump.f64(prefix + "work_done_on", {2}, {s.work_done_on[0].value, s.work_done_on[1].value});
    dump.scalar(prefix + "mass_ejected", s.mass_ejected.value);
    dump.scalar(prefix + "angular_momentum_ejected", s.angular_momentum_ejected.value);
    dump.f64(prefix + "orbital_elements_acc", {10}, elements_vector(s.orbital_elements_acc));
    dump.f64(prefix + "orbital_elements_grav", {10}, elements_vector(s.orbital_elements_grav));
    dump.f64(prefix + "orbital_elements", {10}, elements_vector(s.orbital_elements));
}

static void dump_solver_data(dump_t& dump, const binary::solver_data_t& d)
{
    auto N = d.block_size;
    auto B = d.vertices.size();
    auto index = std::vector<std::int64_t>();
    auto verts = std::vector<double>();
    auto xc    = std::vector<double>();
    auto dA    = std::vector<double>();
    auto br    = std::vector<double>();
    auto U0    = std::vector<double>();

    d.vertices.indexes().sink([&] (auto i)
    {
        index.push_back(i.level);
        index.push_back(i.coordinates[0]);
        index.push_back(i.coordinates[1]);
    });
    d.vertices.sink([&] (auto block)
    {
        for (std::size_t c = 0; c < 2; ++c)
            for (std::size_t i = 0; i <= N; ++i) for (std::size_t j = 0; j <= N; ++j) verts.push_back(block(i, j)[c].value);
    });
    d.cell_centers.sink([&] (auto block)
    {
        for (std::size_t c = 0; c < 2; ++c)
            for (std::size_t i = 0; i < N; ++i) for (std::size_t j = 0; j < N; ++j) xc.push_back(block(i, j)[c].value);
    });
    d.cell_areas.sink([&] (auto block)
    {
        for (std::size_t i = 0; i < N; ++i) for (std::size_t j = 0; j < N; ++j) dA.push_back(block(i, j).value);
    });
    d.buffer_rate_field.sink([&] (auto block)
    {
        for (std::size_t i = 0; i < N; ++i) for (std::size_t j = 0; j < N; ++j) br.push_back(block(i, j).value);
    });
    auto flatten0 = [&] (auto block)
    {
        for (std::size_t i = 0; i < N; ++i) for (std::size_t j = 0; j < N; ++j) U0.push_back(mara::get<0>(block(i, j)).value);
        for (std::size_t i = 0; i < N; ++i) for (std::size_t j = 0; j < N; ++j) U0.push_back(mara::get<1>(block(i, j)).value);
        for (std::size_t i = 0; i < N; ++i) for (std::size_t j = 0; j < N; ++j) U0.push_back(mara::get<2>(block(i, j)).value);
    };
    if (g_q_mode) d.initial_conserved_q.sink(flatten0); else d.initial_conserved_u.sink(flatten0);
    dump.i64("tree_index", {B, 3}, index);
    dump.f64("vertices", {B, 2, N + 1, N + 1}, verts);
    dump.f64("cell_centers", {B, 2, N, N}, xc);
    dump.f64("cell_areas", {B, N, N}, dA);
    dump.f64("buffer_rate_field", {B, N, N}, br);
    dump.f64("initial_conserved_u", {B, 3, N, N}, U0);
    dump.scalar("recommended_time_step", d.recommended_time_step.value);
    dump.scalar("gst_suppr_radius", d.gst_suppr_radius.value);
    dump.scalar("density_floor", d.density_floor.value);
}




//=============================================================================
struct step_report_t
{
    binary::solution_t solution;
    double dt = 0.0;
    bool fell_back = false;
};

// Re-statement of the TU-local binary::next_solution (src/subprog_binary.cpp:258-293).
static step_report_t next_solution(const binary::solution_t& solution, const binary::solver_data_t& solver_data, dump_t* stage_dump, const std::string& prefix)
{
    auto can_fail = [stage_dump, &prefix] (const binary::solution_t& s0, const binary::solver_data_t& solver_data, auto dt, bool safe_mode)
    {
        switch (solver_data.rk_order)
        {
            case 1: return binary::advance(s0, solver_data, dt, safe_mode);
            case 2:
            {
                auto b0 = mara::make_rational(1, 2);
                auto s1 = binary::advance(s0, solver_data, dt, safe_mode);
                if (stage_dump && ! safe_mode) dump_solution(*stage_dump, prefix + "stage1/", s1, solver_data.block_size);
                auto s2 = binary::advance(s1, solver_data, dt, safe_mode);
                if (stage_dump && ! safe_mode) dump_solution(*stage_dump, prefix + "stage2/", s2, solver_data.block_size);
                return s0 * b0 + s2 * (1 - b0);
            }
        }
        throw std::invalid_argument("next_solution");
    };

    auto dt = solver_data.fixed_dt
    ? solver_data.recommended_time_step
    : solver_data.cfl_number * binary::maximum_timestep(solution, solver_data);

    try {
        return {can_fail(solution, solver_data, dt, false), dt.value, false};
    }
    catch (const std::exception& e)
    {
        std::cout << e.what() << std::endl;
        return {can_fail(solution, solver_data, dt * 0.1, true), dt.value * 0.1, true};
    }
}




//=============================================================================
int main(int argc, const char* argv[])
{
    int steps = 0;
    int warmup = 0;
    bool mesh_only = false;
    bool stages = false;
    bool timing = false;
    double max_seconds = 0.0;
    std::string dump_name;
    std::set<int> dump_steps;

    for (int n = 1; n < argc; ++n)
    {
        std::string a = argv[n];
        auto next = [&] { return std::string(n + 1 < argc ? argv[++n] : ""); };

        if      (a == "--steps")     steps = std::stoi(next());
        else if (a == "--warmup")    warmup = std::stoi(next());
        else if (a == "--dump")      dump_name = next();
        else if (a == "--mesh-only") mesh_only = true;
        else if (a == "--stages")    stages = true;
        else if (a == "--timing")    timing = true;
        else if (a == "--max-seconds") max_seconds = std::stod(next());
        else if (a == "--dump-steps")
        {
            auto ss = std::stringstream(next());
            auto tok = std::string();
            while (std::getline(ss, tok, ',')) dump_steps.insert(std::stoi(tok));
        }
    }

    auto run_config  = binary::create_run_config(argc - 1, argv + 1);
    auto solver_data = binary::create_solver_data(run_config);
    g_q_mode         = ! solver_data.conserve_linear_p;
    auto dump        = dump_t(dump_name);
    auto N           = solver_data.block_size;
    auto B           = solver_data.vertices.size();

    dump_solver_data(dump, solver_data);

    if (mesh_only)
    {
        std::printf("blocks=%lu cells=%lu\n", (unsigned long)B, (unsigned long)(B * N * N));
        return 0;
    }

    binary::set_scheme_globals(run_config);
    auto solution = binary::create_solution(run_config);
    auto num_fallbacks = 0;
    auto dts = std::vector<double>();

    if (dump_steps.count(0)) dump_solution(dump, "step0/", solution, N);

    for (int n = 0; n < warmup; ++n)
    {
        solution = next_solution(solution, solver_data, nullptr, "").solution;
    }
    auto t0 = std::chrono::high_resolution_clock::now();

    for (int n = 1; n <= steps; ++n)
    {
        auto prefix = "step" + std::to_string(n) + "/";
        auto want = dump_steps.count(n) > 0;
        auto report = next_solution(solution, solver_data, (stages && want) ? &dump : nullptr, prefix);
        solution = report.solution;
        num_fallbacks += report.fell_back;
        dts.push_back(report.dt);
        if (want) dump_solution(dump, prefix, solution, N);
        if (max_seconds > 0.0 && std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count() > max_seconds)
        {
            steps = n;
            break;
        }
    }
    auto t1 = std::chrono::high_resolution_clock::now();
    auto seconds = std::chrono::duration<double>(t1 - t0).count();

    if (! dts.empty()) dump.f64("dt_history", {dts.size()}, dts);

    // Known-answer line: plain sum of sigma over all cells in traversal order.
    auto sum_sigma = 0.0;
    if (g_q_mode) solution.conserved_q.sink([&] (auto block) { for (auto u : block) sum_sigma += mara::get<0>(u).value; });
    else solution.conserved_u.sink([&] (auto block) { for (auto u : block) sum_sigma += mara::get<0>(u).value; });

    std::printf("blocks=%lu cells=%lu steps=%d fallbacks=%d t=%.17g sum_sigma=%.17g\n",
        (unsigned long)B, (unsigned long)(B * N * N), steps, num_fallbacks, solution.time.value, sum_sigma);

    if (timing && steps > 0)
    {
        std::printf("timing: threads=%d seconds=%.6f mzps=%.6f steps=%d\n",
            run_config.get_int("threaded"), seconds, double(B * N * N) * steps / seconds * 1e-6, steps);
    }
    return 0;
}
