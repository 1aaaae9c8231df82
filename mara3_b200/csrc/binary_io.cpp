#include "binary_io.hpp"
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <sys/stat.h>
#include "h5lite.hpp"

using namespace m3b;
using h5::type_t;

schedule_t::task_t& schedule_t::at(const std::string& name)
{
    auto it = tasks.find(name);
    if (it == tasks.end()) throw std::out_of_range("no task scheduled with the name " + name);
    return it->second;
}

const schedule_t::task_t& schedule_t::at(const std::string& name) const
{
    auto it = tasks.find(name);
    if (it == tasks.end()) throw std::out_of_range("no task scheduled with the name " + name);
    return it->second;
}

std::string m3b::format_tree_index(int level, int i, int j)
{
    // std::setw(1 + std::log10(1 << level)): the double is truncated to int
    const int width = int(1 + std::log10(double(1 << level)));
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%d:%0*d-%0*d", level, width, i, width, j);
    return buf;
}

namespace
{
    /** full_orbital_elements_t as a nested compound (subprog_binary_io.cpp:43-90); elements_t has the same layout. */
    type_t elements_type()
    {
        static_assert(sizeof(elements_t) == 80, "elements_t must be ten doubles");
        auto inner = type_t::compound(32, {type_t::member("separation", 0, type_t::f64()), type_t::member("total_mass", 8, type_t::f64()),
                                           type_t::member("mass_ratio", 16, type_t::f64()), type_t::member("eccentricity", 24, type_t::f64())});
        return type_t::compound(80, {type_t::member("pomega", 0, type_t::f64()), type_t::member("tau", 8, type_t::f64()),
                                     type_t::member("cm_position_x", 16, type_t::f64()), type_t::member("cm_position_y", 24, type_t::f64()),
                                     type_t::member("cm_velocity_x", 32, type_t::f64()), type_t::member("cm_velocity_y", 40, type_t::f64()),
                                     type_t::member("elements", 48, inner)});
    }

    /** time_series_sample_t: members in the order the reference inserts them, offsets of the struct (subprog_binary_io.cpp:96-125). */
    type_t sample_type()
    {
        using s = time_series_sample_t;
        auto v2 = type_t::array(type_t::f64(), 2);
        auto el = elements_type();
        return type_t::compound(sizeof(s), {
            type_t::member("time", offsetof(s, time), type_t::f64()),
            type_t::member("disk_mass", offsetof(s, disk_mass), type_t::f64()),
            type_t::member("disk_angular_momentum", offsetof(s, disk_angular_momentum), type_t::f64()),
            type_t::member("mass_accreted_on", offsetof(s, mass_accreted_on), v2),
            type_t::member("angular_momentum_accreted_on", offsetof(s, angular_momentum_accreted_on), v2),
            type_t::member("integrated_torque_on", offsetof(s, integrated_torque_on), v2),
            type_t::member("work_done_on", offsetof(s, work_done_on), v2),
            type_t::member("mass_ejected", offsetof(s, mass_ejected), type_t::f64()),
            type_t::member("angular_momentum_ejected", offsetof(s, angular_momentum_ejected), type_t::f64()),
            type_t::member("orbital_elements_acc", offsetof(s, orbital_elements_acc), el),
            type_t::member("orbital_elements_grav", offsetof(s, orbital_elements_grav), el),
            type_t::member("orbital_elements", offsetof(s, orbital_elements), el),
            type_t::member("position_of_mass1", offsetof(s, position_of_mass1), v2),
            type_t::member("position_of_mass2", offsetof(s, position_of_mass2), v2)});
    }

    void write_config(h5::writer_t& w, const std::string& group, const config_t& config)
    {
        w.require_group(group);
        for (const auto& item : config.all())
        {
            const auto path = group + "/" + item.first;
            switch (item.second.index())
            {
                case 0: w.write_int(path, std::get<int>(item.second)); break;
                case 1: w.write_double(path, std::get<double>(item.second)); break;
                case 2: w.write_string(path, std::get<std::string>(item.second)); break;
            }
        }
    }

    std::string leaf_name(const solver_data_t& data, int global_block)
    {
        auto index = data.tree->index(global_block);
        return format_tree_index(index.level, int(index.i), int(index.j));
    }
}

time_series_sample_t m3b::make_time_series_sample(binary_solver_t& solver, const solution_t& u)
{
    auto sample = time_series_sample_t();
    auto bodies = two_body_state(u.orbital_elements, u.time);
    double totals[2];
    solver.device().disk_totals(*u.conserved_u, totals);
    sample.time = u.time;
    sample.disk_mass = totals[0];
    sample.disk_angular_momentum = totals[1];
    sample.mass_ejected = u.mass_ejected;
    sample.angular_momentum_ejected = u.angular_momentum_ejected;
    for (int k = 0; k < 2; ++k)
    {
        sample.mass_accreted_on[k] = u.mass_accreted_on[k];
        sample.angular_momentum_accreted_on[k] = u.angular_momentum_accreted_on[k];
        sample.integrated_torque_on[k] = u.integrated_torque_on[k];
        sample.work_done_on[k] = u.work_done_on[k];
    }
    std::memcpy(sample.orbital_elements_acc, &u.orbital_elements_acc, 80);
    std::memcpy(sample.orbital_elements_grav, &u.orbital_elements_grav, 80);
    std::memcpy(sample.orbital_elements, &u.orbital_elements, 80);
    sample.position_of_mass1[0] = bodies.body1.x; sample.position_of_mass1[1] = bodies.body1.y;
    sample.position_of_mass2[0] = bodies.body2.x; sample.position_of_mass2[1] = bodies.body2.y;
    return sample;
}

namespace
{
    /**
     * One product file written by all ranks.  Every rank describes the same groups and datasets (the state is replicated apart
     * from the blocks, whose datasets a rank fills only for the blocks it owns) and so computes the same layout; rank 0 then
     * writes the structure, the small datasets and its own blocks, leaving holes for the others, and once it is done the other
     * ranks store their blocks at their addresses in that file (h5lite writer roles root / part).  Nothing is gathered: at
     * 16384^2 a rank moves its own 0.8 GB instead of rank 0 moving 6.4 GB three times.  The outcome is collective -- if any rank
     * fails (directory not writable, disk full), every rank throws, so that no rank walks on into a step its peers never take.
     */
    template<typename Describe>
    void write_shared_file(const std::string& filename, m3b::binary_solver_t& solver, Describe&& describe)
    {
        auto& device = solver.device();
        const int rank = device.rank(), nranks = device.num_ranks_();
        using role_t = m3b::h5::writer_t::role_t;
        auto message = std::string();
        auto attempt = [&] (auto&& fn) { try { fn(); return 1.0; } catch (const std::exception& e) { message = e.what(); return 0.0; } };
        if (nranks == 1)
        {
            auto w = m3b::h5::writer_t(filename);
            describe(w);
            w.close();
            return;
        }
        auto w = m3b::h5::writer_t(filename, rank == 0 ? role_t::root : role_t::part);
        double ok = attempt([&] { describe(w); });
        if (rank == 0 && ok == 1.0) ok = attempt([&] { w.close(); });
        bool all_ok = true;
        for (double v : device.all_gather_scalar(ok)) all_ok = all_ok && v == 1.0;          // (also orders the root's close before the parts')
        if (all_ok && rank != 0) ok = attempt([&] { w.close(); });
        for (double v : device.all_gather_scalar(all_ok ? ok : 0.0)) all_ok = all_ok && v == 1.0;
        if (! all_ok) throw std::runtime_error("writing " + filename + " failed on " + (message.empty() ? std::string("another rank") : "this rank: " + message));
    }
}

void m3b::write_checkpoint(const std::string& filename, binary_solver_t& solver, const state_t& state)
{
    const auto& data = solver.solver_data();
    const auto& u = state.solution;
    const int N = data.block_size, B = data.num_blocks, first = data.partition.first_owned, owned = data.num_owned;
    const std::size_t NN = std::size_t(N) * N;

    // this rank's blocks: conserved_u/<level:ii-jj> is (N, N) of double[3], the raw image of the reference's
    // std::tuple<sigma, px, py>, which libstdc++ lays out in reverse: (py, px, sigma)  (SURVEY.md 8c caveat 1)
    auto planes = std::vector<double>(std::size_t(owned) * 3 * NN);
    solver.device().download(*u.conserved_u, planes.data());
    auto cells = std::vector<double>(planes.size());
    for (int b = 0; b < owned; ++b)
        for (int q = 0; q < 3; ++q)
        {
            const double* src = planes.data() + (std::size_t(b) * 3 + q) * NN;
            double* dst = cells.data() + std::size_t(b) * NN * 3 + (2 - q);
            for (std::size_t c = 0; c < NN; ++c) dst[3 * c] = src[c];
        }
    planes = std::vector<double>();

    write_shared_file(filename, solver, [&] (h5::writer_t& w)
    {
        auto v2 = type_t::array(type_t::f64(), 2), v3 = type_t::array(type_t::f64(), 3), el = elements_type();

        // ---- /solution
        w.write_double("/solution/time", u.time);
        int iteration[2] = {u.iteration_num, u.iteration_den};
        w.write("/solution/iteration", type_t::array(type_t::i32(), 2), {}, iteration, true);

        // conserve_linear_p = 0: the state is conserved_q = (sigma, Sr, Lz) -- its tuple image is (Lz, Sr, sigma) all the same
        const std::string used = data.conserve_linear_p ? "/solution/conserved_u" : "/solution/conserved_q";
        const std::string unused = data.conserve_linear_p ? "/solution/conserved_q" : "/solution/conserved_u";
        w.require_group(used);
        for (int b = 0; b < B; ++b)
        {
            const bool mine = b >= first && b < first + owned;
            w.write(used + "/" + leaf_name(data, b), v3, {std::uint64_t(N), std::uint64_t(N)},
                    mine ? cells.data() + std::size_t(b - first) * NN * 3 : nullptr, false, mine);
        }
        // the unused variable set is a default tree: one leaf "0:0-0" holding an empty array
        w.write(unused + "/0:0-0", v3, {0, 0}, cells.data());

        w.write("/solution/mass_accreted_on", v2, {}, u.mass_accreted_on, true);
        w.write("/solution/angular_momentum_accreted_on", v2, {}, u.angular_momentum_accreted_on, true);
        w.write("/solution/integrated_torque_on", v2, {}, u.integrated_torque_on, true);
        w.write("/solution/work_done_on", v2, {}, u.work_done_on, true);
        w.write_double("/solution/mass_ejected", u.mass_ejected);
        w.write_double("/solution/angular_momentum_ejected", u.angular_momentum_ejected);
        w.write("/solution/orbital_elements_acc", el, {}, &u.orbital_elements_acc, true);
        w.write("/solution/orbital_elements_grav", el, {}, &u.orbital_elements_grav, true);
        w.write("/solution/orbital_elements", el, {}, &u.orbital_elements, true);

        // ---- /schedule, /time_series, /run_config
        w.require_group("/schedule");
        for (const auto& t : state.schedule.tasks)
        {
            w.write_string("/schedule/" + t.first + "/name", t.second.name);
            w.write_int("/schedule/" + t.first + "/num_times_performed", t.second.num_times_performed);
            w.write_double("/schedule/" + t.first + "/last_performed", t.second.last_performed);
        }
        w.write("/time_series", sample_type(), {std::uint64_t(state.time_series.size())}, state.time_series.data());
        write_config(w, "/run_config", solver.run_config());
    });
}

void m3b::write_diagnostics(const std::string& filename, binary_solver_t& solver, const solution_t& u)
{
    const auto& data = solver.solver_data();
    const int N = data.block_size, B = data.num_blocks, V = N + 1, first = data.partition.first_owned, owned = data.num_owned;
    const std::size_t NN = std::size_t(N) * N;

    // this rank's blocks: sigma, v_r, v_phi (subprog_binary_diagnostics.cpp:48-82) and the vertices (N + 1, N + 1) of (x, y)
    auto fields = std::vector<double>(std::size_t(owned) * 3 * NN);
    solver.device().diagnostic_fields(*u.conserved_u, fields.data());
    auto verts = std::vector<double>(std::size_t(owned) * V * V * 2);
    for (int b = 0; b < owned; ++b)
    {
        const auto& leaf = data.tree->leaf_node(first + b);             // the whole tree is known on every rank
        for (int i = 0; i < V; ++i)
            for (int j = 0; j < V; ++j)
            {
                verts[((std::size_t(b) * V + i) * V + j) * 2 + 0] = leaf.xv[i] * data.domain_radius;
                verts[((std::size_t(b) * V + i) * V + j) * 2 + 1] = leaf.yv[j] * data.domain_radius;
            }
    }

    write_shared_file(filename, solver, [&] (h5::writer_t& w)
    {
        auto v2 = type_t::array(type_t::f64(), 2);
        write_config(w, "/run_config", solver.run_config());
        w.write_double("/time", u.time);
        const char* names[3] = {"sigma", "radial_velocity", "phi_velocity"};
        for (const char* n : names) w.require_group(std::string("/") + n);
        w.require_group("/vertices");
        for (int b = 0; b < B; ++b)
        {
            const bool mine = b >= first && b < first + owned;
            auto idx = leaf_name(data, b);
            w.write("/vertices/" + idx, v2, {std::uint64_t(V), std::uint64_t(V)}, mine ? verts.data() + std::size_t(b - first) * V * V * 2 : nullptr, false, mine);
            for (int q = 0; q < 3; ++q)
                w.write(std::string("/") + names[q] + "/" + idx, type_t::f64(), {std::uint64_t(N), std::uint64_t(N)},
                        mine ? fields.data() + (std::size_t(b - first) * 3 + q) * NN : nullptr, false, mine);
        }
        auto bodies = two_body_state(u.orbital_elements, u.time);
        double p1[2] = {bodies.body1.x, bodies.body1.y}, p2[2] = {bodies.body2.x, bodies.body2.y};
        w.write("/position_of_mass1", v2, {}, p1, true);
        w.write("/position_of_mass2", v2, {}, p2, true);
    });
}

std::map<std::string, std::string> m3b::read_checkpoint_config(const std::string& filename)
{
    auto r = h5::reader_t(filename);
    auto out = std::map<std::string, std::string>();
    for (const auto& key : r.keys("/run_config"))
    {
        const auto path = "/run_config/" + key;
        auto t = r.type(path);
        std::ostringstream ss;
        ss.precision(17);
        if (t.kind == type_t::kind_t::i32) ss << r.read_int(path);
        else if (t.kind == type_t::kind_t::f64) ss << r.read_double(path);
        else ss << r.read_string(path);
        out[key] = ss.str();
    }
    return out;
}

state_t m3b::read_checkpoint(const std::string& filename, binary_solver_t& solver)
{
    const auto& data = solver.solver_data();
    const int N = data.block_size, B = data.num_owned;
    const std::size_t NN = std::size_t(N) * N;
    auto r = h5::reader_t(filename);
    auto v2 = type_t::array(type_t::f64(), 2), v3 = type_t::array(type_t::f64(), 3), el = elements_type();
    auto state = state_t();
    auto& u = state.solution;
    auto get = [&r] (const std::string& path, const type_t& t, void* out) { auto b = r.read(path, t); std::memcpy(out, b.data(), b.size()); };

    u.time = r.read_double("/solution/time");
    int iteration[2];
    get("/solution/iteration", type_t::array(type_t::i32(), 2), iteration);
    u.iteration_num = iteration[0]; u.iteration_den = iteration[1];

    const std::string used = data.conserve_linear_p ? "/solution/conserved_u" : "/solution/conserved_q";
    auto names = r.keys(used);
    if (int(names.size()) != data.num_blocks) throw std::runtime_error("restart file has " + std::to_string(names.size()) + " blocks, the run configuration makes " + std::to_string(data.num_blocks));
    auto planes = std::vector<double>(std::size_t(B) * 3 * NN);
    for (int b = 0; b < B; ++b)
    {
        const auto path = used + "/" + leaf_name(data, data.global_block(b));     // every rank reads its own blocks
        if (! r.exists(path)) throw std::runtime_error("restart file has no block " + path + " (different mesh?)");
        auto shape = r.shape(path);
        if (shape.size() != 2 || int(shape[0]) != N || int(shape[1]) != N) throw std::runtime_error("restart block " + path + " has the wrong shape");
        auto bytes = r.read(path, v3);
        auto cells = reinterpret_cast<const double*>(bytes.data());
        for (std::size_t c = 0; c < NN; ++c)
            for (int q = 0; q < 3; ++q)
                planes[(std::size_t(b) * 3 + q) * NN + c] = cells[c * 3 + (2 - q)];
    }
    u.conserved_u = solver.new_field();
    solver.device().upload(planes.data(), *u.conserved_u);

    get("/solution/mass_accreted_on", v2, u.mass_accreted_on);
    get("/solution/angular_momentum_accreted_on", v2, u.angular_momentum_accreted_on);
    get("/solution/integrated_torque_on", v2, u.integrated_torque_on);
    get("/solution/work_done_on", v2, u.work_done_on);
    u.mass_ejected = r.read_double("/solution/mass_ejected");
    u.angular_momentum_ejected = r.read_double("/solution/angular_momentum_ejected");
    get("/solution/orbital_elements_acc", el, &u.orbital_elements_acc);
    get("/solution/orbital_elements_grav", el, &u.orbital_elements_grav);
    get("/solution/orbital_elements", el, &u.orbital_elements);

    auto count = r.shape("/time_series");
    state.time_series.resize(count.empty() ? 0 : count[0]);
    if (! state.time_series.empty()) get("/time_series", sample_type(), state.time_series.data());

    for (const auto& task : r.keys("/schedule"))
    {
        auto t = schedule_t::task_t();
        t.name = task;
        t.num_times_performed = r.read_int("/schedule/" + task + "/num_times_performed");
        t.last_performed = r.read_double("/schedule/" + task + "/last_performed");
        state.schedule.tasks[task] = t;
    }
    return state;
}




// ============================================================================
namespace
{
    void require_dir(const std::string& dir)
    {
        // mara::filesystem::require_dir: mkdir -p
        std::string partial;
        for (std::size_t k = 0; k <= dir.size(); ++k)
        {
            if (k == dir.size() || dir[k] == '/')
            {
                if (! partial.empty() && partial != ".") ::mkdir(partial.c_str(), 0755);
            }
            if (k < dir.size()) partial += dir[k];
        }
        struct stat st;
        if (::stat(dir.c_str(), &st) != 0 || ! S_ISDIR(st.st_mode)) throw std::runtime_error("cannot create the output directory " + dir);
    }

    std::string numbered(const std::string& outdir, const char* prefix, int count)
    {
        char name[1024];
        std::snprintf(name, sizeof(name), "%s.%04d.h5", prefix, count);
        return outdir.empty() ? std::string(name) : outdir + "/" + name;
    }

    /** binary::run_tasks (subprog_binary.cpp:313-381): diagnostics, time series, checkpoint -- those due in the INCOMING state. */
    void run_tasks(binary_solver_t& solver, state_t& state, bool root)
    {
        const auto due = state.schedule;
        const auto outdir = solver.run_config().get_string("outdir");

        if (due.at("write_diagnostics").is_due)
        {
            auto fname = numbered(outdir, "diagnostics", state.schedule.at("write_diagnostics").num_times_performed);
            write_diagnostics(fname, solver, state.solution);
            if (root) std::printf("write diagnostics: %s\n", fname.c_str());
            state.schedule.mark_as_completed("write_diagnostics");
        }
        if (due.at("record_time_series").is_due)
        {
            state.time_series.push_back(make_time_series_sample(solver, state.solution));
            state.schedule.mark_as_completed("record_time_series");
        }
        if (due.at("write_checkpoint").is_due)
        {
            // the file number is the count BEFORE the task is marked complete; the state stored is the one AFTER
            auto fname = numbered(outdir, "chkpt", state.schedule.at("write_checkpoint").num_times_performed);
            state.schedule.mark_as_completed("write_checkpoint");
            write_checkpoint(fname, solver, state);
            if (root) std::printf("write checkpoint: %s\n", fname.c_str());
        }
    }

    /** binary::next_schedule (subprog_binary.cpp:295-301) with mark_tasks_in (app_schedule.hpp:180-196). */
    void mark_tasks(const config_t& config, double time, schedule_t& schedule)
    {
        const std::pair<const char*, const char*> tasks[3] = {{"write_checkpoint", "cpi"}, {"write_diagnostics", "dfi"}, {"record_time_series", "tsi"}};
        for (const auto& t : tasks)
        {
            const double interval = config.get_double(t.second) * 2 * M_PI;
            if (time - schedule.at(t.first).last_performed >= interval) schedule.mark_as_due(t.first, interval);
        }
    }
}

int m3b::binary_main(int argc, const char* const argv[], int device, int rank, int nranks, const unsigned char* nccl_unique_id)
{
    const bool root = rank == 0;        // several ranks: rank 0 prints and writes; the products are gathered (binary_io.cpp)

    // ---- create_run_config (subprog_binary.cpp:155-164): template <- restart file's run_config <- command line
    auto config = config_t::binary_template();
    std::string restart;
    for (int n = 0; n < argc; ++n)
        if (! std::strncmp(argv[n], "restart=", 8)) restart = argv[n] + 8;
    if (! restart.empty())
        for (const auto& item : read_checkpoint_config(restart))
            if (config.has(item.first)) config.set(item.first, item.second);
    {
        (void) config_t::from_argv(argc, argv);                // validates keys / duplicates / values
        for (int n = 0; n < argc; ++n)
        {
            auto arg = std::string(argv[n]);
            auto eq = arg.find('=');
            if (eq != std::string::npos) config.set(arg.substr(0, eq), arg.substr(eq + 1));
        }
    }
    auto solver = binary_solver_t(config, device, false, false, rank, nranks, nccl_unique_id);
    solver.set_quiet(! root);

    auto state = state_t();
    if (restart.empty())
    {
        state.solution = solver.create_solution();
        for (const char* task : {"write_checkpoint", "write_diagnostics", "record_time_series"}) state.schedule.create_and_mark_as_due(task);
    }
    else state = read_checkpoint(restart, solver);

    {
        // collective outcome: a rank whose peers could not create the output directory must not start stepping (it would wait
        // for their guard zones until the deadline)
        double ok = 1.0;
        auto message = std::string();
        if (root) { try { require_dir(config.get_string("outdir")); } catch (const std::exception& e) { ok = 0.0; message = e.what(); } }
        if (solver.has_device()) for (double v : solver.device().all_gather_scalar(ok)) ok = std::min(ok, v);
        if (ok != 1.0) throw std::runtime_error(message.empty() ? "rank 0 could not create the output directory " + config.get_string("outdir") : message);
    }
    if (root) config.pretty_print(std::cout, "config");
    run_tasks(solver, state, root);

    const double cells = double(solver.solver_data().num_cells());
    auto step = [&] ()
    {
        // next_state: the schedule is marked from the time BEFORE the step (subprog_binary.cpp:295-311)
        const double time_before = state.solution.time;
        double dt = 0.0;
        bool fell_back = false;
        auto st = solver.next_solution(state.solution, &dt, &fell_back);
        if (st != status_ok) throw std::runtime_error(solver.last_error().empty() ? "the step failed" : solver.last_error());
        mark_tasks(config, time_before, state.schedule);
        run_tasks(solver, state, root);
    };
    while (state.solution.time / (2 * M_PI) < config.get_double("tfinal"))
    {
        auto t0 = std::chrono::high_resolution_clock::now();
        step();
        auto ms = 1e-6 * double(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::high_resolution_clock::now() - t0).count());
        if (root)
        {
            std::printf("[%04d] orbits=%3.7lf kzps=%3.2lf\n", state.solution.iteration_num / state.solution.iteration_den,
                        state.solution.time / (2 * M_PI), cells / ms);
            std::fflush(stdout);
        }
    }
    step();     // the reference finishes with tasks(next(state)): one more step and its tasks
    return 0;
}
