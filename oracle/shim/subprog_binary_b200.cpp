/*
 * oracle/shim/subprog_binary_b200.cpp -- the reference-side binding of INTEGRATION.md, made real.
 *
 * This is the file a Mara3 maintainer would add to src/ IN PLACE OF subprog_binary_scheme.cpp: it defines the six
 * symbols the rest of the `binary` subprogram takes from that translation unit (src/subprog_binary.hpp:180-208) --
 *
 *     binary::set_scheme_globals      binary::advance            binary::maximum_timestep
 *     binary::recover_primitive       solution_t::operator+      solution_t::operator*
 *
 * -- on top of the C ABI of include/mara3_b200.h, and nothing else.  oracle/Makefile (target `shim`) compiles it against
 * the reference's own headers where they lie under /root/reference/src, links it with the reference's other four `binary`
 * translation units and oracle/ref_harness.cpp, and tests/test_gpu_shim.py compares that executable (mara_ref_b200: the
 * reference driving libmara3_b200.so) with the unmodified reference (mara_ref) on the same configuration.
 *
 * TEST INFRASTRUCTURE: it lives under oracle/ because it needs the reference's headers; the product never links it.
 * solution_t <-> flat arrays always goes through mara::get<I> (the std::tuple memory image is reversed under libstdc++).
 */
#include <cstdio>
#include <map>
#include <tuple>
#include <sstream>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>
#include "subprog_binary.hpp"
#include "mara3_b200.h"

namespace
{
    m3b_solver_t* solver = nullptr;
    bool q_mode = false;            // conserve_linear_p = 0: the evolved tree is conserved_q = (sigma, Sr, Lz)

    struct flat_t
    {
        std::vector<double> u;      // [block][3][N][N], blocks in the tree's traversal order (== the library's block order)
        double scalars[M3B_NUM_SCALARS];
    };

    template<typename Tree>
    void flatten_tree(const Tree& tree, std::vector<double>& u)
    {
        tree.sink([&u] (auto block)
        {
            const std::size_t ni = block.shape(0), nj = block.shape(1);
            for (std::size_t i = 0; i < ni; ++i) for (std::size_t j = 0; j < nj; ++j) u.push_back(mara::get<0>(block(i, j)).value);
            for (std::size_t i = 0; i < ni; ++i) for (std::size_t j = 0; j < nj; ++j) u.push_back(mara::get<1>(block(i, j)).value);
            for (std::size_t i = 0; i < ni; ++i) for (std::size_t j = 0; j < nj; ++j) u.push_back(mara::get<2>(block(i, j)).value);
        });
    }

    void put_elements(double* o, const mara::full_orbital_elements_t& e)
    {
        const double v[10] = {e.pomega, e.tau, e.cm_position_x, e.cm_position_y, e.cm_velocity_x, e.cm_velocity_y,
                              e.elements.separation, e.elements.total_mass, e.elements.mass_ratio, e.elements.eccentricity};
        for (int k = 0; k < 10; ++k) o[k] = v[k];
    }

    mara::full_orbital_elements_t get_elements(const double* v)
    {
        auto e = mara::full_orbital_elements_t();
        e.pomega = v[0]; e.tau = v[1]; e.cm_position_x = v[2]; e.cm_position_y = v[3]; e.cm_velocity_x = v[4]; e.cm_velocity_y = v[5];
        e.elements.separation = v[6]; e.elements.total_mass = v[7]; e.elements.mass_ratio = v[8]; e.elements.eccentricity = v[9];
        return e;
    }

    flat_t to_flat(const binary::solution_t& s)
    {
        auto f = flat_t();
        if (q_mode) flatten_tree(s.conserved_q, f.u); else flatten_tree(s.conserved_u, f.u);
        double* o = f.scalars;
        o[0] = s.time.value; o[1] = s.iteration.get_numerator(); o[2] = s.iteration.get_denominator();
        for (int k = 0; k < 2; ++k)
        {
            o[3 + k] = s.mass_accreted_on[k].value;      o[5 + k] = s.angular_momentum_accreted_on[k].value;
            o[7 + k] = s.integrated_torque_on[k].value;  o[9 + k] = s.work_done_on[k].value;
        }
        o[11] = s.mass_ejected.value; o[12] = s.angular_momentum_ejected.value;
        put_elements(o + 13, s.orbital_elements_acc); put_elements(o + 23, s.orbital_elements_grav); put_elements(o + 33, s.orbital_elements);
        return f;
    }

    /**
     * A tree with the topology of `like` (square blocks of one size), its cells taken from the flat array.  The position of a
     * block in the flat array is its ordinal in the tree's traversal (sink) order; tree.map() must not be trusted to visit the
     * children in that order (core_sequence.hpp:436-439 expands them as function arguments), so the ordinal is looked up by index.
     */
    template<typename Tree>
    Tree tree_from_flat(const Tree& like, const std::vector<double>& u)
    {
        using key_type = std::tuple<std::size_t, std::size_t, std::size_t>;
        auto key = [] (const mara::tree_index_t<2>& i) { return key_type(i.level, i.coordinates[0], i.coordinates[1]); };
        auto ordinal = std::map<key_type, std::size_t>();
        like.indexes().sink([&] (auto index) { const auto n = ordinal.size(); ordinal[key(index)] = n; });
        const std::size_t nn = u.size() / (3 * ordinal.size());
        auto n1 = std::size_t(0);
        while (n1 * n1 < nn) ++n1;

        return like.indexes().map([&u, &ordinal, key, nn, n1] (auto index)
        {
            using block_type = std::decay_t<decltype(like.at(index))>;
            using cell_type = typename block_type::value_type;
            using T0 = std::decay_t<decltype(mara::get<0>(std::declval<cell_type>()))>;
            using T1 = std::decay_t<decltype(mara::get<1>(std::declval<cell_type>()))>;
            using T2 = std::decay_t<decltype(mara::get<2>(std::declval<cell_type>()))>;
            const std::size_t at = ordinal.at(key(index)) * 3 * nn;
            auto out = nd::make_unique_array<cell_type>(n1, n1);
            for (std::size_t i = 0; i < n1; ++i)
                for (std::size_t j = 0; j < n1; ++j)
                    out(i, j) = mara::make_arithmetic_tuple(T0{u[at + i * n1 + j]}, T1{u[at + nn + i * n1 + j]}, T2{u[at + 2 * nn + i * n1 + j]});
            return std::move(out).shared();
        });
    }

    binary::solution_t from_flat(const binary::solution_t& like, const flat_t& f, mara::rational_number_t iteration)
    {
        const double* o = f.scalars;
        auto s = like;
        s.time = o[0];
        s.iteration = iteration;
        if (q_mode) s.conserved_q = tree_from_flat(like.conserved_q, f.u); else s.conserved_u = tree_from_flat(like.conserved_u, f.u);
        for (int k = 0; k < 2; ++k)
        {
            s.mass_accreted_on[k] = o[3 + k];      s.angular_momentum_accreted_on[k] = o[5 + k];
            s.integrated_torque_on[k] = o[7 + k];  s.work_done_on[k] = o[9 + k];
        }
        s.mass_ejected = o[11]; s.angular_momentum_ejected = o[12];
        s.orbital_elements_acc = get_elements(o + 13); s.orbital_elements_grav = get_elements(o + 23); s.orbital_elements = get_elements(o + 33);
        return s;
    }

    void need_solver()
    {
        if (! solver) throw std::logic_error("subprog_binary_b200: set_scheme_globals has not been called");
    }
}


//=============================================================================
void binary::set_scheme_globals(const mara::config_t& run_config)      // replaces subprog_binary_scheme.cpp:42-49
{
    // every item of the run configuration as "key=value": the library re-creates solver_data_t from the same 39 keys
    auto kv = std::vector<std::string>();
    for (auto item : run_config)
    {
        auto ss = std::ostringstream();
        ss.precision(17);
        std::visit([&ss] (auto v) { ss << v; }, item.second);
        if (item.first == "restart" || item.first == "outdir") continue;       // files are the caller's business
        kv.push_back(item.first + "=" + ss.str());
    }
    auto argv = std::vector<const char*>();
    for (auto& s : kv) argv.push_back(s.c_str());
    if (solver) m3b_solver_destroy(solver);
    solver = m3b_solver_create(int(argv.size()), argv.data(), /*device*/ 0, /*flags*/ 0);
    if (! solver) throw std::runtime_error(std::string("mara3_b200: ") + m3b_global_error());
    m3b_set_quiet(solver, 1);
    q_mode = run_config.get_int("conserve_linear_p") == 0;
}

binary::solution_t binary::advance(const solution_t& solution, const solver_data_t&, mara::unit_time<double> dt, bool safe_mode)
{
    need_solver();
    auto in = to_flat(solution), out = in;
    const int status = m3b_advance_host(solver, in.u.data(), in.scalars, dt.value, safe_mode ? 1 : 0, out.u.data(), out.scalars);
    if (status == M3B_NEGATIVE_DENSITY)
    {
        // validate_u prints the offending cells and throws; next_solution's catch retries in safe mode, as before
        for (int n = 0; n < m3b_num_messages(solver); ++n) std::printf("%s\n", m3b_message(solver, n));
        throw std::runtime_error("negative density in updated state");
    }
    if (status != M3B_OK) throw std::runtime_error(m3b_last_error(solver));
    // iteration is rational in the reference (subprog_binary.hpp: solution_t::iteration); one stage advances it by one
    return from_flat(solution, out, solution.iteration + 1);
}

double binary::maximum_timestep(const solution_t& solution, const solver_data_t&)
{
    need_solver();
    auto f = to_flat(solution);
    auto u = m3b_solution_create(solver);
    if (! u) throw std::runtime_error(m3b_last_error(solver));
    auto dt = 0.0;
    m3b_solution_set_scalars(u, f.scalars);
    int status = m3b_solution_set_conserved(solver, u, f.u.data());
    if (status == M3B_OK) status = m3b_maximum_timestep(solver, u, &dt);
    m3b_solution_destroy(u);
    if (status != M3B_OK) throw std::runtime_error(m3b_last_error(solver));
    return dt;
}

binary::quad_tree_t<mara::iso2d::primitive_t> binary::recover_primitive(const solution_t& solution, const solver_data_t& solver_data)
{
    // only the diagnostic products call this (subprog_binary_diagnostics.cpp:48-82): a host loop is all it needs
    if (q_mode)
    {
        return solution.conserved_q.pair(solver_data.cell_centers).map([] (auto qx)
        {
            auto [q, x] = qx;
            return nd::zip(q, x) | nd::apply([] (auto qq, auto xx) { return mara::iso2d::recover_primitive(qq, xx); }) | nd::to_shared();
        });
    }
    return solution.conserved_u.map([] (auto block)
    {
        return block | nd::map([] (auto u) { return mara::iso2d::recover_primitive(u); }) | nd::to_shared();
    });
}

binary::solution_t binary::solution_t::operator+(const solution_t& other) const
{
    auto a = to_flat(*this), b = to_flat(other);
    for (std::size_t k = 0; k < a.u.size(); ++k) a.u[k] += b.u[k];
    for (int k = 0; k < M3B_NUM_SCALARS; ++k) a.scalars[k] += b.scalars[k];
    return from_flat(*this, a, iteration + other.iteration);
}

binary::solution_t binary::solution_t::operator*(mara::rational_number_t scale) const
{
    auto a = to_flat(*this);
    const auto w = scale.as_double();
    for (auto& v : a.u) v *= w;
    for (int k = 0; k < M3B_NUM_SCALARS; ++k) a.scalars[k] *= w;
    return from_flat(*this, a, iteration * scale);
}
