#include "solver_data.hpp"
#include <algorithm>
#include <cmath>
#include <limits>
#include <stdexcept>

using namespace m3b;

namespace
{
    struct disk_model_t
    {
        double rs, rc, mach, s0, s1, mdot;
        int sign;

        explicit disk_model_t(const config_t& cfg)
        {
            rs   = cfg.get_double("softening_radius");
            rc   = cfg.get_double("disk_radius");
            mach = cfg.get_double("mach_number");
            mdot = cfg.get_double("mdot");
            sign = cfg.get_int("counter_rotate") ? -1 : 1;
            s0   = cfg.get_double("disk_mass") / (17.0618 * rc * rc);    // normalisation: subprog_binary.cpp:115
            s1   = cfg.get_double("ambient_density") * s0;
        }
        double sigma(double r) const
        {
            double x = r / rc;
            return s0 * std::exp(-0.5 * (x - 1) * (x - 1)) + s1;
        }
        double dp_dr(double r) const
        {
            double GM = 1.0, x = r / rc;
            return (GM / mach / mach / (r + rs)) * (x * (1 - x) * (1 - s1 / sigma(r)) - 1.0);
        }
        void evaluate(double x, double y, double prim[3]) const
        {
            double GM = 1.0;
            double r = std::sqrt(x * x + y * y);
            double vp = std::sqrt(GM / (r + rs) + dp_dr(r)) * sign;
            double vr = -mdot / (sigma(r) * 2 * M_PI * r) * (r > 2.0);
            prim[0] = sigma(r);
            prim[1] = vr * (x / r) + vp * (-y / r);
            prim[2] = vr * (y / r) + vp * ( x / r);
        }
    };
}

void m3b::disk_profile(const config_t& run_config, double x, double y, double prim[3])
{
    disk_model_t(run_config).evaluate(x, y, prim);
}

solver_data_t m3b::create_solver_data(const config_t& cfg, int rank, int nranks)
{
    auto d = solver_data_t();
    d.domain_radius       = cfg.get_double("domain_radius");
    d.mach_number         = cfg.get_double("mach_number");
    d.alpha_cutoff_radius = cfg.get_double("alpha_cutoff_radius");
    d.alpha               = cfg.get_double("alpha");
    d.nu                  = cfg.get_double("nu");
    d.sink_rate           = cfg.get_double("sink_rate");
    d.sink_radius         = cfg.get_double("sink_radius");
    d.softening_radius    = cfg.get_double("softening_radius");
    d.plm_theta           = cfg.get_double("plm_theta");
    d.begin_live_binary   = cfg.get_double("begin_live_binary");
    d.axisymmetric_cs2    = cfg.get_int("axisymmetric_cs2");
    d.conserve_linear_p   = cfg.get_int("conserve_linear_p");
    d.fixed_dt            = cfg.get_int("fixed_dt");
    d.rk_order            = cfg.get_int("rk_order");
    d.block_size          = cfg.get_int("block_size");
    d.no_accretion_force  = cfg.get_int("no_accretion_force");
    d.density_floor       = cfg.get_double("density_floor") * cfg.get_double("disk_mass");
    d.cfl_number          = cfg.get_double("cfl_number");

    auto method = cfg.get_string("reconstruct_method");     // parsed but unused by the scheme, as in the reference
    if (method != "plm" && method != "pcm")
        throw std::invalid_argument("invalid reconstruct_method '" + method + "', must be plm or pcm");

    d.initial_elements = elements_t();                      // create_binary_params (subprog_binary.cpp:186-194)
    d.initial_elements.total_mass   = 1.0;
    d.initial_elements.separation   = cfg.get_double("separation");
    d.initial_elements.mass_ratio   = cfg.get_double("mass_ratio");
    d.initial_elements.eccentricity = cfg.get_double("eccentricity");

    d.tree = std::make_shared<quadtree_t>(d.block_size, cfg.get_int("depth"), cfg.get_double("focus_factor"), cfg.get_double("focus_index"));
    d.partition = make_partition(*d.tree, rank, nranks, /*all_general*/ ! d.conserve_linear_p);
    d.num_blocks = d.tree->num_leaves();
    d.num_owned = d.partition.num_owned;
    d.num_local = d.partition.num_local();

    const int N = d.block_size, L = d.num_local;
    const auto disk = disk_model_t(cfg);
    const double buffer_rate = cfg.get_double("buffer_damping_rate");
    double min_dx = std::numeric_limits<double>::infinity(), min_dy = min_dx, max_v = 0.0;

    // the smallest vertex spacing is a property of the whole tree (solver_data.cpp:41-55); 1-d work per block
    for (int g = 0; g < d.num_blocks; ++g)
    {
        const auto& leaf = d.tree->leaf_node(g);
        for (int k = 0; k < N; ++k)
        {
            min_dx = std::min(min_dx, leaf.xv[k + 1] * d.domain_radius - leaf.xv[k] * d.domain_radius);
            min_dy = std::min(min_dy, leaf.yv[k + 1] * d.domain_radius - leaf.yv[k] * d.domain_radius);
        }
    }
    d.xv.resize(std::size_t(L) * (N + 1));
    d.yv.resize(std::size_t(L) * (N + 1));
    d.buffer_rate_field.resize(d.num_local_cells());
    d.initial_conserved_u.resize(3 * d.num_local_cells());

    for (int b = 0; b < L; ++b)
    {
        double* xv = &d.xv[std::size_t(b) * (N + 1)];
        double* yv = &d.yv[std::size_t(b) * (N + 1)];
        const auto& leaf = d.tree->leaf_node(d.global_block(b));

        for (int k = 0; k <= N; ++k)    // (block * domain_radius): subprog_binary.cpp:180-183
        {
            xv[k] = leaf.xv[k] * d.domain_radius;
            yv[k] = leaf.yv[k] * d.domain_radius;
        }
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j)
            {
                // centre = midpoint on axis 0 then on axis 1; the second midpoint averages two equal numbers
                double x = (xv[i] + xv[i + 1]) * 0.5;
                double y = (yv[j] + yv[j + 1]) * 0.5;
                double prim[3];
                disk.evaluate(x, y, prim);
                std::size_t k = (std::size_t(b) * N + i) * N + j;

                d.initial_conserved_u[0 * d.num_local_cells() + k] = prim[0];
                if (d.conserve_linear_p)
                {
                    d.initial_conserved_u[1 * d.num_local_cells() + k] = prim[0] * prim[1];
                    d.initial_conserved_u[2 * d.num_local_cells() + k] = prim[0] * prim[2];
                }
                else    // conserved_q = (sigma, Sr, Lz): to_conserved_angmom_per_area (physics_iso2d.hpp:263-272), held in the same array
                {
                    d.initial_conserved_u[1 * d.num_local_cells() + k] = prim[0] * (x * prim[1] + y * prim[2]);
                    d.initial_conserved_u[2 * d.num_local_cells() + k] = prim[0] * (x * prim[2] - y * prim[1]);
                }
                if (b < d.num_owned) max_v = std::max(max_v, std::sqrt(prim[1] * prim[1] + prim[2] * prim[2]));

                // buffer zone: rate * (1 + tanh(3 (r - domain_radius))) (solver_data.cpp:64-78)
                double r = std::pow(x * x + y * y, 0.5);
                d.buffer_rate_field[k] = buffer_rate * (1.0 + std::tanh(3.0 * (r - d.domain_radius)));
            }
    }
    d.min_spacing = std::min(min_dx, min_dy);
    d.gst_suppr_radius = cfg.get_double("source_term_softening") * d.min_spacing;
    d.max_velocity_local = max_v;
    d.set_max_velocity(max_v);      // exact on one rank; a distributed solver re-does this with the global maximum
    return d;
}

void solver_data_t::set_max_velocity(double global_max)
{
    recommended_time_step = min_spacing / std::max(1.0, global_max) * cfl_number;
}

std::vector<double> solver_data_t::vertices() const
{
    const int N = block_size, V = N + 1;
    auto out = std::vector<double>(std::size_t(num_owned) * 2 * V * V);
    for (int b = 0; b < num_owned; ++b)
        for (int i = 0; i < V; ++i)
            for (int j = 0; j < V; ++j)
            {
                out[((std::size_t(b) * 2 + 0) * V + i) * V + j] = xv[std::size_t(b) * V + i];
                out[((std::size_t(b) * 2 + 1) * V + i) * V + j] = yv[std::size_t(b) * V + j];
            }
    return out;
}

std::vector<double> solver_data_t::cell_centers() const
{
    const int N = block_size, V = N + 1;
    auto out = std::vector<double>(std::size_t(num_owned) * 2 * N * N);
    for (int b = 0; b < num_owned; ++b)
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j)
            {
                out[((std::size_t(b) * 2 + 0) * N + i) * N + j] = (xv[std::size_t(b) * V + i] + xv[std::size_t(b) * V + i + 1]) * 0.5;
                out[((std::size_t(b) * 2 + 1) * N + i) * N + j] = (yv[std::size_t(b) * V + j] + yv[std::size_t(b) * V + j + 1]) * 0.5;
            }
    return out;
}

std::vector<double> solver_data_t::cell_areas() const
{
    const int N = block_size, V = N + 1;
    auto out = std::vector<double>(std::size_t(num_owned) * N * N);
    for (int b = 0; b < num_owned; ++b)
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j)
                out[(std::size_t(b) * N + i) * N + j] =
                    (xv[std::size_t(b) * V + i + 1] - xv[std::size_t(b) * V + i]) *
                    (yv[std::size_t(b) * V + j + 1] - yv[std::size_t(b) * V + j]);
    return out;
}
