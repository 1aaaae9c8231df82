/**
 * solver_data.hpp -- static per-run data of the `binary` hot path (host side).
 *
 * Counterpart of the reference's solver_data_t (Mara3 src/subprog_binary.hpp:74-104)
 * as built by create_solver_data (src/subprog_binary_solver_data.cpp:18-115) and of
 * the initial condition create_solution / create_disk_profile
 * (src/subprog_binary.cpp:105-153, 196-227).  Field values are bit-identical to
 * the reference's (same operation order, same libm); the storage is flattened:
 * blocks in the reference's traversal (Morton) order, structure-of-arrays per
 * field, each block row-major (N, N) with y fastest.
 */
#pragma once
#include <memory>
#include <vector>
#include "config.hpp"
#include "quadtree.hpp"
#include "two_body.hpp"

namespace m3b
{
    struct solver_data_t
    {
        // --- scalars (subprog_binary.hpp:76-97)
        double sink_rate, recommended_time_step, softening_radius, gst_suppr_radius, domain_radius, sink_radius;
        double density_floor, mach_number, alpha, alpha_cutoff_radius, nu, plm_theta, cfl_number, begin_live_binary;
        int rk_order;
        bool axisymmetric_cs2, conserve_linear_p, fixed_dt, no_accretion_force;
        int block_size;
        elements_t initial_elements;

        // --- mesh
        std::shared_ptr<quadtree_t> tree;
        int num_blocks = 0;
        std::vector<double> xv;                  // [B][N+1] block vertex x coordinates (already * domain_radius)
        std::vector<double> yv;                  // [B][N+1]
        std::vector<double> buffer_rate_field;   // [B][N][N]
        std::vector<double> initial_conserved_u; // [3][B][N][N]   (sigma, px, py)

        std::size_t cells_per_block() const { return std::size_t(block_size) * block_size; }
        std::size_t num_cells() const { return cells_per_block() * num_blocks; }
        double spacing(int level) const { return 2.0 * domain_radius / block_size / (1 << level); }

        // --- derived arrays in the reference's layouts (for parity tests, diagnostics and I/O)
        std::vector<double> vertices() const;      // [B][2][N+1][N+1]
        std::vector<double> cell_centers() const;  // [B][2][N][N]
        std::vector<double> cell_areas() const;    // [B][N][N]
    };

    /** create_solver_data (subprog_binary_solver_data.cpp:18-115). */
    solver_data_t create_solver_data(const config_t& run_config);

    /** The disk model at a point (subprog_binary.cpp:105-153): returns (sigma, vx, vy). */
    void disk_profile(const config_t& run_config, double x, double y, double prim[3]);
}
