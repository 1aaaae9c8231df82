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
#include "partition.hpp"
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

        // --- mesh.  Per-block arrays are LOCAL to this rank: the blocks it owns (a Morton-contiguous
        // range of the global leaves, all of them on one GPU) followed by its ghost blocks.
        std::shared_ptr<quadtree_t> tree;        // the whole tree, on every rank
        partition_t partition;
        int num_blocks = 0;                      // leaves of the whole tree
        int num_owned = 0;                       // local blocks [0, num_owned) are updated by this rank
        int num_local = 0;                       // owned + ghosts
        std::vector<double> xv;                  // [L][N+1] block vertex x coordinates (already * domain_radius)
        std::vector<double> yv;                  // [L][N+1]
        std::vector<double> buffer_rate_field;   // [L][N][N]
        std::vector<double> initial_conserved_u; // [3][L][N][N]   (sigma, px, py); conserve_linear_p = 0: conserved_q = (sigma, Sr, Lz)
        double min_spacing = 0.0;                // over the whole tree
        double max_velocity_local = 0.0;         // over this rank's owned cells (reduce over ranks for recommended_time_step)

        int global_block(int local) const { return partition.local_to_global[local]; }
        std::size_t cells_per_block() const { return std::size_t(block_size) * block_size; }
        std::size_t num_cells() const { return cells_per_block() * num_blocks; }            // whole domain
        std::size_t num_local_cells() const { return cells_per_block() * num_local; }
        std::size_t num_owned_cells() const { return cells_per_block() * num_owned; }
        double spacing(int level) const { return 2.0 * domain_radius / block_size / (1 << level); }
        void set_max_velocity(double global_max);   // finishes recommended_time_step (solver_data.cpp:57-62, 102)

        // --- derived arrays of the OWNED blocks in the reference's layouts (parity tests, diagnostics, I/O)
        std::vector<double> vertices() const;      // [B][2][N+1][N+1]
        std::vector<double> cell_centers() const;  // [B][2][N][N]
        std::vector<double> cell_areas() const;    // [B][N][N]
    };

    /** create_solver_data (subprog_binary_solver_data.cpp:18-115) for one rank of `nranks`. */
    solver_data_t create_solver_data(const config_t& run_config, int rank = 0, int nranks = 1);

    /** The disk model at a point (subprog_binary.cpp:105-153): returns (sigma, vx, vy). */
    void disk_profile(const config_t& run_config, double x, double y, double prim[3]);
}
