/**
 * partition.hpp -- Morton-contiguous partition of the leaf blocks over ranks (one rank per GPU)
 * and the guard-zone exchange plan between ranks.
 *
 * The reference runs in one address space: its guard fill reads the neighbour leaf directly
 * (Mara3 src/subprog_binary_scheme.cpp:132-142).  Here the leaves, which the reference
 * already traverses in Morton order (src/core_tree.hpp:156-159), are cut into `nranks`
 * contiguous ranges of (nearly) equal length; a rank stores its own blocks followed by ghost
 * copies of every remote block that one of its blocks touches (faces and corners, periodic), and
 * before each RK stage the two-cell-deep edge strips / 2x2 corners those blocks need are moved
 * rank to rank.  Blocks at a refinement jump read whole neighbour blocks two face-layers deep
 * (partition.cpp: remote_needs), which travel as whole-block regions.  Both sides derive the same ordered lists from the global tree, so no handshake
 * is needed.
 */
#pragma once
#include <vector>
#include "quadtree.hpp"

namespace m3b
{
    /** One strip or corner of a block: which cells move.  (di, dj) is the position of the SOURCE
     *  block relative to the block that needs it, so the cells are the source's rows facing back:
     *  di = -1 -> rows N-2, N-1;  di = +1 -> rows 0, 1;  di = 0 -> all rows (same for dj / columns);
     *  (0, 0) is the whole block. */
    struct halo_region_t
    {
        int block;          // local block id (owned on the sending side, ghost on the receiving side)
        int di, dj;
        int num_cells(int N) const { return (di ? 2 : N) * (dj ? 2 : N); }
    };

    struct partition_t
    {
        int rank = 0, nranks = 1;
        int num_global = 0;                     // leaves of the whole tree
        int first_owned = 0, num_owned = 0;     // this rank owns global leaves [first_owned, first_owned + num_owned)
        std::vector<int> local_to_global;       // owned blocks in Morton order, then ghosts (ascending global id)
        std::vector<int> global_to_local;       // -1 where the block is not stored on this rank
        std::vector<int> owner;                 // rank owning each global leaf
        std::vector<std::vector<halo_region_t>> send;   // [peer] strips of owned blocks, in the peer's receive order
        std::vector<std::vector<halo_region_t>> recv;   // [peer] strips of ghost blocks, in arrival order

        int num_local() const { return int(local_to_global.size()); }
        int num_ghost() const { return num_local() - num_owned; }
        bool is_distributed() const { return nranks > 1; }
    };

    /** First leaf of each rank's range: balanced by leaf count (all leaves hold N^2 cells). */
    std::vector<int> partition_offsets(int num_leaves, int nranks);

    /** Build the partition and exchange plan seen by `rank` (any 2:1 balanced tree). */
    /** all_general: every block is updated by the any-tree kernels (conserve_linear_p = 0): whole-block ghosts everywhere. */
    partition_t make_partition(const quadtree_t& tree, int rank, int nranks, bool all_general = false);
}
