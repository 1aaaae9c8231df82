/**
 * quadtree.hpp -- static block quadtree of the `binary` subprogram (host side).
 *
 * Topology and vertex coordinates reproduce the reference bit for bit:
 *   - tree_index_t with periodic next/prev        (Mara3 src/core_tree.hpp:86-219)
 *   - leaf order = depth first, child n = bx + 2*by, i.e. Morton order with x the
 *     least-significant bit                       (core_tree.hpp:156-159, 334-337)
 *   - create_vertex_quadtree: `depth` refinement passes, the PASS number (not the
 *     node level) feeds the focusing predicate    (mesh_tree_operators.hpp:158-198,
 *                                                  subprog_binary.cpp:166-184)
 *   - ensure_valid_quadtree: 2:1 balance          (mesh_tree_operators.hpp:90-139)
 *
 * The reference stores an (N+1)x(N+1) array of 2-vectors per block; those arrays
 * are exact Cartesian products (midpoint refinement `(a + b) * 0.5` of equal
 * values is exact), so each block here keeps two 1-d coordinate arrays instead.
 */
#pragma once
#include <array>
#include <cstdint>
#include <vector>

namespace m3b
{
    struct tree_index_t
    {
        int level = 0;
        long i = 0;
        long j = 0;

        long extent() const { return 1L << level; }
        tree_index_t next_on(int axis) const { return shifted(axis, +1); }
        tree_index_t prev_on(int axis) const { return shifted(axis, -1); }
        tree_index_t parent() const { return {level - 1, i / 2, j / 2}; }
        tree_index_t child(int n) const { return {level + 1, 2 * i + (n & 1), 2 * j + (n >> 1)}; }
        int orthant_in_parent() const { return int(i % 2) + 2 * int(j % 2); }
        bool operator==(const tree_index_t& o) const { return level == o.level && i == o.i && j == o.j; }

        /** Morton key of this index's lower-left corner at `at_level` (x is the low bit). */
        std::uint64_t morton(int at_level) const;

    private:
        tree_index_t shifted(int axis, long d) const
        {
            long n = extent();
            return axis == 0 ? tree_index_t{level, (i + n + d) % n, j} : tree_index_t{level, i, (j + n + d) % n};
        }
    };

    /** How a block sees the same-level region across one of its faces. */
    enum class neighbor_kind_t : int { same = 0, coarser = 1, finer = 2 };

    struct face_neighbor_t
    {
        neighbor_kind_t kind = neighbor_kind_t::same;
        int leaf[4] = {-1, -1, -1, -1};  // same: leaf[0]; coarser: leaf[0] = parent leaf; finer: the 4 children (n = bx + 2*by)
        int bx = 0, by = 0;              // coarser: orthant of the (virtual) same-level neighbour inside its parent
    };

    class quadtree_t
    {
    public:
        struct node_t
        {
            int child[4] = {-1, -1, -1, -1};
            int leaf = -1;
            std::vector<double> xv, yv;    // unit-square vertex coordinates (leaves only)
            bool is_leaf() const { return child[0] < 0; }
        };

        /** Build the balanced tree for (block_size, depth, focus_factor, focus_index). */
        quadtree_t(int block_size, int depth, double focus_factor, double focus_index);

        int block_size() const { return N; }
        int num_leaves() const { return int(leaf_nodes.size()); }
        int max_level() const;
        const tree_index_t& index(int leaf) const { return leaf_index[leaf]; }
        const node_t& leaf_node(int leaf) const { return nodes[leaf_nodes[leaf]]; }

        /** node id at exactly this index (leaf or not), or -1 */
        int find_node(const tree_index_t& index) const;
        /** leaf id at exactly this index, or -1 */
        int find_leaf(const tree_index_t& index) const;
        const node_t& node(int id) const { return nodes[id]; }
        int root() const { return 0; }

        /** The region across face `side` (0: x-, 1: x+, 2: y-, 3: y+) of a leaf, periodic. */
        face_neighbor_t face_neighbor(int leaf, int side) const;

        /** Same-level leaf at offset (di, dj) in {-1,0,1}^2 (periodic), or -1. */
        int same_level_neighbor(int leaf, int di, int dj) const;

    private:
        void split(int node_id);
        int depth_below(int node_id) const;
        bool over_refined(const tree_index_t& index) const;
        void enumerate(int node_id, tree_index_t index);

        int N;
        std::vector<node_t> nodes;
        std::vector<int> leaf_nodes;
        std::vector<tree_index_t> leaf_index;
    };
}
