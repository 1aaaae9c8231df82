#include "partition.hpp"
#include <algorithm>
#include <set>
#include <stdexcept>
#include <tuple>

using namespace m3b;

std::vector<int> m3b::partition_offsets(int num_leaves, int nranks)
{
    auto offsets = std::vector<int>(nranks + 1);
    for (int r = 0; r <= nranks; ++r) offsets[r] = int((long(num_leaves) * r) / nranks);
    return offsets;
}

namespace
{
    /**
     * The (owner, source block, di, dj) regions rank `r` needs from other ranks, ordered by
     * (owner, source block, di, dj) and without duplicates.  (di, dj) = (0, 0) is the whole block.
     *
     * A block whose eight same-level neighbours are all leaves is updated by the fused kernel and needs
     * their two-cell edge strips / 2x2 corners.  A block that touches a refinement jump is updated by the
     * any-tree kernels, which read primitives AND PLM gradients of its face neighbours through the
     * prolongation / restriction rules (mesh_tree_operators.hpp:223-252): the gradients of those
     * neighbours are computed on this rank too, from THEIR face neighbours -- so the rank stores whole
     * copies of the face neighbours (layer 1) and of the face neighbours of those (layer 2).
     */
    std::vector<std::tuple<int, int, int, int>> remote_needs(const quadtree_t& tree, const std::vector<int>& owner, int first, int count, int r, bool all_general)
    {
        auto needs = std::set<std::tuple<int, int, int, int>>();
        auto whole = std::set<int>();
        auto want_whole = [&] (int leaf) { if (leaf >= 0 && owner[leaf] != r) whole.insert(leaf); };

        for (int b = first; b < first + count; ++b)
        {
            bool regular = ! all_general;
            for (int di = -1; di <= 1; ++di)
                for (int dj = -1; dj <= 1; ++dj)
                    if ((di || dj) && tree.same_level_neighbor(b, di, dj) < 0) regular = false;

            if (regular)
            {
                for (int di = -1; di <= 1; ++di)
                    for (int dj = -1; dj <= 1; ++dj)
                    {
                        if (di == 0 && dj == 0) continue;
                        int n = tree.same_level_neighbor(b, di, dj);
                        if (owner[n] != r) needs.insert({owner[n], n, di, dj});
                    }
                continue;
            }
            for (int side = 0; side < 4; ++side)
            {
                auto f1 = tree.face_neighbor(b, side);
                for (int l1 : f1.leaf)
                {
                    if (l1 < 0) continue;
                    want_whole(l1);
                    for (int side2 = 0; side2 < 4; ++side2)
                        for (int l2 : tree.face_neighbor(l1, side2).leaf) want_whole(l2);
                }
            }
        }
        // a whole block makes its strips redundant
        for (auto it = needs.begin(); it != needs.end(); )
        {
            if (whole.count(std::get<1>(*it))) it = needs.erase(it); else ++it;
        }
        for (int leaf : whole) needs.insert({owner[leaf], leaf, 0, 0});
        return {needs.begin(), needs.end()};
    }
}

partition_t m3b::make_partition(const quadtree_t& tree, int rank, int nranks, bool all_general)
{
    if (nranks < 1 || rank < 0 || rank >= nranks) throw std::invalid_argument("make_partition: bad rank / nranks");

    auto p = partition_t();
    auto offsets = partition_offsets(tree.num_leaves(), nranks);
    p.rank = rank;
    p.nranks = nranks;
    p.num_global = tree.num_leaves();
    p.first_owned = offsets[rank];
    p.num_owned = offsets[rank + 1] - offsets[rank];
    p.owner.resize(p.num_global);
    for (int r = 0; r < nranks; ++r) for (int b = offsets[r]; b < offsets[r + 1]; ++b) p.owner[b] = r;
    p.global_to_local.assign(p.num_global, -1);
    p.send.resize(nranks);
    p.recv.resize(nranks);

    for (int k = 0; k < p.num_owned; ++k)
    {
        p.local_to_global.push_back(p.first_owned + k);
        p.global_to_local[p.first_owned + k] = k;
    }
    if (nranks == 1) return p;
    if (p.num_owned == 0) throw std::invalid_argument("make_partition: more ranks than leaf blocks");

    // what this rank receives: ghosts are numbered after the owned blocks in ascending global id
    auto mine = remote_needs(tree, p.owner, p.first_owned, p.num_owned, rank, all_general);
    auto ghosts = std::set<int>();
    for (auto& [o, n, di, dj] : mine) ghosts.insert(n);
    for (int g : ghosts)
    {
        p.global_to_local[g] = p.num_local();
        p.local_to_global.push_back(g);
    }
    for (auto& [o, n, di, dj] : mine) p.recv[o].push_back({p.global_to_local[n], di, dj});

    // what this rank sends: the same lists as computed by each peer, restricted to blocks owned here
    for (int peer = 0; peer < nranks; ++peer)
    {
        if (peer == rank) continue;
        for (auto& [o, n, di, dj] : remote_needs(tree, p.owner, offsets[peer], offsets[peer + 1] - offsets[peer], peer, all_general))
            if (o == rank) p.send[peer].push_back({p.global_to_local[n], di, dj});
    }
    return p;
}
