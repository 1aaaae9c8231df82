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
    /** The (source block, di, dj) triples rank `r` needs from other ranks, ordered by
     *  (owner, source block, di, dj) and without duplicates. */
    std::vector<std::tuple<int, int, int, int>> remote_needs(const quadtree_t& tree, const std::vector<int>& owner, int first, int count, int r)
    {
        auto needs = std::set<std::tuple<int, int, int, int>>();

        for (int b = first; b < first + count; ++b)
            for (int di = -1; di <= 1; ++di)
                for (int dj = -1; dj <= 1; ++dj)
                {
                    if (di == 0 && dj == 0) continue;
                    int n = tree.same_level_neighbor(b, di, dj);
                    if (n < 0)
                        throw std::invalid_argument("multi-GPU runs need a uniform-level tree in this build: block " + std::to_string(b)
                            + " touches a refinement jump (raise focus_factor, or run nested trees on one GPU)");
                    if (owner[n] != r) needs.insert({owner[n], n, di, dj});
                }
        return {needs.begin(), needs.end()};
    }
}

partition_t m3b::make_partition(const quadtree_t& tree, int rank, int nranks)
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
    auto mine = remote_needs(tree, p.owner, p.first_owned, p.num_owned, rank);
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
        for (auto& [o, n, di, dj] : remote_needs(tree, p.owner, offsets[peer], offsets[peer + 1] - offsets[peer], peer))
            if (o == rank) p.send[peer].push_back({p.global_to_local[n], di, dj});
    }
    return p;
}
