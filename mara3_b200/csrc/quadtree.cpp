#include "quadtree.hpp"
#include <algorithm>
#include <cmath>
#include <functional>
#include <stdexcept>

using namespace m3b;

std::uint64_t tree_index_t::morton(int at_level) const
{
    std::uint64_t x = std::uint64_t(i) << (at_level - level);
    std::uint64_t y = std::uint64_t(j) << (at_level - level);
    std::uint64_t key = 0;
    for (int b = 0; b < 31; ++b)
    {
        key |= ((x >> b) & 1ULL) << (2 * b);
        key |= ((y >> b) & 1ULL) << (2 * b + 1);
    }
    return key;
}

quadtree_t::quadtree_t(int block_size, int depth, double focus_factor, double focus_index) : N(block_size)
{
    if (N < 2) throw std::invalid_argument("quadtree_t: block_size must be >= 2");

    // Root block: linspace(-1, 1, N + 1) on both axes, x0 + (x1 - x0) * i / (count - 1)
    // (core_ndarray.hpp:2544-2551).
    nodes.emplace_back();
    nodes[0].xv.resize(N + 1);
    nodes[0].yv.resize(N + 1);
    for (int k = 0; k <= N; ++k)
    {
        nodes[0].xv[k] = -1.0 + (1.0 - -1.0) * k / N;
        nodes[0].yv[k] = -1.0 + (1.0 - -1.0) * k / N;
    }

    // `depth` passes; in pass p every leaf present at the START of the pass whose centroid
    // radius is below focus_factor / pow(p, focus_index) splits once (pass 0: always).
    for (int pass = 0; pass < depth; ++pass)
    {
        double threshold = focus_factor / std::pow(double(pass), focus_index);
        int count = int(nodes.size());

        for (int id = 0; id < count; ++id)
        {
            if (! nodes[id].is_leaf()) continue;
            double cx = (nodes[id].xv[0] + nodes[id].xv[N]) * 0.5;
            double cy = (nodes[id].yv[0] + nodes[id].yv[N]) * 0.5;
            if (std::sqrt(cx * cx + cy * cy) < threshold) split(id);
        }
    }

    // 2:1 balance: flag every leaf with an over-refined face neighbour on the current
    // tree, split all flagged leaves, repeat until nothing is flagged.
    for (;;)
    {
        leaf_nodes.clear();
        leaf_index.clear();
        enumerate(0, {});
        auto flagged = std::vector<int>();
        for (int l = 0; l < num_leaves(); ++l) if (over_refined(leaf_index[l])) flagged.push_back(leaf_nodes[l]);
        if (flagged.empty()) break;
        for (int id : flagged) split(id);
    }
    for (int l = 0; l < num_leaves(); ++l) nodes[leaf_nodes[l]].leaf = l;
}

// refine_verts<2> on the two 1-d coordinate arrays: fine[2m] = c[m], fine[2m+1] =
// (c[m] + c[m+1]) * 0.5; child bx takes fine[bx*N .. bx*N + N]
// (mesh_prolong_restrict.hpp:148-159, 198-216, 311-322).
void quadtree_t::split(int id)
{
    auto refine = [this] (const std::vector<double>& c)
    {
        auto fine = std::vector<double>(2 * N + 1);
        for (int k = 0; k <= 2 * N; ++k)
        {
            int lo = k / 2, hi = k / 2 + (k % 2 == 0 ? 0 : 1);
            fine[k] = (c[lo] + c[hi]) * 0.5;
        }
        return fine;
    };
    auto fx = refine(nodes[id].xv);
    auto fy = refine(nodes[id].yv);

    for (int n = 0; n < 4; ++n)
    {
        int bx = n & 1, by = n >> 1;
        node_t ch;
        ch.xv.assign(fx.begin() + bx * N, fx.begin() + bx * N + N + 1);
        ch.yv.assign(fy.begin() + by * N, fy.begin() + by * N + N + 1);
        nodes.push_back(std::move(ch));
        nodes[id].child[n] = int(nodes.size()) - 1;
    }
    nodes[id].xv.clear();
    nodes[id].yv.clear();
}

int quadtree_t::depth_below(int id) const
{
    if (nodes[id].is_leaf()) return 0;
    int d = 0;
    for (int n = 0; n < 4; ++n) d = std::max(d, 1 + depth_below(nodes[id].child[n]));
    return d;
}

bool quadtree_t::over_refined(const tree_index_t& index) const
{
    for (int axis = 0; axis < 2; ++axis)
        for (auto nb : {index.next_on(axis), index.prev_on(axis)})
        {
            int id = find_node(nb);
            if (id >= 0 && depth_below(id) > 1) return true;
        }
    return false;
}

void quadtree_t::enumerate(int id, tree_index_t index)
{
    if (nodes[id].is_leaf())
    {
        leaf_nodes.push_back(id);
        leaf_index.push_back(index);
        return;
    }
    for (int n = 0; n < 4; ++n) enumerate(nodes[id].child[n], index.child(n));
}

int quadtree_t::max_level() const
{
    int l = 0;
    for (auto& idx : leaf_index) l = std::max(l, idx.level);
    return l;
}

int quadtree_t::find_node(const tree_index_t& index) const
{
    if (index.level < 0) return -1;
    int id = 0;
    for (int l = index.level - 1; l >= 0; --l)
    {
        if (nodes[id].is_leaf()) return -1;
        id = nodes[id].child[((index.i >> l) & 1) + 2 * ((index.j >> l) & 1)];
    }
    return id;
}

int quadtree_t::find_leaf(const tree_index_t& index) const
{
    int id = find_node(index);
    return (id >= 0 && nodes[id].is_leaf()) ? nodes[id].leaf : -1;
}

face_neighbor_t quadtree_t::face_neighbor(int leaf, int side) const
{
    auto idx = leaf_index[leaf];
    auto nb = side % 2 ? idx.next_on(side / 2) : idx.prev_on(side / 2);
    auto result = face_neighbor_t();

    // Same three cases, in the same order, as get_cell_block (mesh_tree_operators.hpp:223-252).
    if (int l = find_leaf(nb); l >= 0)
    {
        result.kind = neighbor_kind_t::same;
        result.leaf[0] = l;
        return result;
    }
    if (int l = nb.level > 0 ? find_leaf(nb.parent()) : -1; l >= 0)
    {
        result.kind = neighbor_kind_t::coarser;
        result.leaf[0] = l;
        result.bx = int(nb.i % 2);
        result.by = int(nb.j % 2);
        return result;
    }
    int id = find_node(nb);
    if (id < 0 || nodes[id].is_leaf()) throw std::logic_error("quadtree_t::face_neighbor (tree is not 2:1 balanced)");
    result.kind = neighbor_kind_t::finer;
    for (int n = 0; n < 4; ++n)
    {
        const auto& ch = nodes[nodes[id].child[n]];
        if (! ch.is_leaf()) throw std::logic_error("quadtree_t::face_neighbor (tree is not 2:1 balanced)");
        result.leaf[n] = ch.leaf;
    }
    return result;
}

int quadtree_t::same_level_neighbor(int leaf, int di, int dj) const
{
    auto idx = leaf_index[leaf];
    if (di > 0) idx = idx.next_on(0);
    if (di < 0) idx = idx.prev_on(0);
    if (dj > 0) idx = idx.next_on(1);
    if (dj < 0) idx = idx.prev_on(1);
    return find_leaf(idx);
}
