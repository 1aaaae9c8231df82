#include "h5lite.hpp"
#include <fcntl.h>
#include <unistd.h>
#include <sys/types.h>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <functional>
#include <stdexcept>

using namespace m3b::h5;

namespace
{
    constexpr std::uint64_t UNDEF = ~std::uint64_t(0);
    constexpr int LEAF_K = 4, INTERNAL_K = 16;                  // libhdf5's defaults (H5F_CRT_SYM_LEAF_DEF, H5F_CRT_BTREE_RANK)
    constexpr std::uint64_t SNOD_BYTES = 8 + 2 * LEAF_K * 40;
    constexpr std::uint64_t TREE_BYTES = 24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8;
    const unsigned char SIGNATURE[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};

    using bytes_t = std::vector<unsigned char>;

    void put(bytes_t& b, std::uint64_t v, int n) { for (int k = 0; k < n; ++k) b.push_back((unsigned char) (v >> (8 * k))); }
    void put_bytes(bytes_t& b, const void* p, std::size_t n) { auto c = static_cast<const unsigned char*>(p); b.insert(b.end(), c, c + n); }
    void pad8(bytes_t& b) { while (b.size() % 8) b.push_back(0); }
    std::uint64_t round8(std::uint64_t n) { return (n + 7) / 8 * 8; }

    /** Datatype message body (without the message header): class + version, bit field, size, properties. */
    void encode_type(const type_t& t, bytes_t& b)
    {
        switch (t.kind)
        {
        case type_t::kind_t::f64:       // IEEE little-endian double, as H5T_NATIVE_DOUBLE on x86-64
            put(b, 0x11, 1); put(b, 0x20, 1); put(b, 0x3f, 1); put(b, 0, 1); put(b, 8, 4);
            put(b, 0, 2); put(b, 64, 2); put(b, 52, 1); put(b, 11, 1); put(b, 0, 1); put(b, 52, 1); put(b, 1023, 4);
            break;
        case type_t::kind_t::i32:       // two's complement little-endian int, H5T_NATIVE_INT
            put(b, 0x10, 1); put(b, 0x08, 1); put(b, 0, 2); put(b, 4, 4);
            put(b, 0, 2); put(b, 32, 2);
            break;
        case type_t::kind_t::string:    // H5T_C_S1 resized: null-terminated, ASCII
            put(b, 0x13, 1); put(b, 0, 3); put(b, t.size, 4);
            break;
        case type_t::kind_t::array:     // version 2: rank, sizes, permutation indices, base type
            put(b, 0x2a, 1); put(b, 0, 3); put(b, t.size, 4);
            put(b, t.dims.size(), 1); put(b, 0, 3);
            for (auto d : t.dims) put(b, d, 4);
            for (std::size_t k = 0; k < t.dims.size(); ++k) put(b, k, 4);
            encode_type(*t.base, b);
            break;
        case type_t::kind_t::compound:  // version 2: name padded to a multiple of 8, byte offset, member type
            put(b, 0x26, 1); put(b, t.members.size(), 2); put(b, 0, 1); put(b, t.size, 4);
            for (const auto& m : t.members)
            {
                put_bytes(b, m.name.c_str(), m.name.size() + 1);
                for (std::size_t k = m.name.size() + 1; k % 8; ++k) b.push_back(0);     // the NAME is padded to a multiple of 8 bytes
                put(b, m.offset, 4);
                encode_type(*m.type, b);
            }
            break;
        }
    }

    void put_message(bytes_t& b, int type, int flags, const bytes_t& body)
    {
        put(b, type, 2); put(b, round8(body.size()), 2); put(b, flags, 1); put(b, 0, 3);
        put_bytes(b, body.data(), body.size());
        pad8(b);
    }
}




// ============================================================================
type_t type_t::f64() { return type_t(); }
type_t type_t::i32() { type_t t; t.kind = kind_t::i32; t.size = 4; return t; }
type_t type_t::string(std::size_t length) { type_t t; t.kind = kind_t::string; t.size = std::uint32_t(std::max<std::size_t>(1, length)); return t; }

type_t type_t::array(const type_t& base, std::uint32_t n)
{
    type_t t;
    t.kind = kind_t::array;
    t.size = base.size * n;
    t.dims = {n};
    t.base = std::make_shared<type_t>(base);
    return t;
}

type_t type_t::compound(std::uint32_t size, std::vector<member_t> members)
{
    type_t t;
    t.kind = kind_t::compound;
    t.size = size;
    t.members = std::move(members);
    return t;
}

type_t::member_t type_t::member(std::string name, std::size_t offset, const type_t& type)
{
    return {std::move(name), std::uint32_t(offset), std::make_shared<type_t>(type)};
}

bool type_t::operator==(const type_t& o) const
{
    if (kind != o.kind || size != o.size || dims != o.dims || members.size() != o.members.size()) return false;
    if (kind == kind_t::array && ! (*base == *o.base)) return false;
    for (std::size_t k = 0; k < members.size(); ++k)
        if (members[k].name != o.members[k].name || members[k].offset != o.members[k].offset || ! (*members[k].type == *o.members[k].type)) return false;
    return true;
}

std::string type_t::describe() const
{
    switch (kind)
    {
    case kind_t::f64: return "f64";
    case kind_t::i32: return "i32";
    case kind_t::string: return "string[" + std::to_string(size) + "]";
    case kind_t::array: return base->describe() + "[" + std::to_string(dims.empty() ? 0 : dims[0]) + "]";
    case kind_t::compound:
    {
        std::string s = "{";
        for (const auto& m : members) s += m.name + "@" + std::to_string(m.offset) + ":" + m.type->describe() + " ";
        return s + "}";
    }
    }
    return "?";
}




// ============================================================================
struct writer_t::node_t
{
    bool is_group = true;
    std::map<std::string, std::unique_ptr<node_t>> children;        // std::map iterates in strcmp order, the order of a symbol table

    // dataset
    type_t type;
    std::vector<std::uint64_t> shape;
    const void* data = nullptr;         // nullptr: another process stores these bytes (role root / part)
    bytes_t owned;
    std::uint64_t nbytes = 0;
    bool mine = false;                  // role part: this process stores them

    // layout (filled by allocate)
    std::uint64_t header_addr = 0, heap_addr = 0, heap_data_addr = 0, heap_bytes = 0, root_tree_addr = 0, data_addr = UNDEF;
    std::vector<std::uint64_t> name_offset;                         // heap offset of each child's name
    std::vector<std::uint64_t> snod_addr;
    std::vector<std::vector<std::uint64_t>> tree_addr;              // [level][node], level 0 points at symbol-table nodes
    bytes_t header;                                                 // dataset object header, built at allocation
};

writer_t::writer_t(std::string filename, role_t role) : filename(std::move(filename)), root(new node_t), role(role) {}
writer_t::~writer_t() { if (! closed) { try { close(); } catch (...) {} } }

writer_t::node_t* writer_t::descend(const std::string& path, bool create_last_as_group, std::string* leaf_name)
{
    node_t* node = root.get();
    std::size_t p = 0;
    std::vector<std::string> parts;
    while (p < path.size())
    {
        auto q = path.find('/', p);
        if (q == std::string::npos) q = path.size();
        if (q > p) parts.push_back(path.substr(p, q - p));
        p = q + 1;
    }
    for (std::size_t k = 0; k < parts.size(); ++k)
    {
        bool last = k + 1 == parts.size();
        if (last && ! create_last_as_group)
        {
            *leaf_name = parts[k];
            return node;
        }
        auto& child = node->children[parts[k]];
        if (! child) child.reset(new node_t);
        if (! child->is_group) throw std::invalid_argument("h5lite: " + path + " crosses a dataset");
        node = child.get();
    }
    if (! create_last_as_group) throw std::invalid_argument("h5lite: empty dataset name");
    return node;
}

void writer_t::require_group(const std::string& path) { descend(path, true, nullptr); }

void writer_t::write(const std::string& path, const type_t& type, const std::vector<std::uint64_t>& shape, const void* data, bool copy, bool mine)
{
    std::string name;
    node_t* parent = descend(path, false, &name);
    if (parent->children.count(name)) throw std::invalid_argument("h5lite: " + path + " already exists");
    auto d = std::unique_ptr<node_t>(new node_t);
    d->is_group = false;
    d->type = type;
    d->shape = shape;
    d->nbytes = type.size;
    for (auto n : shape) d->nbytes *= n;
    d->mine = mine;
    if (copy && data)
    {
        d->owned.assign(static_cast<const unsigned char*>(data), static_cast<const unsigned char*>(data) + d->nbytes);
        d->data = d->owned.data();
    }
    else d->data = data;
    parent->children[name] = std::move(d);
}

void writer_t::write_double(const std::string& path, double value) { write(path, type_t::f64(), {}, &value, true); }
void writer_t::write_int(const std::string& path, int value) { write(path, type_t::i32(), {}, &value, true); }

void writer_t::write_string(const std::string& path, const std::string& value)
{
    // core_hdf5.hpp:474-483: H5T_C_S1 resized to max(1, length); an empty string is one NUL byte
    auto t = type_t::string(value.size());
    bytes_t b(t.size, 0);
    std::memcpy(b.data(), value.data(), value.size());
    write(path, t, {}, b.data(), true);
}

void writer_t::close()
{
    if (closed) return;
    closed = true;

    // ---- pass 1: addresses, in the order the bytes will be streamed out
    std::uint64_t cursor = 96;      // after the superblock

    struct layout_t
    {
        static void allocate(node_t& n, std::uint64_t& cursor)
        {
            n.header_addr = cursor;
            if (! n.is_group)
            {
                bytes_t space, type, fill, layout, messages;
                put(space, 1, 1); put(space, n.shape.size(), 1); put(space, 0, 1); put(space, 0, 5);
                for (auto d : n.shape) put(space, d, 8);
                encode_type(n.type, type);
                put(fill, 2, 1); put(fill, 2, 1); put(fill, 2, 1); put(fill, 1, 1); put(fill, 0, 4);      // late allocation, fill if set, default value
                put_message(messages, 0x0001, 0, space);
                put_message(messages, 0x0003, 1, type);
                put_message(messages, 0x0005, 1, fill);
                std::uint64_t header_bytes = 16 + messages.size() + 8 + 24;
                n.data_addr = n.nbytes ? cursor + header_bytes : UNDEF;
                put(layout, 3, 1); put(layout, 1, 1); put(layout, n.data_addr, 8); put(layout, n.nbytes, 8);
                put_message(messages, 0x0008, 0, layout);
                put(n.header, 1, 1); put(n.header, 0, 1); put(n.header, 4, 2); put(n.header, 1, 4); put(n.header, messages.size(), 4); put(n.header, 0, 4);
                put_bytes(n.header, messages.data(), messages.size());
                if (n.header.size() != header_bytes) throw std::logic_error("h5lite: header size");
                cursor += header_bytes + round8(n.nbytes);
                return;
            }
            cursor += 40;                                               // object header with one symbol-table message
            n.heap_addr = cursor;
            cursor += 32;
            n.heap_data_addr = cursor;
            std::uint64_t offset = 8;                                   // offset 0 holds the empty name
            for (auto& c : n.children) { n.name_offset.push_back(offset); offset += round8(c.first.size() + 1); }
            n.heap_bytes = offset;
            cursor += n.heap_bytes;

            std::size_t count = n.children.size(), snods = (count + 2 * LEAF_K - 1) / (2 * LEAF_K);
            std::vector<std::size_t> per_level;
            for (std::size_t m = std::max<std::size_t>(1, (snods + 2 * INTERNAL_K - 1) / (2 * INTERNAL_K)); ; m = (m + 2 * INTERNAL_K - 1) / (2 * INTERNAL_K))
            {
                per_level.push_back(m);
                if (m == 1) break;
            }
            n.tree_addr.resize(per_level.size());
            for (std::size_t level = per_level.size(); level-- > 0; )     // root first
                for (std::size_t k = 0; k < per_level[level]; ++k) { n.tree_addr[level].push_back(cursor); cursor += TREE_BYTES; }
            n.root_tree_addr = n.tree_addr.back()[0];
            for (std::size_t k = 0; k < snods; ++k) { n.snod_addr.push_back(cursor); cursor += SNOD_BYTES; }
            for (auto& c : n.children) allocate(*c.second, cursor);
        }
    };
    layout_t::allocate(*root, cursor);
    const std::uint64_t eof = cursor;

    if (role == role_t::part)
    {
        // only this process's datasets, at the addresses every process has computed alike
        const int fd = ::open(filename.c_str(), O_WRONLY);
        if (fd < 0) throw std::runtime_error("h5lite: cannot open " + filename + " (written by the root process) for writing");
        struct part_t
        {
            static void store(const node_t& n, int fd, const std::string& filename)
            {
                if (n.is_group) { for (auto& c : n.children) store(*c.second, fd, filename); return; }
                if (! n.mine || ! n.data || ! n.nbytes) return;
                const char* p = static_cast<const char*>(n.data);
                std::uint64_t done = 0;
                while (done < n.nbytes)
                {
                    const ssize_t k = ::pwrite(fd, p + done, std::size_t(std::min<std::uint64_t>(n.nbytes - done, 1u << 30)), off_t(n.data_addr + done));
                    if (k <= 0) { ::close(fd); throw std::runtime_error("h5lite: write to " + filename + " failed"); }
                    done += std::uint64_t(k);
                }
            }
        };
        part_t::store(*root, fd, filename);
        if (::close(fd) != 0) throw std::runtime_error("h5lite: closing " + filename + " failed");
        return;
    }

    // ---- pass 2: stream
    std::FILE* f = std::fopen(filename.c_str(), "wb");
    if (! f) throw std::runtime_error("h5lite: cannot open " + filename + " for writing");
    std::uint64_t written = 0;
    auto emit = [&] (const void* p, std::size_t n)
    {
        if (n && ! p)
        {
            // leave a hole (the file is extended to its full length at the end)
            if (fseeko(f, off_t(n), SEEK_CUR) != 0) { std::fclose(f); throw std::runtime_error("h5lite: seek in " + filename + " failed"); }
        }
        else if (n && std::fwrite(p, 1, n, f) != n) { std::fclose(f); throw std::runtime_error("h5lite: write to " + filename + " failed"); }
        written += n;
    };
    auto emit_bytes = [&] (const bytes_t& b) { emit(b.data(), b.size()); };

    {
        bytes_t sb;
        put_bytes(sb, SIGNATURE, 8);
        put(sb, 0, 1); put(sb, 0, 1); put(sb, 0, 1); put(sb, 0, 1); put(sb, 0, 1);     // versions: superblock, free space, root entry, reserved, shared header
        put(sb, 8, 1); put(sb, 8, 1); put(sb, 0, 1);                                    // sizes of offsets and lengths
        put(sb, LEAF_K, 2); put(sb, INTERNAL_K, 2); put(sb, 0, 4);
        put(sb, 0, 8); put(sb, UNDEF, 8); put(sb, eof, 8); put(sb, UNDEF, 8);            // base, free space, end of file, driver info
        put(sb, 0, 8); put(sb, root->header_addr, 8); put(sb, 1, 4); put(sb, 0, 4); put(sb, root->root_tree_addr, 8); put(sb, root->heap_addr, 8);
        emit_bytes(sb);
    }

    struct stream_t
    {
        static void write(const node_t& n, const std::function<void(const void*, std::size_t)>& emit, std::uint64_t& written)
        {
            if (written != n.header_addr) throw std::logic_error("h5lite: layout drift");
            if (! n.is_group)
            {
                emit(n.header.data(), n.header.size());
                if (n.data || ! n.nbytes)
                {
                    emit(n.data, n.nbytes);
                    static const unsigned char zeros[8] = {0};
                    emit(zeros, round8(n.nbytes) - n.nbytes);
                }
                else emit(nullptr, round8(n.nbytes));        // a hole: its owner stores the bytes (role root)
                return;
            }
            bytes_t b;
            // object header: version 1, one message (symbol table), reference count 1
            put(b, 1, 1); put(b, 0, 1); put(b, 1, 2); put(b, 1, 4); put(b, 24, 4); put(b, 0, 4);
            bytes_t st; put(st, n.root_tree_addr, 8); put(st, n.heap_addr, 8);
            put_message(b, 0x0011, 0, st);
            // local heap: no free block (free-list head = H5HL_FREE_NULL = 1)
            put_bytes(b, "HEAP", 4); put(b, 0, 4); put(b, n.heap_bytes, 8); put(b, 1, 8); put(b, n.heap_data_addr, 8);
            put(b, 0, 8);
            for (auto& c : n.children) { put_bytes(b, c.first.c_str(), c.first.size() + 1); pad8(b); }

            // B-tree: last name (heap offset) under each symbol-table node, then under each tree node
            const std::size_t count = n.children.size(), per = 2 * LEAF_K, fan = 2 * INTERNAL_K;
            std::vector<std::uint64_t> child_addr = n.snod_addr, last_key;
            for (std::size_t k = 0; k < n.snod_addr.size(); ++k) last_key.push_back(n.name_offset[std::min(count, (k + 1) * per) - 1]);
            std::vector<bytes_t> levels;
            for (std::size_t level = 0; level < n.tree_addr.size(); ++level)
            {
                bytes_t lv;
                std::vector<std::uint64_t> next_last;
                const auto& addr = n.tree_addr[level];
                for (std::size_t m = 0; m < addr.size(); ++m)
                {
                    std::size_t c0 = m * fan, c1 = std::min(child_addr.size(), c0 + fan);
                    bytes_t node;
                    put_bytes(node, "TREE", 4); put(node, 0, 1); put(node, level, 1); put(node, c1 - c0, 2);
                    put(node, m ? addr[m - 1] : UNDEF, 8); put(node, m + 1 < addr.size() ? addr[m + 1] : UNDEF, 8);
                    put(node, c0 ? last_key[c0 - 1] : 0, 8);
                    for (std::size_t c = c0; c < c1; ++c) { put(node, child_addr[c], 8); put(node, last_key[c], 8); }
                    node.resize(TREE_BYTES, 0);
                    put_bytes(lv, node.data(), node.size());
                    next_last.push_back(c1 > c0 ? last_key[c1 - 1] : 0);
                }
                levels.push_back(std::move(lv));
                child_addr = addr;
                last_key = next_last;
            }
            for (std::size_t level = levels.size(); level-- > 0; ) put_bytes(b, levels[level].data(), levels[level].size());

            // symbol-table nodes
            auto it = n.children.begin();
            for (std::size_t k = 0; k < n.snod_addr.size(); ++k)
            {
                bytes_t node;
                std::size_t e0 = k * per, e1 = std::min(count, e0 + per);
                put_bytes(node, "SNOD", 4); put(node, 1, 1); put(node, 0, 1); put(node, e1 - e0, 2);
                for (std::size_t e = e0; e < e1; ++e, ++it)
                {
                    const node_t& c = *it->second;
                    put(node, n.name_offset[e], 8); put(node, c.header_addr, 8);
                    put(node, c.is_group ? 1 : 0, 4); put(node, 0, 4);
                    put(node, c.is_group ? c.root_tree_addr : 0, 8); put(node, c.is_group ? c.heap_addr : 0, 8);
                }
                node.resize(SNOD_BYTES, 0);
                put_bytes(b, node.data(), node.size());
            }
            emit(b.data(), b.size());
            for (auto& c : n.children) write(*c.second, emit, written);
        }
    };
    std::function<void(const void*, std::size_t)> emit_fn = emit;
    stream_t::write(*root, emit_fn, written);
    if (written != eof) { std::fclose(f); throw std::logic_error("h5lite: end-of-file address mismatch"); }
    if (std::fflush(f) != 0 || ::ftruncate(fileno(f), off_t(eof)) != 0) { std::fclose(f); throw std::runtime_error("h5lite: sizing " + filename + " failed"); }
    if (std::fclose(f) != 0) throw std::runtime_error("h5lite: closing " + filename + " failed");
}




// ============================================================================
struct reader_t::impl_t
{
    std::FILE* f = nullptr;
    std::uint64_t base = 0, root_header = 0, file_bytes = 0;

    struct message_t { int type; bytes_t data; };
    struct dataset_t { type_t type; std::vector<std::uint64_t> shape; std::uint64_t address = UNDEF, nbytes = 0; bytes_t compact; bool is_compact = false; };

    bytes_t at(std::uint64_t addr, std::size_t n) const
    {
        bytes_t b(n);
        if (addr == UNDEF || base + addr + n > file_bytes) throw std::runtime_error("h5lite: read past the end of the file");
        if (std::fseek(f, long(base + addr), SEEK_SET) != 0 || (n && std::fread(b.data(), 1, n, f) != n)) throw std::runtime_error("h5lite: read failed");
        return b;
    }
    static std::uint64_t get(const bytes_t& b, std::size_t p, int n)
    {
        if (p + n > b.size()) throw std::runtime_error("h5lite: truncated structure");
        std::uint64_t v = 0;
        for (int k = 0; k < n; ++k) v |= std::uint64_t(b[p + k]) << (8 * k);
        return v;
    }

    std::vector<message_t> messages(std::uint64_t addr) const
    {
        auto h = at(addr, 16);
        if (h[0] != 1) throw std::runtime_error("h5lite: only version-1 object headers are supported");
        std::size_t total = get(h, 2, 2);
        std::vector<std::pair<std::uint64_t, std::uint64_t>> blocks = {{addr + 16, get(h, 8, 4)}};
        std::vector<message_t> out;
        for (std::size_t k = 0; k < blocks.size() && out.size() < total; ++k)
        {
            auto block = at(blocks[k].first, blocks[k].second);
            std::size_t p = 0;
            while (p + 8 <= block.size() && out.size() < total)
            {
                int type = int(get(block, p, 2));
                std::size_t size = get(block, p + 2, 2);
                if (p + 8 + size > block.size()) throw std::runtime_error("h5lite: message overruns its block");
                message_t m{type, bytes_t(block.begin() + p + 8, block.begin() + p + 8 + size)};
                if (type == 0x0010) blocks.push_back({get(m.data, 0, 8), get(m.data, 8, 8)});
                out.push_back(std::move(m));
                p += 8 + size;
            }
        }
        return out;
    }

    // One read per local heap and one walk per group: a restart file holds one dataset per leaf (65 536 at config 5), and every
    // rank looks all of its blocks up -- without these two caches each lookup re-read the group's whole symbol table.
    mutable std::map<std::uint64_t, bytes_t> heap_cache;                                        // heap address -> its data segment
    mutable std::map<std::uint64_t, std::map<std::string, std::uint64_t>> children_cache;       // group header -> name -> object header
    mutable std::map<std::uint64_t, std::vector<std::string>> order_cache;                      // group header -> names in file order

    std::string heap_name(std::uint64_t heap, std::uint64_t offset) const
    {
        auto it = heap_cache.find(heap);
        if (it == heap_cache.end())
        {
            auto h = at(heap, 32);
            if (std::memcmp(h.data(), "HEAP", 4)) throw std::runtime_error("h5lite: bad local heap");
            std::uint64_t size = get(h, 8, 8), data = get(h, 24, 8);
            it = heap_cache.emplace(heap, at(data, size)).first;
        }
        const bytes_t& seg = it->second;
        if (offset >= seg.size()) throw std::runtime_error("h5lite: name offset outside the heap");
        const char* p = reinterpret_cast<const char*>(seg.data()) + offset;
        return std::string(p, strnlen(p, seg.size() - offset));
    }

    void walk(std::uint64_t addr, std::uint64_t heap, std::vector<std::pair<std::string, std::uint64_t>>& out) const
    {
        auto head = at(addr, 8);
        if (! std::memcmp(head.data(), "SNOD", 4))
        {
            std::size_t count = get(head, 6, 2);
            auto body = at(addr + 8, count * 40);
            for (std::size_t k = 0; k < count; ++k) out.push_back({heap_name(heap, get(body, 40 * k, 8)), get(body, 40 * k + 8, 8)});
            return;
        }
        if (std::memcmp(head.data(), "TREE", 4)) throw std::runtime_error("h5lite: bad B-tree node");
        std::size_t used = get(head, 6, 2);
        auto body = at(addr + 24, (2 * used + 1) * 8);
        for (std::size_t k = 0; k < used; ++k) walk(get(body, 16 * k + 8, 8), heap, out);
    }

    /** name -> object header of a group's members (nullptr: not a group); built once per group */
    const std::map<std::string, std::uint64_t>* members(std::uint64_t header) const
    {
        auto it = children_cache.find(header);
        if (it != children_cache.end()) return &it->second;
        for (auto& m : messages(header))
            if (m.type == 0x0011)
            {
                std::vector<std::pair<std::string, std::uint64_t>> out;
                walk(get(m.data, 0, 8), get(m.data, 8, 8), out);
                auto& names = order_cache[header];
                auto& table = children_cache[header];
                for (auto& c : out) { names.push_back(c.first); table[c.first] = c.second; }
                return &table;
            }
        return nullptr;
    }

    std::vector<std::pair<std::string, std::uint64_t>> children(std::uint64_t header) const
    {
        const auto* table = members(header);
        if (! table) throw std::runtime_error("h5lite: not a group");
        std::vector<std::pair<std::string, std::uint64_t>> out;
        for (auto& name : order_cache[header]) out.push_back({name, table->at(name)});
        return out;
    }

    bool find(const std::string& path, std::uint64_t& header) const
    {
        header = root_header;
        std::size_t p = 0;
        while (p < path.size())
        {
            auto q = path.find('/', p);
            if (q == std::string::npos) q = path.size();
            if (q > p)
            {
                auto part = path.substr(p, q - p);
                const auto* table = members(header);
                if (! table) return false;
                auto hit = table->find(part);
                if (hit == table->end()) return false;
                header = hit->second;
            }
            p = q + 1;
        }
        return true;
    }

    static type_t parse_type(const bytes_t& d, std::size_t& p)
    {
        int cls = d.at(p) & 0x0f, version = d.at(p) >> 4;
        std::uint32_t b0 = d.at(p + 1), b1 = d.at(p + 2), size = std::uint32_t(get(d, p + 4, 4));
        p += 8;
        if (cls == 0)
        {
            if ((b0 & 1) || size != 4 || ! (b0 & 0x08)) throw std::runtime_error("h5lite: only little-endian signed 32-bit integers are supported");
            p += 4;
            return type_t::i32();
        }
        if (cls == 1)
        {
            if ((b0 & 1) || size != 8) throw std::runtime_error("h5lite: only little-endian 64-bit floats are supported");
            p += 12;
            return type_t::f64();
        }
        if (cls == 3) return type_t::string(size);
        if (cls == 10)
        {
            std::size_t rank = d.at(p);
            if (rank != 1) throw std::runtime_error("h5lite: only one-dimensional array types are supported");
            std::uint32_t n;
            if (version == 2) { n = std::uint32_t(get(d, p + 4, 4)); p += 4 + 8; }
            else if (version == 3) { n = std::uint32_t(get(d, p + 1, 4)); p += 5; }
            else throw std::runtime_error("h5lite: array datatype version");
            auto base = parse_type(d, p);
            return type_t::array(base, n);
        }
        if (cls == 6)
        {
            std::size_t count = b0 | (b1 << 8);
            std::vector<type_t::member_t> members;
            for (std::size_t k = 0; k < count; ++k)
            {
                std::size_t end = p;
                while (d.at(end)) ++end;
                std::string name(d.begin() + p, d.begin() + end);
                p = version < 3 ? p + (end - p + 8) / 8 * 8 : end + 1;
                std::uint32_t offset;
                if (version == 1) { offset = std::uint32_t(get(d, p, 4)); p += 4 + 1 + 3 + 4 + 4 + 16; }
                else if (version == 2) { offset = std::uint32_t(get(d, p, 4)); p += 4; }
                else { int nb = size < 256 ? 1 : (size < 65536 ? 2 : 4); offset = std::uint32_t(get(d, p, nb)); p += nb; }
                auto t = parse_type(d, p);
                members.push_back(type_t::member(name, offset, t));
            }
            return type_t::compound(size, std::move(members));
        }
        throw std::runtime_error("h5lite: datatype class " + std::to_string(cls) + " is not supported");
    }

    dataset_t dataset(const std::string& path) const
    {
        std::uint64_t header;
        if (! find(path, header)) throw std::runtime_error("h5lite: no object " + path);
        dataset_t ds;
        bool have_space = false, have_type = false, have_layout = false;
        for (auto& m : messages(header))
        {
            if (m.type == 0x0001)
            {
                int version = m.data.at(0);
                std::size_t rank = m.data.at(1), start = version == 1 ? 8 : 4;
                for (std::size_t k = 0; k < rank; ++k) ds.shape.push_back(get(m.data, start + 8 * k, 8));
                have_space = true;
            }
            else if (m.type == 0x0003) { std::size_t p = 0; ds.type = parse_type(m.data, p); have_type = true; }
            else if (m.type == 0x0008)
            {
                int version = m.data.at(0);
                if (version == 3 && m.data.at(1) == 1) { ds.address = get(m.data, 2, 8); ds.nbytes = get(m.data, 10, 8); }
                else if (version == 3 && m.data.at(1) == 0) { std::size_t n = get(m.data, 2, 2); ds.compact.assign(m.data.begin() + 4, m.data.begin() + 4 + n); ds.is_compact = true; }
                else if ((version == 1 || version == 2) && m.data.at(2) == 1) { ds.address = get(m.data, 8, 8); ds.nbytes = UNDEF; }
                else throw std::runtime_error("h5lite: " + path + " has a dataset layout that is not supported (chunked?)");
                have_layout = true;
            }
        }
        if (! (have_space && have_type && have_layout)) throw std::runtime_error("h5lite: " + path + " is not a dataset");
        return ds;
    }
};

reader_t::reader_t(const std::string& filename) : impl(new impl_t)
{
    impl->f = std::fopen(filename.c_str(), "rb");
    if (! impl->f) throw std::runtime_error("h5lite: cannot open " + filename);
    std::fseek(impl->f, 0, SEEK_END);
    impl->file_bytes = std::uint64_t(std::ftell(impl->f));
    bool found = false;
    for (std::uint64_t off = 0; off + 96 <= impl->file_bytes; off = off ? off * 2 : 512)
    {
        impl->base = 0;
        auto sig = impl->at(off, 8);
        if (! std::memcmp(sig.data(), SIGNATURE, 8)) { impl->base = off; found = true; break; }
    }
    if (! found) throw std::runtime_error("h5lite: " + filename + " is not an HDF5 file");
    auto sb = impl->at(0, 96);
    if (sb[8] != 0) throw std::runtime_error("h5lite: superblock version " + std::to_string(sb[8]) + " is not supported (file written with libver=latest?)");
    if (sb[13] != 8 || sb[14] != 8) throw std::runtime_error("h5lite: only 8-byte offsets and lengths are supported");
    impl->root_header = impl_t::get(sb, 64, 8);
}

reader_t::~reader_t() { if (impl->f) std::fclose(impl->f); }

bool reader_t::exists(const std::string& path) const { std::uint64_t h; return impl->find(path, h); }

bool reader_t::is_group(const std::string& path) const
{
    std::uint64_t h;
    if (! impl->find(path, h)) return false;
    for (auto& m : impl->messages(h)) if (m.type == 0x0011) return true;
    return false;
}

std::vector<std::string> reader_t::keys(const std::string& group_path) const
{
    std::uint64_t h;
    if (! impl->find(group_path, h)) throw std::runtime_error("h5lite: no group " + group_path);
    std::vector<std::string> out;
    for (auto& c : impl->children(h)) out.push_back(c.first);
    return out;
}

type_t reader_t::type(const std::string& path) const { return impl->dataset(path).type; }
std::vector<std::uint64_t> reader_t::shape(const std::string& path) const { return impl->dataset(path).shape; }

std::vector<unsigned char> reader_t::read(const std::string& path, const type_t& expected) const
{
    auto ds = impl->dataset(path);
    if (ds.type != expected) throw std::runtime_error("h5lite: " + path + " has type " + ds.type.describe() + ", expected " + expected.describe());
    std::uint64_t n = ds.type.size;
    for (auto d : ds.shape) n *= d;
    if (ds.is_compact) { ds.compact.resize(n); return ds.compact; }
    if (n == 0) return {};
    return impl->at(ds.address, n);
}

double reader_t::read_double(const std::string& path) const { auto b = read(path, type_t::f64()); double v; std::memcpy(&v, b.data(), 8); return v; }
int reader_t::read_int(const std::string& path) const { auto b = read(path, type_t::i32()); int v; std::memcpy(&v, b.data(), 4); return v; }

std::string reader_t::read_string(const std::string& path) const
{
    auto t = type(path);
    if (t.kind != type_t::kind_t::string) throw std::runtime_error("h5lite: " + path + " is not a string");
    auto b = read(path, t);
    return std::string(reinterpret_cast<const char*>(b.data()), strnlen(reinterpret_cast<const char*>(b.data()), b.size()));
}
