/**
 * h5lite.hpp -- the subset of the HDF5 file format that the reference's products use, written and read
 * without libhdf5 (which this image does not have).
 *
 * The reference writes its checkpoints / diagnostics through a thin RAII wrapper over the HDF5 C API
 * (Mara3 src/core_hdf5.hpp:48-57, 474-700) with default property lists, i.e. what libhdf5 calls the
 * "earliest" file format: version-0 superblock, version-1 object headers, groups as symbol tables
 * (v1 B-tree + local heap + symbol-table nodes), contiguous dataset layout, and datatypes of class
 * fixed-point, floating-point, string, array and compound.  This module emits exactly those structures
 * (HDF5 File Format Specification, "Disk Format: Level 0 / 1 / 2"), so h5py / h5dump / the reference's own
 * readers open the files, and parses them back for `restart=`.
 *
 * Checked in tests/test_h5lite.py with an independent Python parser that is itself pinned against a file
 * written by the real library.
 */
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace m3b { namespace h5 {

    /** A datatype: H5T_NATIVE_DOUBLE / H5T_NATIVE_INT / H5T_C_S1 with a size / H5Tarray_create / H5Tcreate(H5T_COMPOUND). */
    struct type_t
    {
        enum class kind_t { f64, i32, string, array, compound };
        struct member_t { std::string name; std::uint32_t offset; std::shared_ptr<type_t> type; };

        kind_t kind = kind_t::f64;
        std::uint32_t size = 8;                 // bytes of one element
        std::vector<std::uint32_t> dims;        // array
        std::shared_ptr<type_t> base;           // array
        std::vector<member_t> members;          // compound

        static type_t f64();
        static type_t i32();
        static type_t string(std::size_t length);                   // fixed length, null-terminated padding (H5T_C_S1)
        static type_t array(const type_t& base, std::uint32_t n);   // one-dimensional array type
        static type_t compound(std::uint32_t size, std::vector<member_t> members);
        static member_t member(std::string name, std::size_t offset, const type_t& type);
        bool operator==(const type_t& other) const;
        bool operator!=(const type_t& other) const { return ! (*this == other); }
        std::string describe() const;
    };

    /**
     * Collects groups and datasets, then lays the file out in one pass on close().  Dataset bytes are NOT copied:
     * the caller keeps them alive until close() (or passes `copy = true` for temporaries).
     */
    class writer_t
    {
    public:
        /**
         * One file written by several processes (one per GPU), each the bytes of its own datasets:
         *   whole  one process writes everything (the default);
         *   root   writes the file's structure and whatever datasets it holds; a dataset given as nullptr is left as a hole
         *          of the right size for its owner;
         *   part   describes the SAME groups and datasets in the same order (nullptr for what it does not own), computes the
         *          same layout, and on close() only stores its own datasets (`mine = true`) at their addresses in the file the
         *          root has created (call it after the root's close(); pwrite, no truncation).
         */
        enum class role_t { whole, root, part };
        explicit writer_t(std::string filename, role_t role = role_t::whole);
        ~writer_t();
        writer_t(const writer_t&) = delete;

        void require_group(const std::string& path);
        void write(const std::string& path, const type_t& type, const std::vector<std::uint64_t>& shape, const void* data, bool copy = false, bool mine = false);

        // conveniences mirroring h5::Group::write for the reference's scalar types
        void write_double(const std::string& path, double value);
        void write_int(const std::string& path, int value);
        void write_string(const std::string& path, const std::string& value);
        void close();

    private:
        struct node_t;
        node_t* descend(const std::string& path, bool create_last_as_group, std::string* leaf_name);
        std::string filename;
        std::unique_ptr<node_t> root;
        bool closed = false;
        role_t role = role_t::whole;
    };

    /** Read side for `restart=`: the same subset, from files written by this module or by libhdf5's defaults. */
    class reader_t
    {
    public:
        explicit reader_t(const std::string& filename);
        ~reader_t();
        reader_t(const reader_t&) = delete;

        bool exists(const std::string& path) const;
        bool is_group(const std::string& path) const;
        std::vector<std::string> keys(const std::string& group_path) const;     // in name order
        type_t type(const std::string& path) const;
        std::vector<std::uint64_t> shape(const std::string& path) const;
        /** Raw element bytes; throws unless the stored type equals `expected`. */
        std::vector<unsigned char> read(const std::string& path, const type_t& expected) const;
        double read_double(const std::string& path) const;
        int read_int(const std::string& path) const;
        std::string read_string(const std::string& path) const;

    private:
        struct impl_t;
        std::unique_ptr<impl_t> impl;
    };
}}
