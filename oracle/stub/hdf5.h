/*
 * Declaration-only stand-in for <hdf5.h>, used ONLY to compile the reference's
 * own translation units (in place, from /root/reference/src) into the parity
 * oracle under oracle/_ref/.  libhdf5 is not installed in this image.  The
 * reference's `binary` hot path (advance / maximum_timestep / create_solver_data)
 * never reaches an HDF5 call unless restart= is given or an output task runs,
 * so every body below aborts.  This file is test infrastructure; nothing under
 * mara3_b200/ includes it.
 *
 * The symbol set is the one used by the reference's h5:: wrapper
 * (reference: src/core_hdf5.hpp:72-990).
 */
#ifndef M3B_ORACLE_HDF5_STUB_H
#define M3B_ORACLE_HDF5_STUB_H
#include <stdlib.h>
#include <stddef.h>
#include <stdint.h>

typedef int64_t hid_t;
typedef int herr_t;
typedef int htri_t;
typedef unsigned long long hsize_t;
typedef long long hssize_t;

typedef struct { const char* desc; } H5E_error2_t;
typedef enum { H5O_TYPE_UNKNOWN = -1, H5O_TYPE_GROUP, H5O_TYPE_DATASET } H5O_type_t;
typedef struct { H5O_type_t type; } H5O_info_t;
typedef struct { int unused; } H5L_info_t;
typedef enum { H5E_WALK_UPWARD = 0 } H5E_direction_t;
typedef enum { H5S_SCALAR = 0, H5S_SIMPLE = 1, H5S_NULL = 2 } H5S_class_t;
typedef enum { H5S_SELECT_SET = 0 } H5S_seloper_t;
typedef enum { H5T_COMPOUND = 6 } H5T_class_t;
typedef enum { H5_INDEX_NAME = 0 } H5_index_t;
typedef enum { H5_ITER_INC = 0, H5_ITER_NATIVE = 2 } H5_iter_order_t;
typedef herr_t (*H5E_walk2_t)(unsigned, const H5E_error2_t*, void*);
typedef herr_t (*H5L_iterate_t)(hid_t, const char*, const H5L_info_t*, void*);

#define H5P_DEFAULT 0
#define H5E_DEFAULT 0
#define H5S_UNLIMITED ((hsize_t)(-1))
#define H5F_ACC_RDONLY 0u
#define H5F_ACC_RDWR 1u
#define H5F_ACC_TRUNC 2u
#define H5P_DATASET_CREATE 1
#define H5T_NATIVE_DOUBLE 1
#define H5T_NATIVE_INT 2
#define H5T_NATIVE_ULONG 3
#define H5T_C_S1 4

#define M3B_STUB(ret, name, args) static inline ret name args { abort(); }
M3B_STUB(hid_t,   H5Eget_current_stack, (void))
M3B_STUB(herr_t,  H5Ewalk, (hid_t a, H5E_direction_t b, H5E_walk2_t c, void* d))
M3B_STUB(herr_t,  H5Eclear, (hid_t a))
M3B_STUB(herr_t,  H5Eclose_stack, (hid_t a))
M3B_STUB(hid_t,   H5Pcreate, (hid_t a))
M3B_STUB(hid_t,   H5Pcopy, (hid_t a))
M3B_STUB(herr_t,  H5Pclose, (hid_t a))
M3B_STUB(herr_t,  H5Pset_chunk, (hid_t a, int b, const hsize_t* c))
M3B_STUB(htri_t,  H5Pequal, (hid_t a, hid_t b))
M3B_STUB(hid_t,   H5Tcopy, (hid_t a))
M3B_STUB(hid_t,   H5Tcreate, (H5T_class_t a, size_t b))
M3B_STUB(herr_t,  H5Tinsert, (hid_t a, const char* b, size_t c, hid_t d))
M3B_STUB(herr_t,  H5Tclose, (hid_t a))
M3B_STUB(htri_t,  H5Tequal, (hid_t a, hid_t b))
M3B_STUB(size_t,  H5Tget_size, (hid_t a))
M3B_STUB(herr_t,  H5Tset_size, (hid_t a, size_t b))
M3B_STUB(hid_t,   H5Tarray_create, (hid_t a, unsigned b, const hsize_t* c))
M3B_STUB(hid_t,   H5Screate, (H5S_class_t a))
M3B_STUB(hid_t,   H5Screate_simple, (int a, const hsize_t* b, const hsize_t* c))
M3B_STUB(hid_t,   H5Scopy, (hid_t a))
M3B_STUB(htri_t,  H5Sextent_equal, (hid_t a, hid_t b))
M3B_STUB(herr_t,  H5Sclose, (hid_t a))
M3B_STUB(int,     H5Sget_simple_extent_ndims, (hid_t a))
M3B_STUB(hssize_t,H5Sget_simple_extent_npoints, (hid_t a))
M3B_STUB(int,     H5Sget_simple_extent_dims, (hid_t a, hsize_t* b, hsize_t* c))
M3B_STUB(hssize_t,H5Sget_select_npoints, (hid_t a))
M3B_STUB(herr_t,  H5Sget_select_bounds, (hid_t a, hsize_t* b, hsize_t* c))
M3B_STUB(herr_t,  H5Sselect_all, (hid_t a))
M3B_STUB(herr_t,  H5Sselect_none, (hid_t a))
M3B_STUB(herr_t,  H5Sselect_hyperslab, (hid_t a, H5S_seloper_t b, const hsize_t* c, const hsize_t* d, const hsize_t* e, const hsize_t* f))
M3B_STUB(herr_t,  H5Fclose, (hid_t a))
M3B_STUB(herr_t,  H5Gclose, (hid_t a))
M3B_STUB(herr_t,  H5Dclose, (hid_t a))
M3B_STUB(herr_t,  H5Literate, (hid_t a, H5_index_t b, H5_iter_order_t c, hsize_t* d, H5L_iterate_t e, void* f))
M3B_STUB(htri_t,  H5Lexists, (hid_t a, const char* b, hid_t c))
M3B_STUB(herr_t,  H5Oget_info_by_name, (hid_t a, const char* b, H5O_info_t* c, hid_t d))
M3B_STUB(hid_t,   H5Gopen, (hid_t a, const char* b, hid_t c))
M3B_STUB(hid_t,   H5Gcreate, (hid_t a, const char* b, hid_t c, hid_t d, hid_t e))
M3B_STUB(hid_t,   H5Dopen, (hid_t a, const char* b, hid_t c))
M3B_STUB(hid_t,   H5Dcreate, (hid_t a, const char* b, hid_t c, hid_t d, hid_t e, hid_t f, hid_t g))
M3B_STUB(long,    H5Lget_name_by_idx, (hid_t a, const char* b, H5_index_t c, H5_iter_order_t d, hsize_t e, char* f, size_t g, hid_t h))
M3B_STUB(hid_t,   H5Dget_space, (hid_t a))
M3B_STUB(hid_t,   H5Dget_type, (hid_t a))
M3B_STUB(herr_t,  H5Dwrite, (hid_t a, hid_t b, hid_t c, hid_t d, hid_t e, const void* f))
M3B_STUB(herr_t,  H5Dread, (hid_t a, hid_t b, hid_t c, hid_t d, hid_t e, void* f))
M3B_STUB(hid_t,   H5Dget_create_plist, (hid_t a))
M3B_STUB(herr_t,  H5Dset_extent, (hid_t a, const hsize_t* b))
M3B_STUB(htri_t,  H5Fis_hdf5, (const char* a))
M3B_STUB(hid_t,   H5Fopen, (const char* a, unsigned b, hid_t c))
M3B_STUB(hid_t,   H5Fcreate, (const char* a, unsigned b, hid_t c, hid_t d))
M3B_STUB(herr_t,  H5Fget_intent, (hid_t a, unsigned* b))
M3B_STUB(long,    H5Fget_name, (hid_t a, char* b, size_t c))
#undef M3B_STUB
#endif
