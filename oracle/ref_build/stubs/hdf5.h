/* Declaration-only stand-in for <hdf5.h> (parallel build flag set).
 *
 * TEST INFRASTRUCTURE ONLY.  The reference's operator translation units include
 * its HDF5 wrapper header transitively; none of the hot-path arithmetic does I/O.
 * Every function here is an inert inline no-op so that those translation units
 * compile from where they lie under /root/reference.  Nothing is copied from it.
 */
#ifndef SB200_STUB_HDF5_H
#define SB200_STUB_HDF5_H
#define H5_HAVE_PARALLEL 1
typedef long hid_t;
typedef int herr_t;
typedef int htri_t;
typedef unsigned long long hsize_t;
typedef long long hssize_t;
typedef int H5D_layout_t;
typedef int H5FD_mpio_xfer_t;
typedef int H5Z_filter_t;
#define H5FD_MPIO_INDEPENDENT 0
#define H5FD_MPIO_COLLECTIVE 1
#define H5F_ACC_RDONLY 0
#define H5F_ACC_RDWR 1
#define H5F_ACC_TRUNC 2
#define H5F_ACC_EXCL 4
#define H5F_SCOPE_GLOBAL 1
#define H5F_SCOPE_LOCAL 0
#define H5P_DEFAULT 0
#define H5P_DATASET_ACCESS 1
#define H5P_DATASET_CREATE 2
#define H5P_DATASET_XFER 3
#define H5P_FILE_ACCESS 4
#define H5P_FILE_CREATE 5
#define H5P_LINK_ACCESS 6
#define H5P_LINK_CREATE 7
#define H5P_GROUP_CREATE 8
#define H5S_ALL 0
#define H5S_SCALAR 0
#define H5S_SELECT_SET 0
#define H5S_UNLIMITED ((hsize_t)(-1))
#define H5T_C_S1 1
#define H5T_NATIVE_DOUBLE 2
#define H5T_NATIVE_INT 3
#define H5T_NATIVE_SHORT 4
#define H5T_NATIVE_UINT 5
#define H5T_NATIVE_ULONG 6
#define H5T_NATIVE_UINT64 7
#define H5T_NATIVE_FLOAT 8
#define H5T_NATIVE_CHAR 9
#define H5T_NATIVE_LONG 10
#define H5T_NATIVE_USHORT 11
#define H5T_NATIVE_UINT32 12
#define H5T_VARIABLE ((size_t)(-1))
#define H5Z_FILTER_DEFLATE 1
#define H5D_CHUNKED 2
#define H5D_CONTIGUOUS 1
#define H5E_DEFAULT 0
#define H5_INDEX_NAME 0
#define H5_ITER_NATIVE 0
#ifdef __cplusplus
#define SB200_H5_NOOP(name) template<class... A> static inline long name( A... ) { return -1; }
SB200_H5_NOOP(H5Aclose) SB200_H5_NOOP(H5Acreate) SB200_H5_NOOP(H5Acreate2) SB200_H5_NOOP(H5Aexists) SB200_H5_NOOP(H5Aget_space)
SB200_H5_NOOP(H5Aget_type) SB200_H5_NOOP(H5Aopen) SB200_H5_NOOP(H5Aopen_name) SB200_H5_NOOP(H5Aread) SB200_H5_NOOP(H5Awrite)
SB200_H5_NOOP(H5Dclose) SB200_H5_NOOP(H5Dcreate) SB200_H5_NOOP(H5Dcreate2) SB200_H5_NOOP(H5Dget_space) SB200_H5_NOOP(H5Dopen)
SB200_H5_NOOP(H5Dopen2) SB200_H5_NOOP(H5Dread) SB200_H5_NOOP(H5Dset_extent) SB200_H5_NOOP(H5Dwrite) SB200_H5_NOOP(H5Dget_type)
SB200_H5_NOOP(H5Fflush) SB200_H5_NOOP(H5Fclose) SB200_H5_NOOP(H5Fcreate) SB200_H5_NOOP(H5Fopen)
SB200_H5_NOOP(H5Gcreate) SB200_H5_NOOP(H5Gcreate2) SB200_H5_NOOP(H5Gopen) SB200_H5_NOOP(H5Gopen2) SB200_H5_NOOP(H5Gclose)
SB200_H5_NOOP(H5Gget_num_objs) SB200_H5_NOOP(H5Gget_objname_by_idx) SB200_H5_NOOP(H5Gget_info)
SB200_H5_NOOP(H5Lcreate_soft) SB200_H5_NOOP(H5Lexists) SB200_H5_NOOP(H5Ldelete) SB200_H5_NOOP(H5Lget_name_by_idx)
SB200_H5_NOOP(H5Oopen) SB200_H5_NOOP(H5Oclose)
SB200_H5_NOOP(H5Pclose) SB200_H5_NOOP(H5Pcreate) SB200_H5_NOOP(H5Pget_dxpl_mpio) SB200_H5_NOOP(H5Pget_layout)
SB200_H5_NOOP(H5Premove_filter) SB200_H5_NOOP(H5Pset_chunk) SB200_H5_NOOP(H5Pset_deflate) SB200_H5_NOOP(H5Pset_dxpl_mpio)
SB200_H5_NOOP(H5Pset_layout) SB200_H5_NOOP(H5Pset_fapl_mpio) SB200_H5_NOOP(H5Pset_alloc_time) SB200_H5_NOOP(H5Pset_create_intermediate_group)
SB200_H5_NOOP(H5Pset_libver_bounds) SB200_H5_NOOP(H5Pset_fill_time)
SB200_H5_NOOP(H5Sclose) SB200_H5_NOOP(H5Screate) SB200_H5_NOOP(H5Screate_simple) SB200_H5_NOOP(H5Sget_simple_extent_dims)
SB200_H5_NOOP(H5Sget_simple_extent_ndims) SB200_H5_NOOP(H5Sget_simple_extent_npoints) SB200_H5_NOOP(H5Sselect_hyperslab)
SB200_H5_NOOP(H5Sselect_none) SB200_H5_NOOP(H5Sselect_elements)
SB200_H5_NOOP(H5Tclose) SB200_H5_NOOP(H5Tcopy) SB200_H5_NOOP(H5Tget_size) SB200_H5_NOOP(H5Tset_size) SB200_H5_NOOP(H5Tequal)
SB200_H5_NOOP(H5Eset_auto) SB200_H5_NOOP(H5Eset_auto2) SB200_H5_NOOP(H5open) SB200_H5_NOOP(H5close)
#endif
#endif
