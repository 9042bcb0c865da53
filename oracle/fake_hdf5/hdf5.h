// Stub of <hdf5.h> for oracle/_ref (test infrastructure).  libhdf5 is not in this image, but 23 OpenMC
// headers include it.  With these declarations the reference's translation units PARSE; oracle/Makefile then
// links single functions out of them (-ffunction-sections + --gc-sections, hidden visibility), so no HDF5
// function is ever referenced, let alone called.  Nothing here restates anything of HDF5's behaviour.
#pragma once
#include <cstddef>
#include <cstdint>
typedef int64_t hid_t;
typedef int herr_t;
typedef int htri_t;
typedef unsigned long long hsize_t;
typedef long long hssize_t;
#define H5T_NATIVE_DOUBLE ((hid_t)1)
#define H5T_NATIVE_INT ((hid_t)2)
#define H5T_NATIVE_INT64 ((hid_t)3)
#define H5T_NATIVE_UINT64 ((hid_t)4)
#define H5T_NATIVE_INT8 ((hid_t)5)
#define H5T_NATIVE_UINT ((hid_t)6)
#define H5T_NATIVE_ULONG ((hid_t)7)
#define H5T_NATIVE_LONG ((hid_t)8)
#define H5T_NATIVE_FLOAT ((hid_t)9)
#define H5T_NATIVE_CHAR ((hid_t)10)
#define H5T_NATIVE_HBOOL ((hid_t)11)
#define H5T_C_S1 ((hid_t)12)
#define H5S_ALL ((hid_t)0)
#define H5P_DEFAULT ((hid_t)0)
