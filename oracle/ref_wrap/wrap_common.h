// oracle/ref_wrap/wrap_common.h -- TEST INFRASTRUCTURE ONLY.
// Helpers to build the shim's nb::ndarray views from raw C pointers.  Each
// wrap_*.cpp #includes ONE reference translation unit verbatim from
// /root/reference/src (never copied into this repo) and re-exports its entry
// points with a C ABI so tests can reach them through ctypes.
#pragma once
#include <nanobind/nanobind.h>
#include <cstddef>
namespace nb = nanobind;
#define A1D(p, n) nb::ndarray<double, nb::ro, nb::ndim<1>>((const double *)(p), {(size_t)(n)})
#define A2D(p, n, m) nb::ndarray<double, nb::ro, nb::ndim<2>>((const double *)(p), {(size_t)(n), (size_t)(m)})
#define A3D(p, n, m, k) nb::ndarray<double, nb::ro, nb::ndim<3>>((const double *)(p), {(size_t)(n), (size_t)(m), (size_t)(k)})
#define A1I(p, n) nb::ndarray<int, nb::ro, nb::ndim<1>>((const int *)(p), {(size_t)(n)})
#define A2I(p, n, m) nb::ndarray<int, nb::ro, nb::ndim<2>>((const int *)(p), {(size_t)(n), (size_t)(m)})
#define W1D(p, n) nb::ndarray<double, nb::ndim<1>>((double *)(p), {(size_t)(n)})
#define W2D(p, n, m) nb::ndarray<double, nb::ndim<2>>((double *)(p), {(size_t)(n), (size_t)(m)})
#define W3D(p, n, m, k) nb::ndarray<double, nb::ndim<3>>((double *)(p), {(size_t)(n), (size_t)(m), (size_t)(k)})
#define W1I(p, n) nb::ndarray<int, nb::ndim<1>>((int *)(p), {(size_t)(n)})
#define W2I(p, n, m) nb::ndarray<int, nb::ndim<2>>((int *)(p), {(size_t)(n), (size_t)(m)})
#define BOXARGS const double *box9, const double *origin3, const int *boundary3
#define BOXPASS A2D(box9, 3, 3), A1D(origin3, 3), A1I(boundary3, 3)
