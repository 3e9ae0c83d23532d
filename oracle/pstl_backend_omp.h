// oracle/pstl_backend_omp.h — TEST/BASELINE INFRASTRUCTURE ONLY (never linked into libnbx.so).
//
// A minimal OpenMP parallel backend for libstdc++'s PSTL so that the UNMODIFIED reference (`/root/reference/src/main.cpp`)
// can run its `std::execution::par / par_unseq` algorithms on all host cores in an image that has no TBB: libstdc++ 13
// ships only the TBB and the serial backend (pstl/parallel_backend.h), and without <tbb/tbb.h> it silently selects the
// serial one (bits/c++config.h:869-876) — the "1 core" baseline of round 1. oracle/Makefile copies the system's pstl/
// headers into oracle/_ref/pstl_shim/pstl/, applies the reference's own one-line patch (scripts/patch_pstl.sh, the C++23
// P2408 back-port that lets views::iota iterators take the parallel path) and replaces pstl/parallel_backend.h by a
// stub that includes this file. Nothing of the reference is copied or modified.
//
// Only what the reference uses runs in parallel — for_each / for_each_n (-> __parallel_for), transform_reduce
// (-> __parallel_transform_reduce) and sort (-> __parallel_stable_sort); the remaining entry points forward to the serial
// backend.
#ifndef NBX_PSTL_BACKEND_OMP_H
#define NBX_PSTL_BACKEND_OMP_H

#include <omp.h>

#include <algorithm>
#include <cstddef>
#include <iterator>
#include <type_traits>
#include <vector>

#include "parallel_backend_serial.h"  // resolved inside the shim copy of pstl/ (this file is installed there)

namespace __pstl {
namespace __omp_backend {

using __serial_backend::__buffer;
using __serial_backend::__cancel_execution;
using __serial_backend::__parallel_invoke;
using __serial_backend::__parallel_merge;
using __serial_backend::__parallel_reduce;
using __serial_backend::__parallel_strict_scan;
using __serial_backend::__parallel_transform_scan;

// first + k for integral indices and random-access iterators alike
template <class _Index>
inline _Index __adv(_Index __first, std::size_t __k) {
  if constexpr (std::is_integral_v<_Index>)
    return _Index(__first + _Index(__k));
  else
    return __first + typename std::iterator_traits<_Index>::difference_type(__k);
}

// number of chunks a range of n items is cut into: a few per thread so that uneven work (tree walks) balances
inline std::size_t __chunks_for(std::size_t __n) {
  const std::size_t __t = std::size_t(omp_get_max_threads());
  return std::max<std::size_t>(1, std::min<std::size_t>(__n, __t * 16));
}

template <class _ExecutionPolicy, class _Index, class _Fp>
void __parallel_for(_ExecutionPolicy&&, _Index __first, _Index __last, _Fp __f) {
  const std::size_t __n = std::size_t(__last - __first);
  if (__n == 0) return;
  const std::size_t __c = __chunks_for(__n);
  if (__c == 1 || omp_in_parallel()) {
    __f(__first, __last);
    return;
  }
#pragma omp parallel for schedule(dynamic, 1)
  for (std::size_t __k = 0; __k < __c; ++__k) {
    const std::size_t __b = __n * __k / __c, __e = __n * (__k + 1) / __c;
    if (__b < __e) __f(__adv(__first, __b), __adv(__first, __e));
  }
}

template <class _ExecutionPolicy, class _Index, class _UnaryOp, class _Tp, class _BinaryOp, class _Reduce>
_Tp __parallel_transform_reduce(_ExecutionPolicy&&, _Index __first, _Index __last, _UnaryOp __u, _Tp __init, _BinaryOp __combine,
                                _Reduce __reduce) {
  const std::size_t __n = std::size_t(__last - __first);
  const std::size_t __c = __chunks_for(__n);
  if (__n < 2 || __c == 1 || omp_in_parallel()) return __reduce(__first, __last, __init);
  // every chunk folds its own items starting from its first transformed item (no identity element is known),
  // the partial results are then combined with `init` in chunk order
  std::vector<_Tp> __part(__c, __init);
#pragma omp parallel for schedule(dynamic, 1)
  for (std::size_t __k = 0; __k < __c; ++__k) {
    const std::size_t __b = __n * __k / __c, __e = __n * (__k + 1) / __c;  // n >= c => b < e
    __part[__k] = __reduce(__adv(__first, __b + 1), __adv(__first, __e), _Tp(__u(__adv(__first, __b))));
  }
  _Tp __r = __init;
  for (std::size_t __k = 0; __k < __c; ++__k) __r = __combine(__r, __part[__k]);
  return __r;
}

template <class _ExecutionPolicy, typename _RandomAccessIterator, typename _Compare, typename _LeafSort>
void __parallel_stable_sort(_ExecutionPolicy&&, _RandomAccessIterator __first, _RandomAccessIterator __last, _Compare __comp,
                            _LeafSort __leaf_sort, std::size_t = 0) {
  const std::size_t __n = std::size_t(__last - __first);
  std::size_t __c = 1;
  while (__c < std::size_t(omp_get_max_threads())) __c <<= 1;  // power of two => a clean merge tree
  if (__n < 4096 || __c == 1 || omp_in_parallel()) {
    __leaf_sort(__first, __last, __comp);
    return;
  }
#pragma omp parallel for schedule(dynamic, 1)
  for (std::size_t __k = 0; __k < __c; ++__k) __leaf_sort(__adv(__first, __n * __k / __c), __adv(__first, __n * (__k + 1) / __c), __comp);
  for (std::size_t __w = 1; __w < __c; __w <<= 1) {  // merge runs of w chunks pairwise (std::inplace_merge is stable)
#pragma omp parallel for schedule(dynamic, 1)
    for (std::size_t __k = 0; __k < __c; __k += 2 * __w)
      std::inplace_merge(__adv(__first, __n * __k / __c), __adv(__first, __n * (__k + __w) / __c), __adv(__first, __n * (__k + 2 * __w) / __c), __comp);
  }
}

}  // namespace __omp_backend
namespace __par_backend = __omp_backend;
}  // namespace __pstl

#endif
