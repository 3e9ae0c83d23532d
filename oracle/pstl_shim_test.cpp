// oracle/pstl_shim_test.cpp — TEST INFRASTRUCTURE: exercises the OpenMP PSTL backend (oracle/pstl_backend_omp.h) through the
// standard parallel algorithms the reference uses (for_each, for_each_n over an iota view, transform_reduce, sort) and checks
// every result against the sequential algorithm. Built and run by tests/test_pstl_shim.py with -isystem oracle/_ref/pstl_shim.
#include <omp.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <execution>
#include <numeric>
#include <random>
#include <ranges>
#include <tuple>
#include <vector>

int main() {
  int bad = 0;
  auto check = [&](bool ok, const char* what) {
    if (!ok) { std::printf("FAIL %s\n", what); ++bad; }
  };
  const std::size_t n = 1'000'003;
  std::mt19937_64 gen{7};
  std::vector<std::uint64_t> keys(n);
  for (auto& k : keys) k = gen() % 100000;  // many ties
  // sort of (key, index) pairs by key only, like src/bvh.h:62-69
  std::vector<std::pair<std::uint64_t, std::size_t>> a(n), b;
  for (std::size_t i = 0; i < n; ++i) a[i] = {keys[i], i};
  b = a;
  std::sort(std::execution::par_unseq, a.begin(), a.end(), [](auto x, auto y) { return x.first < y.first; });
  std::stable_sort(b.begin(), b.end(), [](auto x, auto y) { return x.first < y.first; });
  bool sorted = std::is_sorted(a.begin(), a.end(), [](auto x, auto y) { return x.first < y.first; });
  check(sorted, "sort: keys ascending");
  std::vector<std::size_t> ia(n), ib(n);
  for (std::size_t i = 0; i < n; ++i) { ia[i] = a[i].second; ib[i] = b[i].second; }
  std::sort(ia.begin(), ia.end());
  std::sort(ib.begin(), ib.end());
  check(ia == ib, "sort: a permutation of the input");
  // for_each_n over an iota view (the reference's counting ranges), every index exactly once
  std::vector<std::uint32_t> hits(n, 0);
  auto r = std::views::iota(std::size_t(0), n);
  std::for_each_n(std::execution::par_unseq, r.begin(), n, [h = hits.data()](auto i) { h[i] += 1; });
  check(std::all_of(hits.begin(), hits.end(), [](auto v) { return v == 1; }), "for_each_n: every index once");
  std::atomic<int> maxthreads{0};
  std::for_each(std::execution::par, r.begin(), r.end(), [&](auto) {
    int t = omp_get_num_threads(), cur = maxthreads.load();
    while (t > cur && !maxthreads.compare_exchange_weak(cur, t)) {}
  });
  check(maxthreads.load() == omp_get_max_threads(), "for_each: runs on all threads");
  // transform_reduce with a non-commutative-safe associative op (min/max tuple, src/octree.h:95-106) and a sum
  std::vector<double> x(n);
  for (auto& v : x) v = double(gen() % 2000001) - 1e6;
  auto mm = std::transform_reduce(
      std::execution::par_unseq, r.begin(), r.end(), std::make_tuple(0.0, 0.0),
      [](auto l, auto rr) { return std::make_tuple(std::min(std::get<0>(l), std::get<0>(rr)), std::max(std::get<1>(l), std::get<1>(rr))); },
      [p = x.data()](auto i) { return std::make_tuple(p[i], p[i]); });
  const auto [mn, mx] = std::minmax_element(x.begin(), x.end());
  check(std::get<0>(mm) == std::min(0.0, *mn) && std::get<1>(mm) == std::max(0.0, *mx), "transform_reduce: min/max");
  const std::uint64_t s1 = std::transform_reduce(std::execution::par, r.begin(), r.end(), std::uint64_t(5), std::plus<>{},
                                                 [k = keys.data()](auto i) { return k[i]; });
  const std::uint64_t s2 = std::accumulate(keys.begin(), keys.end(), std::uint64_t(5));
  check(s1 == s2, "transform_reduce: integer sum with init");
  // tiny ranges
  std::vector<int> tiny{3, 1, 2};
  std::sort(std::execution::par, tiny.begin(), tiny.end());
  check(tiny == std::vector<int>{1, 2, 3}, "sort: tiny");
  auto r1 = std::views::iota(0, 1);
  check(std::transform_reduce(std::execution::par, r1.begin(), r1.end(), 10, std::plus<>{}, [](int i) { return i + 1; }) == 11,
        "transform_reduce: one element");
  std::printf(bad ? "PSTL_SHIM_TEST FAIL\n" : "PSTL_SHIM_TEST PASS threads=%d\n", omp_get_max_threads());
  return bad ? 1 : 0;
}
