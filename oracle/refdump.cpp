// refdump — TEST INFRASTRUCTURE ONLY (never linked into, or called by, the product path).
//
// A thin dump harness around the UNMODIFIED reference headers. It is compiled with
// `-I/root/reference/src` from the sources where they lie (see oracle/Makefile); no reference source
// is copied into this repository. Each invocation runs ONE reference function on a state read from a
// raw binary file and writes the resulting arrays to another raw binary file, so that
//   (1) the plain-C restatement in oracle/nbody_oracle.c can be pinned against the real reference, and
//   (2) golden vectors for tests/golden/ can be generated (tests/golden/make_golden.py).
// One process per call on purpose: the reference's hilbert_sort keeps function-static scratch sized by
// the FIRST call's n (bvh.h:38,62,73), so re-use inside one process with a larger n would overflow.
//
// usage: refdump_d{2,3} <op> <f|d> <n> <theta> <steps> <in.bin|-> <out.bin>
//   in.bin : double dt, double G, then T m[n], x[n*D], v[n*D], a[n*D], ao[n*D]
//   state out: same layout.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <span>

#ifndef DIM_SIZE
  #error compile with -DDIM_SIZE=2 or 3
#endif

#include "all_pairs.h"
#include "arguments.h"
#include "bvh.h"
#include "models.h"
#include "octree.h"
#include "timer.h"

namespace {

template <typename T>
void put(FILE* f, T const * p, size_t n) {
  if (n && fwrite(p, sizeof(T), n, f) != n) { perror("fwrite"); exit(2); }
}
template <typename T>
void get(FILE* f, T* p, size_t n) {
  if (n && fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
}

template <typename T, dim_t N>
void write_state(FILE* f, System<T, N>& s) {
  double dt = s.dt, G = s.constant;
  put(f, &dt, 1);
  put(f, &G, 1);
  put(f, s.m.data(), s.size);
  put(f, reinterpret_cast<T*>(s.x.data()), size_t(s.size) * N);
  put(f, reinterpret_cast<T*>(s.v.data()), size_t(s.size) * N);
  put(f, reinterpret_cast<T*>(s.a.data()), size_t(s.size) * N);
  put(f, reinterpret_cast<T*>(s.ao.data()), size_t(s.size) * N);
}

template <typename T, dim_t N>
System<T, N> read_state(char const * path, uint32_t n) {
  FILE* f = fopen(path, "rb");
  if (!f) { perror(path); exit(2); }
  double dt, G;
  get(f, &dt, 1);
  get(f, &G, 1);
  System<T, N> s(n, T(dt), T(G));
  get(f, s.m.data(), n);
  get(f, reinterpret_cast<T*>(s.x.data()), size_t(n) * N);
  get(f, reinterpret_cast<T*>(s.v.data()), size_t(n) * N);
  get(f, reinterpret_cast<T*>(s.a.data()), size_t(n) * N);
  get(f, reinterpret_cast<T*>(s.ao.data()), size_t(n) * N);
  fclose(f);
  return s;
}

template <typename T, dim_t N>
int run(std::string op, uint32_t n, double theta, size_t steps, char const * in, char const * out) {
  FILE* fo = fopen(out, "wb");
  if (!fo) { perror(out); return 2; }

  if (op == "galaxy" || op == "uniform" || op == "plummer") {
    Arguments args;
    args.size = n;
    auto s    = op == "galaxy" ? build_galaxy_model<T, N>(args)
                : op == "uniform" ? build_uniform_model<T, N>(args)
                                  : build_plummer_model<T, N>(args);
    uint32_t sz = s.size;
    put(fo, &sz, 1);
    uint32_t pad = 0;
    put(fo, &pad, 1);
    write_state(fo, s);
    fclose(fo);
    return 0;
  }

  auto s = read_state<T, N>(in, n);

  if (op == "force_all_pairs") {
    all_pairs_force(s);
    write_state(fo, s);
  } else if (op == "force_collapsed") {
    all_pairs_collapsed_force(s);
    write_state(fo, s);
  } else if (op == "accelerate") {
    s.accelerate_step();
    write_state(fo, s);
  } else if (op == "energies") {
    auto [k, g] = s.calc_energies();
    put(fo, &k, 1);
    put(fo, &g, 1);
  } else if (op == "run_all_pairs") {
    for (size_t i = 0; i < steps; ++i) { all_pairs_force(s); s.accelerate_step(); }
    write_state(fo, s);
  } else if (op == "run_collapsed") {
    for (size_t i = 0; i < steps; ++i) { all_pairs_collapsed_force(s); s.accelerate_step(); }
    write_state(fo, s);
  } else if (op == "bbox") {
    auto b = bounding_box(std::span{s.x});
    put(fo, &b.xmin.data[0], N);
    put(fo, &b.xmax.data[0], N);
  } else if (op == "keys") {
    // Same expressions as bvh.h:33-45, evaluated through the reference's own cast<>/hilbert<>.
    auto bbox                      = bounding_box(std::span{s.x});
    uint32_t hilbert_cells_per_dim = N == 2 ? 0xffffffff : 0x1fffff;
    vec<T, N> grid_cell_size       = bbox.lengths() / (T)hilbert_cells_per_dim;
    std::vector<uint64_t> keys(n);
    for (uint32_t i = 0; i < n; ++i) {
      vec<uint32_t, N> cell_idx = cast<uint32_t>((s.x[i] - bbox.xmin) / grid_cell_size);
      keys[i]                   = hilbert(cell_idx);
    }
    put(fo, keys.data(), n);
  } else if (op == "sort") {
    auto bbox = bounding_box(std::span{s.x});
    hilbert_sort(s, bbox);
    write_state(fo, s);
  } else if (op == "bvh_build" || op == "bvh_force" || op == "run_bvh") {
    auto tree     = bvh<T, N>::alloc(s);
    size_t nnodes = bvh<T, N>::nnodes_until_level(tree.last_level + 1);
    // dead nodes' b/bw are never written by the reference (bvh.h:185-188,225-228): zero them so dumps are stable
    std::memset((void*)tree.b, 0, nnodes * sizeof(aabb<T, N>));
    std::memset((void*)tree.bw, 0, nnodes * sizeof(T));
    std::memset((void*)tree.m, 0, nnodes * sizeof(monopole<T, N>));
    size_t k = op == "run_bvh" ? steps : 1;
    for (size_t i = 0; i < k; ++i) {
      auto bbox = bounding_box(std::span{s.x});
      hilbert_sort(s, bbox);
      tree.build_tree(s);
      if (op != "bvh_build") tree.compute_force(s, T(theta));
      if (op == "run_bvh") s.accelerate_step();
    }
    write_state(fo, s);
    if (op == "bvh_build") {
      uint64_t nn = nnodes;
      put(fo, &nn, 1);
      put(fo, reinterpret_cast<T*>(tree.m), nnodes * (N + 1));
      put(fo, tree.bw, nnodes);
      put(fo, reinterpret_cast<T*>(tree.b), nnodes * 2 * N);
    }
  } else if (op == "octree_build" || op == "octree_force" || op == "run_octree") {
    auto tree = octree<T, N>::alloc(s.max_tree_node_size);
    size_t k  = op == "run_octree" ? steps : 1;
    for (size_t i = 0; i < k; ++i) {
      tree.clear(s, i == 0 ? tree.capacity : tree.next_free_child_group->load());
      tree.compute_bounds(s);
      tree.insert(s);
      tree.compute_tree(s);
      if (op != "octree_build") tree.compute_force(s, T(theta));
      if (op == "run_octree") s.accelerate_step();
    }
    write_state(fo, s);
    if (op == "octree_build") {
      uint64_t used = tree.next_free_child_group->load();
      put(fo, &used, 1);
      T side = tree.root_side_length;
      put(fo, &side, 1);
      put(fo, &tree.root_x.data[0], N);
      put(fo, tree.first_child, used);
      put(fo, tree.parent, 1 + used / child_count<N>);
      put(fo, reinterpret_cast<T*>(tree.m), used * (N + 1));
    }
  } else {
    fprintf(stderr, "unknown op %s\n", op.c_str());
    return 2;
  }
  fclose(fo);
  return 0;
}

// hilbert_cells: in.bin = uint32 cells[n*D]; out.bin = uint64 key[n]  (known-answer table for vec.h:266-356)
template <dim_t N>
int run_hilbert_cells(uint32_t n, char const * in, char const * out) {
  FILE* f = fopen(in, "rb");
  if (!f) { perror(in); return 2; }
  std::vector<uint32_t> cells(size_t(n) * N);
  get(f, cells.data(), cells.size());
  fclose(f);
  std::vector<uint64_t> keys(n);
  for (uint32_t i = 0; i < n; ++i) {
    vec<uint32_t, N> c;
    for (dim_t d = 0; d < N; ++d) c[d] = cells[size_t(i) * N + d];
    keys[i] = hilbert(c);
  }
  FILE* fo = fopen(out, "wb");
  put(fo, keys.data(), n);
  fclose(fo);
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc != 8) {
    fprintf(stderr, "usage: %s <op> <f|d> <n> <theta> <steps> <in.bin|-> <out.bin>\n", argv[0]);
    return 2;
  }
  std::string op = argv[1];
  bool single    = argv[2][0] == 'f';
  uint32_t n     = (uint32_t)std::stoul(argv[3]);
  double theta   = std::stod(argv[4]);
  size_t steps   = std::stoul(argv[5]);
  if (op == "hilbert_cells") return run_hilbert_cells<DIM_SIZE>(n, argv[6], argv[7]);
  if (single) return run<float, DIM_SIZE>(op, n, theta, steps, argv[6], argv[7]);
  return run<double, DIM_SIZE>(op, n, theta, steps, argv[6], argv[7]);
}
