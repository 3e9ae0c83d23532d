// nbody_d{2,3} — C++ host driver that keeps the reference's driver surface and runs the hot path on libnbx.so.
//
// Same CLI as the reference (src/arguments.h:40-156): -n, -s, --theta, --precision float|double,
// --algorithm all-pairs|all-pairs-collapsed|octree|bvh, --workload uniform|plummer|galaxy|load <file>, --print-state,
// --print-info, --save pos|energy|all|none, --csv-detailed, --csv-total, --help; compiled once per dimension with
// -DDIM_SIZE=2|3 like the reference (src/main.cpp:5-7). Extension: --gpus N shards the targets over N GPUs of the box.
// Same observable outputs: banner (src/main.cpp:22-39), System::print format (src/system.h:90-97), CSV header/rows
// (src/all_pairs.h:58-66,99-104, src/octree.h:278-283,336-346, src/bvh.h:340-344,405-414), positions.bin / energy.bin
// (src/saving.h:85-122), including the 10 hidden warm-up steps of the default mode (src/arguments.h:26).
// All simulation arithmetic happens behind the C ABI of include/nbx.h; this file only generates the initial state
// (bit-identical to src/models.h through the same libstdc++ engine/distributions), moves state_t arrays in and out,
// and formats output.
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <format>
#include <fstream>
#include <iostream>
#include <limits>
#include <numbers>
#include <optional>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "nbx.h"

#ifndef DIM_SIZE
  #error Must specify spatial dimensions by compiling with -DDIM_SIZE=2 or -DDIM_SIZE=3 .
#endif

namespace {

enum class Workload { Uniform, Plummer, Galaxy, Load };

struct Options {
  std::size_t size         = 1000;  // src/arguments.h:24-37 defaults
  std::size_t steps        = 1;
  std::size_t warmup_steps = 10;
  bool single_precision    = true;
  Workload workload        = Workload::Uniform;
  int algorithm            = NBX_OCTREE;
  bool print_state = false, print_info = false;
  double theta   = 0.5;
  bool save_pos = false, save_energy = false, csv_detailed = false, csv_total = false;
  std::optional<std::string> load_input;
  int gpus = 1;          // extension
  bool dry_run = false;  // extension: build the workload, honour --print-state / --save pos for the initial frame, exit
};

[[noreturn]] void die(const std::string& head, const std::string& options) {
  std::cerr << head << std::endl;
  std::cerr << options << std::endl;
  std::exit(EXIT_FAILURE);
}

Options parse(const std::vector<std::string>& args) {
  Options o;
  for (std::size_t i = 0; i < args.size(); ++i) {
    const std::string& a = args[i];
    auto value = [&]() -> const std::string& {
      if (i + 1 >= args.size()) die("Missing value for argument: '" + a + "'", "");
      return args[++i];
    };
    if (a == "-n") o.size = std::stoi(value());
    else if (a == "-s") o.steps = std::stoi(value());
    else if (a == "--theta") o.theta = std::stod(value());
    else if (a == "--gpus") o.gpus = std::stoi(value());
    else if (a == "--dry-run") o.dry_run = true;
    else if (a == "--csv-detailed") o.csv_detailed = true;
    else if (a == "--csv-total") o.csv_total = true;
    else if (a == "--print-state") o.print_state = true;
    else if (a == "--print-info") o.print_info = true;
    else if (a == "--precision") {
      const std::string& v = value();
      if (v == "float") o.single_precision = true;
      else if (v == "double") o.single_precision = false;
      else die("Unknown precision: \"" + v + "\".", "Options are: double, float (default).");
    } else if (a == "--algorithm") {
      const std::string& v = value();
      if (v == "all-pairs") o.algorithm = NBX_ALL_PAIRS;
      else if (v == "all-pairs-collapsed") o.algorithm = NBX_ALL_PAIRS_COLLAPSED;
      else if (v == "octree") o.algorithm = NBX_OCTREE;
      else if (v == "bvh") o.algorithm = NBX_BVH;
      else die("Unknown algorithm: \"" + v + "\".", "Options are: all-pairs, all-pairs-collapsed, octree (default).");
    } else if (a == "--workload") {
      const std::string& v = value();
      if (v == "plummer") o.workload = Workload::Plummer;
      else if (v == "galaxy") o.workload = Workload::Galaxy;
      else if (v == "uniform") o.workload = Workload::Uniform;
      else if (v == "load") {
        o.load_input = value();
        o.workload   = Workload::Load;
      } else die("Unknown workload: \"" + v + "\".", "Options are: plummer, galaxy, uniform (default).");
    } else if (a == "--save") {
      const std::string& v = value();
      if (v == "pos") o.save_pos = true;
      else if (v == "energy") o.save_energy = true;
      else if (v == "all") o.save_pos = o.save_energy = true;
      else if (v == "none") o.save_pos = o.save_energy = false;
      else die("Unknown save options: \"" + v + "\".", "Options are: pos, energy, all, none (default).");
    } else if (a == "--help" || a == "-h") {
      std::cout << "Help:\n"
                   "-n size\t\tNumber of particles to simulate\n"
                   "-s steps\t\tNumber of steps to run simulation for\n"
                   "--theta t\t\tTheta threshold parameter to use in Octree\n"
                   "--precision double|float(default)\t\tSelects floating-point precision\n"
                   "--algorithm all-pairs|all-pairs-collapsed|bvh|octree(default)<algo>\t\tSelects simulation algorithm\n"
                   "--workload plummer|galaxy|uniform(default)|load <file.bin>\t\tSelects workload\n"
                   "--print-state\t\tPrint the initial and final state of the simulation\n"
                   "--print-info\t\tPrint info every timestep\n"
                   "--save pos|energy|all|none(default) \t\tSelects what data to save every timestep\n"
                   "--gpus N\t\t(nbx extension) shard the targets over N GPUs of this box\n"
                   "--help\t\tDisplay this help message and quit\n";
      std::exit(EXIT_SUCCESS);
    } else {
      std::cout << std::format("Unknown argument: '{}'\n", a);
      std::exit(EXIT_FAILURE);
    }
  }
  if (o.csv_detailed && o.csv_total) {
    std::cerr << "Cannot capture a CSV detailed and coarse trace in the same run. Specify one or the other." << std::endl;
    std::exit(EXIT_FAILURE);
  }
  return o;
}

// ---- host copy of System<T,N> in state_t layout (src/system.h:14-50) ---------------------------------------------------
template <typename T, int N>
struct HostSystem {
  std::uint32_t size;
  T dt, constant;
  std::vector<T> m, x, v, a, ao;  // x.. are packed vec<T,N>
  std::mt19937 gen{42};           // src/system.h:22-25
  std::uniform_real_distribution<> angle_dis{0, 2 * std::numbers::pi};
  std::uniform_real_distribution<> unit_dis{0, 1};
  std::uniform_real_distribution<> sym_dis{-1, 1};
  std::size_t next = 0;

  HostSystem(std::uint32_t n, T dt_, T c_)
      : size(n), dt(dt_), constant(c_), m(n), x(std::size_t(n) * N), v(std::size_t(n) * N), a(std::size_t(n) * N), ao(std::size_t(n) * N) {}

  void add(T mass, const std::array<T, N>& pos, const std::array<T, N>& vel) {
    if (next >= size) {  // galaxy with n = 1 places its second centre body past the System (UB in the reference): drop it
      ++next;
      return;
    }
    m[next] = mass;
    for (int k = 0; k < N; ++k) {
      x[next * N + k] = pos[k];
      v[next * N + k] = vel[k];
    }
    ++next;
  }

  void print() const {  // src/system.h:90-97: components 0 and 1 only, 3 significant digits
    for (std::size_t i = 0; i < size; ++i)
      std::cout << std::format("{:02}: m={: .3e}, p=({: .3e}, {: .3e}), v=({: .3e}, {: .3e}), f=({: .3e}, {: .3e})", i, m[i],
                               x[i * N], x[i * N + 1], v[i * N], v[i * N + 1], a[i * N], a[i * N + 1])
                << std::endl;
  }
};

// ---- workloads (src/models.h). Every expression keeps the reference's types and evaluation order so that the same
// libstdc++ produces the same bytes; built with -ffp-contract=off like the pinned oracle. ---------------------------------
template <typename T, int N>
HostSystem<T, N> make_uniform(const Options& o) {  // src/models.h:12-28
  HostSystem<T, N> s(std::uint32_t(o.size), T(1e-1), T(1));
  for (std::size_t p = 0; p < o.size; ++p) {
    const T mass = 1.0 / static_cast<T>(o.size);
    std::array<T, N> pos{}, vel{};
    for (int k = 0; k < N; ++k) {
      pos[k] = s.sym_dis(s.gen);
      vel[k] = s.sym_dis(s.gen);
    }
    s.add(mass, pos, vel);
  }
  return s;
}

template <typename T, int N>
HostSystem<T, N> make_plummer(const Options& o) {  // src/models.h:30-66 (3-D only, :68-71)
  if constexpr (N != 3) {
    throw std::runtime_error(std::format("Cannot build Plummer model for D={}", N));
  } else {
    HostSystem<T, N> s(std::uint32_t(o.size), T(1), static_cast<T>(6.674e-11));
    for (std::size_t p = 0; p < o.size; ++p) {
      const T mass    = 1.0 / static_cast<T>(o.size);
      const T radius  = 1.0 / std::sqrt(std::pow(s.unit_dis(s.gen), -2.0 / 3.0) - 1);
      const T p_theta = std::acos(s.sym_dis(s.gen));
      const T p_phi   = s.angle_dis(s.gen);
      std::array<T, N> pos{T(std::sin(p_theta) * std::cos(p_phi)), T(std::sin(p_theta) * std::sin(p_phi)), T(std::cos(p_theta))};
      for (auto& c : pos) c *= radius;
      T q = 0.0, g = 0.1;  // rejection sampling of the speed
      while (g > q * q * std::pow(1.0 - q * q, 3.5)) {
        q = s.unit_dis(s.gen);
        g = 0.1 * s.unit_dis(s.gen);
      }
      const T vnorm   = q * std::numbers::sqrt2 * std::pow(radius * radius + 1, -0.25);
      const T v_theta = std::acos(s.sym_dis(s.gen));
      const T v_phi   = s.angle_dis(s.gen);
      std::array<T, N> vel{T(std::sin(v_theta) * std::cos(v_phi)), T(std::sin(v_theta) * std::sin(v_phi)), T(std::cos(v_theta))};
      for (auto& c : vel) c *= vnorm;
      s.add(mass, pos, vel);
    }
    return s;
  }
}

template <typename T, int N>
void add_orbiting_disc(HostSystem<T, N>& s, std::size_t count, T total_mass, T orbit_mass, const std::array<T, N>& centre) {
  constexpr T eps = std::numeric_limits<T>::epsilon();
  for (std::size_t p = 0; p < count; ++p) {  // src/models.h:81-110
    const T mass   = orbit_mass / static_cast<T>(count);
    const T radius = 30 + 20 * s.unit_dis(s.gen);
    const T angle  = s.angle_dis(s.gen);
    std::array<T, N> pos{};
    pos[0] = std::sin(angle);
    pos[1] = std::cos(angle);
    for (auto& c : pos) c *= radius;
    const T vnorm = std::sqrt(s.constant * total_mass / (radius + eps));
    T norm2       = T(0.);
    for (auto c : pos) norm2 += c * c;
    const T scale = vnorm / (std::sqrt(norm2) + eps);
    std::array<T, N> vel{};
    vel[0] = -pos[1];
    vel[1] = pos[0];
    for (auto& c : vel) c *= scale;
    if constexpr (N == 3) {
      pos[2] = 10 * s.sym_dis(s.gen);
      vel[2] = 0.00001 * s.sym_dis(s.gen);
      const T tilt[3][3] = {{T(0.0), T(-1.0), T(0.0)}, {T(0.9), T(0.0), T(0.5)}, {T(0.5), T(0.0), T(0.9)}};
      std::array<T, N> rp{}, rv{};
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          rp[i] += tilt[i][j] * pos[j];
          rv[i] += tilt[i][j] * vel[j];
        }
      pos = rp;
      vel = rv;
    }
    for (int k = 0; k < N; ++k) pos[k] = pos[k] + centre[k];
    s.add(mass, pos, vel);
  }
}

template <typename T, int N>
HostSystem<T, N> make_galaxy(const Options& o) {  // src/models.h:112-136
  const double per_galaxy = o.size / 2.0;
  HostSystem<T, N> s(std::uint32_t(2 * per_galaxy), T(1e1), T(1e-4));
  T centre_mass = 1e4;
  const T offset = 100.0;
  const T signs[2][2] = {{T(-1), T(1 / 2.0)}, {T(1), T(-1 / 2.0)}};
  for (int g = 0; g < 2; ++g) {
    std::array<T, N> centre{};
    centre[0] = signs[g][0];
    centre[1] = signs[g][1];
    for (auto& c : centre) c *= offset;
    s.add(centre_mass, centre, std::array<T, N>{});
    add_orbiting_disc<T, N>(s, std::size_t(per_galaxy - 1), centre_mass + 1, T(1), centre);
    centre_mass /= 10;
  }
  return s;
}

template <typename T, int N>
HostSystem<T, N> load_system(const std::string& file) {  // src/saving.h:25-68: u32 n, u32 dim, f32 dt, f32 G, n x (m,x[dim],v[dim]) f32
  std::ifstream in(file, std::ios::binary);
  std::uint32_t n = 0, dim = 0;
  float dt = 0, G = 0;
  in.read(reinterpret_cast<char*>(&n), 4);
  in.read(reinterpret_cast<char*>(&dim), 4);
  in.read(reinterpret_cast<char*>(&dt), 4);
  in.read(reinterpret_cast<char*>(&G), 4);
  if (dim != N) throw std::runtime_error(std::format("This version is compiled with D={}, but the file provided is D={}", N, dim));
  const std::size_t stride = 1 + 2 * dim;
  std::vector<float> data(std::size_t(n) * stride);
  in.read(reinterpret_cast<char*>(data.data()), data.size() * sizeof(float));
  HostSystem<T, N> s(n, dt, G);
  for (std::size_t i = 0; i < n; ++i) {
    std::array<T, N> pos{}, vel{};
    for (std::uint32_t k = 0; k < dim; ++k) {
      pos[k] = data[i * stride + 1 + k];
      vel[k] = data[i * stride + 1 + N + k];
    }
    s.add(data[i * stride], pos, vel);
  }
  return s;
}

// ---- engines (one per GPU) --------------------------------------------------------------------------------------------------
void check(int rc, const char* what) {
  if (rc != NBX_OK) throw std::runtime_error(std::format("{}: {} ({})", what, nbx_last_error(), rc));
}

template <typename T, int N>
struct Engines {
  std::vector<nbx_engine*> e;
  HostSystem<T, N>& sys;
  Engines(HostSystem<T, N>& s, const Options& o) : sys(s) {
    if (nbx_device_count() < o.gpus) throw std::runtime_error(std::format("need {} CUDA device(s), found {} (nbx has no CPU path)", o.gpus, nbx_device_count()));
    unsigned char id[NBX_UNIQUE_ID_BYTES];
    if (o.gpus > 1) check(nbx_comm_unique_id(id), "nbx_comm_unique_id");
    e.resize(o.gpus, nullptr);
    for (int r = 0; r < o.gpus; ++r) {
      nbx_config c{};
      c.struct_size = sizeof(c);
      c.dim         = N;
      c.precision   = sizeof(T);
      c.algorithm   = o.algorithm;
      c.n           = s.size;
      c.device      = r;
      c.dt          = s.dt;
      c.G           = s.constant;
      c.theta       = static_cast<T>(o.theta);
      c.rank        = r;
      c.world_size  = o.gpus;
      check(nbx_create(&c, &e[r]), "nbx_create");
    }
    if (o.gpus > 1) each([&](int r) { check(nbx_comm_init_rank(e[r], id), "nbx_comm_init_rank"); });
    // every engine ends up with the full replicated state; with several GPUs each one copies only its shard of the bodies
    // over PCIe and the shards are all-gathered over NVLink (a collective: all host threads call it)
    each([&](int r) {
      if (o.gpus > 1) check(nbx_upload_shard(e[r], s.m.data(), s.x.data(), s.v.data(), s.a.data(), s.ao.data()), "nbx_upload_shard");
      else check(nbx_upload(e[r], s.m.data(), s.x.data(), s.v.data(), s.a.data(), s.ao.data()), "nbx_upload");
    });
  }
  ~Engines() {
    for (auto* p : e) nbx_destroy(p);
  }
  template <typename F>
  void each(F&& f) {  // collectives need every rank in flight at once: one host thread per GPU
    if (e.size() == 1) return f(0);
    std::vector<std::thread> th;
    std::vector<std::exception_ptr> err(e.size());
    for (int r = 0; r < int(e.size()); ++r)
      th.emplace_back([&, r] {
        try { f(r); } catch (...) { err[r] = std::current_exception(); }
      });
    for (auto& t : th) t.join();
    for (auto& x : err)
      if (x) std::rethrow_exception(x);
  }
  void step(std::uint32_t k) {
    each([&](int r) {
      check(nbx_step(e[r], k), "nbx_step");
      check(nbx_sync(e[r]), "nbx_sync");
    });
  }
  void download() { check(nbx_download(e[0], sys.m.data(), sys.x.data(), sys.v.data(), sys.a.data(), sys.ao.data()), "nbx_download"); }
};

// ---- Saver (src/saving.h:8-123) ------------------------------------------------------------------------------------------
template <typename T, int N>
struct Saver {
  bool pos, energy;
  std::uint32_t n, steps, data_size = sizeof(T);
  std::ofstream fpos, fen;
  Saver(const Options& o, std::uint32_t size) : pos(o.save_pos), energy(o.save_energy), n(size), steps(std::uint32_t(o.steps)) {
    if (pos) {
      fpos.open("positions.bin", std::ios::out | std::ios::binary);
      std::uint32_t hdr[4] = {std::uint32_t(o.size), steps, data_size, std::uint32_t(N)};
      fpos.write(reinterpret_cast<const char*>(hdr), sizeof(hdr));
    }
    if (energy) {
      fen.open("energy.bin", std::ios::out | std::ios::binary);
      std::uint32_t hdr[2] = {steps, data_size};
      fen.write(reinterpret_cast<const char*>(hdr), sizeof(hdr));
    }
  }
  // Streaming variant for the per-step frames: `begin` snapshots the positions and starts the device->host copy on a
  // second stream (overlapping the next step), `flush` writes the oldest pending frame to positions.bin.
  int pending = 0;
  void begin_frame(Engines<T, N>& eng) {
    if (!pos) return;
    if (pending == 2) flush_frame(eng);
    check(nbx_stream_positions_begin(eng.e[0]), "nbx_stream_positions_begin");
    ++pending;
  }
  void flush_frame(Engines<T, N>& eng) {
    if (!pos || pending == 0) return;
    check(nbx_stream_positions_end(eng.e[0], eng.sys.x.data()), "nbx_stream_positions_end");
    fpos.write(reinterpret_cast<const char*>(eng.sys.x.data()), std::streamsize(n) * data_size * N);
    --pending;
  }
  void save_energy_only(Engines<T, N>& eng) {
    if (!energy) return;
    double k = 0, g = 0;
    check(nbx_calc_energies(eng.e[0], &k, &g), "nbx_calc_energies");
    T kt = T(k), gt = T(g);
    fen.write(reinterpret_cast<const char*>(&kt), sizeof(T));
    fen.write(reinterpret_cast<const char*>(&gt), sizeof(T));
  }
  void save_all(Engines<T, N>& eng, bool state_is_current) {
    if (!pos && !energy) return;
    if (pos) {
      if (!state_is_current) check(nbx_download(eng.e[0], nullptr, eng.sys.x.data(), nullptr, nullptr, nullptr), "nbx_download");
      fpos.write(reinterpret_cast<const char*>(eng.sys.x.data()), std::streamsize(n) * data_size * N);
    }
    if (energy) {
      double k = 0, g = 0;
      check(nbx_calc_energies(eng.e[0], &k, &g), "nbx_calc_energies");
      T kt = T(k), gt = T(g);
      fen.write(reinterpret_cast<const char*>(&kt), sizeof(T));
      fen.write(reinterpret_cast<const char*>(&gt), sizeof(T));
    }
  }
};

using dsec = std::chrono::duration<double>;
template <typename F>
dsec timed(F&& f) {
  auto t0 = std::chrono::steady_clock::now();
  f();
  return dsec(std::chrono::steady_clock::now() - t0);
}

const char* algo_name(int a) {
  switch (a) {
    case NBX_ALL_PAIRS: return "all-pairs";
    case NBX_ALL_PAIRS_COLLAPSED: return "all-pairs-collapsed";
    case NBX_OCTREE: return "octree";
    default: return "bvh";
  }
}

// The body of run_all_pairs / run_octree / run_bvh (src/all_pairs.h:52-106, src/octree.h:266-347, src/bvh.h:327-418)
template <typename T, int N>
void run_algorithm(HostSystem<T, N>& sys, Options o) {
  Engines<T, N> eng(sys, o);
  Saver<T, N> saver(o, sys.size);
  saver.save_all(eng, true);
  const bool tree = o.algorithm == NBX_OCTREE || o.algorithm == NBX_BVH;
  if (o.csv_total && (o.print_state || o.print_info || o.save_pos || o.save_energy)) std::abort();
  // all-pairs prints its header only under --csv-total (src/all_pairs.h:58-66)
  if (tree ? (o.csv_total || o.csv_detailed) : o.csv_total) {
    std::cout << "algorithm,dim,precision,nsteps,nbodies,total [s]";
    if (o.csv_detailed) {
      if (o.algorithm == NBX_OCTREE) std::cout << ",force [s],accel [s],clear [s],bbox [s],insert [s],multipoles [s],force approx [s]";
      else if (o.algorithm == NBX_BVH) std::cout << ",force [s],accel [s],bbox [s],sort [s],multipoles [s],force approx [s]";
      else std::cout << ",force [s],accel [s]";
    }
    std::cout << "\n";
  }
  dsec total(0);
  double ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // accumulated device phase times [s]: nbx PhaseSlot order
  if (o.csv_detailed) {
    eng.each([&](int r) { check(nbx_set_phase_timing(eng.e[r], 1), "nbx_set_phase_timing"); });
    total = timed([&] {
      for (std::size_t step = 0; step < o.steps; ++step) {
        eng.step(1);
        float ms[8];
        int cnt = 0;
        check(nbx_get_phase_ms(eng.e[0], ms, 8, &cnt), "nbx_get_phase_ms");
        for (int k = 0; k < cnt; ++k) ph[k] += ms[k] * 1e-3;
        if (o.print_info && tree) {  // src/octree.h:313-316, src/bvh.h:377 (root monopole mass == sum of the masses)
          if (o.algorithm == NBX_OCTREE) {
            std::uint64_t used = 0;
            check(nbx_octree_get_root(eng.e[0], nullptr, nullptr, &used), "nbx_octree_get_root");
            std::cout << std::format("Tree size: {}\n", used);
          }
          T total_mass = 0;
          for (T mi : sys.m) total_mass += mi;
          std::cout << std::format("Total mass: {: .5f}\n", total_mass);
        }
        // frame k is copied to the host while step k+1 runs (src/saving.h:110-114 writes it synchronously)
        saver.begin_frame(eng);
        saver.save_energy_only(eng);
        if (saver.pending == 2) saver.flush_frame(eng);
      }
      while (saver.pending) saver.flush_frame(eng);
    });
  } else {
    eng.step(std::uint32_t(o.warmup_steps));
    total = timed([&] {
      if (o.steps > o.warmup_steps) eng.step(std::uint32_t(o.steps - o.warmup_steps));
    });
    o.steps -= o.warmup_steps;  // yes, this underflows for -s < 10 exactly like the reference
  }
  eng.download();
  if (o.csv_detailed || o.csv_total) {
    std::cout << std::format("{},{},{},{},{},{:.2f}", algo_name(o.algorithm), N, sizeof(T) * 8, o.steps, sys.size, total.count());
    if (o.csv_detailed) {
      // slots: 0 force, 1 accel, 2 bbox, 3 sort, 4 build, 5 multipoles, 6 traverse, 7 comm
      if (o.algorithm == NBX_OCTREE)
        std::cout << std::format(",{:.2f},{:.2f},{:.2f},{:.2f},{:.2f},{:.2f},{:.2f}", ph[2] + ph[3] + ph[4] + ph[5] + ph[6], ph[1], 0.0, ph[2],
                                 ph[3] + ph[4], ph[5], ph[6]);
      else if (o.algorithm == NBX_BVH)
        std::cout << std::format(",{:.2f},{:.2f},{:.2f},{:.2f},{:.2f},{:.2f}", ph[2] + ph[3] + ph[5] + ph[6], ph[1], ph[2], ph[3], ph[5], ph[6]);
      else std::cout << std::format(",{:.2f},{:.2f}", ph[0], ph[1]);
    }
    std::cout << "\n";
  }
}

template <typename T, int N>
void run_precision(Options o) {  // src/main.cpp:42-65
  auto sys = [&] {
    switch (o.workload) {
      case Workload::Plummer: return make_plummer<T, N>(o);
      case Workload::Uniform: return make_uniform<T, N>(o);
      case Workload::Galaxy: return make_galaxy<T, N>(o);
      case Workload::Load: {
        auto s = load_system<T, N>(o.load_input.value());
        o.size = s.size;
        return s;
      }
    }
    throw std::runtime_error("Unknown simulation type");
  }();
  if (o.print_state) {  // src/main.cpp:19-40
    std::cout << "Starting state:" << std::endl;
    sys.print();
  }
  if (o.dry_run) {  // no GPU touched: used to check the workload generators byte-for-byte against the reference
    if (o.save_pos) {
      std::ofstream f("positions.bin", std::ios::out | std::ios::binary);
      std::uint32_t hdr[4] = {std::uint32_t(o.size), std::uint32_t(o.steps), std::uint32_t(sizeof(T)), std::uint32_t(N)};
      f.write(reinterpret_cast<const char*>(hdr), sizeof(hdr));
      f.write(reinterpret_cast<const char*>(sys.x.data()), std::streamsize(sys.x.size() * sizeof(T)));
      f.write(reinterpret_cast<const char*>(sys.v.data()), std::streamsize(sys.v.size() * sizeof(T)));
      f.write(reinterpret_cast<const char*>(sys.m.data()), std::streamsize(sys.m.size() * sizeof(T)));
    }
    return;
  }
  const bool quiet = o.csv_total || o.csv_detailed;
  if (!quiet) std::cout << "Starting simulation" << std::endl;
  auto t0 = std::chrono::steady_clock::now();
  run_algorithm<T, N>(sys, o);
  auto t1 = std::chrono::steady_clock::now();
  if (o.print_state) {
    std::cout << "Final state:" << std::endl;
    sys.print();
  }
  if (!quiet) std::cout << std::format("Done simulation\nTotal time: {:.2f} ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count());
}

}  // namespace

int main(int argc, char* argv[]) {
  Options o = parse(std::vector<std::string>(argv + 1, argv + argc));
  try {
    if (o.single_precision) run_precision<float, DIM_SIZE>(o);
    else run_precision<double, DIM_SIZE>(o);
  } catch (const std::exception& ex) {
    std::cerr << "nbody: " << ex.what() << std::endl;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
