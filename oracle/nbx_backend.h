// nbx_backend.h — the binding a maintainer of the reference would add as src/nbx_backend.h (INTEGRATION.md §2 prints this
// file verbatim). It is compiled HERE against the reference's own headers (oracle/Makefile target `binding`) so that a
// drift between include/nbx.h and the reference's seam (sim_func_t, src/main.cpp:16-17; System::state_t,
// src/system.h:41-50; Arguments, src/arguments.h:23-38) breaks the build instead of going unnoticed.
// TEST INFRASTRUCTURE: lives under oracle/ because it needs /root/reference; the product never includes it.
#pragma once
#include <nbx.h>  // this repository's include/nbx.h

#include <stdexcept>

#include "arguments.h"
#include "saving.h"
#include "system.h"
#include "timer.h"

template <typename T, dim_t N>
void run_nbx(System<T, N>& system, Arguments arguments) {  // a sim_func_t<T,N>, like run_octree / run_bvh
  nbx_config cfg{};
  cfg.struct_size = sizeof(cfg);
  cfg.dim         = N;                                            // -DDIM_SIZE
  cfg.precision   = sizeof(T);                                    // NBX_F32 / NBX_F64
  cfg.algorithm   = static_cast<int>(arguments.simulation_algo);  // SimulationAlgo order == nbx_algorithm order
  cfg.n           = system.size;
  cfg.dt          = system.dt;
  cfg.G           = system.constant;
  cfg.theta       = arguments.theta;
  cfg.world_size  = 1;
  nbx_engine* e = nullptr;
  auto s        = system.state();  // src/system.h:47-50
  auto ok       = [&](int rc) {
    if (rc != NBX_OK) throw std::runtime_error(nbx_last_error());
  };
  ok(nbx_create(&cfg, &e));
  ok(nbx_upload(e, s.m, s.x, s.v, s.a, s.ao));
  // default mode of the reference drivers: warm-up steps, then the timed steps (src/all_pairs.h:86-97)
  ok(nbx_step(e, arguments.warmup_steps));
  auto dt_total = time([&] {
    if (arguments.steps > arguments.warmup_steps) ok(nbx_step(e, arguments.steps - arguments.warmup_steps));
    ok(nbx_sync(e));
  });
  ok(nbx_download(e, s.m, s.x, s.v, s.a, s.ao));  // state back into the System the caller prints / saves
  nbx_destroy(e);
  arguments.steps -= arguments.warmup_steps;
  if (arguments.csv_total) {  // CSV row exactly as src/all_pairs.h:58-66,99-104
    static const char* names[] = {"all-pairs", "all-pairs-collapsed", "octree", "bvh"};
    std::cout << "algorithm,dim,precision,nsteps,nbodies,total [s]\n";
    std::cout << std::format("{},{},{},{},{},{:.2f}\n", names[cfg.algorithm], N, sizeof(T) * 8, arguments.steps, system.size,
                             dt_total.count());
  }
}
