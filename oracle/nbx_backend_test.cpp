// nbx_backend_test.cpp — proves the drop-in binding with the reference's OWN code: the unmodified src/main.cpp is
// included as a header (its `main` renamed), so `sim_func_t`, `run_simulation`, `parse_args`, the workload builders and
// `System::print` are the reference's; only the bound simulation function is ours (run_nbx, oracle/nbx_backend.h).
// Built by `make -C oracle binding` into oracle/_ref/nbody_nbx_d{2,3} and linked to libnbx.so; tests/test_binding_gpu.py
// diffs its --print-state output against oracle/_ref/nbody_d{2,3} run with the same arguments.
// TEST INFRASTRUCTURE ONLY (needs /root/reference at build time).
#define main reference_main_unused
#include "main.cpp"  // /root/reference/src/main.cpp via -I
#undef main

#include "nbx_backend.h"

template <typename T, dim_t N>
static void run_precision_nbx(Arguments arguments) {  // src/main.cpp:42-65 with the extra `case` of INTEGRATION.md §2
  auto system = [&arguments] {
    switch (arguments.simulation_type) {
      case SimulationType::Plummer: return build_plummer_model<T, N>(arguments);
      case SimulationType::Uniform: return build_uniform_model<T, N>(arguments);
      case SimulationType::Galaxy: return build_galaxy_model<T, N>(arguments);
      case SimulationType::Load: {
        auto system    = Saver<T, N>::load_system(arguments.load_input.value());
        arguments.size = system.size;
        return system;
      }
      default: throw std::runtime_error("Unknown simulation type");
    }
  }();
  sim_func_t<T, N> f = run_nbx<T, N>;  // the seam: std::function<void(System<T,N>&, Arguments)>
  return run_simulation<T, N>(arguments, system, f);
}

int main(int argc, char* argv[]) {
  auto arguments = parse_args(std::vector<std::string>(argv + 1, argv + argc));
  try {
    if (arguments.single_precision) run_precision_nbx<float, DIM_SIZE>(arguments);
    else run_precision_nbx<double, DIM_SIZE>(arguments);
  } catch (const std::exception& ex) {
    std::cerr << "nbody_nbx: " << ex.what() << std::endl;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
