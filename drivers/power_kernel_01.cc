// power_kernel_01 - the power-kernel study of the reference (power_kernel_01.likwid.cc) on the device: dst_0 = A src (Laplace),
// dst_1 = M dst_0 (mass operator), once with the second operator fused into the cell waves of the first ("powero": every cell is
// released on its own; "powerb": cells are released in batches of `n lanes`) and once as two sweeps ("sequential").  Same JSON keys
// (power_kernel_01.likwid.cc:69-82), the same protocol (n repetitions warm-up runs, then n repetitions timed runs, the three vector
// norms printed after each version, 443-477) and the same result table (573-599).
// Differences: the mesh is the hyper-rectangle of subdivided_hyper_cube_balanced without constraints as in the reference, source
// vector sin(x) is replaced by a fixed pseudo-random vector; "n lanes" is the batch size of the "powerb" variant (default 8), not a
// SIMD width; "n components" = 1, "use dg" = false and dim = 3 only; "dof renumbering" has no effect (the library's brick-grouped
// numbering is always on); LIKWID markers are dropped.  Host code over libdasm's C ABI.
//
//   ./power_kernel_01 [input_0.json ...]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <iostream>

#include "../include/dasm/operator.h"
#include "../include/dasm/json.h"

using namespace dasm;

struct Parameters
{
  unsigned int dim              = 3;
  unsigned int fe_degree        = 4;
  unsigned int n_components     = 1;
  unsigned int subdivisions     = 34;
  unsigned int n_lanes          = 0;
  unsigned int cell_granularity = 0;
  unsigned int n_repetitions    = 10;
  bool         dof_renumbering  = false;
  bool         use_dg           = false;
  bool         do_computation   = true;
  std::string  number_type      = "double";

  void
  parse(const std::string &file_name)
  {
    const ptree prm  = ptree::parse_file(file_name);
    dim              = prm.get<unsigned int>("dim", dim);
    fe_degree        = prm.get<unsigned int>("fe degree", fe_degree);
    n_components     = prm.get<unsigned int>("n components", n_components);
    subdivisions     = prm.get<unsigned int>("n subdivisions", subdivisions);
    n_lanes          = prm.get<unsigned int>("n lanes", n_lanes);
    cell_granularity = prm.get<unsigned int>("cell granularity", cell_granularity);
    n_repetitions    = prm.get<unsigned int>("n repetitions", n_repetitions);
    dof_renumbering  = prm.get<bool>("dof renumbering", dof_renumbering);
    use_dg           = prm.get<bool>("use dg", use_dg);
    do_computation   = prm.get<bool>("do computation", do_computation);
    number_type      = prm.get<std::string>("number type", number_type);
    if (number_type != "double" && number_type != "float")
      throw std::runtime_error("number type must be double|float");
  }

  void
  print() const
  {
    std::cout << "{\"dim\": " << dim << ", \"fe degree\": " << fe_degree << ", \"n components\": " << n_components
              << ", \"n subdivisions\": " << subdivisions << ", \"n lanes\": " << n_lanes << ", \"cell granularity\": " << cell_granularity
              << ", \"n repetitions\": " << n_repetitions << ", \"dof renumbering\": " << (dof_renumbering ? "true" : "false")
              << ", \"use dg\": " << (use_dg ? "true" : "false") << ", \"do computation\": " << (do_computation ? "true" : "false")
              << ", \"number type\": \"" << number_type << "\"}" << std::endl;
  }
};

struct Row
{
  unsigned int degree, n_lanes, granularity, n_repetitions, n_procs;
  long long    n_cells, n_dofs;
  double       t_own, t_batch, t_sequential;
};

template <typename Number>
static Row
run(const Parameters &params, Context &ctx)
{
  if (params.n_components != 1 || params.use_dg)
    throw std::runtime_error("ExcNotImplemented: n components = 1 and use dg = false only");
  int n_refine, sub[3];
  check(dasm_decompose_balanced((int)params.subdivisions, &n_refine, sub));
  int    n_cells[3], periodic[3] = {0, 0, 0};
  double length[3];
  for (int d = 0; d < 3; ++d)
    {
      n_cells[d] = sub[d] << n_refine;
      length[d]  = sub[d];
    }
  Mesh                                 mesh(ctx, n_cells, periodic, false, length, DASM_MAP_CARTESIAN);
  LaplaceOperatorMatrixFree<3, Number> op(mesh, params.fe_degree, typename LaplaceOperatorMatrixFree<3, Number>::AdditionalData(true, ""));
  const unsigned int batch = params.n_lanes == 0 ? 8 : params.n_lanes;
  if (params.cell_granularity != 0 && params.cell_granularity < batch)
    throw std::runtime_error("ExcInternalError: cell granularity must not be smaller than the batch size"); // power_kernel_01.likwid.cc:364-365
  dasm_power *own = nullptr, *bat = nullptr;
  check(dasm_power_create(op.handle(), params.cell_granularity, 1, &own));
  check(dasm_power_create(op.handle(), params.cell_granularity, (int)batch, &bat));
  std::cout << mesh.n_cells() << " " << dasm_power_n_waves(own) + 1 << std::endl;
  for (long long w = 0; w < std::min<long long>(dasm_power_n_waves(own), 32); ++w)
    printf("%4lld %4lld\n", w, dasm_power_post_count(own, w));

  Vector<Number> src, dst_0, dst_1;
  op.initialize_dof_vector(src);
  op.initialize_dof_vector(dst_0);
  op.initialize_dof_vector(dst_1);
  {
    std::vector<double> host((size_t)src.locally_owned_size());
    unsigned long long  s = 88172645463325252ull;
    for (auto &v : host)
      {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        v = (double)(s >> 11) / 9007199254740992.0 - 0.5;
      }
    src.upload(host);
  }
  const auto norm = [](const Vector<Number> &v) {
    double s = 0;
    for (const double x : v.download())
      s += x * x;
    return std::sqrt(s);
  };
  const auto measure = [&](dasm_power *p, const int fused) {
    dst_0 = 0.0;
    dst_1 = 0.0;
    for (unsigned int c = 0; c < params.n_repetitions; ++c)
      check(dasm_power_run(p, dst_0.data(), dst_1.data(), src.data(), fused, params.do_computation ? 1 : 0));
    ctx.sync();
    const auto t0 = std::chrono::system_clock::now();
    for (unsigned int c = 0; c < params.n_repetitions; ++c)
      check(dasm_power_run(p, dst_0.data(), dst_1.data(), src.data(), fused, params.do_computation ? 1 : 0));
    ctx.sync();
    const double time = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::system_clock::now() - t0).count() / 1e9;
    std::cout << norm(src) << " " << norm(dst_0) << " " << norm(dst_1) << std::endl;
    return time;
  };
  Row r;
  r.t_own         = measure(own, 1);
  r.t_batch       = measure(bat, 1);
  r.t_sequential  = measure(own, 0);
  r.degree        = params.fe_degree;
  r.n_lanes       = batch;
  r.granularity   = params.cell_granularity;
  r.n_repetitions = params.n_repetitions;
  r.n_procs       = 1;
  r.n_cells       = mesh.n_cells();
  r.n_dofs        = op.m();
  dasm_power_destroy(own);
  dasm_power_destroy(bat);
  return r;
}

static void
write_table(const std::vector<Row> &rows)
{
  // ConvergenceTable::write_text(org_mode_table) of the reference's columns (power_kernel_01.likwid.cc:573-599)
  printf("| degree | n_lanes | granularity | n_repetitions | n_procs | n_cells | n_dofs | s_own | s_batch | t_own | t_batch | t_sequential | tp_own | tp_batch | tp_sequential |\n");
  for (const Row &r : rows)
    {
      const double dofs = 2.0 * (double)r.n_dofs * r.n_repetitions * r.n_procs;
      printf("| %u | %u | %u | %u | %u | %lld | %lld | %.4e | %.4e | %.4e | %.4e | %.4e | %.4e | %.4e | %.4e |\n", r.degree, r.n_lanes, r.granularity,
             r.n_repetitions, r.n_procs, r.n_cells, r.n_dofs, r.t_sequential / r.t_own, r.t_sequential / r.t_batch, r.t_own, r.t_batch, r.t_sequential,
             dofs / r.t_own, dofs / r.t_batch, dofs / r.t_sequential);
    }
  std::cout << std::endl;
}

int
main(int argc, char *argv[])
{
  try
    {
      Context                  ctx(0);
      std::vector<std::string> input_files;
      for (int i = 1; i < argc; ++i)
        input_files.emplace_back(argv[i]);
      if (input_files.empty())
        input_files.push_back("");
      std::vector<Row> table;
      for (const auto &file_name : input_files)
        {
          Parameters params;
          if (file_name != "")
            params.parse(file_name);
          params.print();
          if (params.dim != 3)
            throw std::runtime_error("ExcNotImplemented: only dim = 3 is built");
          table.push_back(params.number_type == "double" ? run<double>(params, ctx) : run<float>(params, ctx));
          write_table(table);
        }
    }
  catch (const std::exception &e)
    {
      std::cerr << "ERROR: " << e.what() << std::endl;
      return 1;
    }
  return 0;
}
