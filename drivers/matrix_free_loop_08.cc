// matrix_free_loop_08 - throughput benchmark of the smoother hot path with the reference driver's interface
// (matrix_free_loop_08.likwid.cc of the reference): same JSON keys (lines 54-72), same label mini-grammar
// (108-144, 244-297), same measurement protocol (n repetitions warm-up calls, then n repetitions timed calls,
// 345-382) and the same machine-readable output line
//   >> label n_dofs repetitions*degree time sizeof(Number) degree n_ghost n_import        (390-395)
// The mesh is the periodic hyper-rectangle of GridGenerator::subdivided_hyper_cube_balanced (160-174).
// Host code is C++ over libdasm's C ABI; LIKWID markers are replaced by CUDA-event timing inside the library.
//
//   ./matrix_free_loop_08 input_0.json [input_1.json ...]
#include <chrono>
#include <iostream>
#include <sstream>

#include "../include/dasm/precondition.h"

using namespace dasm;

struct Parameters
{
  unsigned int dim                  = 3;
  std::string  number_type          = "double";
  unsigned int fe_degree            = 3;
  unsigned int n_subdivision        = 1;
  std::string  preconditioner_types = "post-1-c";
  bool         dof_renumbering      = true;
  bool         compress_indices     = true;
  bool         use_cartesian_mesh   = true;
  std::string  mapping_type         = "default";
  unsigned int n_repetitions        = 10;

  void
  parse(const std::string &file_name)
  {
    const ptree prm      = ptree::parse_file(file_name);
    dim                  = prm.get<unsigned int>("dim", dim);
    number_type          = prm.get<std::string>("number type", number_type);
    fe_degree            = prm.get<unsigned int>("fe degree", fe_degree);
    n_subdivision        = prm.get<unsigned int>("n subdivisions", n_subdivision);
    preconditioner_types = prm.get<std::string>("preconditioner types", preconditioner_types);
    n_repetitions        = prm.get<unsigned int>("n repetitions", n_repetitions);
    dof_renumbering      = prm.get<bool>("dof renumbering", dof_renumbering);
    use_cartesian_mesh   = prm.get<bool>("use cartesian mesh", use_cartesian_mesh);
    mapping_type         = prm.get<std::string>("mapping type", mapping_type);
    if (number_type != "double" && number_type != "float")
      throw std::runtime_error("number type must be double|float");
    if (!dof_renumbering)
      std::cerr << "note: \"dof renumbering\": false has no effect - libdasm always uses its brick-grouped data-locality numbering "
                   "(the counterpart of DoFRenumbering::matrix_free_data_locality, matrix_free_loop_08.likwid.cc:216-222)"
                << std::endl;
  }
};

static std::vector<std::string>
split_string(const std::string &text, const char deliminator, const unsigned int size = 0)
{
  std::stringstream        stream(text);
  std::string              substring;
  std::vector<std::string> list;
  while (std::getline(stream, substring, deliminator))
    list.push_back(substring);
  for (unsigned int i = list.size(); i < size; ++i)
    list.push_back("-");
  return list;
}

// matrix_free_loop_08.likwid.cc:108-144
static void
process_fdm_parameters(const unsigned int offset, const std::vector<std::string> &props, ptree &params, std::string &constness)
{
  const auto type               = props[offset + 0];
  const auto n_overlap          = props[offset + 1];
  const auto weighting_sequence = props[offset + 2];
  const bool overlap_pre_post   = (weighting_sequence == "g") ? (props[offset + 3] == "p") : true;
  constness                     = (weighting_sequence == "g") ? (props[offset + 4]) : std::string("c");
  params.put("weighting type", (type == "add") ? std::string("none") : type);
  if (n_overlap == "v")
    params.put("element centric", false);
  else
    {
      params.put("n overlap", n_overlap);
      params.put("element centric", true);
    }
  params.put("weight sequence",
             weighting_sequence == "g" ? "global" : (weighting_sequence == "l" ? "local" : (weighting_sequence == "dg" ? "DG" : "compressed")));
  params.put("overlap pre post", overlap_pre_post);
}

template <int dim, typename Number>
void
test(const Parameters &params_in, Context &ctx)
{
  using VectorType = Vector<Number>;
  int n_refine, sub[3];
  check(dasm_decompose_balanced((int)params_in.n_subdivision, &n_refine, sub));
  int    n_cells[3], periodic[3] = {1, 1, 1};
  double length[3];
  for (int d = 0; d < 3; ++d)
    {
      n_cells[d] = sub[d] << n_refine;
      length[d]  = sub[d];
    }
  Mesh mesh(ctx, n_cells, periodic, false, length, params_in.use_cartesian_mesh ? DASM_MAP_CARTESIAN : DASM_MAP_SINE);

  typename LaplaceOperatorMatrixFree<dim, Number>::AdditionalData ad_operator;
  ad_operator.compress_indices = params_in.compress_indices;
  if (params_in.mapping_type != "default")
    ad_operator.mapping_type = params_in.mapping_type;
  else
    ad_operator.mapping_type = params_in.use_cartesian_mesh ? "" : "merged"; // reference: "quadratic geometry" (same operator)

  bool info_printed = false;
  for (const auto &label : split_string(params_in.preconditioner_types, ' '))
    {
      const auto   props     = split_string(label, '-', 10);
      const auto   type      = props[0];
      std::string  constness = "c";
      unsigned int factor    = 1;

      LaplaceOperatorMatrixFree<dim, Number> op(mesh, params_in.fe_degree, ad_operator);
      if (!info_printed)
        {
          std::cout << "Info" << std::endl << " - degree: " << params_in.fe_degree << std::endl << " - n dofs: " << op.m() << std::endl << std::endl;
          info_printed = true;
        }
      std::shared_ptr<const ASPoissonPreconditioner<dim, Number>> precondition_fdm;
      std::shared_ptr<const PreconditionerBase<VectorType>>        precondition;
      if (type != "vmult")
        {
          if (type == "cheby")
            {
              ptree params, params_fdm;
              if (props[3] == "diag")
                params_fdm.put("type", "Diagonal");
              else
                {
                  std::string c2;
                  process_fdm_parameters(3, props, params_fdm, c2);
                  params_fdm.put("type", "FDM");
                }
              params.add_child("preconditioner", params_fdm);
              params.put("type", "Chebyshev");
              params.put("degree", std::atoi(props[1].c_str()));
              params.put("optimize", std::atoi(props[2].c_str()));
              factor       = std::atoi(props[1].c_str());
              precondition = create_system_preconditioner<dim, Number>(op, params);
            }
          else
            {
              ptree params;
              process_fdm_parameters(0, props, params, constness);
              precondition_fdm = create_fdm_preconditioner<dim, Number>(op, params);
            }
        }
      VectorType src, dst;
      op.initialize_dof_vector(src);
      op.initialize_dof_vector(dst);
      src = 1.0;

      const auto fu = [&]() {
        if (type == "vmult")
          op.vmult(dst, src);
        else if (precondition_fdm)
          precondition_fdm->vmult(dst, src);
        else if (precondition)
          precondition->step(dst, src);
        else
          throw std::runtime_error("ExcNotImplemented");
      };
      for (unsigned int i = 0; i < params_in.n_repetitions; ++i)
        fu();
      ctx.sync();
      const auto timer = std::chrono::system_clock::now();
      for (unsigned int i = 0; i < params_in.n_repetitions; ++i)
        fu();
      ctx.sync();
      const double time = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::system_clock::now() - timer).count() / 1e9;
      std::cout << ">> " << label << " " << op.m() << " " << params_in.n_repetitions * factor << " " << time << " " << sizeof(Number) << " "
                << params_in.fe_degree << " " << dasm_op_n_ghost(op.handle()) << " " << dasm_op_n_import(op.handle()) << std::endl;
    }
}

int
main(int argc, char *argv[])
{
  try
    {
      if (argc < 2)
        throw std::runtime_error("usage: matrix_free_loop_08 input.json [...]");
      Context ctx(0);
      for (int i = 1; i < argc; ++i)
        {
          Parameters params;
          params.parse(argv[i]);
          if (params.dim == 3 && params.number_type == "float")
            test<3, float>(params, ctx);
          else if (params.dim == 3 && params.number_type == "double")
            test<3, double>(params, ctx);
          else
            throw std::runtime_error("ExcNotImplemented: only dim = 3 is built");
        }
    }
  catch (const std::exception &e)
    {
      std::cerr << "ERROR: " << e.what() << std::endl;
      return 1;
    }
  return 0;
}
