// element_centered_preconditioners_01 - the reference's solver driver for the rows of SURVEY.md section 8(f): Krylov solver + (multigrid)
// preconditioner around the smoother hot path, with the reference driver's JSON interface and output format
// (element_centered_preconditioners_01.cc of the reference: solve() 108-263, test() 266-778, MyMultigrid include/precondition.h:82-186).
//
//   keys: "type" ("matrixfree"), "dim" (3), "degree", "n refinements", "operator mapping type", "operator compress indices",
//         "mesh": {"name": "hypercube" | "kershaw" | "hyperball", "n subdivisions", "n initial refinements", "eps" | "epsy" / "epsz"},
//         "solver": {"type": "CG" | "GMRES", "max iterations", "abs tolerance", "rel tolerance", "max n tmp vectors"},
//         "preconditioner": {"type": "Identity" | "Diagonal" | "FDM" | "Chebyshev" | "AdditiveSchwarzPreconditioner" | "Multigrid",
//                            "mg type": "h" | "p", "mg p sequence": "bisect" | "decrease by one" | "go to one",
//                            "mg smoother": {...}, "mg coarse grid solver": {...}, "one-sided v-cycle"}
//   output: the reference's log lines ("- Create operator:", "- Setting up smoother on level", " - Solving with", "   - n iterations:")
//           and its result table.
// The mesh is libdasm's structured hexahedral mesh (the reference's hyper_cube / subdivided_hyper_cube + global refinement; Kershaw map
// include/kershaw.h); matrix-free level operators are float, the outer operator and the Krylov vectors double
// (LaplaceOperatorMatrixFreeTrait, :787-792); "hyperball" is the unstructured ball of include/dasm/grid_generator.h (FDM with n overlap = 1 on
// degrees >= 2, multigrid h / p / hp / ph through transfers with the parent map of the refinement).  Not available here: dim = 2,
// "type": "matrixbased", AMG.
//
//   ./element_centered_preconditioners_01 input_0.json [input_1.json ...]
#include <chrono>
#include <iostream>
#include <sstream>

#include "../include/dasm/multigrid.h"
#include "../include/dasm/precondition.h"
#include "../include/dasm/restrictors.h"

using namespace dasm;

static std::vector<unsigned int>
create_polynomial_coarsening_sequence(const unsigned int degree, const std::string &type)
{
  // MGTransferGlobalCoarseningTools::create_polynomial_coarsening_sequence
  std::vector<unsigned int> degrees{degree};
  while (degrees.back() > 1)
    {
      const unsigned int d = degrees.back();
      if (type == "bisect")
        degrees.push_back(std::max(d / 2, 1u));
      else if (type == "decrease by one")
        degrees.push_back(d - 1);
      else if (type == "go to one")
        degrees.push_back(1);
      else
        throw std::runtime_error("Multigrid p sequence <" + type + "> is not known!");
    }
  std::reverse(degrees.begin(), degrees.end());
  return degrees;
}

struct Result
{
  long long n_cells = 0, n_dofs = 0;
  int       L = 0, it = 0;
};

template <typename Number>
static void
print_operator(const LaplaceOperatorMatrixFree<3, Number> &op, const long long n_cells, const bool compress, const std::string &mapping)
{
  std::cout << "- Create operator:" << std::endl;
  std::cout << "  - n cells:          " << n_cells << std::endl;
  std::cout << "  - n dofs:           " << op.m() << std::endl;
  std::cout << "  - compress indices: " << (compress ? "true" : "false") << std::endl;
  std::cout << "  - mapping type:     " << mapping << std::endl << std::endl;
}

static int
solve(const LaplaceOperatorMatrixFree<3, double> &op, Vector<double> &x, const Vector<double> &b, const int kind, void *handle, const ptree &params,
      Context &ctx)
{
  const auto max_iterations = params.get<unsigned int>("max iterations", 1000);
  const auto abs_tolerance  = params.get<double>("abs tolerance", 1e-10);
  const auto rel_tolerance  = params.get<double>("rel tolerance", 1e-2);
  const auto type           = params.get<std::string>("type", "");
  std::cout << " - Solving with " << type << std::endl;
  std::cout << "   - max iterations: " << max_iterations << std::endl;
  std::cout << "   - abs tolerance:  " << abs_tolerance << std::endl;
  std::cout << "   - rel tolrance:   " << rel_tolerance << std::endl;
  int solver;
  if (type == "CG")
    solver = DASM_SOLVER_CG;
  else if (type == "GMRES")
    solver = DASM_SOLVER_GMRES;
  else
    throw std::runtime_error("Solver <" + type + "> is not known!");
  const int restart = params.get<int>("max n tmp vectors", 30);
  int       n_it    = 0;
  double    res     = 0;
  auto      run     = [&]() { return dasm_solve(op.handle(), solver, kind, handle, x.data(), b.data(), (int)max_iterations, abs_tolerance, rel_tolerance, restart, &n_it, &res); };
  if (run() != 0) // warm up (the reference solves twice, :221-236)
    {
      std::cout << "   - DID NOT CONVERGE!" << std::endl << std::endl;
      return 999;
    }
  ctx.sync();
  const auto t0 = std::chrono::system_clock::now();
  check(run());
  ctx.sync();
  const double time = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::system_clock::now() - t0).count() / 1e9;
  std::cout << "   - n iterations:   " << n_it << std::endl;
  std::cout << "   - time:           " << time << " #" << std::endl << std::endl;
  return n_it;
}

static Result
test(const ptree &params, Context &ctx)
{
  const unsigned int fe_degree   = params.get<unsigned int>("degree", 1);
  const unsigned int n_refine    = params.get<unsigned int>("n refinements", 6);
  const ptree        solver_prm  = try_get_child(params, "solver");
  const ptree        precon_prm  = try_get_child(params, "preconditioner");
  const auto         precon_type = precon_prm.get<std::string>("type", "");
  const auto         op_mapping  = params.get<std::string>("operator mapping type", "");
  const bool         op_compress = params.get<bool>("operator compress indices", false);
  const ptree        mesh_prm    = try_get_child(params, "mesh");
  const auto         geometry    = mesh_prm.get<std::string>("name", "hypercube");
  if (params.get<unsigned int>("dim", 3) != 3)
    throw std::runtime_error("ExcNotImplemented: libdasm builds the 3-D path");
  if (params.get<std::string>("type", "matrixfree") != "matrixfree")
    throw std::runtime_error("ExcNotImplemented: only the matrix-free operator exists in libdasm");

  int    coarse_cells = 1, map_kind = DASM_MAP_CARTESIAN;
  double map_params[4] = {0, 0, 0, 0};
  unsigned int n_levels_h = n_refine + 1;
  bool         is_ball    = false;
  if (geometry == "hypercube")
    {
      coarse_cells = mesh_prm.get<int>("n subdivisions", 1);
      std::cout << "- Create mesh: hypercube" << std::endl << std::endl;
    }
  else if (geometry == "kershaw")
    {
      double epsy = mesh_prm.get<double>("epsy", 0.0), epsz = mesh_prm.get<double>("epsz", 0.0);
      const int n_initial = mesh_prm.get<int>("n initial refinements", 1);
      coarse_cells        = mesh_prm.get<int>("n subdivisions", 3);
      if (epsy == 0.0 || epsz == 0.0)
        epsy = epsz = mesh_prm.get<double>("eps", 1.0);
      std::cout << "- Create mesh: kershaw" << std::endl;
      std::cout << "  - epsx: " << 1.0 << std::endl;
      std::cout << "  - epsy: " << epsy << std::endl;
      std::cout << "  - epsz: " << epsz << std::endl << std::endl;
      map_kind      = DASM_MAP_KERSHAW;
      map_params[0] = epsy;
      map_params[1] = epsz;
      n_levels_h += n_initial; // subdivided_hyper_cube(n) + n_initial + n_refine global refinements
    }
  else if (geometry == "hyperball")
    {
      // GridGenerator::hyper_ball_balanced + MappingQCache(2) in the reference (:398-402); here include/dasm/grid_generator.h
      is_ball = true;
      std::cout << "- Create mesh: hyperball" << std::endl << std::endl;
    }
  else
    throw std::runtime_error("Geometry with the name <" + geometry + "> is not known!");

  const int    periodic[3] = {0, 0, 0};
  const double length[3]   = {1, 1, 1};
  auto make_mesh = [&](const unsigned int level) {
    const int c = coarse_cells << level;
    const int nc[3] = {c, c, c};
    return std::make_shared<Mesh>(ctx, nc, periodic, /*dirichlet*/ true, length, map_kind, map_params);
  };
  const unsigned int finest = n_levels_h - 1;
  using OperatorType        = LaplaceOperatorMatrixFree<3, double>;
  using LevelOperatorType   = LaplaceOperatorMatrixFree<3, float>;
  std::shared_ptr<Mesh>                          mesh;
  std::vector<std::shared_ptr<UnstructuredMesh>> umeshes(n_levels_h);
  auto make_umesh = [&](const unsigned int level) {
    if (!umeshes[level])
      umeshes[level] = std::make_shared<UnstructuredMesh>(GridGenerator::hyper_ball(level));
    return umeshes[level];
  };
  std::unique_ptr<OperatorType> op_ptr;
  if (is_ball)
    op_ptr.reset(new OperatorType(ctx, *make_umesh(finest), fe_degree, OperatorType::AdditionalData(op_compress, op_mapping)));
  else
    {
      mesh = make_mesh(finest);
      op_ptr.reset(new OperatorType(*mesh, fe_degree, OperatorType::AdditionalData(op_compress, op_mapping)));
    }
  OperatorType &op = *op_ptr;
  print_operator(op, op.n_cells(), op_compress, op_mapping);

  Result result;
  result.n_cells = op.n_cells();
  result.L       = (int)n_levels_h;
  result.n_dofs  = (long long)op.m();

  Vector<double> solution, rhs;
  op.initialize_dof_vector(solution);
  op.initialize_dof_vector(rhs);
  op.rhs(rhs, 1.0); // RightHandSide: f = 1 (:65-81)

  if (precon_type == "Identity")
    {
      std::cout << "- Create system preconditioner: Identity" << std::endl << std::endl;
      result.it = solve(op, solution, rhs, DASM_PRECON_IDENTITY, nullptr, solver_prm, ctx);
    }
  else if (precon_type == "Diagonal")
    {
      std::cout << "- Create system preconditioner: Diagonal" << std::endl << std::endl;
      result.it = solve(op, solution, rhs, DASM_PRECON_DIAGONAL, nullptr, solver_prm, ctx);
    }
  else if (precon_type == "Multigrid")
    {
      std::cout << "- Create system preconditioner: Multigrid" << std::endl;
      const auto mg_type = precon_prm.get<std::string>("mg type", "h");
      const auto mg_seq  = precon_prm.get<std::string>("mg p sequence", "bisect");
      std::cout << " - type:       " << mg_type << std::endl;
      std::cout << " - p sequence: " << mg_seq << std::endl << std::endl;
      const auto mg_degrees = create_polynomial_coarsening_sequence(fe_degree, mg_seq);
      std::vector<std::pair<unsigned int, unsigned int>> levels; // (mesh level, degree)
      if (mg_type == "h")
        for (unsigned int r = 0; r < n_levels_h; ++r)
          levels.emplace_back(r, mg_degrees.back());
      else if (mg_type == "p")
        for (const auto d : mg_degrees)
          levels.emplace_back(finest, d);
      else if (mg_type == "hp")
        {
          for (unsigned int i = 0; i + 1 < mg_degrees.size(); ++i)
            levels.emplace_back(0, mg_degrees[i]);
          for (unsigned int r = 0; r < n_levels_h; ++r)
            levels.emplace_back(r, mg_degrees.back());
        }
      else if (mg_type == "ph")
        {
          for (unsigned int r = 0; r + 1 < n_levels_h; ++r)
            levels.emplace_back(r, mg_degrees.front());
          for (const auto d : mg_degrees)
            levels.emplace_back(finest, d);
        }
      else
        throw std::runtime_error("Multigrid variant <" + mg_type + "> is not known!");
      std::vector<std::shared_ptr<Mesh>>                          meshes(n_levels_h);
      std::vector<std::shared_ptr<LevelOperatorType>>             mg_operators;
      std::vector<std::shared_ptr<const PreconditionerBase<Vector<float>>>> keep;
      std::vector<std::shared_ptr<PreconditionChebyshev<3, float>>>         mg_smoothers;
      for (const auto &lv : levels)
        {
          if (is_ball)
            mg_operators.push_back(std::make_shared<LevelOperatorType>(ctx, *make_umesh(lv.first), lv.second, LevelOperatorType::AdditionalData(op_compress, op_mapping)));
          else
            {
              if (!meshes[lv.first])
                meshes[lv.first] = lv.first == finest ? mesh : make_mesh(lv.first);
              mg_operators.push_back(std::make_shared<LevelOperatorType>(*meshes[lv.first], lv.second, LevelOperatorType::AdditionalData(op_compress, op_mapping)));
            }
          print_operator(*mg_operators.back(), mg_operators.back()->n_cells(), op_compress, op_mapping);
        }
      for (unsigned int l = 0; l < levels.size(); ++l)
        {
          if (l == 0)
            std::cout << "- Setting up coarse-grid solver on level " << l << std::endl << std::endl;
          else
            std::cout << "- Setting up smoother on level " << l << std::endl << std::endl;
          const auto p = create_system_preconditioner<3, float>(*mg_operators[l], try_get_child(precon_prm, l == 0 ? "mg coarse grid solver" : "mg smoother"));
          const auto s = std::dynamic_pointer_cast<const SystemPreconditioner<3, float>>(p);
          if (!s || !s->chebyshev)
            throw std::runtime_error("ExcNotImplemented: multigrid smoothers / coarse-grid solvers must be of type Chebyshev in libdasm");
          keep.push_back(p);
          mg_smoothers.push_back(s->chebyshev);
        }
      std::unique_ptr<PreconditionerGMG<3, float, double>> mg;
      if (is_ball)
        {
          // unstructured levels: the geometric transfers need the parent map of the refinement
          std::vector<std::shared_ptr<MGTwoLevelTransfer<3, float>>> transfers(levels.size());
          for (unsigned int l = 1; l < levels.size(); ++l)
            transfers[l] = std::make_shared<MGTwoLevelTransfer<3, float>>(*mg_operators[l], *mg_operators[l - 1],
                                                                          levels[l].first != levels[l - 1].first ? GridGenerator::ball_parents(levels[l].first) :
                                                                                                                    std::vector<std::uint32_t>());
          mg.reset(new PreconditionerGMG<3, float, double>(mg_operators, mg_smoothers, transfers, precon_prm.get<bool>("one-sided v-cycle", false)));
        }
      else
        mg.reset(new PreconditionerGMG<3, float, double>(mg_operators, mg_smoothers, precon_prm.get<bool>("one-sided v-cycle", false)));
      result.it = solve(op, solution, rhs, DASM_PRECON_MULTIGRID, mg->handle(), solver_prm, ctx);
    }
  else if (precon_type == "AdditiveSchwarzPreconditioner")
    {
      std::cout << "- Create system preconditioner: AdditiveSchwarzPreconditioner" << std::endl << std::endl;
      using Restrictor = Restrictors::ElementCenteredRestrictor<3, double>;
      Restrictor::AdditionalData ad(std::min(precon_prm.get<unsigned int>("n overlap", 1), fe_degree), get_weighting_type(precon_prm),
                                    precon_prm.get<std::string>("restriction type", "element"));
      const auto restrictor = std::make_shared<const Restrictor>(op, ad);
      RestrictedPreconditioner<3, double> precon(restrictor);
      result.it = solve(op, solution, rhs, DASM_PRECON_BLOCK_ASM, precon.handle(), solver_prm, ctx);
    }
  else
    {
      if (precon_type == "")
        throw std::runtime_error("ExcNotImplemented");
      const auto p = create_system_preconditioner<3, double>(op, precon_prm);
      const auto s = std::dynamic_pointer_cast<const SystemPreconditioner<3, double>>(p);
      if (s && s->chebyshev)
        result.it = solve(op, solution, rhs, DASM_PRECON_CHEBYSHEV, s->chebyshev->handle(), solver_prm, ctx);
      else
        result.it = solve(op, solution, rhs, DASM_PRECON_FDM, s->fdm->handle(), solver_prm, ctx);
    }
  return result;
}

int
main(int argc, char *argv[])
{
  try
    {
      Context ctx(0);
      std::vector<Result> table;
      for (int i = 1; i < argc; ++i)
        table.push_back(test(ptree::parse_file(argv[i]), ctx));
      std::cout << "| n_cells | L | n_dofs | it |" << std::endl;
      for (const auto &r : table)
        std::cout << "| " << r.n_cells << " | " << r.L << " | " << r.n_dofs << " | " << r.it << " |" << std::endl;
    }
  catch (const std::exception &e)
    {
      std::cerr << "Exception: " << e.what() << std::endl;
      return 1;
    }
  return 0;
}
