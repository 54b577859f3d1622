// Factories with the reference's JSON vocabulary (include/precondition.h:9-20, precondition.templates.h):
//   get_weighting_type             templates.h:10-29
//   create_fdm_preconditioner      templates.h:162-247
//   create_system_preconditioner   templates.h:251-818, types on the hot path: "Chebyshev" (around "FDM" or "Diagonal")
//                                  and "FDM"; the other types (AMG, AdditiveSchwarzPreconditioner, ...) are rejected
#pragma once
#include <algorithm>
#include <iostream>

#include "json.h"
#include "preconditioners.h"

namespace dasm
{
  inline Restrictors::WeightingType
  get_weighting_type(const ptree &params)
  {
    const auto type = params.get<std::string>("weighting type", "symm");
    if (type == "symm")
      return Restrictors::WeightingType::symm;
    else if (type == "pre")
      return Restrictors::WeightingType::pre;
    else if (type == "post")
      return Restrictors::WeightingType::post;
    else if (type == "ras")
      return Restrictors::WeightingType::ras;
    else if (type == "none")
      return Restrictors::WeightingType::none;
    throw std::runtime_error("Weighting type <" + type + "> is not known!");
  }

  template <int dim, typename Number>
  std::shared_ptr<const ASPoissonPreconditioner<dim, Number>>
  create_fdm_preconditioner(const LaplaceOperatorMatrixFree<dim, Number> &op, const ptree &params, std::ostream *pcout = &std::cout)
  {
    const unsigned int fe_degree   = op.get_fe_degree();
    const unsigned int n_overlap   = std::min(params.get<unsigned int>("n overlap", 1), fe_degree);
    const auto         weight_type = get_weighting_type(params);
    const unsigned int sub_mesh    = params.get<unsigned int>("sub mesh approximation", dim);
    const bool         reuse       = params.get<bool>("reuse partitioner", true);
    const auto         seq         = params.get<std::string>("weight sequence", n_overlap > 1 ? "global" : "compressed");
    const bool         overlap_pp  = params.get<bool>("overlap pre post", true);
    const bool         el_centric  = params.get<bool>("element centric", true);
    if (pcout)
      {
        *pcout << "- Create system preconditioner: FDM" << std::endl;
        *pcout << "    - n overlap:              " << n_overlap << std::endl;
        *pcout << "    - sub mesh approximation: " << sub_mesh << std::endl;
        *pcout << "    - reuse partitioner:      " << (reuse ? "true" : "false") << std::endl << std::endl;
      }
    return std::make_shared<const ASPoissonPreconditioner<dim, Number>>(op, n_overlap, sub_mesh, weight_type, seq, overlap_pp, el_centric);
  }

  template <int dim, typename Number>
  struct SystemPreconditioner : public PreconditionerBase<Vector<Number>>
  {
    std::shared_ptr<const ASPoissonPreconditioner<dim, Number>> fdm;
    std::shared_ptr<PreconditionChebyshev<dim, Number>>         chebyshev;
    void
    vmult(Vector<Number> &dst, const Vector<Number> &src) const override
    {
      if (chebyshev)
        chebyshev->vmult(dst, src);
      else
        fdm->vmult(dst, src);
    }
    void
    step(Vector<Number> &dst, const Vector<Number> &src) const override
    {
      if (chebyshev)
        chebyshev->step(dst, src);
      else
        throw std::runtime_error("ExcNotImplemented");
    }
  };

  template <int dim, typename Number>
  std::shared_ptr<const PreconditionerBase<Vector<Number>>>
  create_system_preconditioner(const LaplaceOperatorMatrixFree<dim, Number> &op, const ptree &params, std::ostream *pcout = &std::cout)
  {
    const auto type = params.get<std::string>("type", "");
    auto       out  = std::make_shared<SystemPreconditioner<dim, Number>>();
    if (type == "Chebyshev")
      {
        const ptree pp = try_get_child(params, "preconditioner");
        const auto  pt = pp.get<std::string>("type", "");
        if (pt == "")
          throw std::runtime_error("ExcNotImplemented");
        typename PreconditionChebyshev<dim, Number>::AdditionalData ad;
        ad.degree          = params.get<unsigned int>("degree", 3);
        ad.smoothing_range = params.get<double>("smoothing range", 20.);
        const auto ev      = params.get<std::string>("ev algorithm", "");
        if (ev == "lanczos")
          ad.eigenvalue_algorithm = DASM_EV_LANCZOS;
        else if (ev == "power iteration")
          ad.eigenvalue_algorithm = DASM_EV_POWER_ITERATION;
        else if (ev != "")
          throw std::runtime_error("Eigen-value algorithm <" + ev + "> is not known!");
        const auto poly = params.get<std::string>("polynomial type", "1st kind");
        if (poly == "1st kind")
          ad.polynomial_type = DASM_POLY_FIRST_KIND;
        else if (poly == "4th kind")
          ad.polynomial_type = DASM_POLY_FOURTH_KIND;
        else
          throw std::runtime_error("Polynomial type <" + poly + "> is not known!");
        if (pt == "Diagonal")
          {
            if (pcout)
              *pcout << "- Create system preconditioner: Diagonal" << std::endl << std::endl;
            ad.optimize    = params.get<unsigned int>("optimize", 3);
            out->chebyshev = std::make_shared<PreconditionChebyshev<dim, Number>>(op, nullptr, ad);
          }
        else if (pt == "FDM")
          {
            const unsigned int n_overlap = pp.get<unsigned int>("n overlap", 1);
            ad.optimize                  = params.get<unsigned int>("optimize", (n_overlap == 1) ? 2 : 1);
            out->fdm                     = create_fdm_preconditioner<dim, Number>(op, pp, pcout);
            out->chebyshev               = std::make_shared<PreconditionChebyshev<dim, Number>>(op, out->fdm.get(), ad);
          }
        else
          throw std::runtime_error("Preconditioner <" + pt + "> is not known!");
        Vector<Number> vec;
        const auto     evs = out->chebyshev->estimate_eigenvalues(vec);
        if (pcout)
          {
            *pcout << "- Create system preconditioner: Chebyshev" << std::endl;
            *pcout << "    - degree: " << ad.degree << std::endl;
            *pcout << "    - min ev: " << evs.min_eigenvalue_estimate << std::endl;
            *pcout << "    - max ev: " << evs.max_eigenvalue_estimate << std::endl;
            *pcout << "    - omega:  " << 2.0 / (evs.min_eigenvalue_estimate + evs.max_eigenvalue_estimate) << std::endl << std::endl;
          }
        return out;
      }
    else if (type == "FDM")
      {
        out->fdm = create_fdm_preconditioner<dim, Number>(op, params, pcout);
        return out;
      }
    throw std::runtime_error("Preconditioner <" + type + "> is not known!");
  }
} // namespace dasm
