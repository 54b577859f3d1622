// C++ mirror of the reference's preconditioner interfaces over libdasm's C ABI.
//   PreconditionerBase        include/preconditioners.h:725-742 (vmult, step throws by default)
//   Restrictors::WeightingType include/restrictors.h:8-15
//   ASPoissonPreconditioner   include/matrix_free.h:63-1568
//   PreconditionChebyshev     deal.II class configured in include/precondition.templates.h:89-158
#pragma once
#include "operator.h"

namespace dasm
{
  namespace Restrictors
  {
    enum class WeightingType
    {
      none = DASM_WEIGHT_NONE,
      pre  = DASM_WEIGHT_PRE,
      post = DASM_WEIGHT_POST,
      ras  = DASM_WEIGHT_RAS,
      symm = DASM_WEIGHT_SYMM
    };
  }

  template <typename VectorType>
  class PreconditionerBase
  {
  public:
    virtual ~PreconditionerBase()                                    = default;
    virtual void vmult(VectorType &dst, const VectorType &src) const = 0;
    virtual void step(VectorType &, const VectorType &) const { throw std::runtime_error("ExcNotImplemented"); }
  };

  template <int dim, typename Number>
  class ASPoissonPreconditioner : public PreconditionerBase<Vector<Number>>
  {
  public:
    using VectorType = Vector<Number>;

    // argument order of the reference constructor (matrix_free.h:73-86) minus the deal.II objects
    ASPoissonPreconditioner(const LaplaceOperatorMatrixFree<dim, Number> &op, const unsigned int n_overlap,
                            const unsigned int               sub_mesh_approximation,
                            const Restrictors::WeightingType weight_type         = Restrictors::WeightingType::post,
                            const std::string                weight_local_global = "global", const bool overlap_pre_post = true,
                            const bool element_centric = true)
    {
      int seq;
      if (weight_local_global == "global")
        seq = DASM_WSEQ_GLOBAL;
      else if (weight_local_global == "local")
        seq = DASM_WSEQ_LOCAL;
      else if (weight_local_global == "dg" || weight_local_global == "DG")
        seq = DASM_WSEQ_DG;
      else if (weight_local_global == "compressed")
        seq = DASM_WSEQ_COMPRESSED;
      else
        throw std::runtime_error("weight sequence <" + weight_local_global + "> is not known!");
      check(dasm_fdm_create(op.handle(), (int)n_overlap, (int)sub_mesh_approximation, (int)weight_type, seq, overlap_pre_post ? 1 : 0,
                            element_centric ? 1 : 0, &h));
      partitioner = op.get_partitioner();
    }
    // matrix_free.h:994-999: the (enlarged) vector layout of the preconditioner = the operator's layout in libdasm
    const std::shared_ptr<const Partitioner> &get_partitioner() const { return partitioner; }
    ~ASPoissonPreconditioner() override { dasm_fdm_destroy(h); }

    SymmetryType::SymmetryType
    is_symmetric() const
    {
      return dasm_fdm_is_symmetric(h) ? SymmetryType::symmetric : SymmetryType::non_symmetric;
    }
    void vmult(VectorType &dst, const VectorType &src) const override { check(dasm_fdm_vmult(h, dst.data(), src.data())); }
    void
    vmult(VectorType &dst, const VectorType &src, const dasm_hook &pre, const dasm_hook &post) const
    {
      check(dasm_fdm_vmult_hooks(h, dst.data(), src.data(), &pre, &post));
    }
    std::size_t  memory_consumption() const { return (std::size_t)dasm_fdm_memory_consumption(h); }
    unsigned int n_fdm_instances() const { return (unsigned int)dasm_fdm_n_instances(h); }
    // diagnostics (no counterpart in the reference): bricks processed by the warp-specialised kernel
    long long    n_fast_bricks() const { return dasm_fdm_n_fast_bricks(h); }
    dasm_fdm *   handle() const { return h; }

  private:
    dasm_fdm *                         h = nullptr;
    std::shared_ptr<const Partitioner> partitioner;
  };

  template <int dim, typename Number>
  class PreconditionChebyshev : public PreconditionerBase<Vector<Number>>
  {
  public:
    using VectorType = Vector<Number>;
    struct AdditionalData
    {
      unsigned int degree              = 3;
      double       smoothing_range     = 20.;
      unsigned int eig_cg_n_iterations = 40;
      int          eigenvalue_algorithm = DASM_EV_DEFAULT;
      int          polynomial_type      = DASM_POLY_FIRST_KIND;
      unsigned int optimize             = 2;
    };
    struct EigenvalueInformation
    {
      double min_eigenvalue_estimate, max_eigenvalue_estimate;
    };

    // fdm == nullptr: point Jacobi (DiagonalMatrixPrePost, preconditioners.h:951-997)
    PreconditionChebyshev(const LaplaceOperatorMatrixFree<dim, Number> &op, const ASPoissonPreconditioner<dim, Number> *fdm,
                          const AdditionalData &ad)
    {
      check(dasm_cheb_create(op.handle(), fdm ? fdm->handle() : nullptr, (int)ad.degree, ad.smoothing_range, ad.polynomial_type,
                             ad.eigenvalue_algorithm, (int)ad.optimize, (int)ad.eig_cg_n_iterations, &h));
    }
    ~PreconditionChebyshev() override { dasm_cheb_destroy(h); }
    EigenvalueInformation
    estimate_eigenvalues(const VectorType &) const
    {
      EigenvalueInformation info;
      check(dasm_cheb_estimate_eigenvalues(h, &info.min_eigenvalue_estimate, &info.max_eigenvalue_estimate));
      return info;
    }
    void vmult(VectorType &dst, const VectorType &src) const override { check(dasm_cheb_vmult(h, dst.data(), src.data())); }
    void step(VectorType &dst, const VectorType &src) const override { check(dasm_cheb_step(h, dst.data(), src.data())); }
    dasm_cheb *handle() const { return h; }

  private:
    dasm_cheb *h = nullptr;
  };
} // namespace dasm
