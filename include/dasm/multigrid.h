// C++ mirror of the reference's multigrid preconditioner over libdasm's C ABI.
//   PreconditionerGMG                 include/multigrid.h:109-537 (constructor, do_update, vmult)
//   MGCoarseGridApplyPreconditioner   include/multigrid.h:17-108  (the coarse level applies its preconditioner once)
//   MGTwoLevelTransfer                deal.II, set up in include/multigrid.h:338-349
//   WrapperForGMG                     include/precondition.h:24-78 (smoother = any PreconditionerBase with vmult / step)
//
// The reference builds the level operators, constraints and DoF handlers from deal.II triangulations
// (element_centered_preconditioners_01.cc:540-740); here a level is a dasm::LaplaceOperatorMatrixFree on a dasm::Mesh with twice the
// cells of the next coarser level ("mg type" h) or the same mesh and a lower degree ("mg type" p), and every level smoother is a
// dasm::PreconditionChebyshev (around FDM or point Jacobi), the smoother type of every multigrid configuration in experiments/.
#pragma once
#include "preconditioners.h"

namespace dasm
{
  template <int dim, typename Number>
  class MGTwoLevelTransfer
  {
  public:
    using VectorType = Vector<Number>;
    MGTwoLevelTransfer(const LaplaceOperatorMatrixFree<dim, Number> &fine, const LaplaceOperatorMatrixFree<dim, Number> &coarse)
    {
      check(dasm_transfer_create(fine.handle(), coarse.handle(), &h));
    }
    // operators on unstructured meshes, geometric transfer: parent[fine cell] = coarse cell | child position << 28
    // (GridGenerator::ball_parents)
    MGTwoLevelTransfer(const LaplaceOperatorMatrixFree<dim, Number> &fine, const LaplaceOperatorMatrixFree<dim, Number> &coarse,
                       const std::vector<std::uint32_t> &parent)
    {
      check(dasm_transfer_create_unstructured(fine.handle(), coarse.handle(), parent.empty() ? nullptr : parent.data(), &h));
    }
    ~MGTwoLevelTransfer() { dasm_transfer_destroy(h); }
    dasm_transfer *handle() const { return h; }
    MGTwoLevelTransfer(const MGTwoLevelTransfer &) = delete;
    void prolongate_and_add(VectorType &dst, const VectorType &src) const { check(dasm_transfer_prolongate_and_add(h, dst.data(), src.data())); }
    void restrict_and_add(VectorType &dst, const VectorType &src) const { check(dasm_transfer_restrict_and_add(h, dst.data(), src.data())); }

  private:
    dasm_transfer *h = nullptr;
  };

  // LevelNumber: number type of the level operators (float in the reference's matrix-free trait,
  // element_centered_preconditioners_01.cc:787-792); OuterNumber: number type of the vectors vmult is called with.
  template <int dim, typename LevelNumber, typename OuterNumber = double>
  class PreconditionerGMG : public PreconditionerBase<Vector<OuterNumber>>
  {
  public:
    using LevelMatrixType = LaplaceOperatorMatrixFree<dim, LevelNumber>;
    using SmootherType    = PreconditionChebyshev<dim, LevelNumber>;
    using VectorType      = Vector<LevelNumber>;
    using VectorTypeOuter = Vector<OuterNumber>;

    // mg_operators[0] is the coarsest level; mg_smoothers[0] the coarse-grid solver (create_mg_coarse_grid_solver), mg_smoothers[l]
    // the level smoother (create_mg_level_smoother) of level l
    PreconditionerGMG(const std::vector<std::shared_ptr<LevelMatrixType>> &mg_operators,
                      const std::vector<std::shared_ptr<SmootherType>> &   mg_smoothers, const bool use_one_sided_v_cycle = false)
      : mg_operators(mg_operators)
      , mg_smoothers(mg_smoothers)
    {
      if (mg_operators.size() != mg_smoothers.size() || mg_operators.empty())
        throw std::runtime_error("ExcDimensionMismatch: one smoother per multigrid level is needed");
      std::vector<dasm_op *>   ops;
      std::vector<dasm_cheb *> sms;
      for (const auto &o : mg_operators)
        ops.push_back(o->handle());
      for (const auto &s : mg_smoothers)
        sms.push_back(s->handle());
      check(dasm_mg_create((int)ops.size(), ops.data(), sms.data(), use_one_sided_v_cycle ? 1 : 0, &h));
    }
    // same with caller-built transfers (entry l between the levels l and l - 1, empty pointers are built by the library): the geometric
    // levels of an unstructured mesh
    PreconditionerGMG(const std::vector<std::shared_ptr<LevelMatrixType>> &mg_operators,
                      const std::vector<std::shared_ptr<SmootherType>> &   mg_smoothers,
                      const std::vector<std::shared_ptr<MGTwoLevelTransfer<dim, LevelNumber>>> &transfers, const bool use_one_sided_v_cycle = false)
      : mg_operators(mg_operators)
      , mg_smoothers(mg_smoothers)
      , mg_transfers(transfers)
    {
      if (mg_operators.size() != mg_smoothers.size() || mg_operators.size() != transfers.size() || mg_operators.empty())
        throw std::runtime_error("ExcDimensionMismatch: one smoother and one transfer entry per multigrid level are needed");
      std::vector<dasm_op *>       ops;
      std::vector<dasm_cheb *>     sms;
      std::vector<dasm_transfer *> trs;
      for (const auto &o : mg_operators)
        ops.push_back(o->handle());
      for (const auto &s : mg_smoothers)
        sms.push_back(s->handle());
      for (const auto &t : transfers)
        trs.push_back(t ? t->handle() : nullptr);
      check(dasm_mg_create_with_transfers((int)ops.size(), ops.data(), sms.data(), trs.data(), use_one_sided_v_cycle ? 1 : 0, &h));
    }
    ~PreconditionerGMG() override { dasm_mg_destroy(h); }

    void
    vmult(VectorTypeOuter &dst, const VectorTypeOuter &src) const override
    {
      ++all_mg_counter;
      check(dasm_mg_vmult_outer(h, dst.data(), src.data(), NumberType<OuterNumber>::value));
    }
    unsigned int n_calls() const { return all_mg_counter; } // print_timings(): "#N of calls of multigrid"
    void         clear_timings() const { all_mg_counter = 0; }
    dasm_mg *    handle() const { return h; }

  private:
    std::vector<std::shared_ptr<LevelMatrixType>> mg_operators;
    std::vector<std::shared_ptr<SmootherType>>    mg_smoothers;
    std::vector<std::shared_ptr<MGTwoLevelTransfer<dim, LevelNumber>>> mg_transfers;
    dasm_mg *                                     h = nullptr;
    mutable unsigned int                          all_mg_counter = 0;
  };
} // namespace dasm
