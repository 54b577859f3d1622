// C++ mirror of the reference's restrictor / exact-block additive Schwarz interface over libdasm's C ABI.
//   Restrictors::WeightingType                  include/restrictors.h:8-15      (enum in preconditioners.h of this mirror)
//   Restrictors::ElementCenteredRestrictor      include/restrictors.h:17-378    (AdditionalData 24-46: n_overlap, weighting_type,
//                                                                                type "element" | "vertex"; read / add with weights)
//   RestrictedPreconditioner                    include/preconditioners.h:744-813 (vmult 775-808)
//   RestrictedMatrixView + gauss_jordan blocks  include/preconditioners.h:528-605
//
// The patches (index lists, weights) live on the device as part of an ASPoissonPreconditioner; the restrictor object exposes them,
// and RestrictedPreconditioner replaces the fast-diagonalisation block inverse by the exact inverse of the restricted operator matrix.
#pragma once
#include <cstdint>

#include "preconditioners.h"

namespace dasm
{
  namespace Restrictors
  {
    template <int dim, typename Number>
    class ElementCenteredRestrictor
    {
    public:
      struct AdditionalData
      {
        AdditionalData(const unsigned int n_overlap = 1, const WeightingType weighting_type = WeightingType::none,
                       const std::string type = "element")
          : n_overlap(n_overlap)
          , weighting_type(weighting_type)
          , type(type)
        {}
        unsigned int  n_overlap;
        WeightingType weighting_type;
        std::string   type; // "element" (cell-centred patches) | "vertex" (vertex-star patches)
      };

      ElementCenteredRestrictor(const LaplaceOperatorMatrixFree<dim, Number> &op, const AdditionalData &ad = AdditionalData())
        : op(op)
        , layout(op, ad.n_overlap, dim, ad.weighting_type, ad.n_overlap > 1 ? "global" : "compressed", true, check_type(ad.type))
      {}

      unsigned int n_blocks() const { return (unsigned int)dasm_mesh_n_cells(dasm_op_mesh(op.handle())); }
      unsigned int n_entries_per_block() const
      {
        const unsigned int m = (unsigned int)dasm_fdm_patch_size_1d(layout.handle());
        return m * m * m;
      }
      // indices (0xFFFFFFFF = entry not part of the patch) and weights of all patches, copied to the host
      void
      get_indices_and_weights(std::vector<std::uint32_t> &indices, std::vector<double> &weights, bool &weights_pre, bool &weights_post) const
      {
        const std::size_t n = (std::size_t)n_blocks() * n_entries_per_block();
        indices.resize(n);
        weights.resize(n);
        int wp = 0, wq = 0;
        check(dasm_fdm_patches_host(layout.handle(), indices.data(), weights.data(), &wp, &wq));
        weights_pre  = wp != 0;
        weights_post = wq != 0;
      }
      const ASPoissonPreconditioner<dim, Number> &get_layout() const { return layout; }

    private:
      static bool
      check_type(const std::string &type)
      {
        if (type != "element" && type != "vertex")
          throw std::runtime_error("Restrictor type <" + type + "> is not known!");
        return type == "element";
      }
      const LaplaceOperatorMatrixFree<dim, Number> &op;
      ASPoissonPreconditioner<dim, Number>          layout;
    };
  } // namespace Restrictors

  // RestrictedPreconditioner<VectorType, InverseMatrixType = exact block inverse, RestrictorType = ElementCenteredRestrictor>
  template <int dim, typename Number>
  class RestrictedPreconditioner : public PreconditionerBase<Vector<Number>>
  {
  public:
    using VectorType = Vector<Number>;
    explicit RestrictedPreconditioner(const std::shared_ptr<const Restrictors::ElementCenteredRestrictor<dim, Number>> &restrictor)
      : restrictor(restrictor)
    {
      check(dasm_asm_create(restrictor->get_layout().handle(), &h));
    }
    ~RestrictedPreconditioner() override { dasm_asm_destroy(h); }
    void        vmult(VectorType &dst, const VectorType &src) const override { check(dasm_asm_vmult(h, dst.data(), src.data())); }
    std::size_t memory_consumption() const { return (std::size_t)dasm_asm_memory_consumption(h); }
    dasm_asm *  handle() const { return h; }

  private:
    std::shared_ptr<const Restrictors::ElementCenteredRestrictor<dim, Number>> restrictor;
    dasm_asm *                                                                 h = nullptr;
  };
} // namespace dasm
