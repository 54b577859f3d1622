// C++ mirror of the reference's operator interface (include/operator.h of the reference) over libdasm's C ABI.
//
//   LaplaceOperatorBase        operator.h:32-60
//   LaplaceOperatorMatrixFree  operator.h:266-1628  (AdditionalData 285-295, vmult 1353-1430,
//                              compute_inverse_diagonal 1512-1524, el/Tvmult throw like 1432-1463)
//
// The reference's operators are templated on deal.II's MatrixFree; here the mesh/DoF infrastructure is the
// library's structured-mesh object (dasm::Mesh) and vectors are device vectors (dasm::Vector<Number>).
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../dasm.h"

#include "grid_generator.h"

namespace dasm
{
  inline void
  check(int rc)
  {
    if (rc != 0)
      throw std::runtime_error(dasm_last_error()); // mirrors AssertThrow
  }

  class Context
  {
  public:
    dasm_ctx *h = nullptr;
    explicit Context(int device = 0) { check(dasm_ctx_create(device, &h)); }
    ~Context() { dasm_ctx_destroy(h); }
    Context(const Context &) = delete;
    void sync() { check(dasm_ctx_sync(h)); }
  };

  class Mesh
  {
  public:
    dasm_mesh *h = nullptr;
    Context &  ctx;
    int        n_cells_dir[3];
    // structured hyper-rectangle; map_kind see dasm_map_kind
    Mesh(Context &ctx, const int n_cells[3], const int periodic[3], bool dirichlet, const double length[3], int map_kind = 0,
         const double *map_params = nullptr, const int *partition = nullptr, int rank = 0)
      : ctx(ctx)
    {
      const int    one[3]  = {1, 1, 1};
      const double zero[4] = {0, 0, 0, 0};
      for (int d = 0; d < 3; ++d)
        n_cells_dir[d] = n_cells[d];
      check(dasm_mesh_create_structured(ctx.h, n_cells, periodic, dirichlet ? 1 : 0, length, map_kind, map_params ? map_params : zero,
                                        partition ? partition : one, rank, &h));
    }
    ~Mesh() { dasm_mesh_destroy(h); }
    Mesh(const Mesh &) = delete;
    long long n_cells() const { return dasm_mesh_n_cells(h); }
  };

  template <typename Number>
  struct NumberType;
  template <>
  struct NumberType<double>
  {
    static constexpr int value = DASM_F64;
  };
  template <>
  struct NumberType<float>
  {
    static constexpr int value = DASM_F32;
  };

  // device vector with the locally-owned + ghost layout of LinearAlgebra::distributed::Vector
  template <typename Number>
  class Vector
  {
  public:
    using value_type = Number;
    Vector() = default;
    ~Vector() { reset(); }
    Vector(const Vector &) = delete;
    void
    reinit(dasm_op *op_)
    {
      reset();
      op = op_;
      check(dasm_op_vec_alloc(op, &ptr));
      n_owned = dasm_op_n_dofs(op);
    }
    void
    reset()
    {
      if (ptr)
        dasm_op_vec_free(op, ptr);
      ptr = nullptr;
    }
    Vector &
    operator=(const Number v)
    {
      std::vector<double> tmp(n_owned, (double)v);
      check(dasm_op_vec_upload(op, ptr, tmp.data()));
      return *this;
    }
    void upload(const std::vector<double> &host) { check(dasm_op_vec_upload(op, ptr, host.data())); }
    std::vector<double>
    download() const
    {
      std::vector<double> out(n_owned);
      check(dasm_op_vec_download(op, out.data(), ptr));
      return out;
    }
    long long locally_owned_size() const { return n_owned; }
    void *    data() { return ptr; }
    const void *data() const { return ptr; }

  private:
    dasm_op * op      = nullptr;
    void *    ptr     = nullptr;
    long long n_owned = 0;
  };

  namespace SymmetryType
  {
    enum SymmetryType
    {
      symmetric,
      non_symmetric,
      undefined
    };
  }

  // Utilities::MPI::Partitioner as far as the hot path uses it (include/matrix_free.h:154-213, include/operator.h:780-849): the layout of
  // a distributed vector.  libdasm gives the operator and every preconditioner built on it ONE layout that already contains the
  // enlarged ghost set of overlapping patches, so the partitioner of a preconditioner is always the one of its operator.
  struct Partitioner
  {
    const void *owner = nullptr; // the operator that defines the layout
    long long   n_locally_owned = 0, n_ghost = 0, n_import = 0, vec_size = 0;
    long long   locally_owned_size() const { return n_locally_owned; }
    long long   n_ghost_indices() const { return n_ghost; }
    long long   n_import_indices() const { return n_import; }
  };

  // operator.h:32-60
  template <int dim, typename VectorType>
  class LaplaceOperatorBase
  {
  public:
    virtual ~LaplaceOperatorBase()                                    = default;
    virtual void vmult(VectorType &dst, const VectorType &src) const  = 0;
    virtual void initialize_dof_vector(VectorType &vec) const         = 0;
    virtual SymmetryType::SymmetryType is_symmetric() const { return SymmetryType::symmetric; }
  };

  template <int dim, typename Number>
  class LaplaceOperatorMatrixFree : public LaplaceOperatorBase<dim, Vector<Number>>
  {
    static_assert(dim == 3, "libdasm builds the 3-D path");

  public:
    static const int dimension = dim;
    using value_type           = Number;
    using vector_type          = Vector<Number>;
    using VectorType           = vector_type;

    struct AdditionalData // operator.h:285-295
    {
      AdditionalData(const bool compress_indices = false, const std::string mapping_type = "")
        : compress_indices(compress_indices)
        , mapping_type(mapping_type)
      {}
      bool        compress_indices;
      std::string mapping_type;
    };

    LaplaceOperatorMatrixFree(Mesh &mesh, const unsigned int fe_degree, const AdditionalData &ad = AdditionalData())
      : mesh(&mesh)
      , fe_degree(fe_degree)
    {
      check(dasm_op_create(mesh.h, (int)fe_degree, NumberType<Number>::value, ad.mapping_type.c_str(), ad.compress_indices ? 1 : 0, &h));
    }
    // operator on an unstructured all-hex mesh given by arrays (grid_generator.h; dasm_op_create_unstructured): the ball of
    // element_centered_preconditioners_01.cc:398-402.  Homogeneous Dirichlet conditions on the whole boundary (:404-413).
    LaplaceOperatorMatrixFree(Context &ctx, const UnstructuredMesh &umesh, const unsigned int fe_degree, const AdditionalData &ad = AdditionalData())
      : mesh(nullptr)
      , fe_degree(fe_degree)
    {
      check(dasm_op_create_unstructured(ctx.h, (int)fe_degree, NumberType<Number>::value, ad.mapping_type.c_str(), umesh.n_vertices(),
                                        umesh.vertices.data(), umesh.n_cells(), umesh.cells.data(),
                                        umesh.support.empty() ? nullptr : umesh.support.data(), 1, &h));
    }
    ~LaplaceOperatorMatrixFree() override { dasm_op_destroy(h); }
    long long n_cells() const { return dasm_op_n_cells(h); }

    virtual bool uses_compressed_indices() const { return dasm_op_uses_compressed_indices(h) != 0; }
    static constexpr bool is_matrix_free() { return true; }

    void initialize_dof_vector(VectorType &vec) const override { vec.reinit(h); }

    void vmult(VectorType &dst, const VectorType &src) const override { check(dasm_op_vmult(h, dst.data(), src.data())); }

    // vmult with the enumerated pre/post operations (see dasm_hook_kind)
    void
    vmult(VectorType &dst, const VectorType &src, const dasm_hook &pre, const dasm_hook &post) const
    {
      check(dasm_op_vmult_hooks(h, dst.data(), src.data(), &pre, &post));
    }

    // the reference's signature with host lambdas over DoF ranges (operator.h:1367-1373): a CUDA kernel cannot
    // call them; empty functions select the plain product, anything else is rejected
    void
    vmult(VectorType &dst, const VectorType &src, const std::function<void(const unsigned int, const unsigned int)> &pre,
          const std::function<void(const unsigned int, const unsigned int)> &post) const
    {
      if (post)
        throw std::runtime_error("ExcNotImplemented: arbitrary host lambdas cannot run inside the device cell loop; "
                                 "use the dasm_hook overload (DASM_HOOK_RESIDUAL / DASM_HOOK_CHEB_UPDATE / DASM_HOOK_SCALE)");
      (void)pre; // the zeroing pre-operation is implied
      vmult(dst, src);
    }

    // rhs(vec, func), operator.h:53-56: constant functions only (the reference's drivers use f = 1)
    void
    rhs(VectorType &vec, const double value = 1.0) const
    {
      if (vec.data() == nullptr)
        vec.reinit(h);
      check(dasm_op_rhs_constant(h, vec.data(), value));
    }
    // get_constraints(), operator.h:46-51: the list of constrained (homogeneous Dirichlet) DoFs
    std::vector<std::uint32_t>
    get_constraints() const
    {
      std::vector<std::uint32_t> out((std::size_t)dasm_op_constrained_dofs(h, nullptr));
      if (!out.empty())
        dasm_op_constrained_dofs(h, out.data());
      return out;
    }

    void Tvmult(VectorType &, const VectorType &) const { throw std::runtime_error("ExcNotImplemented"); }
    Number el(unsigned int, unsigned int) const { throw std::runtime_error("ExcNotImplemented"); }
    unsigned long long m() const { return (unsigned long long)dasm_op_n_global_dofs(h); }
    void compute_inverse_diagonal(VectorType &diagonal) const
    {
      if (diagonal.data() == nullptr)
        diagonal.reinit(h);
      check(dasm_op_inverse_diagonal(h, diagonal.data()));
    }
    // get_partitioner / set_partitioner (operator.h:780-849): the reference re-targets the operator's vector access to the
    // preconditioner's larger ghost layout; here both already share it, so set_partitioner only checks that the layout is this one
    std::shared_ptr<const Partitioner>
    get_partitioner() const
    {
      auto p             = std::make_shared<Partitioner>();
      p->owner           = h;
      p->n_locally_owned = dasm_op_n_dofs(h);
      p->vec_size        = dasm_op_vec_size(h);
      p->n_ghost         = p->vec_size - p->n_locally_owned;
      p->n_import        = dasm_op_n_import(h);
      return p;
    }
    void
    set_partitioner(const std::shared_ptr<const Partitioner> &vector_partitioner) const
    {
      if (!vector_partitioner || vector_partitioner->owner != h)
        throw std::runtime_error("ExcNotImplemented: the partitioner must be the one of a preconditioner built on this operator");
    }
    unsigned int get_fe_degree() const { return fe_degree; }
    Mesh &
    get_mesh() const
    {
      if (mesh == nullptr)
        throw std::runtime_error("the operator lives on an unstructured mesh");
      return *mesh;
    }
    dasm_op *    handle() const { return h; }

  private:
    Mesh *       mesh;
    unsigned int fe_degree;
    dasm_op *    h = nullptr;
  };
} // namespace dasm
