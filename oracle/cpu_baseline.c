/* CPU baseline of the smoother hot path on a periodic Cartesian mesh - TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * Nothing in the product path may link or call this file.  It stands in for the reference's MPI CPU path (deal.II
 * cannot be built here) in bench.py's cpu_baseline / --impl reference legs and is checked against the
 * numpy oracle in tests/test_oracle_c.py.  It is written the way BASELINE.md section 3 describes the reference's path:
 *
 *   - degree as a compile-time constant (the switch at the bottom instantiates n = 2..9), like FEEvaluation<dim, degree>;
 *   - SIMD across cells: 8 cells (one x-run) per batch, every 1-D contraction vectorises over the 8 lanes
 *     (VectorizedArray<double> on AVX-512; include/operator.h:273-278);
 *   - Cartesian fast path: the cell matrix in Kronecker form K(x)M(x)M + M(x)K(x)M + M(x)M(x)K with scaled 1-D
 *     matrices, no per-quadrature-point coefficients (on an affine cell this is exactly the reference's
 *     evaluate(gradients) -> J^-1 J^-T JxW -> integrate(gradients), include/operator.h:866-875);
 *   - FDM apply_inverse (S (x) S (x) S) Lambda^-1 (S (x) S (x) S)^T with the weights applied in the gathered cell
 *     vector (include/matrix_free.h:1023-1062, 1366-1488);
 *   - the Chebyshev vector updates fused into the cell loops through pre / post operations on DoF ranges that run
 *     when a range is touched first / last (include/operator.h:1367-1430, include/matrix_free.h:960-986): cells are
 *     split into contiguous z-slabs, one per thread and colour; the planes interior to a slab are finished right
 *     after the slab, the planes between two slabs after the second colour.  One Chebyshev term then makes 7 vector
 *     passes (read x, read b, write t | read t, read x, read x_old, write x_new), as in SURVEY.md 8(d);
 *   - one contiguous cell range per thread (the role of the MPI partition), OpenMP threads = host cores.
 *
 * DoFs are numbered lexicographically (x fastest) on the periodic (k nx) x (k ny) x (k nz) lattice.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LANES 8
#define INLINE static inline __attribute__((always_inline))
typedef double v8d __attribute__((vector_size(64), aligned(64))); /* one value of 8 cells (AVX-512: one register) */

/* the launcher may have set OMP_NUM_THREADS=1 (torch.distributed.run does): the baseline uses all host cores it is given */
void cpu_baseline_set_threads(int n)
{
#ifdef _OPENMP
  if (n > 0)
    omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int cpu_baseline_max_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* out (+)= M applied along direction dir of the [n][n][n] tensor of 8-cell vectors; M row-major [o][i] */
INLINE void sweep(const int n, const double *restrict M, const v8d *restrict in, v8d *restrict out, const int dir, const int add)
{
  const int s = dir == 0 ? 1 : (dir == 1 ? n : n * n);
  for (int b = 0; b < n; ++b)
    for (int a = 0; a < n; ++a)
      {
        const int base = dir == 0 ? (b * n + a) * n : (dir == 1 ? b * n * n + a : b * n + a);
        v8d       v[9];
        for (int i = 0; i < n; ++i)
          v[i] = in[base + i * s];
        for (int o = 0; o < n; ++o)
          {
            v8d acc = M[o * n] * v[0];
            for (int i = 1; i < n; ++i)
              acc += M[o * n + i] * v[i];
            if (add)
              out[base + o * s] += acc;
            else
              out[base + o * s] = acc;
          }
      }
}

typedef struct
{
  int           k, n, nc[3], nd[3]; /* degree, n = k + 1, cells and DoFs per direction */
  const double *M, *K[3];           /* 1-D mass matrix and the three scaled stiffness matrices (n x n) */
  const double *Mh[3];              /* mass matrices scaled per direction */
  const double *S[3], *ST[3];       /* FDM eigenvectors per direction [i][a] and transposed */
  const double *ilam;               /* 1 / (lx[a] + ly[b] + lz[c]) at (c n + b) n + a */
  const double *w;                  /* weights of the gathered cell vector [n^3] or NULL */
  int           w_pre, w_post;
} Problem;

INLINE size_t dof_row(const Problem *p, const int cy, const int cz, const int j, const int l)
{
  const int Y = (p->k * cy + j) % p->nd[1], Z = (p->k * cz + l) % p->nd[2];
  return ((size_t)Z * p->nd[1] + Y) * p->nd[0];
}

/* gather the batch of cells (cx0 .. cx0 + nb - 1, cy, cz): lane v = cell cx0 + v.  Full batches away from the periodic wrap
 * use the AVX-512 gather / scatter instructions with the constant index vector k v (VectorizedArray::gather / scatter). */
#ifdef __AVX512F__
#include <immintrin.h>
#endif

INLINE void gather(const Problem *p, const int n, const int cx0, const int nb, const int cy, const int cz, const double *restrict x,
                   v8d *restrict uv)
{
  double *restrict u = (double *)uv;
  const int k = n - 1, NX = p->nd[0];
#ifdef __AVX512F__
  if (nb == LANES && k * (cx0 + LANES) < NX)
    {
      const __m256i vidx = _mm256_setr_epi32(0, k, 2 * k, 3 * k, 4 * k, 5 * k, 6 * k, 7 * k);
      for (int l = 0; l < n; ++l)
        for (int j = 0; j < n; ++j)
          {
            const double *row = x + dof_row(p, cy, cz, j, l) + k * cx0;
            for (int i = 0; i < n; ++i)
              _mm512_store_pd(u + ((l * n + j) * n + i) * LANES, _mm512_i32gather_pd(vidx, row + i, 8));
          }
      return;
    }
#endif
  for (int l = 0; l < n; ++l)
    for (int j = 0; j < n; ++j)
      {
        const double *row = x + dof_row(p, cy, cz, j, l);
        for (int i = 0; i < n; ++i)
          for (int v = 0; v < LANES; ++v)
            {
              const int X = (k * (cx0 + (v < nb ? v : 0)) + i) % NX;
              u[((l * n + j) * n + i) * LANES + v] = row[X];
            }
      }
}

INLINE void scatter_add(const Problem *p, const int n, const int cx0, const int nb, const int cy, const int cz, const v8d *restrict uv,
                        double *restrict y)
{
  const double *restrict u = (const double *)uv;
  const int k = n - 1, NX = p->nd[0];
#ifdef __AVX512F__
  if (nb == LANES && k * (cx0 + LANES) < NX)
    {
      /* the node x = k of lane v is the node x = 0 of lane v + 1: merged by a lane shift, then every address is written once */
      const __m256i vidx  = _mm256_setr_epi32(0, k, 2 * k, 3 * k, 4 * k, 5 * k, 6 * k, 7 * k);
      const __m512i shift = _mm512_setr_epi64(0, 0, 1, 2, 3, 4, 5, 6);
      for (int l = 0; l < n; ++l)
        for (int j = 0; j < n; ++j)
          {
            double *      row = y + dof_row(p, cy, cz, j, l) + k * cx0;
            const double *ur  = u + ((l * n + j) * n) * LANES;
            const __m512d uk  = _mm512_load_pd(ur + k * LANES);
            __m512d       u0  = _mm512_add_pd(_mm512_load_pd(ur), _mm512_maskz_permutexvar_pd(0xFE, shift, uk));
            _mm512_i32scatter_pd(row, vidx, _mm512_add_pd(_mm512_i32gather_pd(vidx, row, 8), u0), 8);
            for (int i = 1; i < k; ++i)
              _mm512_i32scatter_pd(row + i, vidx, _mm512_add_pd(_mm512_i32gather_pd(vidx, row + i, 8), _mm512_load_pd(ur + i * LANES)), 8);
            row[k * LANES] += ur[k * LANES + LANES - 1];
          }
      return;
    }
#endif
  for (int l = 0; l < n; ++l)
    for (int j = 0; j < n; ++j)
      {
        double *row = y + dof_row(p, cy, cz, j, l);
        for (int v = 0; v < nb; ++v) /* lane by lane: neighbouring cells of the batch share DoFs */
          for (int i = 0; i < n; ++i)
            row[(k * (cx0 + v) + i) % NX] += u[((l * n + j) * n + i) * LANES + v];
      }
}

/* y += A x on one batch: 7 sweeps */
INLINE void laplace_batch(const Problem *p, const int n, v8d *restrict u, v8d *restrict t0, v8d *restrict t1, v8d *restrict t2)
{
  /* t0 = Mx u, t1 = Kx u */
  sweep(n, p->Mh[0], u, t0, 0, 0);
  sweep(n, p->K[0], u, t1, 0, 0);
  /* t2 = My t1 (-> Kx My), u = My t0 (-> Mx My), t1 = Ky t0 (-> Mx Ky) */
  sweep(n, p->Mh[1], t1, t2, 1, 0);
  sweep(n, p->Mh[1], t0, u, 1, 0);
  sweep(n, p->K[1], t0, t1, 1, 0);
  for (int i = 0; i < n * n * n; ++i)
    t2[i] += t1[i]; /* (Kx My + Mx Ky) u */
  /* result = Mz t2 + Kz u */
  sweep(n, p->Mh[2], t2, t0, 2, 0);
  sweep(n, p->K[2], u, t0, 2, 1);
}

INLINE void fdm_batch(const Problem *p, const int n, v8d *restrict u, v8d *restrict t0)
{
  const int n3 = n * n * n;
  if (p->w && p->w_pre)
    for (int i = 0; i < n3; ++i)
      u[i] *= p->w[i];
  sweep(n, p->ST[0], u, t0, 0, 0);
  sweep(n, p->ST[1], t0, u, 1, 0);
  sweep(n, p->ST[2], u, t0, 2, 0);
  for (int i = 0; i < n3; ++i)
    t0[i] *= p->ilam[i];
  sweep(n, p->S[0], t0, u, 0, 0);
  sweep(n, p->S[1], u, t0, 1, 0);
  sweep(n, p->S[2], t0, u, 2, 0);
  if (p->w && p->w_post)
    for (int i = 0; i < n3; ++i)
      u[i] *= p->w[i];
}

/* pre / post operations on the DoF planes [Z0, Z1) (periodic in Z) */
typedef struct
{
  int           kind;   /* 0: A-sweep  pre t = 0,  post t = b - t;   1: P-sweep  pre z = 0,  post xn = x + f1 (x - xo) + f2 z */
  double        f1, f2;
  double *      dst;    /* t or z (accumulated) */
  const double *b;      /* A-sweep: right-hand side */
  const double *x, *xo; /* P-sweep: current / previous iterate (xo may be NULL) */
  double *      xn;     /* P-sweep: new iterate */
} Hooks;

INLINE void pre_planes(const Problem *p, const Hooks *h, int Z0, int Z1)
{
  const size_t plane = (size_t)p->nd[0] * p->nd[1];
  for (int Z = Z0; Z < Z1; ++Z)
    memset(h->dst + (size_t)((Z % p->nd[2] + p->nd[2]) % p->nd[2]) * plane, 0, plane * sizeof(double));
}

INLINE void post_planes(const Problem *p, const Hooks *h, int Z0, int Z1)
{
  const size_t plane = (size_t)p->nd[0] * p->nd[1];
  for (int Z = Z0; Z < Z1; ++Z)
    {
      const size_t o = (size_t)((Z % p->nd[2] + p->nd[2]) % p->nd[2]) * plane;
      if (h->kind == 0)
        for (size_t i = 0; i < plane; ++i)
          h->dst[o + i] = h->b[o + i] - h->dst[o + i];
      else if (h->xo)
        for (size_t i = 0; i < plane; ++i)
          h->xn[o + i] = h->x[o + i] + h->f1 * (h->x[o + i] - h->xo[o + i]) + h->f2 * h->dst[o + i];
      else
        for (size_t i = 0; i < plane; ++i)
          h->xn[o + i] = h->x[o + i] + h->f1 * h->x[o + i] + h->f2 * h->dst[o + i];
    }
}

/* one cell loop (A-sweep: op 0, P-sweep: op 1) over all cells with fused hooks */
INLINE void cell_loop(const Problem *p, const int n, const int op, const double *restrict src, const Hooks *h)
{
  const int k = n - 1, n3 = n * n * n;
  /* an even number of z-slabs, two per thread: colour 0 = even slabs, colour 1 = odd slabs */
  int n_threads = 1;
#ifdef _OPENMP
  n_threads = omp_get_max_threads();
#endif
  int n_slabs = 2 * n_threads;
  if (n_slabs > p->nc[2])
    n_slabs = p->nc[2] - (p->nc[2] & 1);
  if (n_slabs < 2)
    n_slabs = 2; /* nc[2] >= 2 is required by the caller */
  for (int colour = 0; colour < 2; ++colour)
    {
#pragma omp parallel for schedule(static)
      for (int s = colour; s < n_slabs; s += 2)
        {
          const int z0 = (int)((long)p->nc[2] * s / n_slabs), z1 = (int)((long)p->nc[2] * (s + 1) / n_slabs);
          v8d       u[n3], t0[n3], t1[n3], t2[n3];
          /* planes k z0 .. k z1 (inclusive) are touched; k z0 and k z1 are shared with the neighbouring slabs */
          if (colour == 0)
            pre_planes(p, h, k * z0, k * z1 + 1);
          else
            pre_planes(p, h, k * z0 + 1, k * z1);
          for (int cz = z0; cz < z1; ++cz)
            for (int cy = 0; cy < p->nc[1]; ++cy)
              for (int cx0 = 0; cx0 < p->nc[0]; cx0 += LANES)
                {
                  const int nb = p->nc[0] - cx0 < LANES ? p->nc[0] - cx0 : LANES;
                  gather(p, n, cx0, nb, cy, cz, src, u);
                  if (op == 0)
                    {
                      laplace_batch(p, n, u, t0, t1, t2);
                      scatter_add(p, n, cx0, nb, cy, cz, t0, h->dst);
                    }
                  else
                    {
                      fdm_batch(p, n, u, t0);
                      scatter_add(p, n, cx0, nb, cy, cz, u, h->dst);
                    }
                }
          if (colour == 0)
            post_planes(p, h, k * z0 + 1, k * z1);
          else
            post_planes(p, h, k * z0, k * z1 + 1);
        }
    }
}

#define INSTANTIATE(NN)                                                                          \
  static void cell_loop_##NN(const Problem *p, const int op, const double *src, const Hooks *h) \
  {                                                                                              \
    cell_loop(p, NN, op, src, h);                                                                \
  }
INSTANTIATE(2)
INSTANTIATE(3)
INSTANTIATE(4)
INSTANTIATE(5)
INSTANTIATE(6)
INSTANTIATE(7)
INSTANTIATE(8)
INSTANTIATE(9)

static void run_loop(const Problem *p, const int op, const double *src, const Hooks *h)
{
  switch (p->n)
    {
      case 2: cell_loop_2(p, op, src, h); break;
      case 3: cell_loop_3(p, op, src, h); break;
      case 4: cell_loop_4(p, op, src, h); break;
      case 5: cell_loop_5(p, op, src, h); break;
      case 6: cell_loop_6(p, op, src, h); break;
      case 7: cell_loop_7(p, op, src, h); break;
      case 8: cell_loop_8(p, op, src, h); break;
      case 9: cell_loop_9(p, op, src, h); break;
      default: break;
    }
}

/* One Chebyshev step of `degree` terms: x <- smoother(x, b).
 *   M_ref, K_ref  reference 1-D mass / stiffness matrices (n x n, row-major)
 *   hcell[3]      cell sizes
 *   S[3][n*n]     FDM eigenvectors per direction ([i][a]: component i of eigenvector a), lam[3][n] eigenvalues
 *   w[n^3]        weights of the gathered cell vector (NULL: none), applied before (w_pre) / after (w_post) the inverse
 *   f1[], f2[]    Chebyshev coefficients per term
 *   work          4 vectors of n_dofs doubles
 * returns 0, or 1 for unsupported sizes. */
int cpu_baseline_cheb_step(int k, const int nc[3], const double hcell[3], const double *M_ref, const double *K_ref, const double *S,
                           const double *lam, const double *w, int w_pre, int w_post, int degree, const double *f1, const double *f2,
                           double *x, const double *b, double *work)
{
  const int n = k + 1;
  if (n < 2 || n > 9 || nc[2] < 2)
    return 1;
  Problem p;
  p.k = k;
  p.n = n;
  double Mh[3][81], Kh[3][81], ST[3][81], ilam[729];
  for (int d = 0; d < 3; ++d)
    {
      p.nc[d] = nc[d];
      p.nd[d] = k * nc[d];
      /* K (x) M (x) M scaled by the cell: int = h_d / 2-type factors are already in M_ref / K_ref for the unit interval:
       * mass scales with h, stiffness with 1 / h */
      for (int i = 0; i < n * n; ++i)
        {
          Mh[d][i] = M_ref[i] * hcell[d];
          Kh[d][i] = K_ref[i] / hcell[d];
        }
      for (int i = 0; i < n; ++i)
        for (int a = 0; a < n; ++a)
          ST[d][a * n + i] = S[d * n * n + i * n + a];
      p.Mh[d] = Mh[d];
      p.K[d]  = Kh[d];
      p.S[d]  = S + d * n * n;
      p.ST[d] = ST[d];
    }
  /* the Kronecker terms need the mass factors of the two other directions: fold them in by using Mh in the sweeps and the
   * scaled K above (K_d / h_d, M_e h_e) */
  for (int c = 0; c < n; ++c)
    for (int bq = 0; bq < n; ++bq)
      for (int a = 0; a < n; ++a)
        ilam[(c * n + bq) * n + a] = 1.0 / (lam[a] + lam[n + bq] + lam[2 * n + c]);
  p.M      = M_ref;
  p.ilam   = ilam;
  p.w      = w;
  p.w_pre  = w_pre;
  p.w_post = w_post;
  const size_t N = (size_t)p.nd[0] * p.nd[1] * p.nd[2];
  double *t = work, *z = work + N;
  double *bufs[3] = {x, work + 2 * N, work + 3 * N};
  int     cur = 0, old = -1, nxt = 1;
  for (int term = 0; term < degree; ++term)
    {
      Hooks ha = {0, 0, 0, t, b, NULL, NULL, NULL};
      run_loop(&p, 0, bufs[cur], &ha); /* t = b - A cur */
      const int use_old = (old >= 0 && f1[term] != 0.0);
      Hooks     hp      = {1, old >= 0 ? f1[term] : 0.0, f2[term], z, NULL, bufs[cur], use_old ? bufs[old] : NULL, bufs[nxt]};
      run_loop(&p, 1, t, &hp); /* nxt = cur + f1 (cur - old) + f2 P^-1 t */
      const int prev_old = old;
      old                = cur;
      cur                = nxt;
      nxt                = prev_old < 0 ? 2 : prev_old;
    }
  const double *res = bufs[cur];
  if (res != x)
    memcpy(x, res, N * sizeof(double));
  return 0;
}
