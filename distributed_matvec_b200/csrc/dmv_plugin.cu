// dmv_plugin.cu -- the reference's plugin surface: ls_chpl_init / ls_chpl_finalize (src/library.c:19-34), the four entries of
// ls_chpl_kernels (src/FFI.chpl:233-239) and the PRIMME callback (src/Diagonalize.chpl:134-162), on contexts bound to the
// opaque handles of the Haskell library.
#include "dmv_context.h"

namespace dmv { namespace host {


std::mutex g_bind_mutex;
std::map<const void *, dmv_context *> g_bindings;

} }  // namespace dmv::host

extern "C" {

void ls_chpl_init(void) {}      // no runtime to start (reference src/library.c:19-32 boots the Chapel runtime)
void ls_chpl_finalize(void) {}  // reference src/library.c:34

extern "C++" {
namespace {
dmv_context *bound_context(const void *key, const char *who) {
  dmv_context *ctx = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_bind_mutex);
    auto it = g_bindings.find(key);
    if (it != g_bindings.end()) ctx = it->second;
  }
  if (!ctx) { fprintf(stderr, "%s: handle is not bound to a dmv context (dmv_bind_operator)\n", who); abort(); }
  return ctx;
}
template <typename T>
dmv_external_array external_array(size_t n) {   // convertToExternalArray (BO:232): callee allocates, caller frees
  dmv_external_array a;
  a.elts = n ? malloc(n * sizeof(T)) : nullptr;
  a.num_elts = n;
  a.freer = n ? &free : nullptr;
  if (n && !a.elts) { fprintf(stderr, "out of memory\n"); abort(); }
  return a;
}
}  // namespace
}  // extern "C++"

void ls_chpl_operator_apply_diag(const void *ls_hs_operator_ptr, int64_t count, const uint64_t *alphas,
                                 dmv_external_array *coeffs, int64_t /*num_tasks*/) {
  dmv_context *ctx = bound_context(ls_hs_operator_ptr, "ls_chpl_operator_apply_diag");
  *coeffs = external_array<double>((size_t)count);
  if (dmv_apply_diag(ctx, count, alphas, (double *)coeffs->elts) != 0) { fprintf(stderr, "%s\n", dmv_last_error()); abort(); }
}

void ls_chpl_operator_apply_off_diag(const void *ls_hs_operator_ptr, int64_t count, const uint64_t *alphas,
                                     dmv_external_array *betas, dmv_external_array *coeffs,
                                     dmv_external_array *offsets, int64_t /*num_tasks*/) {
  dmv_context *ctx = bound_context(ls_hs_operator_ptr, "ls_chpl_operator_apply_off_diag");
  const size_t T = (size_t)dmv_max_number_off_diag(ctx);
  *offsets = external_array<int64_t>((size_t)count + 1);
  if (T == 0) {   // BO:269-273
    betas->elts = nullptr; betas->num_elts = 0; betas->freer = nullptr;
    coeffs->elts = nullptr; coeffs->num_elts = 0; coeffs->freer = nullptr;
    memset(offsets->elts, 0, ((size_t)count + 1) * sizeof(int64_t));
    return;
  }
  *betas = external_array<uint64_t>((size_t)count * T);
  *coeffs = external_array<double>((size_t)count * T * 2);
  if (dmv_apply_off_diag(ctx, count, alphas, (uint64_t *)betas->elts, (double *)coeffs->elts,
                         (int64_t *)offsets->elts) != 0) { fprintf(stderr, "%s\n", dmv_last_error()); abort(); }
}

// reference src/StatesEnumeration.chpl:588-603: the bounds are accepted and ignored there too (the whole range of the
// basis is enumerated); returns this locale's block
void ls_chpl_enumerate_representatives(const void *ls_hs_basis_ptr, uint64_t /*lower*/, uint64_t /*upper*/,
                                       dmv_external_array *dest) {
  dmv_context *ctx = bound_context(ls_hs_basis_ptr, "ls_chpl_enumerate_representatives");
  if (dmv_number_states(ctx) < 0 && dmv_basis_build(ctx) != 0) { fprintf(stderr, "%s\n", dmv_last_error()); abort(); }
  *dest = external_array<uint64_t>((size_t)dmv_number_states(ctx));
  if (dmv_get_representatives(ctx, (uint64_t *)dest->elts, nullptr) != 0) { fprintf(stderr, "%s\n", dmv_last_error()); abort(); }
}

int dmv_bind_operator(const void *ls_hs_operator_ptr, dmv_context *ctx) {
  API_BEGIN
  std::lock_guard<std::mutex> lock(g_bind_mutex);
  if (ctx) g_bindings[ls_hs_operator_ptr] = ctx;
  else g_bindings.erase(ls_hs_operator_ptr);
  API_END
}

// reference: src/DistributedMatrixVector.chpl:1095-1110.  Halts (abort) on error like the reference.
void ls_chpl_matrix_vector_product(const void *ls_hs_operator_ptr, int num_vectors, double *x, double *y) {
  dmv_context *ctx = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_bind_mutex);
    auto it = g_bindings.find(ls_hs_operator_ptr);
    if (it != g_bindings.end()) ctx = it->second;
  }
  if (!ctx) { fprintf(stderr, "ls_chpl_matrix_vector_product: operator is not bound to a dmv context\n"); abort(); }
  if (num_vectors != 1) {  // DMV:1101-1102
    fprintf(stderr, "applying the Operator to more than 1 vector is not yet implemented\n");
    abort();
  }
  if (dmv_local_matvec(ctx, DMV_F64, x, y) != 0) { fprintf(stderr, "%s\n", dmv_last_error()); abort(); }
}

// reference: src/Diagonalize.chpl:134-162 -- the matrix-vector callback PRIMME drives (`primme.matrixMatvec`):
// blockSize columns of real(64), leading dimensions ldx / ldy >= nLocal, column k through localMatrixVector.  The
// reference reads the operator from primme->matrix; here the primme_params pointer itself is the handle, bound to a
// context with dmv_bind_operator (no dependence on PRIMME's struct layout).  Contiguous columns go through
// dmv_matvec_batch (four columns share one term walk in k_gather); collective when the context has several ranks.
void ls_chpl_primme_matvec(void *x, int64_t *ldx, void *y, int64_t *ldy, int *block_size, void *primme, int *ierr) {
  dmv_context *ctx = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_bind_mutex);
    auto it = g_bindings.find(primme);
    if (it != g_bindings.end()) ctx = it->second;
  }
  if (!ctx) { fprintf(stderr, "ls_chpl_primme_matvec: primme_params is not bound to a dmv context\n"); abort(); }
  const int64_t n = ctx->n_states;
  if (*ldx < n || *ldy < n) { fprintf(stderr, "ls_chpl_primme_matvec: leading dimension below nLocal\n"); abort(); }   // :143-144
  int rc = 0;
  if (*ldx == n && *ldy == n) {
    rc = dmv_matvec_batch(ctx, DMV_F64, *block_size, x, y);
  } else {
    for (int k = 0; k < *block_size && rc == 0; ++k) {
      const double *xk = reinterpret_cast<const double *>(x) + *ldx * k;
      double *yk = reinterpret_cast<double *>(y) + *ldy * k;
      rc = ctx->num_ranks == 1 ? dmv_local_matvec(ctx, DMV_F64, xk, yk) : dmv_matvec(ctx, DMV_F64, xk, yk);
    }
  }
  if (rc == 0 && is_device_pointer(y)) rc = dmv_synchronize(ctx);
  if (rc != 0) { fprintf(stderr, "%s\n", dmv_last_error()); abort(); }   // the reference halts
  *ierr = 0;                                                              // :160
}

}  // extern "C"
