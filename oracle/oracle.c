/*
 * oracle.c -- CPU restatement of the reference algorithm for the H.x hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (distributed_matvec_b200/) may call,
 * link or import this file; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
 * `--impl reference` legs do, and only as the checker or the timed CPU baseline.
 *
 * PARITY UNPINNED: the reference (twesterhout/distributed-matvec @ 050c3f2) cannot be built in
 * this image (Chapel, GHC, HDF5 and liblattice_symmetries_haskell are absent) and its golden HDF5
 * vectors are downloaded at test time (reference Makefile:128-146), so no reference artefact pins
 * this restatement.  It is pinned instead by independent constructions (tests/test_oracle_pins.py):
 * Kronecker-product matrices (oracle/dense_pin.py; dense to 16 sites, sparse to 20: heisenberg_kagome_16
 * at full size), the Heisenberg definition in numpy on all 2 704 156 states of heisenberg_chain_24,
 * exact basis dimensions, the Bethe-ansatz ground-state energy of every ring (tests/bethe.py, 1e-10;
 * heisenberg_chain_32_symm at full size: tools/oracle_ground_state.py) and ground-state energies from
 * the exact-diagonalisation literature (4x4 torus, 12-site kagome cluster).
 *
 * The arithmetic of term generation / symmetry projection / state indexing lives in the
 * third-party library lattice-symmetries-haskell (release `continuous`, build 14e7319, reference
 * .github/workflows/ci.yml:6,29-31) which is not vendored; its published contract is restated here
 * from the reference's call sites.  Every function cites the reference file:line it follows.
 *
 * Build: see oracle/build.py (gcc -O3 -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double re, im; } c128;

static inline c128 c_mul(c128 a, c128 b) {
  c128 r = {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
  return r;
}
static inline c128 c_scale(c128 a, double s) { c128 r = {a.re * s, a.im * s}; return r; }

/* ---------------------------------------------------------------------------------------------
 * hash64_01 / localeIdxOf            reference: src/StatesEnumeration.chpl:122-136
 * ------------------------------------------------------------------------------------------- */
uint64_t oracle_hash64_01(uint64_t x) {
  x = (x ^ (x >> 30)) * UINT64_C(0xbf58476d1ce4e5b9);
  x = (x ^ (x >> 27)) * UINT64_C(0x94d049bb133111eb);
  x = x ^ (x >> 31);
  return x;
}

int oracle_locale_idx_of(uint64_t state, int num_locales) {
  /* CHPL_COMM == "" (single locale) returns 0: src/StatesEnumeration.chpl:129-132 */
  if (num_locales <= 1) return 0;
  return (int)(oracle_hash64_01(state) % (uint64_t)num_locales);
}

void oracle_locale_idx_of_many(int64_t n, const uint64_t *states, int num_locales, uint8_t *keys) {
  for (int64_t i = 0; i < n; ++i) keys[i] = (uint8_t)oracle_locale_idx_of(states[i], num_locales);
}

/* ---------------------------------------------------------------------------------------------
 * nextStateFixedHamming / nextStateGeneral     reference: src/StatesEnumeration.chpl:31-38
 * ------------------------------------------------------------------------------------------- */
static inline uint64_t next_state_fixed_hamming(uint64_t v) {
  const uint64_t t = v | (v - 1);
  return (t + 1) | (((~t & (t + 1)) - 1) >> (__builtin_ctzll(v) + 1));
}

/* ---------------------------------------------------------------------------------------------
 * Non-branching terms (third-party contract; declared at reference src/FFI.chpl:219-225).
 *   <beta|t|alpha> = v * [alpha & m == r] * (-1)^popcount(alpha & s),  beta = alpha ^ x
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int64_t count;
  const c128 *v;
  const uint64_t *m, *r, *x, *s;
} oracle_terms;

static inline int term_matches(const oracle_terms *t, int64_t k, uint64_t alpha) {
  return (alpha & t->m[k]) == t->r[k];
}
static inline double term_sign(const oracle_terms *t, int64_t k, uint64_t alpha) {
  return (__builtin_popcountll(alpha & t->s[k]) & 1) ? -1.0 : 1.0;
}

/* ls_internal_operator_apply_diag_x1(op, batch, alphas, ys, xs)
 *   declared src/FFI.chpl:219-221; called src/DistributedMatrixVector.chpl:43-45 (ys = y chunk,
 *   xs = x chunk: y[i] = x[i] * Re(sum_t ...)) and src/BatchedOperator.chpl:230 (xs = nil: ys[i] =
 *   the diagonal matrix element itself).  elt = 1: real x/y; elt = 2: complex x/y (our extension,
 *   the diagonal coefficient then stays complex). */
void oracle_apply_diag_x1(int64_t T, const c128 *v, const uint64_t *m, const uint64_t *r,
                          const uint64_t *s, int64_t batch, const uint64_t *alphas, double *ys,
                          const double *xs, int elt) {
  oracle_terms t = {T, v, m, r, NULL, s};
  for (int64_t i = 0; i < batch; ++i) {
    const uint64_t alpha = alphas[i];
    c128 acc = {0.0, 0.0};
    for (int64_t k = 0; k < T; ++k) {
      if (term_matches(&t, k, alpha)) {
        const double sg = term_sign(&t, k, alpha);
        acc.re += sg * v[k].re;
        acc.im += sg * v[k].im;
      }
    }
    if (elt == 1) {
      ys[i] = (xs != NULL) ? xs[i] * acc.re : acc.re;
    } else {
      c128 xi = {1.0, 0.0};
      if (xs != NULL) { xi.re = xs[2 * i]; xi.im = xs[2 * i + 1]; }
      const c128 y = c_mul(acc, xi);
      ys[2 * i] = y.re;
      ys[2 * i + 1] = y.im;
    }
  }
}

/* ls_internal_operator_apply_off_diag_x1(op, batch, alphas, betas, coeffs, offsets, xs)
 *   declared src/FFI.chpl:222-225; called src/BatchedOperator.chpl:99-106,129-136,168-175,256-263.
 *   For each alpha_i and each term that matches, emit (beta, v*sign*xs[i]); offsets is the CSR row
 *   pointer (offsets[count] = total, BO:109).  xs == NULL means "times one" (BO:263).
 *   Terms sharing a flip mask x are emitted separately, in table order. */
int64_t oracle_apply_off_diag_x1(int64_t T, const c128 *v, const uint64_t *m, const uint64_t *r,
                                 const uint64_t *x, const uint64_t *s, int64_t batch,
                                 const uint64_t *alphas, uint64_t *betas, c128 *coeffs,
                                 int64_t *offsets, const double *xs, int elt) {
  oracle_terms t = {T, v, m, r, x, s};
  int64_t n = 0;
  offsets[0] = 0;
  for (int64_t i = 0; i < batch; ++i) {
    const uint64_t alpha = alphas[i];
    c128 xi = {1.0, 0.0};
    if (xs != NULL) {
      if (elt == 1) { xi.re = xs[i]; }
      else { xi.re = xs[2 * i]; xi.im = xs[2 * i + 1]; }
    }
    for (int64_t k = 0; k < T; ++k) {
      if (term_matches(&t, k, alpha)) {
        betas[n] = alpha ^ x[k];
        coeffs[n] = c_mul(c_scale(v[k], term_sign(&t, k, alpha)), xi);
        ++n;
      }
    }
    offsets[i + 1] = n;
  }
  return n;
}

/* ---------------------------------------------------------------------------------------------
 * Symmetry group (third-party contract).  Element g: result bit i = input bit
 * perm[g][i]; if flips[g] the result is then spin-inverted (xor with the n-site mask).
 * Deliberately naive (bit by bit) so that it shares nothing with the GPU's permutation networks.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int n;          /* number of sites */
  int64_t G;      /* group order */
  const int32_t *perms; /* [G][n] */
  const uint8_t *flips; /* [G] */
  const c128 *chars;    /* [G] */
} oracle_group;

static inline uint64_t group_apply(const oracle_group *g, int64_t e, uint64_t s) {
  const int32_t *p = g->perms + e * g->n;
  uint64_t out = 0;
  for (int i = 0; i < g->n; ++i) out |= ((s >> p[i]) & 1u) << i;
  if (g->flips[e]) out ^= (g->n == 64) ? ~UINT64_C(0) : ((UINT64_C(1) << g->n) - 1);
  return out;
}

/* ls_hs_state_info(basis, batch, alphas, 1, betas, 1, characters, norms)
 *   declared src/FFI.chpl:181-184; called src/BatchedOperator.chpl:188-194.
 *   betas[k] = min_g g(alpha_k); characters[k] = conj(chi(g_min)) -- the factor that makes
 *   BO:200 `cs[k] *= tempCoeffs[k] * norms[k] / norms[total + i]` the projected matrix element
 *   (derivation in DESIGN.md); norms[k] = sqrt( (1/|G|) * sum_{g: g(alpha)=alpha} Re chi(g) ). */
void oracle_state_info(int n, int64_t G, const int32_t *perms, const uint8_t *flips,
                       const c128 *chars, int64_t batch, const uint64_t *alphas, uint64_t *betas,
                       c128 *characters, double *norms) {
  oracle_group g = {n, G, perms, flips, chars};
  for (int64_t k = 0; k < batch; ++k) {
    const uint64_t alpha = alphas[k];
    uint64_t best = alpha;
    c128 chi = {1.0, 0.0};
    int have = 0;
    double stab = 0.0;
    for (int64_t e = 0; e < G; ++e) {
      const uint64_t y = group_apply(&g, e, alpha);
      if (!have || y < best) { best = y; chi = chars[e]; have = 1; }
      if (y == alpha) stab += chars[e].re;
    }
    betas[k] = best;
    characters[k].re = chi.re;
    characters[k].im = -chi.im;
    double nn = stab / (double)G;
    if (nn < 0.0 && nn > -1e-12) nn = 0.0;
    norms[k] = (nn > 1e-12) ? sqrt(nn) : 0.0;
  }
}

/* ls_hs_is_representative(basis, batch, alphas, 1, are_representatives, norms)
 *   declared src/FFI.chpl:177-179; called src/ForeignTypes.chpl:137-143 (via
 *   src/StatesEnumeration.chpl:180-188: keep a state iff flag && norm > 0). */
void oracle_is_representative(int n, int64_t G, const int32_t *perms, const uint8_t *flips,
                              const c128 *chars, int64_t batch, const uint64_t *alphas,
                              uint8_t *are_representatives, double *norms) {
  oracle_group g = {n, G, perms, flips, chars};
  for (int64_t k = 0; k < batch; ++k) {
    const uint64_t alpha = alphas[k];
    int is_rep = 1;
    double stab = 0.0;
    for (int64_t e = 0; e < G; ++e) {
      const uint64_t y = group_apply(&g, e, alpha);
      if (y < alpha) { is_rep = 0; break; }
      if (y == alpha) stab += chars[e].re;
    }
    are_representatives[k] = (uint8_t)is_rep;
    double nn = is_rep ? stab / (double)G : 0.0;
    norms[k] = (nn > 1e-12) ? sqrt(nn) : 0.0;
  }
}

/* ---------------------------------------------------------------------------------------------
 * Basis enumeration       reference: src/StatesEnumeration.chpl:158-224 (projected / unprojected)
 * Sequential restatement of one chunk covering [lower, upper]; result ascending.
 * Two-pass usage: call with out == NULL to count.  G == 0 means "no projection".
 * inversion_only: representatives are {s : s < s ^ mask}.
 * ------------------------------------------------------------------------------------------- */
int64_t oracle_enumerate_states(uint64_t lower, uint64_t upper, int fixed_hamming, int n, int64_t G,
                                const int32_t *perms, const uint8_t *flips, const c128 *chars,
                                uint64_t *out, double *out_norms) {
  oracle_group g = {n, G, perms, flips, chars};
  int64_t count = 0;
  if (lower > upper) return 0;
  uint64_t v = lower;
  for (;;) {
    int keep = 1;
    double norm = 1.0;
    if (G > 0) {
      double stab = 0.0;
      for (int64_t e = 0; e < G; ++e) {
        const uint64_t y = group_apply(&g, e, v);
        if (y < v) { keep = 0; break; }
        if (y == v) stab += chars[e].re;
      }
      if (keep) {
        const double nn = stab / (double)G;
        norm = (nn > 1e-12) ? sqrt(nn) : 0.0;
        if (!(norm > 0.0)) keep = 0;  /* SE:186-188: flags[i] && norms[i] > 0 */
      }
    }
    if (keep) {
      if (out != NULL) { out[count] = v; if (out_norms != NULL) out_norms[count] = norm; }
      ++count;
    }
    if (v == upper) break;   /* SE:44-45: incrementing past the bound may overflow */
    v = fixed_hamming ? next_state_fixed_hamming(v) : v + 1;
    if (v > upper) break;
  }
  return count;
}

/* ---------------------------------------------------------------------------------------------
 * The same third-party kernels with the group applied as Benes networks (oracle/networks.py): what a tuned CPU
 * library does, used ONLY by the timed CPU arm (bench.py cpu_baseline / --impl reference) and by the parallel
 * enumeration that prepares its input.  The bit-by-bit versions above stay the checker.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int n; int64_t G; int64_t n_stages;
  const int32_t *delta; const uint64_t *masks; const uint8_t *flips; const c128 *chars;
} oracle_networks;

static inline uint64_t network_apply(const oracle_networks *g, int64_t e, uint64_t s) {
  const uint64_t *m = g->masks + e * g->n_stages;
  for (int64_t st = 0; st < g->n_stages; ++st) {
    const int d = g->delta[st];
    const uint64_t t = ((s >> d) ^ s) & m[st];
    s ^= t ^ (t << d);
  }
  if (g->flips[e]) s ^= (g->n == 64) ? ~UINT64_C(0) : ((UINT64_C(1) << g->n) - 1);
  return s;
}

void oracle_state_info_networks(int n, int64_t G, int64_t n_stages, const int32_t *delta, const uint64_t *masks,
                                const uint8_t *flips, const c128 *chars, int64_t batch, const uint64_t *alphas,
                                uint64_t *betas, c128 *characters, double *norms) {
  const oracle_networks g = {n, G, n_stages, delta, masks, flips, chars};
  for (int64_t k = 0; k < batch; ++k) {
    const uint64_t alpha = alphas[k];
    uint64_t best = alpha;
    c128 chi = {1.0, 0.0};
    int have = 0;
    double stab = 0.0;
    for (int64_t e = 0; e < G; ++e) {
      const uint64_t y = network_apply(&g, e, alpha);
      if (!have || y < best) { best = y; chi = chars[e]; have = 1; }
      if (y == alpha) stab += chars[e].re;
    }
    betas[k] = best;
    characters[k].re = chi.re;
    characters[k].im = -chi.im;
    double nn = stab / (double)G;
    if (nn < 0.0 && nn > -1e-12) nn = 0.0;
    norms[k] = (nn > 1e-12) ? sqrt(nn) : 0.0;
  }
}

/* combinadic rank / unrank of fixed-Hamming-weight states (ls_hs_fixed_hamming_state_to_index / index_to_state,
 * reference src/FFI.chpl:165-166; used by determineEnumerationRanges, src/StatesEnumeration.chpl:80-116) */
static uint64_t g_binom[65][65];
static void binom_init(void) {
  if (g_binom[0][0] == 1) return;
  for (int i = 0; i <= 64; ++i) {
    g_binom[i][0] = 1;
    for (int j = 1; j <= i; ++j) {
      const uint64_t a = g_binom[i - 1][j - 1], b = (j <= i - 1) ? g_binom[i - 1][j] : 0;
      g_binom[i][j] = (a > UINT64_MAX - b) ? UINT64_MAX : a + b;
    }
  }
}
static uint64_t fixed_hamming_rank(uint64_t s) {
  uint64_t r = 0;
  int k = 0;
  while (s) { const int pos = __builtin_ctzll(s); ++k; r += g_binom[pos][k]; s &= s - 1; }
  return r;
}
static uint64_t fixed_hamming_unrank(uint64_t r, int weight) {
  uint64_t s = 0;
  for (int k = weight; k >= 1; --k) {
    int pos = k - 1;
    while (pos + 1 <= 63 && g_binom[pos + 1][k] <= r) ++pos;
    r -= g_binom[pos][k];
    s |= UINT64_C(1) << pos;
  }
  return s;
}

/* enumerateStates on one locale with the candidate range cut into chunks handled by OpenMP threads
 * (reference src/StatesEnumeration.chpl:94-116 determineEnumerationRanges, :158-224 per-chunk filter).
 * Result identical to oracle_enumerate_states; out == NULL counts.  masks == NULL: bit-by-bit group. */
int64_t oracle_enumerate_states_parallel(uint64_t lower, uint64_t upper, int fixed_hamming, int n, int64_t G,
                                         const int32_t *perms, const uint8_t *flips, const c128 *chars,
                                         int64_t n_stages, const int32_t *delta, const uint64_t *masks,
                                         uint64_t *out, double *out_norms) {
  if (lower > upper) return 0;
  binom_init();
  const int weight = __builtin_popcountll(lower);
  const uint64_t first = fixed_hamming ? fixed_hamming_rank(lower) : lower;
  const uint64_t last = fixed_hamming ? fixed_hamming_rank(upper) : upper;
  const uint64_t total = last - first + 1;
  int64_t n_chunks = 4096;
  if ((uint64_t)n_chunks > total) n_chunks = (int64_t)total;
  int64_t *counts = (int64_t *)calloc((size_t)n_chunks + 1, sizeof(int64_t));
  const oracle_group gb = {n, G, perms, flips, chars};
  const oracle_networks gn = {n, G, n_stages, delta, masks, flips, chars};
  for (int pass = 0; pass < (out ? 2 : 1); ++pass) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t c = 0; c < n_chunks; ++c) {
      const uint64_t lo = first + (uint64_t)(((__uint128_t)total * (uint64_t)c) / (uint64_t)n_chunks);
      const uint64_t hi = first + (uint64_t)(((__uint128_t)total * (uint64_t)(c + 1)) / (uint64_t)n_chunks);
      if (hi <= lo) continue;
      uint64_t v = fixed_hamming ? fixed_hamming_unrank(lo, weight) : lo;
      int64_t cnt = 0;
      const int64_t base = pass ? counts[c] : 0;
      for (uint64_t idx = lo; idx < hi; ++idx) {
        int keep = 1;
        double norm = 1.0;
        if (G > 0) {
          double stab = 0.0;
          for (int64_t e = 0; e < G; ++e) {
            const uint64_t y = masks ? network_apply(&gn, e, v) : group_apply(&gb, e, v);
            if (y < v) { keep = 0; break; }
            if (y == v) stab += chars[e].re;
          }
          if (keep) {
            const double nn = stab / (double)G;
            norm = (nn > 1e-12) ? sqrt(nn) : 0.0;
            if (!(norm > 0.0)) keep = 0;
          }
        }
        if (keep) {
          if (pass) { out[base + cnt] = v; if (out_norms) out_norms[base + cnt] = norm; }
          ++cnt;
        }
        if (idx + 1 < hi) v = fixed_hamming ? next_state_fixed_hamming(v) : v + 1;
      }
      if (!pass) counts[c + 1] = cnt;
    }
    if (!pass) {
      counts[0] = 0;
      for (int64_t c = 0; c < n_chunks; ++c) counts[c + 1] += counts[c];
    }
  }
  const int64_t result = counts[n_chunks];
  free(counts);
  return result;
}

/* ---------------------------------------------------------------------------------------------
 * ls_hs_state_index(basis, batch, spins, 1, indices, 1)
 *   declared src/FFI.chpl:173-175; called src/DistributedMatrixVector.chpl:102.
 *   Position in the ascending `representatives` installed by uncheckedSetRepresentatives
 *   (src/ForeignTypes.chpl:74-77, DMV:1084); -1 when absent.
 * ------------------------------------------------------------------------------------------- */
void oracle_state_index(int64_t N, const uint64_t *representatives, int64_t batch,
                        const uint64_t *spins, int64_t *indices) {
  for (int64_t k = 0; k < batch; ++k) {
    const uint64_t key = spins[k];
    int64_t lo = 0, hi = N;
    while (lo < hi) {
      const int64_t mid = lo + (hi - lo) / 2;
      if (representatives[mid] < key) lo = mid + 1; else hi = mid;
    }
    indices[k] = (lo < N && representatives[lo] == key) ? lo : -1;
  }
}

/* ---------------------------------------------------------------------------------------------
 * The distributed matrix-vector product, restated for P logical locales in one address space.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  /* operator */
  int64_t T_off; const c128 *off_v; const uint64_t *off_m, *off_r, *off_x, *off_s;
  int64_t T_diag; const c128 *diag_v; const uint64_t *diag_m, *diag_r, *diag_s;
  int64_t max_off_diag;      /* numberOffDiagTerms(): src/ForeignTypes.chpl:228-229 */
  /* basis */
  int n; int spin_inversion; int has_permutations; int state_index_is_identity;
  int64_t G; const int32_t *perms; const uint8_t *flips; const c128 *chars;
  /* optional: the same group as Benes networks (oracle/networks.py).  Only the TIMED arm sets them; the checker
   * leaves net_masks NULL and permutes bit by bit. */
  int64_t n_stages; const int32_t *net_delta; const uint64_t *net_masks;   /* [G][n_stages] */
} oracle_model;

typedef struct {
  int64_t capacity;
  uint64_t *spins1, *spins2;
  c128 *coeffs1, *coeffs2;
  double *norms;
  uint8_t *keys;
  int64_t *offsets;
} batched_operator;  /* src/BatchedOperator.chpl:40-66 */

static void bo_init(batched_operator *bo, int64_t batch_size, int64_t number_terms) {
  if (number_terms < 1) number_terms = 1;               /* BO:63 */
  bo->capacity = batch_size * (number_terms + 1);       /* BO:64 */
  bo->spins1 = (uint64_t *)malloc(sizeof(uint64_t) * bo->capacity);
  bo->spins2 = (uint64_t *)malloc(sizeof(uint64_t) * bo->capacity);
  bo->coeffs1 = (c128 *)malloc(sizeof(c128) * bo->capacity);
  bo->coeffs2 = (c128 *)malloc(sizeof(c128) * bo->capacity);
  bo->norms = (double *)malloc(sizeof(double) * bo->capacity);
  bo->keys = (uint8_t *)malloc(bo->capacity);
  bo->offsets = (int64_t *)malloc(sizeof(int64_t) * (batch_size + 1));
}
static void bo_free(batched_operator *bo) {
  free(bo->spins1); free(bo->spins2); free(bo->coeffs1); free(bo->coeffs2);
  free(bo->norms); free(bo->keys); free(bo->offsets);
}

/* BatchedOperator.computeOffDiag         reference: src/BatchedOperator.chpl:82-213
 * returns n; betas, coeffs and keys alias the workspace (valid until the next call). */
static int64_t bo_compute_off_diag(batched_operator *bo, const oracle_model *M, int num_locales,
                                   int64_t count, const uint64_t *alphas, const double *xs, int elt,
                                   uint64_t **betas_out, c128 **coeffs_out, uint8_t **keys_out) {
  const int requires_projection = M->has_permutations || M->spin_inversion != 0;
  if (!requires_projection) {                                               /* BO:89-116 */
    const int64_t total = oracle_apply_off_diag_x1(M->T_off, M->off_v, M->off_m, M->off_r, M->off_x,
                                                   M->off_s, count, alphas, bo->spins1, bo->coeffs1,
                                                   bo->offsets, xs, elt);
    oracle_locale_idx_of_many(total, bo->spins1, num_locales, bo->keys);
    *betas_out = bo->spins1; *coeffs_out = bo->coeffs1; *keys_out = bo->keys;
    return total;
  }
  if (!M->has_permutations && M->spin_inversion != 0) {                     /* BO:119-161 */
    const int64_t total = oracle_apply_off_diag_x1(M->T_off, M->off_v, M->off_m, M->off_r, M->off_x,
                                                   M->off_s, count, alphas, bo->spins1, bo->coeffs1,
                                                   bo->offsets, xs, elt);
    const uint64_t mask = (M->n == 64) ? ~UINT64_C(0) : ((UINT64_C(1) << M->n) - 1);   /* BO:140 */
    const double character = (double)M->spin_inversion;                               /* BO:141 */
    for (int64_t i = 0; i < total; ++i) {                                             /* BO:145-152 */
      const uint64_t current = bo->spins1[i];
      const uint64_t inverted = current ^ mask;
      if (inverted < current) {
        bo->spins1[i] = inverted;
        bo->coeffs1[i] = c_scale(bo->coeffs1[i], character);
      }
    }
    oracle_locale_idx_of_many(total, bo->spins1, num_locales, bo->keys);
    *betas_out = bo->spins1; *coeffs_out = bo->coeffs1; *keys_out = bo->keys;
    return total;
  }
  /* BO:163-212: project every beta and every source alpha */
  const int64_t total = oracle_apply_off_diag_x1(M->T_off, M->off_v, M->off_m, M->off_r, M->off_x,
                                                 M->off_s, count, alphas, bo->spins2, bo->coeffs2,
                                                 bo->offsets, xs, elt);
  memcpy(bo->spins2 + total, alphas, (size_t)count * sizeof(uint64_t));              /* BO:181 */
  if (M->net_masks != NULL)   /* timed arm: the group as Benes networks */
    oracle_state_info_networks(M->n, M->G, M->n_stages, M->net_delta, M->net_masks, M->flips, M->chars,
                               total + count, bo->spins2, bo->spins1, bo->coeffs1, bo->norms);
  else
    oracle_state_info(M->n, M->G, M->perms, M->flips, M->chars, total + count, bo->spins2,
                      bo->spins1, bo->coeffs1, bo->norms);                            /* BO:188-194 */
  for (int64_t i = 0; i < count; ++i) {                                               /* BO:198-202 */
    for (int64_t k = bo->offsets[i]; k < bo->offsets[i + 1]; ++k) {
      /* cs[k] *= tempCoeffs[k] * norms[k] / norms[totalCount + i] */
      const c128 t = c_scale(bo->coeffs2[k], bo->norms[k] / bo->norms[total + i]);
      bo->coeffs1[k] = c_mul(bo->coeffs1[k], t);
    }
  }
  oracle_locale_idx_of_many(total, bo->spins1, num_locales, bo->keys);
  *betas_out = bo->spins1; *coeffs_out = bo->coeffs1; *keys_out = bo->keys;
  return total;
}

/* radixOneStep       reference: src/DistributedMatrixVector.chpl:265-311
 * Postcondition restated (not the in-place American-flag swaps): keys/betas/coeffs grouped by key,
 * offsets[k] = start of bucket k, offsets[256] = n.  Order inside a bucket is unspecified in the
 * reference (unstable swaps); a stable counting sort satisfies the same postcondition. */
static void radix_one_step(int64_t n, uint8_t *keys, int64_t offsets[257], uint64_t *betas,
                           c128 *coeffs, uint64_t *tmp_b, c128 *tmp_c) {
  int64_t counts[256];
  memset(counts, 0, sizeof(counts));
  for (int64_t i = 0; i < n; ++i) counts[keys[i]]++;
  offsets[0] = 0;
  for (int k = 0; k < 256; ++k) offsets[k + 1] = offsets[k] + counts[k];
  int64_t cursor[256];
  for (int k = 0; k < 256; ++k) cursor[k] = offsets[k];
  for (int64_t i = 0; i < n; ++i) {
    const int64_t d = cursor[keys[i]]++;
    tmp_b[d] = betas[i];
    tmp_c[d] = coeffs[i];
  }
  memcpy(betas, tmp_b, (size_t)n * sizeof(uint64_t));
  memcpy(coeffs, tmp_c, (size_t)n * sizeof(c128));
  for (int k = 0; k < 256; ++k)
    for (int64_t i = offsets[k]; i < offsets[k + 1]; ++i) keys[i] = (uint8_t)k;
}

static inline void atomic_add_f64(double *p, double v) {
  /* ConcurrentAccessor.localAdd: relaxed atomic fetch-add   reference: src/ConcurrentAccessor.chpl:48-54 */
#pragma omp atomic
  *p += v;
}

/* localProcess        reference: src/DistributedMatrixVector.chpl:73-127
 * returns 0 on success, or 1 + k when element k has c != 0 and is not in the basis (the reference
 * halts: DMV:115-118). */
static int64_t local_process(const oracle_model *M, int num_locales, int64_t N,
                             const uint64_t *representatives, double *y, int elt,
                             const uint64_t *basis_states, const c128 *coeffs, int64_t size) {
  if (size == 0) return 0;                                                     /* DMV:83 */
  /* The identity fast path indexes the LOCAL block with the state itself, which is only meaningful
   * when one locale owns every state; with several locales the hashed blocks need the search. */
  if (M->state_index_is_identity && num_locales == 1) {                        /* DMV:86-95 */
    for (int64_t k = 0; k < size; ++k) {
      const int64_t i = (int64_t)basis_states[k];
      atomic_add_f64(&y[elt * i], coeffs[k].re);                               /* c:coeffType cast */
      if (elt == 2) atomic_add_f64(&y[2 * i + 1], coeffs[k].im);
    }
    return 0;
  }
  int64_t *indices = (int64_t *)malloc(sizeof(int64_t) * size);                /* DMV:98 */
  oracle_state_index(N, representatives, size, basis_states, indices);        /* DMV:102 */
  int64_t status = 0;
  for (int64_t k = 0; k < size; ++k) {                                         /* DMV:107-120 */
    const int64_t i = indices[k];
    const c128 c = coeffs[k];
    const int nonzero = (elt == 1) ? (c.re != 0.0) : (c.re != 0.0 || c.im != 0.0);
    if (nonzero) {
      if (i >= 0) {
        atomic_add_f64(&y[elt * i], c.re);
        if (elt == 2) atomic_add_f64(&y[2 * i + 1], c.im);
      } else if (status == 0) {
        status = 1 + k;
      }
    }
  }
  free(indices);
  return status;
}

/* matrixVectorProduct / localMatrixVector / localOffDiagonalNoQueue for P logical locales.
 *   reference: src/DistributedMatrixVector.chpl:1072-1093, :1055-1070, :856-1053.
 * reps[p], xs[p], ys[p] are locale p's blocks (sizes[p] elements; x/y have `elt` doubles each).
 * Diagonal first (overwrites y when the operator has diagonal terms, DMV:1062-1063), then for every
 * locale and every chunk: computeOffDiag -> radixOneStep -> localProcess on the owner (the
 * PUT + flag handshake of DMV:638-661,818-852 is a plain call here).
 * Chunk sizing follows DMV:871-883,905-907 with num_producer_tasks.
 * Returns 0, or -(1) on an invalid index (DMV:115-118).  OpenMP parallelises over chunks (the
 * reference's producer tasks, DMV:957-1011). */
static int64_t matvec_impl(const oracle_model *M, int P, const int64_t *sizes, const uint64_t *const *reps,
                           const double *const *xs, double *const *ys, int elt, int64_t remote_buffer_size,
                           int num_producer_tasks, int64_t row_lo, int64_t row_hi);

int64_t oracle_matvec(const oracle_model *M, int P, const int64_t *sizes, const uint64_t *const *reps,
                      const double *const *xs, double *const *ys, int elt, int64_t remote_buffer_size,
                      int num_producer_tasks) {
  return matvec_impl(M, P, sizes, reps, xs, ys, elt, remote_buffer_size, num_producer_tasks, 0, -1);
}

/* The same product restricted to the SOURCE rows [row_lo, row_hi) of one locale: the bounded sample the CPU arm of
 * bench.py times (the contributions of those rows to y; chunk sizing as for the whole block). */
int64_t oracle_matvec_rows(const oracle_model *M, int64_t N, const uint64_t *reps, const double *x, double *y,
                           int elt, int64_t remote_buffer_size, int num_producer_tasks, int64_t row_lo,
                           int64_t row_hi) {
  const uint64_t *rp[1] = {reps};
  const double *xp[1] = {x};
  double *yp[1] = {y};
  return matvec_impl(M, 1, &N, rp, xp, yp, elt, remote_buffer_size, num_producer_tasks, row_lo, row_hi);
}

static int64_t matvec_impl(const oracle_model *M, int P, const int64_t *sizes, const uint64_t *const *reps,
                           const double *const *xs, double *const *ys, int elt, int64_t remote_buffer_size,
                           int num_producer_tasks, int64_t row_lo, int64_t row_hi) {
  const int slab = row_hi >= 0;   /* only with P == 1 */
  if (M->T_diag > 0) {
    for (int p = 0; p < P; ++p) {
      const int64_t N = sizes[p];
      const int64_t nchunks = 64;
      const int64_t r0 = slab ? row_lo : 0, r1 = slab ? row_hi : N;
#pragma omp parallel for schedule(dynamic, 1)
      for (int64_t c = 0; c < nchunks; ++c) {
        const int64_t lo = r0 + (r1 - r0) * c / nchunks, hi = r0 + (r1 - r0) * (c + 1) / nchunks;
        if (hi > lo)
          oracle_apply_diag_x1(M->T_diag, M->diag_v, M->diag_m, M->diag_r, M->diag_s, hi - lo,
                               reps[p] + lo, ys[p] + elt * lo, xs[p] + elt * lo, elt);
      }
    }
  }
  if (M->T_off == 0) return 0;
  int64_t failed = 0;
  const int64_t T = M->max_off_diag > 0 ? M->max_off_diag : 1;
  if (remote_buffer_size < T) remote_buffer_size = T;                          /* DMV:871 */
  for (int p = 0; p < P; ++p) {
    const int64_t N = slab ? row_hi - row_lo : sizes[p];          /* source rows handled here */
    const int64_t row0 = slab ? row_lo : 0;
    if (N <= 0) continue;
    int64_t num_chunks = (N * T + remote_buffer_size - 1) / remote_buffer_size; /* DMV:879-883 */
    if (num_chunks < 10 * num_producer_tasks) num_chunks = 10 * num_producer_tasks;
    if (num_chunks > N) num_chunks = N;
    const int64_t chunk_size = (N + num_chunks - 1) / num_chunks;               /* DMV:907 */
#pragma omp parallel
    {
      batched_operator bo;
      bo_init(&bo, chunk_size, M->T_off);   /* worst case: every elementary term emits */
      uint64_t *tmp_b = (uint64_t *)malloc(sizeof(uint64_t) * bo.capacity);
      c128 *tmp_c = (c128 *)malloc(sizeof(c128) * bo.capacity);
      int64_t offsets[257];
#pragma omp for schedule(dynamic, 1)
      for (int64_t c = 0; c < num_chunks; ++c) {
        /* chunks(0 ..# N, numChunks): DMV:905 */
        const int64_t lo = row0 + N * c / num_chunks, hi = row0 + N * (c + 1) / num_chunks;
        if (hi <= lo) continue;
        uint64_t *betas; c128 *coeffs; uint8_t *keys;
        const int64_t n = bo_compute_off_diag(&bo, M, P, hi - lo, reps[p] + lo, xs[p] + elt * lo,
                                              elt, &betas, &coeffs, &keys);     /* DMV:684 */
        radix_one_step(n, keys, offsets, betas, coeffs, tmp_b, tmp_c);          /* DMV:689 */
        for (int q = 0; q < P; ++q) {                                           /* DMV:695-729 */
          const int64_t k = offsets[q], cnt = offsets[q + 1] - k;
          const int64_t st = local_process(M, P, sizes[q], reps[q], ys[q], elt, betas + k, coeffs + k, cnt);
          if (st != 0) {
#pragma omp atomic write
            failed = 1;
          }
        }
      }
      free(tmp_b); free(tmp_c);
      bo_free(&bo);
    }
  }
  return failed ? -1 : 0;
}

/* Exposes BatchedOperator.computeOffDiag for one chunk (for unit tests of the three branches).
 * Outputs are copied into caller arrays of capacity count * T_off. */
int64_t oracle_compute_off_diag(const oracle_model *M, int num_locales, int64_t count,
                                const uint64_t *alphas, const double *xs, int elt, uint64_t *betas,
                                c128 *coeffs, uint8_t *keys, int64_t *offsets) {
  batched_operator bo;
  bo_init(&bo, count, M->T_off);
  uint64_t *b; c128 *c; uint8_t *k;
  const int64_t n = bo_compute_off_diag(&bo, M, num_locales, count, alphas, xs, elt, &b, &c, &k);
  memcpy(betas, b, (size_t)n * sizeof(uint64_t));
  memcpy(coeffs, c, (size_t)n * sizeof(c128));
  memcpy(keys, k, (size_t)n);
  memcpy(offsets, bo.offsets, (size_t)(count + 1) * sizeof(int64_t));
  bo_free(&bo);
  return n;
}

/* y[i] for a SAMPLE of rows of one block, column by column -- the at-size check of bench.py and tests/:
 *   y[i] = D(a_i) x[i] + sum_k conj(c_k) x[index(b_k)],   (b_k, c_k) = computeOffDiag(a_i, xs = 1)
 * i.e. row i of H from column i of H: valid for HERMITIAN operators (every model input of the reference).  It costs
 * |rows| * (T + 1) orbit scans instead of a whole product, so it also runs for bases of 10^7..10^8 states.
 * reps: the whole ascending basis, x its vector.  Returns 0, or -1 when a generated state is missing. */
int64_t oracle_expected_rows(const oracle_model *M, int64_t N, const uint64_t *reps, const double *x, int elt,
                             int64_t count, const int64_t *rows, double *y_out) {
  int64_t failed = 0;
#pragma omp parallel
  {
    batched_operator bo;
    bo_init(&bo, 1, M->T_off);
#pragma omp for schedule(dynamic, 16)
    for (int64_t q = 0; q < count; ++q) {
      const int64_t i = rows[q];
      const uint64_t alpha = reps[i];
      c128 acc = {0.0, 0.0};
      if (M->T_diag > 0) {
        double d[2] = {0.0, 0.0};
        oracle_apply_diag_x1(M->T_diag, M->diag_v, M->diag_m, M->diag_r, M->diag_s, 1, &alpha, d, x + elt * i, elt);
        acc.re = d[0];
        if (elt == 2) acc.im = d[1];
      }
      uint64_t *betas; c128 *coeffs; uint8_t *keys;
      const int64_t n = bo_compute_off_diag(&bo, M, 1, 1, &alpha, NULL, 1, &betas, &coeffs, &keys);
      for (int64_t k = 0; k < n; ++k) {
        const c128 c = coeffs[k];
        if (c.re == 0.0 && c.im == 0.0) continue;
        int64_t j;
        oracle_state_index(N, reps, 1, &betas[k], &j);
        if (j < 0) {
#pragma omp atomic write
          failed = 1;
          continue;
        }
        const c128 cc = {c.re, -c.im};
        if (elt == 1) {
          acc.re += cc.re * x[j];
        } else {
          const c128 xj = {x[2 * j], x[2 * j + 1]};
          const c128 t = c_mul(cc, xj);
          acc.re += t.re; acc.im += t.im;
        }
      }
      y_out[elt * q] = acc.re;
      if (elt == 2) y_out[2 * q + 1] = acc.im;
    }
    bo_free(&bo);
  }
  return failed ? -1 : 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
