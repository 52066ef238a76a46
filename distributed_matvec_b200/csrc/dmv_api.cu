// dmv_api.cu -- the C ABI of libdmv_b200.so (see include/dmv_b200.h) and the per-GPU context.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types only; the library itself is resolved with dlopen at dmv_comm_init

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/dmv_b200.h"
#include "dmv_host.h"

using namespace dmv;

namespace {

thread_local std::string g_last_error;

#define CUDA_CHECK(expr)                                                                        \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      throw std::runtime_error(std::string(#expr) + ": " + cudaGetErrorString(_e));            \
  } while (0)

#define API_BEGIN try {
#define API_END                                         \
  return 0;                                             \
  }                                                     \
  catch (const std::exception &e) {                     \
    g_last_error = e.what();                            \
    return 1;                                           \
  }                                                     \
  catch (...) {                                         \
    g_last_error = "unknown error";                     \
    return 1;                                           \
  }

template <typename T>
struct DevBuf {
  T *ptr = nullptr;
  size_t count = 0;
  void alloc(size_t n) {
    if (n <= count && ptr) return;
    release();
    if (n == 0) n = 1;
    CUDA_CHECK(cudaMalloc(&ptr, n * sizeof(T)));
    count = n;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    count = 0;
  }
  void upload(const std::vector<T> &h, cudaStream_t s) {
    alloc(h.size());
    if (!h.empty()) CUDA_CHECK(cudaMemcpyAsync(ptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  ~DevBuf() { release(); }
};

bool is_device_pointer(const void *p) {
  if (!p) return false;
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

// ---- NCCL through dlopen ------------------------------------------------------------------------
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi &nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
#define LOAD(sym) api.sym = reinterpret_cast<decltype(api.sym)>(dlsym(api.handle, "nccl" #sym))
    LOAD(GetUniqueId); LOAD(CommInitRank); LOAD(CommDestroy); LOAD(GroupStart); LOAD(GroupEnd);
    LOAD(Send); LOAD(Recv); LOAD(AllGather); LOAD(AllReduce); LOAD(GetErrorString);
#undef LOAD
  });
  if (!api.handle || !api.Send) throw std::runtime_error("NCCL (libnccl.so.2) is not available");
  return api;
}
#define NCCL_CHECK(expr)                                                                       \
  do {                                                                                         \
    ncclResult_t _r = (expr);                                                                  \
    if (_r != ncclSuccess)                                                                     \
      throw std::runtime_error(std::string(#expr) + ": " + nccl().GetErrorString(_r));        \
  } while (0)

// stages of one product (the coarse part of the reference's timing tree, DMV:1028-1052; the split of the fused kernels
// into the reference's inner timers -- applyOffDiag / stateInfo / indexing / accessing -- comes from tools/ncu_tree.py)
enum Timing { T_H2D = 0, T_GENERATE, T_EXCHANGE, T_ACCUMULATE, T_D2H, T_TOTAL, T_TABLE_FILL, T_COUNT };
const char *kTimingNames[T_COUNT] = {"h2d", "generate(diag+offdiag+local accumulate)", "exchange(all-to-all)",
                                     "accumulate(remote records)", "d2h", "total",
                                     "table refill (k_rows; part of generate)"};

}  // namespace

namespace {

// Flip-mask groups of an operator in look-up-table form (see LutGroup in dmv_device.cuh).
struct HostTables {
  std::vector<LutGroup> groups;
  std::vector<double> lut_re, lut_c;   // real parts only / interleaved complex
  std::vector<OffTerm> terms;
  std::vector<BpWord> bp;               // non-empty: bit-parallel emit test (see BpWord)
  bool any_generic = false, any_s_out = false;
};
struct DevTables {
  DevBuf<LutGroup> groups;
  DevBuf<double> lut_re, lut_c;
  DevBuf<OffTerm> terms;
  DevBuf<BpWord> bp;
  void upload(const HostTables &h, cudaStream_t s) {
    groups.upload(h.groups, s); lut_re.upload(h.lut_re, s); lut_c.upload(h.lut_c, s); terms.upload(h.terms, s);
    bp.upload(h.bp, s);
  }
};

// support of a group: union of the masks of its terms
uint64_t support_of(const std::vector<OffTerm> &terms) {
  uint64_t m = 0;
  for (const auto &t : terms) m |= t.m;
  return m;
}

HostTables build_tables(const std::map<uint64_t, std::vector<OffTerm>> &by_x) {
  HostTables H;
  // ---- can the whole operator use the bit-parallel emit test?  (every group: support <= 2 bits and a
  // common sign mask outside the support)
  struct Item { uint64_t x; const std::vector<OffTerm> *terms; int p0, p1; };
  std::vector<Item> items;
  bool bp_ok = !by_x.empty();
  for (const auto &kv : by_x) {
    const uint64_t sup = support_of(kv.second);
    const int k = __builtin_popcountll(sup);
    Item it{kv.first, &kv.second, 0, 0};
    if (k > 2) bp_ok = false;
    if (k >= 1) it.p0 = __builtin_ctzll(sup);
    it.p1 = (k == 2) ? 63 - __builtin_clzll(sup) : it.p0;
    const uint64_t s_out = kv.second.front().s & ~sup;
    for (const auto &t : kv.second) if ((t.s & ~sup) != s_out) bp_ok = false;
    items.push_back(it);
  }
  if (bp_ok) {
    // Two orders of the groups are tried.  (1) by the LOWEST site they act on: the groups acting only on high sites get
    // the high bits of the emit mask -- the 32 consecutive rows of a warp share their high bits, so k_gather, walking the
    // mask from the top, keeps its lanes in step (chains: two shifts per operand).  (2) by the distance between the two
    // sites, then position: few distinct shifts "group index - bit position" on two-dimensional lattices, where (1)
    // needs more than kBpPairs of them.
    const int n_words = (int)((items.size() + 63) / 64);
    std::vector<BpWord> words;
    for (int order = 0; order < 2; ++order) {
      std::stable_sort(items.begin(), items.end(), [order](const Item &a, const Item &b) {
        if (order == 0) return a.p0 != b.p0 ? a.p0 < b.p0 : a.p1 < b.p1;
        if (a.p1 - a.p0 != b.p1 - b.p0) return a.p1 - a.p0 < b.p1 - b.p0;
        return a.p0 < b.p0;
      });
      bp_ok = true;
      words.assign((size_t)n_words, BpWord{});
      for (auto &w : words) memset(&w, 0, sizeof(w));
      for (size_t g = 0; g < items.size() && bp_ok; ++g) {
        BpWord &W = words[g / 64];
        const int gl = (int)(g % 64);
        auto add = [&](int pos, BpPair *pairs, int32_t &n) {
          const int d = gl - pos;
          const uint32_t sl = d >= 0 ? (uint32_t)d : 0u, sr = d >= 0 ? 0u : (uint32_t)(-d);
          for (int k = 0; k < n; ++k)
            if (pairs[k].l == sl && pairs[k].r == sr) { pairs[k].m |= 1ull << gl; return; }
          if (n == kBpPairs) { bp_ok = false; return; }
          pairs[n].l = sl; pairs[n].r = sr; pairs[n].m = 1ull << gl; ++n;
        };
        add(items[g].p0, W.p0, W.n0);
        add(items[g].p1, W.p1, W.n1);
      }
      if (bp_ok) break;
    }
    if (bp_ok) {
      for (size_t g = 0; g < items.size(); ++g) {
        const Item &it = items[g];
        LutGroup grp{};
        grp.x = it.x;
        grp.first = (int32_t)H.terms.size();
        grp.count = (int32_t)it.terms->size();
        const uint64_t sup = support_of(*it.terms);
        grp.s_out = it.terms->front().s & ~sup;
        if (grp.s_out) H.any_s_out = true;
        grp.posk = (2ull << 48) | ((uint64_t)it.p1 << 8) | (uint64_t)it.p0;
        grp.lut_offset = (uint32_t)(4 * g);
        for (const auto &t : *it.terms) H.terms.push_back(t);
        for (int idx = 0; idx < 4; ++idx) {
          const int b0 = idx & 1, b1 = idx >> 1;
          double re = 0.0, im = 0.0;
          bool hit = false;
          if (!(it.p0 == it.p1 && b0 != b1)) {
            const uint64_t a = ((uint64_t)b0 << it.p0) | ((uint64_t)b1 << it.p1);
            for (const auto &t : *it.terms)
              if ((a & t.m) == t.r) {
                const double sg = (__builtin_popcountll(a & t.s & sup) & 1) ? -1.0 : 1.0;
                re += sg * t.v_re; im += sg * t.v_im; hit = true;
              }
          }
          if (hit && (re != 0.0 || im != 0.0)) {
            grp.emit_bits |= 1ull << idx;
            words[g / 64].tt[idx] |= 1ull << (g % 64);
          }
          H.lut_re.push_back(re);
          H.lut_c.push_back(re); H.lut_c.push_back(im);
        }
        H.groups.push_back(grp);
      }
      H.bp = words;
      return H;
    }
  }
  // ---- general layout: one LUT of 2^k entries per group (k <= 6), term-by-term evaluation otherwise
  for (const auto &kv : by_x) {
    LutGroup g{};
    g.x = kv.first;
    g.first = (int32_t)H.terms.size();
    g.count = (int32_t)kv.second.size();
    for (const auto &t : kv.second) H.terms.push_back(t);
    const uint64_t support = support_of(kv.second);
    const int k = __builtin_popcountll(support);
    bool lutable = k <= 6;
    const uint64_t s_out = kv.second.front().s & ~support;
    for (const auto &t : kv.second) lutable &= ((t.s & ~support) == s_out);
    if (lutable) {
      int pos[6] = {0, 0, 0, 0, 0, 0}, nb = 0;
      for (int b = 0; b < 64; ++b) if ((support >> b) & 1) pos[nb++] = b;
      g.posk = (uint64_t)k << 48;
      for (int b = 0; b < k; ++b) g.posk |= (uint64_t)pos[b] << (8 * b);
      g.s_out = s_out;
      g.lut_offset = (uint32_t)H.lut_re.size();
      for (int idx = 0; idx < (1 << k); ++idx) {
        uint64_t a = 0;
        for (int b = 0; b < k; ++b) if ((idx >> b) & 1) a |= 1ull << pos[b];
        double re = 0.0, im = 0.0;
        bool hit = false;
        for (const auto &t : kv.second)
          if ((a & t.m) == t.r) {
            const double sg = (__builtin_popcountll(a & t.s & support) & 1) ? -1.0 : 1.0;
            re += sg * t.v_re; im += sg * t.v_im; hit = true;
          }
        if (hit && (re != 0.0 || im != 0.0)) g.emit_bits |= 1ull << idx;
        H.lut_re.push_back(re);
        H.lut_c.push_back(re); H.lut_c.push_back(im);
      }
      if (s_out) H.any_s_out = true;
    } else {
      g.posk = 1ull << 56;
      H.any_generic = true;
    }
    H.groups.push_back(g);
  }
  return H;
}

}  // namespace

struct dmv_context {
  int device = 0, rank = 0, num_ranks = 1;
  // basis
  int n_sites = 0, hamming_weight = -1, spin_inversion = 0;
  bool has_permutations = false;
  Projection proj = PROJ_NONE;
  bool identity_index = false;
  uint64_t site_mask = 0;
  bool complex_coefficients = false;  // operator or characters are complex
  HostOrbitProgram host_orbit;
  DevBuf<uint64_t> d_orbit64;
  DevBuf<int32_t> d_orbit32;
  DevBuf<double> d_chars;
  DevBuf<uint16_t> d_canon_lut;
  DevBuf<uint64_t> d_canon_masks, d_cc_mask;
  DevBuf<uint32_t> d_canon_lut2;
  DevBuf<int32_t> d_cc_begin, d_cc_delta;
  DevBuf<uint16_t> d_tor_lutm;
  DevBuf<uint8_t> d_tor_frow;
  DevBuf<uint32_t> d_tor_luts;
  DevBuf<uint64_t> d_tor_net_mask;
  DevBuf<int32_t> d_tor_net_delta;
  int opt_canon = -1;    // -1 auto (block-rotation canonical form when the chain subgroup allows it), 0 walk the chain
  OrbitProgram orbit{};  // device view
  // operator
  std::vector<DiagTerm> h_diag;
  HostTables h_push, h_pull;          // column-traversal (scatter) / row-traversal (gather) tables
  DevTables d_push, d_pull;
  DevBuf<DiagTerm> d_diag;
  std::vector<DiagClass> h_diag_classes;  // bit-parallel part of the diagonal; h_diag is reordered: rest first
  DevBuf<DiagClass> d_diag_classes;
  int n_diag_rest = 0;
  size_t h_diag_kept = 0;   // number of diagonal terms of the operator (h_diag itself only keeps the non-class rest)
  // options
  int opt_mode = -1;    // -1 auto (pull when one rank owns the basis), 0 push (scatter), 1 pull (gather)
  int opt_index = -1;   // -1 auto, 0 directory search, 2 combinadic rank
  int opt_bitparallel = 1;  // 0: walk the groups one by one even when the bit-parallel test applies
  int opt_gather = -1;      // row traversal kernel: -1 auto (k_gather when it applies), 0 always the queued k_pull
  // k_gather applicability (set at context creation from the row-traversal tables)
  bool gather_ok = false, gather_narrow = false, gather_uniform = false;
  // k_rows applicability (bases with permutation symmetries, trivial characters, real bit-parallel operator) and its
  // hash table over this context's representatives (see table_slot in dmv_device.cuh)
  bool rows_ok = false;
  int opt_rows = -1;        // -1 auto (k_rows when it applies), 0 the queued k_pull
  int opt_rows_ctas = 2;    // k_rows: 2 CTAs per SM (122 registers, default) | 3 (80 registers, spills)
  int opt_gather_walk = 0;  // k_gather: 0 per-lane walk from the top bit (default), 1 group-major warp-uniform walk
                            // (measured slower), 2 per-lane walk from the bottom bit (round 1)
  DevBuf<unsigned char> d_table;
  DevBuf<unsigned char> d_mph_blocks, d_dense;   // dense index: perfect-hash blocks, dense table of (key, value) slots
  PerfectHash mph{};
  bool dense_index = false;
  int opt_rows_index = -1;   // -1 auto / 0 open-addressing table; 1 dense index through a perfect hash (measured slower:
                             // profiles/r02_rows_pipelines.md)
  DevBuf<uint32_t> d_slot_of;
  uint32_t table_slots = 0;
  int table_elt = 0;        // element type the slots are laid out for (0: not built)
  double gather_uni[2] = {0.0, 0.0};
  int index_mode = INDEX_DIRECTORY;
  DevBuf<uint32_t> d_binom, d_lin_a, d_lin_b;
  int lin_bits = 0;
  int binom_stride = 0;
  uint64_t rank_total = 0;
  // representatives of this rank
  int64_t n_states = -1;
  DevBuf<uint64_t> d_reps;
  DevBuf<double> d_norms;
  DevBuf<uint32_t> d_dir;
  uint64_t n_buckets = 0;
  int dir_shift = 0;
  // vectors staged for host callers
  DevBuf<double> d_x, d_y;
  // outgoing / incoming records
  bool planned = false;
  std::vector<int64_t> send_counts;        // [num_ranks]
  std::vector<int64_t> recv_counts;        // [num_ranks] (filled by dmv_comm plan exchange)
  std::vector<int64_t> h_out_offset;       // [num_ranks + 1]
  DevBuf<int64_t> d_out_offset;
  DevBuf<unsigned long long> d_out_count;
  DevBuf<uint64_t> d_out_betas, d_in_betas;
  DevBuf<double> d_out_coeffs, d_in_coeffs;
  int record_width = 2;                    // doubles per coefficient of the current buckets
  int plan_grid = 0;                       // CTAs of the planned launches (exact warp-private regions)
  int row_split = 1;                       // lanes per source state (chosen at plan time from the block size)
  bool peer_direct = false;                // records are stored straight into the peers' incoming buffers
  int ptr_width = 0;                       // record width the destination pointer table was built for
  int opt_exchange = -1;                   // -1 auto (peer-direct when possible), 0 NCCL send/recv, 1 peer-direct
  std::vector<void *> peer_betas, peer_coeffs;   // IPC-mapped incoming buffers of the peers
  std::vector<int64_t> my_offset_in_peer;         // first slot of MY region in every peer's incoming buffer
  DevBuf<int> d_barrier;
  DevBuf<unsigned long long> d_warp_counts;
  DevBuf<int64_t> d_warp_offsets, d_out_capacity;
  DevBuf<uint64_t *> d_out_betas_ptr;
  DevBuf<double *> d_out_coeffs_ptr;
  std::vector<uint64_t *> h_out_betas_ptr;   // where the records for every destination go (local bucket or peer)
  std::vector<double *> h_out_coeffs_ptr;
  int64_t number_terms = 0;
  DevBuf<unsigned long long> d_status;
  // streams
  cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
  static constexpr int kCopyChunks = 8;
  cudaEvent_t ev_chunk[kCopyChunks] = {};
  cudaEvent_t ev[T_COUNT + 2] = {};
  cudaEvent_t ev_fill[2] = {};
  bool fill_timed = false;
  double timings[T_COUNT] = {};
  // communicator
  ncclComm_t comm = nullptr;
  // replicated-x product (exchange = 2, see setup_replicated): a single-rank twin context holding the WHOLE basis,
  // the slot of every global state in the all-gathered x, and the gathered x itself
  std::vector<double> k_off_v, k_diag_v;                      // copies of the creation arguments
  std::vector<uint64_t> k_off_m, k_off_r, k_off_x, k_off_s, k_diag_m, k_diag_r, k_diag_s;
  std::vector<int32_t> k_perms;
  std::vector<uint8_t> k_flips;
  std::vector<double> k_chars;
  int64_t k_group_order = 0;
  dmv_context *global = nullptr;
  DevBuf<uint32_t> d_pos;
  int64_t repl_block = 0;        // slot size per rank in the gathered x (the largest block)
  DevBuf<double> d_xcat;
  bool replicated = false, exchange_decided = false, timeline_replicated = false;
  // peer-direct all-gather of x (launch_push_block): the peers' gathered vectors (two buffers, alternating by epoch) and
  // flag words mapped with CUDA IPC
  bool peer_gather = false;
  int opt_peer_gather = -1;                 // -1 auto, 0 NCCL all-gather
  std::vector<void *> peer_xcat, peer_flagmem;
  DevBuf<unsigned> d_flags, d_push_done;    // [num_ranks] epochs raised by the peers; CTA counter of k_push_block
  DevBuf<void *> d_peer_slot[2];            // [num_ranks] slot `rank` of every rank's buffer b
  DevBuf<unsigned *> d_peer_flags;          // [num_ranks]
  int peer_slot_elt = 0;                    // element width the slot pointers were computed for
  unsigned gather_epoch = 0;
  // record exchange in overlapped ROUNDS (peer-direct records; reference DMV:638-661, 818-852, 957-1011): the rows are cut
  // into R rounds; round r's records land in the owners' buffers while round r + 1 is being generated, and the owner
  // accumulates round r on a second stream as soon as every sender has raised its flag for it
  struct Rounds {
    bool ready = false, tried = false;
    int R = 0, grid = 0, row_split = 1;
    std::vector<int64_t> row_begin;           // [R + 1]
    DevBuf<int64_t> d_warp_offsets;           // [R][warps][P]: first slot of every warp inside MY region of (round, dest)
    DevBuf<int64_t> d_capacity;               // [R][P]
    std::vector<int64_t> in_slice;            // [R + 1]: rounds inside my incoming buffer (records)
    std::vector<int64_t> my_off;              // [R][P]: my region of round r inside rank q's incoming buffer
    int64_t in_total = 0;
    std::vector<int64_t> peer_total;          // [P]: in_total of every rank (start of its second buffer)
    DevBuf<uint64_t> d_in_betas;              // two buffers (alternating products) of in_total records
    DevBuf<double> d_in_coeffs;               // two doubles per record
    std::vector<void *> peer_betas, peer_coeffs, peer_flags;
    DevBuf<unsigned> d_flags;                 // [P] raised by the senders: product * R + round + 1
    DevBuf<unsigned *> d_peer_flags;
    DevBuf<uint64_t *> d_bptr;                // [2][R][P]
    DevBuf<double *> d_cptr;
    int ptr_width = 0;
    unsigned seq = 0;
    cudaStream_t acc_stream = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_done = nullptr;
    int64_t terms = 0;
  } rounds;
  int opt_rounds = -1;                        // -1 auto, 0 / 1 off (generate everything, fence, accumulate), R > 1

  // Lanczos work space (dmv_lanczos)
  DevBuf<double> lz_v[4];
  DevBuf<double> lz_scal;

  ~dmv_context() {
    delete global;
    for (void *q : peer_betas) if (q) cudaIpcCloseMemHandle(q);
    for (void *q : peer_coeffs) if (q) cudaIpcCloseMemHandle(q);
    for (void *q : peer_xcat) if (q) cudaIpcCloseMemHandle(q);
    for (void *q : rounds.peer_betas) if (q) cudaIpcCloseMemHandle(q);
    for (void *q : rounds.peer_coeffs) if (q) cudaIpcCloseMemHandle(q);
    for (void *q : rounds.peer_flags) if (q) cudaIpcCloseMemHandle(q);
    if (rounds.acc_stream) cudaStreamDestroy(rounds.acc_stream);
    if (rounds.ev_begin) cudaEventDestroy(rounds.ev_begin);
    if (rounds.ev_done) cudaEventDestroy(rounds.ev_done);
    for (void *q : peer_flagmem) if (q) cudaIpcCloseMemHandle(q);
    if (comm) nccl().CommDestroy(comm);
    for (auto &e : ev) if (e) cudaEventDestroy(e);
    for (auto &e : ev_fill) if (e) cudaEventDestroy(e);
    for (auto &e : ev_chunk) if (e) cudaEventDestroy(e);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (own_stream) cudaStreamDestroy(own_stream);
  }
};

namespace {

bool use_gather(const dmv_context *ctx) {   // the lean row-gather kernel applies and is not switched off
  return ctx->gather_ok && ctx->opt_gather != 0 && ctx->opt_bitparallel != 0 && ctx->proj != PROJ_GROUP;
}
bool use_rows(const dmv_context *ctx) {   // the pipelined row kernel for bases with permutation symmetries
  return ctx->rows_ok && ctx->opt_rows != 0 && ctx->opt_bitparallel != 0 && ctx->orbit.trivial_characters;
}
bool use_pull(const dmv_context *ctx) {
  // auto: one rank, bit-parallel operator, no permutation symmetries -> k_gather (rows, no atomics, see
  // dmv_gather.cu); everything else -> push (k_generate).  "mode" = 1 forces the row traversal (k_gather
  // when it applies, else the queued k_pull), "mode" = 0 the scatter form.
  if (ctx->num_ranks != 1) return false;
  if (ctx->opt_mode == 1) return true;
  return ctx->opt_mode == -1 && (use_gather(ctx) || use_rows(ctx));
}

void use_device(const dmv_context *ctx) { CUDA_CHECK(cudaSetDevice(ctx->device)); }

bool complex_values(const dmv_context *ctx, int elt) { return elt == DMV_C128 || ctx->complex_coefficients; }

KernelParams base_params(dmv_context *ctx) {
  KernelParams p{};
  p.index.reps = ctx->d_reps.ptr;
  p.index.n = ctx->n_states;
  p.index.dir = ctx->d_dir.ptr;
  p.index.n_buckets = ctx->n_buckets;
  p.index.shift = ctx->dir_shift;
  p.index.mode = ctx->index_mode;
  p.index.binom = ctx->d_binom.ptr;
  p.index.stride = ctx->binom_stride;
  p.index.n_sites = ctx->n_sites;
  p.index.weight = ctx->hamming_weight;
  p.index.site_mask = ctx->site_mask;
  p.index.lin_a = ctx->d_lin_a.ptr;
  p.index.lin_b = ctx->d_lin_b.ptr;
  p.index.lin_bits = ctx->lin_bits;
  p.rank_total = ctx->rank_total;
  p.norms = ctx->d_norms.ptr;
  p.diag = ctx->d_diag.ptr;     p.n_diag = (int)ctx->h_diag_kept;
  p.diag_classes = ctx->d_diag_classes.ptr; p.n_diag_classes = (int)ctx->h_diag_classes.size();
  p.n_diag_rest = ctx->n_diag_rest;
  p.orbit = ctx->orbit;
  p.site_mask = ctx->site_mask;
  p.inversion_character = (double)ctx->spin_inversion;
  p.rank = ctx->rank;
  p.num_ranks = ctx->num_ranks;
  p.out_betas = ctx->d_out_betas.ptr;
  p.out_coeffs = ctx->d_out_coeffs.ptr;
  p.out_offset = ctx->d_out_offset.ptr;
  p.out_count = ctx->d_out_count.ptr;
  p.grid_blocks = ctx->num_ranks > 1 ? ctx->plan_grid : 0;
  p.row_split = ctx->row_split;
  p.warp_offsets = ctx->d_warp_offsets.ptr;
  p.warp_counts = ctx->d_warp_counts.ptr;
  p.out_betas_ptr = ctx->d_out_betas_ptr.ptr;
  p.out_coeffs_ptr = ctx->d_out_coeffs_ptr.ptr;
  p.out_capacity = ctx->d_out_capacity.ptr;
  p.status = ctx->d_status.ptr;
  p.row_begin = 0;
  p.row_end = ctx->n_states;
  p.gather_walk = ctx->opt_gather_walk;
  p.rows_ctas = ctx->opt_rows_ctas;
  return p;
}

// Split the diagonal into bit-parallel classes (m == 0, two sign bits, equal coefficient) and the rest.
void build_diag_classes(dmv_context *ctx) {
  std::vector<DiagTerm> rest;
  std::map<std::pair<double, double>, std::vector<DiagTerm>> by_v;
  for (const auto &d : ctx->h_diag) {
    if (d.m == 0 && d.r == 0 && __builtin_popcountll(d.s) == 2) by_v[{d.v_re, d.v_im}].push_back(d);
    else rest.push_back(d);
  }
  for (auto &kv : by_v) {
    auto &terms = kv.second;
    std::stable_sort(terms.begin(), terms.end(), [](const DiagTerm &a, const DiagTerm &b) {
      const int a0 = __builtin_ctzll(a.s), a1 = 63 - __builtin_clzll(a.s);
      const int b0 = __builtin_ctzll(b.s), b1 = 63 - __builtin_clzll(b.s);
      if (a1 - a0 != b1 - b0) return a1 - a0 < b1 - b0;
      return a0 < b0;
    });
    for (size_t first = 0; first < terms.size(); first += 64) {
      const size_t n = std::min<size_t>(64, terms.size() - first);
      DiagClass D;
      memset(&D, 0, sizeof(D));
      D.v_re = kv.first.first; D.v_im = kv.first.second;
      bool ok = true;
      for (size_t t = 0; t < n && ok; ++t) {
        const uint64_t sbits = terms[first + t].s;
        const int pos[2] = {__builtin_ctzll(sbits), 63 - __builtin_clzll(sbits)};
        BpPair *pairs[2] = {D.p0, D.p1};
        int32_t *cnt[2] = {&D.n0, &D.n1};
        for (int b = 0; b < 2 && ok; ++b) {
          const int d = (int)t - pos[b];
          const uint32_t sl = d >= 0 ? (uint32_t)d : 0u, sr = d >= 0 ? 0u : (uint32_t)(-d);
          int k = 0;
          for (; k < *cnt[b]; ++k)
            if (pairs[b][k].l == sl && pairs[b][k].r == sr) break;
          if (k == *cnt[b]) {
            if (k == kBpPairs) { ok = false; break; }
            pairs[b][k].l = sl; pairs[b][k].r = sr; pairs[b][k].m = 0; ++*cnt[b];
          }
          pairs[b][k].m |= 1ull << t;
        }
      }
      if (ok) {
        D.count = (int32_t)n;
        D.mask = n == 64 ? ~0ull : ((1ull << n) - 1);
        ctx->h_diag_classes.push_back(D);
      } else {
        for (size_t t = 0; t < n; ++t) rest.push_back(terms[first + t]);
      }
    }
  }
  // reorder: the terms evaluated one by one come first
  ctx->n_diag_rest = (int)rest.size();
  std::vector<DiagTerm> reordered = rest;
  ctx->h_diag_kept = ctx->h_diag.size();
  ctx->h_diag = reordered;
}

// point the kernel at the column-traversal (push) or row-traversal (pull) tables
void select_tables(dmv_context *ctx, KernelParams &p, bool pull, bool complex_vals) {
  const HostTables &h = pull ? ctx->h_pull : ctx->h_push;
  DevTables &d = pull ? ctx->d_pull : ctx->d_push;
  p.groups = d.groups.ptr; p.n_groups = (int)h.groups.size();
  p.lut = complex_vals ? d.lut_c.ptr : d.lut_re.ptr; p.n_lut = (int)h.lut_re.size();
  p.terms = d.terms.ptr; p.n_terms = (int)h.terms.size();
  p.any_generic = h.any_generic ? 1 : 0;
  p.any_s_out = h.any_s_out ? 1 : 0;
  p.bp = d.bp.ptr; p.n_bp = (ctx->opt_bitparallel != 0) ? (int)h.bp.size() : 0;
}

void require_states(const dmv_context *ctx) {
  if (ctx->n_states < 0) throw std::runtime_error("basis is not built");  // src/ForeignTypes.chpl:113-114
}

void check_status(dmv_context *ctx) {
  unsigned long long st[4];
  CUDA_CHECK(cudaMemcpyAsync(st, ctx->d_status.ptr, sizeof(st), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  if (st[0] != 0 || st[2] != 0 || st[3] != 0) {
    CUDA_CHECK(cudaMemsetAsync(ctx->d_status.ptr, 0, 4 * sizeof(unsigned long long), ctx->stream));
    char buf[256];
    if (st[3] != 0)
      snprintf(buf, sizeof(buf), "peer-direct all-gather of x: a rank did not raise its flag within the time limit");
    else if (st[2] != 0)
      snprintf(buf, sizeof(buf), "outgoing bucket overflow (%llu records): plan is stale", st[2]);
    else  // message of the reference: DMV:116-118
      snprintf(buf, sizeof(buf), "invalid index: -1 for state %llu (%llu such records): the operator does "
               "not respect the basis symmetries or the representatives are incomplete", st[1], st[0]);
    throw std::runtime_error(buf);
  }
}

// binomial table for the combinadic ranking of fixed-Hamming-weight states
struct Binomials {
  uint64_t c[65][65];
  Binomials() {
    for (int n = 0; n <= 64; ++n)
      for (int k = 0; k <= 64; ++k) {
        if (k == 0 || k == n) c[n][k] = (k <= n) ? 1 : 0;
        else if (k > n) c[n][k] = 0;
        else {
          const unsigned __int128 v = (unsigned __int128)c[n - 1][k - 1] + c[n - 1][k];
          c[n][k] = v > (unsigned __int128)~0ull ? ~0ull : (uint64_t)v;
        }
      }
  }
};
const Binomials &binom() { static Binomials b; return b; }

// Which state -> index kernel applies (the reference's per-basis `state_index_kernel`, FFI:90-93).
//   auto (-1): identity when it applies; two-table Lin lookup for full fixed-Hamming bases on one rank
//   (<= 40 sites); directory search otherwise.  0 forces the directory, 2 the combinadic rank, 3 Lin.
void select_index_mode(dmv_context *ctx) {
  ctx->index_mode = INDEX_DIRECTORY;
  if (ctx->identity_index && ctx->num_ranks == 1) { ctx->index_mode = INDEX_IDENTITY; return; }
  const int n = ctx->n_sites, w = ctx->hamming_weight;
  const int want = ctx->opt_index;
  if (want == 0) return;
  const bool eligible = ctx->num_ranks == 1 && w >= 0 && ctx->proj != PROJ_GROUP;
  if (!eligible) return;
  const uint64_t total = binom().c[n][w];
  const uint64_t expect = (ctx->proj == PROJ_INVERSION) ? total / 2 : total;
  if ((uint64_t)ctx->n_states != expect || total >= (1ull << 32)) return;
  StateIndex ix{};
  ix.reps = ctx->d_reps.ptr; ix.n = ctx->n_states; ix.n_sites = n; ix.weight = w; ix.site_mask = ctx->site_mask;
  if (want == 2) {
    const int stride = w + 2;
    std::vector<uint32_t> table((size_t)n * stride);
    for (int pos = 0; pos < n; ++pos)
      for (int k = 0; k < stride; ++k)
        table[(size_t)pos * stride + k] = (uint32_t)std::min<uint64_t>(binom().c[pos][k], 0xffffffffull);
    ctx->d_binom.upload(table, ctx->stream);
    ctx->binom_stride = stride;
    ix.mode = INDEX_RANK; ix.binom = ctx->d_binom.ptr; ix.stride = stride;
  } else {
    if (n > 40) return;
    // Lin tables: states ascending = (hi, lo) lexicographic; index = Ja[hi] + Jb[lo]
    const int lb = n / 2, hb = n - lb;
    std::vector<uint32_t> ja((size_t)1 << hb), jb((size_t)1 << lb);
    uint64_t running = 0;
    for (uint64_t hi = 0; hi < (1ull << hb); ++hi) {
      const int k = w - __builtin_popcountll(hi);
      ja[hi] = (uint32_t)std::min<uint64_t>(running, 0xffffffffull);
      if (k >= 0 && k <= lb) running += binom().c[lb][k];
    }
    std::vector<uint32_t> counter(lb + 1, 0);
    for (uint64_t lo = 0; lo < (1ull << lb); ++lo) jb[lo] = counter[__builtin_popcountll(lo)]++;
    ctx->d_lin_a.upload(ja, ctx->stream);
    ctx->d_lin_b.upload(jb, ctx->stream);
    ctx->lin_bits = lb;
    ix.mode = INDEX_LIN; ix.lin_a = ctx->d_lin_a.ptr; ix.lin_b = ctx->d_lin_b.ptr; ix.lin_bits = lb;
  }
  ctx->rank_total = total;
  // the block must be exactly the first `expect` fixed-weight states: index(reps[i]) == i for all i
  CUDA_CHECK(cudaMemsetAsync(ctx->d_status.ptr, 0, 4 * sizeof(unsigned long long), ctx->stream));
  launch_verify_rank(ix, ctx->d_status.ptr, ctx->stream);
  unsigned long long bad = 0;
  CUDA_CHECK(cudaMemcpyAsync(&bad, ctx->d_status.ptr, sizeof(bad), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  CUDA_CHECK(cudaMemsetAsync(ctx->d_status.ptr, 0, 4 * sizeof(unsigned long long), ctx->stream));
  if (bad == 0) ctx->index_mode = ix.mode;
}

void install_directory(dmv_context *ctx) {
  const int64_t n = ctx->n_states;
  uint64_t max_rep = 0;
  if (n > 0) CUDA_CHECK(cudaMemcpyAsync(&max_rep, ctx->d_reps.ptr + (n - 1), 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  int bits = 0;
  while (bits < 64 && (max_rep >> bits) != 0) ++bits;
  // about 4 states per bucket on average, at least 2^10 and at most 2^26 buckets
  int want = 10;
  while (want < 26 && (1ll << (want + 2)) < n) ++want;
  int shift = bits > want ? bits - want : 0;
  ctx->dir_shift = shift;
  ctx->n_buckets = (max_rep >> shift) + 1;
  ctx->d_dir.alloc(2 * ctx->n_buckets + 2);
  launch_build_directory(ctx->d_reps.ptr, n, ctx->d_dir.ptr, ctx->n_buckets, shift, ctx->stream);
  ctx->planned = false;
  ctx->table_elt = 0;
  // a new block also invalidates the exchange set-up: the replicated-x twin / slot table and the record plan
  ctx->exchange_decided = false;
  ctx->replicated = false;
  ctx->repl_block = 0;
  if (ctx->global) { delete ctx->global; ctx->global = nullptr; }
  ctx->d_pos.release();
  for (auto &q : ctx->peer_xcat) if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
  for (auto &q : ctx->peer_flagmem) if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
  ctx->peer_gather = false;
  ctx->peer_slot_elt = 0;
  ctx->d_xcat.release();
  for (auto *v : {&ctx->rounds.peer_betas, &ctx->rounds.peer_coeffs, &ctx->rounds.peer_flags})
    for (auto &q : *v) if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
  ctx->rounds.ready = false;
  ctx->rounds.tried = false;
  std::fill(ctx->recv_counts.begin(), ctx->recv_counts.end(), -1);
  select_index_mode(ctx);
}

void upload_orbit(dmv_context *ctx) {
  const HostOrbitProgram &H = ctx->host_orbit;
  std::vector<uint64_t> h64 = H.benes_mask;
  h64.insert(h64.end(), H.step_mask.begin(), H.step_mask.end());
  if (h64.size() & 1) h64.push_back(0);   // 16-byte alignment of the packed 32-bit steps
  const size_t off_pack64 = h64.size();
  h64.insert(h64.end(), H.step_pack64.begin(), H.step_pack64.end());
  if (h64.size() & 1) h64.push_back(0);
  const size_t off_pack32 = h64.size();
  for (size_t i = 0; i + 1 < H.step_pack32.size(); i += 2)
    h64.push_back((uint64_t)H.step_pack32[i] | ((uint64_t)H.step_pack32[i + 1] << 32));
  std::vector<int32_t> h32 = H.benes_delta;
  h32.insert(h32.end(), H.step_shift.begin(), H.step_shift.end());
  ctx->d_orbit64.upload(h64, ctx->stream);
  ctx->d_orbit32.upload(h32, ctx->stream);
  ctx->d_chars.upload(H.characters, ctx->stream);
  OrbitProgram P = H.view();
  P.benes_mask = ctx->d_orbit64.ptr;
  P.step_mask = ctx->d_orbit64.ptr + H.benes_mask.size();
  P.benes_delta = ctx->d_orbit32.ptr;
  P.step_shift = ctx->d_orbit32.ptr + H.benes_delta.size();
  P.characters = reinterpret_cast<const double2 *>(ctx->d_chars.ptr);
  P.simple = H.simple;
  P.step_pack64 = ctx->d_orbit64.ptr + off_pack64;
  P.step_pack32 = H.step_pack32.empty() ? nullptr : reinterpret_cast<const uint4 *>(ctx->d_orbit64.ptr + off_pack32);
  ctx->d_canon_lut.upload(H.canon_lut, ctx->stream);
  ctx->d_canon_masks.upload(H.canon_masks, ctx->stream);
  P.canon_lut = ctx->d_canon_lut.ptr;
  P.canon_masks = ctx->d_canon_masks.ptr;
  ctx->d_canon_lut2.upload(H.canon_lut2, ctx->stream);
  ctx->d_cc_begin.upload(H.cc_begin, ctx->stream);
  ctx->d_cc_mask.upload(H.cc_mask, ctx->stream);
  ctx->d_cc_delta.upload(H.cc_delta, ctx->stream);
  P.canon_lut2 = H.canon_lut2.empty() ? nullptr : ctx->d_canon_lut2.ptr;
  P.cc_begin = ctx->d_cc_begin.ptr;
  P.cc_mask = ctx->d_cc_mask.ptr;
  P.cc_delta = ctx->d_cc_delta.ptr;
  ctx->d_tor_lutm.upload(H.tor_lutm, ctx->stream);
  ctx->d_tor_luts.upload(H.tor_luts, ctx->stream);
  ctx->d_tor_net_mask.upload(H.tor_net_mask, ctx->stream);
  ctx->d_tor_net_delta.upload(H.tor_net_delta, ctx->stream);
  P.tor_lutm = ctx->d_tor_lutm.ptr;
  P.tor_luts = ctx->d_tor_luts.ptr;
  ctx->d_tor_frow.upload(H.tor_frow, ctx->stream);
  P.tor_frow = ctx->d_tor_frow.ptr;
  P.tor_net_mask = ctx->d_tor_net_mask.ptr;
  P.tor_net_delta = ctx->d_tor_net_delta.ptr;
  if (ctx->opt_canon >= 0) { P.tor_mode = 0; P.chain_dihedral = 0; }   // 1: round-1 forms (coset chain / four run searches)
  if (ctx->opt_canon == 2) { P.canon_lut2 = nullptr; P.cc_n = 0; P.cc_stages = 0; }   // first version: single-block LUT, independent networks
  if (ctx->opt_canon == 0) P.canon_mode = 0;
  ctx->orbit = P;
}

// rank of a fixed-weight state among states of the same weight in ascending order
// (what ls_hs_fixed_hamming_state_to_index computes, reference src/FFI.chpl:165)
uint64_t fixed_hamming_rank(uint64_t s) {
  uint64_t r = 0;
  int k = 0;
  while (s) {
    const int pos = __builtin_ctzll(s);
    ++k;
    r += binom().c[pos][k];
    s &= s - 1;
  }
  return r;
}
uint64_t fixed_hamming_unrank(uint64_t r, int weight) {  // ls_hs_fixed_hamming_index_to_state
  uint64_t s = 0;
  for (int k = weight; k >= 1; --k) {
    int pos = k - 1;
    while (pos + 1 <= 63 && binom().c[pos + 1][k] <= r) ++pos;
    r -= binom().c[pos][k];
    s |= 1ull << pos;
  }
  return s;
}

void zero_y_if_diag(dmv_context *ctx, int elt, void *y) {
  // DMV:1062-1063: with diagonal terms y is overwritten by D x, otherwise it is accumulated into
  if (ctx->h_diag_kept > 0)
    CUDA_CHECK(cudaMemsetAsync(y, 0, (size_t)ctx->n_states * 8 * elt, ctx->stream));
}

struct VecStage {  // x / y either used in place (device pointers) or staged through context buffers
  const void *x_dev; void *y_dev; bool y_host; void *y_user; size_t bytes;
  const void *x_host_pending;   // host x whose upload is pipelined with generation (push traversal)
};
VecStage stage_vectors(dmv_context *ctx, int elt, const void *x, void *y) {
  VecStage v{};
  v.bytes = (size_t)ctx->n_states * 8 * elt;
  CUDA_CHECK(cudaEventRecord(ctx->ev[0], ctx->stream));
  if (is_device_pointer(x)) v.x_dev = x;
  else {
    ctx->d_x.alloc((size_t)ctx->n_states * elt);
    v.x_dev = ctx->d_x.ptr;
    // the column traversal only reads x[i] of the rows it is generating: upload in row chunks on a copy
    // stream and start generating as soon as the first chunk has landed (see do_generate)
    if (!use_pull(ctx) && ctx->num_ranks == 1 && ctx->n_states >= (1 << 16)) v.x_host_pending = x;
    else CUDA_CHECK(cudaMemcpyAsync(ctx->d_x.ptr, x, v.bytes, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (is_device_pointer(y)) { v.y_dev = y; v.y_host = false; }
  else {
    ctx->d_y.alloc((size_t)ctx->n_states * elt);
    v.y_dev = ctx->d_y.ptr; v.y_host = true; v.y_user = y;
    if (ctx->h_diag_kept == 0)  // y is accumulated into: bring the caller's y over
      CUDA_CHECK(cudaMemcpyAsync(ctx->d_y.ptr, y, v.bytes, cudaMemcpyHostToDevice, ctx->stream));
  }
  CUDA_CHECK(cudaEventRecord(ctx->ev[1], ctx->stream));
  return v;
}
void finish_vectors(dmv_context *ctx, const VecStage &v) {
  CUDA_CHECK(cudaEventRecord(ctx->ev[4], ctx->stream));
  if (v.y_host) CUDA_CHECK(cudaMemcpyAsync(v.y_user, v.y_dev, v.bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaEventRecord(ctx->ev[5], ctx->stream));
}

void upload_out_pointers(dmv_context *ctx) {
  ctx->d_out_betas_ptr.upload(ctx->h_out_betas_ptr, ctx->stream);
  ctx->d_out_coeffs_ptr.upload(ctx->h_out_coeffs_ptr, ctx->stream);
}

// One counting pass.  Record counts do not depend on x, and the grid-stride tile loop is deterministic,
// so the pass yields (a) the exact number of records for every destination and (b) for num_ranks <= 32
// the exact share of every warp, from which each warp gets a private, exactly sized slice of every
// destination region (prefix sums): the real pass needs no slot-claim atomics at all.
void do_plan(dmv_context *ctx) {
  require_states(ctx);
  const int P = ctx->num_ranks;
  const bool exact_regions = P <= 32;
  ctx->row_split = choose_row_split(ctx->n_states, (int)ctx->h_push.groups.size());
  ctx->plan_grid = planned_grid(ctx->n_states, ctx->row_split);
  const size_t n_warps = (size_t)ctx->plan_grid * kWarpsPerCta;
  ctx->d_out_count.alloc(P);
  ctx->d_warp_counts.alloc(n_warps * P);
  CUDA_CHECK(cudaMemsetAsync(ctx->d_out_count.ptr, 0, sizeof(unsigned long long) * P, ctx->stream));
  CUDA_CHECK(cudaMemsetAsync(ctx->d_warp_counts.ptr, 0, sizeof(unsigned long long) * n_warps * P, ctx->stream));
  KernelParams p = base_params(ctx);
  p.grid_blocks = ctx->plan_grid;
  select_tables(ctx, p, false, ctx->complex_coefficients);
  // counting pass: element type does not matter
  launch_generate(p, ctx->proj, ctx->complex_coefficients, false, /*count_only=*/true, ctx->stream);
  std::vector<unsigned long long> counts(P, 0);
  std::vector<int64_t> warp_offsets(n_warps * P, 0);
  if (exact_regions) {
    std::vector<unsigned long long> wc(n_warps * P);
    CUDA_CHECK(cudaMemcpyAsync(wc.data(), ctx->d_warp_counts.ptr, sizeof(unsigned long long) * wc.size(),
                               cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    for (int d = 0; d < P; ++d)
      for (size_t w = 0; w < n_warps; ++w) {
        warp_offsets[w * P + d] = (int64_t)counts[d];
        counts[d] += wc[w * P + d];
      }
  } else {
    CUDA_CHECK(cudaMemcpyAsync(counts.data(), ctx->d_out_count.ptr, sizeof(unsigned long long) * P,
                               cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  }
  ctx->d_warp_offsets.upload(warp_offsets, ctx->stream);
  ctx->send_counts.assign(P, 0);
  ctx->number_terms = 0;
  for (int d = 0; d < P; ++d) { ctx->send_counts[d] = (int64_t)counts[d]; ctx->number_terms += (int64_t)counts[d]; }
  ctx->h_out_offset.assign(P + 1, 0);
  std::vector<int64_t> capacity(P, 0);
  for (int d = 0; d < P; ++d) {
    capacity[d] = (d == ctx->rank) ? 0 : ctx->send_counts[d];
    ctx->h_out_offset[d + 1] = ctx->h_out_offset[d] + capacity[d];
  }
  ctx->d_out_offset.upload(ctx->h_out_offset, ctx->stream);
  ctx->d_out_capacity.upload(capacity, ctx->stream);
  const int64_t total_out = ctx->h_out_offset[P];
  ctx->d_out_betas.alloc((size_t)total_out);
  ctx->d_out_coeffs.alloc((size_t)total_out * 2);
  // by default the records of destination d go to the local bucket d (sent with NCCL afterwards);
  // the coefficient base assumes the widest record (re-derived per product, see do_generate)
  ctx->h_out_betas_ptr.assign(P, nullptr);
  ctx->h_out_coeffs_ptr.assign(P, nullptr);
  ctx->peer_direct = false;
  ctx->ptr_width = 0;
  ctx->recv_counts.assign(P, -1);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  ctx->planned = true;
}

// hash table of k_rows over ctx's representatives: keys once per basis and element type, values once per product
void ensure_table(dmv_context *ctx, int elt) {
  if (ctx->table_elt == elt) return;
  const int64_t n = ctx->n_states;
  cudaStream_t st = ctx->stream;
  const bool ce = elt == DMV_C128;
  // ---- dense index: two perfect-hash levels of 4 bits per state; what they cannot place goes to the table below
  const uint64_t *left_keys = ctx->d_reps.ptr;
  int64_t n_left = n;
  DevBuf<uint64_t> d_left[2];
  ctx->dense_index = ctx->opt_rows_index == 1 && n >= 1;
  ctx->mph = PerfectHash{};
  if (ctx->dense_index) {
    if (n >= 2147483647ll) throw std::runtime_error("k_rows: more than 2^31 states");
    std::vector<unsigned long long> bits;      // seen & ~collide of both levels, 3 words per block
    uint32_t nb[2] = {0, 0};
    DevBuf<unsigned long long> d_count;
    d_count.alloc(1);
    for (int level = 0; level < 2 && n_left > 0; ++level) {
      nb[level] = (uint32_t)std::max<int64_t>(1, (4 * n_left + kMphBits - 1) / kMphBits);
      const size_t words = (size_t)nb[level] * 3;
      DevBuf<unsigned long long> d_seen, d_coll;
      d_seen.alloc(words); d_coll.alloc(words);
      CUDA_CHECK(cudaMemsetAsync(d_seen.ptr, 0, words * 8, st));
      CUDA_CHECK(cudaMemsetAsync(d_coll.ptr, 0, words * 8, st));
      CUDA_CHECK(cudaMemsetAsync(d_count.ptr, 0, 8, st));
      launch_mph_mark(left_keys, n_left, level, nb[level], d_seen.ptr, d_coll.ptr, st);
      d_left[level].alloc((size_t)std::max<int64_t>(1, n_left));
      launch_mph_compact(left_keys, n_left, level, nb[level], d_coll.ptr, d_left[level].ptr, d_count.ptr, st);
      std::vector<unsigned long long> seen(words), coll(words);
      unsigned long long cnt = 0;
      CUDA_CHECK(cudaMemcpyAsync(seen.data(), d_seen.ptr, words * 8, cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaMemcpyAsync(coll.data(), d_coll.ptr, words * 8, cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaMemcpyAsync(&cnt, d_count.ptr, 8, cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
      for (size_t w = 0; w < words; ++w) bits.push_back(seen[w] & ~coll[w]);
      left_keys = d_left[level].ptr;
      n_left = (int64_t)cnt;
    }
    // blocks { w0, w1, w2, prefix }: prefix = number of set bits before the block, over both levels
    const size_t n_blocks = (size_t)nb[0] + nb[1];
    std::vector<unsigned long long> blocks(n_blocks * 4);
    unsigned long long prefix = 0;
    for (size_t b = 0; b < n_blocks; ++b) {
      blocks[4 * b + 3] = prefix;
      for (int k = 0; k < 3; ++k) {
        blocks[4 * b + k] = bits[3 * b + k];
        prefix += (unsigned long long)__builtin_popcountll(bits[3 * b + k]);
      }
    }
    if ((int64_t)prefix + n_left != n) throw std::runtime_error("k_rows: perfect hash lost states");
    ctx->d_mph_blocks.alloc(blocks.size() * 8);
    CUDA_CHECK(cudaMemcpyAsync(ctx->d_mph_blocks.ptr, blocks.data(), blocks.size() * 8, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    ctx->mph.blocks = ctx->d_mph_blocks.ptr;
    ctx->mph.n_blocks0 = nb[0];
    ctx->mph.n_blocks1 = nb[1];
    ctx->mph.n_dense = (uint32_t)prefix;
    const size_t dense_bytes = (size_t)std::max<unsigned long long>(1, prefix) * (ce ? 32 : 16);
    ctx->d_dense.alloc(dense_bytes);
    CUDA_CHECK(cudaMemsetAsync(ctx->d_dense.ptr, 0xff, dense_bytes, st));
  }
  // ---- open-addressing table over the states that are left (all of them without the dense index)
  // complex128: one-slot buckets, 8 per state (1.07 probes per look-up) while the table stays below a quarter of the
  // free memory, else 4 or 2 per state; float64: two-slot buckets, 2 per state
  size_t free_b = 0, total_b = 0;
  CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
  int64_t per_state = ce ? 8 : 2;
  while (per_state > 2 && (double)per_state * n_left * 32.0 > 0.25 * (double)free_b) per_state /= 2;
  if (per_state * n_left + 16 >= 2147483647ll) throw std::runtime_error("k_rows: table of more than 2^31 buckets");
  const uint32_t slots = (uint32_t)std::max<int64_t>(16, per_state * n_left);
  ctx->d_table.alloc((size_t)slots * 32);
  ctx->d_slot_of.alloc((size_t)std::max<int64_t>(1, n));
  CUDA_CHECK(cudaMemsetAsync(ctx->d_table.ptr, 0xff, (size_t)slots * 32, st));
  if (ctx->dense_index) {
    DevBuf<uint32_t> d_tmp;
    d_tmp.alloc((size_t)std::max<int64_t>(1, n_left));
    launch_table_insert(left_keys, n_left, ctx->d_table.ptr, slots, ce ? 1 : 2, d_tmp.ptr, st);
    launch_mph_slots(ctx->d_reps.ptr, n, ctx->mph, ctx->d_table.ptr, slots, ce ? 1 : 2, ctx->d_slot_of.ptr,
                     ctx->d_status.ptr, st);
    CUDA_CHECK(cudaStreamSynchronize(st));
  } else {
    launch_table_insert(ctx->d_reps.ptr, n, ctx->d_table.ptr, slots, ce ? 1 : 2, ctx->d_slot_of.ptr, st);
  }
  ctx->table_slots = slots;
  ctx->table_elt = elt;
}

// y[rows] <- rows of H through k_rows.  `basis` owns the table (this rank's context, or the twin holding the whole
// basis in the replicated-x product), x_all is indexed like basis' states (through pos when given), p names the rows.
void rows_product(dmv_context *basis, KernelParams &p, int elt, const void *x_all, const uint32_t *pos,
                  cudaStream_t stream, bool fill = true, dmv_context *timer = nullptr) {
  if (!timer) timer = basis;   // whose event timeline the refill belongs to (the rank's context in the replicated form)
  cudaStream_t keep = basis->stream;
  basis->stream = stream;
  ensure_table(basis, elt);
  basis->stream = keep;
  if (fill) {   // (a product cut into row chunks refreshes the values once, with its first chunk)
    CUDA_CHECK(cudaEventRecord(timer->ev_fill[0], stream));
    launch_table_fill(basis->n_states, elt == DMV_C128, x_all, basis->d_norms.ptr, pos, basis->d_slot_of.ptr,
                      basis->d_reps.ptr, basis->d_table.ptr, basis->dense_index ? basis->d_dense.ptr : nullptr, stream);
    CUDA_CHECK(cudaEventRecord(timer->ev_fill[1], stream));
    timer->fill_timed = true;
  }
  select_tables(basis, p, true, false);
  p.uni_re = basis->gather_uni[0]; p.uni_im = basis->gather_uni[1];
  p.table = basis->d_table.ptr;
  p.table_slots = basis->table_slots;
  p.mph = basis->mph;
  p.dense = basis->dense_index ? basis->d_dense.ptr : nullptr;
  p.row_split = 1;
  launch_rows(p, elt == DMV_C128, stream);
}

void do_generate(dmv_context *ctx, int elt, const void *x_dev, void *y_dev,
                 const void *x_host_pending = nullptr, int64_t row_begin = 0, int64_t row_end = 0) {
  if (use_pull(ctx)) {   // one rank owns the basis: traverse by rows (gather), see k_gather / k_pull
    KernelParams p = base_params(ctx);
    p.x = x_dev;
    p.y = y_dev;
    if (row_end > row_begin) { p.row_begin = row_begin; p.row_end = row_end; }
    if (use_gather(ctx)) {
      select_tables(ctx, p, true, ctx->complex_coefficients);
      p.row_split = choose_row_split(ctx->n_states, (int)ctx->h_pull.groups.size());
      p.uni_re = ctx->gather_uni[0]; p.uni_im = ctx->gather_uni[1];
      launch_gather(p, ctx->proj == PROJ_INVERSION, ctx->complex_coefficients, elt == DMV_C128,
                    ctx->gather_narrow, ctx->index_mode == INDEX_LIN, ctx->gather_uniform, ctx->stream);
      return;
    }
    if (use_rows(ctx)) {
      rows_product(ctx, p, elt, x_dev, nullptr, ctx->stream, /*fill=*/row_begin == 0);
      return;
    }
    select_tables(ctx, p, true, complex_values(ctx, elt));
    launch_pull(p, ctx->proj, complex_values(ctx, elt), elt == DMV_C128, ctx->stream);
    return;
  }
  if (!ctx->planned) do_plan(ctx);
  zero_y_if_diag(ctx, elt, y_dev);
  if (ctx->num_ranks > 1)
    CUDA_CHECK(cudaMemsetAsync(ctx->d_out_count.ptr, 0, sizeof(unsigned long long) * ctx->num_ranks, ctx->stream));
  KernelParams p = base_params(ctx);
  p.x = x_dev;
  p.y = y_dev;
  const bool cv = complex_values(ctx, elt);
  if (ctx->num_ranks > 1 && ctx->peer_direct && ctx->ptr_width != (cv ? 2 : 1)) {
    // peer-direct: destination d's records are stored straight into d's incoming buffer over NVLink
    const int width = cv ? 2 : 1;
    for (int d = 0; d < ctx->num_ranks; ++d) {
      if (d == ctx->rank) { ctx->h_out_betas_ptr[d] = nullptr; ctx->h_out_coeffs_ptr[d] = nullptr; continue; }
      ctx->h_out_betas_ptr[d] = reinterpret_cast<uint64_t *>(ctx->peer_betas[d]) + ctx->my_offset_in_peer[d];
      ctx->h_out_coeffs_ptr[d] = reinterpret_cast<double *>(ctx->peer_coeffs[d]) + ctx->my_offset_in_peer[d] * width;
    }
    upload_out_pointers(ctx);
    p.out_betas_ptr = ctx->d_out_betas_ptr.ptr;
    p.out_coeffs_ptr = ctx->d_out_coeffs_ptr.ptr;
    ctx->ptr_width = width;
  }
  if (ctx->num_ranks > 1 && !ctx->peer_direct && ctx->ptr_width != (cv ? 2 : 1)) {
    // local buckets: destination d's records start at out_offset[d] (coefficients: width doubles each)
    const int width = cv ? 2 : 1;
    for (int d = 0; d < ctx->num_ranks; ++d) {
      ctx->h_out_betas_ptr[d] = ctx->d_out_betas.ptr + ctx->h_out_offset[d];
      ctx->h_out_coeffs_ptr[d] = ctx->d_out_coeffs.ptr + ctx->h_out_offset[d] * width;
    }
    upload_out_pointers(ctx);
    p.out_betas_ptr = ctx->d_out_betas_ptr.ptr;
    p.out_coeffs_ptr = ctx->d_out_coeffs_ptr.ptr;
    ctx->ptr_width = width;
  }
  ctx->record_width = cv ? 2 : 1;
  select_tables(ctx, p, false, cv);
  if (!x_host_pending) {
    launch_generate(p, ctx->proj, cv, elt == DMV_C128, false, ctx->stream);
    return;
  }
  // pipelined: chunk k of x is copied while chunk k-1 is being generated
  const int chunks = dmv_context::kCopyChunks;
  const int64_t n = ctx->n_states, per = ((n + chunks - 1) / chunks + 31) / 32 * 32;
  const size_t esz = (size_t)8 * elt;
  CUDA_CHECK(cudaEventRecord(ctx->ev_chunk[0], ctx->stream));       // copy stream starts after prior work
  CUDA_CHECK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_chunk[0], 0));
  for (int k = 0; k < chunks; ++k) {
    const int64_t b = std::min<int64_t>(n, (int64_t)k * per), e = std::min<int64_t>(n, b + per);
    if (e <= b) break;
    CUDA_CHECK(cudaMemcpyAsync(reinterpret_cast<char *>(ctx->d_x.ptr) + b * esz,
                               reinterpret_cast<const char *>(x_host_pending) + b * esz, (size_t)(e - b) * esz,
                               cudaMemcpyHostToDevice, ctx->copy_stream));
    CUDA_CHECK(cudaEventRecord(ctx->ev_chunk[k], ctx->copy_stream));
    CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[k], 0));
    p.row_begin = b;
    p.row_end = e;
    launch_generate(p, ctx->proj, cv, elt == DMV_C128, false, ctx->stream);
  }
}

void do_accumulate(dmv_context *ctx, int elt, int64_t count, const uint64_t *betas, const double *coeffs,
                   void *y_dev) {
  KernelParams p = base_params(ctx);
  p.y = y_dev;
  launch_accumulate(p, ctx->proj, complex_values(ctx, elt), elt == DMV_C128, count, betas, coeffs, ctx->stream);
}

void collect_timings(dmv_context *ctx) {
  auto ms = [&](int a, int b) { float t = 0; cudaEventElapsedTime(&t, ctx->ev[a], ctx->ev[b]); return (double)t; };
  ctx->timings[T_H2D] = ms(0, 1);
  if (ctx->timeline_replicated) {   // the exchange (all-gather of x) comes first, then the row gather
    ctx->timings[T_EXCHANGE] = ms(1, 6);
    ctx->timings[T_GENERATE] = ms(6, 2);
    ctx->timings[T_ACCUMULATE] = 0.0;
  } else {
  ctx->timings[T_GENERATE] = ms(1, 2);
  ctx->timings[T_EXCHANGE] = ms(2, 3);
  ctx->timings[T_ACCUMULATE] = ms(3, 4);
  }
  ctx->timings[T_D2H] = ms(4, 5);
  ctx->timings[T_TOTAL] = ms(0, 5);
  ctx->timings[T_TABLE_FILL] = 0.0;
  if (ctx->fill_timed) {
    float t = 0;
    if (cudaEventElapsedTime(&t, ctx->ev_fill[0], ctx->ev_fill[1]) == cudaSuccess) ctx->timings[T_TABLE_FILL] = t;
    ctx->fill_timed = false;
  }
}

std::mutex g_bind_mutex;
std::map<const void *, dmv_context *> g_bindings;

}  // namespace

// ---- small kernels exposed with host-or-device pointers ----------------------------------------
namespace {
template <typename T>
struct InArg {  // device view of an input array
  DevBuf<T> buf; const T *ptr;
  InArg(const T *p, size_t n, cudaStream_t s) {
    if (is_device_pointer(p)) ptr = p;
    else { buf.alloc(n); if (n) CUDA_CHECK(cudaMemcpyAsync(buf.ptr, p, n * sizeof(T), cudaMemcpyHostToDevice, s)); ptr = buf.ptr; }
  }
};
template <typename T>
struct OutArg {  // device view of an output array, copied back by finish()
  DevBuf<T> buf; T *ptr; T *user; size_t n; bool host;
  OutArg(T *p, size_t n_) : user(p), n(n_) {
    host = !is_device_pointer(p);
    if (host) { buf.alloc(n); ptr = buf.ptr; } else ptr = p;
  }
  void finish(cudaStream_t s, size_t used = (size_t)-1) {
    if (host && user) { const size_t m = used == (size_t)-1 ? n : used; if (m) CUDA_CHECK(cudaMemcpyAsync(user, ptr, m * sizeof(T), cudaMemcpyDeviceToHost, s)); }
  }
};
}  // namespace



// One-time exchange of the plan: every rank learns how many records each peer sends it; then, when
// possible, the peers' incoming buffers are mapped (CUDA IPC over NVLink) so that k_generate can store
// remote records directly where the owner will read them.
void setup_exchange(dmv_context *ctx) {
  NcclApi &N = nccl();
  const int P = ctx->num_ranks;
  for (auto &q : ctx->peer_betas) if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
  for (auto &q : ctx->peer_coeffs) if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
  DevBuf<int64_t> d_send, d_all;
  d_send.upload(ctx->send_counts, ctx->stream);
  d_all.alloc((size_t)P * P);
  NCCL_CHECK(N.AllGather(d_send.ptr, d_all.ptr, (size_t)P, ncclInt64, ctx->comm, ctx->stream));
  std::vector<int64_t> all((size_t)P * P);   // all[r * P + q]: records r emits for q (own ones included)
  CUDA_CHECK(cudaMemcpyAsync(all.data(), d_all.ptr, all.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  int64_t total_in = 0;
  for (int q = 0; q < P; ++q) {
    ctx->recv_counts[q] = (q == ctx->rank) ? 0 : all[(size_t)q * P + ctx->rank];
    total_in += ctx->recv_counts[q];
  }
  ctx->d_in_betas.alloc((size_t)total_in);
  ctx->d_in_coeffs.alloc((size_t)total_in * 2);
  ctx->d_barrier.alloc(1);
  CUDA_CHECK(cudaMemsetAsync(ctx->d_barrier.ptr, 0, sizeof(int), ctx->stream));
  ctx->peer_direct = false;
  if (ctx->opt_exchange == 0 || P > 32) return;

  // ---- try to map the peers' incoming buffers
  struct Handles { cudaIpcMemHandle_t betas, coeffs; int ok; int pad[15]; };
  static_assert(sizeof(Handles) % 8 == 0, "handle block");
  Handles mine{};
  mine.ok = (cudaIpcGetMemHandle(&mine.betas, ctx->d_in_betas.ptr) == cudaSuccess &&
             cudaIpcGetMemHandle(&mine.coeffs, ctx->d_in_coeffs.ptr) == cudaSuccess) ? 1 : 0;
  cudaGetLastError();
  DevBuf<char> d_mine, d_handles;
  d_mine.alloc(sizeof(Handles));
  d_handles.alloc(sizeof(Handles) * P);
  CUDA_CHECK(cudaMemcpyAsync(d_mine.ptr, &mine, sizeof(Handles), cudaMemcpyHostToDevice, ctx->stream));
  NCCL_CHECK(N.AllGather(d_mine.ptr, d_handles.ptr, sizeof(Handles), ncclChar, ctx->comm, ctx->stream));
  std::vector<Handles> handles(P);
  CUDA_CHECK(cudaMemcpyAsync(handles.data(), d_handles.ptr, sizeof(Handles) * P, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  int ok = 1;
  for (int q = 0; q < P; ++q) ok &= handles[q].ok;
  ctx->peer_betas.assign(P, nullptr);
  ctx->peer_coeffs.assign(P, nullptr);
  if (ok) {
    for (int q = 0; q < P && ok; ++q) {
      if (q == ctx->rank) continue;
      if (cudaIpcOpenMemHandle(&ctx->peer_betas[q], handles[q].betas, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
          cudaIpcOpenMemHandle(&ctx->peer_coeffs[q], handles[q].coeffs, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        ok = 0;
        cudaGetLastError();
      }
    }
  }
  // everybody must agree
  int agree = ok;
  CUDA_CHECK(cudaMemcpyAsync(ctx->d_barrier.ptr, &agree, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  NCCL_CHECK(N.AllReduce(ctx->d_barrier.ptr, ctx->d_barrier.ptr, 1, ncclInt32, ncclMin, ctx->comm, ctx->stream));
  CUDA_CHECK(cudaMemcpyAsync(&agree, ctx->d_barrier.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  if (!agree) {
    for (auto &q : ctx->peer_betas) if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
    for (auto &q : ctx->peer_coeffs) if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
    if (ctx->opt_exchange == 1) throw std::runtime_error("peer-direct exchange requested but CUDA IPC mapping failed");
    return;
  }
  // my region inside peer q's incoming buffer: after the regions of the ranks before me (q itself sends nothing)
  ctx->my_offset_in_peer.assign(P, 0);
  for (int q = 0; q < P; ++q) {
    int64_t off = 0;
    for (int r = 0; r < ctx->rank; ++r)
      if (r != q) off += all[(size_t)r * P + q];
    ctx->my_offset_in_peer[q] = off;
  }
  ctx->peer_direct = true;
  ctx->ptr_width = 0;
}

// -------------------------------------------------------------------------------------------------
// Replicated-x product.  With 180 GB of HBM per GPU every basis of BASELINE.json fits on ONE device many times
// over, so for operators k_gather applies to, the ranks can trade the reference's record exchange (24 bytes per
// off-diagonal term over NVLink, DMV:313-436) for one all-gather of x (E bytes per STATE): every rank keeps the
// whole sorted basis (a single-rank twin context), gathers x from all ranks into slots of equal size, and computes
// ITS rows by the atomics-free row traversal.  The hash partition of x, y and the representatives -- the layout the
// callers see (SE:129-156) -- is unchanged.  Local part of the set-up; no communication here.
void setup_replicated(dmv_context *ctx) {
  require_states(ctx);
  const int P = ctx->num_ranks;
  if (P > 32) throw std::runtime_error("replicated-x product supports at most 32 ranks");
  if (!ctx->global) {
    dmv_basis_desc b{};
    b.number_sites = ctx->n_sites; b.hamming_weight = ctx->hamming_weight; b.spin_inversion = ctx->spin_inversion;
    if (ctx->proj == PROJ_GROUP) {
      b.has_permutations = 1; b.group_order = ctx->k_group_order;
      b.perms = ctx->k_perms.data(); b.flips = ctx->k_flips.data(); b.characters = ctx->k_chars.data();
    }
    dmv_operator_desc o{};
    o.n_off = (int64_t)ctx->k_off_m.size(); o.off_v = ctx->k_off_v.data();
    o.off_m = ctx->k_off_m.data(); o.off_r = ctx->k_off_r.data(); o.off_x = ctx->k_off_x.data(); o.off_s = ctx->k_off_s.data();
    o.n_diag = (int64_t)ctx->k_diag_m.size(); o.diag_v = ctx->k_diag_v.data();
    o.diag_m = ctx->k_diag_m.data(); o.diag_r = ctx->k_diag_r.data(); o.diag_s = ctx->k_diag_s.data();
    // rough size check before enumerating: reps + directory + positions + gathered x
    double states = 1.0;
    if (ctx->hamming_weight >= 0) states = (double)binom().c[ctx->n_sites][ctx->hamming_weight];
    else states = std::ldexp(1.0, ctx->n_sites);
    if (ctx->spin_inversion != 0 && ctx->proj != PROJ_GROUP) states *= 0.5;
    if (ctx->proj == PROJ_GROUP) states = 1.5 * states / (double)std::max<int64_t>(1, ctx->k_group_order) + 1e4;
    size_t free_b = 0, total_b = 0;
    CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
    if (states * 48.0 > 0.5 * (double)free_b) throw std::runtime_error("replicated-x product: the whole basis does not fit");
    dmv_context *g = nullptr;
    if (dmv_context_create(&b, &o, ctx->device, 0, 1, &g) != 0) throw std::runtime_error(g_last_error);
    ctx->global = g;
    g->opt_rows = ctx->opt_rows;
    g->opt_gather_walk = ctx->opt_gather_walk;
    g->opt_rows_index = ctx->opt_rows_index;
    g->opt_rows_ctas = ctx->opt_rows_ctas;
    if (ctx->opt_canon != g->opt_canon && g->proj == PROJ_GROUP) { g->opt_canon = ctx->opt_canon; upload_orbit(g); }
    if (dmv_basis_build(g) != 0) throw std::runtime_error(g_last_error);
  }
  dmv_context *g = ctx->global;
  CUDA_CHECK(cudaStreamSynchronize(g->stream));
  const int64_t n = g->n_states;
  // ---- slot of every global state: owner r = hash % P (SE:129-136), index inside r's ascending block
  const int64_t chunk = 256, n_chunks = (n + chunk - 1) / chunk;
  DevBuf<unsigned long long> d_counts, d_base;
  d_counts.alloc((size_t)n_chunks * P);
  d_base.alloc((size_t)n_chunks * P);
  ctx->d_pos.alloc((size_t)n);
  launch_owner_positions(g->d_reps.ptr, nullptr, n, P, chunk, false, d_counts.ptr, nullptr, 0, nullptr, ctx->stream);
  std::vector<unsigned long long> counts((size_t)n_chunks * P), base((size_t)n_chunks * P);
  CUDA_CHECK(cudaMemcpyAsync(counts.data(), d_counts.ptr, counts.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  std::vector<unsigned long long> total(P, 0);
  for (int64_t c = 0; c < n_chunks; ++c)
    for (int r = 0; r < P; ++r) { base[(size_t)c * P + r] = total[r]; total[r] += counts[(size_t)c * P + r]; }
  if ((int64_t)total[ctx->rank] != ctx->n_states)
    throw std::runtime_error("replicated-x product: this rank's block is not the hash partition of the full basis");
  int64_t block = 0;
  for (int r = 0; r < P; ++r) block = std::max<int64_t>(block, (int64_t)total[r]);
  block = (block + 1) / 2 * 2;
  if ((double)block * P >= 4294967295.0) throw std::runtime_error("replicated-x product: more than 2^32 slots");
  d_base.upload(base, ctx->stream);
  launch_owner_positions(g->d_reps.ptr, nullptr, n, P, chunk, true, d_counts.ptr, d_base.ptr, block, ctx->d_pos.ptr, ctx->stream);
  ctx->repl_block = block;
  ctx->d_xcat.alloc((size_t)block * P * 2 * 2);   // two buffers of P slots (alternating products), 16 bytes per element
  CUDA_CHECK(cudaMemsetAsync(ctx->d_xcat.ptr, 0, (size_t)block * P * 2 * 2 * sizeof(double), ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

// y (this rank's block) <- rows of H applied to the gathered x (slot r * repl_block holds rank r's block)
void replicated_rows(dmv_context *ctx, int elt, const void *x_cat, void *y_dev) {
  dmv_context *g = ctx->global;
  KernelParams p = base_params(g);
  p.x = x_cat;
  p.y = y_dev;
  p.status = ctx->d_status.ptr;
  p.row_states = ctx->d_reps.ptr;
  p.row_begin = 0;
  p.row_end = ctx->n_states;
  p.pos = ctx->d_pos.ptr;
  p.x_row_offset = (int64_t)ctx->rank * ctx->repl_block;
  if (use_gather(g)) {
    select_tables(g, p, true, g->complex_coefficients);
    p.row_split = choose_row_split(ctx->n_states, (int)g->h_pull.groups.size());
    p.uni_re = g->gather_uni[0]; p.uni_im = g->gather_uni[1];
    launch_gather(p, g->proj == PROJ_INVERSION, g->complex_coefficients, elt == DMV_C128, g->gather_narrow,
                  g->index_mode == INDEX_LIN, g->gather_uniform, ctx->stream);
    return;
  }
  p.row_norms = ctx->d_norms.ptr;
  if (use_rows(g)) {   // bases with permutation symmetries: hash table over the whole basis, filled from the gathered x
    rows_product(g, p, elt, x_cat, ctx->d_pos.ptr, ctx->stream, true, ctx);
    return;
  }
  // operators outside the bit-parallel test / non-trivial characters: the queued row traversal
  if (p.index.mode == INDEX_RANK) p.index.mode = INDEX_DIRECTORY;   // the incremental rank needs row index == rank
  p.row_split = 1;
  select_tables(g, p, true, complex_values(g, elt));
  launch_pull(p, g->proj, complex_values(g, elt), elt == DMV_C128, ctx->stream);
}

// Collective set-up of the overlapped record exchange: per-round counting passes, exchange of the counts, incoming
// buffers laid out round-major, CUDA IPC mapping of buffers and flags.  Leaves rounds.ready false when it does not apply
// (one round, IPC impossible): the caller then uses the one-shot exchange.
void setup_rounds(dmv_context *ctx) {
  dmv_context::Rounds &Q = ctx->rounds;
  NcclApi &N = nccl();
  const int P = ctx->num_ranks;
  Q.tried = true;
  Q.ready = false;
  int R = ctx->opt_rounds;
  if (R < 0) R = ctx->n_states >= (1 << 18) ? 4 : 1;
  // every rank must use the same number of rounds
  ctx->d_barrier.alloc(1);
  CUDA_CHECK(cudaMemcpyAsync(ctx->d_barrier.ptr, &R, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  NCCL_CHECK(N.AllReduce(ctx->d_barrier.ptr, ctx->d_barrier.ptr, 1, ncclInt32, ncclMin, ctx->comm, ctx->stream));
  CUDA_CHECK(cudaMemcpyAsync(&R, ctx->d_barrier.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  if (R <= 1 || P > 32 || ctx->opt_exchange == 0) return;
  Q.R = R;
  Q.row_split = 1;
  Q.row_begin.assign(R + 1, 0);
  for (int r = 0; r <= R; ++r) Q.row_begin[r] = std::min<int64_t>(ctx->n_states, (ctx->n_states * r / R + 31) / 32 * 32);
  Q.row_begin[R] = ctx->n_states;
  Q.grid = planned_grid((ctx->n_states + R - 1) / R, 1);
  const size_t n_warps = (size_t)Q.grid * kWarpsPerCta;
  // ---- counting pass per round: exact share of every warp for every destination
  std::vector<int64_t> offsets((size_t)R * n_warps * P, 0), counts((size_t)R * P, 0);
  ctx->d_warp_counts.alloc(n_warps * P);
  ctx->d_out_count.alloc(P);
  std::vector<unsigned long long> wc(n_warps * P);
  Q.terms = 0;
  for (int r = 0; r < R; ++r) {
    CUDA_CHECK(cudaMemsetAsync(ctx->d_warp_counts.ptr, 0, sizeof(unsigned long long) * n_warps * P, ctx->stream));
    KernelParams p = base_params(ctx);
    p.grid_blocks = Q.grid;
    p.row_split = 1;
    p.row_begin = Q.row_begin[r];
    p.row_end = Q.row_begin[r + 1];
    p.warp_counts = ctx->d_warp_counts.ptr;
    select_tables(ctx, p, false, ctx->complex_coefficients);
    launch_generate(p, ctx->proj, ctx->complex_coefficients, false, /*count_only=*/true, ctx->stream);
    CUDA_CHECK(cudaMemcpyAsync(wc.data(), ctx->d_warp_counts.ptr, sizeof(unsigned long long) * wc.size(),
                               cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    for (int d = 0; d < P; ++d)
      for (size_t w = 0; w < n_warps; ++w) {
        offsets[((size_t)r * n_warps + w) * P + d] = counts[(size_t)r * P + d];
        counts[(size_t)r * P + d] += (int64_t)wc[w * P + d];
      }
    for (int d = 0; d < P; ++d) Q.terms += counts[(size_t)r * P + d];
  }
  Q.d_warp_offsets.upload(offsets, ctx->stream);
  if (!ctx->planned) ctx->number_terms = Q.terms;
  std::vector<int64_t> capacity((size_t)R * P);
  for (int r = 0; r < R; ++r)
    for (int d = 0; d < P; ++d) capacity[(size_t)r * P + d] = d == ctx->rank ? 0 : counts[(size_t)r * P + d];
  Q.d_capacity.upload(capacity, ctx->stream);
  // ---- everybody's counts: all[s][r][d]
  DevBuf<int64_t> d_send, d_all;
  d_send.upload(counts, ctx->stream);
  d_all.alloc((size_t)P * R * P);
  NCCL_CHECK(N.AllGather(d_send.ptr, d_all.ptr, (size_t)R * P, ncclInt64, ctx->comm, ctx->stream));
  std::vector<int64_t> all((size_t)P * R * P);
  CUDA_CHECK(cudaMemcpyAsync(all.data(), d_all.ptr, all.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  auto cnt = [&](int s, int r, int d) { return all[((size_t)s * R + r) * P + d]; };
  // incoming buffer of rank q, round-major: [round 0: sources 0 .. P-1 (without q)] [round 1: ...] ...
  auto region = [&](int q, int r, int src) {   // first record of (round r, source src) inside q's buffer
    int64_t off = 0;
    for (int rr = 0; rr < r; ++rr)
      for (int s = 0; s < P; ++s) if (s != q) off += cnt(s, rr, q);
    for (int s = 0; s < src; ++s) if (s != q) off += cnt(s, r, q);
    return off;
  };
  Q.in_slice.assign(R + 1, 0);
  for (int r = 0; r <= R; ++r) Q.in_slice[r] = region(ctx->rank, r, 0);
  Q.in_total = Q.in_slice[R];
  Q.peer_total.assign(P, 0);
  for (int q = 0; q < P; ++q) Q.peer_total[q] = region(q, R, 0);
  Q.my_off.assign((size_t)R * P, 0);
  for (int r = 0; r < R; ++r)
    for (int q = 0; q < P; ++q) if (q != ctx->rank) Q.my_off[(size_t)r * P + q] = region(q, r, ctx->rank);
  Q.d_in_betas.alloc((size_t)std::max<int64_t>(1, 2 * Q.in_total));
  Q.d_in_coeffs.alloc((size_t)std::max<int64_t>(1, 4 * Q.in_total));
  Q.d_flags.alloc(P);
  CUDA_CHECK(cudaMemsetAsync(Q.d_flags.ptr, 0, sizeof(unsigned) * P, ctx->stream));
  Q.seq = 0;
  // ---- map the peers' buffers and flags
  for (auto *v : {&Q.peer_betas, &Q.peer_coeffs, &Q.peer_flags})
    for (auto &q : *v) if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
  struct Handles { cudaIpcMemHandle_t betas, coeffs, flags; int ok; int pad[15]; };
  static_assert(sizeof(Handles) % 8 == 0, "handle block");
  Handles mine{};
  mine.ok = (cudaIpcGetMemHandle(&mine.betas, Q.d_in_betas.ptr) == cudaSuccess &&
             cudaIpcGetMemHandle(&mine.coeffs, Q.d_in_coeffs.ptr) == cudaSuccess &&
             cudaIpcGetMemHandle(&mine.flags, Q.d_flags.ptr) == cudaSuccess) ? 1 : 0;
  cudaGetLastError();
  DevBuf<char> d_mine, d_handles;
  d_mine.alloc(sizeof(Handles));
  d_handles.alloc(sizeof(Handles) * P);
  CUDA_CHECK(cudaMemcpyAsync(d_mine.ptr, &mine, sizeof(Handles), cudaMemcpyHostToDevice, ctx->stream));
  NCCL_CHECK(N.AllGather(d_mine.ptr, d_handles.ptr, sizeof(Handles), ncclChar, ctx->comm, ctx->stream));
  std::vector<Handles> handles(P);
  CUDA_CHECK(cudaMemcpyAsync(handles.data(), d_handles.ptr, sizeof(Handles) * P, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  int ok = 1;
  for (int q = 0; q < P; ++q) ok &= handles[q].ok;
  Q.peer_betas.assign(P, nullptr); Q.peer_coeffs.assign(P, nullptr); Q.peer_flags.assign(P, nullptr);
  for (int q = 0; q < P && ok; ++q) {
    if (q == ctx->rank) continue;
    if (cudaIpcOpenMemHandle(&Q.peer_betas[q], handles[q].betas, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
        cudaIpcOpenMemHandle(&Q.peer_coeffs[q], handles[q].coeffs, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
        cudaIpcOpenMemHandle(&Q.peer_flags[q], handles[q].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      ok = 0;
      cudaGetLastError();
    }
  }
  int agree = ok;
  CUDA_CHECK(cudaMemcpyAsync(ctx->d_barrier.ptr, &agree, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  NCCL_CHECK(N.AllReduce(ctx->d_barrier.ptr, ctx->d_barrier.ptr, 1, ncclInt32, ncclMin, ctx->comm, ctx->stream));
  CUDA_CHECK(cudaMemcpyAsync(&agree, ctx->d_barrier.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  if (!agree) {
    for (auto *v : {&Q.peer_betas, &Q.peer_coeffs, &Q.peer_flags})
      for (auto &q : *v) if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
    return;
  }
  std::vector<unsigned *> flags(P);
  for (int q = 0; q < P; ++q) flags[q] = q == ctx->rank ? Q.d_flags.ptr : reinterpret_cast<unsigned *>(Q.peer_flags[q]);
  Q.d_peer_flags.upload(flags, ctx->stream);
  if (!Q.acc_stream) CUDA_CHECK(cudaStreamCreateWithFlags(&Q.acc_stream, cudaStreamNonBlocking));
  if (!Q.ev_begin) CUDA_CHECK(cudaEventCreateWithFlags(&Q.ev_begin, cudaEventDisableTiming));
  if (!Q.ev_done) CUDA_CHECK(cudaEventCreateWithFlags(&Q.ev_done, cudaEventDisableTiming));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  Q.ptr_width = 0;
  Q.ready = true;
}

// where my records of (buffer, round, destination) go: [2][R][P] pointers into the peers' incoming buffers
void upload_round_pointers(dmv_context *ctx, int width) {
  dmv_context::Rounds &Q = ctx->rounds;
  const int P = ctx->num_ranks, R = Q.R;
  std::vector<uint64_t *> bp((size_t)2 * R * P, nullptr);
  std::vector<double *> cp((size_t)2 * R * P, nullptr);
  for (int b = 0; b < 2; ++b)
    for (int r = 0; r < R; ++r)
      for (int q = 0; q < P; ++q) {
        if (q == ctx->rank) continue;
        const int64_t first = (int64_t)b * Q.peer_total[q] + Q.my_off[(size_t)r * P + q];
        bp[((size_t)b * R + r) * P + q] = reinterpret_cast<uint64_t *>(Q.peer_betas[q]) + first;
        cp[((size_t)b * R + r) * P + q] = reinterpret_cast<double *>(Q.peer_coeffs[q]) + first * width;
      }
  Q.d_bptr.upload(bp, ctx->stream);
  Q.d_cptr.upload(cp, ctx->stream);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  Q.ptr_width = width;
}

// One product through the overlapped rounds (x, y device pointers).  Main stream: generate round 0, raise flags,
// generate round 1, ...; second stream: wait for every sender's flag of round r, accumulate its slice.  Returns with the
// main stream waiting for the last accumulate.
void rounds_product(dmv_context *ctx, int elt, const void *x_dev, void *y_dev) {
  dmv_context::Rounds &Q = ctx->rounds;
  const int P = ctx->num_ranks, R = Q.R;
  const bool cv = complex_values(ctx, elt);
  const int width = cv ? 2 : 1;
  if (Q.ptr_width != width) upload_round_pointers(ctx, width);
  ctx->record_width = width;
  zero_y_if_diag(ctx, elt, y_dev);
  CUDA_CHECK(cudaEventRecord(Q.ev_begin, ctx->stream));
  CUDA_CHECK(cudaStreamWaitEvent(Q.acc_stream, Q.ev_begin, 0));
  const int b = (int)(Q.seq & 1u);
  const size_t n_warps = (size_t)Q.grid * kWarpsPerCta;
  for (int r = 0; r < R; ++r) {
    KernelParams p = base_params(ctx);
    p.x = x_dev;
    p.y = y_dev;
    p.grid_blocks = Q.grid;
    p.row_split = 1;
    p.row_begin = Q.row_begin[r];
    p.row_end = Q.row_begin[r + 1];
    p.warp_offsets = Q.d_warp_offsets.ptr + (size_t)r * n_warps * P;
    p.out_capacity = Q.d_capacity.ptr + (size_t)r * P;
    p.out_betas_ptr = Q.d_bptr.ptr + ((size_t)b * R + r) * P;
    p.out_coeffs_ptr = Q.d_cptr.ptr + ((size_t)b * R + r) * P;
    select_tables(ctx, p, false, cv);
    launch_generate(p, ctx->proj, cv, elt == DMV_C128, false, ctx->stream);
    const unsigned value = Q.seq * (unsigned)R + (unsigned)r + 1u;
    launch_raise_flags(Q.d_peer_flags.ptr, P, ctx->rank, value, ctx->stream);
    // owner side, second stream: every sender has delivered round r -> search + accumulate its slice
    launch_wait_flags(Q.d_flags.ptr, P, value, ctx->d_status.ptr, Q.acc_stream);
    const int64_t first = (int64_t)b * Q.in_total + Q.in_slice[r], count = Q.in_slice[r + 1] - Q.in_slice[r];
    if (count > 0) {
      KernelParams pa = base_params(ctx);
      pa.y = y_dev;
      launch_accumulate(pa, ctx->proj, cv, elt == DMV_C128, count, Q.d_in_betas.ptr + first,
                        Q.d_in_coeffs.ptr + first * width, Q.acc_stream);
    }
  }
  ++Q.seq;
  CUDA_CHECK(cudaEventRecord(ctx->ev[2], ctx->stream));   // end of generation
  CUDA_CHECK(cudaEventRecord(ctx->ev[3], ctx->stream));
  CUDA_CHECK(cudaEventRecord(Q.ev_done, Q.acc_stream));
  CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, Q.ev_done, 0));   // what is left of the accumulate is the exposed part
}

// Collective: map every rank's gathered-x buffers and flag words into every other rank (CUDA IPC over NVLink) so that
// the all-gather of x becomes one kernel of peer stores + flags (launch_push_block).  Falls back to the NCCL all-gather
// when any rank cannot map.
void setup_peer_gather(dmv_context *ctx) {
  NcclApi &N = nccl();
  const int P = ctx->num_ranks;
  ctx->peer_gather = false;
  ctx->d_flags.alloc(P);
  ctx->d_push_done.alloc(1);
  CUDA_CHECK(cudaMemsetAsync(ctx->d_flags.ptr, 0, sizeof(unsigned) * P, ctx->stream));
  CUDA_CHECK(cudaMemsetAsync(ctx->d_push_done.ptr, 0, sizeof(unsigned), ctx->stream));
  ctx->gather_epoch = 0;
  struct Handles { cudaIpcMemHandle_t xcat, flags; int ok; int pad[15]; };
  static_assert(sizeof(Handles) % 8 == 0, "handle block");
  Handles mine{};
  mine.ok = (ctx->opt_peer_gather != 0 && cudaIpcGetMemHandle(&mine.xcat, ctx->d_xcat.ptr) == cudaSuccess &&
             cudaIpcGetMemHandle(&mine.flags, ctx->d_flags.ptr) == cudaSuccess) ? 1 : 0;
  cudaGetLastError();
  DevBuf<char> d_mine, d_handles;
  d_mine.alloc(sizeof(Handles));
  d_handles.alloc(sizeof(Handles) * P);
  CUDA_CHECK(cudaMemcpyAsync(d_mine.ptr, &mine, sizeof(Handles), cudaMemcpyHostToDevice, ctx->stream));
  NCCL_CHECK(N.AllGather(d_mine.ptr, d_handles.ptr, sizeof(Handles), ncclChar, ctx->comm, ctx->stream));
  std::vector<Handles> handles(P);
  CUDA_CHECK(cudaMemcpyAsync(handles.data(), d_handles.ptr, sizeof(Handles) * P, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  int ok = 1;
  for (int q = 0; q < P; ++q) ok &= handles[q].ok;
  for (auto &q : ctx->peer_xcat) if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
  for (auto &q : ctx->peer_flagmem) if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
  ctx->peer_xcat.assign(P, nullptr);
  ctx->peer_flagmem.assign(P, nullptr);
  if (ok) {
    for (int q = 0; q < P && ok; ++q) {
      if (q == ctx->rank) continue;
      if (cudaIpcOpenMemHandle(&ctx->peer_xcat[q], handles[q].xcat, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
          cudaIpcOpenMemHandle(&ctx->peer_flagmem[q], handles[q].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        ok = 0;
        cudaGetLastError();
      }
    }
  }
  int agree = ok;   // everybody must agree; the all-reduce is also the barrier after which flags may be raised
  ctx->d_barrier.alloc(1);
  CUDA_CHECK(cudaMemcpyAsync(ctx->d_barrier.ptr, &agree, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  NCCL_CHECK(N.AllReduce(ctx->d_barrier.ptr, ctx->d_barrier.ptr, 1, ncclInt32, ncclMin, ctx->comm, ctx->stream));
  CUDA_CHECK(cudaMemcpyAsync(&agree, ctx->d_barrier.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  if (!agree) {
    for (auto &q : ctx->peer_xcat) if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
    for (auto &q : ctx->peer_flagmem) if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
    return;
  }
  std::vector<unsigned *> flags(P);
  for (int q = 0; q < P; ++q)
    flags[q] = q == ctx->rank ? ctx->d_flags.ptr : reinterpret_cast<unsigned *>(ctx->peer_flagmem[q]);
  ctx->d_peer_flags.upload(flags, ctx->stream);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  ctx->peer_slot_elt = 0;
  ctx->peer_gather = true;
}

// slot `rank` of buffer b of every rank's gathered vector, for elements of `elt` doubles
void upload_peer_slots(dmv_context *ctx, int elt) {
  const int P = ctx->num_ranks;
  const size_t buffer_doubles = (size_t)ctx->repl_block * P * 2;   // buffers are sized for 16-byte elements
  for (int b = 0; b < 2; ++b) {
    std::vector<void *> slots(P);
    for (int q = 0; q < P; ++q) {
      double *base = q == ctx->rank ? ctx->d_xcat.ptr : reinterpret_cast<double *>(ctx->peer_xcat[q]);
      slots[q] = base + b * buffer_doubles + (size_t)ctx->rank * ctx->repl_block * elt;
    }
    ctx->d_peer_slot[b].upload(slots, ctx->stream);
  }
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  ctx->peer_slot_elt = elt;
}

// Collective: which exchange the distributed product uses.  exchange = -1 (auto) prefers the replicated-x product
// when k_gather applies and the whole basis fits, else the record exchange (peer-direct / NCCL, see setup_exchange).
void decide_exchange(dmv_context *ctx) {
  NcclApi &N = nccl();
  int ok = 0;
  std::string why;
  const bool want = (ctx->opt_exchange == 2 || ctx->opt_exchange == -1) && ctx->opt_mode != 0;
  if (want && ctx->num_ranks <= 32) {
    try { setup_replicated(ctx); ok = 1; } catch (const std::exception &e) { why = e.what(); ok = 0; }
  } else {
    why = "switched off (exchange / mode options) or more than 32 ranks";
  }
  ctx->d_barrier.alloc(1);
  CUDA_CHECK(cudaMemcpyAsync(ctx->d_barrier.ptr, &ok, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  NCCL_CHECK(N.AllReduce(ctx->d_barrier.ptr, ctx->d_barrier.ptr, 1, ncclInt32, ncclMin, ctx->comm, ctx->stream));
  int agree = 0;
  CUDA_CHECK(cudaMemcpyAsync(&agree, ctx->d_barrier.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  ctx->replicated = agree != 0;
  ctx->exchange_decided = true;
  if (ctx->replicated) setup_peer_gather(ctx);
  if (!ctx->replicated) {
    delete ctx->global; ctx->global = nullptr;
    ctx->d_pos.release(); ctx->d_xcat.release();
    if (ctx->opt_exchange == 2)
      throw std::runtime_error("replicated-x exchange requested but not possible on every rank: " + why);
  }
}

// -------------------------------------------------------------------------------------------------
// Block <-> hashed redistribution of vectors (arrFromBlockToHashed, reference src/BlockToHashed.chpl:87-208;
// arrFromHashedToBlock, src/HashedToBlock.chpl:67-153).  "Block" = the global array in sorted-state order cut into
// contiguous chunks, one per rank; "hashed" = every rank holds the elements of the states it owns, ascending.
// positions: slot of element i of a chunk in the ordering "grouped by owner, stable": offsets[mask[i]] + #{j < i :
// mask[j] == mask[i]}; counts[r] = elements owned by r.  One counting pass, host prefix sums, one writing pass.
void hashed_positions(dmv_context *ctx, int64_t count, const uint8_t *d_masks, int P, std::vector<int64_t> &counts,
                      uint32_t *d_pos) {
  if (P > 32) throw std::runtime_error("block <-> hashed redistribution supports at most 32 ranks");
  counts.assign(P, 0);
  if (count <= 0) return;
  if (count >= (1ll << 32)) throw std::runtime_error("chunks of more than 2^32 elements are not supported");
  const int64_t chunk = 256, n_chunks = (count + chunk - 1) / chunk;
  DevBuf<unsigned long long> d_counts, d_base;
  d_counts.alloc((size_t)n_chunks * P);
  launch_owner_positions(nullptr, d_masks, count, P, chunk, false, d_counts.ptr, nullptr, 0, nullptr, ctx->stream);
  std::vector<unsigned long long> c((size_t)n_chunks * P), base((size_t)n_chunks * P);
  CUDA_CHECK(cudaMemcpyAsync(c.data(), d_counts.ptr, c.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  for (int64_t k = 0; k < n_chunks; ++k)
    for (int r = 0; r < P; ++r) counts[r] += (int64_t)c[(size_t)k * P + r];
  std::vector<unsigned long long> run(P, 0);
  unsigned long long off = 0;
  for (int r = 0; r < P; ++r) { run[r] = off; off += (unsigned long long)counts[r]; }
  for (int64_t k = 0; k < n_chunks; ++k)
    for (int r = 0; r < P; ++r) { base[(size_t)k * P + r] = run[r]; run[r] += c[(size_t)k * P + r]; }
  d_base.upload(base, ctx->stream);
  launch_owner_positions(nullptr, d_masks, count, P, chunk, true, d_counts.ptr, d_base.ptr, 0, d_pos, ctx->stream);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));   // d_base is released on return
}

// all[r * P + q] = what rank r reported for q (collective)
std::vector<int64_t> all_gather_counts(dmv_context *ctx, const std::vector<int64_t> &mine) {
  NcclApi &N = nccl();
  const int P = ctx->num_ranks;
  DevBuf<int64_t> d_mine, d_all;
  d_mine.upload(mine, ctx->stream);
  d_all.alloc((size_t)P * P);
  NCCL_CHECK(N.AllGather(d_mine.ptr, d_all.ptr, (size_t)P, ncclInt64, ctx->comm, ctx->stream));
  std::vector<int64_t> all((size_t)P * P);
  CUDA_CHECK(cudaMemcpyAsync(all.data(), d_all.ptr, all.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return all;
}

// -------------------------------------------------------------------------------------------------
// Lowest eigenpair of a symmetric tridiagonal matrix (diagonal a[0..k), off-diagonal b[0..k-1)): Sturm bisection for
// the eigenvalue, inverse iteration for the vector.  Host side of dmv_lanczos; k is at most a few hundred.
double tridiagonal_lowest(const std::vector<double> &a, const std::vector<double> &b, std::vector<double> &vec) {
  const int k = (int)a.size();
  double lo = a[0], hi = a[0];
  for (int i = 0; i < k; ++i) {
    const double r = (i > 0 ? std::fabs(b[i - 1]) : 0.0) + (i + 1 < k ? std::fabs(b[i]) : 0.0);
    lo = std::min(lo, a[i] - r);
    hi = std::max(hi, a[i] + r);
  }
  auto below = [&](double x) {   // number of eigenvalues < x
    int count = 0;
    double q = a[0] - x;
    for (int i = 0;; ++i) {
      if (q < 0.0) ++count;
      if (i + 1 == k) break;
      if (std::fabs(q) < 1e-300) q = q < 0 ? -1e-300 : 1e-300;
      q = a[i + 1] - x - b[i] * b[i] / q;
    }
    return count;
  };
  for (int it = 0; it < 200 && hi - lo > 4e-16 * std::max(1.0, std::max(std::fabs(lo), std::fabs(hi))); ++it) {
    const double mid = 0.5 * (lo + hi);
    if (below(mid) >= 1) hi = mid; else lo = mid;
  }
  const double theta = 0.5 * (lo + hi);
  // inverse iteration on (T - shift I): LU of a tridiagonal matrix with partial pivoting (the dgttrf / dgttrs scheme)
  vec.assign(k, 1.0 / std::sqrt((double)k));
  const double scale = std::max(1.0, std::max(std::fabs(lo), std::fabs(hi)));
  const double shift = theta - 1e-13 * scale;
  if (k > 1) {
    std::vector<double> dl(k - 1), d(k), du(k - 1), du2(k > 2 ? k - 2 : 0, 0.0);
    std::vector<int> piv(k - 1);
    for (int i = 0; i < k; ++i) d[i] = a[i] - shift;
    for (int i = 0; i + 1 < k; ++i) { dl[i] = b[i]; du[i] = b[i]; }
    const double tiny = 1e-300;
    for (int i = 0; i + 1 < k; ++i) {
      if (std::fabs(d[i]) >= std::fabs(dl[i])) {
        if (std::fabs(d[i]) < tiny) d[i] = tiny;
        const double f = dl[i] / d[i];
        dl[i] = f;
        d[i + 1] -= f * du[i];
        piv[i] = i;
      } else {
        const double f = d[i] / dl[i];
        d[i] = dl[i];
        dl[i] = f;
        const double t = du[i];
        du[i] = d[i + 1];
        d[i + 1] = t - f * d[i + 1];
        if (i + 2 < k) { du2[i] = du[i + 1]; du[i + 1] = -f * du[i + 1]; }
        piv[i] = i + 1;
      }
    }
    if (std::fabs(d[k - 1]) < tiny) d[k - 1] = tiny;
    for (int rep = 0; rep < 4; ++rep) {
      std::vector<double> x = vec;
      for (int i = 0; i + 1 < k; ++i) {
        if (piv[i] == i) x[i + 1] -= dl[i] * x[i];
        else { const double t = x[i]; x[i] = x[i + 1]; x[i + 1] = t - dl[i] * x[i]; }
      }
      x[k - 1] /= d[k - 1];
      if (k > 1) x[k - 2] = (x[k - 2] - du[k - 2] * x[k - 1]) / d[k - 2];
      for (int i = k - 3; i >= 0; --i) x[i] = (x[i] - du[i] * x[i + 1] - du2[i] * x[i + 2]) / d[i];
      double nrm = 0.0;
      for (double v : x) nrm += v * v;
      nrm = std::sqrt(nrm);
      if (!(nrm > 0.0) || !std::isfinite(nrm)) break;
      for (int i = 0; i < k; ++i) vec[i] = x[i] / nrm;
    }
  }
  if (k == 1) vec[0] = 1.0;
  return theta;
}

// =================================================================================================
extern "C" {

void ls_chpl_init(void) {}      // no runtime to start (reference src/library.c:19-32 boots the Chapel runtime)
void ls_chpl_finalize(void) {}  // reference src/library.c:34
const char *dmv_last_error(void) { return g_last_error.c_str(); }
int dmv_version(void) { return 100; }
int64_t dmv_launch_count(void) { return launch_counter(); }

int dmv_context_create(const dmv_basis_desc *basis, const dmv_operator_desc *op, int device, int rank,
                       int num_ranks, dmv_context **out) {
  API_BEGIN
  if (!basis || !op || !out) throw std::runtime_error("null argument");
  if (basis->number_sites <= 0 || basis->number_sites > 64)
    throw std::runtime_error("bases with more than 64 bits are not yet implemented");  // DMV:1099-1100
  if (num_ranks < 1 || num_ranks > 256 || rank < 0 || rank >= num_ranks)
    throw std::runtime_error("need 0 <= rank < num_ranks <= 256");                     // DMV:664: uint8 keys
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0) {
    cudaGetLastError();
    throw std::runtime_error("no CUDA device: libdmv_b200 has no CPU fallback");
  }
  if (device < 0 || device >= n_dev) throw std::runtime_error("bad device ordinal");
  std::unique_ptr<dmv_context> ctx(new dmv_context());
  ctx->device = device; ctx->rank = rank; ctx->num_ranks = num_ranks;
  use_device(ctx.get());
  CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  ctx->stream = ctx->own_stream;
  CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  for (auto &e : ctx->ev) CUDA_CHECK(cudaEventCreate(&e));
  for (auto &e : ctx->ev_fill) CUDA_CHECK(cudaEventCreate(&e));
  for (auto &e : ctx->ev_chunk) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  ctx->n_sites = basis->number_sites;
  ctx->hamming_weight = basis->hamming_weight;
  ctx->spin_inversion = basis->spin_inversion;
  ctx->has_permutations = basis->has_permutations != 0;
  ctx->site_mask = basis->number_sites == 64 ? ~0ull : ((1ull << basis->number_sites) - 1);
  if (ctx->has_permutations) ctx->proj = PROJ_GROUP;                   // BO:163
  else if (ctx->spin_inversion != 0) ctx->proj = PROJ_INVERSION;       // BO:119
  else ctx->proj = PROJ_NONE;                                          // BO:89
  ctx->identity_index = (ctx->proj == PROJ_NONE && ctx->hamming_weight < 0);

  ctx->k_off_v.assign(op->off_v, op->off_v + 2 * op->n_off);
  ctx->k_off_m.assign(op->off_m, op->off_m + op->n_off); ctx->k_off_r.assign(op->off_r, op->off_r + op->n_off);
  ctx->k_off_x.assign(op->off_x, op->off_x + op->n_off); ctx->k_off_s.assign(op->off_s, op->off_s + op->n_off);
  ctx->k_diag_v.assign(op->diag_v, op->diag_v + 2 * op->n_diag);
  ctx->k_diag_m.assign(op->diag_m, op->diag_m + op->n_diag); ctx->k_diag_r.assign(op->diag_r, op->diag_r + op->n_diag);
  ctx->k_diag_s.assign(op->diag_s, op->diag_s + op->n_diag);
  if (basis->has_permutations && basis->group_order > 0 && basis->perms && basis->flips && basis->characters) {
    ctx->k_group_order = basis->group_order;
    ctx->k_perms.assign(basis->perms, basis->perms + basis->group_order * basis->number_sites);
    ctx->k_flips.assign(basis->flips, basis->flips + basis->group_order);
    ctx->k_chars.assign(basis->characters, basis->characters + 2 * basis->group_order);
  }
  bool cplx = false;
  // ---- operator: group off-diagonal terms by flip mask
  std::map<uint64_t, std::vector<OffTerm>> by_x;
  for (int64_t t = 0; t < op->n_off; ++t) {
    OffTerm o{op->off_m[t], op->off_r[t], op->off_s[t], op->off_v[2 * t], op->off_v[2 * t + 1]};
    if (op->off_x[t] == 0) throw std::runtime_error("off-diagonal term with zero flip mask");
    if (o.v_im != 0.0) cplx = true;
    by_x[op->off_x[t]].push_back(o);
  }
  std::map<uint64_t, std::vector<OffTerm>> by_x_rows;
  for (auto &kv : by_x)
    for (auto &o : kv.second) {
      // row traversal: <b|t|b^x> = v (-1)^popc(x&s) [b & m == r ^ (x & m)] (-1)^popc(b & s)
      const uint64_t x = kv.first;
      const double sg = (__builtin_popcountll(x & o.s) & 1) ? -1.0 : 1.0;
      by_x_rows[x].push_back(OffTerm{o.m, o.r ^ (x & o.m), o.s, sg * o.v_re, sg * o.v_im});
    }
  ctx->h_push = build_tables(by_x);
  ctx->h_pull = build_tables(by_x_rows);
  for (int64_t t = 0; t < op->n_diag; ++t) {
    DiagTerm d{op->diag_m[t], op->diag_r[t], op->diag_s[t], op->diag_v[2 * t], op->diag_v[2 * t + 1]};
    if (d.v_im != 0.0) cplx = true;
    ctx->h_diag.push_back(d);
  }
  {
    // k_gather needs the bit-parallel emit test on the row tables and coefficients that depend on the support
    // bits only; it runs in 32-bit registers when sites and groups fit, and skips the LUT when every emitting
    // (group, support bits) pair carries the same coefficient (every Heisenberg-type operator)
    const HostTables &H = ctx->h_pull;
    ctx->gather_ok = !H.bp.empty() && !H.any_s_out && !H.any_generic;
    ctx->gather_narrow = basis->number_sites <= 32 && H.groups.size() <= 32;
    bool first = true, uniform = ctx->gather_ok;
    for (size_t g = 0; g < H.groups.size() && uniform; ++g)
      for (int idx = 0; idx < 4; ++idx)
        if ((H.groups[g].emit_bits >> idx) & 1ull) {
          const double re = H.lut_c[2 * (4 * g + idx)], im = H.lut_c[2 * (4 * g + idx) + 1];
          if (first) { ctx->gather_uni[0] = re; ctx->gather_uni[1] = im; first = false; }
          else if (re != ctx->gather_uni[0] || im != ctx->gather_uni[1]) { uniform = false; break; }
        }
    ctx->gather_uniform = uniform && !first;
  }
  build_diag_classes(ctx.get());
  ctx->d_diag_classes.upload(ctx->h_diag_classes, ctx->stream);
  ctx->d_push.upload(ctx->h_push, ctx->stream);
  ctx->d_pull.upload(ctx->h_pull, ctx->stream);
  ctx->d_diag.upload(ctx->h_diag, ctx->stream);

  // ---- symmetry group
  if (ctx->proj == PROJ_GROUP) {
    if (basis->group_order <= 0 || !basis->perms || !basis->flips || !basis->characters)
      throw std::runtime_error("basis with permutation symmetries needs the group tables");
    ctx->host_orbit = compile_orbit_program(basis->number_sites, basis->group_order, basis->perms,
                                            basis->flips, basis->characters);
    for (double v : ctx->host_orbit.characters) (void)v;
    for (size_t e = 0; e < ctx->host_orbit.characters.size(); e += 2)
      if (ctx->host_orbit.characters[e + 1] != 0.0) cplx = true;
    upload_orbit(ctx.get());
  }
  ctx->complex_coefficients = cplx;
  ctx->rows_ok = ctx->proj == PROJ_GROUP && !cplx && ctx->host_orbit.trivial_characters &&
                 !ctx->h_pull.bp.empty() && !ctx->h_pull.any_generic;
  ctx->d_status.alloc(4);
  CUDA_CHECK(cudaMemsetAsync(ctx->d_status.ptr, 0, 4 * sizeof(unsigned long long), ctx->stream));
  ctx->d_out_count.alloc(num_ranks);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  *out = ctx.release();
  API_END
}

int dmv_context_destroy(dmv_context *ctx) {
  API_BEGIN
  if (ctx) {
    {
      std::lock_guard<std::mutex> lock(g_bind_mutex);
      for (auto it = g_bindings.begin(); it != g_bindings.end();)
        it = (it->second == ctx) ? g_bindings.erase(it) : std::next(it);
    }
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    delete ctx;
  }
  API_END
}

int dmv_set_stream(dmv_context *ctx, void *cuda_stream, int use_own_stream) {
  API_BEGIN
  use_device(ctx);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  // a NULL handle is the legacy default stream, which is what torch's default stream is
  ctx->stream = use_own_stream ? ctx->own_stream : reinterpret_cast<cudaStream_t>(cuda_stream);
  API_END
}

int dmv_set_option(dmv_context *ctx, const char *name, int64_t value) {
  API_BEGIN
  use_device(ctx);
  const std::string key(name ? name : "");
  if (key == "mode") {
    if (value < -1 || value > 1) throw std::runtime_error("mode: -1 auto, 0 push, 1 pull");
    ctx->opt_mode = (int)value;
  } else if (key == "index") {
    if (value != -1 && value != 0 && value != 2 && value != 3)
      throw std::runtime_error("index: -1 auto, 0 directory, 2 combinadic rank, 3 Lin tables");
    ctx->opt_index = (int)value;
    if (ctx->n_states >= 0) { CUDA_CHECK(cudaStreamSynchronize(ctx->stream)); select_index_mode(ctx); }
  } else if (key == "exchange") {
    if (value < -1 || value > 2)
      throw std::runtime_error("exchange: -1 auto, 0 NCCL send/recv, 1 peer-direct records, 2 replicated x (all-gather)");
    ctx->opt_exchange = (int)value;
    ctx->planned = false;
    ctx->exchange_decided = false;
    ctx->replicated = false;
    ctx->rounds.tried = false;
    ctx->rounds.ready = false;
  } else if (key == "gather") {
    if (value < -1 || value > 0) throw std::runtime_error("gather: -1 auto, 0 off (queued k_pull for mode = 1)");
    ctx->opt_gather = (int)value;
  } else if (key == "rows_ctas") {
    ctx->opt_rows_ctas = value == 3 ? 3 : 2;
    if (ctx->global) ctx->global->opt_rows_ctas = ctx->opt_rows_ctas;
  } else if (key == "rows_index") {
    if (value < -1 || value > 1) throw std::runtime_error("rows_index: -1 auto / 0 open-addressing table, 1 dense index (perfect hash)");
    ctx->opt_rows_index = (int)value;
    ctx->table_elt = 0;
    if (ctx->global) { ctx->global->opt_rows_index = (int)value; ctx->global->table_elt = 0; }
  } else if (key == "rounds") {
    if (value < -1 || value > 64) throw std::runtime_error("rounds: -1 auto, 0 / 1 one-shot exchange, R <= 64 overlapped rounds");
    ctx->opt_rounds = (int)value;
    ctx->rounds.tried = false;
    ctx->rounds.ready = false;
  } else if (key == "gather_walk") {
    ctx->opt_gather_walk = (value >= 0 && value <= 2) ? (int)value : 0;
    if (ctx->global) ctx->global->opt_gather_walk = ctx->opt_gather_walk;
  } else if (key == "peer_gather") {
    if (value < -1 || value > 0) throw std::runtime_error("peer_gather: -1 auto, 0 NCCL all-gather of x");
    ctx->opt_peer_gather = (int)value;
    ctx->exchange_decided = false;
  } else if (key == "rows") {
    if (value < -1 || value > 0) throw std::runtime_error("rows: -1 auto, 0 off (queued k_pull / k_generate for symmetric bases)");
    ctx->opt_rows = (int)value;
    if (ctx->global) ctx->global->opt_rows = (int)value;
  } else if (key == "canon") {
    ctx->opt_canon = (value >= 0 && value <= 2) ? (int)value : -1;
    if (ctx->proj == PROJ_GROUP) { CUDA_CHECK(cudaStreamSynchronize(ctx->stream)); upload_orbit(ctx); }
  } else if (key == "bitparallel") {
    ctx->opt_bitparallel = value != 0;
    ctx->planned = false;
  } else {
    throw std::runtime_error("unknown option '" + key + "'");
  }
  API_END
}

int64_t dmv_get_info(const dmv_context *ctx, const char *name) {
  const std::string key(name ? name : "");
  if (!ctx) return -1;
  if (key == "index_mode") return ctx->index_mode;
  if (key == "pull") return use_pull(ctx) ? 1 : 0;
  if (key == "gather")
    return ((use_pull(ctx) && use_gather(ctx)) || (ctx->replicated && ctx->global && use_gather(ctx->global))) ? 1 : 0;
  if (key == "gather_narrow") return ctx->gather_narrow ? 1 : 0;
  if (key == "gather_uniform") return ctx->gather_uniform ? 1 : 0;
  if (key == "peer_direct") return ctx->peer_direct ? 1 : 0;
  if (key == "replicated") return ctx->replicated ? 1 : 0;
  if (key == "replicated_block") return ctx->repl_block;
  if (key == "global_states") return ctx->global ? ctx->global->n_states : -1;
  if (key == "projection") return (int64_t)ctx->proj;
  if (key == "n_groups") return (int64_t)ctx->h_push.groups.size();
  if (key == "bp_words") return (int64_t)ctx->h_push.bp.size();
  if (key == "bp_pairs") { int64_t n = 0; for (auto &w : ctx->h_push.bp) n += w.n0 + w.n1; return n; }
  if (key == "canon_mode") return ctx->orbit.canon_mode;
  if (key == "torus_mode") return ctx->orbit.tor_mode;
  if (key == "rows")
    return ((use_pull(ctx) && !use_gather(ctx) && use_rows(ctx)) ||
            (ctx->replicated && ctx->global && !use_gather(ctx->global) && use_rows(ctx->global))) ? 1 : 0;
  if (key == "rows_ok") return ctx->rows_ok ? 1 : 0;
  if (key == "rows_dense") return ctx->dense_index ? (int64_t)ctx->mph.n_dense : (ctx->global && ctx->global->dense_index ? (int64_t)ctx->global->mph.n_dense : 0);
  if (key == "rounds") return ctx->rounds.ready ? ctx->rounds.R : 0;
  if (key == "peer_gather") return (ctx->replicated && ctx->peer_gather) ? 1 : 0;
  if (key == "complex_coefficients") return ctx->complex_coefficients ? 1 : 0;
  if (key == "canon_k") return ctx->host_orbit.canon_k;
  if (key == "orbit_n_q") return ctx->host_orbit.n_q;
  if (key == "orbit_n_t") return ctx->host_orbit.n_t;
  if (key == "orbit_n_stages") return ctx->host_orbit.n_stages;
  if (key == "group_order") return ctx->host_orbit.group_order;
  if (key == "n_buckets") return (int64_t)ctx->n_buckets;
  return -1;
}

int dmv_synchronize(dmv_context *ctx) {
  API_BEGIN
  use_device(ctx);
  check_status(ctx);   // synchronises the stream and surfaces device-side errors (DMV:115-118)
  API_END
}

int dmv_set_representatives(dmv_context *ctx, const uint64_t *representatives, int64_t count,
                            const double *norms) {
  API_BEGIN
  use_device(ctx);
  if (count < 0 || (count > 0 && !representatives)) throw std::runtime_error("bad representatives");
  if (count >= (1ll << 32)) throw std::runtime_error("more than 2^32 states per rank are not supported");
  ctx->d_reps.alloc((size_t)count);
  if (count > 0)
    CUDA_CHECK(cudaMemcpyAsync(ctx->d_reps.ptr, representatives, (size_t)count * 8, cudaMemcpyDefault, ctx->stream));
  ctx->n_states = count;
  if (ctx->proj == PROJ_GROUP) {
    ctx->d_norms.alloc((size_t)count);
    if (norms) {
      if (count > 0)
        CUDA_CHECK(cudaMemcpyAsync(ctx->d_norms.ptr, norms, (size_t)count * 8, cudaMemcpyDefault, ctx->stream));
    } else {
      launch_compute_norms(ctx->orbit, count, ctx->d_reps.ptr, ctx->d_norms.ptr, ctx->stream);
    }
  }
  install_directory(ctx);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  API_END
}

int dmv_basis_build(dmv_context *ctx) {
  API_BEGIN
  use_device(ctx);
  const int n = ctx->n_sites, w = ctx->hamming_weight;
  const bool fixed = w >= 0;
  // candidate range (mirror of ls_hs_min/max_state_estimate, reference src/ForeignTypes.chpl:102-109);
  // with spin inversion the top site is never set in a representative
  uint64_t lo, hi;
  const bool inv = ctx->spin_inversion != 0;
  if (fixed) {
    if (w > n) throw std::runtime_error("hamming weight exceeds the number of sites");
    lo = w == 0 ? 0 : ((w == 64) ? ~0ull : ((1ull << w) - 1));
    const int top = (inv && n - 1 >= w) ? n - 1 : n;
    hi = w == 0 ? 0 : (((w == 64) ? ~0ull : ((1ull << w) - 1)) << (top - w));
  } else {
    lo = 0;
    hi = inv ? (ctx->site_mask >> 1) : ctx->site_mask;
  }
  const uint64_t first_rank = fixed ? fixed_hamming_rank(lo) : lo;
  const uint64_t last_rank = fixed ? fixed_hamming_rank(hi) : hi;
  const uint64_t total = last_rank - first_rank + 1;
  uint64_t chunk_len = total / (148ull * 128 * 16);
  chunk_len = std::min<uint64_t>(std::max<uint64_t>(chunk_len, 64), 4096);
  const int64_t n_chunks = (int64_t)((total + chunk_len - 1) / chunk_len);
  std::vector<uint64_t> h_first((size_t)n_chunks), h_last((size_t)n_chunks);
  for (int64_t c = 0; c < n_chunks; ++c) {
    const uint64_t r0 = first_rank + (uint64_t)c * chunk_len;
    const uint64_t r1 = std::min(r0 + chunk_len - 1, last_rank);
    h_first[c] = fixed ? fixed_hamming_unrank(r0, w) : r0;
    h_last[c] = fixed ? fixed_hamming_unrank(r1, w) : r1;
  }
  DevBuf<uint64_t> d_first, d_last;
  DevBuf<unsigned long long> d_count, d_offset;
  d_first.upload(h_first, ctx->stream);
  d_last.upload(h_last, ctx->stream);
  d_count.alloc((size_t)n_chunks);
  d_offset.alloc((size_t)n_chunks);
  launch_enumerate(ctx->orbit, ctx->proj, ctx->site_mask, fixed, ctx->rank, ctx->num_ranks, n_chunks,
                   d_first.ptr, d_last.ptr, d_count.ptr, d_offset.ptr, nullptr, nullptr, false, ctx->stream);
  std::vector<unsigned long long> h_count((size_t)n_chunks), h_offset((size_t)n_chunks);
  CUDA_CHECK(cudaMemcpyAsync(h_count.data(), d_count.ptr, sizeof(unsigned long long) * n_chunks,
                             cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  unsigned long long acc = 0;
  for (int64_t c = 0; c < n_chunks; ++c) { h_offset[c] = acc; acc += h_count[c]; }
  if (acc >= (1ull << 32)) throw std::runtime_error("more than 2^32 states per rank are not supported");
  d_offset.upload(h_offset, ctx->stream);
  ctx->d_reps.alloc((size_t)acc);
  if (ctx->proj == PROJ_GROUP) ctx->d_norms.alloc((size_t)acc);
  launch_enumerate(ctx->orbit, ctx->proj, ctx->site_mask, fixed, ctx->rank, ctx->num_ranks, n_chunks,
                   d_first.ptr, d_last.ptr, d_count.ptr, d_offset.ptr, ctx->d_reps.ptr,
                   ctx->proj == PROJ_GROUP ? ctx->d_norms.ptr : nullptr, true, ctx->stream);
  ctx->n_states = (int64_t)acc;
  install_directory(ctx);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  API_END
}

int64_t dmv_number_states(const dmv_context *ctx) { return ctx ? ctx->n_states : -1; }

int dmv_get_representatives(dmv_context *ctx, uint64_t *representatives, double *norms) {
  API_BEGIN
  use_device(ctx);
  require_states(ctx);
  if (representatives && ctx->n_states > 0)
    CUDA_CHECK(cudaMemcpyAsync(representatives, ctx->d_reps.ptr, (size_t)ctx->n_states * 8, cudaMemcpyDefault, ctx->stream));
  if (norms && ctx->n_states > 0) {
    if (ctx->proj == PROJ_GROUP)
      CUDA_CHECK(cudaMemcpyAsync(norms, ctx->d_norms.ptr, (size_t)ctx->n_states * 8, cudaMemcpyDefault, ctx->stream));
    else {
      std::vector<double> ones((size_t)ctx->n_states, ctx->proj == PROJ_INVERSION ? std::sqrt(0.5) : 1.0);
      CUDA_CHECK(cudaMemcpyAsync(norms, ones.data(), ones.size() * 8, cudaMemcpyDefault, ctx->stream));
      CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
  }
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  API_END
}

int dmv_state_index(dmv_context *ctx, int64_t count, const uint64_t *spins, int64_t *indices) {
  API_BEGIN
  use_device(ctx);
  require_states(ctx);
  InArg<uint64_t> in(spins, (size_t)count, ctx->stream);
  OutArg<int64_t> out(indices, (size_t)count);
  KernelParams p = base_params(ctx);
  launch_state_index(p.index, count, in.ptr, out.ptr, ctx->stream);
  out.finish(ctx->stream);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  API_END
}

int dmv_state_info(dmv_context *ctx, int64_t count, const uint64_t *alphas, uint64_t *betas,
                   double *characters, double *norms) {
  API_BEGIN
  use_device(ctx);
  InArg<uint64_t> in(alphas, (size_t)count, ctx->stream);
  OutArg<uint64_t> ob(betas, (size_t)count);
  OutArg<double> oc(characters, (size_t)count * 2);
  OutArg<double> on(norms, (size_t)count);
  launch_state_info(ctx->orbit, ctx->proj, ctx->site_mask, (double)ctx->spin_inversion, count, in.ptr,
                    ob.ptr, oc.ptr, on.ptr, ctx->stream);
  ob.finish(ctx->stream); oc.finish(ctx->stream); on.finish(ctx->stream);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  API_END
}

int dmv_locale_idx_of(dmv_context *ctx, int64_t count, const uint64_t *states, int num_locales, uint8_t *keys) {
  API_BEGIN
  use_device(ctx);
  InArg<uint64_t> in(states, (size_t)count, ctx->stream);
  OutArg<uint8_t> out(keys, (size_t)count);
  launch_locale_idx(count, in.ptr, num_locales, out.ptr, ctx->stream);
  out.finish(ctx->stream);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  API_END
}

int64_t dmv_max_number_off_diag(const dmv_context *ctx) { return ctx ? (int64_t)ctx->h_push.groups.size() : -1; }

int dmv_plan(dmv_context *ctx, int64_t *send_counts) {
  API_BEGIN
  use_device(ctx);
  do_plan(ctx);
  if (send_counts) std::copy(ctx->send_counts.begin(), ctx->send_counts.end(), send_counts);
  API_END
}

int64_t dmv_number_terms(const dmv_context *ctx) { return ctx ? ctx->number_terms : -1; }

int dmv_generate(dmv_context *ctx, int elt, const void *x, void *y) {
  API_BEGIN
  use_device(ctx);
  require_states(ctx);
  if (elt != DMV_F64 && elt != DMV_C128) throw std::runtime_error("elt must be DMV_F64 or DMV_C128");
  if (!is_device_pointer(x) || !is_device_pointer(y))
    throw std::runtime_error("dmv_generate needs device pointers (y is accumulated into by later steps)");
  do_generate(ctx, elt, x, y);
  check_status(ctx);
  API_END
}

int dmv_outgoing(dmv_context *ctx, int dest, const uint64_t **betas, const double **coeffs, int64_t *count) {
  API_BEGIN
  if (!ctx->planned) throw std::runtime_error("no plan");
  if (dest < 0 || dest >= ctx->num_ranks) throw std::runtime_error("bad destination");
  const int64_t off = ctx->h_out_offset[dest];
  if (betas) *betas = ctx->d_out_betas.ptr + off;
  if (coeffs) *coeffs = ctx->d_out_coeffs.ptr + off * ctx->record_width;
  if (count) *count = ctx->h_out_offset[dest + 1] - off;
  API_END
}

int dmv_accumulate(dmv_context *ctx, int elt, int64_t count, const uint64_t *betas, const double *coeffs, void *y) {
  API_BEGIN
  use_device(ctx);
  require_states(ctx);
  if (!is_device_pointer(y)) throw std::runtime_error("dmv_accumulate needs a device y");
  const int width = complex_values(ctx, elt) ? 2 : 1;
  InArg<uint64_t> b(betas, (size_t)count, ctx->stream);
  InArg<double> c(coeffs, (size_t)count * width, ctx->stream);
  do_accumulate(ctx, elt, count, b.ptr, c.ptr, y);
  check_status(ctx);
  API_END
}

int dmv_local_matvec(dmv_context *ctx, int elt, const void *x, void *y) {
  API_BEGIN
  use_device(ctx);
  require_states(ctx);
  if (ctx->num_ranks != 1) throw std::runtime_error("dmv_local_matvec needs num_ranks == 1; use dmv_matvec");
  if (elt != DMV_F64 && elt != DMV_C128) throw std::runtime_error("elt must be DMV_F64 or DMV_C128");
  if (x == y) throw std::runtime_error("x and y must not alias");
  VecStage v = stage_vectors(ctx, elt, x, y);
  const bool host_result = v.y_host;
  if (use_pull(ctx) && (use_gather(ctx) || use_rows(ctx)) && v.y_host && ctx->n_states >= (1 << 16)) {
    // row traversal into a host y: every row chunk is final as soon as its launch ends, so its D2H copy
    // (copy stream) overlaps the gather of the next chunk
    const int chunks = dmv_context::kCopyChunks;
    const int64_t n = ctx->n_states, per = ((n + chunks - 1) / chunks + 31) / 32 * 32;
    const size_t esz = (size_t)8 * elt;
    for (int k = 0; k < chunks; ++k) {
      const int64_t b = std::min<int64_t>(n, (int64_t)k * per), e = std::min<int64_t>(n, b + per);
      if (e <= b) break;
      do_generate(ctx, elt, v.x_dev, v.y_dev, nullptr, b, e);
      CUDA_CHECK(cudaEventRecord(ctx->ev_chunk[k], ctx->stream));
      CUDA_CHECK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_chunk[k], 0));
      CUDA_CHECK(cudaMemcpyAsync(reinterpret_cast<char *>(v.y_user) + b * esz,
                                 reinterpret_cast<const char *>(v.y_dev) + b * esz, (size_t)(e - b) * esz,
                                 cudaMemcpyDeviceToHost, ctx->copy_stream));
    }
    CUDA_CHECK(cudaEventRecord(ctx->ev[2], ctx->stream));
    CUDA_CHECK(cudaEventRecord(ctx->ev[3], ctx->stream));
    CUDA_CHECK(cudaEventRecord(ctx->ev[4], ctx->stream));
    CUDA_CHECK(cudaEventRecord(ctx->ev_chunk[0], ctx->copy_stream));
    CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[0], 0));
    CUDA_CHECK(cudaEventRecord(ctx->ev[5], ctx->stream));
  } else {
  do_generate(ctx, elt, v.x_dev, v.y_dev, v.x_host_pending);
  CUDA_CHECK(cudaEventRecord(ctx->ev[2], ctx->stream));
  CUDA_CHECK(cudaEventRecord(ctx->ev[3], ctx->stream));
  finish_vectors(ctx, v);
  }
  if (host_result || !is_device_pointer(x)) {
    // host callers get a finished result (and the error check) on return
    check_status(ctx);
    collect_timings(ctx);
  }
  API_END
}

int dmv_comm_unique_id(void *id128) {
  API_BEGIN
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  NCCL_CHECK(nccl().GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  API_END
}

int dmv_comm_init(dmv_context *ctx, const void *id128) {
  API_BEGIN
  use_device(ctx);
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  NCCL_CHECK(nccl().CommInitRank(&ctx->comm, ctx->num_ranks, id, ctx->rank));
  API_END
}

int dmv_matvec(dmv_context *ctx, int elt, const void *x, void *y) {
  API_BEGIN
  use_device(ctx);
  require_states(ctx);
  if (elt != DMV_F64 && elt != DMV_C128) throw std::runtime_error("elt must be DMV_F64 or DMV_C128");
  const int P = ctx->num_ranks;
  if (P == 1) {
    const int rc = dmv_local_matvec(ctx, elt, x, y);
    if (rc) throw std::runtime_error(g_last_error);
    return 0;
  }
  if (!ctx->comm) throw std::runtime_error("dmv_matvec on several ranks needs dmv_comm_init");
  NcclApi &N = nccl();
  if (!ctx->exchange_decided) decide_exchange(ctx);
  if (ctx->replicated) {
    // ---- replicated-x product: all-gather x into equal slots, then this rank's rows by the row traversal
    if (x == y) throw std::runtime_error("x and y must not alias");
    const size_t esz = (size_t)8 * elt, bytes = (size_t)ctx->n_states * esz;
    CUDA_CHECK(cudaEventRecord(ctx->ev[0], ctx->stream));
    void *y_dev = y;
    const bool y_host = !is_device_pointer(y);
    if (y_host) {
      ctx->d_y.alloc((size_t)ctx->n_states * elt);
      y_dev = ctx->d_y.ptr;
      if (ctx->h_diag_kept == 0) CUDA_CHECK(cudaMemcpyAsync(y_dev, y, bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    const double *x_cat = ctx->d_xcat.ptr;
    if (ctx->peer_gather) {
      // ---- peer-direct: my block goes straight into slot `rank` of every rank's buffer (epoch parity picks the buffer:
      // a rank raises its flag for epoch e + 1 only after it has consumed buffer e)
      const void *x_dev = x;
      if (!is_device_pointer(x)) {
        ctx->d_x.alloc((size_t)ctx->n_states * elt);
        CUDA_CHECK(cudaMemcpyAsync(ctx->d_x.ptr, x, bytes, cudaMemcpyHostToDevice, ctx->stream));
        x_dev = ctx->d_x.ptr;
      }
      CUDA_CHECK(cudaEventRecord(ctx->ev[1], ctx->stream));
      if (ctx->peer_slot_elt != elt) upload_peer_slots(ctx, elt);
      const unsigned epoch = ++ctx->gather_epoch;
      const int b = (int)(epoch & 1u);
      const int64_t n_doubles = ctx->n_states * elt;
      const bool wide = (n_doubles % 2 == 0) && (reinterpret_cast<uintptr_t>(x_dev) % 16 == 0) &&
                        ((size_t)ctx->repl_block * elt) % 2 == 0;
      launch_push_block(x_dev, n_doubles, P, ctx->d_peer_slot[b].ptr, ctx->d_push_done.ptr, ctx->d_peer_flags.ptr,
                        ctx->rank, epoch, wide, ctx->stream);
      launch_wait_flags(ctx->d_flags.ptr, P, epoch, ctx->d_status.ptr, ctx->stream);
      x_cat = ctx->d_xcat.ptr + (size_t)b * ctx->repl_block * P * 2;
    } else {
      char *slot = reinterpret_cast<char *>(ctx->d_xcat.ptr) + (size_t)ctx->rank * ctx->repl_block * esz;
      CUDA_CHECK(cudaMemcpyAsync(slot, x, bytes, cudaMemcpyDefault, ctx->stream));   // host or device x
      CUDA_CHECK(cudaEventRecord(ctx->ev[1], ctx->stream));
      NCCL_CHECK(N.AllGather(slot, ctx->d_xcat.ptr, (size_t)ctx->repl_block * elt, ncclDouble, ctx->comm, ctx->stream));
    }
    CUDA_CHECK(cudaEventRecord(ctx->ev[6], ctx->stream));
    replicated_rows(ctx, elt, x_cat, y_dev);
    CUDA_CHECK(cudaEventRecord(ctx->ev[2], ctx->stream));
    CUDA_CHECK(cudaEventRecord(ctx->ev[3], ctx->stream));
    CUDA_CHECK(cudaEventRecord(ctx->ev[4], ctx->stream));
    if (y_host) CUDA_CHECK(cudaMemcpyAsync(y, y_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaEventRecord(ctx->ev[5], ctx->stream));
    ctx->timeline_replicated = true;
    if (y_host || !is_device_pointer(x)) {
      check_status(ctx);
      collect_timings(ctx);
    }
    return 0;
  }
  ctx->timeline_replicated = false;
  if (!ctx->rounds.tried) setup_rounds(ctx);
  if (ctx->rounds.ready) {
    // ---- record exchange in overlapped rounds (peer-direct NVLink stores + per-round flags)
    VecStage v = stage_vectors(ctx, elt, x, y);
    if (v.x_host_pending) {   // (single-rank pipelining of the upload does not apply here)
      CUDA_CHECK(cudaMemcpyAsync(ctx->d_x.ptr, v.x_host_pending, v.bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    rounds_product(ctx, elt, v.x_dev, v.y_dev);
    finish_vectors(ctx, v);
    if (v.y_host || !is_device_pointer(x)) {
      check_status(ctx);
      collect_timings(ctx);
    }
    return 0;
  }
  if (!ctx->planned) do_plan(ctx);
  if (ctx->recv_counts[0] < 0) setup_exchange(ctx);
  auto barrier = [&]() {
    NCCL_CHECK(N.AllReduce(ctx->d_barrier.ptr, ctx->d_barrier.ptr, 1, ncclInt32, ncclMax, ctx->comm, ctx->stream));
  };
  // peer-direct: nobody may overwrite my incoming buffer before I have consumed the previous product
  if (ctx->peer_direct) barrier();
  VecStage v = stage_vectors(ctx, elt, x, y);
  do_generate(ctx, elt, v.x_dev, v.y_dev, v.x_host_pending);
  CUDA_CHECK(cudaEventRecord(ctx->ev[2], ctx->stream));
  const int width = ctx->record_width;
  int64_t total_in = 0;
  if (ctx->peer_direct) {
    // the records are already in the peers' incoming buffers (NVLink stores issued by k_generate, overlapped
    // with generation); the all-reduce is the "every sender has finished" fence
    barrier();
    for (int q = 0; q < P; ++q) total_in += ctx->recv_counts[q];
  } else {
  NCCL_CHECK(N.GroupStart());
  {
    int64_t in_off = 0;
    for (int q = 0; q < P; ++q) {
      if (q == ctx->rank) continue;
      const int64_t off = ctx->h_out_offset[q], cnt = ctx->h_out_offset[q + 1] - off;
      if (cnt > 0) {
        NCCL_CHECK(N.Send(ctx->d_out_betas.ptr + off, (size_t)cnt, ncclUint64, q, ctx->comm, ctx->stream));
        NCCL_CHECK(N.Send(ctx->d_out_coeffs.ptr + off * width, (size_t)cnt * width, ncclDouble, q, ctx->comm, ctx->stream));
      }
      const int64_t rc = ctx->recv_counts[q];
      if (rc > 0) {
        NCCL_CHECK(N.Recv(ctx->d_in_betas.ptr + in_off, (size_t)rc, ncclUint64, q, ctx->comm, ctx->stream));
        NCCL_CHECK(N.Recv(ctx->d_in_coeffs.ptr + in_off * width, (size_t)rc * width, ncclDouble, q, ctx->comm, ctx->stream));
      }
      in_off += rc;
    }
    total_in = in_off;
  }
  NCCL_CHECK(N.GroupEnd());
  }
  CUDA_CHECK(cudaEventRecord(ctx->ev[3], ctx->stream));
  do_accumulate(ctx, elt, total_in, ctx->d_in_betas.ptr, ctx->d_in_coeffs.ptr, v.y_dev);
  finish_vectors(ctx, v);
  if (v.y_host || !is_device_pointer(x)) {
    check_status(ctx);
    collect_timings(ctx);
  }
  API_END
}

// ---- block <-> hashed redistribution ("next" row f2)
int dmv_hashed_positions(dmv_context *ctx, int64_t count, const uint8_t *masks, int num_ranks, int64_t *counts,
                         uint32_t *positions) {
  API_BEGIN
  use_device(ctx);
  if (count < 0 || num_ranks < 1) throw std::runtime_error("bad arguments");
  InArg<uint8_t> m(masks, (size_t)count, ctx->stream);
  OutArg<uint32_t> out(positions, (size_t)count);
  std::vector<int64_t> c;
  hashed_positions(ctx, count, m.ptr, num_ranks, c, out.ptr);
  if (counts) std::copy(c.begin(), c.end(), counts);
  out.finish(ctx->stream);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  API_END
}

int dmv_permute(dmv_context *ctx, int elt, int64_t count, const uint32_t *positions, const void *in, void *out,
                int gather) {
  API_BEGIN
  use_device(ctx);
  if (elt != 1 && elt != 2) throw std::runtime_error("elt must be 1 (8-byte) or 2 (16-byte elements)");
  if (in == out) throw std::runtime_error("in and out must not alias");
  InArg<uint32_t> p(positions, (size_t)count, ctx->stream);
  InArg<double> i(reinterpret_cast<const double *>(in), (size_t)count * elt, ctx->stream);
  OutArg<double> o(reinterpret_cast<double *>(out), (size_t)count * elt);
  launch_permute(count, elt, p.ptr, i.ptr, o.ptr, gather != 0, ctx->stream);
  o.finish(ctx->stream);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  API_END
}

int dmv_block_to_hashed(dmv_context *ctx, int elt, int64_t chunk_count, const uint8_t *masks_chunk,
                        const void *block_chunk, void *hashed, int64_t hashed_count) {
  API_BEGIN
  use_device(ctx);
  if (elt != 1 && elt != 2) throw std::runtime_error("elt must be 1 (8-byte) or 2 (16-byte elements)");
  const int P = ctx->num_ranks;
  InArg<uint8_t> m(masks_chunk, (size_t)chunk_count, ctx->stream);
  InArg<double> in(reinterpret_cast<const double *>(block_chunk), (size_t)chunk_count * elt, ctx->stream);
  OutArg<double> out(reinterpret_cast<double *>(hashed), (size_t)hashed_count * elt);
  DevBuf<uint32_t> d_pos;
  DevBuf<double> d_grouped;
  d_pos.alloc((size_t)chunk_count);
  d_grouped.alloc((size_t)chunk_count * elt);
  std::vector<int64_t> counts;
  hashed_positions(ctx, chunk_count, m.ptr, P, counts, d_pos.ptr);
  launch_permute(chunk_count, elt, d_pos.ptr, in.ptr, d_grouped.ptr, false, ctx->stream);
  if (P == 1) {
    if (hashed_count != chunk_count) throw std::runtime_error("hashed block size does not match the masks");
    CUDA_CHECK(cudaMemcpyAsync(out.ptr, d_grouped.ptr, (size_t)chunk_count * elt * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  } else {
    if (!ctx->comm) throw std::runtime_error("dmv_block_to_hashed on several ranks needs dmv_comm_init");
    NcclApi &N = nccl();
    const std::vector<int64_t> all = all_gather_counts(ctx, counts);   // all[r * P + q]: chunk r holds for owner q
    int64_t incoming = 0;
    for (int r = 0; r < P; ++r) incoming += all[(size_t)r * P + ctx->rank];
    if (incoming != hashed_count) throw std::runtime_error("hashed block size does not match the masks");
    NCCL_CHECK(N.GroupStart());
    int64_t send_off = 0, recv_off = 0;
    for (int q = 0; q < P; ++q) {
      const int64_t sc = counts[q], rc = all[(size_t)q * P + ctx->rank];
      if (q == ctx->rank) {
        if (sc > 0) CUDA_CHECK(cudaMemcpyAsync(out.ptr + recv_off * elt, d_grouped.ptr + send_off * elt, (size_t)sc * elt * 8,
                                               cudaMemcpyDeviceToDevice, ctx->stream));
      } else {
        if (sc > 0) NCCL_CHECK(N.Send(d_grouped.ptr + send_off * elt, (size_t)sc * elt, ncclDouble, q, ctx->comm, ctx->stream));
        if (rc > 0) NCCL_CHECK(N.Recv(out.ptr + recv_off * elt, (size_t)rc * elt, ncclDouble, q, ctx->comm, ctx->stream));
      }
      send_off += sc;
      recv_off += rc;
    }
    NCCL_CHECK(N.GroupEnd());
  }
  out.finish(ctx->stream);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  API_END
}

int dmv_hashed_to_block(dmv_context *ctx, int elt, int64_t chunk_count, const uint8_t *masks_chunk,
                        const void *hashed, int64_t hashed_count, void *block_chunk) {
  API_BEGIN
  use_device(ctx);
  if (elt != 1 && elt != 2) throw std::runtime_error("elt must be 1 (8-byte) or 2 (16-byte elements)");
  const int P = ctx->num_ranks;
  InArg<uint8_t> m(masks_chunk, (size_t)chunk_count, ctx->stream);
  InArg<double> in(reinterpret_cast<const double *>(hashed), (size_t)hashed_count * elt, ctx->stream);
  OutArg<double> out(reinterpret_cast<double *>(block_chunk), (size_t)chunk_count * elt);
  DevBuf<uint32_t> d_pos;
  DevBuf<double> d_grouped;
  d_pos.alloc((size_t)chunk_count);
  d_grouped.alloc((size_t)chunk_count * elt);
  std::vector<int64_t> counts;   // counts[q]: positions of MY chunk owned by q = what q sends me
  hashed_positions(ctx, chunk_count, m.ptr, P, counts, d_pos.ptr);
  if (P == 1) {
    if (hashed_count != chunk_count) throw std::runtime_error("hashed block size does not match the masks");
    CUDA_CHECK(cudaMemcpyAsync(d_grouped.ptr, in.ptr, (size_t)chunk_count * elt * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  } else {
    if (!ctx->comm) throw std::runtime_error("dmv_hashed_to_block on several ranks needs dmv_comm_init");
    NcclApi &N = nccl();
    const std::vector<int64_t> all = all_gather_counts(ctx, counts);   // all[r * P + q]: chunk r needs from owner q
    int64_t outgoing = 0;
    for (int r = 0; r < P; ++r) outgoing += all[(size_t)r * P + ctx->rank];
    if (outgoing != hashed_count) throw std::runtime_error("hashed block size does not match the masks");
    NCCL_CHECK(N.GroupStart());
    int64_t send_off = 0, recv_off = 0;
    for (int q = 0; q < P; ++q) {
      // my hashed block is ascending in global position: the part for chunk q follows the parts for chunks < q
      const int64_t sc = all[(size_t)q * P + ctx->rank], rc = counts[q];
      if (q == ctx->rank) {
        if (sc > 0) CUDA_CHECK(cudaMemcpyAsync(d_grouped.ptr + recv_off * elt, in.ptr + send_off * elt, (size_t)sc * elt * 8,
                                               cudaMemcpyDeviceToDevice, ctx->stream));
      } else {
        if (sc > 0) NCCL_CHECK(N.Send(in.ptr + send_off * elt, (size_t)sc * elt, ncclDouble, q, ctx->comm, ctx->stream));
        if (rc > 0) NCCL_CHECK(N.Recv(d_grouped.ptr + recv_off * elt, (size_t)rc * elt, ncclDouble, q, ctx->comm, ctx->stream));
      }
      send_off += sc;
      recv_off += rc;
    }
    NCCL_CHECK(N.GroupEnd());
  }
  launch_permute(chunk_count, elt, d_pos.ptr, d_grouped.ptr, out.ptr, true, ctx->stream);
  out.finish(ctx->stream);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  API_END
}

// Replicated-x product without a communicator (the host owns the all-gather): set-up, then rows of H applied to
// a caller-assembled x_cat (rank r's block at r * dmv_get_info("replicated_block") elements).  Device pointers.
int dmv_replicated_setup(dmv_context *ctx) {
  API_BEGIN
  use_device(ctx);
  setup_replicated(ctx);
  API_END
}

int dmv_replicated_product(dmv_context *ctx, int elt, const void *x_cat, void *y) {
  API_BEGIN
  use_device(ctx);
  require_states(ctx);
  if (!ctx->global || ctx->repl_block <= 0) throw std::runtime_error("dmv_replicated_setup has not run");
  if (elt != DMV_F64 && elt != DMV_C128) throw std::runtime_error("elt must be DMV_F64 or DMV_C128");
  if (!is_device_pointer(x_cat) || !is_device_pointer(y)) throw std::runtime_error("dmv_replicated_product needs device pointers");
  replicated_rows(ctx, elt, x_cat, y);
  check_status(ctx);
  API_END
}

// ---- several vectors per call (the reference's numVectors > 1, "not yet implemented" there: DMV:1101-1102, and what
// PRIMME's blockSize > 1 would use, src/Diagonalize.chpl:154-158).  x, y: num_vectors arrays of dmv_number_states
// elements, one after the other (the [numVectors, N] layout of the reference's BlockVector).  On one rank with device
// pointers and an operator k_gather applies to, four vectors share one walk over the terms and one index look-up per
// term; every other case is the loop over single products.
int dmv_matvec_batch(dmv_context *ctx, int elt, int num_vectors, const void *x, void *y) {
  API_BEGIN
  use_device(ctx);
  require_states(ctx);
  if (elt != DMV_F64 && elt != DMV_C128) throw std::runtime_error("elt must be DMV_F64 or DMV_C128");
  if (num_vectors < 1) throw std::runtime_error("num_vectors must be positive");
  if (x == y) throw std::runtime_error("x and y must not alias");
  const size_t vec_bytes = (size_t)ctx->n_states * 8 * elt;
  const char *xb = reinterpret_cast<const char *>(x);
  char *yb = reinterpret_cast<char *>(y);
  int k = 0;
  if (ctx->num_ranks == 1 && use_pull(ctx) && use_gather(ctx) && is_device_pointer(x) && is_device_pointer(y)) {
    for (; k + 4 <= num_vectors; k += 4) {
      KernelParams p = base_params(ctx);
      p.x = xb + (size_t)k * vec_bytes;
      p.y = yb + (size_t)k * vec_bytes;
      p.batch = 4;
      p.batch_stride = ctx->n_states;
      select_tables(ctx, p, true, ctx->complex_coefficients);
      p.row_split = choose_row_split(ctx->n_states, (int)ctx->h_pull.groups.size());
      p.uni_re = ctx->gather_uni[0]; p.uni_im = ctx->gather_uni[1];
      launch_gather(p, ctx->proj == PROJ_INVERSION, ctx->complex_coefficients, elt == DMV_C128, ctx->gather_narrow,
                    ctx->index_mode == INDEX_LIN, ctx->gather_uniform, ctx->stream);
    }
  }
  for (; k < num_vectors; ++k) {
    const int rc = ctx->num_ranks == 1 ? dmv_local_matvec(ctx, elt, xb + (size_t)k * vec_bytes, yb + (size_t)k * vec_bytes)
                                       : dmv_matvec(ctx, elt, xb + (size_t)k * vec_bytes, yb + (size_t)k * vec_bytes);
    if (rc) throw std::runtime_error(g_last_error);
  }
  API_END
}

// ---- Lanczos ground-state solver on the device ("next" row f3): the consumer of the product.  The reference hands its
// matvec to PRIMME (src/Diagonalize.chpl:134-225); here the three-term recurrence, its dot products (NCCL all-reduce
// across ranks) and the Ritz-vector accumulation all stay in HBM, only alpha_j / beta_j (two doubles) visit the host.
int dmv_lanczos(dmv_context *ctx, int elt, int max_iters, double tol, uint64_t seed, double *eigenvalue,
                void *eigenvector, int *iterations, double *residual) {
  API_BEGIN
  use_device(ctx);
  require_states(ctx);
  if (elt != DMV_F64 && elt != DMV_C128) throw std::runtime_error("elt must be DMV_F64 or DMV_C128");
  if (max_iters < 1) throw std::runtime_error("max_iters must be positive");
  const int P = ctx->num_ranks;
  if (P > 1 && !ctx->comm) throw std::runtime_error("dmv_lanczos on several ranks needs dmv_comm_init");
  const int64_t n = ctx->n_states;
  const size_t words = (size_t)n * elt;
  const bool ce = elt == DMV_C128;
  for (auto &b : ctx->lz_v) b.alloc(words);
  ctx->lz_scal.alloc(8);
  double *scal = ctx->lz_scal.ptr;
  cudaStream_t st = ctx->stream;
  auto reduce = [&](int count) {   // sum the first `count` scalars over the ranks, bring them to the host
    if (P > 1) NCCL_CHECK(nccl().AllReduce(scal, scal, (size_t)count, ncclDouble, ncclSum, ctx->comm, st));
    double h[4] = {0, 0, 0, 0};
    CUDA_CHECK(cudaMemcpyAsync(h, scal, sizeof(double) * count, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    return std::vector<double>(h, h + count);
  };
  auto product = [&](const double *x, double *y) {
    CUDA_CHECK(cudaMemsetAsync(y, 0, words * 8, st));   // operators without a diagonal accumulate into y (DMV:1062-1069)
    const int rc = P == 1 ? dmv_local_matvec(ctx, elt, x, y) : dmv_matvec(ctx, elt, x, y);
    if (rc) throw std::runtime_error(g_last_error);
  };
  auto start_vector = [&](double *v) {
    launch_fill((int64_t)words, seed, (uint64_t)ctx->rank << 40, v, st);
    CUDA_CHECK(cudaMemsetAsync(scal, 0, 8 * sizeof(double), st));
    launch_dot(n, ce, v, v, scal, st);
    const double nrm = std::sqrt(reduce(1)[0]);
    if (!(nrm > 0.0)) throw std::runtime_error("empty basis");
    launch_scale((int64_t)words, 1.0 / nrm, v, v, false, st);
  };
  std::vector<double> alphas, betas, ritz;
  double theta = 0.0, res = 0.0;
  // every rank must take the same stopping decision: the Krylov space is exhausted at the GLOBAL dimension
  int64_t n_global = n;
  if (P > 1) {
    const double mine = (double)n;
    CUDA_CHECK(cudaMemcpyAsync(scal, &mine, sizeof(double), cudaMemcpyHostToDevice, st));
    n_global = (int64_t)std::llround(reduce(1)[0]);
  }
  {
    double *v = ctx->lz_v[0].ptr, *u = ctx->lz_v[1].ptr, *w = ctx->lz_v[2].ptr;
    start_vector(v);
    double beta_prev = 0.0;
    for (int j = 0; j < max_iters; ++j) {
      product(v, w);
      CUDA_CHECK(cudaMemsetAsync(scal, 0, 8 * sizeof(double), st));
      launch_dot(n, ce, v, w, scal, st);
      const double alpha = reduce(1)[0];
      const double coef[2] = {alpha, beta_prev};
      CUDA_CHECK(cudaMemcpyAsync(scal + 4, coef, sizeof(coef), cudaMemcpyHostToDevice, st));
      CUDA_CHECK(cudaMemsetAsync(scal, 0, sizeof(double), st));
      launch_lanczos_update(n, ce, w, v, j > 0 ? u : nullptr, scal + 4, scal, st);
      const double beta = std::sqrt(std::max(0.0, reduce(1)[0]));
      alphas.push_back(alpha);
      theta = tridiagonal_lowest(alphas, betas, ritz);
      res = std::fabs(beta * ritz.back());
      const bool done = res <= tol * std::max(1.0, std::fabs(theta)) || beta <= 1e-14 * std::max(1.0, std::fabs(alpha)) ||
                        (int64_t)alphas.size() >= n_global;
      if (done || j + 1 == max_iters) break;
      betas.push_back(beta);
      launch_scale((int64_t)words, 1.0 / beta, w, w, false, st);
      double *t = u; u = v; v = w; w = t;   // v_prev <- v, v <- w / beta, old v_prev becomes scratch
      beta_prev = beta;
    }
  }
  if (eigenvalue) *eigenvalue = theta;
  if (iterations) *iterations = (int)alphas.size();
  if (residual) *residual = res;
  if (eigenvector) {
    // second pass with the stored alpha / beta (no dot products): Ritz vector = sum_j s_j v_j
    double *v = ctx->lz_v[0].ptr, *u = ctx->lz_v[1].ptr, *w = ctx->lz_v[2].ptr, *acc = ctx->lz_v[3].ptr;
    start_vector(v);
    CUDA_CHECK(cudaMemsetAsync(acc, 0, words * 8, st));
    const int k = (int)alphas.size();
    for (int j = 0; j < k; ++j) {
      launch_scale((int64_t)words, ritz[j], v, acc, true, st);
      if (j + 1 == k) break;
      product(v, w);
      const double coef[2] = {alphas[j], j > 0 ? betas[j - 1] : 0.0};
      CUDA_CHECK(cudaMemcpyAsync(scal + 4, coef, sizeof(coef), cudaMemcpyHostToDevice, st));
      launch_lanczos_update(n, ce, w, v, j > 0 ? u : nullptr, scal + 4, scal, st);
      CUDA_CHECK(cudaStreamSynchronize(st));   // coef lives on the host stack
      launch_scale((int64_t)words, 1.0 / betas[j], w, w, false, st);
      double *t = u; u = v; v = w; w = t;
    }
    CUDA_CHECK(cudaMemcpyAsync(eigenvector, acc, words * 8, cudaMemcpyDefault, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
  }
  check_status(ctx);
  API_END
}

int dmv_last_timings(dmv_context *ctx, double *ms, int capacity) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  if (cudaStreamSynchronize(ctx->stream) == cudaSuccess) {
    try { collect_timings(ctx); } catch (...) {}
  }
  for (int i = 0; i < T_COUNT && i < capacity; ++i) ms[i] = ctx->timings[i];
  return T_COUNT;
}
const char *dmv_timing_name(int i) { return (i >= 0 && i < T_COUNT) ? kTimingNames[i] : ""; }

int dmv_compute_off_diag(dmv_context *ctx, int64_t count, const uint64_t *alphas, const void *xs, int elt,
                         int64_t *n, uint64_t *betas, double *coeffs, uint8_t *keys) {
  API_BEGIN
  // BatchedOperator.computeOffDiag (reference src/BatchedOperator.chpl:82-213) through the same kernel
  // as the product: the given alphas play the role of the source block and every record is written to
  // one flat output (emit_all) together with its locale key.
  use_device(ctx);
  if (elt != DMV_F64 && elt != DMV_C128) throw std::runtime_error("elt must be DMV_F64 or DMV_C128");
  const size_t cap = (size_t)count * std::max<size_t>(1, ctx->h_push.groups.size());
  InArg<uint64_t> a(alphas, (size_t)count, ctx->stream);
  InArg<double> x(reinterpret_cast<const double *>(xs), (size_t)count * elt, ctx->stream);
  OutArg<uint64_t> ob(betas, cap);
  OutArg<double> oc(coeffs, cap * 2);
  OutArg<uint8_t> ok(keys, cap);
  DevBuf<double> d_src_norms;
  DevBuf<int64_t> d_off;
  DevBuf<unsigned long long> d_cnt;
  std::vector<int64_t> off = {0, (int64_t)cap};
  d_off.upload(off, ctx->stream);
  d_cnt.alloc(1);
  CUDA_CHECK(cudaMemsetAsync(d_cnt.ptr, 0, sizeof(unsigned long long), ctx->stream));
  KernelParams p = base_params(ctx);
  p.index.reps = a.ptr; p.index.n = count; p.index.mode = INDEX_DIRECTORY;
  if (ctx->proj == PROJ_GROUP) {  // norms of the sources: BO:178-194 appends the alphas to state_info
    d_src_norms.alloc((size_t)count);
    launch_compute_norms(ctx->orbit, count, a.ptr, d_src_norms.ptr, ctx->stream);
    p.norms = d_src_norms.ptr;
  }
  p.x = x.ptr; p.y = nullptr;
  p.emit_all = 1;
  p.out_betas = ob.ptr; p.out_coeffs = oc.ptr; p.out_keys = ok.ptr;
  p.out_offset = d_off.ptr; p.out_count = d_cnt.ptr;
  p.row_begin = 0; p.row_end = count;
  select_tables(ctx, p, false, true);
  launch_generate(p, ctx->proj, /*complex values*/ true, elt == DMV_C128, false, ctx->stream);
  unsigned long long total = 0;
  CUDA_CHECK(cudaMemcpyAsync(&total, d_cnt.ptr, sizeof(total), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  if (total > cap) throw std::runtime_error("dmv_compute_off_diag: output overflow");
  ob.finish(ctx->stream, (size_t)total); oc.finish(ctx->stream, (size_t)total * 2); ok.finish(ctx->stream, (size_t)total);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  if (n) *n = (int64_t)total;
  API_END
}

// ---- plugin kernels: ls_chpl_operator_apply_diag / _apply_off_diag (reference src/BatchedOperator.chpl:217-275)
int dmv_apply_diag(dmv_context *ctx, int64_t count, const uint64_t *alphas, double *coeffs) {
  API_BEGIN
  use_device(ctx);
  if (ctx->proj != PROJ_NONE) throw std::runtime_error("bases that require projection are not yet supported");  // BO:226-227
  if (count < 0) throw std::runtime_error("negative count");
  InArg<uint64_t> a(alphas, (size_t)count, ctx->stream);
  OutArg<double> out(coeffs, (size_t)count);
  KernelParams p = base_params(ctx);
  select_tables(ctx, p, false, true);
  launch_apply_diag(p, count, a.ptr, out.ptr, ctx->stream);
  out.finish(ctx->stream);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  API_END
}

int dmv_apply_off_diag(dmv_context *ctx, int64_t count, const uint64_t *alphas, uint64_t *betas, double *coeffs,
                       int64_t *offsets) {
  API_BEGIN
  use_device(ctx);
  if (ctx->proj != PROJ_NONE) throw std::runtime_error("bases that require projection are not yet supported");  // BO:247-248
  if (count < 0) throw std::runtime_error("negative count");
  const size_t cap = (size_t)count * std::max<size_t>(1, ctx->h_push.groups.size());
  InArg<uint64_t> a(alphas, (size_t)count, ctx->stream);
  OutArg<uint64_t> ob(betas, cap);
  OutArg<double> oc(coeffs, cap * 2);
  OutArg<int64_t> oo(offsets, (size_t)count + 1);
  DevBuf<int64_t> d_counts;
  d_counts.alloc((size_t)count + 1);
  KernelParams p = base_params(ctx);
  select_tables(ctx, p, false, true);
  launch_apply_off_diag(p, count, a.ptr, nullptr, d_counts.ptr, nullptr, nullptr, false, ctx->stream);
  std::vector<int64_t> h((size_t)count + 1, 0);
  if (count > 0)
    CUDA_CHECK(cudaMemcpyAsync(h.data(), d_counts.ptr, (size_t)count * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  int64_t acc = 0;
  for (int64_t i = 0; i < count; ++i) { const int64_t c = h[i]; h[i] = acc; acc += c; }   // CSR row pointer (BO:109)
  h[count] = acc;
  CUDA_CHECK(cudaMemcpyAsync(oo.ptr, h.data(), ((size_t)count + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  launch_apply_off_diag(p, count, a.ptr, oo.ptr, nullptr, ob.ptr, oc.ptr, true, ctx->stream);
  ob.finish(ctx->stream, (size_t)acc); oc.finish(ctx->stream, (size_t)acc * 2); oo.finish(ctx->stream);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  API_END
}

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

// host-only self-check entry for the tridiagonal solver behind dmv_lanczos (no device needed)
int dmv_debug_tridiagonal_lowest(int k, const double *diag, const double *offdiag, double *eigenvalue, double *vector) {
  API_BEGIN
  if (k < 1) throw std::runtime_error("empty matrix");
  std::vector<double> a(diag, diag + k), b(offdiag, offdiag + (k - 1)), v;
  *eigenvalue = tridiagonal_lowest(a, b, v);
  if (vector) std::copy(v.begin(), v.end(), vector);
  API_END
}

int dmv_debug_compile_group(const dmv_basis_desc *basis, int64_t *info, int64_t count,
                            const uint64_t *states, uint64_t *reps, int32_t *stab) {
  API_BEGIN
  HostOrbitProgram H = compile_orbit_program(basis->number_sites, basis->group_order, basis->perms,
                                             basis->flips, basis->characters);
  if (info) {
    info[0] = H.n_q; info[1] = H.n_stages; info[2] = H.n_t; info[3] = H.n_left; info[4] = H.n_right;
    info[5] = H.has_flip;
    if (count < 0) {   // extended query (count = -1): info must hold 12 entries
      info[6] = H.canon_mode; info[7] = H.canon_k; info[8] = H.canon_r;
      info[9] = H.canon_lut2.empty() ? 0 : 1;
      info[10] = H.cc_begin.empty() ? 0 : (int64_t)H.cc_begin.size() - 1;
      info[11] = (int64_t)H.cc_mask.size();
    }
    if (count < -1) {  // count = -2: info holds 16 entries
      info[12] = H.tor_mode; info[13] = H.tor_rho_n; info[14] = H.tor_tau_n; info[15] = H.chain_dihedral;
    }
  }
  OrbitProgram P = H.view();
  for (int64_t k = 0; k < count; ++k) {
    const OrbitResult r = orbit_scan<true, false>(P, states[k]);
    if (P.canon_mode && orbit_min_canon(P, states[k]) != r.rep)
      throw std::runtime_error("canonical form disagrees with the chain walk");
    if (P.tor_mode == 2 && P.canon_k == P.canon_r && (P.canon_k == 4 || P.canon_k == 6)) {
      const uint64_t got = P.canon_k == 6 ? orbit_min_torus_sq<6>(P, states[k]) : orbit_min_torus_sq<4>(P, states[k]);
      if (got != r.rep) throw std::runtime_error("square-torus canonical form disagrees with the chain walk");
    }
    if (P.tor_mode || P.chain_dihedral) {
      OrbitProgram P1 = P;
      P1.tor_mode = 0;
      P1.chain_dihedral = 0;
      if (orbit_min_canon(P1, states[k]) != r.rep)
        throw std::runtime_error("block-rotation canonical form disagrees with the chain walk");
    }
    if (reps) reps[k] = r.rep;
    if (stab) stab[k] = r.stab;
  }
  API_END
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
