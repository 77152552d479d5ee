// ksched.cu — sm_100a kernels and the C-ABI of include/ksched.h.
//
//  K0  sort_keys / gather_rows   FFD order of the queue (queue.go:35-110) and the dense, FFD-ordered
//                                P x 256 B pod-row matrix the feasibility kernel streams.
//  K1  feasibility_kernel        dense pods x (template, instance type) bitmask F, bit-sliced over columns:
//                                Requirements.Compatible/Intersects (requirements.go:123-206), fits /
//                                hasOffering (node.go:143-159), Taints.Tolerates (taints.go:28) as AND/OR of
//                                precomputed column bitsets; warp ballot/ffs gives the per-pod best column.
//  K2  pack_kernel               the sequential first-fit of Scheduler.add (scheduler.go:174-219) with
//                                Node.Add / ExistingNode.Add (node.go:62-107, existingnode.go:77-130),
//                                topology domains (topology.go:120-167, topologygroup.go:88-243), queue
//                                requeue / relaxation (scheduler.go:104-124, queue.go:44-68) in ONE persistent
//                                CTA: candidate nodes are examined in parallel, the reference's scan order is
//                                recovered with a block-wide argmin on (pod count, stable tie-break).
// Integer / bitmask work only — no tensor cores by design (BASELINE.json north_star).
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdint.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cub/cub.cuh>
#include <string>
#include <mutex>
#include <vector>

#include "ksched.h"
#include "reqmask.cuh"
#include "launch.cuh"

using ksched::KeyMeta;
using ksched::Req;

#define CUDA_TRY(h, expr)                                                                         \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      (h)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                              \
      return KSCHED_ERR_CUDA;                                                                     \
    }                                                                                             \
  } while (0)

namespace {

constexpr int kMaxCG = 8;        // topology groups that may constrain one pod class
constexpr int kMaxTouched = 8;   // requirement keys one Add may touch (pod keys + topology keys)
constexpr int kPackThreads = 512;
constexpr int kFreshMemoSlots = 8192;  // PackState::fd_*
constexpr uint64_t kNoBest = ~0ull;

// ------------------------------------------------------------------------------------------------
// Device-side catalog: instance-type columns as bit-sliced tables (one bit per column, u32 words).
// ------------------------------------------------------------------------------------------------
struct DevCatalog {
  int n_keys, n_res, n_types, n_templates, W32;
  const ksched_keyinfo* keys;        // [n_keys]
  const int64_t* key_int_values;     // [n_keys][64]
  const ksched_template* templates;  // [n_templates]
  const ksched_type_row* types;      // [n_types]
  const int64_t* capacity;           // [n_types][8]
  const float* price32;              // [n_types]
  const int16_t* valrow;             // [n_keys*64] row in valset or -1
  const uint32_t* valset;            // [rows][W32] type has a positive requirement on key containing value
  const uint32_t* absent;            // [n_keys][W32] type has no requirement on key
  const uint32_t* negempty;          // [n_keys][W32] type requirement on key is DoesNotExist
  uint32_t type_relevant;            // bit k: some type defines key k
  const int16_t* offrow;             // [64] row in offset table or -1
  const uint32_t* offset;            // [rows][W32] type has an available offering (ct*16+zone)
  const uint32_t* anyoffer;          // [W32]
  const uint32_t* member;            // [n_templates][W32]
  const int64_t* alloc_sorted;       // [n_res][n_types] ascending
  const uint32_t* fitset;            // [n_res][n_types+1][W32]  rank -> types with alloc >= alloc_sorted[rank]
  const int32_t* perm_desc;          // [n_res][n_types] types by descending allocatable
  const int64_t* alloc_rt;           // [n_res][n_types] allocatable, resource-major
  const uint32_t* domset;            // [n_types][W32] types whose allocatable vector (first 4 resources) is dominated by the row's type
  int zone_key, ct_key;
  const uint64_t* offer_keys;        // [n_types][64] launch-choice keys (ksched_catalog.offering_keys) or nullptr
  const uint32_t* input_index;       // [n_types] provider input order of the column
};

__device__ __forceinline__ KeyMeta key_meta(const DevCatalog& c, int k) {
  return KeyMeta{c.keys[k].int_mask, c.key_int_values ? c.key_int_values + (size_t)k * 64 : nullptr};
}

// number of types with alloc_r < q  (lower bound)
__device__ __forceinline__ int fit_rank(const int64_t* alloc_sorted, int n_types, int r, int64_t q) {
  const int64_t* a = alloc_sorted + (size_t)r * n_types;
  int lo = 0, hi = n_types;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] < q) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Column set {t : instanceType.Requirements.Intersects(node) holds on key k} for one u32 word
// (requirements.go:189-206 with the type as receiver). `allowed` = dictionary values the node requirement
// admits, `neg` = its operator is NotIn/DoesNotExist.
__device__ __forceinline__ uint32_t key_typeset_word(const DevCatalog& c, int k, uint64_t allowed, bool neg, int w) {
  uint32_t s = c.absent[(size_t)k * c.W32 + w];
  if (neg) s |= c.negempty[(size_t)k * c.W32 + w];
  uint64_t m = allowed;
  while (m) {
    int b = __ffsll((long long)m) - 1;
    m &= m - 1;
    int row = c.valrow[k * 64 + b];
    if (row >= 0) s |= c.valset[(size_t)row * c.W32 + w];
  }
  return s;
}
// hasOffering (node.go:151-159) for one word: zmask / cmask = admitted zone / capacity-type value bits
__device__ __forceinline__ uint32_t offer_word(const DevCatalog& c, uint32_t zmask, uint32_t cmask, bool unconstrained, int w) {
  if (unconstrained) return c.anyoffer[w];
  uint32_t s = 0;
  uint32_t cm = cmask & 0xF;
  while (cm) {
    int ct = __ffs(cm) - 1;
    cm &= cm - 1;
    uint32_t zm = zmask & 0xFFFF;
    while (zm) {
      int z = __ffs(zm) - 1;
      zm &= zm - 1;
      int row = c.offrow[ct * 16 + z];
      if (row >= 0) s |= c.offset[(size_t)row * c.W32 + w];
    }
  }
  return s;
}

// ------------------------------------------------------------------------------------------------
// K0: queue order
// ------------------------------------------------------------------------------------------------
__global__ void sort_keys_kernel(int n, const ksched_pod_row* __restrict__ classes, const uint32_t* __restrict__ pod_class,
                                 const int64_t* __restrict__ ts, const uint32_t* __restrict__ uid_rank,
                                 uint64_t* k_cpu, uint64_t* k_mem, uint64_t* k_tie, uint32_t* idx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const ksched_pod_row& row = classes[pod_class[i]];
  // descending cpu / memory -> ascending on the complemented value (milli-units are < 2^62)
  k_cpu[i] = ~(uint64_t)(row.requests[0] + (1ll << 62));
  k_mem[i] = ~(uint64_t)(row.requests[1] + (1ll << 62));
  k_tie[i] = ((uint64_t)(ts[i] + (1ll << 32)) << 30) | (uint64_t)uid_rank[i];
  idx[i] = (uint32_t)i;
}
__global__ void gather_u64_kernel(int n, const uint64_t* __restrict__ src, const uint32_t* __restrict__ idx, uint64_t* dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}
// FFD-ordered dense pod-row matrix: row j = class row of the j-th pod of the queue. One warp per row,
// 256-byte coalesced loads and stores.
__global__ void gather_rows_kernel(int n, const ksched_pod_row* __restrict__ classes, const uint32_t* __restrict__ pod_class,
                                   const uint32_t* __restrict__ order, uint64_t* __restrict__ rows) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int j = warp; j < n; j += nwarps) {
    const uint64_t* src = reinterpret_cast<const uint64_t*>(&classes[pod_class[order[j]]]);
    rows[(size_t)j * KSCHED_ROW_WORDS + lane] = src[lane];
  }
}

// ------------------------------------------------------------------------------------------------
// K1: feasibility. One warp per pod row; lane l owns column words l, l+32, ...
// ------------------------------------------------------------------------------------------------
struct K1Params {
  DevCatalog cat;
  const uint64_t* rows;  // [n_pods][32] FFD order
  int n_pods;
  const uint32_t* itype_sets;  // [n][W32]
  uint32_t* F;                 // [n_pods][n_templates][W32]
  unsigned long long* best;    // [n_pods]
  int word_begin, word_end;    // column shard (u32 words) this device computes
  int alloc_in_smem;           // the sorted allocatable arrays fit in the CTA's shared memory
  long long* dbg;              // optional cycle counters (KSCHED_PROFILE_K1)
  int tables_in_smem;          // valset / absent / negempty / offset / anyoffer / member are staged too
  int n_valrows, n_offrows;
};

// Parameters in constant memory and the CTA's catalog view (table pointers redirected to the staged shared-memory
// copies) in shared memory: the out-of-line row evaluation reads both with immediate addresses instead of through
// references to a kernel parameter / a stack copy (see g_k2 below for what that costs).
__constant__ K1Params g_k1;
__shared__ DevCatalog g_k1cat;

// One row's feasibility against every (template, column): the rarely taken path of feasibility_kernel (rows that differ
// from their predecessor), kept out of line so that the batched streaming loop stays small.
__device__ __noinline__ unsigned long long k1_compute_row(int j, uint64_t word, uint32_t* cache, bool cacheable) {
  const DevCatalog& c = g_k1cat;
  const K1Params& p = g_k1;
  const int lane = threadIdx.x & 31;
  const int V = c.n_templates, W32 = c.W32, NK = c.n_keys, R = c.n_res, T = c.n_types;
  const int wpl = (W32 + 31) >> 5;
    const uint64_t meta = __shfl_sync(0xffffffffu, word, 24);
    const uint64_t tolerated = __shfl_sync(0xffffffffu, word, 25);
    const uint32_t pod_res_present = (uint32_t)__shfl_sync(0xffffffffu, word, 28);
    const uint32_t itype_req = (uint32_t)__shfl_sync(0xffffffffu, word, 29);
    unsigned long long best = kNoBest;
    for (int v = 0; v < V; ++v) {
      const ksched_template& tm = c.templates[v];
      uint32_t* out = p.F + ((size_t)j * V + v) * W32;
      bool ok = (tolerated >> tm.taintset) & 1;  // Taints.Tolerates
      // lanes 8..23 own one requirement key each: Compatible + merge (node.go:73-81 on a fresh node)
      uint64_t allowed = 0;
      bool neg = false, present = false, compat = true;
      const int k = lane - 8;
      if (k >= 0 && k < NK) {
        Req pod;
        pod.present = (meta >> (KSCHED_META_PRESENT_SHIFT + k)) & 1;
        pod.complement = (meta >> (KSCHED_META_COMPLEMENT_SHIFT + k)) & 1;
        pod.has_gt = pod.has_lt = false; pod.gt = pod.lt = 0;
        pod.values = word;
        Req node = ksched::req_load(tm.reqs, nullptr, k);
        KeyMeta km = key_meta(c, k);
        compat = ksched::key_compatible(node, pod, c.keys[k].well_known != 0, km);
        Req merged = ksched::key_add(node, pod, km);
        present = merged.present;
        if (present) {
          allowed = ksched::req_allowed(merged, c.keys[k].dict_mask, km);
          neg = ksched::req_op_negative(merged);
        }
      }
      ok = ok && (__ballot_sync(0xffffffffu, !compat) == 0);
      const uint32_t present_keys = (__ballot_sync(0xffffffffu, present) >> 8) & c.type_relevant;
      // resources: lanes 0..7 own one resource each (Fits, resources.go:138-145)
      int rank = 0;
      bool res_used = false;
      if (lane < R) {
        uint32_t pres = pod_res_present | tm.daemon_res_present;
        res_used = (pres >> lane) & 1;
        if (res_used) rank = fit_rank(c.alloc_sorted, T, lane, (int64_t)word + tm.daemon_requests[lane]);
      }
      const uint32_t res_mask = __ballot_sync(0xffffffffu, res_used);
      uint64_t zallowed = 0xFFFF, callowed = 0xF;
      bool zc_unconstrained = true;
      if (c.zone_key >= 0) {
        bool zp = __shfl_sync(0xffffffffu, (int)present, 8 + c.zone_key);
        uint64_t za = __shfl_sync(0xffffffffu, allowed, 8 + c.zone_key);
        if (zp) { zallowed = za; zc_unconstrained = false; }
      }
      if (c.ct_key >= 0) {
        bool cp = __shfl_sync(0xffffffffu, (int)present, 8 + c.ct_key);
        uint64_t ca = __shfl_sync(0xffffffffu, allowed, 8 + c.ct_key);
        if (cp) { callowed = ca; zc_unconstrained = false; }
      }
      bool any = false;
      int first_word = -1;
      uint32_t first_bits = 0;
      for (int wi = 0; wi < wpl; ++wi) {
        const int w = wi * 32 + lane;
        const bool mine = w < W32 && w >= p.word_begin && w < p.word_end;
        uint32_t s = 0;
        if (ok && mine) s = c.member[(size_t)v * W32 + w];
        uint32_t pk = present_keys;
        while (pk) {  // uniform loop: every lane walks the same keys
          int kk = __ffs(pk) - 1;
          pk &= pk - 1;
          uint64_t a = __shfl_sync(0xffffffffu, allowed, 8 + kk);
          bool ng = __shfl_sync(0xffffffffu, (int)neg, 8 + kk);
          if (s) s &= key_typeset_word(c, kk, a, ng, w);
        }
        if (s) s &= offer_word(c, (uint32_t)zallowed, (uint32_t)callowed, zc_unconstrained, w);
        uint32_t rm = res_mask, fit = 0xFFFFFFFFu;
        while (rm) {  // the loads do not depend on each other (nor on s): one memory round trip for all resources
          int r = __ffs(rm) - 1;
          rm &= rm - 1;
          int rk = __shfl_sync(0xffffffffu, rank, r);
          if (w < W32) fit &= c.fitset[((size_t)r * (T + 1) + rk) * W32 + w];
        }
        s &= fit;
        if (s && itype_req != KSCHED_NONE) s &= p.itype_sets[(size_t)itype_req * W32 + w];
        if (mine) out[w] = s;
        if (cacheable) cache[v * wpl + wi] = s;
        uint32_t nz = __ballot_sync(0xffffffffu, s != 0);
        if (nz && !any) {
          any = true;
          int src = __ffs(nz) - 1;
          first_word = wi * 32 + src;
          first_bits = __shfl_sync(0xffffffffu, s, src);
        }
      }
      if (any) {
        int t = first_word * 32 + __ffs(first_bits) - 1;  // columns are in price order: first set bit = cheapest
        unsigned long long key = ((unsigned long long)__float_as_uint(c.price32[t]) << 32) | ((unsigned long long)v << 24) | (unsigned)t;
        best = key < best ? key : best;
      }
    }
  return best;
}


constexpr int kK1Threads = 512;
constexpr int kK1Cache = 16;  // per-lane words of the previous row's result kept for identical consecutive rows

struct K1Params;

// Shared-memory staging: the small tables every row evaluation walks with DEPENDENT loads (templates, key info, the
// value->row maps, the sorted allocatable arrays of the Fits binary search) are copied once per CTA; the wide column
// bitsets (valset / fitset / member / offset) stay in global memory and are read one coalesced word per lane.
// Each warp owns a CONTIGUOUS chunk of the FFD-ordered pod-row matrix: consecutive rows are very often identical in
// every field that matters to feasibility (same deployment), and then the previous result is written out again.
__global__ void __launch_bounds__(kK1Threads, 1) feasibility_kernel() {
  extern __shared__ __align__(16) unsigned char k1_smem[];
#ifdef KSCHED_PROFILE_K1
  long long t_start = clock64(), t_stage = 0, t_compute = 0, n_compute = 0;
#endif
  const K1Params& p = g_k1;
  DevCatalog c = p.cat;
  const int V = c.n_templates, W32 = c.W32, NK = c.n_keys, R = c.n_res, T = c.n_types;
  {
    unsigned char* ptr = k1_smem;
    ksched_template* s_tmpl = reinterpret_cast<ksched_template*>(ptr); ptr += sizeof(ksched_template) * V;
    ksched_keyinfo* s_keys = reinterpret_cast<ksched_keyinfo*>(ptr); ptr += sizeof(ksched_keyinfo) * KSCHED_MAX_KEYS;
    int64_t* s_alloc = reinterpret_cast<int64_t*>(ptr); ptr += p.alloc_in_smem ? sizeof(int64_t) * R * T : 0;
    int16_t* s_valrow = reinterpret_cast<int16_t*>(ptr); ptr += sizeof(int16_t) * KSCHED_MAX_KEYS * 64;
    int16_t* s_offrow = reinterpret_cast<int16_t*>(ptr);
    // cp.async (LDGSTS): every 4-byte element of every table is requested before anything is waited on, so the whole
    // staging costs about one memory round trip instead of one per table.
    auto stage4 = [&](void* dst, const void* src, int n_words) {
      const uint32_t* g = reinterpret_cast<const uint32_t*>(src);
      int done = 0;
      if ((((uintptr_t)dst | (uintptr_t)src) & 15) == 0) {  // 16-byte copies for the aligned bulk (4x fewer requests)
        const int n16 = n_words >> 2;
        for (int i = threadIdx.x; i < n16; i += blockDim.x) {
          const unsigned d = (unsigned)__cvta_generic_to_shared(reinterpret_cast<uint32_t*>(dst) + 4 * i);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g + 4 * i));
        }
        done = n16 << 2;
      }
      for (int i = done + threadIdx.x; i < n_words; i += blockDim.x) {
        const unsigned d = (unsigned)__cvta_generic_to_shared(reinterpret_cast<uint32_t*>(dst) + i);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(g + i));
      }
    };
    stage4(s_tmpl, c.templates, (int)(sizeof(ksched_template) * V / 4));
    stage4(s_keys, c.keys, (int)(sizeof(ksched_keyinfo) * NK / 4));
    if (p.alloc_in_smem) stage4(s_alloc, c.alloc_sorted, R * T * 2);
    stage4(s_valrow, c.valrow, NK * 64 / 2);
    stage4(s_offrow, c.offrow, 32);
    if (p.tables_in_smem) {
      // the narrow column bitsets: one 4*W32-byte row per (key,value) / key / offering / template
      uint32_t* sp = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(s_offrow) + sizeof(int16_t) * 64);
      auto stage = [&](const uint32_t*& tbl, int rows) {
        stage4(sp, tbl, rows * W32);
        tbl = sp;
        sp += rows * W32;
      };
      stage(c.valset, p.n_valrows);
      stage(c.absent, NK);
      stage(c.negempty, NK);
      stage(c.offset, p.n_offrows);
      stage(c.anyoffer, 1);
      stage(c.member, V);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    c.templates = s_tmpl;
    c.keys = s_keys;
    if (p.alloc_in_smem) c.alloc_sorted = s_alloc;
    c.valrow = s_valrow;
    c.offrow = s_offrow;
    if (threadIdx.x == 0) g_k1cat = c;
  }
  __syncthreads();
#ifdef KSCHED_PROFILE_K1
  t_stage = clock64() - t_start;
#endif

  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int chunk = (p.n_pods + nwarps - 1) / nwarps;
  const int j0 = warp * chunk, j1 = min(p.n_pods, j0 + chunk);
  const int wpl = (W32 + 31) >> 5;  // column words per lane
  const bool cacheable = V * wpl <= kK1Cache;
  // fields of the row that feasibility depends on (requests, requirement masks, meta, tolerations, res_present, itype_req)
  const uint64_t cmp_mask = lane <= 25 ? ~0ull : ((lane == 28 || lane == 29) ? 0xFFFFFFFFull : 0ull);
  uint32_t cache[kK1Cache];
  unsigned long long best_prev = kNoBest;
  uint64_t prev_word = 0;
  bool have_prev = false;

  // rows are fetched kBatch at a time (kBatch independent 256-byte loads in flight per warp), then consumed in order
  constexpr int kBatch = 16;
  uint64_t wbuf[kBatch];
  for (int jb = j0; jb < j1; jb += kBatch) {
#pragma unroll
    for (int b = 0; b < kBatch; ++b) wbuf[b] = jb + b < j1 ? __ldg(p.rows + (size_t)(jb + b) * KSCHED_ROW_WORDS + lane) : 0;
    if (have_prev && cacheable) {
      // the usual case: the whole batch repeats the cached row (same deployment) -> one vote, then nothing but stores
      uint64_t diff = 0;
#pragma unroll
      for (int b = 0; b < kBatch; ++b) diff |= jb + b < j1 ? ((wbuf[b] ^ prev_word) & cmp_mask) : 0;
      if (__all_sync(0xffffffffu, diff == 0)) {
        const int nb = min(kBatch, j1 - jb);
        const size_t row_stride = (size_t)V * W32;
        for (int v = 0; v < V; ++v)
          for (int wi = 0; wi < wpl; ++wi) {
            const int w = wi * 32 + lane;
            if (w < W32 && w >= p.word_begin && w < p.word_end) {
              const uint32_t val = cache[v * wpl + wi];
              uint32_t* dst = p.F + ((size_t)jb * V + v) * W32 + w;
              for (int b = 0; b < nb; ++b, dst += row_stride) *dst = val;
            }
          }
        if (lane < nb) p.best[jb + lane] = best_prev;
        continue;
      }
    }
#pragma unroll
    for (int b = 0; b < kBatch; ++b) {
    const int j = jb + b;
    if (j >= j1) break;
    const uint64_t word = wbuf[b];  // lane l holds u64 word l of the 256-byte row
    const bool same = have_prev && cacheable && __all_sync(0xffffffffu, ((word ^ prev_word) & cmp_mask) == 0);
    if (same) {
      for (int v = 0; v < V; ++v)
        for (int wi = 0; wi < wpl; ++wi) {
          const int w = wi * 32 + lane;
          if (w < W32 && w >= p.word_begin && w < p.word_end) p.F[((size_t)j * V + v) * W32 + w] = cache[v * wpl + wi];
        }
      if (lane == 0) p.best[j] = best_prev;
      continue;
    }
    prev_word = word;
    have_prev = true;
#ifdef KSCHED_PROFILE_K1
    long long tc0 = clock64();
#endif
    const unsigned long long best = k1_compute_row(j, word, cache, cacheable);
#ifdef KSCHED_PROFILE_K1
    t_compute += clock64() - tc0; ++n_compute;
#endif
    best_prev = best;
    if (lane == 0) p.best[j] = best;
    }
  }
#ifdef KSCHED_PROFILE_K1
  if (p.dbg && lane == 0) {
    atomicMax((unsigned long long*)&p.dbg[0], (unsigned long long)t_stage);
    atomicMax((unsigned long long*)&p.dbg[1], (unsigned long long)t_compute);
    atomicMax((unsigned long long*)&p.dbg[2], (unsigned long long)(clock64() - t_start));
    atomicMax((unsigned long long*)&p.dbg[3], (unsigned long long)n_compute);
    atomicAdd((unsigned long long*)&p.dbg[4], (unsigned long long)n_compute);
    atomicAdd((unsigned long long*)&p.dbg[5], (unsigned long long)t_compute);
  }
#endif
}

// ------------------------------------------------------------------------------------------------
// K2: pack
// ------------------------------------------------------------------------------------------------
// One (class, group) relation with the group's immutable fields folded in (built at upload): everything
// Topology.Record and the per-step topology build need about the relation arrives with ONE load.
struct RelX {
  uint32_t group, flags;
  uint8_t key, type, has_filter, pad;
  int32_t max_skew;
  int32_t host_row;
  uint32_t pad2[3];
};
static_assert(sizeof(RelX) == 32, "RelX is loaded as two 16-byte vectors");

struct PackState {
  const RelX* relx;                  // [n_class_topo], parallel to class_topo
  // problem (read-only)
  const ksched_pod_row* classes;
  const ksched_topo_group* groups;
  const ksched_class_topo* class_topo;
  const ksched_reqset* filter_terms;
  const uint32_t* itype_sets;        // [n][W32]
  const uint8_t* itype_complement;
  const int32_t* hostname_reqs;      // [n][2]
  const uint32_t* order;             // FFD order: queue position -> pod
  const uint64_t* rows;              // [n_pods][32] dense pod-row matrix in FFD order (K0)
  uint32_t* pod_pos;                 // pod -> FFD position (row of F / best)
  int use_F;                         // F / best cover every column on this device (not column-sharded)
  const uint32_t* F;                 // [n_pods][V][W32] in FFD order (nullptr: compute fresh-node types dynamically)
  const unsigned long long* best;    // [n_pods] FFD order (after allreduce when sharded) or nullptr
  int n_pods, n_classes, n_existing, n_groups, max_new;
  long long min_req[KSCHED_MAX_RES]; // min over all classes of requests[r] (0 if some class lacks r)
  // mutable
  uint32_t* pod_class;               // [n_pods] current class
  int32_t* relax_level;              // [n_pods]
  int32_t* assign;                   // [n_pods]
  int32_t* place_seq;                // [n_pods]
  uint32_t* queue;                   // [n_pods+1] circular
  int32_t* last_len;                 // [n_pods]
  uint32_t* last_epoch;              // [n_pods]
  // existing nodes (SoA)
  int64_t* ex_req;                   // [8][n_existing]
  const int64_t* ex_avail;           // [8][n_existing]
  uint32_t* ex_req_present;
  const uint32_t* ex_avail_present;
  uint64_t* ex_vals;                 // [16][n_existing]
  uint64_t* ex_meta;
  const uint32_t* ex_taintset;
  const uint32_t* ex_itype;
  uint64_t* ex_hp;
  uint8_t* ex_closed;
  // new nodes (SoA, capacity max_new)
  uint8_t* nn_tmpl;
  int32_t* nn_count;
  int32_t* nn_tb;
  int64_t* nn_req;                   // [8][max_new]
  uint32_t* nn_req_present;
  uint64_t* nn_vals;                 // [16][max_new]
  uint64_t* nn_meta;
  uint32_t* nn_opts;                 // [W32][max_new]
  uint64_t* nn_hp;
  // hot state of open nodes beyond the shared-memory window (pack_kernel.cuh: Hot)
  unsigned long long* ov_key;
  long long* ov_q;
  long long* ov_bound;
  long long* ov_bound2;
  int* ov_node;
  unsigned short* ov_flags;
  unsigned* ov_absorbed;
  unsigned* ov_rejected;
  // fresh-node outcome memo per (class, template)
  uint8_t* fc_state;                 // 0 unknown, 1 cached, 2 cached: no surviving type
  uint32_t* fc_opts;                 // [n_classes*V][W32]
  long long* fc_bound;               // [n_classes*V][4]
  long long* fc_bound2;
  uint8_t* fc_dom;
  uint8_t* fc_front_state;           // fc_bound / fc_bound2 / fc_dom hold the front of the K1-row option set
  uint64_t* fc_vals;                 // [n_classes*V][16] requirement masks of the fresh node
  uint64_t* fc_meta;
  long long* fc_q;                   // [n_classes*V][8] requests (daemon overhead + pod)
  uint32_t* fc_qp;
  // fresh-node outcome memo for topology-constrained classes: a hash table keyed by (class, template, the new node's
  // requirement masks) -- the option set of NewNode+Add is a function of exactly those (node.go:62-107) when the
  // provisioner has no limits. Direct mapped, full tag compare, so a collision only costs a recomputation.
  int fd_cap;                        // power of two
  uint8_t* fd_state;                 // 0 empty, 1 options cached, 2 cached: no surviving type
  uint32_t* fd_fc;                   // tag: class * V + template
  uint64_t* fd_meta;                 // tag
  uint64_t* fd_vals;                 // tag [fd_cap][16]
  uint32_t* fd_opts;                 // [fd_cap][W32]
  long long* fd_bound;               // [fd_cap][4]
  long long* fd_bound2;
  uint8_t* fd_dom;
  int count_visited;                 // keep the exact nodes_visited statistic (costs a pass over all in-flight nodes per pod)
  int alloc_in_smem;
  // topology counters
  int32_t* grp_cnt;                  // [n_groups][64]
  uint64_t* grp_registered;          // [n_groups]
  uint16_t* grp_host;                // [n_hostgroups][n_existing+max_new]
  const int32_t* grp_host_row;       // [n_groups] row in grp_host or -1
  int32_t* grp_host_total;           // [n_groups] schedulable-slot domains with count > 0 (+ extra_nonzero_domains)
  uint8_t* grp_active;               // [n_groups] 0 while a relaxation-created group does not exist yet (topology.go:86-117)
  int32_t* grp_min_slot;             // [n_groups] hostname slots below this were never Register()ed with the group
  int64_t* remaining;                // [V][8] provisioner limits
  // outputs / counters: [0]=n_new [1]=n_unscheduled [2]=nodes_visited [3]=add_calls [4]=error [5]=steps
  long long* counters;
};

struct K2Params {
  DevCatalog cat;
  PackState st;
};
// The pack kernel and every out-of-line device function it calls read their parameters from constant memory with
// immediate offsets (a reference to a __grid_constant__ kernel parameter handed to a __noinline__ function degrades every
// field access to a generic load with ~100 cycles of latency, on a code path that is one dependent chain).
// One copy per device: run_pack() orders launches of different handles on the same device behind each other.
__constant__ K2Params g_k2;
#define KS_K2 const DevCatalog& c = g_k2.cat; const PackState& s = g_k2.st; (void)c; (void)s;

struct Touched {
  int n;
  int8_t key[kMaxTouched];
  Req merged[kMaxTouched];  // node ∩ pod          (node.go:81)
  Req fin[kMaxTouched];     // ... ∩ topology      (node.go:90)
  bool changed[kMaxTouched];
};

struct PodTopo {  // per (pod step, constraining group): node-independent part of TopologyGroup.Get
  int n;
  int32_t group[kMaxCG];
  uint32_t flags[kMaxCG];
  int32_t min_count[kMaxCG];   // spread: domainMinCount over the pod's domains
  uint64_t options[kMaxCG];    // affinity / anti-affinity: admissible domains (mask keys)
  uint8_t bootstrap[kMaxCG];   // affinity: no domain has a matching pod and the pod selects itself
  uint64_t pod_allowed[kMaxCG]; // pod's own admissible domains for the key
  // the group's own fields, copied once per step so that the per-node checks read shared memory only
  uint8_t gkey[kMaxCG], gtype[kMaxCG];
  int32_t gskew[kMaxCG], host_row[kMaxCG], min_slot[kMaxCG];
  uint64_t registered[kMaxCG];
  // spread over a mask key: the registered domains within max-skew, in (count, domain id) order. The domain the
  // reference picks for a node is the first entry the node's requirement admits (topologygroup.go:157-183).
  uint8_t n_sorted[kMaxCG];
  uint8_t sorted[kMaxCG][64];
  uint64_t ok_mask[kMaxCG];    // the same domains as a set
  int overflow;
};

// Per-CTA working set of the pack kernel's generic step, at file scope so that every out-of-line function reaches it with
// immediate shared-memory addresses instead of pointers handed down the call chain.
__shared__ ksched_pod_row g_row;  // the pod of the current step (copied from its FFD row / class row by warp 0)
__shared__ PodTopo g_pt;          // its topology constraints
#define KS_ROW const ksched_pod_row& row = g_row; (void)row;
#define KS_PT PodTopo& pt = g_pt; (void)pt;

__device__ __forceinline__ Req load_soa(const uint64_t* vals, uint64_t meta, int stride, int idx, int k) {
  Req r;
  r.present = (meta >> (KSCHED_META_PRESENT_SHIFT + k)) & 1;
  r.complement = (meta >> (KSCHED_META_COMPLEMENT_SHIFT + k)) & 1;
  r.has_gt = r.has_lt = false;
  r.gt = r.lt = 0;
  r.values = r.present ? vals[(size_t)k * stride + idx] : 0;
  return r;
}
__device__ __forceinline__ Req pod_req(const ksched_pod_row& row, int k) {
  Req r;
  r.present = (row.meta >> (KSCHED_META_PRESENT_SHIFT + k)) & 1;
  r.complement = (row.meta >> (KSCHED_META_COMPLEMENT_SHIFT + k)) & 1;
  r.has_gt = r.has_lt = false;
  r.gt = r.lt = 0;
  r.values = row.values[k];
  return r;
}
__device__ __forceinline__ bool req_equal(const Req& a, const Req& b) {
  return a.present == b.present && a.complement == b.complement && a.values == b.values;
}

// TopologyGroup.Get for a mask-key group (topologygroup.go:88-243). node_dom = the node's requirement for the
// key after the pod's own requirements were merged (topology.go:156-159). Returns false when Len()==0.
__device__ __noinline__ bool topo_domains_mask(int j, const Req& node_dom, uint64_t* out) {
  KS_K2
  KS_PT
  const int k = pt.gkey[j];
  KeyMeta km = key_meta(c, k);
  const uint64_t registered = pt.registered[j];
  const uint64_t node_allowed = node_dom.present ? ksched::req_allowed(node_dom, c.keys[k].dict_mask, km) : c.keys[k].dict_mask;
  if (pt.gtype[j] == 0) {  // nextDomainTopologySpread: min (count, domain id) among the admissible domains of the node
    const int ns = pt.n_sorted[j];
    for (int i = 0; i < ns; ++i) {
      const int d = pt.sorted[j][i];
      if ((node_allowed >> d) & 1) { *out = 1ull << d; return true; }
    }
    return false;
  }
  if (pt.gtype[j] == 1) {  // nextDomainAffinity
    uint64_t opts = pt.options[j];
    if (pt.bootstrap[j]) {
      uint64_t inter = registered & pt.pod_allowed[j] & node_allowed;  // podDomains.Intersection(nodeDomains).Has
      if (inter) opts |= inter & (~inter + 1);
      uint64_t pm = registered & pt.pod_allowed[j];
      if (pm) opts |= pm & (~pm + 1);
    }
    if (!opts) return false;
    *out = opts;
    return true;
  }
  uint64_t opts = pt.options[j];  // nextDomainAntiAffinity
  if (!opts) return false;
  *out = opts;
  return true;
}

// hostname-key groups: the node's hostname domain is its slot.
__device__ __noinline__ bool topo_hostname_ok(int j, int slot, bool pod_allows_slot) {
  KS_K2
  KS_PT
  const int stride = s.n_existing + s.max_new;
  const int32_t cnt = s.grp_host[(size_t)pt.host_row[j] * stride + slot];
  // a group created by a later Topology.Update only knows hostnames registered after that, plus those it counted pods on
  if (slot < pt.min_slot[j] && !(slot < s.n_existing && cnt > 0)) return false;
  const int type = pt.gtype[j];
  if (type == 0) {  // spread, min is 0 for hostname (topologygroup.go:186-188); candidate = the node's own hostname
    int64_t c2 = (int64_t)cnt + ((pt.flags[j] & KSCHED_TOPO_SELECTS) ? 1 : 0);
    return c2 <= (int64_t)pt.gskew[j];
  }
  if (type == 1) {  // affinity
    if (cnt > 0 && pod_allows_slot) return true;
    if (pt.bootstrap[j]) return pod_allows_slot;  // first loop picks the node's own (registered) hostname
    return false;
  }
  return cnt == 0 && pod_allows_slot;  // anti-affinity
}

// Does the pod's hostname requirement admit this slot? (requirement on kubernetes.io/hostname, never well-known)
__device__ __forceinline__ bool hostname_allows(const PackState& s, const ksched_pod_row& row, int slot, bool is_existing) {
  if (row.hostname_req == KSCHED_NONE) return true;
  const int32_t comp = s.hostname_reqs[row.hostname_req * 2], target = s.hostname_reqs[row.hostname_req * 2 + 1];
  const bool same = is_existing && target == slot;
  return comp ? !same : same;
}

// Requirement phase of Node.Add / ExistingNode.Add: Compatible(pod) + merge, topology tighten + Compatible + merge.
// vals/meta/stride/idx describe the node's requirement set. Returns false on reject.
__device__ __noinline__ bool requirements_phase(const uint64_t* vals, uint64_t meta, int stride, int idx, int slot, bool is_existing, Touched& t) {
  KS_K2
  KS_ROW KS_PT
  t.n = 0;
  if (!hostname_allows(s, row, slot, is_existing)) return false;
  uint32_t podkeys = (uint32_t)(row.meta >> KSCHED_META_PRESENT_SHIFT) & 0xFFFF;
  while (podkeys) {
    int k = __ffs(podkeys) - 1;
    podkeys &= podkeys - 1;
    Req node = load_soa(vals, meta, stride, idx, k);
    Req pod = pod_req(row, k);
    KeyMeta km = key_meta(c, k);
    if (!ksched::key_compatible(node, pod, c.keys[k].well_known != 0, km)) return false;
    if (t.n >= kMaxTouched) return false;
    Req merged = ksched::key_add(node, pod, km);
    t.key[t.n] = (int8_t)k;
    t.merged[t.n] = merged;
    t.fin[t.n] = merged;
    t.changed[t.n] = !req_equal(merged, node);
    ++t.n;
  }
  for (int j = 0; j < pt.n; ++j) {
    if (pt.gkey[j] == KSCHED_KEY_HOSTNAME) {
      if (!topo_hostname_ok(j, slot, hostname_allows(s, row, slot, is_existing))) return false;
      continue;
    }
    const int k = pt.gkey[j];
    int ti = -1;
    for (int i = 0; i < t.n; ++i) if (t.key[i] == k) ti = i;
    if (ti < 0) {
      if (t.n >= kMaxTouched) return false;
      ti = t.n++;
      Req node = load_soa(vals, meta, stride, idx, k);
      t.key[ti] = (int8_t)k;
      t.merged[ti] = node;
      t.fin[ti] = node;
      t.changed[ti] = false;
    }
    uint64_t dom;
    if (!topo_domains_mask(j, t.merged[ti], &dom)) return false;
    Req d{dom, 0, 0, true, false, false, false};
    KeyMeta km = key_meta(c, k);
    t.fin[ti] = ksched::key_add(t.fin[ti], d, km);  // requirements.Add(domains) topology.go:164
  }
  // nodeRequirements.Compatible(topologyRequirements) node.go:87 — only topology keys can differ
  for (int i = 0; i < t.n; ++i) {
    if (req_equal(t.fin[i], t.merged[i])) continue;
    const int k = t.key[i];
    if (!ksched::key_compatible(t.merged[i], t.fin[i], c.keys[k].well_known != 0, key_meta(c, k))) return false;
    t.fin[i] = ksched::key_add(t.merged[i], t.fin[i], key_meta(c, k));
    t.changed[i] = true;
  }
  return true;
}

// Surviving instance types of a node for one word (filterInstanceTypesByRequirements node.go:137-141):
// previous options ∧ Fits ∧ (keys whose requirement changed) ∧ hasOffering (if zone / capacity-type changed).
struct TypeCtx {
  int rank[KSCHED_MAX_RES];
  uint32_t res_mask;
  int nkeys;
  int8_t key[KSCHED_MAX_KEYS];
  uint64_t allowed[KSCHED_MAX_KEYS];
  bool neg[KSCHED_MAX_KEYS];
  bool offer_needed, offer_unconstrained;
  uint32_t zmask, cmask;
  uint32_t itype_req;
};
__device__ __noinline__ void build_type_ctx(const Touched& t, const long long* q,
                               uint32_t q_present, const uint64_t* vals, uint64_t meta, int stride, int idx, bool fresh,
                               const int64_t* alloc_sorted, TypeCtx& x, bool with_ranks = true) {
  KS_K2
  KS_ROW
  x.res_mask = q_present;
  if (with_ranks)
    for (int r = 0; r < c.n_res; ++r) x.rank[r] = ((q_present >> r) & 1) ? fit_rank(alloc_sorted, c.n_types, r, q[r]) : 0;
  x.nkeys = 0;
  x.offer_needed = fresh;
  auto add_key = [&](int k, const Req& f) {
    if (!((c.type_relevant >> k) & 1) || !f.present) return;
    KeyMeta km = key_meta(c, k);
    x.key[x.nkeys] = (int8_t)k;
    x.allowed[x.nkeys] = ksched::req_allowed(f, c.keys[k].dict_mask, km);
    x.neg[x.nkeys] = ksched::req_op_negative(f);
    ++x.nkeys;
  };
  if (fresh) {
    // every key of the new node's requirement set is evaluated from scratch
    uint32_t done = 0;
    for (int i = 0; i < t.n; ++i) { add_key(t.key[i], t.fin[i]); done |= 1u << t.key[i]; }
    uint32_t rest = ((uint32_t)(meta >> KSCHED_META_PRESENT_SHIFT) & 0xFFFF) & ~done;
    while (rest) {
      int k = __ffs(rest) - 1;
      rest &= rest - 1;
      add_key(k, load_soa(vals, meta, stride, idx, k));
    }
  } else {
    for (int i = 0; i < t.n; ++i) {
      if (!t.changed[i]) continue;
      add_key(t.key[i], t.fin[i]);
      if (t.key[i] == c.zone_key || t.key[i] == c.ct_key) x.offer_needed = true;
    }
  }
  x.zmask = 0xFFFF; x.cmask = 0xF; x.offer_unconstrained = true;
  if (x.offer_needed) {
    auto final_req = [&](int k) -> Req {
      for (int i = 0; i < t.n; ++i) if (t.key[i] == k) return t.fin[i];
      return load_soa(vals, meta, stride, idx, k);
    };
    if (c.zone_key >= 0) {
      Req z = final_req(c.zone_key);
      if (z.present) { x.zmask = (uint32_t)ksched::req_allowed(z, c.keys[c.zone_key].dict_mask, key_meta(c, c.zone_key)); x.offer_unconstrained = false; }
    }
    if (c.ct_key >= 0) {
      Req ct = final_req(c.ct_key);
      if (ct.present) { x.cmask = (uint32_t)ksched::req_allowed(ct, c.keys[c.ct_key].dict_mask, key_meta(c, c.ct_key)); x.offer_unconstrained = false; }
    }
  }
  x.itype_req = row.itype_req;
}
// resource part (Fits) and requirement part (keys / offerings / instance-type requirement) of the per-word filter
__device__ __forceinline__ uint32_t type_word_res(const DevCatalog& c, const TypeCtx& x, uint32_t sw, int w) {
  uint32_t rm = x.res_mask;
  while (sw && rm) {
    int r = __ffs(rm) - 1;
    rm &= rm - 1;
    sw &= c.fitset[((size_t)r * (c.n_types + 1) + x.rank[r]) * c.W32 + w];
  }
  return sw;
}
__device__ __forceinline__ uint32_t type_word_keys(const DevCatalog& c, const PackState& s, const TypeCtx& x, uint32_t sw, int w) {
  for (int i = 0; i < x.nkeys && sw; ++i) sw &= key_typeset_word(c, x.key[i], x.allowed[i], x.neg[i], w);
  if (sw && x.offer_needed) sw &= offer_word(c, x.zmask, x.cmask, x.offer_unconstrained, w);
  if (sw && x.itype_req != KSCHED_NONE) sw &= s.itype_sets[(size_t)x.itype_req * c.W32 + w];
  return sw;
}
__device__ __forceinline__ uint32_t type_word(const DevCatalog& c, const PackState& s, const TypeCtx& x, uint32_t base, int w) {
  return type_word_keys(c, s, x, type_word_res(c, x, base, w), w);
}

// Node-independent part of the pod's topology constraints for this step, one constraining (class, group) relation j.
// Static half: what the relation and its group are (a function of the pod class only).
__device__ __forceinline__ void fill_pod_topo_static(const RelX& x, PodTopo& pt, int j) {
  pt.group[j] = (int)x.group;
  pt.flags[j] = x.flags;
  pt.gkey[j] = x.key;
  pt.gtype[j] = x.type;
  pt.gskew[j] = x.max_skew;
  pt.host_row[j] = x.host_row;
}
// Dynamic half: everything derived from the group's counters, re-read every step.
__device__ __noinline__ void fill_pod_topo_dynamic(int j) {
  KS_K2
  KS_ROW KS_PT
  const int gi = pt.group[j];
  const uint32_t flags = pt.flags[j];
  const int gkey = pt.gkey[j], gtype = pt.gtype[j], gskew = pt.gskew[j];
  pt.min_count[j] = 0;
  pt.options[j] = 0;
  pt.bootstrap[j] = 0;
  pt.pod_allowed[j] = 0;
  pt.min_slot[j] = s.grp_min_slot[gi];
  pt.registered[j] = 0;
  pt.n_sorted[j] = 0;
  pt.ok_mask[j] = 0;
  if (gkey == KSCHED_KEY_HOSTNAME) {
    if (gtype == 1) {
      // options.Len()==0 <=> no admissible hostname has a matching pod (hostname requirements on the pod are
      // restricted to a single existing slot, handled in topo_hostname_ok)
      pt.bootstrap[j] = (s.grp_host_total[gi] == 0) && (flags & KSCHED_TOPO_SELECTS);
    }
    return;
  }
  const int k = gkey;
  KeyMeta km = key_meta(c, k);
  Req pd = pod_req(row, k);
  const uint64_t pod_allowed = pd.present ? ksched::req_allowed(pd, c.keys[k].dict_mask, km) : c.keys[k].dict_mask;
  pt.pod_allowed[j] = pod_allowed;
  const uint64_t registered = s.grp_registered[gi];
  pt.registered[j] = registered;
  uint64_t m = registered & pod_allowed;
  if (gtype == 0) {
    // counts of every registered domain, loaded once (independent loads), ids in ascending order
    int32_t cnts[64];
    uint8_t ids[64];
    int nd = 0;
    for (uint64_t all = registered; all; all &= all - 1) {
      const int d = __ffsll((long long)all) - 1;
      ids[nd] = (uint8_t)d;
      cnts[nd] = s.grp_cnt[(size_t)gi * 64 + d];
      ++nd;
    }
    int32_t mn = INT32_MAX;  // domainMinCount over the pod's own domains (topologygroup.go:186-203)
    for (int i = 0; i < nd; ++i)
      if (((m >> ids[i]) & 1) && cnts[i] < mn) mn = cnts[i];
    pt.min_count[j] = mn;
    // registered domains within the skew bound, insertion-sorted by (count, id)
    const int self = (flags & KSCHED_TOPO_SELECTS) ? 1 : 0;
    int ns = 0;
    int32_t sc[64];
    uint64_t okm = 0;
    for (int i = 0; i < nd; ++i) {
      const int64_t cnt = (int64_t)cnts[i] + self;
      if (cnt - (int64_t)mn > (int64_t)gskew) continue;
      okm |= 1ull << ids[i];
      int q = ns++;
      while (q > 0 && sc[q - 1] > (int32_t)cnt) { sc[q] = sc[q - 1]; pt.sorted[j][q] = pt.sorted[j][q - 1]; --q; }
      sc[q] = (int32_t)cnt;
      pt.sorted[j][q] = ids[i];
    }
    pt.n_sorted[j] = (uint8_t)ns;
    pt.ok_mask[j] = okm;
  } else if (gtype == 1) {
    uint64_t opts = 0;
    while (m) {
      int d = __ffsll((long long)m) - 1;
      m &= m - 1;
      if (s.grp_cnt[(size_t)gi * 64 + d] > 0) opts |= 1ull << d;
    }
    pt.options[j] = opts;
    pt.bootstrap[j] = (opts == 0) && (flags & KSCHED_TOPO_SELECTS);
  } else {
    uint64_t opts = 0;
    while (m) {
      int d = __ffsll((long long)m) - 1;
      m &= m - 1;
      if (s.grp_cnt[(size_t)gi * 64 + d] == 0) opts |= 1ull << d;
    }
    pt.options[j] = opts;
  }
}

// Called by the 32 lanes of warp 0: one (class, group) relation per lane, constraining ones compacted in order.
__device__ void build_pod_topo() {
  KS_K2
  KS_ROW KS_PT
  const int lane = threadIdx.x & 31;
  const uint32_t begin = row.topo_begin, end = row.topo_end;
  int n = 0, overflow = 0;
  for (uint32_t base = begin; base < end; base += 32) {  // uniform trip count
    const uint32_t e = base + lane;
    RelX x{};
    if (e < end) x = s.relx[e];
    const bool cons = e < end && (x.flags & KSCHED_TOPO_CONSTRAINS);
    const unsigned bal = __ballot_sync(0xffffffffu, cons);
    const int j = n + __popc(bal & ((1u << lane) - 1));
    if (cons && j < kMaxCG) {
      fill_pod_topo_static(x, pt, j);
      fill_pod_topo_dynamic(j);
    }
    n += __popc(bal);
    if (n > kMaxCG) { overflow = 1; n = kMaxCG; }
  }
  if (lane == 0) { pt.n = n; pt.overflow = overflow; }
}
// The previous step's pod had the same class: the relations are the same, only the counters moved.
__device__ void refresh_pod_topo() {
  KS_K2
  KS_ROW KS_PT
  const int lane = threadIdx.x & 31;
  if (lane < pt.n) fill_pod_topo_dynamic(lane);
}

// A NECESSARY condition of requirements_phase for an in-flight node, cheap enough to run on every candidate: the
// hostname groups exactly, spread groups over mask keys through the set of admissible domains. The node that wins the
// argmin is then checked in full (and excluded if it fails).
__device__ __forceinline__ bool topo_prefilter(const uint64_t* vals, uint64_t meta,
                                               int stride, int idx, int slot) {
  KS_K2
  KS_PT
  for (int j = 0; j < pt.n; ++j) {
    const int k = pt.gkey[j];
    if (k == KSCHED_KEY_HOSTNAME) {
      if (!topo_hostname_ok(j, slot, true)) return false;
      continue;
    }
    if (pt.gtype[j] != 0) {
      if (!pt.options[j] && !pt.bootstrap[j]) return false;
      continue;
    }
    const Req node = load_soa(vals, meta, stride, idx, k);
    const uint64_t node_allowed = node.present ? ksched::req_allowed(node, c.keys[k].dict_mask, key_meta(c, k)) : c.keys[k].dict_mask;
    if (!(node_allowed & pt.ok_mask[j])) return false;
  }
  return true;
}

// TopologyNodeFilter.MatchesRequirements (topologynodefilter.go:57-70): any term Compatible with the node requirements
__device__ __noinline__ bool filter_matches(const ksched_topo_group& g, const uint64_t* vals, uint64_t meta,
                               int stride, int idx) {
  KS_K2
  if (g.filter_begin == g.filter_end) return true;
  for (uint32_t f = g.filter_begin; f < g.filter_end; ++f) {
    const ksched_reqset& term = s.filter_terms[f];
    bool ok = true;
    uint32_t keys = (uint32_t)(term.meta >> KSCHED_META_PRESENT_SHIFT) & 0xFFFF;
    while (keys && ok) {
      int k = __ffs(keys) - 1;
      keys &= keys - 1;
      Req node = load_soa(vals, meta, stride, idx, k);
      Req inc = ksched::req_load(term, nullptr, k);
      ok = ksched::key_compatible(node, inc, c.keys[k].well_known != 0, key_meta(c, k));
    }
    if (ok) return true;
  }
  return false;
}

// Topology.Record (topology.go:120-143) for ONE (class, group) relation, after the node's requirements were committed.
// Relations of one class name distinct groups, so different threads may record different relations concurrently.
__device__ __noinline__ void topo_record_entry(uint32_t e, const uint64_t* vals, uint64_t meta, int stride, int idx, int slot) {
  KS_K2
  const int hstride = s.n_existing + s.max_new;
  const RelX x = s.relx[e];
  const int gi = (int)x.group;
  if (!(x.flags & (KSCHED_TOPO_RECORDS | KSCHED_TOPO_RECORDS_INVERSE))) return;
  // the loads the record needs are issued together (one memory round trip, not three)
  const uint8_t active = s.grp_active[gi];
  uint16_t* const host_cell = x.key == KSCHED_KEY_HOSTNAME ? &s.grp_host[(size_t)x.host_row * hstride + slot] : nullptr;
  const uint32_t host_old = host_cell ? *host_cell : 0;
  const int32_t host_total = host_cell ? s.grp_host_total[gi] : 0;
  if (!active) return;  // the group does not exist yet
  bool rec = false, all_values = false;
  if (x.flags & KSCHED_TOPO_RECORDS) {
    if (!x.has_filter || filter_matches(s.groups[gi], vals, meta, stride, idx)) { rec = true; all_values = (x.type == 2); }
  }
  const bool rec_inv = (x.flags & KSCHED_TOPO_RECORDS_INVERSE) != 0;
  if (x.key == KSCHED_KEY_HOSTNAME) {  // the node's hostname requirement is always In [its own hostname]
    const int times = (rec ? 1 : 0) + (rec_inv ? 1 : 0);
    if (!times) return;
    if (host_old == 0) s.grp_host_total[gi] = host_total + 1;
    const uint32_t now = host_old + times;
    *host_cell = (uint16_t)(now > 0xFFFF ? 0xFFFF : now);
    return;
  }
  for (int pass = 0; pass < 2; ++pass) {
    const bool doit = pass == 0 ? rec : rec_inv;
    const bool allv = pass == 0 ? all_values : true;
    if (!doit) continue;
    Req r = load_soa(vals, meta, stride, idx, x.key);
    uint64_t v = 0;
    if (allv) v = r.present ? r.values : 0;  // domains.Values(): members, or the excluded set of a complement
    else if (r.present && ksched::req_len_one(r)) v = r.values;
    while (v) {
      int d = __ffsll((long long)v) - 1;
      v &= v - 1;
      s.grp_cnt[(size_t)gi * 64 + d]++;
      s.grp_registered[gi] |= 1ull << d;
    }
  }
}
// every thread of the CTA: relation tid, tid + blockDim, ... (the commit this records must be visible: call after a barrier)
__device__ __forceinline__ void topo_record_block(const uint64_t* vals, uint64_t meta,
                                                  int stride, int idx, int slot) {
  KS_K2
  KS_ROW
  for (uint32_t e = row.topo_begin + threadIdx.x; e < row.topo_end; e += blockDim.x) topo_record_entry(e, vals, meta, stride, idx, slot);
}

}  // namespace

#include "pack_kernel.cuh"

namespace {

// Launch choice per finished new node (launch.cuh): one warp per node, lanes over the words of its final option set.
__global__ void launch_choice_kernel(DevCatalog c, const long long* counters, const uint64_t* nn_vals, const uint64_t* nn_meta,
                                     const uint32_t* nn_opts, int max_new, ksched_launch_choice* out) {
  const int n_new = (int)counters[0];
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int n = warp; n < n_new; n += nwarps) {
    // zones / capacity types the node's requirements admit (Offerings.Requirements, types.go:120-126)
    uint32_t zmask = 0xFFFF, cmask = 0xF;
    const uint64_t meta = nn_meta[n];
    if (c.zone_key >= 0) {
      const Req z = load_soa(nn_vals, meta, max_new, n, c.zone_key);
      if (z.present) zmask = (uint32_t)ksched::req_allowed(z, c.keys[c.zone_key].dict_mask, key_meta(c, c.zone_key));
    }
    if (c.ct_key >= 0) {
      const Req ct = load_soa(nn_vals, meta, max_new, n, c.ct_key);
      if (ct.present) cmask = (uint32_t)ksched::req_allowed(ct, c.keys[c.ct_key].dict_mask, key_meta(c, c.ct_key));
    }
    unsigned long long best = ~0ull, best_off = ksched::kNoOffering;
    int best_col = -1;
    for (int w = lane; w < c.W32; w += 32) {
      uint32_t m = nn_opts[(size_t)w * max_new + n];
      while (m) {
        const int t = w * 32 + __ffs(m) - 1;
        m &= m - 1;
        const uint64_t ok = ksched::offering_min_key(c.offer_keys + (size_t)t * 64, zmask, cmask);
        if (ok == ksched::kNoOffering) continue;
        const unsigned long long key = ksched::option_key(ok, c.input_index[t]);
        if (key < best) { best = key; best_off = ok; best_col = t; }
      }
    }
    const unsigned long long wbest = warp_min_u64(best);
    const unsigned owner = __ballot_sync(0xffffffffu, best == wbest && best != ~0ull);  // keys are unique per option
    if (owner) {
      const int src = __ffs(owner) - 1;
      best_col = __shfl_sync(0xffffffffu, best_col, src);
      best_off = __shfl_sync(0xffffffffu, best_off, src);
    } else {
      best_col = -1;
    }
    if (lane == 0) {
      ksched_launch_choice o;
      o.type_column = best_col;
      o.offering_slot = best_col >= 0 ? (int32_t)(best_off & 0xFF) : -1;
      o.price_rank = best_col >= 0 ? (uint32_t)(best_off >> 16) : 0;
      o.offering_index = best_col >= 0 ? (uint32_t)((best_off >> 8) & 0xFF) : 0;
      out[n] = o;
    }
  }
}

// L2 flush helper: write a buffer larger than L2 between timed iterations
__global__ void flush_kernel(uint32_t* buf, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) buf[i] = (uint32_t)i;
}

template <class T>
struct DevBuf {
  T* ptr = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap && ptr) return cudaSuccess;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    size_t want = n ? n : 1;
    cudaError_t e = cudaMalloc(&ptr, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
};

}  // namespace

struct ksched_handle {
  K1Params k1_host;  // staging copies of the kernels' constant-memory parameters (must outlive the async copies)
  K2Params k2_host;  // staging copy of the pack kernel's constant-memory parameters (must outlive the async copy)
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  // catalog
  bool have_catalog = false;
  DevCatalog cat{};
  int W64 = 0;
  DevBuf<ksched_keyinfo> d_keys;
  DevBuf<int64_t> d_key_int, d_capacity, d_alloc_sorted;
  DevBuf<ksched_template> d_templates;
  DevBuf<ksched_type_row> d_types;
  DevBuf<float> d_price32;
  DevBuf<int16_t> d_valrow, d_offrow;
  DevBuf<uint32_t> d_valset, d_absent, d_negempty, d_offset, d_anyoffer, d_member, d_fitset, d_domset, d_input_index;
  DevBuf<uint64_t> d_offer_keys;
  DevBuf<ksched_launch_choice> d_launch;
  std::vector<ksched_template> h_templates;
  int n_valrows = 1, n_offrows = 1;
  // problem
  bool uploaded = false;
  bool sorted = false;  // the FFD-ordered pod-row matrix of the uploaded problem exists
  int n_pods = 0, n_classes = 0, n_existing = 0, n_groups = 0, max_new = 0, n_hostgroups = 0;
  DevBuf<ksched_pod_row> d_classes;
  DevBuf<uint32_t> d_pod_pos, d_pod_class0, d_pod_class, d_uid_rank, d_order, d_idx_tmp, d_itype_sets, d_queue, d_last_epoch;
  DevBuf<int64_t> d_ts;
  DevBuf<uint64_t> d_k_cpu, d_k_mem, d_k_tie, d_k_tmp, d_rows;
  DevBuf<uint8_t> d_cub_tmp, d_itype_comp, d_ex_closed, d_nn_tmpl;
  DevBuf<ksched_topo_group> d_groups;
  DevBuf<ksched_class_topo> d_class_topo;
  DevBuf<RelX> d_relx;
  DevBuf<ksched_reqset> d_filter_terms;
  DevBuf<int32_t> d_hostname_reqs, d_relax, d_assign, d_place_seq, d_last_len, d_nn_count, d_nn_tb, d_ov_node, d_perm_desc, d_grp_cnt,
      d_grp_cnt0, d_grp_host_row, d_grp_host_total, d_grp_host_total0;
  DevBuf<uint32_t> d_F;
  DevBuf<unsigned long long> d_best;
  DevBuf<int64_t> d_ex_req, d_ex_req0, d_ex_avail, d_nn_req, d_remaining, d_alloc_rt;
  DevBuf<long long> d_ov_q, d_ov_bound, d_ov_bound2, d_fc_bound, d_fc_bound2;
  DevBuf<unsigned long long> d_ov_key;
  DevBuf<unsigned short> d_ov_flags;
  DevBuf<uint8_t> d_fc_state, d_fc_dom, d_fc_front_state, d_grp_active, d_grp_active0;
  DevBuf<int32_t> d_grp_min_slot;
  DevBuf<uint32_t> d_fc_opts, d_fc_qp, d_ov_absorbed, d_ov_rejected;
  DevBuf<uint64_t> d_fc_vals, d_fc_meta;
  DevBuf<long long> d_fc_q;
  DevBuf<uint8_t> d_fd_state, d_fd_dom;
  DevBuf<uint32_t> d_fd_fc, d_fd_opts;
  DevBuf<uint64_t> d_fd_meta, d_fd_vals;
  DevBuf<long long> d_fd_bound, d_fd_bound2;
  int count_visited = 1;
  DevBuf<uint32_t> d_ex_req_present, d_ex_req_present0, d_ex_avail_present, d_ex_taintset, d_ex_itype, d_nn_req_present, d_nn_opts;
  DevBuf<uint64_t> d_ex_vals, d_ex_vals0, d_ex_meta, d_ex_meta0, d_ex_hp, d_ex_hp0, d_nn_vals, d_nn_meta, d_nn_hp, d_grp_registered,
      d_grp_registered0;
  DevBuf<uint16_t> d_grp_host, d_grp_host0;
  DevBuf<long long> d_counters, d_k1dbg;
  DevBuf<uint32_t> d_flush;
  size_t cub_tmp_bytes = 0;
  int64_t min_req[KSCHED_MAX_RES] = {0};
  // sharding / nccl
  int rank = 0, world = 1;
  ncclComm_t comm = nullptr;
  // timings
  ksched_timings tm{};
  cudaEvent_t ev[8] = {nullptr};
};

template <class T>
static cudaError_t upload(ksched_handle* h, DevBuf<T>& buf, const T* src, size_t n) {
  cudaError_t e = buf.ensure(n);
  if (e != cudaSuccess) return e;
  if (n == 0) return cudaSuccess;
  h->tm.h2d_bytes += (int64_t)(n * sizeof(T));
  return cudaMemcpyAsync(buf.ptr, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream);
}
template <class T>
static cudaError_t upload_vec(ksched_handle* h, DevBuf<T>& buf, const std::vector<T>& v) { return upload(h, buf, v.data(), v.size()); }

static int type_words64(int n_types) { int w = (n_types + 63) / 64; return w ? w : 1; }

extern "C" {

int ksched_abi_version(void) { return KSCHED_ABI_VERSION; }
int ksched_type_words(int n_types) { return type_words64(n_types); }

int ksched_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return KSCHED_ERR_NO_DEVICE;
  return n;
}

int ksched_create(int device_ordinal, ksched_handle** out) {
  if (!out) return KSCHED_ERR_INVALID;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device_ordinal >= n) return KSCHED_ERR_NO_DEVICE;
  ksched_handle* h = new ksched_handle();
  h->device = device_ordinal;
  if (cudaSetDevice(device_ordinal) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    return KSCHED_ERR_CUDA;
  }
  for (auto& e : h->ev) cudaEventCreate(&e);
  {
    // CUDA loads kernels lazily on first launch: pull ours in here, not inside the first Solve
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, pack_kernel);
    cudaFuncGetAttributes(&fa, feasibility_kernel);
    cudaFuncGetAttributes(&fa, finalize_options_kernel);
    cudaFuncGetAttributes(&fa, gather_rows_kernel);
    cudaFuncGetAttributes(&fa, sort_keys_kernel);
    cudaFuncGetAttributes(&fa, gather_u64_kernel);
  }
  *out = h;
  return KSCHED_OK;
}

void ksched_destroy(ksched_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->comm) ncclCommDestroy(h->comm);
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;  // every DevBuf frees its allocation
}

const char* ksched_last_error(const ksched_handle* h) { return h ? h->err.c_str() : "null handle"; }


int ksched_load_catalog(ksched_handle* h, const ksched_catalog* cat) {
  if (!h || !cat) return KSCHED_ERR_INVALID;
  if (cat->n_keys > KSCHED_MAX_KEYS || cat->n_res > KSCHED_MAX_RES || cat->n_templates > KSCHED_MAX_TEMPLATES || cat->n_templates < 1 ||
      cat->n_types < 0) { h->err = "catalog dimensions out of range"; return KSCHED_ERR_INVALID; }
  if (cat->type_bounds) { h->err = "instance types with Gt/Lt requirements are not supported"; return KSCHED_ERR_UNSUPPORTED; }
  // device code carries no Gt/Lt bounds at all (reqmask.cuh compiles them out): refuse any requirement set that has one
  for (int v = 0; v < cat->n_templates; ++v)
    if (cat->template_bounds || (cat->templates[v].reqs.meta >> KSCHED_META_HASGT_SHIFT)) {
      h->err = "provisioners with Gt/Lt requirements are not supported on the device path";
      return KSCHED_ERR_UNSUPPORTED;
    }
  for (int t = 0; t < cat->n_types; ++t)
    if (cat->types[t].meta >> KSCHED_META_HASGT_SHIFT) { h->err = "instance types with Gt/Lt requirements are not supported"; return KSCHED_ERR_UNSUPPORTED; }
  CUDA_TRY(h, cudaSetDevice(h->device));
  const int T = cat->n_types, NK = cat->n_keys, V = cat->n_templates, R = cat->n_res;
  const int W64 = type_words64(T), W32 = W64 * 2;
  h->W64 = W64;
  // ---- bit-sliced tables, built on the host (amortised: instance types change rarely)
  std::vector<int16_t> valrow((size_t)NK * 64, -1), offrow(64, -1);
  std::vector<uint32_t> valset, absent((size_t)std::max(NK, 1) * W32, 0), negempty((size_t)std::max(NK, 1) * W32, 0), offset, anyoffer(W32, 0),
      member((size_t)V * W32, 0);
  uint32_t type_relevant = 0;
  int zone_key = -1, ct_key = -1;
  for (int k = 0; k < NK; ++k) {
    if (cat->keys[k].is_zone) zone_key = k;
    if (cat->keys[k].is_capacity_type) ct_key = k;
  }
  std::vector<float> price32(std::max(T, 1), 0.f);
  for (int t = 0; t < T; ++t) {
    const ksched_type_row& row = cat->types[t];
    if (t > 0 && row.min_price < cat->types[t - 1].min_price) { h->err = "instance types must be in ascending price order"; return KSCHED_ERR_INVALID; }
    price32[t] = (float)row.min_price;
    const int w = t >> 5;
    const uint32_t bit = 1u << (t & 31);
    if ((row.meta >> KSCHED_META_COMPLEMENT_SHIFT) & 0xFFFF || (row.meta >> KSCHED_META_HASGT_SHIFT)) {
      h->err = "instance types with complement / bounded requirements are not supported";
      return KSCHED_ERR_UNSUPPORTED;
    }
    for (int k = 0; k < NK; ++k) {
      bool present = (row.meta >> (KSCHED_META_PRESENT_SHIFT + k)) & 1;
      if (!present) { absent[(size_t)k * W32 + w] |= bit; continue; }
      type_relevant |= 1u << k;
      uint64_t v = row.values[k];
      if (!v) { negempty[(size_t)k * W32 + w] |= bit; continue; }
      while (v) {
        int b = __builtin_ctzll(v);
        v &= v - 1;
        int16_t& r = valrow[(size_t)k * 64 + b];
        if (r < 0) { r = (int16_t)(valset.size() / W32); valset.resize(valset.size() + W32, 0); }
        valset[(size_t)r * W32 + w] |= bit;
      }
    }
    uint64_t o = row.offerings;
    if (o) anyoffer[w] |= bit;
    while (o) {
      int b = __builtin_ctzll(o);
      o &= o - 1;
      int16_t& r = offrow[b];
      if (r < 0) { r = (int16_t)(offset.size() / W32); offset.resize(offset.size() + W32, 0); }
      offset[(size_t)r * W32 + w] |= bit;
    }
    for (int v = 0; v < V; ++v) if ((row.template_members >> v) & 1) member[(size_t)v * W32 + w] |= bit;
  }
  if (valset.empty()) valset.resize(W32, 0);
  if (offset.empty()) offset.resize(W32, 0);
  std::vector<int32_t> perm_desc((size_t)std::max(R, 1) * std::max(T, 1), 0);
  std::vector<int64_t> alloc_rt((size_t)std::max(R, 1) * std::max(T, 1), 0);
  std::vector<int64_t> alloc_sorted((size_t)std::max(R, 1) * std::max(T, 1), 0);
  std::vector<uint32_t> fitset((size_t)std::max(R, 1) * (T + 1) * W32, 0);
  {
    std::vector<int> perm(T);
    for (int r = 0; r < R; ++r) {
      for (int t = 0; t < T; ++t) perm[t] = t;
      std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return cat->types[a].allocatable[r] < cat->types[b].allocatable[r]; });
      for (int i = 0; i < T; ++i) {
        alloc_sorted[(size_t)r * T + i] = cat->types[perm[i]].allocatable[r];
        perm_desc[(size_t)r * T + (T - 1 - i)] = perm[i];
        alloc_rt[(size_t)r * T + i] = cat->types[i].allocatable[r];
      }
      uint32_t* base = &fitset[(size_t)r * (T + 1) * W32];
      for (int i = T - 1; i >= 0; --i) {
        uint32_t* cur = base + (size_t)i * W32;
        std::memcpy(cur, base + (size_t)(i + 1) * W32, W32 * sizeof(uint32_t));
        cur[perm[i] >> 5] |= 1u << (perm[i] & 31);
      }
    }
  }
  CUDA_TRY(h, upload(h, h->d_keys, cat->keys, (size_t)NK));
  if (cat->key_int_values) CUDA_TRY(h, upload(h, h->d_key_int, cat->key_int_values, (size_t)NK * 64));
  CUDA_TRY(h, upload(h, h->d_templates, cat->templates, (size_t)V));
  CUDA_TRY(h, upload(h, h->d_types, cat->types, (size_t)T));
  CUDA_TRY(h, upload(h, h->d_capacity, cat->type_capacity, (size_t)T * KSCHED_MAX_RES));
  CUDA_TRY(h, upload_vec(h, h->d_price32, price32));
  CUDA_TRY(h, upload_vec(h, h->d_valrow, valrow));
  CUDA_TRY(h, upload_vec(h, h->d_offrow, offrow));
  CUDA_TRY(h, upload_vec(h, h->d_valset, valset));
  CUDA_TRY(h, upload_vec(h, h->d_absent, absent));
  CUDA_TRY(h, upload_vec(h, h->d_negempty, negempty));
  CUDA_TRY(h, upload_vec(h, h->d_offset, offset));
  CUDA_TRY(h, upload_vec(h, h->d_anyoffer, anyoffer));
  CUDA_TRY(h, upload_vec(h, h->d_member, member));
  CUDA_TRY(h, upload_vec(h, h->d_alloc_sorted, alloc_sorted));
  CUDA_TRY(h, upload_vec(h, h->d_perm_desc, perm_desc));
  CUDA_TRY(h, upload_vec(h, h->d_alloc_rt, alloc_rt));
  {
    // dominance bitsets over the first min(R,4) resources (pack kernel: Pareto front of a node's options)
    const int RH = std::min(R, 4);
    std::vector<uint32_t> domset((size_t)std::max(T, 1) * W32, 0);
    for (int a = 0; a < T; ++a)
      for (int b = 0; b < T; ++b) {
        bool dom = true;
        for (int r = 0; r < RH && dom; ++r) dom = cat->types[a].allocatable[r] >= cat->types[b].allocatable[r];
        if (dom) domset[(size_t)a * W32 + (b >> 5)] |= 1u << (b & 31);
      }
    CUDA_TRY(h, upload_vec(h, h->d_domset, domset));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  }
  CUDA_TRY(h, upload_vec(h, h->d_fitset, fitset));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->h_templates.assign(cat->templates, cat->templates + V);
  h->n_valrows = (int)(valset.size() / W32);
  h->n_offrows = (int)(offset.size() / W32);
  DevCatalog& c = h->cat;
  c.n_keys = NK; c.n_res = R; c.n_types = T; c.n_templates = V; c.W32 = W32;
  c.keys = h->d_keys.ptr;
  c.key_int_values = cat->key_int_values ? h->d_key_int.ptr : nullptr;
  c.templates = h->d_templates.ptr;
  c.types = h->d_types.ptr;
  c.capacity = h->d_capacity.ptr;
  c.price32 = h->d_price32.ptr;
  c.valrow = h->d_valrow.ptr; c.valset = h->d_valset.ptr; c.absent = h->d_absent.ptr; c.negempty = h->d_negempty.ptr;
  c.type_relevant = type_relevant;
  c.offrow = h->d_offrow.ptr; c.offset = h->d_offset.ptr; c.anyoffer = h->d_anyoffer.ptr; c.member = h->d_member.ptr;
  c.alloc_sorted = h->d_alloc_sorted.ptr; c.fitset = h->d_fitset.ptr;
  c.perm_desc = h->d_perm_desc.ptr; c.alloc_rt = h->d_alloc_rt.ptr; c.domset = h->d_domset.ptr;
  c.offer_keys = nullptr;
  c.input_index = nullptr;
  if (cat->offering_keys && T > 0) {
    std::vector<uint32_t> input_index(T);
    for (int t = 0; t < T; ++t) input_index[t] = cat->types[t].input_index;
    CUDA_TRY(h, upload(h, h->d_offer_keys, cat->offering_keys, (size_t)T * 64));
    CUDA_TRY(h, upload_vec(h, h->d_input_index, input_index));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));  // input_index is a stack vector
    c.offer_keys = h->d_offer_keys.ptr;
    c.input_index = h->d_input_index.ptr;
  }
  c.zone_key = zone_key; c.ct_key = ct_key;
  h->have_catalog = true;
  h->uploaded = false;
  return KSCHED_OK;
}

int ksched_set_shard(ksched_handle* h, int rank, int world) {
  if (!h || world < 1 || rank < 0 || rank >= world) return KSCHED_ERR_INVALID;
  h->rank = rank;
  h->world = world;
  return KSCHED_OK;
}

int ksched_shard_range(int n_words32, int rank, int world, int* begin, int* end) {
  if (n_words32 < 0 || world < 1 || rank < 0 || rank >= world || !begin || !end) return KSCHED_ERR_INVALID;
  const int per = (n_words32 + world - 1) / world;
  *begin = std::min(n_words32, rank * per);
  *end = std::min(n_words32, *begin + per);
  return KSCHED_OK;
}

int ksched_nccl_unique_id(void* out128) {
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return KSCHED_ERR_NCCL;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  std::memcpy(out128, &id, 128);
  return KSCHED_OK;
}
int ksched_nccl_init(ksched_handle* h, const void* id128, int rank, int world) {
  if (!h || !id128) return KSCHED_ERR_INVALID;
  if (cudaSetDevice(h->device) != cudaSuccess) return KSCHED_ERR_CUDA;
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  ncclResult_t r = ncclCommInitRank(&h->comm, world, id, rank);
  if (r != ncclSuccess) { h->err = std::string("ncclCommInitRank: ") + ncclGetErrorString(r); return KSCHED_ERR_NCCL; }
  h->rank = rank;
  h->world = world;
  return KSCHED_OK;
}

// ---- upload one problem (pods / nodes / topology) and keep pristine copies of everything the pack kernel mutates
int ksched_upload(ksched_handle* h, const ksched_problem* pb) {
  if (!h || !pb) return KSCHED_ERR_INVALID;
  if (!h->have_catalog) { h->err = "ksched_load_catalog must be called first"; return KSCHED_ERR_INVALID; }
  if (pb->class_bounds || pb->existing_bounds) { h->err = "Gt/Lt requirements are not carried on the device path yet"; return KSCHED_ERR_UNSUPPORTED; }
  for (int i = 0; i < pb->n_classes; ++i)
    if (pb->classes[i].meta >> KSCHED_META_HASGT_SHIFT) { h->err = "Gt/Lt requirements are not carried on the device path yet"; return KSCHED_ERR_UNSUPPORTED; }
  for (int i = 0; i < pb->n_existing; ++i)
    if (pb->existing[i].reqs.meta >> KSCHED_META_HASGT_SHIFT) { h->err = "Gt/Lt requirements are not carried on the device path yet"; return KSCHED_ERR_UNSUPPORTED; }
  for (int i = 0; i < pb->n_filter_terms; ++i)
    if (pb->filter_terms[i].meta >> KSCHED_META_HASGT_SHIFT) { h->err = "Gt/Lt requirements are not carried on the device path yet"; return KSCHED_ERR_UNSUPPORTED; }
  CUDA_TRY(h, cudaSetDevice(h->device));
  h->tm.h2d_bytes = 0;
  const DevCatalog& c = h->cat;
  const int P = pb->n_pods, NC = pb->n_classes, NE = pb->n_existing, NG = pb->n_groups, W32 = c.W32, V = c.n_templates;
  const int MAXN = std::max(1, std::min(pb->max_new_nodes > 0 ? pb->max_new_nodes : P, std::max(P, 1)));
  h->n_pods = P; h->n_classes = NC; h->n_existing = NE; h->n_groups = NG; h->max_new = MAXN;
  h->count_visited = pb->count_nodes_visited;
  for (int c2 = 0; c2 < NC; ++c2) {
    const ksched_pod_row& row = pb->classes[c2];
    if (row.relax_next != KSCHED_NONE && row.relax_next >= (uint32_t)NC) { h->err = "relax_next out of range"; return KSCHED_ERR_INVALID; }
    if (row.topo_end < row.topo_begin || row.topo_end > (uint32_t)pb->n_class_topo) { h->err = "class_topo range out of bounds"; return KSCHED_ERR_INVALID; }
    int ncg = 0, touched = __builtin_popcountll((row.meta >> KSCHED_META_PRESENT_SHIFT) & 0xFFFF);
    for (uint32_t e = row.topo_begin; e < row.topo_end; ++e) {
      if (pb->class_topo[e].group >= (uint32_t)NG) { h->err = "class_topo group out of range"; return KSCHED_ERR_INVALID; }
      if (pb->class_topo[e].flags & KSCHED_TOPO_CONSTRAINS) { ++ncg; if (pb->groups[pb->class_topo[e].group].key != KSCHED_KEY_HOSTNAME) ++touched; }
    }
    if (ncg > kMaxCG || touched > kMaxTouched) { h->err = "a pod class is constrained by more topology groups / label keys than the pack kernel carries"; return KSCHED_ERR_UNSUPPORTED; }
  }
  // min request per resource over all classes (node closing rule)
  for (int r = 0; r < KSCHED_MAX_RES; ++r) {
    int64_t mn = INT64_MAX;
    for (int c2 = 0; c2 < NC; ++c2) {
      const ksched_pod_row& row = pb->classes[c2];
      int64_t v = ((row.res_present >> r) & 1) ? row.requests[r] : 0;
      mn = std::min(mn, v);
    }
    h->min_req[r] = (NC == 0 || mn == INT64_MAX || mn < 0) ? 0 : mn;
  }
  CUDA_TRY(h, upload(h, h->d_classes, pb->classes, (size_t)NC));
  CUDA_TRY(h, upload(h, h->d_pod_class0, pb->pod_class, (size_t)P));
  CUDA_TRY(h, h->d_pod_class.ensure(P));
  CUDA_TRY(h, upload(h, h->d_ts, pb->pod_timestamp, (size_t)P));
  CUDA_TRY(h, upload(h, h->d_uid_rank, pb->pod_uid_rank, (size_t)P));
  CUDA_TRY(h, upload(h, h->d_groups, pb->groups, (size_t)NG));
  CUDA_TRY(h, upload(h, h->d_class_topo, pb->class_topo, (size_t)pb->n_class_topo));
  CUDA_TRY(h, upload(h, h->d_filter_terms, pb->filter_terms, (size_t)pb->n_filter_terms));
  CUDA_TRY(h, upload(h, h->d_hostname_reqs, pb->hostname_reqs, (size_t)pb->n_hostname_reqs * 2));
  CUDA_TRY(h, upload(h, h->d_itype_comp, pb->itype_req_complement, (size_t)pb->n_itype_reqs));
  CUDA_TRY(h, upload(h, h->d_itype_sets, reinterpret_cast<const uint32_t*>(pb->itype_req_sets), (size_t)pb->n_itype_reqs * W32));
  // existing nodes -> SoA
  {
    std::vector<int64_t> req((size_t)8 * std::max(NE, 1), 0), avail((size_t)8 * std::max(NE, 1), 0);
    std::vector<uint32_t> reqp(std::max(NE, 1), 0), availp(std::max(NE, 1), 0), ts(std::max(NE, 1), 0), it(std::max(NE, 1), 0);
    std::vector<uint64_t> vals((size_t)16 * std::max(NE, 1), 0), meta(std::max(NE, 1), 0), hp(std::max(NE, 1), 0);
    for (int e = 0; e < NE; ++e) {
      const ksched_existing_node& n = pb->existing[e];
      for (int r = 0; r < 8; ++r) { req[(size_t)r * NE + e] = n.requests[r]; avail[(size_t)r * NE + e] = n.available[r]; }
      for (int k = 0; k < 16; ++k) vals[(size_t)k * NE + e] = n.reqs.values[k];
      meta[e] = n.reqs.meta & 0xFFFFFFFFull;
      reqp[e] = n.requests_present; availp[e] = n.available_present; ts[e] = n.taintset; it[e] = n.itype; hp[e] = n.hostport_entries;
    }
    CUDA_TRY(h, upload_vec(h, h->d_ex_req0, req));
    CUDA_TRY(h, upload_vec(h, h->d_ex_avail, avail));
    CUDA_TRY(h, upload_vec(h, h->d_ex_req_present0, reqp));
    CUDA_TRY(h, upload_vec(h, h->d_ex_avail_present, availp));
    CUDA_TRY(h, upload_vec(h, h->d_ex_taintset, ts));
    CUDA_TRY(h, upload_vec(h, h->d_ex_itype, it));
    CUDA_TRY(h, upload_vec(h, h->d_ex_vals0, vals));
    CUDA_TRY(h, upload_vec(h, h->d_ex_meta0, meta));
    CUDA_TRY(h, upload_vec(h, h->d_ex_hp0, hp));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));  // host vectors go out of scope
    size_t ne = (size_t)std::max(NE, 1);
    CUDA_TRY(h, h->d_ex_req.ensure(8 * ne));
    CUDA_TRY(h, h->d_ex_req_present.ensure(ne));
    CUDA_TRY(h, h->d_ex_vals.ensure(16 * ne));
    CUDA_TRY(h, h->d_ex_meta.ensure(ne));
    CUDA_TRY(h, h->d_ex_hp.ensure(ne));
    CUDA_TRY(h, h->d_ex_closed.ensure(ne));
  }
  // topology counters
  {
    std::vector<int32_t> host_row(std::max(NG, 1), -1), host_total(std::max(NG, 1), 0), cnt((size_t)std::max(NG, 1) * 64, 0);
    std::vector<uint64_t> registered(std::max(NG, 1), 0);
    int nh = 0;
    for (int g = 0; g < NG; ++g) {
      if (pb->groups[g].key == KSCHED_KEY_HOSTNAME) host_row[g] = nh++;
      else if (pb->groups[g].key >= c.n_keys) { h->err = "topology group key out of range"; return KSCHED_ERR_INVALID; }
      registered[g] = pb->groups[g].registered;
      for (int d = 0; d < 64; ++d) cnt[(size_t)g * 64 + d] = pb->group_domain_counts[(size_t)g * 64 + d];
    }
    h->n_hostgroups = nh;
    const size_t stride = (size_t)NE + MAXN;
    std::vector<uint16_t> host((size_t)std::max(nh, 1) * stride, 0);
    for (int g = 0; g < NG; ++g) {
      if (host_row[g] < 0) continue;
      int total = pb->groups[g].extra_nonzero_domains;
      for (int e = 0; e < NE; ++e) {
        int32_t v = pb->group_existing_counts[(size_t)g * std::max(NE, 1) + e];
        host[(size_t)host_row[g] * stride + e] = (uint16_t)std::min(v, 0xFFFF);
        if (v > 0) ++total;
      }
      host_total[g] = total;
    }
    std::vector<uint8_t> active(std::max(NG, 1), 1);
    for (int g = 0; g < NG; ++g) active[g] = pb->groups[g].dormant ? 0 : 1;
    CUDA_TRY(h, upload_vec(h, h->d_grp_active0, active));
    CUDA_TRY(h, h->d_grp_active.ensure(std::max(NG, 1)));
    CUDA_TRY(h, h->d_grp_min_slot.ensure(std::max(NG, 1)));
    CUDA_TRY(h, upload_vec(h, h->d_grp_host_row, host_row));
    {
      std::vector<RelX> relx((size_t)std::max(pb->n_class_topo, 1));
      for (int e = 0; e < pb->n_class_topo; ++e) {
        const ksched_class_topo& ct = pb->class_topo[e];
        if ((int)ct.group >= NG) { h->err = "class_topo names a group out of range"; return KSCHED_ERR_INVALID; }
        const ksched_topo_group& g = pb->groups[ct.group];
        RelX x{};
        x.group = ct.group; x.flags = ct.flags; x.key = g.key; x.type = g.type; x.has_filter = g.filter_begin != g.filter_end;
        x.max_skew = g.max_skew; x.host_row = host_row[ct.group];
        relx[e] = x;
      }
      CUDA_TRY(h, upload_vec(h, h->d_relx, relx));
    }
    CUDA_TRY(h, upload_vec(h, h->d_grp_host_total0, host_total));
    CUDA_TRY(h, upload_vec(h, h->d_grp_cnt0, cnt));
    CUDA_TRY(h, upload_vec(h, h->d_grp_registered0, registered));
    CUDA_TRY(h, upload_vec(h, h->d_grp_host0, host));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, h->d_grp_host_total.ensure(std::max(NG, 1)));
    CUDA_TRY(h, h->d_grp_cnt.ensure((size_t)std::max(NG, 1) * 64));
    CUDA_TRY(h, h->d_grp_registered.ensure(std::max(NG, 1)));
    CUDA_TRY(h, h->d_grp_host.ensure((size_t)std::max(nh, 1) * stride));
  }
  // scratch / state
  const size_t p1 = (size_t)std::max(P, 1);
  CUDA_TRY(h, h->d_order.ensure(p1)); CUDA_TRY(h, h->d_idx_tmp.ensure(p1)); CUDA_TRY(h, h->d_pod_pos.ensure(p1));
  CUDA_TRY(h, h->d_k_cpu.ensure(p1)); CUDA_TRY(h, h->d_k_mem.ensure(p1)); CUDA_TRY(h, h->d_k_tie.ensure(p1)); CUDA_TRY(h, h->d_k_tmp.ensure(p1));
  CUDA_TRY(h, h->d_rows.ensure(p1 * KSCHED_ROW_WORDS));
  CUDA_TRY(h, h->d_F.ensure(p1 * V * W32));
  CUDA_TRY(h, h->d_best.ensure(p1));
  CUDA_TRY(h, h->d_relax.ensure(p1)); CUDA_TRY(h, h->d_assign.ensure(p1)); CUDA_TRY(h, h->d_place_seq.ensure(p1));
  CUDA_TRY(h, h->d_queue.ensure(p1 + 1)); CUDA_TRY(h, h->d_last_len.ensure(p1)); CUDA_TRY(h, h->d_last_epoch.ensure(p1));
  const size_t mn = (size_t)MAXN;
  CUDA_TRY(h, h->d_nn_tmpl.ensure(mn)); CUDA_TRY(h, h->d_nn_count.ensure(mn)); CUDA_TRY(h, h->d_nn_tb.ensure(mn));
  CUDA_TRY(h, h->d_nn_req.ensure(8 * mn)); CUDA_TRY(h, h->d_nn_req_present.ensure(mn));
  CUDA_TRY(h, h->d_nn_vals.ensure(16 * mn)); CUDA_TRY(h, h->d_nn_meta.ensure(mn));
  CUDA_TRY(h, h->d_nn_opts.ensure((size_t)W32 * mn)); CUDA_TRY(h, h->d_nn_hp.ensure(mn));
  CUDA_TRY(h, h->d_ov_key.ensure(mn)); CUDA_TRY(h, h->d_ov_q.ensure(4 * mn)); CUDA_TRY(h, h->d_ov_bound.ensure(4 * mn)); CUDA_TRY(h, h->d_ov_bound2.ensure(4 * mn));
  CUDA_TRY(h, h->d_ov_node.ensure(mn)); CUDA_TRY(h, h->d_ov_flags.ensure(mn));
  CUDA_TRY(h, h->d_ov_absorbed.ensure(mn)); CUDA_TRY(h, h->d_ov_rejected.ensure(mn));
  {
    const size_t nfc = (size_t)std::max(NC, 1) * V;
    CUDA_TRY(h, h->d_fc_state.ensure(nfc)); CUDA_TRY(h, h->d_fc_dom.ensure(nfc)); CUDA_TRY(h, h->d_fc_front_state.ensure(nfc)); CUDA_TRY(h, h->d_fc_bound.ensure(nfc * 4)); CUDA_TRY(h, h->d_fc_bound2.ensure(nfc * 4));
    CUDA_TRY(h, h->d_fc_opts.ensure(nfc * W32));
    CUDA_TRY(h, h->d_fc_vals.ensure(nfc * KSCHED_MAX_KEYS)); CUDA_TRY(h, h->d_fc_meta.ensure(nfc)); CUDA_TRY(h, h->d_fc_q.ensure(nfc * KSCHED_MAX_RES));
    CUDA_TRY(h, h->d_fc_qp.ensure(nfc));
    const size_t nfd = kFreshMemoSlots;
    CUDA_TRY(h, h->d_fd_state.ensure(nfd)); CUDA_TRY(h, h->d_fd_dom.ensure(nfd)); CUDA_TRY(h, h->d_fd_fc.ensure(nfd)); CUDA_TRY(h, h->d_fd_meta.ensure(nfd));
    CUDA_TRY(h, h->d_fd_vals.ensure(nfd * KSCHED_MAX_KEYS)); CUDA_TRY(h, h->d_fd_opts.ensure(nfd * W32));
    CUDA_TRY(h, h->d_fd_bound.ensure(nfd * 4)); CUDA_TRY(h, h->d_fd_bound2.ensure(nfd * 4));
  }
  CUDA_TRY(h, h->d_remaining.ensure((size_t)V * KSCHED_MAX_RES));
  CUDA_TRY(h, h->d_counters.ensure(32));
  {
    size_t need = 0, n2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, need, (uint64_t*)nullptr, (uint64_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)p1, 0, 64, h->stream);
    n2 = need;
    CUDA_TRY(h, h->d_cub_tmp.ensure(n2 + 256));
    h->cub_tmp_bytes = n2 + 256;
  }
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->uploaded = true;
  h->sorted = false;
  return KSCHED_OK;
}

static int run_sort(ksched_handle* h) {
  const int P = h->n_pods;
  if (P == 0) return KSCHED_OK;
  const int threads = 256, blocks = (P + threads - 1) / threads;
  sort_keys_kernel<<<blocks, threads, 0, h->stream>>>(P, h->d_classes.ptr, h->d_pod_class.ptr, h->d_ts.ptr, h->d_uid_rank.ptr, h->d_k_cpu.ptr,
                                                      h->d_k_mem.ptr, h->d_k_tie.ptr, h->d_order.ptr);
  // LSD over three 64-bit keys with a stable radix sort: tie-break key first, cpu last
  size_t tmp = h->cub_tmp_bytes;
  uint32_t *ia = h->d_order.ptr, *ib = h->d_idx_tmp.ptr;
  CUDA_TRY(h, cub::DeviceRadixSort::SortPairs(h->d_cub_tmp.ptr, tmp, h->d_k_tie.ptr, h->d_k_tmp.ptr, ia, ib, P, 0, 64, h->stream));
  gather_u64_kernel<<<blocks, threads, 0, h->stream>>>(P, h->d_k_mem.ptr, ib, h->d_k_tie.ptr);
  tmp = h->cub_tmp_bytes;
  CUDA_TRY(h, cub::DeviceRadixSort::SortPairs(h->d_cub_tmp.ptr, tmp, h->d_k_tie.ptr, h->d_k_tmp.ptr, ib, ia, P, 0, 64, h->stream));
  gather_u64_kernel<<<blocks, threads, 0, h->stream>>>(P, h->d_k_cpu.ptr, ia, h->d_k_tie.ptr);
  tmp = h->cub_tmp_bytes;
  CUDA_TRY(h, cub::DeviceRadixSort::SortPairs(h->d_cub_tmp.ptr, tmp, h->d_k_tie.ptr, h->d_k_tmp.ptr, ia, ib, P, 0, 64, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_order.ptr, ib, (size_t)P * sizeof(uint32_t), cudaMemcpyDeviceToDevice, h->stream));
  const int gblocks = std::min((P + 7) / 8, 148 * 8);
  gather_rows_kernel<<<gblocks, 256, 0, h->stream>>>(P, h->d_classes.ptr, h->d_pod_class.ptr, h->d_order.ptr, h->d_rows.ptr);
  h->tm.sort_launches = 7;
  h->sorted = true;
  return KSCHED_OK;
}

static void fill_k1(ksched_handle* h, K1Params& k1) {
  k1.cat = h->cat;
  k1.rows = h->d_rows.ptr;
  k1.n_pods = h->n_pods;
  k1.itype_sets = h->d_itype_sets.ptr;
  k1.F = h->d_F.ptr;
  k1.best = h->d_best.ptr;
  const int W32 = h->cat.W32;
  if (h->world > 1) {
    ksched_shard_range(W32, h->rank, h->world, &k1.word_begin, &k1.word_end);
  } else {
    k1.word_begin = 0;
    k1.word_end = W32;
  }
}

static std::mutex g_k1_mu;
static cudaEvent_t g_k1_done[64] = {};

static int run_feasibility(ksched_handle* h) {
  if (h->n_pods == 0) return KSCHED_OK;
  K1Params k1;
  fill_k1(h, k1);
  const DevCatalog& c = h->cat;
  const size_t alloc_bytes = (size_t)c.n_res * c.n_types * sizeof(int64_t);
  k1.alloc_in_smem = alloc_bytes <= (size_t)(128 << 10) ? 1 : 0;
  size_t smem = sizeof(ksched_template) * c.n_templates + sizeof(ksched_keyinfo) * KSCHED_MAX_KEYS + (k1.alloc_in_smem ? alloc_bytes : 0) +
                sizeof(int16_t) * KSCHED_MAX_KEYS * 64 + sizeof(int16_t) * 64;
  k1.n_valrows = h->n_valrows;
  k1.n_offrows = h->n_offrows;
  const size_t table_bytes = (size_t)(h->n_valrows + 2 * c.n_keys + h->n_offrows + 1 + c.n_templates) * c.W32 * sizeof(uint32_t);
  k1.tables_in_smem = smem + table_bytes <= (size_t)(160 << 10) ? 1 : 0;
  if (k1.tables_in_smem) smem += table_bytes;
  // one CTA per SM at most; every warp takes a contiguous chunk of >= 8 rows
  const int warps_per_block = kK1Threads / 32;
  const int want_warps = std::max(1, (h->n_pods + 7) / 8);
  const int blocks = std::max(1, std::min(148, (want_warps + warps_per_block - 1) / warps_per_block));
  CUDA_TRY(h, cudaFuncSetAttribute(feasibility_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k1.dbg = nullptr;
#ifdef KSCHED_PROFILE_K1
  CUDA_TRY(h, h->d_k1dbg.ensure(8));
  CUDA_TRY(h, cudaMemsetAsync(h->d_k1dbg.ptr, 0, 8 * sizeof(long long), h->stream));
  k1.dbg = h->d_k1dbg.ptr;
#endif
  {
    std::lock_guard<std::mutex> lock(g_k1_mu);  // one __constant__ parameter block per device: see run_pack
    if (h->device < 0 || h->device >= 64) { h->err = "device ordinal out of range"; return KSCHED_ERR_INVALID; }
    cudaEvent_t& done = g_k1_done[h->device];
    if (!done) CUDA_TRY(h, cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream, done, 0));
    h->k1_host = k1;
    CUDA_TRY(h, cudaMemcpyToSymbolAsync(g_k1, &h->k1_host, sizeof(K1Params), 0, cudaMemcpyHostToDevice, h->stream));
    feasibility_kernel<<<blocks, kK1Threads, smem, h->stream>>>();
    CUDA_TRY(h, cudaEventRecord(done, h->stream));
  }
#ifdef KSCHED_PROFILE_K1
  {
    long long dbg[8];
    cudaMemcpyAsync(dbg, h->d_k1dbg.ptr, sizeof dbg, cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
    fprintf(stderr, "[k1 profile] blocks=%d max cycles: stage=%lld compute=%lld total=%lld max_computes_per_warp=%lld | total computes=%lld avg compute cycles=%lld\n",
            blocks, dbg[0], dbg[1], dbg[2], dbg[3], dbg[4], dbg[4] ? dbg[5] / dbg[4] : 0);
  }
#endif
  h->tm.feasibility_launches = 1;
  const long long C = (long long)h->cat.n_templates * h->cat.n_types;
  h->tm.feasibility_bytes = (long long)h->n_pods * 256 + C * 256 + (long long)h->n_pods * C / 8;
  return KSCHED_OK;
}

static int reset_state(ksched_handle* h) {
  const int P = h->n_pods, NE = std::max(h->n_existing, 1), NG = std::max(h->n_groups, 1);
  CUDA_TRY(h, cudaMemcpyAsync(h->d_pod_class.ptr, h->d_pod_class0.ptr, (size_t)P * 4, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(h, cudaMemsetAsync(h->d_relax.ptr, 0, (size_t)std::max(P, 1) * 4, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_ex_req.ptr, h->d_ex_req0.ptr, (size_t)8 * NE * 8, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_ex_req_present.ptr, h->d_ex_req_present0.ptr, (size_t)NE * 4, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_ex_vals.ptr, h->d_ex_vals0.ptr, (size_t)16 * NE * 8, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_ex_meta.ptr, h->d_ex_meta0.ptr, (size_t)NE * 8, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_ex_hp.ptr, h->d_ex_hp0.ptr, (size_t)NE * 8, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(h, cudaMemsetAsync(h->d_ex_closed.ptr, 0, (size_t)NE, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_grp_cnt.ptr, h->d_grp_cnt0.ptr, (size_t)NG * 64 * 4, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_grp_registered.ptr, h->d_grp_registered0.ptr, (size_t)NG * 8, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_grp_host_total.ptr, h->d_grp_host_total0.ptr, (size_t)NG * 4, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_grp_active.ptr, h->d_grp_active0.ptr, (size_t)NG, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(h, cudaMemsetAsync(h->d_grp_min_slot.ptr, 0, (size_t)NG * 4, h->stream));
  const size_t hs = (size_t)std::max(h->n_hostgroups, 1) * ((size_t)h->n_existing + h->max_new);
  CUDA_TRY(h, cudaMemcpyAsync(h->d_grp_host.ptr, h->d_grp_host0.ptr, hs * 2, cudaMemcpyDeviceToDevice, h->stream));
  std::vector<int64_t> rem((size_t)h->cat.n_templates * KSCHED_MAX_RES);
  for (int v = 0; v < h->cat.n_templates; ++v)
    for (int r = 0; r < KSCHED_MAX_RES; ++r) rem[(size_t)v * KSCHED_MAX_RES + r] = h->h_templates[v].remaining[r];
  CUDA_TRY(h, cudaMemcpyAsync(h->d_remaining.ptr, rem.data(), rem.size() * 8, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));  // rem is a stack vector
  CUDA_TRY(h, cudaMemsetAsync(h->d_counters.ptr, 0, 32 * sizeof(long long), h->stream));
  CUDA_TRY(h, cudaMemsetAsync(h->d_fc_state.ptr, 0, (size_t)std::max(h->n_classes, 1) * h->cat.n_templates, h->stream));
  CUDA_TRY(h, cudaMemsetAsync(h->d_fc_front_state.ptr, 0, (size_t)std::max(h->n_classes, 1) * h->cat.n_templates, h->stream));
  CUDA_TRY(h, cudaMemsetAsync(h->d_fd_state.ptr, 0, (size_t)kFreshMemoSlots, h->stream));
  return KSCHED_OK;
}

// g_k2 is one __constant__ object per device: a launch may only overwrite it after the previous pack kernel on that
// device (possibly of another handle, on another stream) has finished. Ordered on the device, never on the host.
static std::mutex g_k2_mu;
static cudaEvent_t g_k2_done[64] = {};

static int run_pack(ksched_handle* h) {
  K2Params& k2 = h->k2_host;
  k2.cat = h->cat;
  PackState& s = k2.st;
  s.relx = h->d_relx.ptr;
  s.classes = h->d_classes.ptr; s.groups = h->d_groups.ptr; s.class_topo = h->d_class_topo.ptr; s.filter_terms = h->d_filter_terms.ptr;
  s.itype_sets = h->d_itype_sets.ptr; s.itype_complement = h->d_itype_comp.ptr; s.hostname_reqs = h->d_hostname_reqs.ptr;
  s.order = h->d_order.ptr; s.rows = h->d_rows.ptr; s.F = h->d_F.ptr; s.best = h->d_best.ptr;
  s.pod_pos = h->d_pod_pos.ptr; s.use_F = h->world == 1 ? 1 : 0;
  s.n_pods = h->n_pods; s.n_classes = h->n_classes; s.n_existing = h->n_existing; s.n_groups = h->n_groups; s.max_new = h->max_new;
  for (int r = 0; r < KSCHED_MAX_RES; ++r) s.min_req[r] = h->min_req[r];
  s.pod_class = h->d_pod_class.ptr; s.relax_level = h->d_relax.ptr; s.assign = h->d_assign.ptr; s.place_seq = h->d_place_seq.ptr;
  s.queue = h->d_queue.ptr; s.last_len = h->d_last_len.ptr; s.last_epoch = h->d_last_epoch.ptr;
  s.ex_req = h->d_ex_req.ptr; s.ex_avail = h->d_ex_avail.ptr; s.ex_req_present = h->d_ex_req_present.ptr; s.ex_avail_present = h->d_ex_avail_present.ptr;
  s.ex_vals = h->d_ex_vals.ptr; s.ex_meta = h->d_ex_meta.ptr; s.ex_taintset = h->d_ex_taintset.ptr; s.ex_itype = h->d_ex_itype.ptr;
  s.ex_hp = h->d_ex_hp.ptr; s.ex_closed = h->d_ex_closed.ptr;
  s.nn_tmpl = h->d_nn_tmpl.ptr; s.nn_count = h->d_nn_count.ptr; s.nn_tb = h->d_nn_tb.ptr; s.nn_req = h->d_nn_req.ptr;
  s.nn_req_present = h->d_nn_req_present.ptr;
  s.nn_vals = h->d_nn_vals.ptr; s.nn_meta = h->d_nn_meta.ptr; s.nn_opts = h->d_nn_opts.ptr; s.nn_hp = h->d_nn_hp.ptr;
  s.ov_key = h->d_ov_key.ptr; s.ov_q = h->d_ov_q.ptr; s.ov_bound = h->d_ov_bound.ptr; s.ov_bound2 = h->d_ov_bound2.ptr; s.ov_node = h->d_ov_node.ptr; s.ov_flags = h->d_ov_flags.ptr;
  s.fc_state = h->d_fc_state.ptr; s.fc_opts = h->d_fc_opts.ptr; s.fc_bound = h->d_fc_bound.ptr; s.fc_bound2 = h->d_fc_bound2.ptr; s.fc_dom = h->d_fc_dom.ptr; s.fc_front_state = h->d_fc_front_state.ptr;
  s.fc_vals = h->d_fc_vals.ptr; s.fc_meta = h->d_fc_meta.ptr; s.fc_q = h->d_fc_q.ptr; s.fc_qp = h->d_fc_qp.ptr;
  s.ov_absorbed = h->d_ov_absorbed.ptr; s.ov_rejected = h->d_ov_rejected.ptr;
  s.fd_cap = kFreshMemoSlots; s.fd_state = h->d_fd_state.ptr; s.fd_fc = h->d_fd_fc.ptr; s.fd_meta = h->d_fd_meta.ptr; s.fd_vals = h->d_fd_vals.ptr;
  s.fd_opts = h->d_fd_opts.ptr; s.fd_bound = h->d_fd_bound.ptr; s.fd_bound2 = h->d_fd_bound2.ptr; s.fd_dom = h->d_fd_dom.ptr;
  s.count_visited = h->count_visited;
  s.grp_cnt = h->d_grp_cnt.ptr; s.grp_registered = h->d_grp_registered.ptr; s.grp_host = h->d_grp_host.ptr;
  s.grp_host_row = h->d_grp_host_row.ptr; s.grp_host_total = h->d_grp_host_total.ptr; s.remaining = h->d_remaining.ptr;
  s.grp_active = h->d_grp_active.ptr; s.grp_min_slot = h->d_grp_min_slot.ptr;
  s.counters = h->d_counters.ptr;
  const size_t alloc_bytes = (size_t)h->cat.n_res * h->cat.n_types * sizeof(int64_t);
  // 227 KB per CTA on sm_100a: the hot node window + ~4 KB of static shared memory come first
  const size_t smem_left = (size_t)(227 << 10) - sizeof(HotSmem) - (size_t)(6 << 10);
  s.alloc_in_smem = alloc_bytes <= smem_left ? 1 : 0;
  const size_t smem = sizeof(HotSmem) + (s.alloc_in_smem ? alloc_bytes : 0);
  CUDA_TRY(h, cudaFuncSetAttribute(pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // block size: the chain is latency-bound on ONE thread's commit; more warps only help when there are many candidate
  // nodes to examine per pod (existing nodes, large in-flight sets)
  int threads = h->n_existing >= 2048 ? kPackThreads : (h->n_existing >= 256 ? 256 : 128);
  if (const char* e = getenv("KSCHED_PACK_THREADS")) { int v = atoi(e); if (v >= 32 && v <= kPackThreads && v % 32 == 0) threads = v; }
  {
    std::lock_guard<std::mutex> lock(g_k2_mu);
    if (h->device < 0 || h->device >= 64) { h->err = "device ordinal out of range"; return KSCHED_ERR_INVALID; }
    cudaEvent_t& done = g_k2_done[h->device];
    if (!done) CUDA_TRY(h, cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream, done, 0));  // no-op for an event that was never recorded
    CUDA_TRY(h, cudaMemcpyToSymbolAsync(g_k2, &k2, sizeof(K2Params), 0, cudaMemcpyHostToDevice, h->stream));
    pack_kernel<<<1, threads, smem, h->stream>>>();
    CUDA_TRY(h, cudaEventRecord(done, h->stream));
  }
  finalize_options_kernel<<<148, 256, 0, h->stream>>>(h->cat, h->d_counters.ptr, h->d_nn_req.ptr, h->d_nn_req_present.ptr, h->d_nn_opts.ptr, h->max_new);
  h->tm.pack_launches = 2;
  return KSCHED_OK;
}

static int flush_l2(ksched_handle* h) {
  const size_t n = (size_t)64 << 20;  // 256 MiB of u32 > 126 MB L2
  CUDA_TRY(h, h->d_flush.ensure(n));
  flush_kernel<<<148 * 4, 512, 0, h->stream>>>(h->d_flush.ptr, n);
  return KSCHED_OK;
}

static float ev_us(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  return ms * 1000.f;
}

int ksched_run_resident(ksched_handle* h, int do_flush) {
  if (!h || !h->uploaded) return KSCHED_ERR_INVALID;
  CUDA_TRY(h, cudaSetDevice(h->device));
  int rc;
  if ((rc = reset_state(h)) != KSCHED_OK) return rc;
  if (do_flush && (rc = flush_l2(h)) != KSCHED_OK) return rc;
  CUDA_TRY(h, cudaEventRecord(h->ev[0], h->stream));
  if ((rc = run_sort(h)) != KSCHED_OK) return rc;
  CUDA_TRY(h, cudaEventRecord(h->ev[1], h->stream));
  if ((rc = run_feasibility(h)) != KSCHED_OK) return rc;
  CUDA_TRY(h, cudaEventRecord(h->ev[2], h->stream));
  h->tm.allreduce_us = 0;
  if (h->world > 1 && h->comm) {
    ncclResult_t r = ncclAllReduce(h->d_best.ptr, h->d_best.ptr, (size_t)h->n_pods, ncclUint64, ncclMin, h->comm, h->stream);
    if (r != ncclSuccess) { h->err = std::string("ncclAllReduce: ") + ncclGetErrorString(r); return KSCHED_ERR_NCCL; }
  }
  CUDA_TRY(h, cudaEventRecord(h->ev[3], h->stream));
  if ((rc = run_pack(h)) != KSCHED_OK) return rc;
  CUDA_TRY(h, cudaEventRecord(h->ev[4], h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  CUDA_TRY(h, cudaGetLastError());
  h->tm.sort_us = ev_us(h->ev[0], h->ev[1]);
  h->tm.feasibility_us = ev_us(h->ev[1], h->ev[2]);
  h->tm.allreduce_us = ev_us(h->ev[2], h->ev[3]);
  h->tm.pack_us = ev_us(h->ev[3], h->ev[4]);
  h->tm.total_us = ev_us(h->ev[0], h->ev[4]);
  return KSCHED_OK;
}

int ksched_run_feasibility_only(ksched_handle* h, int do_flush, float* elapsed_us) {
  if (!h || !h->uploaded) return KSCHED_ERR_INVALID;
  CUDA_TRY(h, cudaSetDevice(h->device));
  int rc;
  if (!h->sorted) {  // K0 first: the kernel streams the FFD-ordered pod-row matrix
    CUDA_TRY(h, cudaMemcpyAsync(h->d_pod_class.ptr, h->d_pod_class0.ptr, (size_t)h->n_pods * 4, cudaMemcpyDeviceToDevice, h->stream));
    if ((rc = run_sort(h)) != KSCHED_OK) return rc;
  }
  if (do_flush && (rc = flush_l2(h)) != KSCHED_OK) return rc;
  CUDA_TRY(h, cudaEventRecord(h->ev[5], h->stream));
  if ((rc = run_feasibility(h)) != KSCHED_OK) return rc;
  CUDA_TRY(h, cudaEventRecord(h->ev[6], h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  CUDA_TRY(h, cudaGetLastError());
  if (elapsed_us) *elapsed_us = ev_us(h->ev[5], h->ev[6]);
  h->tm.feasibility_us = ev_us(h->ev[5], h->ev[6]);
  return KSCHED_OK;
}

int ksched_download(ksched_handle* h, const ksched_problem* pb, ksched_result* res) {
  if (!h || !pb || !res || !h->uploaded) return KSCHED_ERR_INVALID;
  CUDA_TRY(h, cudaSetDevice(h->device));
  const int P = h->n_pods, NE = h->n_existing, MAXN = h->max_new, W32 = h->cat.W32, W64 = h->W64, V = h->cat.n_templates;
  long long counters[32];
  CUDA_TRY(h, cudaMemcpyAsync(counters, h->d_counters.ptr, sizeof counters, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
#ifdef KSCHED_PROFILE_PACK
  fprintf(stderr, "[pack profile] generic: topo=%lld existing=%lld eval=%lld commit=%lld fresh=%lld fail=%lld | fastblock=%lld genericcall=%lld | "
                  "n_generic=%lld inflight_placed=%lld fresh_steps=%lld failures=%lld paths[rej,cached,row,dyn,cachedempty]=%lld,%lld,%lld,%lld,%lld steps=%lld\n",
          counters[8], counters[9], counters[10], counters[11], counters[12], counters[13], counters[14], counters[15], counters[17], counters[18],
          counters[19], counters[20], counters[21], counters[22], counters[23], counters[24], counters[25], counters[5]);
  fprintf(stderr, "[pack profile] fresh: decide=%lld words=%lld commit=%lld rest=%lld | in-flight verify=%lld\n", counters[26], counters[27], counters[28],
          counters[12], counters[29]);
  fprintf(stderr, "[pack profile] in-flight commit: winner+barrier=%lld record=%lld rest=%lld\n", counters[30], counters[31], counters[11]);
#endif
  if (counters[4] != 0) {
    h->err = counters[4] == KSCHED_ERR_OVERFLOW ? "new-node capacity exceeded" : "a pod is constrained by more topology groups than the kernel supports";
    return (int)counters[4];
  }
  const int n_new = (int)counters[0];
  h->tm.d2h_bytes = (int64_t)sizeof counters + (int64_t)P * 4 * ((res->assign != nullptr) + (res->relax_level != nullptr) + (res->place_seq != nullptr)) +
                    (int64_t)n_new * (1 + 4 + 4 + 8 + 64 + 128 + (int64_t)W32 * 4) + (res->existing_reqs ? (int64_t)NE * 136 : 0) +
                    (res->launch ? (int64_t)n_new * (int64_t)sizeof(ksched_launch_choice) : 0) +
                    (res->feasibility ? (int64_t)P * V * W32 * 4 + (int64_t)P * 4 : 0) + (res->best_column ? (int64_t)P * 12 : 0);
  res->n_new_nodes = n_new;
  res->n_unscheduled = (int)counters[1];
  res->nodes_visited = counters[2];
  res->add_calls = counters[3];
  h->tm.pack_steps = counters[5];
  if (res->assign) CUDA_TRY(h, cudaMemcpyAsync(res->assign, h->d_assign.ptr, (size_t)P * 4, cudaMemcpyDeviceToHost, h->stream));
  if (res->relax_level) CUDA_TRY(h, cudaMemcpyAsync(res->relax_level, h->d_relax.ptr, (size_t)P * 4, cudaMemcpyDeviceToHost, h->stream));
  if (res->place_seq) CUDA_TRY(h, cudaMemcpyAsync(res->place_seq, h->d_place_seq.ptr, (size_t)P * 4, cudaMemcpyDeviceToHost, h->stream));
  if (n_new > 0 && res->new_nodes && res->new_node_types) {
    std::vector<uint8_t> tmpl(n_new);
    std::vector<int32_t> count(n_new);
    std::vector<int64_t> req((size_t)8 * n_new);
    std::vector<uint32_t> reqp(n_new), opts((size_t)W32 * n_new);
    std::vector<uint64_t> vals((size_t)16 * n_new), meta(n_new);
    CUDA_TRY(h, cudaMemcpyAsync(tmpl.data(), h->d_nn_tmpl.ptr, n_new, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(count.data(), h->d_nn_count.ptr, (size_t)n_new * 4, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(reqp.data(), h->d_nn_req_present.ptr, (size_t)n_new * 4, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(meta.data(), h->d_nn_meta.ptr, (size_t)n_new * 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpy2DAsync(req.data(), (size_t)n_new * 8, h->d_nn_req.ptr, (size_t)MAXN * 8, (size_t)n_new * 8, 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpy2DAsync(vals.data(), (size_t)n_new * 8, h->d_nn_vals.ptr, (size_t)MAXN * 8, (size_t)n_new * 8, 16, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpy2DAsync(opts.data(), (size_t)n_new * 4, h->d_nn_opts.ptr, (size_t)MAXN * 4, (size_t)n_new * 4, W32, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    for (int n = 0; n < n_new; ++n) {
      ksched_new_node& o = res->new_nodes[n];
      std::memset(&o, 0, sizeof o);
      o.template_index = tmpl[n];
      o.pod_count = count[n];
      o.requests_present = reqp[n];
      for (int r = 0; r < 8; ++r) o.requests[r] = req[(size_t)r * n_new + n];
      for (int k = 0; k < 16; ++k) o.reqs.values[k] = vals[(size_t)k * n_new + n];
      o.reqs.meta = meta[n];
      uint32_t* dst = reinterpret_cast<uint32_t*>(res->new_node_types + (size_t)n * W64);
      for (int w = 0; w < W32; ++w) dst[w] = opts[(size_t)w * n_new + n];
    }
  }
  if (res->launch && n_new > 0) {
    if (!h->cat.offer_keys) { h->err = "ksched_result.launch needs ksched_catalog.offering_keys"; return KSCHED_ERR_INVALID; }
    CUDA_TRY(h, h->d_launch.ensure((size_t)n_new));
    launch_choice_kernel<<<148, 256, 0, h->stream>>>(h->cat, h->d_counters.ptr, h->d_nn_vals.ptr, h->d_nn_meta.ptr, h->d_nn_opts.ptr, MAXN,
                                                     h->d_launch.ptr);
    CUDA_TRY(h, cudaMemcpyAsync(res->launch, h->d_launch.ptr, (size_t)n_new * sizeof(ksched_launch_choice), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  }
  if (res->existing_reqs && NE > 0) {
    std::vector<uint64_t> vals((size_t)16 * NE), meta(NE);
    CUDA_TRY(h, cudaMemcpyAsync(vals.data(), h->d_ex_vals.ptr, vals.size() * 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(meta.data(), h->d_ex_meta.ptr, meta.size() * 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    for (int e = 0; e < NE; ++e) {
      for (int k = 0; k < 16; ++k) res->existing_reqs[e].values[k] = vals[(size_t)k * NE + e];
      res->existing_reqs[e].meta = meta[e];
    }
  }
  if (res->feasibility) {
    // device rows are in FFD (queue) order; the caller gets them in its own pod order
    std::vector<uint32_t> order(P), F((size_t)P * V * W32);
    CUDA_TRY(h, cudaMemcpyAsync(order.data(), h->d_order.ptr, (size_t)P * 4, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(F.data(), h->d_F.ptr, F.size() * 4, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    const size_t rowb = (size_t)V * W32;
    for (int j = 0; j < P; ++j) std::memcpy(reinterpret_cast<uint32_t*>(res->feasibility) + (size_t)order[j] * rowb, &F[(size_t)j * rowb], rowb * 4);
  }
  if (res->best_column) {
    std::vector<uint32_t> order(P);
    std::vector<unsigned long long> best(P);
    CUDA_TRY(h, cudaMemcpyAsync(order.data(), h->d_order.ptr, (size_t)P * 4, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(best.data(), h->d_best.ptr, (size_t)P * 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    for (int j = 0; j < P; ++j) res->best_column[order[j]] = best[j];
  }
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return KSCHED_OK;
}

int ksched_solve(ksched_handle* h, const ksched_problem* pb, ksched_result* res) {
  if (!h || !pb || !res) return KSCHED_ERR_INVALID;
  auto t0 = std::chrono::steady_clock::now();
  int rc = ksched_upload(h, pb);
  if (rc != KSCHED_OK) return rc;
  auto t1 = std::chrono::steady_clock::now();
  rc = ksched_run_resident(h, 0);
  if (rc != KSCHED_OK) return rc;
  auto t2 = std::chrono::steady_clock::now();
  rc = ksched_download(h, pb, res);
  auto t3 = std::chrono::steady_clock::now();
  h->tm.upload_us = std::chrono::duration<double, std::micro>(t1 - t0).count();
  h->tm.download_us = std::chrono::duration<double, std::micro>(t3 - t2).count();
  return rc;
}

int ksched_get_timings(const ksched_handle* h, ksched_timings* out) {
  if (!h || !out) return KSCHED_ERR_INVALID;
  *out = h->tm;
  return KSCHED_OK;
}

}  // extern "C"
