// Requirement set algebra on dictionary bitmasks — ONE implementation compiled for host (encoder,
// CPU golden-vector tests) and device (feasibility + pack kernels).
//
// Mirrors pkg/scheduling/requirement.go and requirements.go operation by operation: a requirement is
// {complement, values, greaterThan, lessThan} exactly as in requirement.go:36-42, with `values` a bitmask
// over the key's dictionary instead of a sets.String.
#pragma once
#include <stdint.h>

#include "ksched.h"

#if defined(__CUDACC__)
#define KS_HD __host__ __device__ __forceinline__
#else
#define KS_HD inline
#endif

namespace ksched {

struct Req {
  uint64_t values;  // members (complement=0) or excluded members (complement=1)
  int64_t gt, lt;
  bool present, complement, has_gt, has_lt;
};

struct KeyMeta {
  uint64_t int_mask;          // dictionary entries that parse as integers
  const int64_t* int_values;  // [64] or nullptr when the key has no integer entries / no bounds in play
  // Region form of Gt/Lt (what the DEVICE carries instead of gt / lt, see ksched_key_regions in ksched.h): a bounded
  // complement requirement excludes the region bits outside (gt, lt) and the dictionary values outside it. nullptr: the
  // key has no region bits (host algebra, and every key without thresholds / complement instance types).
  const ksched_key_regions* regions;
};

KS_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __popcll(x);
#else
  return __builtin_popcountll(x);
#endif
}

KS_HD Req req_absent() { return Req{0, 0, 0, false, true, false, false}; }
// Requirements.Get of an undefined key: Exists (requirements.go:115-121)
KS_HD Req req_exists() { return Req{0, 0, 0, true, true, false, false}; }

KS_HD Req req_load(const ksched_reqset& s, const ksched_bounds* b, int k) {
  Req r;
  uint64_t m = s.meta;
  r.present = (m >> (KSCHED_META_PRESENT_SHIFT + k)) & 1;
  r.complement = (m >> (KSCHED_META_COMPLEMENT_SHIFT + k)) & 1;
  r.has_gt = (m >> (KSCHED_META_HASGT_SHIFT + k)) & 1;
  r.has_lt = (m >> (KSCHED_META_HASLT_SHIFT + k)) & 1;
  r.values = s.values[k];
  r.gt = (r.has_gt && b) ? b->gt[k] : 0;
  r.lt = (r.has_lt && b) ? b->lt[k] : 0;
  return r;
}
KS_HD void req_store(ksched_reqset& s, ksched_bounds* b, int k, const Req& r) {
  uint64_t bit = 1ull << k;
  uint64_t clear = ~((bit << KSCHED_META_PRESENT_SHIFT) | (bit << KSCHED_META_COMPLEMENT_SHIFT) |
                     (bit << KSCHED_META_HASGT_SHIFT) | (bit << KSCHED_META_HASLT_SHIFT));
  uint64_t m = s.meta & clear;
  if (r.present) m |= bit << KSCHED_META_PRESENT_SHIFT;
  if (r.present && r.complement) m |= bit << KSCHED_META_COMPLEMENT_SHIFT;
  if (r.present && r.has_gt) m |= bit << KSCHED_META_HASGT_SHIFT;
  if (r.present && r.has_lt) m |= bit << KSCHED_META_HASLT_SHIFT;
  s.meta = m;
  s.values[k] = r.present ? r.values : 0;
  if (b) {
    b->gt[k] = r.has_gt ? r.gt : 0;
    b->lt[k] = r.has_lt ? r.lt : 0;
  }
}

// dictionary entries inside (gt, lt): withinIntPtrs (requirement.go:227-243) as a mask
KS_HD uint64_t within_mask(bool has_gt, int64_t gt, bool has_lt, int64_t lt, const KeyMeta& km) {
#if defined(__CUDA_ARCH__)
  // Device code never sees gt / lt: the encoder hands every bounded requirement over in region form (ksched_key_regions).
  return ~0ull;
#endif
  if (!has_gt && !has_lt) return ~0ull;
  uint64_t out = 0;
  uint64_t m = km.int_mask;  // non-integer values are invalid once bounds are set
  while (m) {
#if defined(__CUDA_ARCH__)
    int b = __ffsll((long long)m) - 1;
#else
    int b = __builtin_ctzll(m);
#endif
    m &= m - 1;
    int64_t v = km.int_values ? km.int_values[b] : 0;
    if (has_gt && gt >= v) continue;
    if (has_lt && lt <= v) continue;
    out |= 1ull << b;
  }
  return out;
}

// Requirement.Len() == 0 (requirement.go:199-204): complement sets are never empty
KS_HD bool req_len_zero(const Req& r) { return !r.complement && r.values == 0; }

// ---- region form (ksched_key_regions): number of excluded regions at the low / high end of a complement requirement
KS_HD int region_low(uint64_t values, const ksched_key_regions& g) {
  const uint64_t f = (values & g.region_mask) >> g.region_shift;
  int n = 0;
  while (n <= g.n_thresholds && ((f >> n) & 1)) ++n;
  return n;
}
KS_HD int region_high(uint64_t values, const ksched_key_regions& g) {
  const uint64_t f = (values & g.region_mask) >> g.region_shift;
  int n = 0;
  while (n <= g.n_thresholds && ((f >> (g.n_thresholds - n)) & 1)) ++n;
  return n;
}
// The values a complement requirement really excludes (Requirement.values of requirement.go:36-42; Intersection keeps only
// excluded values inside the bounds, :139-143): in region form `values` also holds everything the bounds cut away.
KS_HD uint64_t req_excluded(const Req& r, const KeyMeta& km) {
  if (!km.regions || !km.regions->region_mask) return r.values;
  const ksched_key_regions& g = *km.regions;
  return r.values & ~g.region_mask & g.above[region_low(r.values, g)] & g.below[region_high(r.values, g)];
}
// Operator() in {NotIn, DoesNotExist} (requirement.go:186-197); a complement set with bounds but no excluded value is Exists
KS_HD bool req_op_negative(const Req& r, const KeyMeta& km) { return r.complement ? (req_excluded(r, km) != 0) : (r.values == 0); }
// Len() == 1
KS_HD bool req_len_one(const Req& r) { return !r.complement && popc64(r.values) == 1; }

// Host algebra -> region form for one requirement (the conversion the encoder applies at the C-ABI boundary). Returns false
// when a bound is not one of the key's thresholds.
inline bool req_to_region_form(const Req& r, const ksched_key_regions& g, uint64_t dict_mask, Req* out) {
  *out = r;
  if (!r.present || (!r.has_gt && !r.has_lt)) return true;
  const int m = g.n_thresholds;
  int low = 0, high = 0;
  if (r.has_gt) { int i = 0; while (i < m && g.thresholds[i] != r.gt) ++i; if (i == m) return false; low = i + 1; }
  if (r.has_lt) { int i = 0; while (i < m && g.thresholds[i] != r.lt) ++i; if (i == m) return false; high = m - i; }
  uint64_t v = r.values;
  for (int i = 0; i < low; ++i) v |= 1ull << (g.region_shift + i);
  for (int i = 0; i < high; ++i) v |= 1ull << (g.region_shift + m - i);
  v |= dict_mask & ~(g.above[low] & g.below[high]);
  *out = Req{v, 0, 0, true, true, false, false};
  return true;
}
// Fill a key's region tables from its thresholds (ascending, distinct) and its integer dictionary entries.
inline void regions_build(ksched_key_regions* g, const int64_t* thresholds, int m, int n_dict_values, uint64_t dict_mask, const KeyMeta& km) {
  *g = ksched_key_regions{};
  g->region_shift = n_dict_values;
  g->n_thresholds = m;
  g->region_mask = ((1ull << (m + 1)) - 1) << n_dict_values;
  for (int i = 0; i < m; ++i) g->thresholds[i] = thresholds[i];
  for (int q = 0; q <= m; ++q) {
    g->above[q] = q == 0 ? dict_mask : within_mask(true, g->thresholds[q - 1], false, 0, km) & dict_mask;
    g->below[q] = q == 0 ? dict_mask : within_mask(false, 0, true, g->thresholds[m - q], km) & dict_mask;
  }
}

// Requirement.Intersection (requirement.go:117-150) in REGION form: the bounds are excluded region bits, so max(gt) / min(lt) /
// "drop members outside the bounds" are one OR / AND-NOT. This is the device's algebra; the host keeps {values, gt, lt} and the
// encoder converts at the boundary (tests/test_region_form.py checks that the two commute).
KS_HD Req req_intersect_regions(const Req& a, const Req& b, const KeyMeta& km) {
  Req r;
  r.present = true;
  r.complement = a.complement && b.complement;
  r.has_gt = r.has_lt = false;
  r.gt = r.lt = 0;
  if (a.complement && b.complement) {
    r.values = a.values | b.values;
    // greaterThan >= lessThan -> DoesNotExist (requirement.go:124-126): no region is left between the bounds
    if (km.regions && km.regions->region_mask && (r.values & km.regions->region_mask) == km.regions->region_mask) { r.complement = false; r.values = 0; }
  }
  else if (a.complement && !b.complement) r.values = b.values & ~a.values;
  else if (!a.complement && b.complement) r.values = a.values & ~b.values;
  else r.values = a.values & b.values;
  return r;
}

// Requirement.Intersection (requirement.go:117-150)
KS_HD Req req_intersect(const Req& a, const Req& b, const KeyMeta& km) {
#if defined(__CUDA_ARCH__)
  return req_intersect_regions(a, b, km);
#endif
  Req r;
  r.present = true;
  r.complement = a.complement && b.complement;
  r.has_gt = a.has_gt || b.has_gt;
  r.has_lt = a.has_lt || b.has_lt;
  r.gt = a.has_gt ? (b.has_gt ? (a.gt > b.gt ? a.gt : b.gt) : a.gt) : b.gt;
  r.lt = a.has_lt ? (b.has_lt ? (a.lt < b.lt ? a.lt : b.lt) : a.lt) : b.lt;
  if (r.has_gt && r.has_lt && r.gt >= r.lt) return Req{0, 0, 0, true, false, false, false};  // DoesNotExist
  uint64_t v;
  if (a.complement && b.complement) v = a.values | b.values;
  else if (a.complement && !b.complement) v = b.values & ~a.values;
  else if (!a.complement && b.complement) v = a.values & ~b.values;
  else v = a.values & b.values;
  v &= within_mask(r.has_gt, r.gt, r.has_lt, r.lt, km);
  r.values = v;
  if (!r.complement) { r.has_gt = r.has_lt = false; r.gt = r.lt = 0; }
  return r;
}

// Dictionary entries b with Requirement.Has(b) (requirement.go:171-176)
KS_HD uint64_t req_allowed(const Req& r, uint64_t dict_mask, const KeyMeta& km) {
  uint64_t w = within_mask(r.has_gt, r.gt, r.has_lt, r.lt, km);
  return (r.complement ? (~r.values & dict_mask) : r.values) & w;
}

// One key of Requirements.Intersects (requirements.go:189-206), `existing` is the receiver.
// Returns true when the key does NOT produce an error.
KS_HD bool key_intersects(const Req& existing, const Req& incoming, const KeyMeta& km) {
  if (!existing.present || !incoming.present) return true;
  Req i = req_intersect(existing, incoming, km);
  if (!req_len_zero(i)) return true;
  return req_op_negative(incoming, km) && req_op_negative(existing, km);
}
// One key of Requirements.Compatible (requirements.go:123-133), `node` is the receiver.
KS_HD bool key_compatible(const Req& node, const Req& incoming, bool well_known, const KeyMeta& km) {
  if (!incoming.present) return true;
  if (!well_known && !node.present && !req_op_negative(incoming, km)) return false;
  return key_intersects(node, incoming, km);
}
// Requirements.Add for one key (requirements.go:87-94): incoming.Intersection(existing)
KS_HD Req key_add(const Req& existing, const Req& incoming, const KeyMeta& km) {
  if (!incoming.present) return existing;
  if (!existing.present) return incoming;
  return req_intersect(incoming, existing, km);
}

}  // namespace ksched
