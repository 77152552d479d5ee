// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product; only tests/, smoke() and
// bench.py's cpu_baseline / --impl reference legs may load it.
//
// CPU restatement of the reference's label-requirement set algebra:
//   pkg/scheduling/requirement.go   (Requirement)
//   pkg/scheduling/requirements.go  (Requirements)
// Strings and std::set, exactly like the Go code's sets.String; no bitmasks here so that
// the product's dictionary/bitmask encoding is checked against an independent form.
// Iteration order wherever Go iterates a map: ascending string order.
#pragma once
#include <cstdint>
#include <limits>
#include <map>
#include <optional>
#include <set>
#include <string>
#include <vector>

#include "../karpenter-core_b200/host/model.h"

namespace oracle {

using kmodel::Op;

// v1alpha5.NormalizedLabels (pkg/apis/v1alpha5/labels.go:103-109)
inline std::string normalize_key(const std::string& k) {
  if (k == "failure-domain.beta.kubernetes.io/zone") return "topology.kubernetes.io/zone";
  if (k == "beta.kubernetes.io/arch") return "kubernetes.io/arch";
  if (k == "beta.kubernetes.io/os") return "kubernetes.io/os";
  if (k == "beta.kubernetes.io/instance-type") return "node.kubernetes.io/instance-type";
  if (k == "failure-domain.beta.kubernetes.io/region") return "topology.kubernetes.io/region";
  return k;
}

constexpr const char* kHostname = "kubernetes.io/hostname";
constexpr const char* kZone = "topology.kubernetes.io/zone";
constexpr const char* kCapacityType = "karpenter.sh/capacity-type";
constexpr const char* kProvisionerName = "karpenter.sh/provisioner-name";
constexpr const char* kInitialized = "karpenter.sh/initialized";
constexpr const char* kInstanceType = "node.kubernetes.io/instance-type";

// strconv.Atoi restated: optional sign, decimal digits only, must fit in int64.
inline bool atoi_strict(const std::string& s, int64_t* out) {
  if (s.empty()) return false;
  size_t i = 0;
  bool neg = false;
  if (s[0] == '+' || s[0] == '-') { neg = s[0] == '-'; i = 1; }
  if (i >= s.size()) return false;
  __int128 v = 0;
  for (; i < s.size(); ++i) {
    if (s[i] < '0' || s[i] > '9') return false;
    v = v * 10 + (s[i] - '0');
    if (v > (__int128)std::numeric_limits<int64_t>::max() + 1) return false;
  }
  if (neg) v = -v;
  if (v > std::numeric_limits<int64_t>::max() || v < std::numeric_limits<int64_t>::min()) return false;
  *out = (int64_t)v;
  return true;
}

// requirement.go:227-243
inline bool within_int_ptrs(const std::string& value, const std::optional<int64_t>& gt, const std::optional<int64_t>& lt) {
  if (!gt && !lt) return true;
  int64_t v;
  if (!atoi_strict(value, &v)) return false;
  if (gt && *gt >= v) return false;
  if (lt && *lt <= v) return false;
  return true;
}

struct Requirement {
  std::string key;
  bool complement = true;
  std::set<std::string> values;
  std::optional<int64_t> greater_than, less_than;

  // requirement.go:44-68
  static Requirement New(const std::string& key_in, Op op, const std::vector<std::string>& vals = {}) {
    Requirement r;
    r.key = normalize_key(key_in);
    r.complement = !(op == Op::In || op == Op::DoesNotExist);
    if (op == Op::In || op == Op::NotIn) r.values.insert(vals.begin(), vals.end());
    int64_t v = 0;
    if (op == Op::Gt) { atoi_strict(vals.at(0), &v); r.greater_than = v; }
    if (op == Op::Lt) { atoi_strict(vals.at(0), &v); r.less_than = v; }
    return r;
  }

  // requirement.go:199-204
  int64_t Len() const {
    if (complement) return std::numeric_limits<int64_t>::max() - (int64_t)values.size();
    return (int64_t)values.size();
  }
  // requirement.go:186-197
  Op Operator() const {
    if (complement) {
      if (Len() < std::numeric_limits<int64_t>::max()) return Op::NotIn;
      return Op::Exists;
    }
    if (Len() > 0) return Op::In;
    return Op::DoesNotExist;
  }
  // requirement.go:171-176
  bool Has(const std::string& v) const {
    if (complement) return !values.count(v) && within_int_ptrs(v, greater_than, less_than);
    return values.count(v) && within_int_ptrs(v, greater_than, less_than);
  }
  // requirement.go:117-150
  Requirement Intersection(const Requirement& o) const {
    bool comp = complement && o.complement;
    std::optional<int64_t> gt = greater_than, lt = less_than;
    if (o.greater_than && (!gt || *o.greater_than > *gt)) gt = o.greater_than;
    if (o.less_than && (!lt || *o.less_than < *lt)) lt = o.less_than;
    if (gt && lt && *gt >= *lt) return New(key, Op::DoesNotExist);
    std::set<std::string> vals;
    if (complement && o.complement) {
      vals = values;
      vals.insert(o.values.begin(), o.values.end());
    } else if (complement && !o.complement) {
      for (auto& v : o.values) if (!values.count(v)) vals.insert(v);
    } else if (!complement && o.complement) {
      for (auto& v : values) if (!o.values.count(v)) vals.insert(v);
    } else {
      for (auto& v : values) if (o.values.count(v)) vals.insert(v);
    }
    for (auto it = vals.begin(); it != vals.end();) {
      if (!within_int_ptrs(*it, gt, lt)) it = vals.erase(it); else ++it;
    }
    if (!comp) { gt.reset(); lt.reset(); }
    Requirement r;
    r.key = key; r.values = std::move(vals); r.complement = comp; r.greater_than = gt; r.less_than = lt;
    return r;
  }
  bool operator==(const Requirement& o) const {
    return key == o.key && complement == o.complement && values == o.values &&
           greater_than == o.greater_than && less_than == o.less_than;
  }
  // requirement.go:206-225 (values always in sorted order, as sets.String.List())
  std::string String() const {
    static const char* names[] = {"In", "NotIn", "Exists", "DoesNotExist", "Gt", "Lt"};
    Op op = Operator();
    std::string s = key + " " + names[(int)op];
    if (op == Op::In || op == Op::NotIn) {
      s += " [";
      size_t n = 0;
      for (auto& v : values) {
        if (n == 5 && values.size() > 5) { s += " and " + std::to_string(values.size() - 5) + " others"; break; }
        if (n) s += " ";
        s += v;
        ++n;
      }
      s += "]";
    }
    if (greater_than) s += " >" + std::to_string(*greater_than);
    if (less_than) s += " <" + std::to_string(*less_than);
    return s;
  }
  // full-fidelity rendering used for result comparison (no 5-value truncation)
  std::string Canonical() const {
    static const char* names[] = {"In", "NotIn", "Exists", "DoesNotExist", "Gt", "Lt"};
    std::string s = names[(int)Operator()];
    s += " [";
    bool first = true;
    for (auto& v : values) { if (!first) s += " "; s += v; first = false; }
    s += "]";
    if (greater_than) s += " >" + std::to_string(*greater_than);
    if (less_than) s += " <" + std::to_string(*less_than);
    return s;
  }
};

struct Requirements {
  std::map<std::string, Requirement> m;

  Requirements() = default;
  // requirements.go:87-94
  void Add(const Requirement& req) {
    auto it = m.find(req.key);
    if (it != m.end()) it->second = req.Intersection(it->second);
    else m.emplace(req.key, req);
  }
  void AddAll(const Requirements& o) { for (auto& kv : o.m) Add(kv.second); }
  bool Has(const std::string& k) const { return m.count(k) > 0; }
  // requirements.go:115-121
  Requirement Get(const std::string& k) const {
    auto it = m.find(k);
    if (it == m.end()) return Requirement::New(k, Op::Exists);
    return it->second;
  }
  // requirements.go:45-51
  static Requirements FromNodeSelector(const std::vector<kmodel::NodeSelectorRequirement>& rs) {
    Requirements r;
    for (auto& x : rs) r.Add(Requirement::New(x.key, x.op, x.values));
    return r;
  }
  // requirements.go:54-59
  static Requirements FromLabels(const kmodel::Labels& labels) {
    Requirements r;
    for (auto& kv : labels) r.Add(Requirement::New(kv.first, Op::In, {kv.second}));
    return r;
  }

  // requirements.go:189-206; returns true when there is NO error
  bool Intersects(const Requirements& incoming) const {
    for (auto& kv : m) {
      auto it = incoming.m.find(kv.first);
      if (it == incoming.m.end()) continue;
      const Requirement& existing = kv.second;
      const Requirement& inc = it->second;
      if (existing.Intersection(inc).Len() == 0) {
        Op io = inc.Operator();
        if (io == Op::NotIn || io == Op::DoesNotExist) {
          Op eo = existing.Operator();
          if (eo == Op::NotIn || eo == Op::DoesNotExist) continue;
        }
        return false;
      }
    }
    return true;
  }
  // requirements.go:123-133; returns true when there is NO error
  bool Compatible(const Requirements& incoming, const std::set<std::string>& well_known) const {
    bool ok = true;
    for (auto& kv : incoming.m) {
      if (well_known.count(kv.first)) continue;
      Op op = kv.second.Operator();
      if (Has(kv.first) || op == Op::NotIn || op == Op::DoesNotExist) continue;
      ok = false;
    }
    return Intersects(incoming) && ok;
  }
};

// v1alpha5.WellKnownLabels (labels.go:84-92) + provider additions
inline std::set<std::string> well_known_labels(const std::vector<std::string>& extra) {
  std::set<std::string> s = {kProvisionerName, kZone, "topology.kubernetes.io/region", kInstanceType,
                             "kubernetes.io/arch", "kubernetes.io/os", kCapacityType};
  s.insert(extra.begin(), extra.end());
  return s;
}

}  // namespace oracle
