// ORACLE — TEST INFRASTRUCTURE ONLY. C entry points for ctypes (tests, smoke, bench cpu_baseline).
#include <cstring>
#include <sstream>

#include "oracle.h"

using namespace oracle;

// spec mini-language: "In:A,B" "NotIn:1,2" "Exists" "DoesNotExist" "Gt:1" "Lt:9"
static Requirement parse_spec(const std::string& key, const std::string& spec) {
  std::string op = spec, vals;
  auto c = spec.find(':');
  if (c != std::string::npos) { op = spec.substr(0, c); vals = spec.substr(c + 1); }
  std::vector<std::string> v;
  if (c != std::string::npos) {
    std::stringstream ss(vals);
    std::string x;
    while (std::getline(ss, x, ',')) v.push_back(x);
  }
  kmodel::Op o = kmodel::Op::In;
  if (op == "In") o = kmodel::Op::In;
  else if (op == "NotIn") o = kmodel::Op::NotIn;
  else if (op == "Exists") o = kmodel::Op::Exists;
  else if (op == "DoesNotExist") o = kmodel::Op::DoesNotExist;
  else if (op == "Gt") o = kmodel::Op::Gt;
  else if (op == "Lt") o = kmodel::Op::Lt;
  return Requirement::New(key, o, v);
}
static std::string render(const Requirement& r) {
  std::string s = std::string("complement=") + (r.complement ? "1" : "0") + ";values=";
  bool first = true;
  for (auto& v : r.values) { if (!first) s += ","; s += v; first = false; }
  s += ";gt=" + (r.greater_than ? std::to_string(*r.greater_than) : "");
  s += ";lt=" + (r.less_than ? std::to_string(*r.less_than) : "");
  return s;
}
static int put(const std::string& s, char* out, int cap) {
  if ((int)s.size() + 1 > cap) return -1;
  std::memcpy(out, s.c_str(), s.size() + 1);
  return (int)s.size();
}

extern "C" {
void oracle_set_fast(int on) { set_fast(on != 0); }
int oracle_req_render(const char* a, char* out, int cap) { return put(render(parse_spec("key", a)), out, cap); }
int oracle_req_intersection(const char* a, const char* b, char* out, int cap) {
  return put(render(parse_spec("key", a).Intersection(parse_spec("key", b))), out, cap);
}
int oracle_req_has(const char* a, const char* v) { return parse_spec("key", a).Has(v) ? 1 : 0; }
int oracle_req_operator(const char* a) { return (int)parse_spec("key", a).Operator(); }
long long oracle_req_len(const char* a) { return parse_spec("key", a).Len(); }
int oracle_req_string(const char* a, const char* b, char* out, int cap) {
  Requirement r = parse_spec("key", a);
  if (b && *b) r = r.Intersection(parse_spec("key", b));
  return put(r.String(), out, cap);
}
// Requirements.Compatible on one key. a/b == "" means "no requirement for the key".
int oracle_reqs_compatible(const char* key, const char* a, const char* b, int key_is_well_known) {
  Requirements ra, rb;
  if (a && *a) ra.Add(parse_spec(key, a));
  if (b && *b) rb.Add(parse_spec(key, b));
  std::set<std::string> wk = well_known_labels({});
  if (key_is_well_known) wk.insert(normalize_key(key)); else wk.erase(normalize_key(key));
  return ra.Compatible(rb, wk) ? 1 : 0;
}
int oracle_normalize_key(const char* k, char* out, int cap) { return put(normalize_key(k), out, cap); }

int oracle_solve(const kmodel::Problem* P, const int* candidates, int ncand, kmodel::Result* out) {
  std::vector<int> c(candidates, candidates + ncand);
  solve(*P, c, *out);
  return out->error.empty() ? 0 : -1;
}

// consolidate: returns action; fills scalars; options/probes copied up to caps
// one probe: computeConsolidation over the `count` cheapest-to-disrupt candidates; returns the action, fills the options
int oracle_consolidate_probe(const kmodel::Problem* P, int count, int* options, int options_cap, int* n_options, char* err, int err_cap) {
  ConsolidationResult r;
  consolidate(*P, r, count);
  *n_options = (int)r.replacement_options.size();
  for (int i = 0; i < *n_options && i < options_cap; ++i) options[i] = r.replacement_options[i];
  put(r.error, err, err_cap);
  return r.error.empty() ? r.action : -1;
}

int oracle_consolidate(const kmodel::Problem* P, int* nodes_removed, int* simulations, int* options, int options_cap,
                       int* n_options, int* probes, int* probe_actions, int probes_cap, int* n_probes, char* err, int err_cap) {
  ConsolidationResult r;
  consolidate(*P, r);
  *nodes_removed = r.nodes_removed;
  *simulations = r.simulations;
  *n_options = (int)r.replacement_options.size();
  for (int i = 0; i < *n_options && i < options_cap; ++i) options[i] = r.replacement_options[i];
  *n_probes = (int)r.probes.size();
  for (int i = 0; i < *n_probes && i < probes_cap; ++i) { probes[i] = r.probes[i]; probe_actions[i] = r.probe_actions[i]; }
  put(r.error, err, err_cap);
  return r.error.empty() ? r.action : -1;
}

// single-node consolidation: returns the action; node = Problem.nodes index removed (-1 none); probes = positions tried
int oracle_consolidate_single(const kmodel::Problem* P, int only_candidate, int* node, int* options, int options_cap, int* n_options,
                              int* n_probes, char* err, int err_cap) {
  ConsolidationResult r;
  consolidate_single(*P, r, only_candidate);
  *node = r.single_node;
  *n_options = (int)r.replacement_options.size();
  for (int i = 0; i < *n_options && i < options_cap; ++i) options[i] = r.replacement_options[i];
  *n_probes = (int)r.probes.size();
  put(r.error, err, err_cap);
  return r.error.empty() ? r.action : -1;
}
// candidate ranking: order = Problem.nodes indices in disruption order, cost parallel; returns the number of candidates
int oracle_rank_candidates(const kmodel::Problem* P, int* order, double* cost, int cap) {
  std::vector<int> o;
  std::vector<double> c;
  rank_candidates(*P, &o, &c);
  for (size_t i = 0; i < o.size() && (int)i < cap; ++i) { order[i] = o[i]; cost[i] = c[i]; }
  return (int)o.size();
}
}
