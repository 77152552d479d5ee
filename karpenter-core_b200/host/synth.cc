// Deterministic synthetic clusters for BASELINE.json's configurations C1..C5.
// Instance-type shapes restate the reference's fake generators
// (pkg/cloudprovider/fake/instancetype.go:48-164): InstanceTypes(n) and InstanceTypesAssorted();
// pod request mix restates scheduling_benchmark_test.go:280-288. PRNG: SplitMix64.
#include "synth.h"

#include <cstdio>
#include <stdexcept>

namespace kmodel {
namespace {

struct SplitMix64 {
  uint64_t s;
  explicit SplitMix64(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
  }
  uint64_t below(uint64_t n) { return next() % n; }
  double unit() { return (double)(next() >> 11) / (double)(1ULL << 53); }
};

const int64_t kGi = 1024LL * 1024 * 1024 * 1000;  // milli-bytes
const int64_t kMi = 1024LL * 1024 * 1000;

double price_from(int64_t cpu_milli, int64_t mem_milli) {  // fake/instancetype.go:174-187
  return 0.1 * ((double)cpu_milli / 1000.0) + 0.1 * ((double)mem_milli / 1000.0) / 1e9;
}

InstanceType make_type(const std::string& name, int64_t cpu, int64_t mem_gi, int64_t pods, const std::string& arch,
                       const std::vector<std::string>& oses, const std::vector<Offering>& offerings) {
  InstanceType it;  // NewInstanceType fake/instancetype.go:48-106
  it.name = name;
  it.capacity = {{"cpu", cpu * 1000}, {"memory", mem_gi * kGi}, {"pods", pods * 1000}};
  it.offerings = offerings;
  it.kube_reserved = {{"cpu", 100}, {"memory", 10 * kMi}};
  std::vector<std::string> zones, cts;
  for (auto& o : offerings)
    if (o.available) { zones.push_back(o.zone); cts.push_back(o.capacity_type); }
  bool large = cpu > 4 && mem_gi > 8;
  it.requirements = {
      {"node.kubernetes.io/instance-type", Op::In, {name}},
      {"kubernetes.io/arch", Op::In, {arch}},
      {"kubernetes.io/os", Op::In, oses},
      {"topology.kubernetes.io/zone", Op::In, zones},
      {"karpenter.sh/capacity-type", Op::In, cts},
      {"size", Op::In, {large ? "large" : "small"}},
      large ? NodeSelectorRequirement{"special", Op::In, {"optional"}} : NodeSelectorRequirement{"special", Op::DoesNotExist, {}},
      {"integer", Op::In, {std::to_string(cpu)}},
  };
  return it;
}

std::vector<Offering> default_offerings(double price) {  // fake/instancetype.go:61-69
  return {{"spot", "test-zone-1", price, true},      {"spot", "test-zone-2", price, true},
          {"on-demand", "test-zone-1", price, true}, {"on-demand", "test-zone-2", price, true},
          {"on-demand", "test-zone-3", price, true}};
}

void add_incrementing_types(Problem& P, int total) {  // InstanceTypes(total) :151-164
  for (int i = 0; i < total; ++i) {
    int64_t cpu = i + 1, mem = (i + 1) * 2, pods = (i + 1) * 10;
    P.instance_types.push_back(make_type("fake-it-" + std::to_string(i), cpu, mem, pods, "amd64",
                                         {"darwin", "linux", "windows"}, default_offerings(price_from(cpu * 1000, mem * kGi))));
  }
}

// InstanceTypesAssorted :109-143, first `total` in generation order. all_offerings: C4/C5 variant where every
// type is offered in all 3 zones x {spot,on-demand}.
void add_assorted_types(Problem& P, int total, bool all_offerings, int64_t pods) {
  const int cpus[] = {1, 2, 4, 8, 16, 32, 64};
  const int mems[] = {1, 2, 4, 8, 16, 32, 64, 128};
  const char* zones[] = {"test-zone-1", "test-zone-2", "test-zone-3"};
  const char* cts[] = {"spot", "on-demand"};
  const char* oses[] = {"linux", "windows"};
  const char* archs[] = {"amd64", "arm64"};
  int n = 0;
  for (int cpu : cpus)
    for (int mem : mems)
      for (auto* zone : zones)
        for (auto* ct : cts)
          for (auto* os : oses)
            for (auto* arch : archs) {
              if (n++ >= total) return;
              char name[160];
              std::snprintf(name, sizeof name, "%d-cpu-%d-mem-%s-%s-%s-%s", cpu, mem, arch, os, zone, ct);
              double price = price_from(cpu * 1000LL, mem * kGi);
              std::vector<Offering> ofs;
              if (all_offerings) {
                for (auto* z : zones)
                  for (auto* c : cts) ofs.push_back({c, z, price, true});
              } else {
                ofs.push_back({ct, zone, price, true});
              }
              P.instance_types.push_back(make_type(name, cpu, mem, pods, arch, {os}, ofs));
            }
}

const int64_t kCpu[] = {100, 250, 500, 1000, 1500};
const int64_t kMem[] = {100, 256, 512, 1024, 2048, 4096};

Pod base_pod(int index, int64_t cpu, int64_t mem_mi) {
  Pod p;
  char buf[32];
  std::snprintf(buf, sizeof buf, "pod-%07d", index);
  p.name = buf;
  p.uid = buf;
  Container c;
  c.requests = {{"cpu", cpu}, {"memory", mem_mi * kMi}};
  p.containers.push_back(c);
  return p;
}

LabelSelector match_label(const std::string& k, const std::string& v) {
  LabelSelector s;
  s.is_nil = false;
  s.match_labels[k] = v;
  return s;
}

}  // namespace

Problem* synth_problem(int config, int64_t n_pods, int64_t n_types, uint64_t seed, int64_t n_nodes) {
  auto prob = std::make_unique<Problem>();
  Problem& P = *prob;
  P.extra_well_known_labels = {"size", "special", "integer"};  // fake/instancetype.go:40-46
  SplitMix64 rng(seed);
  auto all_types = [&](Provisioner& pr) {
    for (size_t i = 0; i < P.instance_types.size(); ++i) pr.instance_types.push_back((int32_t)i);
  };
  switch (config) {
    case 1:
    case 2: {
      add_incrementing_types(P, (int)n_types);
      Provisioner pr;
      pr.name = "default";
      pr.has_limits = true;  // benchmark: Limits: map{} (non-nil, empty) scheduling_benchmark_test.go:117
      all_types(pr);
      P.provisioners.push_back(pr);
      for (int64_t i = 0; i < n_pods; ++i) {
        int64_t cpu = kCpu[rng.below(5)], mem = kMem[rng.below(6)];
        P.pods.push_back(base_pod((int)i, cpu, mem));
      }
      break;
    }
    case 3: {
      add_assorted_types(P, (int)n_types, false, 5);
      Provisioner a, b, c;
      a.name = "general"; a.weight = 30;
      b.name = "batch"; b.weight = 20;
      b.taints = {{"dedicated", "batch", "NoSchedule"}};
      b.labels = {{"tier", "batch"}};
      c.name = "accel"; c.weight = 10;
      c.taints = {{"gpu", "true", "NoSchedule"}};
      c.requirements = {{"kubernetes.io/arch", Op::In, {"arm64"}}};
      all_types(a); all_types(b); all_types(c);
      P.provisioners = {a, b, c};
      const int64_t per = 250;
      int64_t deployments = (n_pods + per - 1) / per;
      int index = 0;
      for (int64_t d = 0; d < deployments; ++d) {
        int64_t cpu = kCpu[rng.below(5)], mem = kMem[rng.below(6)];
        Labels sel;
        if (rng.unit() < 0.6) {
          switch (rng.below(5)) {
            case 0: { const char* z[] = {"test-zone-1", "test-zone-2", "test-zone-3"}; sel["topology.kubernetes.io/zone"] = z[rng.below(3)]; break; }
            case 1: { const char* a2[] = {"amd64", "arm64"}; sel["kubernetes.io/arch"] = a2[rng.below(2)]; break; }
            case 2: { const char* o[] = {"linux", "windows"}; sel["kubernetes.io/os"] = o[rng.below(2)]; break; }
            case 3: { const char* ct[] = {"spot", "on-demand"}; sel["karpenter.sh/capacity-type"] = ct[rng.below(2)]; break; }
            default: sel["tier"] = "batch";
          }
        }
        bool tol_b = rng.unit() < 0.3, tol_c = rng.unit() < 0.1;
        for (int64_t k = 0; k < per && index < n_pods; ++k, ++index) {
          Pod p = base_pod(index, cpu, mem);
          p.labels = {{"app", "d" + std::to_string(d)}};
          p.node_selector = sel;
          if (tol_b) p.tolerations.push_back({"dedicated", "Equal", "batch", "NoSchedule"});
          if (tol_c) p.tolerations.push_back({"gpu", "Exists", "", "NoSchedule"});
          P.pods.push_back(std::move(p));
        }
      }
      break;
    }
    case 4: {
      add_assorted_types(P, (int)n_types, true, 5);
      Provisioner pr;
      pr.name = "default";
      pr.requirements = {{"topology.kubernetes.io/zone", Op::In, {"test-zone-1", "test-zone-2", "test-zone-3"}}};
      all_types(pr);
      P.provisioners.push_back(pr);
      const int64_t per = 250;
      int64_t deployments = (n_pods + per - 1) / per;
      int index = 0;
      for (int64_t d = 0; d < deployments; ++d) {
        int64_t cpu = kCpu[rng.below(5)], mem = kMem[rng.below(6)];
        std::string app = "d" + std::to_string(d);
        bool anti = rng.unit() < 0.25, zone_spread = rng.unit() < 0.5, host_spread = rng.unit() < 0.25;
        const int32_t skews[] = {1, 2, 4};
        int32_t host_skew = skews[rng.below(3)];
        for (int64_t k = 0; k < per && index < n_pods; ++k, ++index) {
          Pod p = base_pod(index, cpu, mem);
          p.labels = {{"app", app}};
          if (anti) {
            PodAffinityTerm t;
            t.selector = match_label("app", app);
            t.topology_key = "kubernetes.io/hostname";
            p.pod_anti_affinity_required.push_back(t);
          }
          if (zone_spread) p.topology_spread.push_back({1, "topology.kubernetes.io/zone", false, match_label("app", app)});
          if (host_spread) p.topology_spread.push_back({host_skew, "kubernetes.io/hostname", false, match_label("app", app)});
          P.pods.push_back(std::move(p));
        }
      }
      break;
    }
    case 5: {
      // Deviation from SURVEY 8d (documented in DESIGN.md): instance types get pods=110 instead of the fake
      // generator's default 5, otherwise 10 bound pods per node could never fit anywhere and every probe is trivial.
      add_assorted_types(P, (int)n_types, true, 110);
      Provisioner pr;
      pr.name = "default";
      all_types(pr);
      P.provisioners.push_back(pr);
      std::vector<int> big;
      for (size_t i = 0; i < P.instance_types.size(); ++i)
        if (P.instance_types[i].capacity.at("cpu") >= 16000) big.push_back((int)i);
      if (big.empty()) throw std::runtime_error("config 5 needs instance types with >=16 cpu (n_types too small)");
      const char* zones[] = {"test-zone-1", "test-zone-2", "test-zone-3"};
      int index = 0;
      for (int64_t n = 0; n < n_nodes; ++n) {
        const InstanceType& it = P.instance_types[big[rng.below(big.size())]];
        StateNode node;
        char buf[32];
        std::snprintf(buf, sizeof buf, "node-%05d", (int)n);
        node.name = buf;
        const char* zone = zones[rng.below(3)];
        bool spot = rng.below(2) == 0;
        node.labels = {{"karpenter.sh/provisioner-name", "default"},
                       {"karpenter.sh/initialized", "true"},
                       {"node.kubernetes.io/instance-type", it.name},
                       {"topology.kubernetes.io/zone", zone},
                       {"karpenter.sh/capacity-type", spot ? "spot" : "on-demand"},
                       {"kubernetes.io/hostname", buf},
                       {"kubernetes.io/arch", it.requirements[1].values[0]},
                       {"kubernetes.io/os", it.requirements[2].values[0]}};
        node.capacity = it.capacity;
        node.allocatable = it.capacity;
        node.allocatable["cpu"] -= 100;
        node.allocatable["memory"] -= 10 * kMi;
        int64_t per_node = n_nodes > 0 ? (n_pods / n_nodes) : 0;
        for (int64_t k = 0; k < per_node; ++k, ++index) {
          Pod p = base_pod(index, kCpu[rng.below(5)], kMem[rng.below(6)]);
          p.node_name = node.name;
          node.pods.push_back(std::move(p));
        }
        node.candidate = true;
        node.disruption_cost = (double)node.pods.size();  // helpers.go:159-165 with default eviction cost 1.0
        P.nodes.push_back(std::move(node));
      }
      P.simulation_mode = true;
      break;
    }
    default:
      throw std::runtime_error("unknown synthetic config");
  }
  return prob.release();
}

}  // namespace kmodel
