// ORACLE — TEST INFRASTRUCTURE ONLY (see requirements.h header).
#pragma once
#include <string>
#include <vector>

#include "../karpenter-core_b200/host/model.h"
#include "requirements.h"

namespace oracle {

// Result-neutral shortcuts for full-size runs (see oracle.cc header: FAST MODE). Default off = literal restatement.
void set_fast(bool on);

// pkg/utils/resources/resources.go:25-33
kmodel::ResourceList requests_for_pods(const std::vector<const kmodel::Pod*>& pods);

// One Scheduler.Solve (scheduler.go:96) over the pods/nodes a reconcile would hand it:
// candidates == {}  -> Provisioner.Reconcile's view (provisioner.go:119-144)
// candidates != {}  -> simulateScheduling's view (deprovisioning/helpers.go:42-93); Result.assign
// is indexed over [P.pods..., candidates' pods in order..., deleting nodes' pods].
void solve(const kmodel::Problem& P, const std::vector<int>& candidates, kmodel::Result& out);

struct ConsolidationResult {
  int action = 0;         // 0 do nothing, 1 delete, 2 replace (consolidation.go actions)
  int nodes_removed = 0;  // largest feasible prefix of the cost-ordered candidates
  std::vector<int> replacement_options;  // instance type indices after price filters
  std::vector<int> candidate_order;      // node indices in disruption-cost order
  std::vector<int> probes, probe_actions;  // binary-search trace (prefix length, action)
  int simulations = 0;
  int single_node = -1;   // consolidate_single: Problem.nodes index of the node the command removes
  std::string error;
};
// MultiNodeConsolidation.firstNNodeConsolidationOption (multinodeconsolidation.go:74-114)
// only_count > 0: a single computeConsolidation over that many candidates instead of the search
void consolidate(const kmodel::Problem& P, ConsolidationResult& out, int only_count = 0);
// SingleNodeConsolidation.ComputeCommand (singlenodeconsolidation.go:43-84): candidates in disruption-cost order, the first one
// whose computeConsolidation yields delete / replace wins. probes = positions tried. only_candidate >= 0: just that position.
void consolidate_single(const kmodel::Problem& P, ConsolidationResult& out, int only_candidate = -1);
// candidateNodes + sortAndFilterCandidates (helpers.go:171-249,339-366, consolidation.go:85-118, pdblimits.go:55-68)
void rank_candidates(const kmodel::Problem& P, std::vector<int>* order, std::vector<double>* cost);
double pod_eviction_cost(const kmodel::Pod& p);  // helpers.go:125-146

}  // namespace oracle
