// Internal declarations shared by the host-only readers (chain_io.cc: FSM text; egs_io.cc: Kaldi binary / text objects).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "context.h"

struct tdnnf_host_graph;
struct tdnnf_host_num_graph;

namespace tdnnf {

// An acceptor in the chain convention: ilabel = pdf-id + 1, tropical weights (-log probability).
struct FsmArc {
  int src, dst, ilabel;
  float weight;
};
struct Fsm {
  int start = -1, num_states = 0;
  std::vector<FsmArc> arcs;
  std::map<int, float> finals;
};

// AT&T FSM text (`fstprint`) -> Fsm; err names the offending line.
int parse_fsm(const char* text, size_t len, Fsm* f, std::string* err);
// DenominatorGraph::SetTransitions + SetInitialProbs (kaldi: chain/chain-den-graph.cc) over an Fsm.
int build_host_den_graph(const Fsm& f, int num_pdfs, tdnnf_host_graph** out);
// One Fsm per sequence -> the arrays of tdnnf_num_graph_create.
int build_host_num_graph(const std::vector<Fsm>& fsms, int num_pdfs, tdnnf_host_num_graph** out);

}  // namespace tdnnf
