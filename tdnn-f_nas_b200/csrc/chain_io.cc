// Readers for the two graph inputs of the chain objective, host only (no CUDA): the data formats on the near side of the
// path.
//   * den.fst in AT&T FSM text form (`fstprint den.fst`): -> DenominatorGraph arrays, i.e. kaldi chain/chain-den-graph.cc
//     SetTransitions + SetInitialProbs (upstream Kaldi, not shipped with the reference: restated from SURVEY.md B.2);
//   * the per-sequence numerator FSTs of an unconstrained ("e2e") Supervision in the same text form -> the arrays
//     tdnnf_num_graph_create takes (kaldi chain/chain-generic-numerator.cc reads them from Supervision::e2e_fsts).
// FSM text: one arc per line "src dst ilabel olabel [weight]", one line per final state "state [weight]"; the source of
// the first line is the start state; weights are tropical (-log probability, default 0); ilabel = pdf-id + 1.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "chain_io.h"

using namespace tdnnf;

namespace tdnnf {

int parse_fsm(const char* text, size_t len, Fsm* f, std::string* err) {
  std::istringstream in(std::string(text, len));
  std::string line;
  int lineno = 0;
  while (std::getline(in, line)) {
    ++lineno;
    std::istringstream ls(line);
    std::vector<std::string> tok;
    std::string t;
    while (ls >> t) tok.push_back(t);
    if (tok.empty()) continue;
    char* end = nullptr;
    auto to_int = [&](const std::string& s, int* v) {
      const long x = strtol(s.c_str(), &end, 10);
      if (*end != '\0' || x < 0 || x > 0x7fffffff) return false;
      *v = (int)x;
      return true;
    };
    auto to_float = [&](const std::string& s, float* v) {
      if (s == "Infinity" || s == "inf") {
        *v = INFINITY;
        return true;
      }
      *v = strtof(s.c_str(), &end);
      return *end == '\0';
    };
    bool ok = true;
    if (tok.size() <= 2) {  // final state
      int s = 0;
      float w = 0.f;
      ok = to_int(tok[0], &s) && (tok.size() == 1 || to_float(tok[1], &w));
      if (ok) {
        f->finals[s] = w;
        if (f->start < 0) f->start = s;
        f->num_states = std::max(f->num_states, s + 1);
      }
    } else if (tok.size() == 4 || tok.size() == 5) {
      FsmArc a;
      int olabel = 0;
      a.weight = 0.f;
      ok = to_int(tok[0], &a.src) && to_int(tok[1], &a.dst) && to_int(tok[2], &a.ilabel) && to_int(tok[3], &olabel) &&
           (tok.size() == 4 || to_float(tok[4], &a.weight));
      if (ok) {
        if (f->start < 0) f->start = a.src;
        f->num_states = std::max(f->num_states, std::max(a.src, a.dst) + 1);
        f->arcs.push_back(a);
      }
    } else {
      ok = false;
    }
    if (!ok) {
      *err = "FSM text line " + std::to_string(lineno) + " is not 'src dst ilabel olabel [weight]' or 'state [weight]': " + line;
      return TDNNF_ERR_INVALID;
    }
  }
  if (f->start < 0) {
    *err = "empty FST";
    return TDNNF_ERR_INVALID;
  }
  return TDNNF_OK;
}

}  // namespace tdnnf

struct tdnnf_host_graph {
  int num_states = 0, num_pdfs = 0, num_transitions = 0;  // num_transitions = forward list + backward list
  std::vector<int32_t> fwd_ranges, bwd_ranges, pdf, state;
  std::vector<float> prob, init;
};

struct tdnnf_host_num_graph {
  int num_seqs = 0, num_arcs = 0;  // arcs per direction
  std::vector<int32_t> state_offsets, fwd_ranges, bwd_ranges, pdf, state;
  std::vector<float> logprob, final_logprob;
};

extern "C" int tdnnf_den_graph_parse_fst_text(const char* text, uint64_t len, int num_pdfs, tdnnf_host_graph** out) {
  TDNNF_REQUIRE(text && out && num_pdfs > 0, "bad argument");
  Fsm f;
  std::string err;
  int rc = parse_fsm(text, (size_t)len, &f, &err);
  if (rc) return fail(rc, "den.fst: " + err);
  return build_host_den_graph(f, num_pdfs, out);
}

int tdnnf::build_host_den_graph(const Fsm& f, int num_pdfs, tdnnf_host_graph** out) {
  const int N = f.num_states;
  const size_t A = f.arcs.size();
  TDNNF_REQUIRE(A > 0, "den.fst has no arcs");
  tdnnf_host_graph* g = new tdnnf_host_graph();
  g->num_states = N;
  g->num_pdfs = num_pdfs;
  g->num_transitions = (int)(2 * A);
  // SetTransitions: forward list grouped by source state (arc order kept), backward list grouped by destination
  std::vector<int> out_cnt(N, 0), in_cnt(N, 0);
  for (const FsmArc& a : f.arcs) {
    if (a.ilabel < 1 || a.ilabel > num_pdfs) {
      delete g;
      return fail(TDNNF_ERR_INVALID, "den.fst: ilabel " + std::to_string(a.ilabel) + " is not a pdf-id + 1 in [1, num_pdfs]");
    }
    out_cnt[a.src]++;
    in_cnt[a.dst]++;
  }
  std::vector<int> fb(N + 1, 0), bb(N + 1, 0);
  for (int s = 0; s < N; ++s) {
    fb[s + 1] = fb[s] + out_cnt[s];
    bb[s + 1] = bb[s] + in_cnt[s];
  }
  g->prob.assign(2 * A, 0.f);
  g->pdf.assign(2 * A, 0);
  g->state.assign(2 * A, 0);
  std::vector<int> fpos(fb.begin(), fb.end() - 1), bpos(bb.begin(), bb.end() - 1);
  // arcs of one source state need not be contiguous in the text: place them by state, in file order
  std::vector<std::vector<int>> by_src(N);
  for (size_t i = 0; i < A; ++i) by_src[f.arcs[i].src].push_back((int)i);
  for (int s = 0; s < N; ++s) {
    for (int i : by_src[s]) {
      const FsmArc& a = f.arcs[i];
      const float p = std::exp(-a.weight);
      const int fi = fpos[s]++, bi = (int)A + bpos[a.dst]++;
      g->prob[fi] = p; g->pdf[fi] = a.ilabel - 1; g->state[fi] = a.dst;
      g->prob[bi] = p; g->pdf[bi] = a.ilabel - 1; g->state[bi] = s;
    }
  }
  g->fwd_ranges.resize(2 * (size_t)N);
  g->bwd_ranges.resize(2 * (size_t)N);
  for (int s = 0; s < N; ++s) {
    g->fwd_ranges[2 * s] = fb[s];
    g->fwd_ranges[2 * s + 1] = fb[s + 1];
    g->bwd_ranges[2 * s] = (int)A + bb[s];
    g->bwd_ranges[2 * s + 1] = (int)A + bb[s + 1];
  }
  // SetInitialProbs: 100 steps of the state distribution from the start state, every state's outgoing mass normalised
  // together with its final probability, the distribution renormalised after each step; initial_probs = the average.
  std::vector<double> norm(N, 0.0), cur(N, 0.0), next(N, 0.0), avg(N, 0.0);
  for (int s = 0; s < N; ++s)
    for (int i = fb[s]; i < fb[s + 1]; ++i) norm[s] += g->prob[i];
  for (const auto& kv : f.finals) norm[kv.first] += std::exp(-(double)kv.second);
  cur[f.start] = 1.0;
  const int num_iters = 100;
  for (int iter = 0; iter < num_iters; ++iter) {
    for (int s = 0; s < N; ++s) avg[s] += cur[s] / num_iters;
    for (int s = 0; s < N; ++s) {
      if (cur[s] == 0.0 || norm[s] <= 0.0) continue;
      const double p = cur[s] / norm[s];
      for (int i = fb[s]; i < fb[s + 1]; ++i) next[g->state[i]] += p * g->prob[i];
    }
    double sum = 0.0;
    for (int s = 0; s < N; ++s) sum += next[s];
    for (int s = 0; s < N; ++s) {
      cur[s] = sum > 0.0 ? next[s] / sum : 0.0;
      next[s] = 0.0;
    }
  }
  g->init.resize(N);
  for (int s = 0; s < N; ++s) g->init[s] = (float)avg[s];
  *out = g;
  return TDNNF_OK;
}

extern "C" int tdnnf_host_graph_dims(const tdnnf_host_graph* g, int* num_states, int* num_pdfs, int* num_transitions) {
  TDNNF_REQUIRE(g && num_states && num_pdfs && num_transitions, "null argument");
  *num_states = g->num_states;
  *num_pdfs = g->num_pdfs;
  *num_transitions = g->num_transitions;
  return TDNNF_OK;
}

extern "C" int tdnnf_host_graph_arrays(const tdnnf_host_graph* g, const int32_t** fwd_ranges, const int32_t** bwd_ranges,
                                       const float** prob, const int32_t** pdf, const int32_t** state, const float** initial_probs) {
  TDNNF_REQUIRE(g && fwd_ranges && bwd_ranges && prob && pdf && state && initial_probs, "null argument");
  *fwd_ranges = g->fwd_ranges.data();
  *bwd_ranges = g->bwd_ranges.data();
  *prob = g->prob.data();
  *pdf = g->pdf.data();
  *state = g->state.data();
  *initial_probs = g->init.data();
  return TDNNF_OK;
}

extern "C" int tdnnf_host_graph_free(tdnnf_host_graph* g) {
  delete g;
  return TDNNF_OK;
}

extern "C" int tdnnf_den_graph_create_from_host(tdnnf_ctx* ctx, const tdnnf_host_graph* g, tdnnf_den_graph** out) {
  TDNNF_REQUIRE(ctx && g && out, "null argument");
  return tdnnf_den_graph_create(ctx, g->num_states, g->num_pdfs, g->num_transitions, g->fwd_ranges.data(), g->bwd_ranges.data(),
                                g->prob.data(), g->pdf.data(), g->state.data(), g->init.data(), out);
}

// ---- numerator: one FSM text per sequence
extern "C" int tdnnf_num_graph_parse_fst_texts(const char* const* texts, const uint64_t* lens, int num_seqs, int num_pdfs,
                                               tdnnf_host_num_graph** out) {
  TDNNF_REQUIRE(texts && lens && out && num_seqs > 0 && num_pdfs > 0, "bad argument");
  std::vector<Fsm> fsms((size_t)num_seqs);
  for (int s = 0; s < num_seqs; ++s) {
    std::string err;
    int rc = parse_fsm(texts[s], (size_t)lens[s], &fsms[s], &err);
    if (rc) return fail(rc, "numerator FST " + std::to_string(s) + ": " + err);
  }
  return build_host_num_graph(fsms, num_pdfs, out);
}

int tdnnf::build_host_num_graph(const std::vector<Fsm>& fsms, int num_pdfs, tdnnf_host_num_graph** out) {
  const int num_seqs = (int)fsms.size();
  TDNNF_REQUIRE(out && num_seqs > 0 && num_pdfs > 0, "bad argument");
  tdnnf_host_num_graph* g = new tdnnf_host_num_graph();
  g->num_seqs = num_seqs;
  g->state_offsets.push_back(0);
  struct Arc { int src, dst, pdf; float lp; };
  std::vector<Arc> arcs;
  for (int s = 0; s < num_seqs; ++s) {
    const Fsm& f = fsms[s];
    if (f.start < 0 || f.num_states <= 0) {
      delete g;
      return fail(TDNNF_ERR_INVALID, "numerator FST " + std::to_string(s) + ": empty FST");
    }
    const int base = g->state_offsets.back();
    // the start state becomes the sequence's first state (tdnnf_num_graph_create's convention)
    auto local = [&](int st) { return st == f.start ? 0 : (st == 0 ? f.start : st); };
    for (const FsmArc& a : f.arcs) {
      if (a.ilabel < 1 || a.ilabel > num_pdfs) {
        delete g;
        return fail(TDNNF_ERR_INVALID, "numerator FST " + std::to_string(s) + ": ilabel is not a pdf-id + 1");
      }
      arcs.push_back(Arc{base + local(a.src), base + local(a.dst), a.ilabel - 1, -a.weight});
    }
    g->final_logprob.resize((size_t)base + f.num_states, -1.0e30f);
    for (const auto& kv : f.finals) g->final_logprob[(size_t)base + local(kv.first)] = std::isinf(kv.second) ? -1.0e30f : -kv.second;
    g->state_offsets.push_back(base + f.num_states);
  }
  const int N = g->state_offsets.back();
  const size_t A = arcs.size();
  g->num_arcs = (int)A;
  std::vector<int> fb(N + 1, 0), bb(N + 1, 0);
  for (const Arc& a : arcs) {
    fb[a.src + 1]++;
    bb[a.dst + 1]++;
  }
  for (int s = 0; s < N; ++s) {
    fb[s + 1] += fb[s];
    bb[s + 1] += bb[s];
  }
  g->logprob.assign(2 * A, 0.f);
  g->pdf.assign(2 * A, 0);
  g->state.assign(2 * A, 0);
  std::vector<int> fpos(fb.begin(), fb.end() - 1), bpos(bb.begin(), bb.end() - 1);
  for (const Arc& a : arcs) {  // stable within a state: file order
    const int fi = fpos[a.src]++, bi = (int)A + bpos[a.dst]++;
    g->logprob[fi] = a.lp; g->pdf[fi] = a.pdf; g->state[fi] = a.dst;
    g->logprob[bi] = a.lp; g->pdf[bi] = a.pdf; g->state[bi] = a.src;
  }
  g->fwd_ranges.resize(2 * (size_t)N);
  g->bwd_ranges.resize(2 * (size_t)N);
  for (int s = 0; s < N; ++s) {
    g->fwd_ranges[2 * s] = fb[s];
    g->fwd_ranges[2 * s + 1] = fb[s + 1];
    g->bwd_ranges[2 * s] = (int)A + bb[s];
    g->bwd_ranges[2 * s + 1] = (int)A + bb[s + 1];
  }
  *out = g;
  return TDNNF_OK;
}

extern "C" int tdnnf_host_num_graph_arrays(const tdnnf_host_num_graph* g, int* num_seqs, int* num_arcs, const int32_t** state_offsets,
                                           const int32_t** fwd_ranges, const int32_t** bwd_ranges, const float** arc_logprob,
                                           const int32_t** arc_pdf, const int32_t** arc_state, const float** final_logprob) {
  TDNNF_REQUIRE(g && num_seqs && num_arcs && state_offsets && fwd_ranges && bwd_ranges && arc_logprob && arc_pdf && arc_state &&
                    final_logprob, "null argument");
  *num_seqs = g->num_seqs;
  *num_arcs = g->num_arcs;
  *state_offsets = g->state_offsets.data();
  *fwd_ranges = g->fwd_ranges.data();
  *bwd_ranges = g->bwd_ranges.data();
  *arc_logprob = g->logprob.data();
  *arc_pdf = g->pdf.data();
  *arc_state = g->state.data();
  *final_logprob = g->final_logprob.data();
  return TDNNF_OK;
}

extern "C" int tdnnf_host_num_graph_free(tdnnf_host_num_graph* g) {
  delete g;
  return TDNNF_OK;
}

extern "C" int tdnnf_num_graph_create_from_host(tdnnf_ctx* ctx, const tdnnf_host_num_graph* g, tdnnf_num_graph** out) {
  TDNNF_REQUIRE(ctx && g && out, "null argument");
  return tdnnf_num_graph_create(ctx, g->num_seqs, g->state_offsets.data(), g->num_arcs, g->fwd_ranges.data(), g->bwd_ranges.data(),
                                g->logprob.data(), g->pdf.data(), g->state.data(), g->final_logprob.data(), out);
}
