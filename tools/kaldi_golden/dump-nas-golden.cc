// dump-nas-golden: golden vectors of the TDNN-F_NAS components and of the chain denominator, written by the REFERENCE
// ITSELF.  This file is not part of the library and is not built by it: it is for a maintainer who has a Kaldi tree with the
// TDNN-F_NAS patches applied (the one thing this repository's build container lacks), so that the CPU oracle here -- and
// through it the CUDA path -- can be pinned against the reference's own arithmetic instead of a restatement of it.
// It has never been compiled (no Kaldi here): expect to fix an include or two.
//
//   cp dump-nas-golden.cc $KALDI/src/nnet3bin/ ; add dump-nas-golden to BINFILES in $KALDI/src/nnet3bin/Makefile ; make
//   python tools/kaldi_golden/make_inputs.py /tmp/nas_golden_in          # in this repository: den.fst + seeds
//   $KALDI/src/nnet3bin/dump-nas-golden /tmp/nas_golden_in /path/to/repo/tests/golden/kaldi
//   python -m pytest tests/test_kaldi_golden.py -q                        # in this repository: oracle vs the dumped vectors
//
// A CPU build of Kaldi is enough (CuMatrix falls back to the host); every case is deterministic (no Gumbel noise, no
// uniform sampling: those draw from Kaldi's global RNG and are covered here by replayed draws instead).
// Layout written: <out>/<case>/case.txt (type, config line, sizes), params.txt (Vectorize), indexes.txt
// (PrecomputedIndexes::Write, text), and per step k: in.k.txt out.k.txt memo.k.txt out_deriv.k.txt out_deriv_after.k.txt
// in_deriv.k.txt delta.k.txt (Vectorize of to_update after Backprop, to_update zeroed before every step).
#include <cmath>
#include <fstream>
#include <set>
#include <sstream>

#include "base/kaldi-common.h"
#include "chain/chain-den-graph.h"
#include "chain/chain-denominator.h"
#include "chain/chain-training.h"
#include "fstext/fstext-lib.h"
#include "nnet3/nnet-component-itf.h"
#include "nnet3/nnet-convolutional-component.h"
#include "nnet3/nnet-normalize-component.h"
#include "nnet3/nnet-simple-component.h"
#include "util/common-utils.h"

namespace kaldi {
namespace nnet3 {

// values in (-amp, amp) that depend on (row, col, seed) only: no RNG, same on every platform
static void Fill(CuMatrixBase<BaseFloat> *m, int32 seed, BaseFloat amp) {
  Matrix<BaseFloat> h(m->NumRows(), m->NumCols());
  for (int32 r = 0; r < h.NumRows(); r++)
    for (int32 c = 0; c < h.NumCols(); c++)
      h(r, c) = amp * std::sin(0.7368 * r + 1.2345 * c + 0.618 * seed + 0.05 * r * c);
  m->CopyFromMat(h);
}
static void FillVec(VectorBase<BaseFloat> *v, int32 seed, BaseFloat amp) {
  for (int32 i = 0; i < v->Dim(); i++) (*v)(i) = amp * std::sin(0.4321 * i + 0.618 * seed);
}
// text form with 9 significant digits (the default stream precision of 6 would round the inputs the checker re-uses)
static void WriteMat(const std::string &dir, const std::string &name, int32 step, const CuMatrixBase<BaseFloat> &m) {
  std::ostringstream path;
  path << dir << "/" << name << "." << step << ".txt";
  std::ofstream os(path.str().c_str());
  os.precision(9);
  Matrix<BaseFloat>(m).Write(os, false);
}
static void WriteVec(const std::string &path, const VectorBase<BaseFloat> &v) {
  std::ofstream os(path.c_str());
  os.precision(9);
  Vector<BaseFloat>(v).Write(os, false);
}
static void MakeDir(const std::string &dir) {
  if (system(("mkdir -p " + dir).c_str()) != 0) KALDI_ERR << "cannot create " << dir;
}

// One component through `steps` minibatches: Propagate, Backprop into a same-typed `to_update` (zeroed before each step,
// natural-gradient state carried over: that is what delta_nnet_ is in NnetChainTrainer).
static void DumpComponent(const std::string &out_root, const std::string &case_name, const std::string &type,
                          const std::string &config, int32 num_seq, int32 t_out_begin, int32 t_out_end, int32 t_step,
                          int32 steps, BaseFloat alpha_amp) {
  std::string dir = out_root + "/" + case_name;
  MakeDir(dir);
  Component *comp = Component::NewComponentOfType(type);
  KALDI_ASSERT(comp != NULL);
  ConfigLine cfl;
  if (!cfl.ParseLine(type + " " + config)) KALDI_ERR << "bad config " << config;
  comp->InitFromConfig(&cfl);
  UpdatableComponent *uc = dynamic_cast<UpdatableComponent*>(comp);
  if (uc != NULL) {
    Vector<BaseFloat> params(uc->NumParameters());
    uc->Vectorize(&params);  // keep the scale InitFromConfig chose, replace the values
    BaseFloat rms = std::sqrt(VecVec(params, params) / params.Dim());
    FillVec(&params, 11, rms > 0 ? 1.7 * rms : 0.1);
    TdnnDARTSV3Component *darts = dynamic_cast<TdnnDARTSV3Component*>(comp);
    if (darts != NULL) {
      // Vectorize order: linear_params_ rows, then bias_params_ = [alpha (one per offset), bias (output-dim)]
      // P = D_out * D_in * n + n + D_out  =>  n = (P - D_out) / (D_out * D_in + 1)
      int32 n = (params.Dim() - darts->OutputDim()) / (darts->OutputDim() * darts->InputDim() + 1),
            linear = darts->OutputDim() * darts->InputDim() * n;
      for (int32 i = 0; i < n; i++) params(linear + i) = alpha_amp * std::sin(1.3 * i + 0.4);
    }
    uc->UnVectorize(params);
    WriteVec(dir + "/params.txt", params);
  }
  // the regular grid nnet3 hands to the component: t-major, n fastest
  std::vector<Index> in_idx, out_idx;
  ComponentPrecomputedIndexes *pi = NULL;
  int32 in_rows, out_rows;
  if (!(comp->Properties() & kSimpleComponent)) {
    for (int32 t = t_out_begin; t < t_out_end; t += t_step)
      for (int32 n = 0; n < num_seq; n++) out_idx.push_back(Index(n, t));
    MiscComputationInfo misc;
    std::set<Index> needed;
    for (size_t i = 0; i < out_idx.size(); i++) {
      std::vector<Index> req;
      comp->GetInputIndexes(misc, out_idx[i], &req);
      needed.insert(req.begin(), req.end());
    }
    in_idx.assign(needed.begin(), needed.end());  // Index::operator< sorts by t, then x, then n
    comp->ReorderIndexes(&in_idx, &out_idx);
    pi = comp->PrecomputeIndexes(misc, in_idx, out_idx, true);
    std::ofstream os((dir + "/indexes.txt").c_str());
    pi->Write(os, false);
    in_rows = in_idx.size();
    out_rows = out_idx.size();
  } else {
    in_rows = out_rows = num_seq * ((t_out_end - t_out_begin + t_step - 1) / t_step);
  }
  {
    std::ofstream os((dir + "/case.txt").c_str());
    os << "type " << type << "\nconfig " << config << "\nnum_sequences " << num_seq << "\nin_rows " << in_rows << "\nout_rows "
       << out_rows << "\nsteps " << steps << "\nproperties " << comp->Properties() << "\ninfo " << comp->Info() << "\n";
  }
  Component *to_update = comp->Copy();
  UpdatableComponent *uu = dynamic_cast<UpdatableComponent*>(to_update);
  for (int32 k = 0; k < steps; k++) {
    CuMatrix<BaseFloat> in(in_rows, comp->InputDim()), out(out_rows, comp->OutputDim());
    Fill(&in, 100 + k, 1.0);
    if (comp->Properties() & kPropagateAdds) Fill(&out, 300 + k, 0.25);  // the caller's running sum
    WriteMat(dir, "out_before", k, out);
    void *memo = comp->Propagate(pi, in, &out);
    WriteMat(dir, "in", k, in);
    WriteMat(dir, "out", k, out);
    if (memo != NULL && dynamic_cast<TdnnDARTSV3Component*>(comp) != NULL) {
      std::ostringstream path;
      path << dir << "/memo." << k << ".txt";
      WriteVec(path.str(), Vector<BaseFloat>(*static_cast<CuVector<BaseFloat>*>(memo)));
    }
    CuMatrix<BaseFloat> out_deriv(out_rows, comp->OutputDim()), in_deriv(in_rows, comp->InputDim());
    Fill(&out_deriv, 200 + k, 1.0 / out_rows);
    if (comp->Properties() & kBackpropAdds) Fill(&in_deriv, 400 + k, 0.125);
    WriteMat(dir, "out_deriv", k, out_deriv);
    WriteMat(dir, "in_deriv_before", k, in_deriv);
    if (uu != NULL) uu->Scale(0.0);
    comp->Backprop("golden", pi, in, out, out_deriv, memo, to_update, &in_deriv);
    WriteMat(dir, "out_deriv_after", k, out_deriv);  // {Gumbel}SoftmaxFlops writes the FLOPs penalty into it
    WriteMat(dir, "in_deriv", k, in_deriv);
    if (uu != NULL) {
      Vector<BaseFloat> delta(uu->NumParameters());
      uu->Vectorize(&delta);
      std::ostringstream path;
      path << dir << "/delta." << k << ".txt";
      WriteVec(path.str(), delta);
    }
    if (memo != NULL) comp->DeleteMemo(memo);
  }
  delete pi;
  delete to_update;
  delete comp;
}

static void DumpDenominator(const std::string &in_root, const std::string &out_root, const std::string &case_name,
                            int32 num_pdfs, int32 num_seq, int32 frames, BaseFloat leaky) {
  std::string dir = out_root + "/" + case_name;
  MakeDir(dir);
  fst::StdVectorFst den_fst;
  fst::ReadFstKaldi(in_root + "/" + case_name + ".den.fst", &den_fst);
  chain::DenominatorGraph graph(den_fst, num_pdfs);
  chain::ChainTrainingOptions opts;
  opts.leaky_hmm_coefficient = leaky;
  CuMatrix<BaseFloat> nnet_output(num_seq * frames, num_pdfs), deriv(num_seq * frames, num_pdfs);
  Fill(&nnet_output, 7, 2.0);
  chain::DenominatorComputation den(opts, graph, num_seq, nnet_output);
  BaseFloat logprob = den.Forward();
  bool ok = den.Backward(-1.0, &deriv);
  WriteMat(dir, "nnet_output", 0, nnet_output);
  WriteMat(dir, "deriv", 0, deriv);
  WriteVec(dir + "/initial_probs.txt", Vector<BaseFloat>(graph.InitialProbs()));
  std::ofstream os((dir + "/case.txt").c_str());
  os.precision(9);
  os << "type Denominator\nnum_pdfs " << num_pdfs << "\nnum_sequences " << num_seq << "\nframes " << frames << "\nleaky " << leaky
     << "\nnum_states " << graph.NumStates() << "\nlogprob " << logprob << "\nok " << (ok ? 1 : 0) << "\n";
}

}  // namespace nnet3
}  // namespace kaldi

int main(int argc, char *argv[]) {
  using namespace kaldi;
  using namespace kaldi::nnet3;
  try {
    const char *usage = "Golden vectors of the TDNN-F_NAS components.\nUsage: dump-nas-golden <inputs-dir> <out-dir>\n";
    ParseOptions po(usage);
    po.Read(argc, argv);
    if (po.NumArgs() != 2) {
      po.PrintUsage();
      exit(1);
    }
    std::string in_root = po.GetArg(1), out_root = po.GetArg(2);
    const std::string softmax = "use-gumbel=false use-entropy=false free-select=false uniform-sample=false";
    // the two orientations of the supernet's blocks (generate_config.py: offsets -6..0 into the bottleneck, 0..6 out of it)
    DumpComponent(out_root, "darts_linear_softmax", "TdnnDARTSV3Component",
                  "input-dim=96 output-dim=40 time-offsets=-6,-5,-4,-3,-2,-1,0 use-bias=true learning-rate=0.02 rank-in=20 rank-out=30 "
                  "update-alpha=true update-theta=true " + softmax, 8, 0, 24, 1, 6, 1.0);
    DumpComponent(out_root, "darts_affine_softmax", "TdnnDARTSV3Component",
                  "input-dim=40 output-dim=96 time-offsets=0,1,2,3,4,5,6 use-bias=true learning-rate=0.02 rank-in=20 rank-out=30 "
                  "update-alpha=true update-theta=true " + softmax, 8, 0, 24, 1, 6, 1.0);
    DumpComponent(out_root, "darts_affine_alpha_only", "TdnnDARTSV3Component",
                  "input-dim=40 output-dim=96 time-offsets=0,1,2,3,4,5,6 use-bias=true learning-rate=0.02 rank-in=20 rank-out=30 "
                  "update-alpha=true update-theta=false Temp-Proportion=0.5 " + softmax, 8, 0, 24, 3, 6, 0.5);
    DumpComponent(out_root, "softmax_flops", "SoftmaxFlopsComponent", "dim=8 scale=0.001", 16, 0, 12, 1, 1, 0.0);
    DumpComponent(out_root, "softmax_flops_eta_0p1", "SoftmaxFlopsComponent", "dim=8 scale=0.1", 16, 0, 12, 1, 1, 0.0);
    DumpComponent(out_root, "copyn_1_to_30", "CopyNComponent", "input-dim=1 output-dim=30 scale=1.0", 16, 0, 12, 1, 1, 0.0);
    DumpDenominator(in_root, out_root, "den_small", 37, 6, 9, 0.1);
    DumpDenominator(in_root, out_root, "den_medium", 211, 16, 17, 0.1);
    return 0;
  } catch (const std::exception &e) {
    std::cerr << e.what();
    return -1;
  }
}
