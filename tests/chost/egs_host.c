/* A plain C99 host of the C ABI (no Python, no torch, no C++): reads an archive of NnetChainExample and a binary den.fst
 * from disk, forms one minibatch and the graphs, and prints what it got -- what a Kaldi-side data loader would do before
 * tdnnf_num_graph_update / tdnnf_chain_objf_and_deriv.  tests/test_c_host.py compiles it with
 * gcc against libtdnnf_nas_b200.so, runs it and compares the printout with the ctypes binding's view of the same files. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tdnnf_nas_b200.h"

#define OK_OR_DIE(call)                                                  \
  do {                                                                   \
    if ((call) != TDNNF_OK) {                                            \
      fprintf(stderr, "%s: %s\n", #call, tdnnf_last_error());            \
      return 1;                                                          \
    }                                                                    \
  } while (0)

static char* slurp(const char* path, uint64_t* len) {
  FILE* f = fopen(path, "rb");
  char* buf;
  long n;
  if (!f) return NULL;
  fseek(f, 0, SEEK_END);
  n = ftell(f);
  fseek(f, 0, SEEK_SET);
  buf = (char*)malloc((size_t)n + 1);
  if (buf && fread(buf, 1, (size_t)n, f) != (size_t)n) {
    free(buf);
    buf = NULL;
  }
  fclose(f);
  *len = (uint64_t)n;
  return buf;
}

int main(int argc, char** argv) {
  uint64_t ark_len = 0, den_len = 0;
  char *ark, *den;
  tdnnf_chain_egs* egs = NULL;
  tdnnf_host_graph* hg = NULL;
  tdnnf_host_num_graph* hn = NULL;
  tdnnf_ctx* ctx = NULL;
  int num_pdfs, count = 0, T_in = 0, S = 0, D = 0, first_t = 0, T = 0, i, states = 0, pdfs = 0, trans = 0, nseq = 0, narcs = 0;
  const int32_t *so, *fr, *br, *ap, *as;
  const float *lp, *fl;
  float weight = 0.f, *x, *dw;
  double sum = 0.0, dwsum = 0.0;
  if (argc != 4) {
    fprintf(stderr, "usage: %s egs.ark den.fst num_pdfs\n", argv[0]);
    return 2;
  }
  num_pdfs = atoi(argv[3]);
  ark = slurp(argv[1], &ark_len);
  den = slurp(argv[2], &den_len);
  if (!ark || !den) {
    fprintf(stderr, "cannot read the input files\n");
    return 2;
  }
  printf("abi %d\n", tdnnf_abi_version());
  OK_OR_DIE(tdnnf_chain_egs_read_ark(ark, ark_len, 0, &egs));
  OK_OR_DIE(tdnnf_chain_egs_count(egs, &count));
  OK_OR_DIE(tdnnf_chain_egs_merge_input(egs, 0, count, "input", NULL, 0, &T_in, &S, &D, &first_t));
  x = (float*)malloc(sizeof(float) * (size_t)T_in * S * D);
  OK_OR_DIE(tdnnf_chain_egs_merge_input(egs, 0, count, "input", x, (int64_t)T_in * S * D, &T_in, &S, &D, &first_t));
  for (i = 0; i < T_in * S * D; ++i) sum += (double)x[i] * ((i % 7) + 1);
  OK_OR_DIE(tdnnf_chain_egs_merge_supervision(egs, 0, count, "output", num_pdfs, NULL, 0, &S, &T, &weight, NULL));
  dw = (float*)malloc(sizeof(float) * (size_t)T * S);
  OK_OR_DIE(tdnnf_chain_egs_merge_supervision(egs, 0, count, "output", num_pdfs, dw, T * S, &S, &T, &weight, &hn));
  for (i = 0; i < T * S; ++i) dwsum += dw[i];
  OK_OR_DIE(tdnnf_host_num_graph_arrays(hn, &nseq, &narcs, &so, &fr, &br, &lp, &ap, &as, &fl));
  OK_OR_DIE(tdnnf_den_graph_parse_fst_binary(den, den_len, num_pdfs, &hg));
  OK_OR_DIE(tdnnf_host_graph_dims(hg, &states, &pdfs, &trans));
  printf("examples %d\ninput %d %d %d first_t %d checksum %.6f\n", count, T_in, S, D, first_t, sum);
  printf("supervision seqs %d frames %d weight %.3f deriv_weight_sum %.6f\n", S, T, weight, dwsum);
  printf("numerator seqs %d arcs %d states %d\n", nseq, narcs, so[nseq]);
  printf("denominator states %d pdfs %d transitions %d\n", states, pdfs, trans);
  /* the device side: a context exists only on an sm_100 GPU; anywhere else the library must say so and do nothing */
  if (tdnnf_ctx_create(0, &ctx) == TDNNF_OK) {
    tdnnf_den_graph* dg = NULL;
    tdnnf_num_graph* ng = NULL;
    OK_OR_DIE(tdnnf_den_graph_create_from_host(ctx, hg, &dg));
    OK_OR_DIE(tdnnf_num_graph_create_from_host(ctx, hn, &ng));
    printf("device graphs created\n");
    tdnnf_num_graph_destroy(ng);
    tdnnf_den_graph_destroy(dg);
    tdnnf_ctx_destroy(ctx);
  } else {
    printf("no device: %s\n", tdnnf_last_error());
  }
  tdnnf_host_graph_free(hg);
  tdnnf_host_num_graph_free(hn);
  tdnnf_chain_egs_free(egs);
  free(x);
  free(dw);
  free(ark);
  free(den);
  return 0;
}
