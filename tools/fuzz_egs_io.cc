// Mutation fuzzer for the host-only readers (csrc/egs_io.cc, csrc/chain_io.cc) under AddressSanitizer / UBSan: every
// mutated archive must come back as an error code or as a structure that can be walked and merged, never as a memory error.
//   python - <<'PY'            # seed files from the test writer
//   import numpy as np; from tests import egs_ref as W; r = np.random.default_rng(1)
//   ex = [W.random_example(r, "a", coding="cm1"), W.random_example(r, "b", e2e=False, coding="sparse"), W.random_example(r, "c", num_sequences=2, coding="cm2")]
//   open("/tmp/bin.ark", "wb").write(W.ark(ex, True)); open("/tmp/txt.ark", "wb").write(W.ark(ex, False))
//   f = W.random_fst(r, 20, 29); open("/tmp/den_v.fst", "wb").write(W.fst_vector(f)); open("/tmp/den_c.fst", "wb").write(W.fst_compact_acceptor(f))
//   PY
//   g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-sanitize-recover=undefined -I include -I tdnn-f_nas_b200/csrc \
//       -I /usr/local/cuda/include tools/fuzz_egs_io.cc tdnn-f_nas_b200/csrc/egs_io.cc tdnn-f_nas_b200/csrc/chain_io.cc -o /tmp/fuzz_egs_io
//   /tmp/fuzz_egs_io /tmp/bin.ark /tmp/txt.ark /tmp/den_v.fst /tmp/den_c.fst      # 160 000 mutants, ~45 s; prints "ok N err M"
//   (FUZZ_SEED / FUZZ_ITERS in the environment: another seed, iterations per file)
// (round 2: one finding, a signed overflow in the index vector's one-byte t step, fixed; clean since: 160 000 + 600 000
// mutants over seven seed files, two seeds.)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <random>
#include "context.h"
namespace tdnnf { int fail(int code, const std::string& msg) { return code; } }
extern "C" int tdnnf_den_graph_create(tdnnf_ctx*, int, int, int, const int32_t*, const int32_t*, const float*, const int32_t*, const int32_t*, const float*, tdnnf_den_graph**) { return 1; }
extern "C" int tdnnf_num_graph_create(tdnnf_ctx*, int, const int32_t*, int, const int32_t*, const int32_t*, const float*, const int32_t*, const int32_t*, const float*, tdnnf_num_graph**) { return 1; }
int main(int argc, char** argv) {
  std::mt19937 rng(getenv("FUZZ_SEED") ? atoi(getenv("FUZZ_SEED")) : 1);
  const int iters = getenv("FUZZ_ITERS") ? atoi(getenv("FUZZ_ITERS")) : 40000;
  long ok = 0, err = 0;
  for (int a = 1; a < argc; ++a) {
    FILE* f = fopen(argv[a], "rb");
    std::vector<char> base(1 << 20);
    size_t n = fread(base.data(), 1, base.size(), f);
    fclose(f);
    base.resize(n);
    for (int it = 0; it < iters; ++it) {
      std::vector<char> b = base;
      int k = 1 + rng() % 5;
      for (int j = 0; j < k; ++j) {
        int mode = rng() % 4;
        size_t pos = rng() % b.size();
        if (mode == 0) b[pos] = (char)(rng() & 255);
        else if (mode == 1) b[pos] ^= (char)(1 << (rng() % 8));
        else if (mode == 2 && pos + 4 < b.size()) { int v = (rng() % 3 == 0) ? 0x7fffffff : (int)(rng() % 100000) - 50; memcpy(&b[pos], &v, 4); }
        else b.resize(pos + 1);
      }
      // exact-size heap copy so ASAN sees over-reads
      char* p = (char*)malloc(b.size());
      memcpy(p, b.data(), b.size());
      tdnnf_chain_egs* e = nullptr;
      if (tdnnf_chain_egs_read_ark(p, b.size(), 0, &e) == 0) {
        ++ok;
        int cnt = 0;
        tdnnf_chain_egs_count(e, &cnt);
        for (int i = 0; i < cnt; ++i) {
          int ni = 0, no = 0; const char* key;
          tdnnf_chain_egs_example(e, i, &key, nullptr, &ni, &no);
          for (int j = 0; j < ni; ++j) {
            const char* name; int T, S, D, t0;
            tdnnf_chain_egs_input(e, i, j, &name, nullptr, nullptr, nullptr, nullptr);
            if (tdnnf_chain_egs_merge_input(e, i, 1, name, nullptr, 0, &T, &S, &D, &t0) == 0) {
              std::vector<float> out((size_t)T * S * D);
              tdnnf_chain_egs_merge_input(e, i, 1, name, out.data(), out.size(), &T, &S, &D, &t0);
            }
          }
          for (int j = 0; j < no; ++j) {
            const char* name; int ns, fps, ld;
            tdnnf_chain_egs_supervision(e, i, j, &name, nullptr, &ns, &fps, &ld, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
            std::vector<float> dw((size_t)ns * fps);
            tdnnf_host_num_graph* g = nullptr;
            int S2, T2; float w;
            if (tdnnf_chain_egs_merge_supervision(e, i, 1, name, ld, dw.data(), (int)dw.size(), &S2, &T2, &w, &g) == 0) tdnnf_host_num_graph_free(g);
          }
        }
        tdnnf_chain_egs_free(e);
      } else ++err;
      tdnnf_host_graph* hg = nullptr;
      if (tdnnf_den_graph_parse_fst_binary(p, b.size(), 29, &hg) == 0) tdnnf_host_graph_free(hg);
      free(p);
    }
  }
  printf("ok %ld err %ld\n", ok, err);
}
