// Reader for the training examples of the chain recipes, host only (no CUDA): the data format on the near side of the
// path ("egs readers (NnetChainExample)").  Everything here is upstream Kaldi / OpenFst on-disk format, none
// of it shipped with the reference: restated from the published formats, pinned only by the round trips of tests/egs_ref.py
// (an independent Python writer) -- never by a file Kaldi itself wrote.  What is read:
//   * a Kaldi table archive: "key " then "\0B" + binary object, or a text object, repeated (kaldi: util/kaldi-holder-inl.h);
//   * NnetChainExample = <Nnet3ChainEg> <NumInputs> n NnetIo* <NumOutputs> m NnetChainSupervision* </Nnet3ChainEg>
//     (kaldi: nnet3/nnet-chain-example.cc), NnetIo = <NnetIo> name <I1V> indexes GeneralMatrix </NnetIo>
//     (nnet3/nnet-example.cc), the index vector with its one-byte delta coding (nnet3/nnet-common.cc);
//   * GeneralMatrix: full ("FM"/"DM"), compressed ("CM" one byte + per-column quartile headers, "CM2" two bytes, "CM3"
//     one byte; matrix/compressed-matrix.cc) and sparse ("SM"; matrix/sparse-matrix.cc), binary and text;
//   * chain::Supervision (chain/chain-supervision.cc): <Weight> <NumSequences> <FramesPerSeq> <LabelDim> <End2End>, then ONE
//     FST (constrained egs) or <Fsts> one per sequence </Fsts> (the `--constrained false` egs of the recipes,
//     run_TDNN_DARTSV3_fbk_stride_pretrain.sh:195), <AlignmentPdfs>; FSTs as text (fstprint lines ended by an empty line)
//     or OpenFst binary "compact_acceptor" (CompactFst over AcceptorCompactor<StdArc>: uint32 state offsets, 12-byte
//     (label, weight, nextstate) elements, a leading label -1 element = the final weight); <DW> / <DW2> derivative weights;
//   * den.fst as chain-make-den-fst writes it: OpenFst binary "vector" / "standard".
// and what is made of it: the minibatch nnet3-chain-merge-egs would form from a run of examples (MergeChainExamples: rows
// t-major with the sequence index fastest, derivative weights likewise, the per-sequence FSTs in order) as the input
// matrix, the derivative weights and the numerator graph arrays the C ABI takes.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "chain_io.h"

using namespace tdnnf;

namespace {

struct ParseError {
  std::string msg;
};
[[noreturn]] void bad(const std::string& m) { throw ParseError{m}; }

// Bounds-checked cursor over the caller's buffer with Kaldi's basic-type framing (base/io-funcs-inl.h).
class Cursor {
 public:
  Cursor(const char* p, size_t n) : p_(reinterpret_cast<const unsigned char*>(p)), n_(n) {}
  size_t pos() const { return pos_; }
  bool at_end() const { return pos_ >= n_; }
  int peek() const { return pos_ < n_ ? p_[pos_] : -1; }
  int get() {
    if (pos_ >= n_) bad("unexpected end of data");
    return p_[pos_++];
  }
  void read(void* dst, size_t k) {
    if (k > n_ - pos_) bad("unexpected end of data (wanted " + std::to_string(k) + " bytes at offset " + std::to_string(pos_) + ")");
    memcpy(dst, p_ + pos_, k);
    pos_ += k;
  }
  const unsigned char* take(size_t k) {
    if (k > n_ - pos_) bad("unexpected end of data (wanted " + std::to_string(k) + " bytes at offset " + std::to_string(pos_) + ")");
    const unsigned char* r = p_ + pos_;
    pos_ += k;
    return r;
  }
  void skip_ws() {
    while (pos_ < n_ && isspace(p_[pos_])) ++pos_;
  }
  // ReadToken: leading white space skipped (istream >> string), the token, then exactly one white-space character.
  std::string token() {
    skip_ws();
    std::string s;
    while (pos_ < n_ && !isspace(p_[pos_])) s.push_back((char)p_[pos_++]);
    if (s.empty()) bad("expected a token, found end of data");
    if (pos_ >= n_ || !isspace(p_[pos_])) bad("token '" + s.substr(0, 40) + "' is not followed by white space");
    ++pos_;
    return s;
  }
  void expect(const char* tok) {
    const size_t at = pos_;
    const std::string s = token();
    if (s != tok) bad(std::string("expected token ") + tok + ", found '" + s.substr(0, 40) + "' at offset " + std::to_string(at));
  }
  // PeekToken: the first character of the next token, the one after '<' when there is one.
  // (looks ahead without consuming: the newline that opens a text-form FST stays where read_text_fst expects it)
  int peek_token(bool binary) const {
    size_t q = pos_;
    if (!binary)
      while (q < n_ && isspace(p_[q])) ++q;
    if (q < n_ && p_[q] == '<') ++q;
    return q < n_ ? p_[q] : -1;
  }
  // a white-space delimited word in text mode (numbers); stops before brackets
  std::string word() {
    skip_ws();
    std::string s;
    while (pos_ < n_ && !isspace(p_[pos_]) && p_[pos_] != ']' && p_[pos_] != ';') s.push_back((char)p_[pos_++]);
    if (s.empty()) bad("expected a number at offset " + std::to_string(pos_));
    return s;
  }
  int32_t i32(bool binary) {
    if (binary) {
      const int sz = get();
      if (sz != 4) bad("expected a 4-byte integer, size byte is " + std::to_string(sz) + " at offset " + std::to_string(pos_ - 1));
      int32_t v;
      read(&v, 4);
      return v;
    }
    const std::string w = word();
    char* end = nullptr;
    const long v = strtol(w.c_str(), &end, 10);
    if (*end != '\0' || v < INT32_MIN || v > INT32_MAX) bad("'" + w.substr(0, 40) + "' is not an int32");
    return (int32_t)v;
  }
  float f32(bool binary) {
    if (binary) {
      const int sz = get();
      if (sz == 4) {
        float v;
        read(&v, 4);
        return v;
      }
      if (sz == 8) {
        double v;
        read(&v, 8);
        return (float)v;
      }
      bad("expected a float, size byte is " + std::to_string(sz));
    }
    return to_float(word());
  }
  bool boolean(bool binary) {
    if (!binary) skip_ws();
    const int c = get();
    if (c != 'T' && c != 'F') bad("expected T or F");
    return c == 'T';
  }
  static float to_float(const std::string& w) {
    char* end = nullptr;
    const float v = strtof(w.c_str(), &end);  // accepts inf / nan / infinity as Kaldi's text reader does
    if (end == w.c_str() || *end != '\0') bad("'" + w.substr(0, 40) + "' is not a number");
    return v;
  }

 private:
  const unsigned char* p_;
  size_t n_, pos_ = 0;
};

// Every count read from the data is checked against what is left of the buffer before anything is allocated.
void check_count(const Cursor& c, size_t total, int64_t count, size_t min_bytes_each, const char* what) {
  if (count < 0 || (min_bytes_each > 0 && (uint64_t)count > (uint64_t)(total - c.pos()) / min_bytes_each))
    bad(std::string(what) + ": count " + std::to_string(count) + " does not fit in the data");
}

struct Dense {
  int rows = 0, cols = 0;
  std::vector<float> v;
};

// ---- kaldi: matrix/kaldi-vector.cc, kaldi-matrix.cc
void read_float_vector(Cursor& c, size_t total, bool binary, std::vector<float>* out) {
  out->clear();
  if (binary) {
    const std::string tok = c.token();
    if (tok != "FV" && tok != "DV") bad("expected FV or DV, found '" + tok.substr(0, 40) + "'");
    const int32_t dim = c.i32(true);
    const size_t es = tok == "FV" ? 4 : 8;
    check_count(c, total, dim, es, "vector");
    const unsigned char* p = c.take((size_t)dim * es);
    out->resize((size_t)dim);
    for (int32_t i = 0; i < dim; ++i) {
      if (es == 4) {
        memcpy(&(*out)[i], p + 4 * (size_t)i, 4);
      } else {
        double d;
        memcpy(&d, p + 8 * (size_t)i, 8);
        (*out)[i] = (float)d;
      }
    }
    return;
  }
  c.skip_ws();
  if (c.get() != '[') bad("text vector does not start with [");
  for (;;) {
    c.skip_ws();
    if (c.peek() == ']') {
      c.get();
      break;
    }
    out->push_back(Cursor::to_float(c.word()));
  }
}

void read_full_matrix(Cursor& c, size_t total, bool binary, Dense* m) {
  if (binary) {
    const std::string tok = c.token();
    if (tok != "FM" && tok != "DM") bad("expected FM or DM, found '" + tok.substr(0, 40) + "'");
    const int32_t rows = c.i32(true), cols = c.i32(true);
    if (rows < 0 || cols < 0) bad("negative matrix dimension");
    const size_t es = tok == "FM" ? 4 : 8;
    check_count(c, total, (int64_t)rows * cols, es, "matrix");
    const size_t cnt = (size_t)rows * cols;
    const unsigned char* p = c.take(cnt * es);
    m->rows = rows;
    m->cols = cols;
    m->v.resize(cnt);
    if (es == 4) {
      if (cnt) memcpy(m->v.data(), p, cnt * 4);
    } else {
      for (size_t i = 0; i < cnt; ++i) {
        double d;
        memcpy(&d, p + 8 * i, 8);
        m->v[i] = (float)d;
      }
    }
    return;
  }
  // text: " [\n  a b c\n  d e f ]"; a row ends at a newline or ';'
  c.skip_ws();
  if (c.get() != '[') bad("text matrix does not start with [");
  m->rows = m->cols = 0;
  m->v.clear();
  int in_row = 0;
  auto end_row = [&]() {
    if (in_row == 0) return;
    if (m->rows == 0) m->cols = in_row;
    else if (in_row != m->cols) bad("text matrix rows have different lengths");
    ++m->rows;
    in_row = 0;
  };
  for (;;) {
    const int ch = c.peek();
    if (ch < 0) bad("text matrix is not closed by ]");
    if (ch == ']') {
      c.get();
      end_row();
      break;
    }
    if (ch == '\n' || ch == ';') {
      c.get();
      end_row();
    } else if (isspace(ch)) {
      c.get();
    } else {
      m->v.push_back(Cursor::to_float(c.word()));
      ++in_row;
    }
  }
}

// ---- kaldi: matrix/compressed-matrix.{h,cc}
void read_compressed_matrix(Cursor& c, size_t total, Dense* m) {
  const std::string tok = c.token();
  int format = 0;
  if (tok == "CM") format = 1;
  else if (tok == "CM2") format = 2;
  else if (tok == "CM3") format = 3;
  else bad("expected CM, CM2 or CM3, found '" + tok.substr(0, 40) + "'");
  struct {
    float min_value, range;
    int32_t num_rows, num_cols;
  } h;  // GlobalHeader without its leading `format` word, which the token carries
  c.read(&h, 16);
  if (h.num_rows < 0 || h.num_cols < 0) bad("negative compressed-matrix dimension");
  const int R = h.num_rows, C = h.num_cols;
  m->rows = R;
  m->cols = C;
  const size_t cnt = (size_t)R * C;
  if (format == 1) {
    check_count(c, total, C, 8, "compressed matrix column headers");
    const unsigned char* ph = c.take((size_t)C * 8);
    check_count(c, total, (int64_t)cnt, 1, "compressed matrix");
    const unsigned char* pd = c.take(cnt);
    m->v.resize(cnt);
    for (int j = 0; j < C; ++j) {
      uint16_t q[4];
      memcpy(q, ph + 8 * (size_t)j, 8);
      float p[4];
      for (int k = 0; k < 4; ++k) p[k] = h.min_value + h.range * 1.52590218966964e-05f * q[k];  // Uint16ToFloat
      const unsigned char* col = pd + (size_t)j * R;  // bytes are column-major
      for (int i = 0; i < R; ++i) {
        const unsigned char b = col[i];
        float x;  // CharToFloat: three linear pieces between the 0 / 25 / 75 / 100 percentiles
        if (b <= 64) x = p[0] + (p[1] - p[0]) * b * (1 / 64.0f);
        else if (b <= 192) x = p[1] + (p[2] - p[1]) * (b - 64) * (1 / 128.0f);
        else x = p[2] + (p[3] - p[2]) * (b - 192) * (1 / 63.0f);
        m->v[(size_t)i * C + j] = x;
      }
    }
  } else if (format == 2) {
    check_count(c, total, (int64_t)cnt, 2, "compressed matrix");
    const unsigned char* pd = c.take(cnt * 2);
    m->v.resize(cnt);
    const float inc = h.range * (1.0f / 65535.0f);
    for (size_t i = 0; i < cnt; ++i) {
      uint16_t u;
      memcpy(&u, pd + 2 * i, 2);
      m->v[i] = h.min_value + u * inc;
    }
  } else {
    check_count(c, total, (int64_t)cnt, 1, "compressed matrix");
    const unsigned char* pd = c.take(cnt);
    m->v.resize(cnt);
    const float inc = h.range * (1.0f / 255.0f);
    for (size_t i = 0; i < cnt; ++i) m->v[i] = h.min_value + pd[i] * inc;
  }
}

// ---- kaldi: matrix/sparse-matrix.cc
void read_sparse_matrix(Cursor& c, size_t total, bool binary, Dense* m) {
  std::vector<std::vector<std::pair<int32_t, float>>> rows;
  int dim = 0;
  if (binary) {
    c.expect("SM");
    const int32_t nr = c.i32(true);
    check_count(c, total, nr, 3, "sparse matrix");
    rows.resize((size_t)nr);
    for (int32_t r = 0; r < nr; ++r) {
      c.expect("SV");
      const int32_t d = c.i32(true), ne = c.i32(true);
      if (d < 0) bad("negative sparse-vector dimension");
      if (r == 0) dim = d;
      else if (d != dim) bad("sparse matrix rows have different dimensions");
      check_count(c, total, ne, 10, "sparse vector");
      for (int32_t k = 0; k < ne; ++k) {
        const int32_t i = c.i32(true);
        const float v = c.f32(true);
        if (i < 0 || i >= d) bad("sparse-vector index out of range");
        rows[r].push_back({i, v});
      }
    }
  } else {
    // "rows=N dim=D [ i v i v ] dim=D [ ... ] ..."
    std::string w = c.token();
    if (w.compare(0, 5, "rows=") != 0) bad("expected rows=N, found '" + w.substr(0, 40) + "'");
    const long nr = strtol(w.c_str() + 5, nullptr, 10);
    check_count(c, total, nr, 8, "sparse matrix");
    rows.resize((size_t)nr);
    for (long r = 0; r < nr; ++r) {
      w = c.token();
      if (w.compare(0, 4, "dim=") != 0) bad("expected dim=D, found '" + w.substr(0, 40) + "'");
      const long d = strtol(w.c_str() + 4, nullptr, 10);
      if (d < 0 || d > INT32_MAX) bad("bad sparse-vector dimension");
      if (r == 0) dim = (int)d;
      else if (d != dim) bad("sparse matrix rows have different dimensions");
      c.skip_ws();
      if (c.get() != '[') bad("sparse vector does not start with [");
      for (;;) {
        c.skip_ws();
        if (c.peek() == ']') {
          c.get();
          break;
        }
        const int32_t i = c.i32(false);
        const float v = Cursor::to_float(c.word());
        if (i < 0 || i >= d) bad("sparse-vector index out of range");
        rows[r].push_back({i, v});
      }
    }
  }
  if ((uint64_t)rows.size() * (uint64_t)dim > (1ull << 31)) bad("sparse matrix too large to expand");
  m->rows = (int)rows.size();
  m->cols = m->rows ? dim : 0;
  m->v.assign((size_t)m->rows * m->cols, 0.f);
  for (int r = 0; r < m->rows; ++r)
    for (const auto& iv : rows[r]) m->v[(size_t)r * dim + iv.first] = iv.second;
}

// kaldi: matrix/sparse-matrix.cc GeneralMatrix::Read
void read_general_matrix(Cursor& c, size_t total, bool binary, Dense* m) {
  if (binary) {
    const int pk = c.peek();
    if (pk == 'C') read_compressed_matrix(c, total, m);
    else if (pk == 'S') read_sparse_matrix(c, total, true, m);
    else read_full_matrix(c, total, true, m);
  } else {
    c.skip_ws();
    if (c.peek() == 'r') read_sparse_matrix(c, total, false, m);
    else read_full_matrix(c, total, false, m);
  }
}

// ---- kaldi: nnet3/nnet-common.cc ReadIndexVector (+ ReadIndexVectorElementBinary)
void read_index_vector(Cursor& c, size_t total, bool binary, std::vector<int32_t>* out /* n t x per index */) {
  c.expect("<I1V>");
  const int32_t size = c.i32(binary);
  check_count(c, total, size, 1, "index vector");
  out->resize(3 * (size_t)size);
  for (int32_t i = 0; i < size; ++i) {
    int32_t* cur = out->data() + 3 * (size_t)i;
    if (!binary) {
      c.expect("<I1>");
      cur[0] = c.i32(false);
      cur[1] = c.i32(false);
      cur[2] = c.i32(false);
      continue;
    }
    const int ch = (signed char)c.get();
    if (std::abs(ch) < 125) {  // one byte: t (first element, n = x = 0) or the step in t from the previous element
      if (i == 0) {
        cur[0] = 0;
        cur[1] = ch;
        cur[2] = 0;
      } else {
        const int64_t t = (int64_t)cur[-2] + ch;
        if (t < INT32_MIN || t > INT32_MAX) bad("index vector: t leaves the int32 range");
        cur[0] = cur[-3];
        cur[1] = (int32_t)t;
        cur[2] = cur[-1];
      }
    } else {
      if (ch != 127) bad("index vector: unexpected marker byte " + std::to_string(ch));
      cur[0] = c.i32(true);
      cur[1] = c.i32(true);
      cur[2] = c.i32(true);
    }
  }
}

// base/io-funcs-inl.h ReadIntegerVector
template <typename T>
void read_integer_vector(Cursor& c, size_t total, bool binary, std::vector<T>* out) {
  out->clear();
  if (binary) {
    const int sz = c.get();
    if (sz != (int)sizeof(T)) bad("integer vector: element size " + std::to_string(sz) + ", expected " + std::to_string(sizeof(T)));
    int32_t n;
    c.read(&n, 4);
    check_count(c, total, n, sizeof(T), "integer vector");
    out->resize((size_t)n);
    if (n) c.read(out->data(), (size_t)n * sizeof(T));
    return;
  }
  c.skip_ws();
  if (c.get() != '[') bad("text integer vector does not start with [");
  for (;;) {
    c.skip_ws();
    if (c.peek() == ']') {
      c.get();
      break;
    }
    out->push_back((T)c.i32(false));
  }
}

// ---- OpenFst binary files (fst/fst.h FstHeader::Read, fst/vector-fst.h, fst/compact-fst.h)
struct FstHeader {
  std::string fsttype, arctype;
  int32_t version = 0, flags = 0;
  uint64_t properties = 0;
  int64_t start = -1, numstates = 0, numarcs = 0;
};
std::string read_fst_string(Cursor& c, size_t total) {
  int32_t n;
  c.read(&n, 4);
  check_count(c, total, n, 1, "FST header string");
  if (n > 256) bad("FST header string is implausibly long");
  const unsigned char* p = c.take((size_t)n);
  return std::string(reinterpret_cast<const char*>(p), (size_t)n);
}
void read_fst_header(Cursor& c, size_t total, FstHeader* h) {
  int32_t magic;
  c.read(&magic, 4);
  if (magic != 2125659606) bad("not an OpenFst binary FST (magic number " + std::to_string(magic) + ")");
  h->fsttype = read_fst_string(c, total);
  h->arctype = read_fst_string(c, total);
  c.read(&h->version, 4);
  c.read(&h->flags, 4);
  c.read(&h->properties, 8);
  c.read(&h->start, 8);
  c.read(&h->numstates, 8);
  c.read(&h->numarcs, 8);
  if (h->arctype != "standard") bad("FST arc type '" + h->arctype + "' (only 'standard' = tropical float weights is read)");
  if (h->flags & 3) bad("FST carries symbol tables: not read");
  if (h->flags & 4) bad("FST written with --fst_align: not read");
  if (h->numstates < 0 || h->numstates > INT32_MAX) bad("bad FST state count");
}

void read_binary_fst(Cursor& c, size_t total, Fsm* f) {
  FstHeader h;
  read_fst_header(c, total, &h);
  const int N = (int)h.numstates;
  f->arcs.clear();
  f->finals.clear();
  f->num_states = N;
  f->start = N > 0 ? (int)h.start : -1;
  if (N > 0 && (h.start < 0 || h.start >= N)) bad("FST start state out of range");
  auto add_arc = [&](int s, int32_t il, float w, int32_t next) {
    if (next < 0 || next >= N) bad("FST arc to a state that does not exist");
    if (il < 0) bad("FST arc with a negative label");
    f->arcs.push_back(FsmArc{s, next, il, w});
  };
  if (h.fsttype == "compact_acceptor") {
    if (h.version < 2) bad("aligned CompactFst file version: not read");
    check_count(c, total, (int64_t)N + 1, 4, "CompactFst state table");
    std::vector<uint32_t> off((size_t)N + 1);
    c.read(off.data(), off.size() * 4);
    const uint64_t ncompacts = off[N];
    check_count(c, total, (int64_t)ncompacts, 12, "CompactFst elements");
    const unsigned char* p = c.take((size_t)ncompacts * 12);
    f->arcs.reserve((size_t)ncompacts);
    for (int s = 0; s < N; ++s) {
      if (off[s] > off[s + 1] || off[s + 1] > ncompacts) bad("CompactFst state table is not monotone");
      for (uint32_t k = off[s]; k < off[s + 1]; ++k) {
        int32_t label, next;
        float w;
        memcpy(&label, p + 12 * (size_t)k, 4);
        memcpy(&w, p + 12 * (size_t)k + 4, 4);
        memcpy(&next, p + 12 * (size_t)k + 8, 4);
        if (label == -1 && k == off[s]) {  // the final weight of s
          if (!std::isinf(w)) f->finals[s] = w;
        } else {
          add_arc(s, label, w, next);
        }
      }
    }
  } else if (h.fsttype == "vector") {
    for (int s = 0; s < N; ++s) {
      float fw;
      int64_t na;
      c.read(&fw, 4);
      c.read(&na, 8);
      if (!std::isinf(fw) && !std::isnan(fw)) f->finals[s] = fw;
      check_count(c, total, na, 16, "VectorFst arcs");
      for (int64_t k = 0; k < na; ++k) {
        struct {
          int32_t il, ol;
          float w;
          int32_t next;
        } a;
        c.read(&a, 16);
        add_arc(s, a.il, a.w, a.next);
      }
    }
  } else {
    bad("FST type '" + h.fsttype + "' (read: compact_acceptor, vector)");
  }
}

// fstext/kaldi-fst-io.cc ReadFstKaldi, text mode: a newline, fstprint lines, an empty line.
void read_text_fst(Cursor& c, Fsm* f) {
  while (c.peek() == ' ' || c.peek() == '\t' || c.peek() == '\r') c.get();
  if (c.get() != '\n') bad("text FST: expected a newline before the first line");
  std::string text;
  for (;;) {
    std::string line;
    bool eof = false;
    for (;;) {
      const int ch = c.peek();
      if (ch < 0) {
        eof = true;
        break;
      }
      c.get();
      if (ch == '\n') break;
      line.push_back((char)ch);
    }
    bool blank = true;
    for (char ch : line) blank = blank && isspace((unsigned char)ch);
    if (blank) break;  // the terminating empty line (or the end of the data)
    text += line;
    text.push_back('\n');
    if (eof) break;
  }
  std::string err;
  if (parse_fsm(text.data(), text.size(), f, &err) != TDNNF_OK) bad("text FST: " + err);
}

struct EgFst {
  Fsm f;
  std::vector<int32_t> arcs3, final_states;  // flattened for the accessor
  std::vector<float> arc_w, final_w;
  void flatten() {
    arcs3.clear();
    arc_w.clear();
    final_states.clear();
    final_w.clear();
    for (const FsmArc& a : f.arcs) {
      arcs3.push_back(a.src);
      arcs3.push_back(a.dst);
      arcs3.push_back(a.ilabel);
      arc_w.push_back(a.weight);
    }
    for (const auto& kv : f.finals) {
      final_states.push_back(kv.first);
      final_w.push_back(kv.second);
    }
  }
};

struct EgIo {
  std::string name;
  std::vector<int32_t> indexes;
  Dense m;
};

struct EgSup {
  std::string name;
  std::vector<int32_t> indexes;
  float weight = 1.f;
  int num_sequences = 1, frames_per_seq = 0, label_dim = 0;
  bool e2e = false;
  std::vector<EgFst> fsts;  // one (constrained) or num_sequences (unconstrained / e2e)
  std::vector<int32_t> alignment_pdfs;
  std::vector<float> deriv_weights;
};

struct Example {
  std::string key;
  bool binary = false;
  std::vector<EgIo> inputs;
  std::vector<EgSup> outputs;
};

void read_one_fst(Cursor& c, size_t total, bool binary, EgFst* out) {
  if (binary) read_binary_fst(c, total, &out->f);
  else read_text_fst(c, &out->f);
  out->flatten();
}

// kaldi: chain/chain-supervision.cc Supervision::Read
void read_supervision(Cursor& c, size_t total, bool binary, EgSup* s) {
  c.expect("<Supervision>");
  c.expect("<Weight>");
  s->weight = c.f32(binary);
  c.expect("<NumSequences>");
  s->num_sequences = c.i32(binary);
  c.expect("<FramesPerSeq>");
  s->frames_per_seq = c.i32(binary);
  c.expect("<LabelDim>");
  s->label_dim = c.i32(binary);
  if (s->num_sequences <= 0 || s->frames_per_seq <= 0 || s->label_dim <= 0) bad("Supervision with a non-positive dimension");
  s->e2e = false;
  if (c.peek_token(binary) == 'E') {  // files older than the unconstrained egs have no <End2End>
    c.expect("<End2End>");
    s->e2e = c.boolean(binary);
  }
  if (!s->e2e) {
    s->fsts.resize(1);
    read_one_fst(c, total, binary, &s->fsts[0]);
  } else {
    c.expect("<Fsts>");
    check_count(c, total, s->num_sequences, 2, "Supervision FSTs");
    s->fsts.resize((size_t)s->num_sequences);
    for (auto& f : s->fsts) read_one_fst(c, total, binary, &f);
    c.expect("</Fsts>");
  }
  if (c.peek_token(binary) == 'A') {
    c.expect("<AlignmentPdfs>");
    read_integer_vector<int32_t>(c, total, binary, &s->alignment_pdfs);
  }
  c.expect("</Supervision>");
}

// kaldi: nnet3/nnet-chain-example.cc NnetChainSupervision::Read, NnetChainExample::Read; nnet3/nnet-example.cc NnetIo::Read
void read_example(Cursor& c, size_t total, bool binary, Example* ex) {
  c.expect("<Nnet3ChainEg>");
  c.expect("<NumInputs>");
  const int32_t ni = c.i32(binary);
  check_count(c, total, ni, 16, "example inputs");
  if (ni < 1) bad("example without inputs");
  ex->inputs.resize((size_t)ni);
  for (EgIo& io : ex->inputs) {
    c.expect("<NnetIo>");
    io.name = c.token();
    read_index_vector(c, total, binary, &io.indexes);
    read_general_matrix(c, total, binary, &io.m);
    c.expect("</NnetIo>");
    if ((size_t)io.m.rows * 3 != io.indexes.size())
      bad("input '" + io.name + "': " + std::to_string(io.indexes.size() / 3) + " indexes for " + std::to_string(io.m.rows) + " rows");
  }
  c.expect("<NumOutputs>");
  const int32_t no = c.i32(binary);
  check_count(c, total, no, 16, "example outputs");
  ex->outputs.resize((size_t)no);
  for (EgSup& s : ex->outputs) {
    c.expect("<NnetChainSup>");
    s.name = c.token();
    read_index_vector(c, total, binary, &s.indexes);
    read_supervision(c, total, binary, &s);
    const std::string tok = c.token();
    if (tok == "<DW>") {  // WriteVectorAsChar: bytes of 255 * weight in binary, a plain vector in text
      if (binary) {
        std::vector<unsigned char> b;
        read_integer_vector<unsigned char>(c, total, true, &b);
        s.deriv_weights.resize(b.size());
        for (size_t i = 0; i < b.size(); ++i) s.deriv_weights[i] = b[i] * (1.0f / 255.0f);
      } else {
        read_float_vector(c, total, false, &s.deriv_weights);
      }
    } else if (tok == "<DW2>") {
      read_float_vector(c, total, binary, &s.deriv_weights);
    } else {
      bad("expected <DW> or <DW2>, found '" + tok.substr(0, 40) + "'");
    }
    c.expect("</NnetChainSup>");
    const size_t frames = (size_t)s.num_sequences * s.frames_per_seq;
    if (s.indexes.size() != 3 * frames)
      bad("supervision '" + s.name + "': " + std::to_string(s.indexes.size() / 3) + " indexes for " + std::to_string(frames) + " frames");
    if (!s.deriv_weights.empty() && s.deriv_weights.size() != frames) bad("supervision '" + s.name + "': derivative weights do not match the frames");
  }
  c.expect("</Nnet3ChainEg>");
}

}  // namespace

struct tdnnf_chain_egs {
  std::vector<Example> ex;
};

extern "C" int tdnnf_chain_egs_read_ark(const char* buf, uint64_t len, int max_examples, tdnnf_chain_egs** out) {
  TDNNF_REQUIRE(buf && out, "null argument");
  tdnnf_chain_egs* e = new tdnnf_chain_egs();
  Cursor c(buf, (size_t)len);
  try {
    for (;;) {
      c.skip_ws();
      if (c.at_end() || (max_examples > 0 && (int)e->ex.size() >= max_examples)) break;
      e->ex.emplace_back();
      Example& ex = e->ex.back();
      ex.key = c.token();
      ex.binary = false;
      if (c.peek() == 0) {  // "\0B" announces a binary object
        c.get();
        if (c.get() != 'B') bad("key '" + ex.key + "': \\0 is not followed by B");
        ex.binary = true;
      }
      try {
        read_example(c, (size_t)len, ex.binary, &ex);
      } catch (ParseError& pe) {
        pe.msg = "example '" + ex.key.substr(0, 80) + "' (" + std::to_string(e->ex.size() - 1) + "): " + pe.msg;
        throw;
      }
    }
  } catch (const ParseError& pe) {
    delete e;
    std::string msg = pe.msg;  // it may quote the data: keep the message printable
    for (char& ch : msg)
      if ((unsigned char)ch < 0x20 || (unsigned char)ch > 0x7e) ch = '?';
    return fail(TDNNF_ERR_INVALID, "egs archive: " + msg);
  } catch (const std::bad_alloc&) {
    delete e;
    return fail(TDNNF_ERR_INVALID, "egs archive: out of memory");
  }
  *out = e;
  return TDNNF_OK;
}

extern "C" int tdnnf_chain_egs_free(tdnnf_chain_egs* e) {
  delete e;
  return TDNNF_OK;
}

extern "C" int tdnnf_chain_egs_count(const tdnnf_chain_egs* e, int* num_examples) {
  TDNNF_REQUIRE(e && num_examples, "null argument");
  *num_examples = (int)e->ex.size();
  return TDNNF_OK;
}

#define EGS_EXAMPLE(e, i)                                                                 \
  TDNNF_REQUIRE((e) != nullptr, "null argument");                                         \
  TDNNF_REQUIRE((i) >= 0 && (size_t)(i) < (e)->ex.size(), "example index out of range");  \
  const Example& ex = (e)->ex[(size_t)(i)]

extern "C" int tdnnf_chain_egs_example(const tdnnf_chain_egs* e, int i, const char** key, int* binary, int* num_inputs, int* num_outputs) {
  EGS_EXAMPLE(e, i);
  if (key) *key = ex.key.c_str();
  if (binary) *binary = ex.binary ? 1 : 0;
  if (num_inputs) *num_inputs = (int)ex.inputs.size();
  if (num_outputs) *num_outputs = (int)ex.outputs.size();
  return TDNNF_OK;
}

extern "C" int tdnnf_chain_egs_input(const tdnnf_chain_egs* e, int i, int j, const char** name, int* rows, int* cols,
                                     const int32_t** indexes, const float** data) {
  EGS_EXAMPLE(e, i);
  TDNNF_REQUIRE(j >= 0 && (size_t)j < ex.inputs.size(), "input index out of range");
  const EgIo& io = ex.inputs[(size_t)j];
  if (name) *name = io.name.c_str();
  if (rows) *rows = io.m.rows;
  if (cols) *cols = io.m.cols;
  if (indexes) *indexes = io.indexes.data();
  if (data) *data = io.m.v.data();
  return TDNNF_OK;
}

extern "C" int tdnnf_chain_egs_supervision(const tdnnf_chain_egs* e, int i, int j, const char** name, float* weight, int* num_sequences,
                                           int* frames_per_seq, int* label_dim, int* e2e, int* num_fsts, const int32_t** indexes,
                                           int* num_deriv_weights, const float** deriv_weights, int* num_alignment_pdfs,
                                           const int32_t** alignment_pdfs) {
  EGS_EXAMPLE(e, i);
  TDNNF_REQUIRE(j >= 0 && (size_t)j < ex.outputs.size(), "supervision index out of range");
  const EgSup& s = ex.outputs[(size_t)j];
  if (name) *name = s.name.c_str();
  if (weight) *weight = s.weight;
  if (num_sequences) *num_sequences = s.num_sequences;
  if (frames_per_seq) *frames_per_seq = s.frames_per_seq;
  if (label_dim) *label_dim = s.label_dim;
  if (e2e) *e2e = s.e2e ? 1 : 0;
  if (num_fsts) *num_fsts = (int)s.fsts.size();
  if (indexes) *indexes = s.indexes.data();
  if (num_deriv_weights) *num_deriv_weights = (int)s.deriv_weights.size();
  if (deriv_weights) *deriv_weights = s.deriv_weights.data();
  if (num_alignment_pdfs) *num_alignment_pdfs = (int)s.alignment_pdfs.size();
  if (alignment_pdfs) *alignment_pdfs = s.alignment_pdfs.data();
  return TDNNF_OK;
}

extern "C" int tdnnf_chain_egs_fst(const tdnnf_chain_egs* e, int i, int j, int k, int* start, int* num_states, int* num_arcs,
                                   const int32_t** arcs, const float** arc_weights, int* num_finals, const int32_t** final_states,
                                   const float** final_weights) {
  EGS_EXAMPLE(e, i);
  TDNNF_REQUIRE(j >= 0 && (size_t)j < ex.outputs.size(), "supervision index out of range");
  const EgSup& s = ex.outputs[(size_t)j];
  TDNNF_REQUIRE(k >= 0 && (size_t)k < s.fsts.size(), "FST index out of range");
  const EgFst& f = s.fsts[(size_t)k];
  if (start) *start = f.f.start;
  if (num_states) *num_states = f.f.num_states;
  if (num_arcs) *num_arcs = (int)f.f.arcs.size();
  if (arcs) *arcs = f.arcs3.data();
  if (arc_weights) *arc_weights = f.arc_w.data();
  if (num_finals) *num_finals = (int)f.final_states.size();
  if (final_states) *final_states = f.final_states.data();
  if (final_weights) *final_weights = f.final_w.data();
  return TDNNF_OK;
}

// ---- the minibatch of examples [first, first + count): kaldi nnet3/nnet-chain-example.cc MergeChainExamples
namespace {

// sequences an example contributes: its supervision says; an example without outputs: the largest n of its first input + 1
int example_num_sequences(const Example& ex) {
  if (!ex.outputs.empty()) return ex.outputs[0].num_sequences;
  int mx = 0;
  const auto& ix = ex.inputs[0].indexes;
  for (size_t r = 0; r + 2 < ix.size(); r += 3) mx = std::max(mx, ix[r]);
  return mx + 1;
}

int find_named(const tdnnf_chain_egs* e, int first, int count, const char* name, bool output, std::vector<int>* which, std::string* err) {
  which->clear();
  for (int i = first; i < first + count; ++i) {
    const Example& ex = e->ex[(size_t)i];
    int found = -1;
    const size_t n = output ? ex.outputs.size() : ex.inputs.size();
    for (size_t j = 0; j < n; ++j)
      if ((output ? ex.outputs[j].name : ex.inputs[j].name) == name) found = (int)j;
    if (found < 0) {
      *err = "example '" + ex.key + "' has no " + (output ? "supervision" : "input") + " named '" + name + "'";
      return TDNNF_ERR_INVALID;
    }
    which->push_back(found);
  }
  return TDNNF_OK;
}

}  // namespace

#define EGS_RANGE(e, first, count)                                                                         \
  TDNNF_REQUIRE((e) != nullptr, "null argument");                                                          \
  TDNNF_REQUIRE((first) >= 0 && (count) > 0 && (size_t)(first) + (size_t)(count) <= (e)->ex.size(), "example range out of bounds")

extern "C" int tdnnf_chain_egs_merge_input(const tdnnf_chain_egs* e, int first, int count, const char* name, float* out,
                                           int64_t out_floats, int* num_t, int* num_seqs, int* cols, int* first_t) {
  EGS_RANGE(e, first, count);
  TDNNF_REQUIRE(name, "null argument");
  std::vector<int> which;
  std::string err;
  if (find_named(e, first, count, name, false, &which, &err)) return fail(TDNNF_ERR_INVALID, err);
  int S = 0, T = -1, D = -1, t0 = 0;
  std::vector<int> base((size_t)count);
  for (int i = 0; i < count; ++i) {
    base[(size_t)i] = S;
    S += example_num_sequences(e->ex[(size_t)(first + i)]);
  }
  // every sequence must bring the same number of frames; rows go to (rank of t within the sequence) * S + sequence
  struct Row { int seq, t; const float* src; };
  std::vector<Row> rows;
  for (int i = 0; i < count; ++i) {
    const Example& ex = e->ex[(size_t)(first + i)];
    const EgIo& io = ex.inputs[(size_t)which[(size_t)i]];
    if (D < 0) D = io.m.cols;
    TDNNF_REQUIRE(io.m.cols == D, "input '" + std::string(name) + "': examples differ in dimension");
    const int ns = example_num_sequences(ex);
    for (int r = 0; r < io.m.rows; ++r) {
      const int32_t* ix = io.indexes.data() + 3 * (size_t)r;
      TDNNF_REQUIRE(ix[0] >= 0 && ix[0] < ns, "input '" + std::string(name) + "': n index outside the example's sequences");
      TDNNF_REQUIRE(ix[2] == 0, "input '" + std::string(name) + "': x index is not 0");
      rows.push_back(Row{base[(size_t)i] + ix[0], ix[1], io.m.v.data() + (size_t)r * D});
    }
  }
  std::vector<std::vector<std::pair<int, const float*>>> per_seq((size_t)S);
  for (const Row& r : rows) per_seq[(size_t)r.seq].push_back({r.t, r.src});
  for (int s = 0; s < S; ++s) {
    auto& v = per_seq[(size_t)s];
    std::sort(v.begin(), v.end(), [](const std::pair<int, const float*>& a, const std::pair<int, const float*>& b) { return a.first < b.first; });
    for (size_t k = 1; k < v.size(); ++k) TDNNF_REQUIRE(v[k].first != v[k - 1].first, "input '" + std::string(name) + "': a frame appears twice");
    if (T < 0) {
      T = (int)v.size();
      t0 = v.empty() ? 0 : v[0].first;
    }
    TDNNF_REQUIRE((int)v.size() == T, "input '" + std::string(name) + "': sequences differ in their number of frames");
  }
  if (num_t) *num_t = T;
  if (num_seqs) *num_seqs = S;
  if (cols) *cols = D;
  if (first_t) *first_t = t0;
  if (!out) return TDNNF_OK;  // dimensions only
  TDNNF_REQUIRE(out_floats >= (int64_t)T * S * D, "output buffer too small");
  for (int s = 0; s < S; ++s)
    for (int k = 0; k < T; ++k) memcpy(out + ((size_t)k * S + s) * D, per_seq[(size_t)s][(size_t)k].second, sizeof(float) * (size_t)D);
  return TDNNF_OK;
}

extern "C" int tdnnf_chain_egs_merge_supervision(const tdnnf_chain_egs* e, int first, int count, const char* name, int num_pdfs,
                                                 float* deriv_weights, int deriv_weights_floats, int* num_seqs, int* frames_per_seq,
                                                 float* weight, tdnnf_host_num_graph** num_graph) {
  EGS_RANGE(e, first, count);
  TDNNF_REQUIRE(name, "null argument");
  std::vector<int> which;
  std::string err;
  if (find_named(e, first, count, name, true, &which, &err)) return fail(TDNNF_ERR_INVALID, err);
  int S = 0, T = -1, label_dim = -1;
  double wsum = 0.0;
  for (int i = 0; i < count; ++i) {
    const EgSup& s = e->ex[(size_t)(first + i)].outputs[(size_t)which[(size_t)i]];
    if (T < 0) {
      T = s.frames_per_seq;
      label_dim = s.label_dim;
    }
    TDNNF_REQUIRE(s.frames_per_seq == T, "supervision '" + std::string(name) + "': examples differ in frames per sequence");
    TDNNF_REQUIRE(s.label_dim == label_dim, "supervision '" + std::string(name) + "': examples differ in label dimension");
    S += s.num_sequences;
    wsum += (double)s.weight * s.num_sequences;
  }
  if (num_seqs) *num_seqs = S;
  if (frames_per_seq) *frames_per_seq = T;
  if (weight) *weight = (float)(wsum / S);
  if (deriv_weights) {  // t-major, sequence fastest (the order of the merged supervision's indexes: sorted by t, then n)
    TDNNF_REQUIRE(deriv_weights_floats >= T * S, "derivative-weight buffer too small");
    int base = 0;
    for (int i = 0; i < count; ++i) {
      const EgSup& s = e->ex[(size_t)(first + i)].outputs[(size_t)which[(size_t)i]];
      for (int t = 0; t < T; ++t)
        for (int n = 0; n < s.num_sequences; ++n)
          deriv_weights[(size_t)t * S + base + n] = s.deriv_weights.empty() ? 1.0f : s.deriv_weights[(size_t)t * s.num_sequences + n];
      base += s.num_sequences;
    }
  }
  if (num_graph) {
    TDNNF_REQUIRE(num_pdfs > 0, "num_pdfs must be positive");
    TDNNF_REQUIRE(label_dim == num_pdfs, "supervision '" + std::string(name) + "': label dimension " + std::to_string(label_dim) +
                                             " is not num_pdfs " + std::to_string(num_pdfs));
    std::vector<Fsm> fsms;
    for (int i = 0; i < count; ++i) {
      const Example& ex = e->ex[(size_t)(first + i)];
      const EgSup& s = ex.outputs[(size_t)which[(size_t)i]];
      // a constrained supervision holds ONE FST over all its sequences (frame-indexed, sequence after sequence): only
      // the single-sequence case is a per-sequence acceptor the generic numerator can take as it is
      TDNNF_REQUIRE(s.e2e || s.num_sequences == 1, "example '" + ex.key + "': a merged constrained supervision (one FST over " +
                                                        std::to_string(s.num_sequences) + " sequences) is not read; pass the unmerged examples");
      for (const EgFst& f : s.fsts) fsms.push_back(f.f);
    }
    return build_host_num_graph(fsms, num_pdfs, num_graph);
  }
  return TDNNF_OK;
}

// den.fst as chain-make-den-fst writes it (OpenFst binary VectorFst<StdArc>; a compact acceptor is read too)
extern "C" int tdnnf_den_graph_parse_fst_binary(const char* buf, uint64_t len, int num_pdfs, tdnnf_host_graph** out) {
  TDNNF_REQUIRE(buf && out && num_pdfs > 0, "bad argument");
  Fsm f;
  try {
    Cursor c(buf, (size_t)len);
    read_binary_fst(c, (size_t)len, &f);
  } catch (const ParseError& pe) {
    return fail(TDNNF_ERR_INVALID, "den.fst: " + pe.msg);
  } catch (const std::bad_alloc&) {
    return fail(TDNNF_ERR_INVALID, "den.fst: out of memory");
  }
  TDNNF_REQUIRE(f.start >= 0, "den.fst: empty FST");
  return build_host_den_graph(f, num_pdfs, out);
}
