// elem_host.hpp -- C++14 host side of `RNAelem train / scan` above the librelem C ABI (include/relem.h).
//
// everything the reference does on the host around the per-sequence DP, restated so
// that the same command line produces the same minibatches, the same shuffled negatives, the same Adam trajectory and
// byte-compatible train.model / train.interim / scan.raw files, while the DP itself runs in the CUDA kernels.
// Nothing in this file computes a DP cell; there is no CPU path for that.
//
// Reference behaviour followed (file:line in /root/reference/RNAelem):
//   messages / streams            util.hpp:95-180
//   FASTQ + minibatch reader      fastq_io.hpp:23-167
//   k-let preserving shuffle      ushuffle/ushuffle.c (uShuffle, Jiang et al. 2008: Euler tour over a random
//                                 arborescence drawn with Wilson's algorithm) driven by glibc rand(), util.hpp:411
//   position weights              motif_model.hpp:62-70
//   parameter vector              motif_model.hpp:136-168, profile_hmm.hpp:103-111,286-313
//   Adam                          optimizer.hpp:72-173
//   model files                   motif_io.hpp:29-262
#ifndef RELEM_ELEM_HOST_HPP
#define RELEM_ELEM_HOST_HPP
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace relem {

using V = std::vector<double>;
using VV = std::vector<V>;
using VI = std::vector<int>;
constexpr double kInf = std::numeric_limits<double>::infinity();

// ------------------------------------------------------------------------------------------ text formatting
// vectors print as [a,b,c] (nested: [[..],[..]]) with the stream's precision -- default 6 significant digits,
// which is what makes train.model / scan.raw byte-compatible (util.hpp:95-102).
template <class T>
std::ostream& operator<<(std::ostream& os, const std::vector<T>& v) {
  os << '[';
  for (size_t k = 0; k < v.size(); ++k) {
    if (k) os << ',';
    os << v[k];
  }
  return os << ']';
}

inline void put_line(std::ostream& os) { os << std::endl; }
template <class T, class... R>
void put_line(std::ostream& os, const T& head, const R&... rest) {
  os << head;
  if (sizeof...(rest)) os << ' ';
  put_line(os, rest...);
}
template <class... A>
void cry(const A&... a) { put_line(std::cerr, a...); }
template <class... A>
[[noreturn]] void die(const A&... a) {
  cry(a...);
  throw std::runtime_error("die");   // util.hpp:121-124; main() turns it into exit status 1
}
template <class... A>
void check(bool ok, const A&... a) { if (!ok) die(a...); }

template <class T>
std::string to_text(const T& t) { std::ostringstream o; o << t; return o.str(); }
template <class T>
std::string glue(const std::string& key, const T& t) { return key + to_text(t); }

// numbered output channels: 0 is always the null sink; 1..3 are --out1..3 ("~COUT~", "~CERR~", "~NULL~" or a path;
// the same path given twice shares one file), util.hpp:128-166
class OutputSet {
  struct NullBuf : std::streambuf { int overflow(int c) override { return c; } };
  NullBuf nullbuf_;
  std::ostream null_{&nullbuf_};
  std::map<std::string, std::unique_ptr<std::ofstream>> files_;
  std::vector<std::string> name_;
 public:
  explicit OutputSet(int n = 4) : name_(n + 1, "~NULL~") {}
  void bind(int id, const std::string& target) {
    if (target != "~NULL~" && target != "~COUT~" && target != "~CERR~" && !files_.count(target)) {
      std::unique_ptr<std::ofstream> f(new std::ofstream(target));
      check(!!*f, "cannot open:", target);
      files_[target] = std::move(f);
    }
    name_.at(id) = target;
  }
  std::ostream& at(int id) {
    const std::string& n = name_.at(id);
    if (n == "~NULL~") return null_;
    if (n == "~COUT~") return std::cout;
    if (n == "~CERR~") return std::cerr;
    return *files_[n];
  }
  template <class... A>
  void dat(int id, const A&... a) { if (id >= 0) put_line(at(id), a...); }
};

// ------------------------------------------------------------------------------------------------ sequences
inline int base_code(char c) {   // bio_sequence.hpp:30-41
  switch (c) {
    case 'A': case 'a': return 1;
    case 'C': case 'c': return 2;
    case 'G': case 'g': return 3;
    case 'T': case 't': case 'U': case 'u': return 4;
    default: return 0;
  }
}
template <class Codes>
std::string codes_to_text(const Codes& seq) {
  std::string s(seq.size(), 'N');
  for (size_t k = 0; k < seq.size(); ++k) s[k] = "NACGU"[seq[k]];
  return s;
}

struct Read {                  // ~3 bytes per base: a scan input of 10^6 x 200-nt reads stays well under 1 GB
  std::string id;              // whole header line including '@'
  std::vector<uint8_t> seq;    // base codes 0..4
  std::vector<int16_t> qual;   // quality values (char - 33), L+1 of them in RNAelem's FASTQ dialect
};

// Strict 4-lines-per-record FASTQ (fastq_io.hpp:64-114).  A record counts only if the stream is still good after
// its fourth line, so a final record without a trailing newline is dropped exactly as the reference drops it.
class FastqReader {
  std::vector<Read> rec_;
  VI order_;          // permutation applied by the epoch shuffles
  int next_ = 0;
  int n_shuffles_ = 0;
 public:
  void open(const std::string& path, const std::string& encoding = "sanger") {
    int base = (encoding == "sanger" || encoding == "illumina1.8") ? 33
             : (encoding == "solexa" || encoding == "illumina1.3" || encoding == "illumina1.5") ? 64 : -1;
    check(base != -1, "wrong encoding:", encoding);
    std::ifstream in(path);
    check(!!in, "could not open:", path);
    rec_.clear();
    std::string id, seq, plus, qual;
    while (!in.eof()) {
      std::getline(in, id);
      std::getline(in, seq);
      std::getline(in, plus);
      std::getline(in, qual);
      if (in.eof()) break;
      Read r;
      r.id = id;
      r.seq.resize(seq.size());
      for (size_t k = 0; k < seq.size(); ++k) r.seq[k] = uint8_t(base_code(seq[k]));
      r.qual.resize(qual.size());
      for (size_t k = 0; k < qual.size(); ++k) r.qual[k] = int16_t(int(qual[k]) - base);
      rec_.push_back(std::move(r));
    }
    order_.resize(rec_.size());
    for (size_t k = 0; k < rec_.size(); ++k) order_[k] = int(k);
    next_ = 0;
    n_shuffles_ = 0;
  }
  const Read& get() { return rec_[order_[next_++]]; }
  // the reference shuffles six parallel index arrays with identically seeded engines (fastq_io.hpp:115-124);
  // permuting one record order with that engine is the same permutation
  void shuffle() {
    std::mt19937 eng;
    eng.seed(n_shuffles_++);
    std::shuffle(order_.begin(), order_.end(), eng);
  }
  bool at_end() const { return next_ == int(rec_.size()); }
  void skip(int n = 1) { next_ += n; }
  void rewind() { next_ = 0; }
  int consumed() const { return next_; }
  int size() const { return int(rec_.size()); }
};

// minibatch view (fastq_io.hpp:132-167): a batch ends after batch_size reads or at the end of the epoch; rewinding
// at the end of an epoch reshuffles
class FastqBatchReader {
  FastqReader rd_;
  int batch_ = 0, in_batch_ = 0, epochs_ = 0;
 public:
  void open(const std::string& path) { rd_.open(path); in_batch_ = 0; epochs_ = 0; }
  void set_batch_size(int n) { batch_ = n < 0 ? rd_.size() : n; }
  const Read& get() { ++in_batch_; return rd_.get(); }
  bool batch_done() const { return batch_ <= in_batch_ || rd_.at_end(); }
  bool epoch_done() const { return rd_.at_end(); }
  void next_batch() {
    if (epoch_done()) { rd_.shuffle(); rd_.rewind(); ++epochs_; }
    in_batch_ = 0;
  }
  void skip(int n) { in_batch_ += n; rd_.skip(n); }
  int in_batch() const { return in_batch_; }
  int epochs() const { return epochs_; }
  int size() const { return rd_.size(); }
  int batch_size() const { return batch_; }
  int consumed_in_epoch() const { return rd_.consumed(); }
};

// RNAelem::set_ws (motif_model.hpp:62-70): L+1 quality values -> L log position weights relative to the modal
// quality (last maximum of the histogram, util.hpp:231-241) and the trailing "contains the motif" flag
// (returned: true when the flag value is 0, i.e. the quality string ends in '!').
template <class Quals>
bool quality_to_weights(const Quals& q, V& ws) {
  VI hist(127 - 33, 0);
  for (int v : q) hist.at(v) += 1;
  int mode = 0, best = std::numeric_limits<int>::lowest();
  for (int k = 0; k < int(hist.size()); ++k)
    if (best <= hist[k]) { mode = k; best = hist[k]; }
  ws.clear();
  for (size_t k = 0; k + 1 < q.size(); ++k) ws.push_back(std::log((0.01 + double(q[k])) / (0.01 + mode)));
  return q.back() == 0;
}

// ------------------------------------------------------------------------------------- k-let preserving shuffle
// uShuffle (Jiang, Anderson, Gillespie, Mayne; BMC Bioinformatics 2008), as vendored by the reference in
// ushuffle/ushuffle.c:146-273.  The (k-1)-mers are the vertices of a multigraph (numbered by first occurrence), every
// position contributes one edge to its successor.  A uniformly random arborescence rooted at the last (k-1)-mer
// (Wilson's loop-erased walks, vertices visited in id order) fixes the LAST edge each vertex leaves by; the other edges
// are permuted (Fisher-Yates from the top); the Euler walk from the first (k-1)-mer spells the shuffled sequence.
// Random numbers are consumed in exactly the reference's order so that, with the same generator, the negatives are
// the same sequences.
template <class Rand>
std::string klet_shuffle(const std::string& s, int k, Rand&& rnd) {
  const int n = int(s.size());
  std::string t(s);
  if (k >= n) return t;
  auto permute = [&](auto first, int len) {
    for (int i = len - 1; i > 0; --i) {
      int j = int(rnd() % (i + 1));
      std::swap(first[i], first[j]);
    }
  };
  if (k <= 1) { permute(t.begin(), n); return t; }
  const int nlets = n - k + 2;
  std::map<std::string, int> ids;
  VI let_vertex(nlets), first_pos;
  for (int p = 0; p < nlets; ++p) {
    auto ins = ids.emplace(s.substr(p, k - 1), int(first_pos.size()));
    if (ins.second) first_pos.push_back(p);
    let_vertex[p] = ins.first->second;
  }
  const int nv = int(first_pos.size());
  const int root = let_vertex[nlets - 1];
  std::vector<VI> succ(nv);
  for (int p = 0; p + 1 < nlets; ++p) succ[let_vertex[p]].push_back(let_vertex[p + 1]);
  std::vector<char> in_tree(nv, 0);
  VI exit_edge(nv, 0);
  in_tree[root] = 1;
  for (int v = 0; v < nv; ++v) {
    int u = v;
    while (!in_tree[u]) {
      exit_edge[u] = int(rnd() % long(succ[u].size()));
      u = succ[u][exit_edge[u]];
    }
    u = v;
    while (!in_tree[u]) { in_tree[u] = 1; u = succ[u][exit_edge[u]]; }
  }
  for (int v = 0; v < nv; ++v) {
    VI& e = succ[v];
    if (v != root) {
      std::swap(e.back(), e[exit_edge[v]]);
      permute(e.begin(), int(e.size()) - 1);
    } else {
      permute(e.begin(), int(e.size()));
    }
  }
  VI used(nv, 0);
  int u = 0, w = k - 1;
  while (used[u] < int(succ[u].size())) {
    int v = succ[u][used[u]++];
    t[w++] = s[first_pos[v] + k - 2];
    u = v;
  }
  return t;
}

// the trainer's negative for one positive at objective evaluation `iter` (motif_trainer.hpp:145-152): glibc's
// generator seeded with (count of the first base) + iter.  The reference calls srand()/rand() under its input mutex;
// here the same stream comes from glibc's re-entrant interface (srand/rand ARE srandom/random on the default
// 128-byte additive-feedback state, which initstate_r/random_r reproduce on a private state), so the negatives of a
// minibatch can be generated by several host threads and are still the reference's sequences bit for bit.
inline std::string shuffled_negative(const std::string& s, int k, int iter) {
  check(s.size() < 9999, "sequence too long for negative generation:", s.size());   // const_options.hpp MAX_SEQLEN
  struct random_data rd;
  char state[128];
  std::memset(&rd, 0, sizeof rd);
  std::memset(state, 0, sizeof state);
  initstate_r(unsigned(int(std::count(s.begin(), s.end(), s.empty() ? '\0' : s[0])) + iter), state, sizeof state, &rd);
  return klet_shuffle(s, k, [&rd] { int32_t v = 0; random_r(&rd, &v); return long(v); });
}

// --------------------------------------------------------------------------------------------- motif model
// reference's pairwise log-sum-exp (util.hpp:195-209) -- the softmax normaliser must round the same way
inline double lse2(double x, double y) {
  return (-kInf == y) ? x : (-kInf == x) ? y : x < y ? y + std::log1p(std::exp(x - y)) : x + std::log1p(std::exp(y - x));
}
inline double lse(const V& v) {
  double s = -kInf;
  for (double e : v) s = lse2(s, e);
  return s;
}

// the parameters and hyper-parameters `RNAelem` holds on the host (motif_model.hpp:24-168); the automaton itself
// lives in librelem (relem_set_pattern) -- only the shape of theta is needed here
struct MotifModel {
  std::string pattern;         // as given (after the '_' -> '.' rewrite)
  bool no_rss = false, no_prf = false, no_ene = false;
  bool theta_softmax = false;
  VV s, theta;                 // row 0 = background (4), one row of 4 per '.', one row of 6 per ')'
  double lambda[2] = {0., 0.};
  double rho_s = 0., rho_theta = 0., rho_lambda = 0., tau = 0., lambda_prior = 0.;
  std::string ene_param = "~T2004~";
  int max_span = 50, max_iloop = 30;
  double min_bpp = 1e-4;

  // ProfileHMM::set_reg_pattern (profile_hmm.hpp:188-204): runs of '*' collapse, leading / trailing '*' go
  static std::string regular_pattern(const std::string& p) {
    std::string r;
    for (char c : p) if (!(c == '*' && !r.empty() && r.back() == '*')) r += c;
    size_t a = r.find_first_not_of('*');
    if (a == std::string::npos) return std::string();   // nothing but '*': the reference's erase(0, npos) leaves ""
    r.erase(0, a);
    size_t b = r.find_last_not_of('*');
    if (b != std::string::npos) r.erase(b + 1);
    return r;
  }
  std::string reg_pattern() const { return regular_pattern(pattern); }

  // set_motif_pattern (motif_model.hpp:80-97) + set_s_theta (profile_hmm.hpp:286-313): s = 0, theta uniform
  void set_pattern(const std::string& p, bool norss, bool noprf) {
    check(!p.empty(), "empty motif");
    check(!(norss && noprf), "no-rss, no-profile are exclusive.");
    pattern = p; no_rss = norss; no_prf = noprf;
    s.assign(1, V(4, 0.));
    for (char c : reg_pattern()) {
      if (c == ')') s.push_back(V(6, 0.));
      else if (c == '.') s.push_back(V(4, 0.));
    }
    if (norss) {
      check(p.find(')') == std::string::npos, "search pattern must not include pair when no-rss mode");
      std::string shown(p);
      std::replace(shown.begin(), shown.end(), '.', '_');
      cry("motif pattern:", shown);
    } else {
      cry("motif pattern:", p);
    }
    softmax();
  }
  void softmax() {   // ProfileHMM::calc_theta
    theta.clear();
    for (const V& row : s) {
      double tot = lse(row);
      V t(row.size());
      for (size_t j = 0; j < row.size(); ++j) t[j] = row[j] - tot;
      theta.push_back(t);
    }
  }
  int n_theta() const { int n = 0; for (const V& r : theta) n += int(r.size()); return n; }
  void pack(V& x) const {   // pack_params
    x.clear();
    for (const V& r : theta_softmax ? s : theta) x.insert(x.end(), r.begin(), r.end());
    x.push_back(lambda[0]); x.push_back(lambda[1]);
  }
  void unpack(const V& x) {   // unpack_params
    size_t k = 0;
    for (V& r : theta_softmax ? s : theta) for (double& v : r) v = x[k++];
    if (theta_softmax) softmax();
    lambda[0] = x[k++]; lambda[1] = x[k++];
  }
  V theta_flat() const { V t; for (const V& r : theta) t.insert(t.end(), r.begin(), r.end()); return t; }
};

inline VV exp_rows(const VV& a) {
  VV b(a);
  for (V& r : b) for (double& v : r) v = std::exp(v);
  return b;
}

// train.model (RNAelemWriter::write, motif_io.hpp:29-57) and the one-line train.interim record (58-87)
inline void write_model(OutputSet& out, int id, MotifModel& m) {
  std::string p = m.reg_pattern();
  if (m.no_rss) std::replace(p.begin(), p.end(), '.', '_');
  out.dat(id, "pattern:", p);
  if (m.theta_softmax) { out.dat(id, "s:", m.s); m.softmax(); }
  else out.dat(id, "theta:", m.theta);
  out.dat(id, "exp-theta:", exp_rows(m.theta));
  out.dat(id, "ene-param:", m.ene_param);
  out.dat(id, "max-span:", m.max_span);
  out.dat(id, "max-internal-loop:", m.max_iloop);
  out.dat(id, "theta-softmax:", m.theta_softmax);
  if (m.theta_softmax) out.dat(id, "rho-s:", m.rho_s);
  else out.dat(id, "rho-theta:", m.rho_theta);
  out.dat(id, "rho-lambda:", m.rho_lambda);
  out.dat(id, "tau:", m.tau);
  out.dat(id, "lambda:", V{m.lambda[0], m.lambda[1]});
  out.dat(id, "lambda-prior:", m.lambda_prior);
  out.dat(id, "min-bpp:", m.min_bpp);
  out.dat(id, "no-rss:", m.no_rss);
  out.dat(id, "no-profile:", m.no_prf);
  out.dat(id, "no-energy:", m.no_ene);
}

inline void write_interim(OutputSet& out, int id, MotifModel& m) {
  std::string p = m.pattern;   // the unregularised pattern here (motif_io.hpp:60)
  if (m.no_rss) std::replace(p.begin(), p.end(), '.', '_');
  if (m.theta_softmax) m.softmax();
  std::vector<std::string> f = {
      glue("pattern:", p),
      m.theta_softmax ? glue("s:", m.s) : glue("theta:", m.theta),
      glue("exp-theta:", exp_rows(m.theta)),
      glue("ene-param:", m.ene_param),
      glue("max-span:", m.max_span),
      glue("max-internal-loop:", m.max_iloop),
      glue("theta-softmax:", m.theta_softmax),
      m.theta_softmax ? glue("rho-s:", m.rho_s) : glue("rho-theta:", m.rho_theta),
      glue("rho-lambda:", m.rho_lambda),
      glue("tau:", m.tau),
      glue("lambda:", V{m.lambda[0], m.lambda[1]}),
      glue("lambda-prior:", m.lambda_prior),
      glue("min-bpp:", m.min_bpp),
      glue("no-rss:", m.no_rss),
      glue("no-profile:", m.no_prf),
      glue("no-energy:", m.no_ene)};
  std::string line;
  for (size_t k = 0; k < f.size(); ++k) line += (k ? " " : "") + f[k];
  out.dat(id, "interim:", line);
}

// RNAelemReader::read_model (motif_io.hpp:118-262)
namespace detail {
inline std::string strip(const std::string& s, const std::string& drop = " \n") {
  size_t a = s.find_first_not_of(drop), b = s.find_last_not_of(drop);
  return (a == std::string::npos || b == std::string::npos || b < a) ? std::string() : s.substr(a, b - a + 1);
}
template <class T>
T parse(const std::string& s) {
  if (s.empty()) return T();
  T v = T();
  std::istringstream in(s);
  in >> v;
  return v;
}
inline V parse_list(const std::string& s) {   // "a,b,c"
  V out;
  if (strip(s).empty()) return out;
  size_t a = 0, b;
  while ((b = s.find(',', a)) != std::string::npos) { out.push_back(parse<double>(s.substr(a, b - a))); a = b + 1; }
  out.push_back(parse<double>(s.substr(a)));
  return out;
}
inline VV parse_rows(const std::string& s) {   // "[[..],[..]]"
  VV rows;
  size_t lo = s.find_first_of('['), hi = s.find_last_of(']');
  if (lo == std::string::npos || hi == std::string::npos) return rows;
  size_t open = std::string::npos;
  for (size_t k = lo + 1; k < hi; ++k) {
    if (s[k] == '[') open = k;
    else if (s[k] == ']' && open != std::string::npos) rows.push_back(parse_list(s.substr(open + 1, k - open - 1)));
  }
  return rows;
}
}  // namespace detail

inline void read_model(const std::string& path, MotifModel& m) {
  std::ifstream in(path);
  check(!!in, "couldn't open:", path);
  VV w;
  std::string pattern;
  V lam{0., 0.};
  bool norss = false, noprf = false;
  unsigned seen = 0;
  while (!in.eof()) {
    std::string line;
    std::getline(in, line);
    std::vector<std::string> kv;
    for (size_t a = 0, b;; a = b + 2) {
      b = line.find(": ", a);
      kv.push_back(line.substr(a, b == std::string::npos ? b : b - a));
      if (b == std::string::npos) break;
    }
    if (kv.size() < 2) continue;
    check(kv.size() == 2, "fail to parse:", path);
    const std::string key = detail::strip(kv[0]), &val = kv[1];
    if (key == "pattern") { pattern = detail::strip(val); seen |= 1u << 0; }
    else if (key == "s" || key == "theta") { w = detail::parse_rows(val); seen |= 1u << 1; }
    else if (key == "ene-param") { m.ene_param = detail::strip(val); seen |= 1u << 2; }
    else if (key == "max-span") { m.max_span = detail::parse<int>(val); seen |= 1u << 3; }
    else if (key == "rho-s") { m.rho_s = detail::parse<double>(val); seen |= 1u << 4; }
    else if (key == "rho-theta") { m.rho_theta = detail::parse<double>(val); seen |= 1u << 4; }
    else if (key == "rho-lambda") { m.rho_lambda = detail::parse<double>(val); seen |= 1u << 5; }
    else if (key == "tau") { m.tau = detail::parse<double>(val); seen |= 1u << 6; }
    else if (key == "lambda") {
      size_t a = val.find_first_of('['), b = val.find_last_of(']');
      lam = detail::parse_list(val.substr(a + 1, b - a - 1));
      seen |= 1u << 7;
    }
    else if (key == "lambda-prior") m.lambda_prior = detail::parse<double>(val);
    else if (key == "min-bpp") { m.min_bpp = detail::parse<double>(val); seen |= 1u << 8; }
    else if (key == "max-internal-loop") { m.max_iloop = detail::parse<int>(val); seen |= 1u << 9; }
    else if (key == "no-rss") norss = detail::parse<bool>(val);
    else if (key == "no-profile") noprf = detail::parse<bool>(val);
    else if (key == "no-energy") m.no_ene = detail::parse<bool>(val);
    else if (key == "exp-theta") {}
    else if (key == "theta-softmax") { m.theta_softmax = detail::parse<bool>(val); seen |= 1u << 10; }
    else cry("unused:", key);
  }
  check(seen == (1u << 11) - 1, "motif file broken:", path);
  if (norss) std::replace(pattern.begin(), pattern.end(), '_', '.');
  m.set_pattern(pattern, norss, noprf);
  for (size_t i = 0; i < m.s.size(); ++i)
    for (size_t j = 0; j < m.s[i].size(); ++j) {
      check(i < w.size() && j < w[i].size(), "motif file broken:", path);
      (m.theta_softmax ? m.s : m.theta)[i][j] = w[i][j];
    }
  if (m.theta_softmax) m.softmax();
  check(lam.size() >= 2, "motif file broken:", path);
  m.lambda[0] = lam[0]; m.lambda[1] = lam[1];
}

// ---------------------------------------------------------------------------------------------------- Adam
// optimizer.hpp:72-173, including its bias-correction schedule (the powers of beta start one step ahead: the first
// update divides by 1-beta^2), L2 regularisation added to y and gr before the update and clipping to the bounds after
class Adam {
  double alpha_ = 0.001, beta1_ = 0.9, beta2_ = 0.999, eps_ = 1e-8, m0_ = 0., v0_ = 0.;
  V m_, v_, x_, lo_, hi_, rho_;
  VI bound_, rgl_;
  int t_ = 0;
 public:
  void set_hp(double m0, double v0, double alpha, double beta1, double beta2, double eps) {
    m0_ = m0; v0_ = v0; alpha_ = alpha; beta1_ = beta1; beta2_ = beta2; eps_ = eps;
  }
  void set_bounds(const V& lo, const V& hi, const VI& type) { lo_ = lo; hi_ = hi; bound_ = type; }
  void set_regularization(const VI& type, const V& rho) { rgl_ = type; rho_ = rho; }
  double rgl_term(const V& x) const {
    double r = 0.;
    for (size_t i = 0; i < x.size() && i < rgl_.size(); ++i) {
      if (rgl_[i] == 1) r += rho_[i] * std::abs(x[i]);
      else if (rgl_[i] == 2) r += rho_[i] * x[i] * x[i] / 2.;
    }
    return r;
  }
  int itercount() const { return t_ - 1; }
  const V& x() const { return x_; }
  template <class F>
  void minimize(F& f, const V& x0, int max_iter) {
    const size_t n = x0.size();
    t_ = 0;
    m_.assign(n, m0_); v_.assign(n, v0_);
    x_ = x0;
    lo_.resize(n, -kInf); hi_.resize(n, kInf); bound_.resize(n, 0); rgl_.resize(n, 0); rho_.resize(n, 0.);
    V gr(n, 0.);
    double y = 0., b1t = beta1_, b2t = beta2_, gg;
    do {
      ++t_;
      f(x_, y, gr);
      double pen = 0.;
      for (size_t i = 0; i < n; ++i) {
        if (rgl_[i] == 1) { pen += rho_[i] * std::abs(x_[i]); gr[i] += rho_[i] * (0 < x_[i] ? 1 : -1); }
        else if (rgl_[i] == 2) { pen += rho_[i] * x_[i] * x_[i] / 2.; gr[i] += rho_[i] * x_[i]; }
      }
      y += pen;
      b1t *= beta1_; b2t *= beta2_;
      for (size_t i = 0; i < n; ++i) {
        m_[i] += (1. - beta1_) * (gr[i] - m_[i]);
        v_[i] += (1. - beta2_) * (gr[i] * gr[i] - v_[i]);
        double mhat = m_[i] / (1. - b1t), vhat = v_[i] / (1. - b2t);
        x_[i] -= alpha_ * mhat / (std::sqrt(vhat) + eps_);
      }
      for (size_t i = 0; i < n; ++i) {
        if ((bound_[i] & 1) && x_[i] < lo_[i]) x_[i] = lo_[i];
        if ((bound_[i] & 2) && hi_[i] < x_[i]) x_[i] = hi_[i];
      }
      gg = 0.;
      for (double g : gr) gg += g * g;   // the reference's "norm2" is the SUM of squares (util.hpp:253-258)
    } while (!(gg < (y + 1.) * 1.e-8) && t_ < max_iter);
  }
};

// f(i) for i in [0, n) on a few host threads (f must only touch slot i)
template <class F>
void parallel_for(int n, F f) {
  const int nt = int(std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency())));
  if (n < 64 || nt == 1) { for (int i = 0; i < n; ++i) f(i); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t)
    th.emplace_back([=] { for (int i = t; i < n; i += nt) f(i); });
  for (auto& x : th) x.join();
}

// contiguous shard of `total` items for worker k of n (ArrayJobManager::assigned_range, arrayjob_manager.hpp:141-149)
inline void shard_range(long total, int n, int k, long& from, long& to) {
  long q = total / n, r = total - q * n;
  from = k * q + std::min<long>(k, r);
  to = from + q + (k < r ? 1 : 0);
}

}  // namespace relem
#endif
