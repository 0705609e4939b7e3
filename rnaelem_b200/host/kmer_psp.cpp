// kmer_psp.cpp -- the preprocessing step of the `elem` pipeline at GPU-box scale: what
// script/kmer-psp.py does (k-mer enrichment between a positive and a negative FASTA by Fisher's exact test, the
// enriched / depleted k-mers turned into per-base pseudo-qualities of the FASTQ that `RNAelem train` reads), as a
// multi-threaded C++14 host tool.  The Python script walks every (sequence, significant k-mer) pair with a regular
// expression and calls scipy's Fisher test once per k-mer: hours at 1 M sequences, i.e. longer than the training it
// feeds once the DP runs on the GPU.  Same command line, same stdout (byte for byte) and the same stderr lines:
//
//     kmer-psp [-t threads] <positive.fa> [<negative.fa>]  > positive.fq
//
// Semantics kept from script/kmer-psp.py (line numbers of that script):
//   * FASTA = alternating header / sequence lines, both stripped; reading stops at the first empty one (:11-19);
//   * a sequence contributes the set of its k-mers at offsets 0 .. len-k-1 -- the last window is NOT counted (:23,:30);
//   * k-mers seen in both sets get a two-sided Fisher exact p on [[nP, nN], [nT-nP, nF-nN]] (:40-41); enriched
//     (nN < nP) or depleted (otherwise) when p < 0.05 (:42-51);
//   * k = the length in 3..10 whose most significant ENRICHED k-mer has the smallest p (strictly smaller wins, :75-81);
//     k = -1 (no enriched k-mer at any length) gives flat qualities;
//   * every non-overlapping occurrence (left to right, like re.finditer, :56-61 -- here ALL windows count, the last one
//     included) of an enriched k-mer adds 1 to the bases it covers, of a depleted one subtracts 1; base quality 10;
//   * quality characters: chr(clamp(33 + q, 33, 126)), and a final '!' = "this read contains the motif" (:62-67).
// Two-sided p as scipy.stats.fisher_exact computes it: the sum of the hypergeometric probabilities that do not exceed
// the observed one (relative tolerance 1e-7 for ties), 1 when a margin is empty.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

struct Record { std::string ann, seq; };

std::string strip(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && std::isspace((unsigned char)s[a])) ++a;
  while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
  return s.substr(a, b - a);
}

std::vector<Record> parse_fa(const std::string& path) {
  std::ifstream f(path.c_str());
  if (!f) { std::fprintf(stderr, "kmer-psp: cannot open %s\n", path.c_str()); std::exit(1); }
  std::vector<Record> v;
  std::string a, s;
  for (;;) {
    if (!std::getline(f, a)) break;
    if (!std::getline(f, s)) s.clear();
    a = strip(a); s = strip(s);
    if (a.empty() || s.empty()) break;
    v.push_back(Record{a, s});
  }
  return v;
}

// k-mers as integers: 6 bits per character, characters numbered in order of first appearance (1..63)
struct Alphabet {
  int code[256];
  char sym[64];
  int n = 0;
  Alphabet() { std::fill(code, code + 256, 0); }
  void learn(const std::string& s) {
    for (unsigned char c : s)
      if (!code[c]) {
        if (n >= 63) { std::fprintf(stderr, "kmer-psp: more than 63 distinct sequence characters\n"); std::exit(1); }
        code[c] = ++n; sym[n] = (char)c;
      }
  }
  uint64_t encode(const char* p, int k) const {
    uint64_t v = 0;
    for (int t = 0; t < k; ++t) v |= (uint64_t)code[(unsigned char)p[t]] << (6 * t);
    return v;
  }
  std::string decode(uint64_t v, int k) const {
    std::string s;
    for (int t = 0; t < k; ++t) s += sym[(v >> (6 * t)) & 63];
    return s;
  }
};

typedef std::unordered_map<uint64_t, uint32_t> Counts;

template <class F> void parallel_chunks(size_t n, int threads, F f) {
  threads = (int)std::max<size_t>(1, std::min<size_t>((size_t)threads, (n + 255) / 256));
  if (threads == 1) { f(0, 0, n); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < threads; ++t) th.emplace_back([=] { f(t, n * t / threads, n * (t + 1) / threads); });
  for (auto& x : th) x.join();
}

// number of sequences that contain each k-mer (windows 0 .. len-k-1)
Counts count_kmers(const std::vector<Record>& recs, const Alphabet& ab, int k, int threads) {
  std::vector<Counts> part((size_t)threads);
  parallel_chunks(recs.size(), threads, [&](int t, size_t a, size_t b) {
    Counts& c = part[(size_t)t];
    std::vector<uint64_t> w;
    for (size_t r = a; r < b; ++r) {
      const std::string& s = recs[r].seq;
      w.clear();
      for (long i = 0; i < (long)s.size() - k; ++i) w.push_back(ab.encode(s.data() + i, k));
      std::sort(w.begin(), w.end());
      w.erase(std::unique(w.begin(), w.end()), w.end());
      for (uint64_t v : w) ++c[v];
    }
  });
  Counts tot;
  for (Counts& c : part) {
    if (tot.empty()) { tot.swap(c); continue; }
    for (auto& kv : c) tot[kv.first] += kv.second;
  }
  return tot;
}

// two-sided Fisher exact p of [[a, b], [c, d]]: the hypergeometric probabilities that do not exceed the observed one
// (ties within 1e-7 relative).  One lgamma evaluation for the observed table, then the neighbouring terms by the ratio
// pmf(x+1) / pmf(x) = (n1-x)(n-x) / ((x+1)(n2-n+x+1)) until they no longer matter: O(standard deviation) steps.
double fisher_two_sided(long a, long b, long c, long d) {
  const long n1 = a + b, n2 = c + d, n = a + c, N = n1 + n2;
  if (n1 == 0 || n2 == 0 || n == 0 || b + d == 0) return 1.0;
  const long lo = std::max(0L, n - n2), hi = std::min(n, n1);
  auto lchoose = [](long nn, long kk) { return std::lgamma((long double)nn + 1) - std::lgamma((long double)kk + 1) - std::lgamma((long double)(nn - kk) + 1); };
  const long double pobs = std::exp(lchoose(n1, a) + lchoose(n2, n - a) - lchoose(N, n));
  const long double cut = pobs * (1.0L + 1e-7L), tiny = pobs * 1e-22L;
  auto up = [&](long x, long double q) { return q * (long double)(n1 - x) * (long double)(n - x) / ((long double)(x + 1) * (long double)(n2 - n + x + 1)); };     // pmf(x+1)
  auto down = [&](long x, long double q) { return q * (long double)x * (long double)(n2 - n + x) / ((long double)(n1 - x + 1) * (long double)(n - x + 1)); };     // pmf(x-1)
  long double p = pobs;
  // the pmf is unimodal: in each direction it either falls at once (all terms count until negligible) or first rises
  // above the observed value (terms skipped) and counts again once it has fallen back below it
  {
    long double q = pobs;
    bool beyond = false;   // past the hump on this side
    for (long x = a; x < hi; ++x) {
      const long double q1 = up(x, q);
      if (q1 <= q) beyond = true;
      q = q1;
      if (q <= cut) { p += q; if (beyond && q < tiny) break; }
    }
  }
  {
    long double q = pobs;
    bool beyond = false;
    for (long x = a; x > lo; --x) {
      const long double q1 = down(x, q);
      if (q1 <= q) beyond = true;
      q = q1;
      if (q <= cut) { p += q; if (beyond && q < tiny) break; }
    }
  }
  return (double)std::min<long double>(p, 1.0L);
}

struct Sig { uint64_t code; double p; };

void calc_pval(const Counts& nP, const Counts& nN, long nT, long nF, const Alphabet& ab, int k, double thresh, int threads,
               std::vector<Sig>& rich, std::vector<Sig>& poor) {
  std::vector<std::pair<uint64_t, uint32_t>> both;
  for (auto& kv : nP) if (nN.count(kv.first)) both.push_back(kv);
  std::vector<double> pv(both.size());
  parallel_chunks(both.size(), threads, [&](int, size_t a, size_t b) {
    for (size_t t = a; t < b; ++t) {
      long p1 = both[t].second, n1 = nN.at(both[t].first);
      pv[t] = fisher_two_sided(p1, n1, nT - p1, nF - n1);
    }
  });
  rich.clear(); poor.clear();
  for (size_t t = 0; t < both.size(); ++t) {
    if (!(pv[t] < thresh)) continue;
    const bool up = nN.at(both[t].first) < both[t].second;
    std::fprintf(stderr, "%c%s\t%f\n", up ? '+' : '-', ab.decode(both[t].first, k).c_str(), pv[t]);
    (up ? rich : poor).push_back(Sig{both[t].first, pv[t]});
  }
}

}  // namespace

int main(int argc, char** argv) {
  int threads = (int)std::max(1u, std::thread::hardware_concurrency());
  std::vector<std::string> pos_args;
  for (int a = 1; a < argc; ++a) {
    if (!std::strcmp(argv[a], "-t") && a + 1 < argc) threads = std::max(1, std::atoi(argv[++a]));
    else pos_args.push_back(argv[a]);
  }
  if (pos_args.empty() || pos_args.size() > 2) {
    std::fprintf(stderr, "usage: kmer-psp [-t threads] <positive.fa> [<negative.fa>]\n");
    return 1;
  }
  const int kmin = 3, kmax = 10, base = 10;
  const double thresh = 5e-2;
  std::vector<Record> P = parse_fa(pos_args[0]), N;
  Alphabet ab;
  for (auto& r : P) ab.learn(r.seq);
  int k = -1;
  std::vector<Sig> rich, poor;
  if (pos_args.size() == 2) {
    N = parse_fa(pos_args[1]);
    for (auto& r : N) ab.learn(r.seq);
    const long nT = (long)P.size(), nF = (long)N.size();
    double min_pval = 1.;
    for (int kk = kmin; kk <= kmax; ++kk) {
      Counts nP = count_kmers(P, ab, kk, threads), nN = count_kmers(N, ab, kk, threads);
      calc_pval(nP, nN, nT, nF, ab, kk, thresh, threads, rich, poor);
      if (rich.empty()) continue;
      double p = rich[0].p;
      for (auto& s : rich) p = std::min(p, s.p);
      if (p < min_pval) { k = kk; min_pval = p; }
    }
    std::fprintf(stderr, "k:%d\n", k);
    rich.clear(); poor.clear();
    if (k > 0) {
      Counts nP = count_kmers(P, ab, k, threads), nN = count_kmers(N, ab, k, threads);
      calc_pval(nP, nN, nT, nF, ab, k, thresh, threads, rich, poor);
    }
  }
  std::unordered_set<uint64_t> up, down;
  for (auto& s : rich) up.insert(s.code);
  for (auto& s : poor) down.insert(s.code);

  // records formatted in parallel, written in input order
  std::vector<std::string> out(P.size());
  parallel_chunks(P.size(), threads, [&](int, size_t a, size_t b) {
    std::vector<int> q;
    std::unordered_map<uint64_t, long> next_ok;
    for (size_t r = a; r < b; ++r) {
      const std::string& s = P[r].seq;
      q.assign(s.size(), base);
      if (k > 0 && (!up.empty() || !down.empty())) {
        next_ok.clear();
        for (long i = 0; i + k <= (long)s.size(); ++i) {
          const uint64_t v = ab.encode(s.data() + i, k);
          const int delta = up.count(v) ? 1 : down.count(v) ? -1 : 0;
          if (!delta) continue;
          long& nx = next_ok[v];   // 0 for a k-mer not yet matched in this sequence
          if (i < nx) continue;    // overlaps the previous occurrence of the same k-mer: re.finditer skips it
          nx = i + k;
          for (int t = 0; t < k; ++t) q[(size_t)(i + t)] += delta;
        }
      }
      std::string& o = out[r];
      o.reserve(P[r].ann.size() + 2 * s.size() + 8);
      o += '@'; o += P[r].ann.substr(1); o += '\n'; o += s; o += "\n+\n";
      for (int v : q) o += (char)std::max(33, std::min(126, 33 + v));
      o += "!\n";
    }
  });
  for (auto& o : out) std::fwrite(o.data(), 1, o.size(), stdout);
  return 0;
}
