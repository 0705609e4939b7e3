// integration/relem_host.hpp -- the reference-side binding of librelem: what a maintainer of iyak/RNAelem adds to
// run the per-sequence inside / outside / Viterbi hot path on a B200 and leave everything else as it is.
//
// Two call sites change in the reference (recipe: oracle/Makefile target `gpu`, which applies integration/trainer.sed
// and integration/scanner.sed to build-time copies of the two headers and compiles the otherwise unmodified main.cpp
// into oracle/_ref/RNAelem_gpu):
//
//   RNAelemTrainer::operator()   motif_trainer.hpp:616-621   ClassThread<RNAelemTrainDP> ct(...); ct(fn,gr);
//       -> relem_host::estep(*_motif,_qr,_mode,_cnt,_kmer_shuf,_from,_to,_sum_eff,fn,gr);
//   RNAelemScanner::scan         motif_scanner.hpp:943-946   ClassThread<RNAelemScanDP> ct(...); ct(_EN);
//       -> relem_host::scan(*_motif,_qr,_out,_EN);
//
// Everything around them keeps running on the host unchanged: option parsing, FastqBatchReader (minibatch order,
// epoch shuffles), Adam AND L-BFGS-B (optimizer.hpp), regularisation and bounds, train.model / train.interim writers,
// the `dat`/`cry` output channels.  This file (context + scan) and relem_host_train.hpp (E-step) use the reference's own types (RNAelem, FastqBatchReader, V, VV, VI,
// ushuffle) -- they are compiled as part of the reference -- and only the C ABI of include/relem.h on the other side.
#ifndef RELEM_HOST_HPP
#define RELEM_HOST_HPP
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

extern "C" {
#include "relem.h"
}

namespace iyak {
namespace relem_host {

struct Binding {
  relem_ctx* ctx = nullptr;
  std::string model_key;   // energy / pattern configuration the context currently holds
  ~Binding() { if (ctx) relem_destroy(ctx); }
};
inline Binding& binding() { static Binding b; return b; }

inline void ok(int rc, const char* what) {
  if (rc != RELEM_OK) die("relem:", what, "failed:", relem_last_error(binding().ctx));
}

// make the context reflect the model: energy set / pattern when they changed, parameters on every call
inline relem_ctx* context(RNAelem& m) {
  Binding& b = binding();
  if (!b.ctx) {
    int dev = 0;
    if (const char* e = std::getenv("RELEM_DEVICE")) dev = std::atoi(e);
    if (relem_create(&b.ctx, dev) != RELEM_OK) die("relem:", relem_last_error(nullptr));
  }
  std::string key = paste1(m.em.param_fname, m.em.max_pair(), m.em.max_iloop(), m.em.min_BPP(), m.em.no_ene(),
                           m.mm.pattern(), m.no_rss(), m.no_prf());
  if (key != b.model_key) {
    ok(relem_set_energy(b.ctx, m.em.param_fname.c_str(), m.em.max_pair(), m.em.max_iloop(), m.em.min_BPP(),
                        m.em.no_ene()), "relem_set_energy");
    ok(relem_set_pattern(b.ctx, m.mm.pattern().c_str(), m.no_rss(), m.no_prf()), "relem_set_pattern");
    b.model_key = key;
  }
  V theta;
  for (auto& row : m.mm.theta()) theta.insert(theta.end(), row.begin(), row.end());
  double lam[2] = {m._lambda[0], m._lambda[1]};
  ok(relem_set_params(b.ctx, theta.data(), (int)theta.size(), lam, m.tau()), "relem_set_params");
  return b.ctx;
}

struct Packed {
  std::vector<uint8_t> seq, kind;
  std::vector<int64_t> off{0};
  V ws;
  std::vector<int32_t> gate;
  std::vector<std::string> id;
  int n() const { return (int)kind.size(); }
  void add(VI const& codes, V const& w, int kd, int g, std::string const& name) {
    for (int c : codes) seq.push_back((uint8_t)c);
    ws.insert(ws.end(), w.begin(), w.begin() + codes.size());
    off.push_back((int64_t)seq.size());
    kind.push_back((uint8_t)kd); gate.push_back(g); id.push_back(name);
  }
};

// body of RNAelemScanner::scan's thread fan-out (motif_scanner.hpp:943-946): all reads through relem_scan, records
// written in input order with the reference's own `dat` (motif_scanner.hpp:240-251)
template <class Reader>
inline void scan(RNAelem& m, Reader& qr, int out, VV& ENg) {
  relem_ctx* c = context(m);
  const long chunk = 4096;
  const int M = m.M;
  std::string id, rss;
  VI seq, qual;
  while (!qr.is_end()) {
    Packed b;
    std::vector<VI> seqs;
    while (!qr.is_end() && b.n() < chunk) {
      qr.get_read(id, seq, qual, rss);
      m.set_ws(qual);
      b.add(seq, m._ws, RELEM_POS_WITHOUT, -1, id);
      seqs.push_back(seq);
    }
    const int ns = b.n();
    const size_t tl = b.seq.size();
    V ps(tl), pe(tl + ns), pi(tl), exist(ns), en;
    for (auto& r : ENg) en.insert(en.end(), r.size(), 0.);
    std::vector<int32_t> psihat(tl), ys(ns), ye(ns);
    std::string rs(tl + 1, ' ');
    relem_scan_out o;
    std::memset(&o, 0, sizeof o);
    o.PysL = ps.data(); o.PyeL = pe.data(); o.PyiL = pi.data(); o.psihat = psihat.data(); o.rss = &rs[0];
    o.Ys = ys.data(); o.Ye = ye.data(); o.exist_prob = exist.data(); o.EN = en.data();
    ok(relem_scan(c, ns, b.seq.data(), b.off.data(), b.ws.data(), &o), "relem_scan");
    for (int n = 0; n < ns; ++n) {
      size_t a = (size_t)b.off[n], L = (size_t)b.off[n + 1] - a;
      VI psi(psihat.begin() + a, psihat.begin() + a + L);
      dat(out, "id:", b.id[n]);
      dat(out, "start:", V(ps.begin() + a, ps.begin() + a + L));
      dat(out, "end:", V(pe.begin() + a + n, pe.begin() + a + n + L + 1));
      dat(out, "inner:", V(pi.begin() + a, pi.begin() + a + L));
      dat(out, "psihat:", psi);
      dat(out, "motif region:", ys[n], "-", ye[n]);
      dat(out, "exist prob:", exist[n]);
      dat(out, "seq:", seq_int2str(seqs[n]));
      dat(out, "rss:", rs.substr(a, L));
      std::string s = "";
      for (int h : psi) s += (0 == h or M - 1 == h ? ' ' : m.mm.node(h));
      dat(out, "mot:", s);
    }
    size_t k = 0;
    for (auto& r : ENg) for (auto& x : r) x += en[k++];
  }
}

}  // namespace relem_host
}  // namespace iyak
#endif
