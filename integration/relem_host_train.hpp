// integration/relem_host_train.hpp -- E-step half of the reference-side binding (see relem_host.hpp): included by the
// patched motif_trainer.hpp after its TR_* mode bits and the ushuffle declarations.
#ifndef RELEM_HOST_TRAIN_HPP
#define RELEM_HOST_TRAIN_HPP
#include "relem_host.hpp"

namespace iyak {
namespace relem_host {

// body of RNAelemTrainer::operator()'s thread fan-out (motif_trainer.hpp:616-621): the reads of the minibatch in
// reader order, each followed by the negative RNAelemTrainDP shuffles from it (:145-152), one relem_estep call, then
// the update block of :248-271 (softmax chain rule, lambda slots)
template <class Reader>
inline void estep(RNAelem& m, Reader& qr, unsigned mode, int iter_cnt, int kmer_shuf, int from, int to,
                  double& sum_eff, double& fn, V& gr) {
  relem_ctx* c = context(m);
  const bool shuffle = !(mode & TR_NO_SHUFFLE), lr = (mode & TR_LIK_RATIO) != 0;
  Packed b;
  std::string id, rss;
  VI seq, qual, neg;
  while (!qr.is_end()) {
    qr.get_read(id, seq, qual, rss);
    check(size(seq) + 1 == size(qual), "bad seq format.", id, size(seq), size(qual));
    if (mode & TR_ARRAYEVAL) {
      if (qr.cnt() < from + 1) continue;
      if (to + 1 <= qr.cnt()) break;
    }
    m.set_ws(qual);   // log position weights + the "contains motif" flag in the last entry (motif_model.hpp:62-70)
    const bool with_motif = !(-inf < m._ws.back());
    const int me = b.n();
    b.add(seq, m._ws, with_motif ? RELEM_POS_WITH : lr ? RELEM_LR_WITHOUT : RELEM_POS_WITHOUT, -1, id);
    if (shuffle) {
      std::string s; seq_int2str(seq, s);
      srand((int)count(s.begin(), s.end(), s[0]) + iter_cnt);
      ushuffle::set_randfunc(long_rand);
      char neg_s[MAX_SEQLEN] = "";
      ushuffle::shuffle(s.c_str(), neg_s, size(s), kmer_shuf);
      seq_str2int(neg_s, neg);
      b.add(neg, V(size(neg) + 1, 0.), lr ? RELEM_LR_NEG : RELEM_NEG, me, id);   // qualities all 0 -> weights ln 1
    }
  }
  const int ns = b.n();
  V en((size_t)[&] { size_t t = 0; for (auto& r : m.mm.theta()) t += r.size(); return t; }(), 0.);
  std::vector<uint8_t> skipped(ns, 0);
  relem_estep_out o;
  std::memset(&o, 0, sizeof o);
  o.EN_diff = en.data();
  o.skipped = skipped.data();
  if (ns > 0)
    ok(relem_estep(c, ns, b.seq.data(), b.off.data(), b.ws.data(), b.kind.data(), b.gate.data(), &o), "relem_estep");
  if (0 == iter_cnt)
    for (int n = 0; n < ns; ++n) if (skipped[n] == 1) cry("skipped:", b.id[n]);
  // update block (motif_trainer.hpp:248-271)
  int k = 0;
  size_t e = 0;
  if (m.theta_softmax()) {
    for (auto& row : m.mm.theta()) {
      double tot = 0.;
      for (size_t j = 0; j < row.size(); ++j) tot += en[e + j];
      for (size_t j = 0; j < row.size(); ++j) {
        double tmp = en[e + j], p = exp(row[j]);
        gr[k++] += (1 - p) * tmp - p * (tot - tmp);
      }
      e += row.size();
    }
  } else {
    for (size_t j = 0; j < en.size(); ++j) gr[k++] += en[j];
  }
  // energy counts are filed by lambda VALUE in the reference (:380-381): equal lambdas share the first slot
  if (m._lambda[0] == m._lambda[1]) { gr[k] += o.EH_diff[0] + o.EH_diff[1]; }
  else { gr[k] += o.EH_diff[0]; gr[k + 1] += o.EH_diff[1]; }
  fn += o.fn;
  sum_eff += o.sum_eff;
}

}  // namespace relem_host
}  // namespace iyak
#endif
