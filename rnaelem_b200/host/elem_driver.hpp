// elem_driver.hpp -- RNAelemTrainer / RNAelemScanner on top of librelem: the two seams of SURVEY.md 8(b).
//
//   RNAelemTrainer::operator()(x, fn, gr)   motif_trainer.hpp:595-633   objective functor handed to Adam
//   RNAelemTrainer::train(model)            motif_trainer.hpp:562-593
//   RNAelemScanner::scan(model)             motif_scanner.hpp:938-949   + record formatting 237-252
//
// The minibatch is packed on the host in reading order (positive, then the negative shuffled from it), evaluated by
// relem_estep on one GPU -- or sharded contiguously over the GPUs of the box, one host thread and one context per
// GPU, with relem_allreduce_sum (NCCL) adding the P+3 partial sums -- and the softmax chain rule, the lambda slots and
// the regulariser are applied on the host exactly where the reference applies them.
#ifndef RELEM_ELEM_DRIVER_HPP
#define RELEM_ELEM_DRIVER_HPP
#include <chrono>
#include <cstdint>
#include <cstring>
#include <future>
#include <thread>

#include "elem_host.hpp"
extern "C" {
#include "relem.h"
}

namespace relem {

enum TrainMode : unsigned { TR_NORMAL = 0, TR_NO_SHUFFLE = 1u << 5, TR_LIK_RATIO = 1u << 6 };

// one context per GPU of this process; rank 0 is the context whose results the host uses
class DeviceGroup {
  std::vector<relem_ctx*> ctx_;
 public:
  DeviceGroup() = default;
  DeviceGroup(const DeviceGroup&) = delete;
  DeviceGroup& operator=(const DeviceGroup&) = delete;
  ~DeviceGroup() { for (relem_ctx* c : ctx_) relem_destroy(c); }
  int size() const { return int(ctx_.size()); }
  relem_ctx* ctx(int k) { return ctx_[k]; }
  void ok(int k, int rc, const char* what) {
    if (rc != RELEM_OK) die("relem:", what, "failed:", relem_last_error(ctx_[k]));
  }
  void open(int ngpu) {
    check(ngpu >= 1, "need at least one GPU");
    for (int k = 0; k < ngpu; ++k) {
      relem_ctx* c = nullptr;
      if (relem_create(&c, k) != RELEM_OK) die("relem:", relem_last_error(nullptr));
      ctx_.push_back(c);
    }
    if (ngpu > 1) {
      uint8_t id[128];
      if (relem_comm_unique_id(id) != RELEM_OK) die("relem: relem_comm_unique_id failed");
      each([&](int k) { ok(k, relem_comm_init(ctx_[k], id, k, ngpu), "relem_comm_init"); });
    }
  }
  // run f(rank) on one host thread per context (a context is bound to its thread's GPU; the collective needs all
  // ranks in flight at once)
  template <class F>
  void each(F f) {
    if (ctx_.size() == 1) { f(0); return; }
    std::vector<std::thread> th;
    std::vector<std::string> err(ctx_.size());
    for (int k = 0; k < int(ctx_.size()); ++k)
      th.emplace_back([&, k] { try { f(k); } catch (std::exception& e) { err[k] = e.what(); } });
    for (auto& t : th) t.join();
    for (auto& e : err) if (!e.empty()) throw std::runtime_error(e);
  }
  void set_model(const MotifModel& m) {
    each([&](int k) {
      ok(k, relem_set_energy(ctx_[k], m.ene_param.c_str(), m.max_span, m.max_iloop, m.min_bpp, m.no_ene), "relem_set_energy");
      ok(k, relem_set_pattern(ctx_[k], m.pattern.c_str(), m.no_rss, m.no_prf), "relem_set_pattern");
    });
    int nt = 0;
    ok(0, relem_model_dims(ctx_[0], nullptr, nullptr, nullptr, &nt), "relem_model_dims");
    check(nt == m.n_theta(), "relem: theta shape mismatch", nt, m.n_theta());
    push_params(m);
  }
  void push_params(const MotifModel& m) {
    V th = m.theta_flat();
    each([&](int k) { ok(k, relem_set_params(ctx_[k], th.data(), int(th.size()), m.lambda, m.tau), "relem_set_params"); });
  }
  std::string node_chars() {
    int n = relem_hmm_get(ctx_[0], 6, nullptr);
    check(n > 0, "relem: relem_hmm_get failed");
    VI c(n);
    relem_hmm_get(ctx_[0], 6, c.data());
    return std::string(c.begin(), c.end());
  }
};

// a packed minibatch in the layout relem_batch_create takes
struct PackedBatch {
  std::vector<uint8_t> seq, kind;
  std::vector<int64_t> off{0};
  V ws;
  std::vector<int32_t> gate;
  std::vector<std::string> id;
  int n() const { return int(kind.size()); }
  template <class Codes>
  void add(const Codes& codes, const V& w, int kd, int g, const std::string& name) {
    for (int c : codes) seq.push_back(uint8_t(c));
    ws.insert(ws.end(), w.begin(), w.end());
    off.push_back(int64_t(seq.size()));
    kind.push_back(uint8_t(kd)); gate.push_back(g); id.push_back(name);
  }
};

class RNAelemTrainer {
  unsigned mode_;
  DeviceGroup& dev_;
  OutputSet& out_;
  FastqBatchReader qr_;
  Adam adam_;
  MotifModel* motif_ = nullptr;
  int max_iter_ = 30, kmer_shuf_ = 2, cnt_ = 0;
  double lambda_init_ = 1., sum_eff_ = 0.;
  std::chrono::system_clock::time_point t0_;
 public:
  RNAelemTrainer(unsigned mode, DeviceGroup& dev, OutputSet& out) : mode_(mode), dev_(dev), out_(out) {}
  void set_fq_name(const std::string& f) { qr_.open(f); }
  void set_conditions(int max_iter, double /*epsilon: L-BFGS-B only*/, double lambda_init, int kmer_shuf, int batch_size) {
    max_iter_ = max_iter; kmer_shuf_ = kmer_shuf; lambda_init_ = lambda_init;
    adam_.set_hp(0, 0, 0.1, 0.9, 0.999, 1.e-8);
    qr_.set_batch_size(batch_size);
    cry("batch size:", batch_size);
  }
  int evaluations() const { return cnt_; }

  // One full-batch objective evaluation of the model as it is: `fn:` on channel 1, `gr:` on channel 2, 17 digits
  // (RNAelemTrainer::eval, motif_eval.hpp:22-53).  In the reference this sub-command never sets a batch size and
  // evaluates zero reads; here it does what it was written for, at evaluation count 0, with
  // shuffled negatives unless --no-shuffle.
  void eval(MotifModel& model) {
    motif_ = &model;
    V x, gr;
    model.pack(x);
    dev_.set_model(model);
    qr_.set_batch_size(-1);
    cnt_ = 0;
    double fn = 0.;
    (*this)(x, fn, gr);
    for (int id : {1, 2}) out_.at(id).precision(17);
    out_.dat(1, "fn:", fn);
    out_.dat(2, "gr:", gr);
    for (int id : {1, 2}) out_.at(id).precision(6);
  }

  void train(MotifModel& model) {
    check(!(mode_ & TR_NO_SHUFFLE), "--no-shuffle training needs the L-BFGS-B optimizer, which is not available in this build");
    motif_ = &model;
    model.lambda[0] = model.lambda[1] = lambda_init_;
    V x;
    model.pack(x);
    const int nth = model.n_theta();
    V lo(nth, -kInf), hi(nth + 2, kInf);               // set_bounds: theta >= log 0, lambda >= 0
    lo.push_back(0.); lo.push_back(0.);
    adam_.set_bounds(lo, hi, VI(nth + 2, 1));
    V rho(nth, model.theta_softmax ? model.rho_s : model.rho_theta);   // set_regularization: L2 everywhere
    rho.push_back(model.rho_lambda); rho.push_back(model.rho_lambda);
    adam_.set_regularization(VI(nth + 2, 2), rho);
    dev_.set_model(model);
    t0_ = std::chrono::system_clock::now();
    cnt_ = 0;
    adam_.minimize(*this, x, max_iter_);
    model.unpack(adam_.x());
    std::chrono::duration<double> dt = std::chrono::system_clock::now() - t0_;
    cry("wall clock time per eval:", dt.count() / cnt_);
  }

  // the objective: fn = sum over the minibatch of ln Zo - ln Zx, gr = expected-count differences
  int operator()(const V& x, double& fn, V& gr) {
    const bool prof = std::getenv("RELEM_HOST_PROFILE") != nullptr;   // stderr: host-side time per stage of this call
    auto tp0 = std::chrono::steady_clock::now();
    auto lapse = [&tp0] {
      auto t = std::chrono::steady_clock::now();
      double ms = std::chrono::duration<double, std::milli>(t - tp0).count();
      tp0 = t;
      return ms;
    };
    if (qr_.size() - qr_.consumed_in_epoch() < qr_.batch_size()) qr_.skip(qr_.size() - qr_.consumed_in_epoch());
    motif_->unpack(x);
    if (qr_.epoch_done()) write_interim(out_, 3, *motif_);
    fn = 0.;
    gr.assign(x.size(), 0.);
    sum_eff_ = 0.;
    qr_.next_batch();

    // reads of the minibatch in reader order; weights and negatives are per-read work done on several host threads
    std::vector<const Read*> rs;
    while (!qr_.batch_done()) {
      const Read& r = qr_.get();
      check(r.seq.size() + 1 == r.qual.size(), "bad seq format.", r.id, r.seq.size(), r.qual.size());
      rs.push_back(&r);
    }
    const int nr = int(rs.size());
    std::vector<V> wv(nr);
    std::vector<VI> negv(nr);
    std::vector<char> flagged(nr, 0);
    std::vector<std::string> fail(nr);
    const int kshuf = kmer_shuf_, iter = cnt_;
    const bool shuffle = !(mode_ & TR_NO_SHUFFLE);   // without negatives a read is a unit of its own
    parallel_for(nr, [&, kshuf, iter, shuffle](int i) {
      try {
        flagged[i] = quality_to_weights(rs[i]->qual, wv[i]) ? 1 : 0;
        if (!shuffle) return;
        std::string neg = shuffled_negative(codes_to_text(rs[i]->seq), kshuf, iter);
        negv[i].resize(neg.size());
        for (size_t k = 0; k < neg.size(); ++k) negv[i][k] = base_code(neg[k]);
      } catch (std::exception& e) { fail[i] = e.what(); }
    });
    for (auto& f : fail) if (!f.empty()) throw std::runtime_error(f);
    const bool lr = (mode_ & TR_LIK_RATIO) != 0;   // likelihood-ratio objective: motif_trainer.hpp:156-202
    const double ms_pack = lapse();
    dev_.push_params(*motif_);
    const double ms_push = lapse();

    // Every rank packs and evaluates its contiguous block of reads (a read travels with the negative shuffled from
    // it, so a gate never crosses ranks), then the partial sums meet in one all-reduce.
    const int nth = motif_->n_theta(), nw = dev_.size();
    V en(nth, 0.);
    double eh[2] = {0., 0.};
    std::vector<V> part(nw);
    std::vector<std::vector<std::string>> skipped_ids(nw);
    dev_.each([&](int k) {
      long p0, p1;
      shard_range(nr, nw, k, p0, p1);
      PackedBatch b;
      for (long i = p0; i < p1; ++i) {
        int me = b.n();
        b.add(rs[i]->seq, wv[i], flagged[i] ? RELEM_POS_WITH : lr ? RELEM_LR_WITHOUT : RELEM_POS_WITHOUT, -1, rs[i]->id);
        if (shuffle) b.add(negv[i], V(negv[i].size(), 0.), lr ? RELEM_LR_NEG : RELEM_NEG, me, rs[i]->id);   // qualities all 0 -> weight 0
      }
      const int ns = b.n();
      std::vector<uint8_t> skipped(ns, 0);
      V& v = part[k];
      v.assign(nth + 6, 0.);   // fn, sum_eff, n_skipped, EN_diff[nth], EH_diff[2], number of ranks whose E-step failed
      relem_estep_out o;
      std::memset(&o, 0, sizeof o);
      o.EN_diff = v.data() + 3;
      o.skipped = skipped.data();
      // A rank whose E-step fails must still enter the collective (the other ranks are already waiting in it): the
      // failure travels as one more summed element and every rank stops after the all-reduce.
      int rc = RELEM_OK;
      std::string why;
      if (ns > 0) rc = relem_estep(dev_.ctx(k), ns, b.seq.data(), b.off.data(), b.ws.data(), b.kind.data(), b.gate.data(), &o);
      if (const char* e = std::getenv("RELEM_TEST_FAIL_RANK"))   // fault injection for tests/test_host_cli.py
        if (std::atoi(e) == k && rc == RELEM_OK) { rc = RELEM_EINVAL; why = "injected failure (RELEM_TEST_FAIL_RANK)"; }
      if (rc != RELEM_OK && why.empty()) why = relem_last_error(dev_.ctx(k));
      if (rc != RELEM_OK) { v.assign(nth + 6, 0.); v[5 + nth] = 1.; }
      else {
        v[0] = o.fn; v[1] = o.sum_eff; v[2] = double(o.n_skipped);
        v[3 + nth] = o.EH_diff[0]; v[4 + nth] = o.EH_diff[1];
        for (int n = 0; n < ns; ++n) if (skipped[n] == 1) skipped_ids[k].push_back(b.id[n]);
      }
      if (nw > 1) dev_.ok(k, relem_allreduce_sum(dev_.ctx(k), v.data(), int(v.size())), "relem_allreduce_sum");
      if (rc != RELEM_OK) die("relem: relem_estep failed on GPU", k, ":", why);
    });
    check(part[0][5 + nth] == 0., "relem: the E-step failed on", int(part[0][5 + nth]), "GPU(s)");
    const double ms_estep = lapse();
    if (prof) {
      const char* nm[8]; float ms[8]; int nl[8];
      int ne = relem_last_timing(dev_.ctx(0), nm, ms, nl, 8);
      double kms = 0.;
      for (int k = 0; k < ne; ++k) kms += ms[k];
      cry("host profile: pack", ms_pack, "ms, set_params", ms_push, "ms, relem_estep", ms_estep, "ms of which kernels", kms, "ms");
    }
    const V& tot = part[0];
    fn = tot[0]; sum_eff_ = tot[1];
    for (int k = 0; k < nth; ++k) en[k] = tot[3 + k];
    eh[0] = tot[3 + nth]; eh[1] = tot[4 + nth];
    if (cnt_ == 0)
      for (const auto& ids : skipped_ids) for (const std::string& id : ids) cry("skipped:", id);

    // RNAelemTrainDP's update block (motif_trainer.hpp:248-271)
    int k = 0;
    if (motif_->theta_softmax) {
      for (const V& row : motif_->theta) {
        double tsum = 0.;
        for (size_t j = 0; j < row.size(); ++j) tsum += en[k + j];
        for (size_t j = 0; j < row.size(); ++j) {
          double d = en[k + j], p = std::exp(row[j]);
          gr[k + j] += (1 - p) * d - p * (tsum - d);
        }
        k += int(row.size());
      }
    } else {
      for (; k < nth; ++k) gr[k] += en[k];
    }
    // the reference files an energy count under lambda[0] whenever the transition's lambda VALUE equals lambda[0]
    // (motif_trainer.hpp:380-381): with equal lambdas everything lands in the first slot
    if (motif_->lambda[0] == motif_->lambda[1]) { gr[k] += eh[0] + eh[1]; }
    else { gr[k] += eh[0]; gr[k + 1] += eh[1]; }

    if (adam_.itercount() == 0) cry("considered BP:", sum_eff_ / qr_.in_batch());
    ++cnt_;
    double gg = 0.;
    for (double g : gr) gg += g * g;
    cry("iter:", adam_.itercount(), ", y:", fn, ", |gr|:", gg, ", p|x|:", adam_.rgl_term(x));
    return 0;
  }
};

class RNAelemScanner {
  DeviceGroup& dev_;
  OutputSet& out_;
  FastqReader qr_;
  int out_id_ = 1;
  long chunk_;   // reads per relem_scan call and GPU
 public:
  // 16 384 reads per call keep the device chunks large (2 048 reads per call reach 4.3 k reads/s, 8 192 and more 5.4 k,
  // tools/scan_probe.py) and let 8 GPUs take 100 000 reads in one round; RELEM_SCAN_CHUNK overrides
  RNAelemScanner(DeviceGroup& dev, OutputSet& out, long chunk = 16384) : dev_(dev), out_(out), chunk_(chunk) {
    if (const char* e = std::getenv("RELEM_SCAN_CHUNK")) chunk_ = std::max(1L, std::atol(e));
  }
  void set_fq_name(const std::string& f) { qr_.open(f); }
  void set_out_id(int id) { out_id_ = id; }

  void scan(MotifModel& model) {
    auto t0 = std::chrono::system_clock::now();
    dev_.set_model(model);
    const std::string node = dev_.node_chars();
    const int M = int(node.size()), nth = model.n_theta(), nw = dev_.size();
    V en_total(nth, 0.);
    qr_.rewind();
    std::vector<const Read*> reads;
    while (!qr_.at_end()) reads.push_back(&qr_.get());

    struct Result {
      PackedBatch b;
      V ps, pe, pi, exist;
      std::vector<int32_t> psihat, ys, ye;
      std::string rss;
      V en;
    };
    // the ten lines of reads [n0, n1) of one rank's block (motif_scanner.hpp:240-251)
    auto format_records = [&node, M](const Result& r, int n0, int n1) {
      std::ostringstream os;
      for (int n = n0; n < n1; ++n) {
        size_t a = size_t(r.b.off[n]), L = size_t(r.b.off[n + 1]) - a;
        VI psi(r.psihat.begin() + a, r.psihat.begin() + a + L), codes(r.b.seq.begin() + a, r.b.seq.begin() + a + L);
        std::string mot;
        for (int h : psi) mot += (h == 0 || h == M - 1) ? ' ' : node[h];
        put_line(os, "id:", r.b.id[n]);
        put_line(os, "start:", V(r.ps.begin() + a, r.ps.begin() + a + L));
        put_line(os, "end:", V(r.pe.begin() + a + n, r.pe.begin() + a + n + L + 1));
        put_line(os, "inner:", V(r.pi.begin() + a, r.pi.begin() + a + L));
        put_line(os, "psihat:", psi);
        put_line(os, "motif region:", r.ys[n], "-", r.ye[n]);
        put_line(os, "exist prob:", r.exist[n]);
        put_line(os, "seq:", codes_to_text(codes));
        put_line(os, "rss:", r.rss.substr(a, L));
        put_line(os, "mot:", mot);
      }
      return os.str();
    };
    // Rounds of nw * chunk reads: rank k scans its contiguous block.  Turning ~600 doubles per read into text costs
    // about as much host time as the GPU needs for the DP, so the records of a round are formatted by a few host
    // threads while the GPUs already work on the next round, and written in input order one round behind.
    const int n_fmt = int(std::max(1u, std::min(16u, std::thread::hardware_concurrency())));
    std::vector<std::future<std::string>> pending;
    auto drain = [&] {
      std::ostream& os = out_.at(out_id_);
      for (auto& f : pending) os << f.get();
      os.flush();
      pending.clear();
    };
    for (size_t base = 0; base < reads.size(); base += size_t(nw) * chunk_) {
      long n_round = long(std::min(reads.size() - base, size_t(nw) * chunk_));
      auto res = std::make_shared<std::vector<Result>>(nw);
      dev_.each([&](int k) {
        long a, z;
        shard_range(n_round, nw, k, a, z);
        Result& r = (*res)[k];
        V w;
        for (long q = a; q < z; ++q) {
          const Read& rd = *reads[base + q];
          check(rd.seq.size() + 1 == rd.qual.size(), "bad seq format.", rd.id, rd.seq.size(), rd.qual.size());
          quality_to_weights(rd.qual, w);
          r.b.add(rd.seq, w, RELEM_POS_WITHOUT, -1, rd.id);
        }
        int ns = r.b.n();
        if (!ns) return;
        size_t tl = r.b.seq.size();
        r.ps.assign(tl, 0.); r.pe.assign(tl + ns, 0.); r.pi.assign(tl, 0.); r.exist.assign(ns, 0.);
        r.psihat.assign(tl, 0); r.ys.assign(ns, 0); r.ye.assign(ns, 0); r.rss.assign(tl + 1, ' '); r.en.assign(nth, 0.);
        relem_scan_out o;
        std::memset(&o, 0, sizeof o);
        o.PysL = r.ps.data(); o.PyeL = r.pe.data(); o.PyiL = r.pi.data(); o.psihat = r.psihat.data();
        o.rss = &r.rss[0]; o.Ys = r.ys.data(); o.Ye = r.ye.data(); o.exist_prob = r.exist.data(); o.EN = r.en.data();
        dev_.ok(k, relem_scan(dev_.ctx(k), ns, r.b.seq.data(), r.b.off.data(), r.b.ws.data(), &o), "relem_scan");
      });
      drain();   // the previous round's text was produced while this round ran on the GPUs
      for (int k = 0; k < nw; ++k) {
        const int ns = (*res)[k].b.n(), per = (ns + n_fmt - 1) / std::max(1, n_fmt);
        for (int n0 = 0; n0 < ns; n0 += per)
          pending.push_back(std::async(std::launch::async, [res, k, n0, per, ns, format_records] {
            return format_records((*res)[k], n0, std::min(ns, n0 + per));
          }));
        for (int t = 0; t < nth && !(*res)[k].en.empty(); ++t) en_total[t] += (*res)[k].en[t];
      }
    }
    drain();
    VV en_rows;
    size_t k = 0;
    for (const V& row : model.theta) { en_rows.push_back(V(en_total.begin() + k, en_total.begin() + k + row.size())); k += row.size(); }
    cry("E[N]:", en_rows);
    std::chrono::duration<double> dt = std::chrono::system_clock::now() - t0;
    cry("scan end:", dt.count());
  }
};

}  // namespace relem
#endif
