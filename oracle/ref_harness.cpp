// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, not product code.
//
// A thin driver around the UNMODIFIED reference headers (included by path from
// /root/reference/RNAelem at build time; nothing is copied into this repo).
// It runs the reference's own classes and prints, at 17 significant digits,
// the quantities the parity tests compare against:
//
//   tables <T2004|A2007>              every EnergyParam table after the reference's parse
//   hmm    <pattern>                  the ProfileHMM automaton (states, transition lists, quads)
//   bpp    <model> <fq>               per sequence: canonical pairs, lnBPP, bp_ok mask, bpp_eff
//   estep  <model> <fq> <shuf> <iter> per sequence Z's and expected counts, then fn / gr of
//                                     RNAelemTrainer::operator() (motif_trainer.hpp:595-633)
//   scan   <model> <fq>               RNAelemScanner::scan records (motif_scanner.hpp:938-949) at 17 digits
//   dump   <model> <fq> <idx> <out>   inside/outside tables of one sequence (binary doubles)
//
// Built by oracle/Makefile into oracle/_ref/ (git-ignored, shipped to the GPU box).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <chrono>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <istream>
#include <limits>
#include <memory>
#include <mutex>
#include <numeric>
#include <ostream>
#include <queue>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <sys/types.h>
#include <thread>
#include <unistd.h>
#include <unordered_map>
#include <vector>

// built with g++ -fno-access-control so the dumps can read the reference's private tables
#include "const_options.hpp"
#include "util.hpp"
#include "profile_hmm.hpp"
#include "motif_model.hpp"
#include "motif_trainer.hpp"
#include "motif_scanner.hpp"
#include "motif_io.hpp"

using namespace iyak;

static void pv(const char* name, const double* p, size_t n) {
  printf("%s %zu", name, n);
  for (size_t i = 0; i < n; ++i) printf(" %.17g", p[i]);
  printf("\n");
}
static void pvv(const char* name, VV const& v) {
  size_t n = 0;
  for (auto& r : v) n += r.size();
  printf("%s %zu", name, n);
  for (auto& r : v) for (double x : r) printf(" %.17g", x);
  printf("\n");
}

static int cmd_tables(const char* which) {
  EnergyParam ep;
  ep.use_default(!strcmp(which, "A2007") ? EnergyParam::A2007 : EnergyParam::T2004);
  pv("hairpin", ep._hairpin, 31);
  pv("mismatch_h", &ep._mismatch_h[0][0][0], 7 * 5 * 5);
  pv("mismatch_i", &ep._mismatch_i[0][0][0], 7 * 5 * 5);
  pv("mismatch_m", &ep._mismatch_m[0][0][0], 7 * 5 * 5);
  pv("mismatch_1ni", &ep._mismatch_1ni[0][0][0], 7 * 5 * 5);
  pv("mismatch_23i", &ep._mismatch_23i[0][0][0], 7 * 5 * 5);
  pv("mismatch_ext", &ep._mismatch_ext[0][0][0], 7 * 5 * 5);
  pv("triloop", ep._triloop, 40);
  pv("tetraloop", ep._tetraloop, 40);
  pv("hexaloop", ep._hexaloop, 40);
  pv("stack", &ep._stack[0][0], 49);
  pv("bulge", ep._bulge, 31);
  pv("term_au", &ep._term_au, 1);
  pv("int11", &ep._int_11[0][0][0][0], 8 * 8 * 5 * 5);
  pv("int21", &ep._int_21[0][0][0][0][0], 8 * 8 * 5 * 5 * 5);
  pv("int22", &ep._int_22[0][0][0][0][0][0], 8 * 8 * 5 * 5 * 5 * 5);
  pv("internal", ep._internal, 31);
  pv("dangle5", &ep._dangle5[0][0], 40);
  pv("dangle3", &ep._dangle3[0][0], 40);
  pv("ninio", ep._ninio, 31);
  pv("mlintern", &ep._mlintern, 1);
  pv("mlclosing", &ep._mlclosing, 1);
  pv("ml_base", &ep._ml_base, 1);
  pv("lxc37", &ep._lxc37, 1);
  printf("triloops \"%s\"\n", ep._triloops.c_str());
  printf("tetraloops \"%s\"\n", ep._tetraloops.c_str());
  printf("hexaloops \"%s\"\n", ep._hexaloops.c_str());
  return 0;
}

static int cmd_hmm(const char* pattern) {
  ProfileHMM mm;
  mm.build(pattern);
  int M = (int)mm.size();
  printf("M %d\n", M);
  printf("nodes");
  for (int h = 0; h < M; ++h) printf(" %c", (char)mm.node(h));
  printf("\ntheta_id");
  for (int h = 0; h < M; ++h) printf(" %d", mm.theta_id(h));
  printf("\ntheta_rows");
  for (auto& r : mm.theta()) printf(" %zu", r.size());
  printf("\nS %zu\n", mm.state().size());
  printf("states");
  for (auto& s : mm.state()) printf(" %d:%d:%d", s.id, s.l, s.r);
  printf("\nloop_states");
  for (auto& s : mm.loop_state()) printf(" %d", s.id);
  printf("\n");
  for (auto& s : mm.state()) {
    printf("right %d", s.id);
    for (auto& t : mm.loop_right_trans(s.id)) printf(" %d", t.id);
    printf("\nleft %d", s.id);
    for (auto& t : mm.loop_left_trans(s.id)) printf(" %d", t.id);
    printf("\npair %d", s.id);
    for (auto& t : mm.pair_trans(s.id)) printf(" %d", t.id);
    printf("\n");
  }
  printf("quads %zu", mm.loop_loop_states().size());
  for (auto& q : mm.loop_loop_states()) printf(" %d,%d,%d,%d", q[0].id, q[1].id, q[2].id, q[3].id);
  printf("\nreachable");
  for (int a = 0; a < M; ++a) for (int b = 0; b < M; ++b) printf(" %d", (int)mm.reachable(a, b));
  printf("\n");
  return 0;
}

static void read_model(RNAelem& model, const char* fname) {
  RNAelemReader reader;
  reader.set_model_fname(fname);
  reader.read_model(model);
}

static int cmd_bpp(const char* model_fname, const char* fq) {
  RNAelem model;
  read_model(model, model_fname);
  FastqReader qr;
  qr.set_fq_fname(fq);
  while (!qr.is_end()) {
    string id, rss; VI seq, qual;
    qr.get_read(id, seq, qual, rss);
    EnergyModel& em = model.em;
    // pass 1: unfiltered (what fill_bpp_tables sees before filtering): lnBPP of every canonical pair
    double keep = em._min_BPP;
    em.set_min_BPP(keep);
    em.set_seq(seq);
    int L = em.L, W = em.W;
    printf("seq %s L %d W %d C %d bpp_eff %.17g\n", id.c_str(), L, W, em.C, em.bpp_eff());
    printf("bp_ok");
    for (int i = 0; i <= L; ++i) for (int d = 0; d <= W; ++d) if (em._bp_ok[i][d]) printf(" %d,%d", i, d);
    printf("\nleft_ok");
    for (int i = 0; i <= L; ++i) for (int d = 0; d <= W; ++d) if (em._left_bp_ok[i][d]) printf(" %d,%d", i, d);
    printf("\n");
    if (keep > 0) {
      // the energy-only tables still hold the unfiltered inside/outside: print lnBPP for canonical pairs
      printf("lnbpp");
      for (int i = 0; i <= L; ++i)
        for (int j = i + 5; j <= std::min(L, i + W); ++j)
          if (0 < bp[seq[i]][seq[j - 1]]) printf(" %d,%d,%.17g", i, j - i, em.lnBPP(i, j));
      printf("\nlnZ %.17g\n", em.inside_o(L));
    }
  }
  return 0;
}

// per-sequence E-step driver: mirrors RNAelemTrainDP::operator() (motif_trainer.hpp:124-272,
// non-lik-ratio branch) one sequence at a time so that the intermediate values can be printed.
struct OneSeqDP : public RNAelemTrainDP {
  using RNAelemTrainDP::RNAelemTrainDP;
  void run(const char* tag, string const& id, VI& seq, VI const& qual, bool neg) {
    VV ENo, ENx; V EHo{0., 0.}, EHx{0., 0.};
    _m.mm.clear_emit_count(ENo);
    _m.mm.clear_emit_count(ENx);
    _m.set_seq(seq);
    _m.set_ws(qual);
    init_inside_tables();
    init_outside_tables(true, true);
    _m.compute_inside(InsideFun(this, ws()));
    double Ztt = part_func(true, true), Ztf = part_func(true, false), Zft = part_func(false, true);
    printf("%s %s L %d bpp_eff %.17g Ztt %.17g Ztf %.17g Zft %.17g\n", tag, id.c_str(), _m.L,
           _m.no_rss() ? 0. : _m.em.bpp_eff(), Ztt, Ztf, Zft);
    bool ok = neg ? std::isfinite(Ztt) : (std::isfinite(Ztt) && std::isfinite(Ztf));
    if (!ok) { printf("skipped 1\n"); return; }
    printf("skipped 0\n");
    _m.compute_outside(OutsideFun(this, ws(), Ztt, EHo, ENo));
    double Zx;
    bool restricted_ari = !neg && !(-inf < ws().back());   // positives asserted to contain the motif
    if (restricted_ari) { init_outside_tables(true, false); Zx = Ztf; }
    else { init_outside_tables(false, true); Zx = Zft; }
    _m.compute_outside(OutsideFun(this, ws(), Zx, EHx, ENx));
    printf("Zo %.17g Zx %.17g\n", Ztt, Zx);
    pvv("ENo", ENo); pvv("ENx", ENx);
    pv("EHo", EHo.data(), 2); pv("EHx", EHx.data(), 2);
  }
};

static int cmd_estep(const char* model_fname, const char* fq, int shuf, int iter, int kmer) {
  RNAelem model;
  read_model(model, model_fname);
  unsigned mode = TR_NORMAL | (shuf ? 0 : TR_NO_SHUFFLE);
  {
    RNAelemTrainer t(mode, 1);
    t.set_fq_name(fq);
    OneSeqDP dp(model, 0, 0, t._sum_eff, t._mx_input, t._mx_update, t._qr, mode, iter, kmer);
    FastqReader qr;
    qr.set_fq_fname(fq);
    while (!qr.is_end()) {
      string id, rss; VI seq, qual, neg;
      qr.get_read(id, seq, qual, rss);
      dp.run("pos", id, seq, qual, false);
      if (shuf) {
        string s; seq_int2str(seq, s);
        srand((int)count(s.begin(), s.end(), s[0]) + iter);
        ushuffle::set_randfunc(long_rand);
        static char neg_s[MAX_SEQLEN + 1];
        memset(neg_s, 0, sizeof(neg_s));
        ushuffle::shuffle(s.c_str(), neg_s, size(s), kmer);
        seq_str2int(neg_s, neg);
        printf("negseq %s\n", neg_s);
        VI q0(qual.size(), 0);
        dp.run("neg", id, neg, q0, true);
      }
    }
  }
  // the real objective functor, full batch, iteration counter = iter
  RNAelemTrainer t(mode, 1);
  t.set_fq_name(fq);
  t.set_conditions(1, 1e-5, 0, kmer, -1, "~NULL~");
  t._cnt = iter;
  t._motif = &model;
  V x, gr; double fn = 0;
  model.pack_params(x);
  t(x, fn, gr);
  printf("fn %.17g\n", fn);
  pv("gr", gr.data(), gr.size());
  printf("sum_eff %.17g\n", t._sum_eff);
  return 0;
}

static int cmd_scan(const char* model_fname, const char* fq) {
  RNAelem model;
  read_model(model, model_fname);
  std::cout.precision(17);
  std::cerr.precision(17);
  set_ostream(1, "~COUT~");
  RNAelemScanner scan(1);
  scan.set_out_id(1);
  scan.set_fq_name(fq);
  scan.scan(model);
  return 0;
}

static int cmd_dump(const char* model_fname, const char* fq, int idx, const char* out) {
  RNAelem model;
  read_model(model, model_fname);
  unsigned mode = TR_NORMAL | TR_NO_SHUFFLE;
  RNAelemTrainer t(mode, 1);
  t.set_fq_name(fq);
  RNAelemTrainDP dp(model, 0, 0, t._sum_eff, t._mx_input, t._mx_update, t._qr, mode, 0, 2);
  FastqReader qr;
  qr.set_fq_fname(fq);
  string id, rss; VI seq, qual;
  for (int n = 0; n <= idx; ++n) qr.get_read(id, seq, qual, rss);
  dp._m.set_seq(seq);
  dp._m.set_ws(qual);
  dp.init_inside_tables();
  dp.init_outside_tables(true, true);
  dp._m.compute_inside(RNAelemTrainDP::InsideFun(&dp, dp.ws()));
  VV EN; V EH{0., 0.};
  dp._m.mm.clear_emit_count(EN);
  double Z = dp.part_func(true, true);
  dp._m.compute_outside(RNAelemTrainDP::OutsideFun(&dp, dp.ws(), Z, EH, EN));
  FILE* f = fopen(out, "wb");
  int hdr[4] = {dp._m.L, dp._m.W, dp._m.E - 1, dp._m.S};
  fwrite(hdr, sizeof(int), 4, f);
  for (int pass = 0; pass < 2; ++pass) {
    auto& T = pass ? dp._outside : dp._inside;
    for (auto& a : T) for (auto& b : a) for (auto& c : b) fwrite(c.data(), sizeof(double), c.size(), f);
    auto& O = pass ? dp._outside_o : dp._inside_o;
    for (auto& a : O) fwrite(a.data(), sizeof(double), a.size(), f);
  }
  fclose(f);
  return 0;
}

int main(int argc, char** argv) {
  init_ostream(4);
  try {
    if (argc >= 3 && !strcmp(argv[1], "tables")) return cmd_tables(argv[2]);
    if (argc >= 3 && !strcmp(argv[1], "hmm")) return cmd_hmm(argv[2]);
    if (argc >= 4 && !strcmp(argv[1], "bpp")) return cmd_bpp(argv[2], argv[3]);
    if (argc >= 6 && !strcmp(argv[1], "estep"))
      return cmd_estep(argv[2], argv[3], atoi(argv[4]), atoi(argv[5]), argc >= 7 ? atoi(argv[6]) : 2);
    if (argc >= 4 && !strcmp(argv[1], "scan")) return cmd_scan(argv[2], argv[3]);
    if (argc >= 6 && !strcmp(argv[1], "dump")) return cmd_dump(argv[2], argv[3], atoi(argv[4]), argv[5]);
  } catch (std::exception& e) {
    fprintf(stderr, "ref_harness: %s\n", e.what());
    return 1;
  }
  fprintf(stderr, "usage: ref_harness tables|hmm|bpp|estep|scan|dump ...\n");
  return 2;
}
