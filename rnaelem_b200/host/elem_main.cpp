// elem_main.cpp -- the `RNAelem` command line (train / scan / train+write+scan) with the DP on B200.
//
// Same sub-commands, options, defaults, output channels and file formats as the reference binary
// (RNAelem/main.cpp:19-163, application.hpp:76-410), so `script/elem` can call this binary in its place.  Host work
// (reading, negative generation, Adam, writers) is elem_host.hpp / elem_driver.hpp; every DP pass is a librelem call.
// One addition: --gpus N (or RELEM_GPUS) shards each minibatch / scan over N GPUs of the box.  `-t/--thread` is
// accepted and ignored (the batch runs on the GPU).  `eval` prints the full-batch objective and gradient of a model
// with 17 digits.  Sub-commands that only exist for the reference's Grid Engine fan-out or drawing (array-eval, develop,
// logo) are refused.
#include <cctype>
#include <cstdio>
#include <cstdlib>

#include "elem_driver.hpp"

namespace {
using namespace relem;

struct Options {
  std::map<std::string, std::string> val;
  std::vector<std::string> args;
  std::string str(const std::string& k) const { return val.at(k); }
  int num(const std::string& k) const { return detail::parse<int>(val.at(k)); }
  double real(const std::string& k) const { return detail::parse<double>(val.at(k)); }
  bool flag(const std::string& k) const { return val.at(k) == "1"; }
};

struct Spec { const char* shrt; const char* lng; const char* dest; const char* dflt; bool is_flag; };
const Spec SPECS[] = {
    {"-f", "--fastq", "seq_fname", "~NONE~", false},
    {"-m", "--motif-pattern", "pattern", "~NONE~", false},
    {"-q", "--motif-model", "model_fname", "~NONE~", false},
    {nullptr, "--pict", "pic_fname", "~NONE~", false},
    {"-i", "--max-iter", "max_iter", "100", false},
    {nullptr, "--out1", "out1", "~COUT~", false},
    {nullptr, "--out2", "out2", "~COUT~", false},
    {nullptr, "--out3", "out3", "~COUT~", false},
    {nullptr, "--energy-param", "ene_param_fname", "~T2004~", false},
    {"-w", "--max-span", "max_span", "50", false},
    {"-c", "--max-internal-loop", "max_iloop", "30", false},
    {nullptr, "--epsilon", "eps", "1e-5", false},
    {nullptr, "--rho-s", "rho_s", "0.1", false},
    {nullptr, "--rho-theta", "rho_theta", "0.1", false},
    {nullptr, "--rho-lambda", "rho_lambda", "0.1", false},
    {nullptr, "--tau", "tau", "0.1", false},
    {nullptr, "--lambda-init", "lambda_init", "0", false},
    {nullptr, "--lambda-prior", "lambda_prior", "0", false},
    {"-p", "--min-bpp", "min_bpp", "1e-4", false},
    {nullptr, "--param-set", "param_set", "", false},
    {"-a", "--array", "array", "1", false},
    {nullptr, "--tmp", "tmp", "~NULL~", false},
    {nullptr, "--sge-option-file", "sge_opt_fname", "~DEFAULT~", false},
    {nullptr, "--font", "font", "~DEFAULT~", false},
    {nullptr, "--no-rss", "no_rss", "0", true},
    {nullptr, "--no-profile", "no_prf", "0", true},
    {nullptr, "--no-energy", "no_ene", "0", true},
    {"-t", "--thread", "thread", "1", false},
    {nullptr, "--no-shuffle", "no_shuffle", "0", true},
    {nullptr, "--theta-softmax", "theta_softmax", "0", true},
    {nullptr, "--kmer-shuf", "kmer_shuf", "2", false},
    {nullptr, "--lik-ratio", "lik_ratio", "0", true},
    {nullptr, "--batch-size", "batch_size", "100", false},
    {nullptr, "--gpus", "gpus", "", false},
};

Options parse_command_line(int argc, const char* const* argv) {
  Options o;
  for (const Spec& s : SPECS) o.val[s.dest] = s.dflt;
  for (int k = 1; k < argc; ++k) {
    std::string a = argv[k], inline_val;
    bool has_inline = false;
    if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
      size_t eq = a.find('=');
      if (eq != std::string::npos) { inline_val = a.substr(eq + 1); a = a.substr(0, eq); has_inline = true; }
    }
    if (a.size() < 2 || a[0] != '-' || (a[1] != '-' && std::isdigit((unsigned char)a[1]))) { o.args.push_back(a); continue; }
    const Spec* hit = nullptr;
    for (const Spec& s : SPECS)
      if (a == s.lng || (s.shrt && a == s.shrt)) hit = &s;
    check(hit != nullptr, "no such option:", a);
    if (hit->is_flag) { o.val[hit->dest] = "1"; continue; }
    if (has_inline) { o.val[hit->dest] = inline_val; continue; }
    check(k + 1 < argc, a, "option requires an argument");
    o.val[hit->dest] = argv[++k];
  }
  return o;
}

void model_from_options(const Options& o, MotifModel& m) {
  std::string pattern = o.str("pattern");
  bool no_rss = o.flag("no_rss");
  if (pattern.find('_') != std::string::npos) {   // application.hpp:402-407
    check(pattern.find('(') == std::string::npos && pattern.find(')') == std::string::npos,
          "patten cannot be mixture of _ & 'base pair'");
    no_rss = true;
    std::replace(pattern.begin(), pattern.end(), '_', '.');
  }
  m.theta_softmax = o.flag("theta_softmax");
  m.rho_s = o.real("rho_s"); m.rho_theta = o.real("rho_theta"); m.rho_lambda = o.real("rho_lambda");
  m.tau = o.real("tau"); m.lambda_prior = o.real("lambda_prior");
  m.ene_param = o.str("ene_param_fname"); m.max_span = o.num("max_span"); m.max_iloop = o.num("max_iloop");
  m.min_bpp = o.real("min_bpp"); m.no_ene = o.flag("no_ene");
  m.set_pattern(pattern, no_rss, o.flag("no_prf"));
}

int run(int argc, const char* const* argv) {
  Options o = parse_command_line(argc, argv);
  enum { NORMAL, TRAIN, SCAN, GENNEG, EVAL } mode = NORMAL;
  if (!o.args.empty()) {
    const std::string& c = o.args[0];
    if (c == "train") mode = TRAIN;
    else if (c == "scan") mode = SCAN;
    else if (c == "gen-neg") mode = GENNEG;
    else if (c == "eval") mode = EVAL;
    else if (c == "array-eval" || c == "develop" || c == "logo")
      die("sub-command not available in the B200 build:", c);
    else die("unknown sub-command:", o.args);
  }
  OutputSet out(4);
  out.bind(1, o.str("out1")); out.bind(2, o.str("out2")); out.bind(3, o.str("out3"));
  check(o.str("seq_fname") != "~NONE~", "require input filename (sequence)");
  if (mode == SCAN || mode == EVAL) check(o.str("model_fname") != "~NONE~", "require input filename (motif model)");
  check(o.num("array") <= 1, "Grid Engine array jobs are replaced by --gpus in the B200 build");
  check(o.str("param_set").empty(), "--param-set (masked training) is not available in the B200 build");

  if (mode == GENNEG) {   // main.cpp:131-152
    FastqReader qr;
    qr.open(o.str("seq_fname"));
    for (int it = 0; it < o.num("max_iter"); ++it) {
      qr.rewind();
      while (!qr.at_end()) {
        const Read& r = qr.get();
        std::string neg = shuffled_negative(codes_to_text(r.seq), o.num("kmer_shuf"), it);
        out.dat(1, ">iter:" + to_text(it) + ";seq:" + to_text(qr.consumed()) + ";orig:\"" + r.id + "\"");
        out.dat(1, neg);
      }
    }
    return 0;
  }

  int ngpu = 1;
  if (!o.str("gpus").empty()) ngpu = o.num("gpus");
  else if (const char* e = std::getenv("RELEM_GPUS")) ngpu = std::atoi(e);
  DeviceGroup dev;
  dev.open(ngpu);

  MotifModel model;
  if (o.str("model_fname") != "~NONE~") read_model(o.str("model_fname"), model);
  else if (mode != SCAN) model_from_options(o, model);

  if (mode == EVAL) {   // main.cpp:31-46
    unsigned tr = TR_NORMAL;
    if (o.flag("no_shuffle")) tr |= TR_NO_SHUFFLE;
    if (o.flag("lik_ratio")) tr |= TR_LIK_RATIO;
    RNAelemTrainer trainer(tr, dev, out);
    trainer.set_fq_name(o.str("seq_fname"));
    trainer.set_conditions(1, o.real("eps"), 0., o.num("kmer_shuf"), -1);
    trainer.eval(model);
    return 0;
  }
  if (mode == NORMAL || mode == TRAIN) {
    unsigned tr = TR_NORMAL;
    if (o.flag("no_shuffle")) tr |= TR_NO_SHUFFLE;
    if (o.flag("lik_ratio")) tr |= TR_LIK_RATIO;
    RNAelemTrainer trainer(tr, dev, out);
    trainer.set_fq_name(o.str("seq_fname"));
    trainer.set_conditions(o.num("max_iter"), o.real("eps"), o.real("lambda_init"), o.num("kmer_shuf"), o.num("batch_size"));
    trainer.train(model);
    // `RNAelem train` writes its model to channel 0, the null sink (main.cpp:118-119); only the sub-command-less
    // form puts it on --out1 and then scans the training reads into --out2 (main.cpp:77-84)
    write_model(out, mode == NORMAL ? 1 : 0, model);
  }
  if (mode == NORMAL || mode == SCAN) {
    RNAelemScanner scanner(dev, out);
    scanner.set_out_id(mode == NORMAL ? 2 : 1);
    scanner.set_fq_name(o.str("seq_fname"));
    scanner.scan(model);
  }
  return 0;
}
}  // namespace

int main(int argc, const char* argv[]) {
  try {
    return run(argc, argv);
  } catch (std::runtime_error& e) {   // main.cpp:154-161
    std::cerr << e.what() << std::endl;
    return 1;
  }
}
