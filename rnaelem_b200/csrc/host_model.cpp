// host_model.cpp -- see host_model.hpp.
#include "host_model.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <sstream>
#include <stdexcept>

#include "energy_data.inc"

namespace relem {

static const double NEG_INF = -std::numeric_limits<double>::infinity();

// ------------------------------------------------------------------------------------------- energy: ints
#define COPY_SET(NAME)                                                                         \
  do {                                                                                         \
    std::memcpy(out.stack, RELEM_##NAME##_stack, sizeof(out.stack));                           \
    std::memcpy(out.mm_h, RELEM_##NAME##_mm_h, sizeof(out.mm_h));                              \
    std::memcpy(out.mm_i, RELEM_##NAME##_mm_i, sizeof(out.mm_i));                              \
    std::memcpy(out.mm_1ni, RELEM_##NAME##_mm_1ni, sizeof(out.mm_1ni));                        \
    std::memcpy(out.mm_23i, RELEM_##NAME##_mm_23i, sizeof(out.mm_23i));                        \
    std::memcpy(out.mm_m, RELEM_##NAME##_mm_m, sizeof(out.mm_m));                              \
    std::memcpy(out.mm_ext, RELEM_##NAME##_mm_ext, sizeof(out.mm_ext));                        \
    std::memcpy(out.dangle5, RELEM_##NAME##_dangle5, sizeof(out.dangle5));                     \
    std::memcpy(out.dangle3, RELEM_##NAME##_dangle3, sizeof(out.dangle3));                     \
    std::memcpy(out.int11, RELEM_##NAME##_int11, sizeof(out.int11));                           \
    std::memcpy(out.int21, RELEM_##NAME##_int21, sizeof(out.int21));                           \
    std::memcpy(out.int22, RELEM_##NAME##_int22, sizeof(out.int22));                           \
    std::memcpy(out.hairpin, RELEM_##NAME##_hairpin, sizeof(out.hairpin));                     \
    std::memcpy(out.bulge, RELEM_##NAME##_bulge, sizeof(out.bulge));                           \
    std::memcpy(out.interior, RELEM_##NAME##_interior, sizeof(out.interior));                  \
    out.ninio_f = RELEM_##NAME##_scalars[0];                                                   \
    out.ninio_max = RELEM_##NAME##_scalars[1];                                                 \
    out.ml_base = RELEM_##NAME##_scalars[2];                                                   \
    out.ml_closing = RELEM_##NAME##_scalars[3];                                                \
    out.ml_intern = RELEM_##NAME##_scalars[4];                                                 \
    out.term_au = RELEM_##NAME##_scalars[5];                                                   \
    out.lxc37 = RELEM_##NAME##_lxc37;                                                          \
    out.tri.assign(RELEM_##NAME##_tri_seq, RELEM_##NAME##_tri_seq + RELEM_##NAME##_ntri);      \
    out.tri_e.assign(RELEM_##NAME##_tri_e, RELEM_##NAME##_tri_e + RELEM_##NAME##_ntri);        \
    out.tetra.assign(RELEM_##NAME##_tetra_seq, RELEM_##NAME##_tetra_seq + RELEM_##NAME##_ntetra); \
    out.tetra_e.assign(RELEM_##NAME##_tetra_e, RELEM_##NAME##_tetra_e + RELEM_##NAME##_ntetra);   \
    out.hexa.assign(RELEM_##NAME##_hexa_seq, RELEM_##NAME##_hexa_seq + RELEM_##NAME##_nhexa);  \
    out.hexa_e.assign(RELEM_##NAME##_hexa_e, RELEM_##NAME##_hexa_e + RELEM_##NAME##_nhexa);    \
  } while (0)

bool builtin_energy_ints(const std::string& name, EnergyInts& out) {
  static_assert(sizeof(out.stack) == sizeof(RELEM_T2004_stack), "stack");
  static_assert(sizeof(out.mm_m) == sizeof(RELEM_T2004_mm_m), "mm_m");
  static_assert(sizeof(out.int11) == sizeof(RELEM_T2004_int11), "int11");
  static_assert(sizeof(out.int21) == sizeof(RELEM_T2004_int21), "int21");
  static_assert(sizeof(out.int22) == sizeof(RELEM_T2004_int22), "int22");
  if (name == "~T2004~") { COPY_SET(T2004); return true; }
  if (name == "~A2007~") { COPY_SET(A2007); return true; }
  return false;
}

// --- text reader for ViennaRNA-2.0 format files, following the reading rules of energy_param.hpp:159-183
// (get_array: whitespace words; a line shorter than 2 chars ends the fill; a word containing "/*" ends the
// line; INF / DEF words) and :519-640 (which section fills which sub-block).
namespace {
struct LineStream {
  std::vector<std::string> lines;
  size_t pos = 0;
  bool getline(std::string& s) {
    if (pos >= lines.size()) return false;
    s = lines[pos++];
    return true;
  }
};
std::vector<std::string> words_of(const std::string& s) {
  std::vector<std::string> w;
  std::istringstream iss(s);
  for (std::string t; iss >> t;) w.push_back(t);
  return w;
}
void fill_ints(LineStream& st, int* dst, int size) {
  int n = 0;
  for (int k = 0; k < size; ++k) dst[k] = EnergyInts::INF;
  while (n < size) {
    std::string s;
    if (!st.getline(s) || s.size() < 2) break;
    std::vector<std::string> w = words_of(s);
    int prev = n;
    for (; n < size && n - prev < (int)w.size(); ++n) {
      const std::string& t = w[n - prev];
      if (t.find("/*") != std::string::npos) break;
      if (t == "INF") dst[n] = EnergyInts::INF;
      else if (t == "DEF") dst[n] = -50;
      else dst[n] = std::atoi(t.c_str());
    }
  }
}
std::string section_of(const std::string& s) {
  if (s.empty() || s[0] != '#') return "";
  std::vector<std::string> w = words_of(s);
  return w.size() > 1 ? w[1] : "";
}
// first data line of a scalar section: skips lines containing '*', stops at an empty line
bool scalar_line(LineStream& st, std::vector<std::string>& w) {
  std::string s;
  while (st.getline(s)) {
    if (s.empty()) return false;
    if (s.find('*') != std::string::npos) continue;
    w = words_of(s);
    return true;
  }
  return false;
}
}  // namespace

bool parse_param_text(const std::string& text, EnergyInts& o, std::string& err) {
  LineStream st;
  {
    std::istringstream iss(text);
    for (std::string s; std::getline(iss, s);) st.lines.push_back(s);
  }
  // defaults for everything a file may omit: forbidden
  auto fill_inf = [](int* p, size_t n) { for (size_t k = 0; k < n; ++k) p[k] = EnergyInts::INF; };
  fill_inf(&o.stack[0][0], 36); fill_inf(&o.mm_h[0][0], 150); fill_inf(&o.mm_i[0][0], 150);
  fill_inf(&o.mm_1ni[0][0], 150); fill_inf(&o.mm_23i[0][0], 150); fill_inf(&o.mm_m[0][0], 175);
  fill_inf(&o.mm_ext[0][0], 175); fill_inf(&o.dangle5[0][0], 35); fill_inf(&o.dangle3[0][0], 35);
  fill_inf(&o.int11[0][0][0], 1225); fill_inf(&o.int21[0][0][0], 6125); fill_inf(&o.int22[0][0][0], 9216);
  fill_inf(o.hairpin, 31); fill_inf(o.bulge, 31); fill_inf(o.interior, 31);
  o.ninio_f = o.ninio_max = o.ml_base = o.ml_closing = o.ml_intern = o.term_au = 0;
  o.lxc37 = 107.856;
  o.tri.clear(); o.tetra.clear(); o.hexa.clear(); o.tri_e.clear(); o.tetra_e.clear(); o.hexa_e.clear();
  // pass 1: lxc37 from the Misc section when its line has more than 4 words (energy_param.hpp:504-517,:475-476)
  {
    std::string s;
    while (st.getline(s)) {
      if (section_of(s) == "Misc") {
        while (st.getline(s)) {
          if (s.empty()) break;
          if (s.find('*') != std::string::npos) continue;
          std::vector<std::string> w = words_of(s);
          if (w.size() <= 2) { err = "read_misc"; return false; }
          if (w.size() > 4) o.lxc37 = std::atof(w[4].c_str());
        }
        break;
      }
    }
  }
  st.pos = 0;
  std::string s;
  while (st.getline(s)) {
    std::string t = section_of(s);
    if (t.empty()) continue;
    if (t == "stack") for (int a = 0; a < 6; ++a) fill_ints(st, o.stack[a], 6);
    else if (t == "mismatch_hairpin") for (int a = 0; a < 6; ++a) fill_ints(st, o.mm_h[a], 25);
    else if (t == "mismatch_interior") for (int a = 0; a < 6; ++a) fill_ints(st, o.mm_i[a], 25);
    else if (t == "mismatch_interior_1n") for (int a = 0; a < 6; ++a) fill_ints(st, o.mm_1ni[a], 25);
    else if (t == "mismatch_interior_23") for (int a = 0; a < 6; ++a) fill_ints(st, o.mm_23i[a], 25);
    else if (t == "mismatch_multi") for (int a = 0; a < 7; ++a) fill_ints(st, o.mm_m[a], 25);
    else if (t == "mismatch_exterior") for (int a = 0; a < 7; ++a) fill_ints(st, o.mm_ext[a], 25);
    else if (t == "dangle5") for (int a = 0; a < 7; ++a) fill_ints(st, o.dangle5[a], 5);
    else if (t == "dangle3") for (int a = 0; a < 7; ++a) fill_ints(st, o.dangle3[a], 5);
    else if (t == "int11") { for (int a = 0; a < 7; ++a) for (int b = 0; b < 7; ++b) fill_ints(st, o.int11[a][b], 25); }
    else if (t == "int21") { for (int a = 0; a < 7; ++a) for (int b = 0; b < 7; ++b) fill_ints(st, o.int21[a][b], 125); }
    else if (t == "int22") {
      for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b)
        for (int c = 0; c < 64; ++c) fill_ints(st, o.int22[a][b] + 4 * c, 4);
    }
    else if (t == "hairpin") fill_ints(st, o.hairpin, 31);
    else if (t == "bulge") fill_ints(st, o.bulge, 31);
    else if (t == "interior") fill_ints(st, o.interior, 31);
    else if (t == "NINIO") {
      std::vector<std::string> w;
      if (scalar_line(st, w)) {
        if (w.size() <= 2) { err = "read_ninio"; return false; }
        o.ninio_f = std::atoi(w[0].c_str()); o.ninio_max = std::atoi(w[2].c_str());
      }
    } else if (t == "ML_params") {
      std::vector<std::string> w;
      if (scalar_line(st, w)) {
        if (w.size() <= 4) { err = "read_ml"; return false; }
        o.ml_base = std::atoi(w[0].c_str()); o.ml_closing = std::atoi(w[2].c_str());
        o.ml_intern = std::atoi(w[4].c_str());
      }
    } else if (t == "Misc") {
      std::string u;
      while (st.getline(u)) {
        if (u.empty()) break;
        if (u.find('*') != std::string::npos) continue;
        std::vector<std::string> w = words_of(u);
        if (w.size() <= 2) { err = "read_misc"; return false; }
        o.term_au = std::atoi(w[2].c_str());
      }
    } else if (t == "Triloops" || t == "Tetraloops" || t == "Hexaloops") {
      std::vector<std::string>& names = t == "Triloops" ? o.tri : t == "Tetraloops" ? o.tetra : o.hexa;
      std::vector<int>& es = t == "Triloops" ? o.tri_e : t == "Tetraloops" ? o.tetra_e : o.hexa_e;
      names.clear(); es.clear();
      std::string u;
      while (st.getline(u)) {
        if (u.empty()) break;
        if (u.find('*') != std::string::npos) continue;
        std::vector<std::string> w = words_of(u);
        if (w.size() <= 1) { err = "read_string"; return false; }
        names.push_back(w[0]); es.push_back(std::atoi(w[1].c_str()));
      }
    }
  }
  return true;
}

// ----------------------------------------------------------------------------------------- energy: weights
namespace {
const double kGas = 1.98717, kK0 = 273.15;
const double kT = (37 + kK0) * kGas;
double smooth_int(int a) {  // Vienna "smooth" of energy_param.hpp:95-106
  double z = double(a);
  if (z / 10. < -1.2283697) return 0.;
  if (0.8660254 < z / 10.) return z;
  return 10. * 0.38490018 * (1. + std::sin(z / 10. - 0.34242663)) * (1. + std::sin(z / 10. - 0.34242663));
}
double lw(int z, bool smo = false) {  // log weight of an energy in dcal/mol (energy_param.hpp:108-114,:175-180)
  if (z == EnergyInts::INF) return NEG_INF;
  if (smo) return smooth_int(-z) * 10. / kT;
  return -z * 10. / kT;
}
template <class T> void fill_neg_inf(T& arr) {
  double* p = reinterpret_cast<double*>(&arr);
  for (size_t k = 0; k < sizeof(arr) / sizeof(double); ++k) p[k] = NEG_INF;
}
}  // namespace

void EnergyTables::build(const EnergyInts& e) {
  fill_neg_inf(hairpin); fill_neg_inf(mismatch_h); fill_neg_inf(mismatch_i); fill_neg_inf(mismatch_m);
  fill_neg_inf(mismatch_1ni); fill_neg_inf(mismatch_23i); fill_neg_inf(mismatch_ext); fill_neg_inf(stack);
  fill_neg_inf(bulge); fill_neg_inf(int11); fill_neg_inf(int21); fill_neg_inf(int22); fill_neg_inf(internal);
  fill_neg_inf(dangle5); fill_neg_inf(dangle3); fill_neg_inf(ninio);
  for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) stack[a + 1][b + 1] = lw(e.stack[a][b]);
  for (int a = 0; a < 6; ++a) for (int k = 0; k < 25; ++k) {
    mismatch_h[a + 1][k / 5][k % 5] = lw(e.mm_h[a][k]);
    mismatch_i[a + 1][k / 5][k % 5] = lw(e.mm_i[a][k]);
    mismatch_1ni[a + 1][k / 5][k % 5] = lw(e.mm_1ni[a][k]);
    mismatch_23i[a + 1][k / 5][k % 5] = lw(e.mm_23i[a][k]);
  }
  for (int a = 0; a < 7; ++a) for (int k = 0; k < 25; ++k) {
    mismatch_m[a + 1][k / 5][k % 5] = lw(e.mm_m[a][k], true);
    mismatch_ext[a + 1][k / 5][k % 5] = lw(e.mm_ext[a][k], true);
  }
  for (int a = 0; a < 7; ++a) for (int k = 0; k < 5; ++k) {
    dangle5[a + 1][k] = lw(e.dangle5[a][k], true);
    dangle3[a + 1][k] = lw(e.dangle3[a][k], true);
  }
  for (int a = 0; a < 7; ++a) for (int b = 0; b < 7; ++b) {
    for (int k = 0; k < 25; ++k) int11[a + 1][b + 1][k / 5][k % 5] = lw(e.int11[a][b][k]);
    for (int k = 0; k < 125; ++k) int21[a + 1][b + 1][k / 25][(k / 5) % 5][k % 5] = lw(e.int21[a][b][k]);
  }
  // The reference clears only the first 8*8*5*5*5 = 8 000 of the 40 000 entries of int22 before it reads the file
  // (energy_param.hpp:597-598); the entries with an unknown base ('N') beyond that are never written and read as
  // +0.0 -- log-weight 0, i.e. no penalty -- in its binaries (fresh zero pages; tests/golden/tables_*.npz holds the
  // dump).  Reads with N bases next to a 2x2 interior loop only agree with the reference if this is reproduced.
  {
    double* p22 = &int22[0][0][0][0][0][0];
    for (int k = 8000; k < 40000; ++k) p22[k] = 0.;
  }
  for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) for (int k = 0; k < 256; ++k)
    int22[a + 1][b + 1][1 + k / 64][1 + (k / 16) % 4][1 + (k / 4) % 4][1 + k % 4] = lw(e.int22[a][b][k]);
  for (int d = 0; d <= 30; ++d) {
    hairpin[d] = lw(e.hairpin[d]);
    bulge[d] = lw(e.bulge[d]);
    internal[d] = lw(e.interior[d]);
    ninio[d] = lw(std::min(e.ninio_max, d * e.ninio_f));
  }
  ml_base = lw(e.ml_base);
  mlclosing = lw(e.ml_closing);
  mlintern = lw(e.ml_intern);
  term_au = lw(e.term_au);
  lxc37 = e.lxc37;
  tri = e.tri; tetra = e.tetra; hexa = e.hexa;
  tri_w.clear(); tetra_w.clear(); hexa_w.clear();
  for (int z : e.tri_e) tri_w.push_back(lw(z));
  for (int z : e.tetra_e) tetra_w.push_back(lw(z));
  for (int z : e.hexa_e) hexa_w.push_back(lw(z));
}

double EnergyTables::hairpin_len(int d) const {
  if (d <= 30) return hairpin[d];
  const double maxloop_div = 1. / 30, kT_div = 1. / kT;
  return hairpin[30] - lxc37 * std::log(double(d) * maxloop_div) * 10. * kT_div;
}

// ------------------------------------------------------------------------------------------------ automaton
static void closure(std::vector<std::vector<char>>& m) {
  int n = (int)m.size();
  for (int k = 0; k < n; ++k)
    for (int i = 0; i < n; ++i)
      if (m[i][k])
        for (int j = 0; j < n; ++j)
          if (m[k][j]) m[i][j] = 1;
}

void ProfileHMM::build(const std::string& pat) {
  if (pat.empty()) throw std::runtime_error("empty motif");
  pattern = pat;
  // regularised pattern: runs of '*' collapse, leading/trailing '*' are dropped (profile_hmm.hpp:188-204)
  reg_pattern.clear();
  for (char c : pat)
    if (!(c == '*' && !reg_pattern.empty() && reg_pattern.back() == '*')) reg_pattern.push_back(c);
  size_t b = reg_pattern.find_first_not_of('*');
  reg_pattern.erase(0, b == std::string::npos ? reg_pattern.size() : b);
  size_t e = reg_pattern.find_last_not_of('*');
  if (e != std::string::npos) reg_pattern.erase(e + 1);

  node.assign(1, 'z');
  for (char c : reg_pattern) node.push_back(c);
  node.push_back('o');
  M = (int)node.size();

  pair.assign(M, -1);
  {
    std::vector<int> open;
    for (int h = 0; h < M; ++h) {
      if (node[h] == '(') open.push_back(h);
      else if (node[h] == ')') {
        if (open.empty()) throw std::runtime_error("unmatched brackets");
        int hl = open.back(); open.pop_back();
        pair[hl] = h; pair[h] = hl;
      }
    }
    if (!open.empty()) throw std::runtime_error("unmatched brackets");
  }
  // node graph: predecessor h-1, h-2 when h-1 is '*' (a '*' may be skipped), and a self loop
  edge_to.assign(M, std::vector<int>());
  edge_from.assign(M, std::vector<int>());
  for (int h = 0; h < M; ++h) {
    if (h > 0) {
      if (node[h - 1] == '*') { edge_to[h].push_back(h - 2); edge_from[h - 2].push_back(h); }
      edge_to[h].push_back(h - 1); edge_from[h - 1].push_back(h);
    }
    if (node[h] != '<' && node[h] != '>') { edge_to[h].push_back(h); edge_from[h].push_back(h); }
  }
  // emission rows
  theta_id.assign(M, -1);
  row_size.assign(1, 4);
  for (int h = 0; h < M; ++h) {
    switch (node[h]) {
      case ')': theta_id[h] = (int)row_size.size(); row_size.push_back(6); break;
      case '.': theta_id[h] = (int)row_size.size(); row_size.push_back(4); break;
      case '*': case 'z': case 'o': theta_id[h] = 0; break;
      case '<': case '>': case '(': break;
      default: throw std::runtime_error(std::string("bad motif char: ") + char(node[h]));
    }
  }
  // reachability between nodes (which intervals are states)
  reachable.assign(M, std::vector<char>(M, 0));
  reachable_as_loop.assign(M, std::vector<char>(M, 0));
  for (int h = 0; h < M; ++h) {
    int c = node[h];
    if (c == ')') { for (int h1 : edge_to[pair[h]]) reachable[h1][h] = 1; }
    else if (c == '>') { for (int h1 : edge_to[pair[h]]) { reachable[h1][h] = 1; reachable_as_loop[h1][h] = 1; } }
    else if (c == '(' || c == '<') {}
    else { for (int h1 : edge_to[h]) { reachable[h1][h] = 1; reachable_as_loop[h1][h] = 1; } }
    reachable[h][h] = 1; reachable_as_loop[h][h] = 1;
  }
  closure(reachable);
  closure(reachable_as_loop);
  // interval states, numbered by right end ascending then left end descending
  state.clear();
  for (int hr = 0; hr < M; ++hr)
    for (int hl = hr; hl >= 0; --hl)
      if (reachable[hl][hr]) state.push_back(IntervalState{(int)state.size(), hl, hr});
  S = (int)state.size();
  n2s.assign(M, std::vector<int>(M, -1));
  for (auto& s : state) n2s[s.l][s.r] = s.id;
  loop_state.clear();
  for (auto& s : state) if (reachable_as_loop[s.l][s.r]) loop_state.push_back(s.id);
  auto is_loop_node = [&](int h) { int c = node[h]; return c == 'z' || c == '.' || c == '*' || c == 'o'; };
  auto is_bg_node = [&](int h) { int c = node[h]; return c == 'z' || c == 'o' || c == '*'; };
  auto need = [&](int l, int r) {
    int id = (l >= 0 && r >= 0 && l < M && r < M) ? n2s[l][r] : -1;
    if (id < 0) throw std::runtime_error("nodes to state failed");
    return id;
  };
  right.assign(S, std::vector<int>());
  for (auto& s : state)
    if (is_loop_node(s.r))
      for (int h : edge_to[s.r])
        if (s.l <= h && reachable[s.l][h]) right[s.id].push_back(need(s.l, h));
  left.assign(S, std::vector<int>());
  for (auto& s : state)
    if (is_loop_node(s.l))
      for (int h : edge_to[s.l])
        if (h <= s.r && reachable[h][s.r]) left[need(h, s.r)].push_back(s.id);
  pairt.assign(S, std::vector<int>());
  for (int hr = 0; hr < M; ++hr)
    if (node[hr] == ')') {
      int kl = pair[hr];
      for (int hl : edge_to[kl]) {
        int s = need(hl, hr);
        for (int kr : edge_to[hr])
          if (reachable[kl][kr]) pairt[s].push_back(need(kl, kr));
      }
    }
  for (auto& s : state)
    if (is_bg_node(s.r))
      for (int hl : edge_from[s.l])
        if (is_bg_node(hl))
          for (int hr : edge_to[s.r])
            if (reachable[hl][hr]) pairt[s.id].push_back(need(hl, hr));
  quads.clear();
  for (int i2 : loop_state)
    for (int i3 : loop_state) {
      const IntervalState& s2 = state[i2];
      const IntervalState& s3 = state[i3];
      if (s3.r < s2.l || !reachable[s2.r][s3.l] || !reachable[s2.l][s3.r]) continue;
      quads.push_back(std::vector<int>{need(s2.l, s3.r), need(s2.r, s3.l), i2, i3});
    }
}

void FlatHMM::from(const ProfileHMM& h) {
  M = h.M; S = h.S;
  st_l.resize(S); st_r.resize(S); is_loop.assign(S, 0);
  for (auto& s : h.state) { st_l[s.id] = s.l; st_r[s.id] = s.r; }
  for (int id : h.loop_state) is_loop[id] = 1;
  auto csr = [&](const std::vector<std::vector<int>>& v, std::vector<int>& off, std::vector<int>& idx) {
    off.assign(1, 0); idx.clear();
    for (auto& r : v) { idx.insert(idx.end(), r.begin(), r.end()); off.push_back((int)idx.size()); }
  };
  csr(h.right, right_off, right_idx);
  csr(h.left, left_off, left_idx);
  csr(h.pairt, pair_off, pair_idx);
  quad_off.assign(S + 1, 0); quad_s1.clear(); quad_s2.clear(); quad_s3.clear();
  for (int s = 0; s < S; ++s) {
    for (auto& q : h.quads) if (q[0] == s) { quad_s1.push_back(q[1]); quad_s2.push_back(q[2]); quad_s3.push_back(q[3]); }
    quad_off[s + 1] = (int)quad_s1.size();
  }
  split_off.assign(S + 1, 0); split_left.clear(); split_right.clear();
  for (int s = 0; s < S; ++s) {
    for (int x = st_l[s]; x <= st_r[s]; ++x)
      if (h.reachable[st_l[s]][x] && h.reachable[x][st_r[s]]) {
        split_left.push_back(h.n2s[st_l[s]][x]); split_right.push_back(h.n2s[x][st_r[s]]);
      }
    split_off[s + 1] = (int)split_left.size();
  }
  auto targets = [&](const std::vector<int>& off, std::vector<int>& tgt) {
    tgt.clear();
    for (int s = 0; s < S; ++s) for (int a = off[s]; a < off[s + 1]; ++a) tgt.push_back(s);
  };
  targets(right_off, right_tgt); targets(left_off, left_tgt); targets(pair_off, pair_tgt);
  targets(quad_off, quad_tgt); targets(split_off, split_tgt);
  node = h.node; theta_id = h.theta_id;
  theta_off.assign(1, 0);
  for (int r : h.row_size) theta_off.push_back(theta_off.back() + r);
  s00 = h.n2s[0][0];
  s0M2 = M >= 2 ? h.n2s[0][M - 2] : -1;
  s0M1 = h.n2s[0][M - 1];
}

void FlatHMM::null_model() {
  M = 1; S = 1;
  st_l.assign(1, 0); st_r.assign(1, 0); is_loop.assign(1, 1);
  right_off = {0, 1}; right_idx = {0};
  left_off = {0, 1}; left_idx = {0};
  pair_off = {0, 1}; pair_idx = {0};
  quad_off = {0, 1}; quad_s1 = {0}; quad_s2 = {0}; quad_s3 = {0};
  split_off = {0, 1}; split_left = {0}; split_right = {0};
  right_tgt = {0}; left_tgt = {0}; pair_tgt = {0}; quad_tgt = {0}; split_tgt = {0};
  node = {'z'}; theta_id = {0}; theta_off = {0, 4};
  s00 = 0; s0M2 = -1; s0M1 = -1;
}

void quality_to_ws(const int* q, int n, double* ws) {
  int cnt[127 - 33];
  std::memset(cnt, 0, sizeof(cnt));
  for (int i = 0; i < n; ++i) {
    if (q[i] < 0 || q[i] >= 127 - 33) throw std::runtime_error("bad quality value");
    cnt[q[i]] += 1;
  }
  int mode = 0, best = std::numeric_limits<int>::lowest();
  for (int i = 0; i < 127 - 33; ++i)
    if (best <= cnt[i]) { mode = i; best = cnt[i]; }  // last maximum wins (util.hpp:231-241)
  for (int i = 0; i < n - 1; ++i) ws[i] = std::log((0.01 + double(q[i])) / (0.01 + mode));
  ws[n - 1] = q[n - 1] == 0 ? NEG_INF : 0.;
}

}  // namespace relem
