// host_model.hpp -- host-side (C++14) model construction for librelem: energy tables and the motif automaton,
// flattened into the arrays the sm_100a kernels read.  Model set-up only; no DP runs on the host.
//
// What it mirrors in the reference (behaviour, not code):
//   EnergyParam tables + parse rules          RNAelem/energy_param.hpp:61-114,159-183,423-640
//   ProfileHMM automaton                      RNAelem/profile_hmm.hpp:188-463
#ifndef RELEM_HOST_MODEL_HPP
#define RELEM_HOST_MODEL_HPP
#include <cstdint>
#include <string>
#include <vector>

namespace relem {

// ---------------------------------------------------------------------------------------------------------
// Integer (dcal/mol) parameter set exactly as the reference's reader sees it: only the sub-blocks it reads,
// RELEM_EINF for "INF".  Filled either from the built-in sets (energy_data.inc) or by parse_param_text().
struct EnergyInts {
  static const int INF = 1000000;
  int stack[6][6];          // pair types 1..6
  int mm_h[6][25], mm_i[6][25], mm_1ni[6][25], mm_23i[6][25];  // types 1..6, [5][5] bases
  int mm_m[7][25], mm_ext[7][25];                              // types 1..7 (the reference reads 7 rows)
  int dangle5[7][5], dangle3[7][5];                            // types 1..7
  int int11[7][7][25];       // types 1..7 x 1..7
  int int21[7][7][125];
  int int22[6][6][256];      // types 1..6, bases 1..4 each
  int hairpin[31], bulge[31], interior[31];
  int ninio_f, ninio_max, ml_base, ml_closing, ml_intern, term_au;
  double lxc37;
  std::vector<std::string> tri, tetra, hexa;   // special hairpins incl. closing pair
  std::vector<int> tri_e, tetra_e, hexa_e;
};
bool builtin_energy_ints(const std::string& name, EnergyInts& out);  // "~T2004~" / "~A2007~"
bool parse_param_text(const std::string& text, EnergyInts& out, std::string& err);

// log-Boltzmann weights (-E*10/kT, kT at 37 C) in the reference's array shapes (energy_param.hpp:61-85).
// -inf = forbidden.  Everything the DP needs is plain lookups + additions on these.
struct EnergyTables {
  double hairpin[31];
  double mismatch_h[7][5][5], mismatch_i[7][5][5], mismatch_m[8][5][5], mismatch_1ni[7][5][5],
      mismatch_23i[7][5][5], mismatch_ext[8][5][5];
  double stack[7][7];
  double bulge[31];
  double term_au;
  double int11[8][8][5][5];
  double int21[8][8][5][5][5];
  double int22[8][8][5][5][5][5];
  double internal[31];
  double dangle5[8][5], dangle3[8][5];
  double ninio[31];
  double mlintern, mlclosing, ml_base, lxc37;
  std::vector<std::string> tri, tetra, hexa;
  std::vector<double> tri_w, tetra_w, hexa_w;
  void build(const EnergyInts& e);
  // hairpin length term for loop size d (energy_param.hpp:715-719), any d >= 0
  double hairpin_len(int d) const;
};

// ---------------------------------------------------------------------------------------------------------
struct IntervalState { int id, l, r; };

struct ProfileHMM {
  std::string pattern, reg_pattern;
  int M = 0, S = 0;
  std::vector<int> node;       // chars: 'z' pattern... 'o'
  std::vector<int> pair;       // partner of a bracket node or -1
  std::vector<int> theta_id;   // row of theta per node, -1 for '('
  std::vector<int> row_size;   // theta row sizes (row 0 = background)
  std::vector<std::vector<int>> edge_to, edge_from;
  std::vector<std::vector<char>> reachable, reachable_as_loop;
  std::vector<IntervalState> state;
  std::vector<std::vector<int>> n2s;               // [l][r] -> id or -1
  std::vector<int> loop_state;                     // ids
  std::vector<std::vector<int>> right, left, pairt;  // transition lists (ids), reference list order
  std::vector<std::vector<int>> quads;             // (s, s1, s2, s3) ids, reference list order
  int n_theta() const { int n = 0; for (int r : row_size) n += r; return n; }
  // throws std::runtime_error on a malformed pattern
  void build(const std::string& pattern);
};

// Flattened automaton for the device (all int32).  Per target state s:
//   right/left/pair: CSR over s
//   quad lists grouped by target s (order within s = reference list order): q_s1,q_s2,q_s3
//   split lists grouped by s, h ascending: sp_left=(s.l,h), sp_right=(h,s.r)
struct FlatHMM {
  int M, S;
  std::vector<int> st_l, st_r, is_loop;
  std::vector<int> right_off, right_idx, left_off, left_idx, pair_off, pair_idx;
  std::vector<int> right_tgt, left_tgt, pair_tgt, quad_tgt, split_tgt;  // target state of every flat entry
  std::vector<int> quad_off, quad_s1, quad_s2, quad_s3;
  std::vector<int> split_off, split_left, split_right;
  std::vector<int> node, theta_id, theta_off;  // theta_off[row] = offset of the row in theta_flat
  int s00, s0M2, s0M1;                          // n2s(0,0), n2s(0,M-2), n2s(0,M-1) (or -1)
  void from(const ProfileHMM& h);
  void null_model();  // one state, no emissions: the energy-only grammar of EnergyModel::calc_BPP
};

// RNAelem::set_ws (motif_model.hpp:62-70): quality values (L+1, already minus the base) -> ws (L+1 doubles,
// the last one is the "contains motif" flag: -inf for quality 0, else 0).
void quality_to_ws(const int* qual, int n, double* ws);

}  // namespace relem
#endif
