// Host front-end of comap_b200: the slice of Bio++'s application layer that CoMap's option
// files use on the hot path, re-implemented without Bio++.
//
// The reference links libbpp-core/seq/phyl >= 3.0.0 (CMakeLists.txt:113) for all of this;
// here the same option keys, input formats and conventions are parsed into the plain
// arrays the C ABI (include/comap_b200.h) takes.  Call sites mirrored: CoMap.cpp:120-152
// (BppApplication, getTree, getSubstitutionCount), CoETools.cpp:78-362 (readData).
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace host {

struct Error : std::runtime_error {
  explicit Error(const std::string& m) : std::runtime_error(m) {}
};

// ---------------------------------------------------------------- options (BppO syntax)
using Params = std::map<std::string, std::string>;

// comap param=FILE key=value ... ; later definitions and the command line win; $(var) is
// substituted; --seed=N / --noninteractive / --warning=N are application switches.
struct Application {
  Params params;
  uint64_t seed = 0;
  bool seed_given = false;
  std::string base_dir; // directory of the option file: relative input paths resolve against the cwd, as Bio++ does
};
void parse_option_text(const std::string& text, Params& out, const std::string& dir, int depth = 0);
Application parse_command_line(int argc, const char* const* argv);

std::string get_string(const Params& p, const std::string& key, const std::string& def);
bool get_bool(const Params& p, const std::string& key, bool def);
double get_double(const Params& p, const std::string& key, double def);
long get_int(const Params& p, const std::string& key, long def);
// ApplicationTools::getAFilePath: "none" (or an empty value) means no file
std::string get_path(const Params& p, const std::string& key, const std::string& def);

// KeyvalTools::parseProcedure: "Name(k=v, k2=Inner(a=b))" -> name + arguments
struct Procedure {
  std::string name;
  Params args;
};
Procedure parse_procedure(const std::string& desc);

// ---------------------------------------------------------------- alphabets / alignments
struct Alphabet {
  std::string name;          // DNA | RNA | Protein
  std::string states;        // resolved states in model order
  uint32_t mask_of(char c) const; // bitmask of compatible states; 0 = not in the alphabet
  bool is_gap(char c) const { return c == '-'; }
  bool is_unknown(char c) const; // the alphabet's "unknown" character (N / X / ?)
  bool is_resolved(char c) const;
};
Alphabet make_alphabet(const std::string& name);

struct Alignment {
  std::vector<std::string> names;
  std::vector<std::string> seqs;                       // equal lengths
  std::map<std::string, std::vector<std::pair<int, int>>> selections; // Mase site selections, 1-based inclusive
  size_t length() const { return seqs.empty() ? 0 : seqs[0].size(); }
};
Alignment read_alignment(const std::string& path, const std::string& format_desc);
Alignment parse_mase(const std::string& text);
Alignment parse_fasta(const std::string& text);
Alignment parse_phylip(const std::string& text, bool sequential, bool extended);

// SequenceApplicationTools::getSitesToAnalyse + CoETools.cpp:347-360 (input.remove_const):
// returns the 0-based columns kept, in order.  `after_selection` receives the columns that
// survive sites_to_use (the "sites to analyse" before constant sites are dropped).
std::vector<int> select_sites(const Alignment& aln, const Alphabet& alpha, const Params& p,
                              const std::string& format_desc, std::vector<int>* after_selection);
bool site_is_constant(const Alignment& aln, const Alphabet& alpha, int col);
bool site_is_complete(const Alignment& aln, const Alphabet& alpha, int col);

// ---------------------------------------------------------------- tree
struct Tree {
  std::vector<int32_t> parent; // post-order ids, root last (parent -1)
  std::vector<double> brlen;
  std::vector<std::string> name; // leaf names ("" for inner nodes)
  std::vector<int> leaves;       // node ids of the leaves in id order (= alignment row order)
  int n_root_children = 0;
  bool was_unrooted = false;
};
Tree parse_newick(const std::string& text);

// ---------------------------------------------------------------- model + rates
struct Model {
  std::string name;
  int A = 0;
  std::vector<double> Q, pi; // generator row-major, normalised to one substitution per unit time
  std::string warning;       // printed by the front-end after "Substitution model" (e.g. an unverified bundled table)
};
Model make_model(const std::string& desc, const Alphabet& alpha, const std::string& data_dir);

struct RateDist {
  std::string name;
  std::vector<double> rates, probs;
  // the continuous distribution behind the classes (simulations.continuous = yes): 1 constant, 2 gamma, 3 invariant + gamma
  int cont_kind = 1;
  double alpha = 1., p_inv = 0.;
};
RateDist make_rate_distribution(const std::string& desc);
// nijt=...(weight=Diff(index1=Volume, symmetrical=no)): AlphabetIndex2 weights of the weighted
// substitution count, A*A row-major w[x][y] (empty: no weights).  *symmetric reports isSymmetric().
std::vector<double> make_count_weights(const std::string& desc, const Alphabet& alpha, bool* symmetric,
                                       const std::string& data_dir);

// regularised lower incomplete gamma P(a, x) and its inverse (used by Gamma(n, alpha))
double pgamma(double x, double a);
double qgamma(double p, double a);

std::string read_file(const std::string& path);
std::string trim(const std::string& s);
std::string lower(const std::string& s);

} // namespace host
