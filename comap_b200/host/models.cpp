// Substitution models and rate distributions in Bio++'s parameterisation
// (PhylogeneticsApplicationTools::getSubstitutionModel / getRateDistribution, called at
// CoETools.cpp:113,122; conventions in SURVEY.md appendix A).
#include "bpp.h"
#include <cmath>
#include <cstdlib>
#include <sstream>

namespace host {

namespace {

// normalise Q (rows sum to zero, one expected substitution per unit time)
void finish(Model& m) {
  const int A = m.A;
  double s = 0.;
  for (int i = 0; i < A; i++) s += m.pi[i];
  for (int i = 0; i < A; i++) m.pi[i] /= s;
  for (int i = 0; i < A; i++) {
    double r = 0.;
    for (int j = 0; j < A; j++)
      if (j != i) r += m.Q[i * A + j];
    m.Q[i * A + i] = -r;
  }
  double scale = 0.;
  for (int i = 0; i < A; i++) scale -= m.pi[i] * m.Q[i * A + i];
  for (auto& q : m.Q) q /= scale;
}

double arg(const Procedure& p, const char* key, double def) {
  auto it = p.args.find(key);
  if (it == p.args.end()) return def;
  return std::atof(it->second.c_str());
}

// Bio++ nucleotide frequency parameters -> (A, C, G, T)
void freqs(double theta, double theta1, double theta2, double* pi) {
  pi[0] = theta1 * (1. - theta);
  pi[1] = (1. - theta2) * theta;
  pi[2] = theta2 * theta;
  pi[3] = (1. - theta1) * (1. - theta);
}

// exchangeabilities s (symmetric, A x A) times target frequency
void from_exchangeabilities(Model& m, const std::vector<double>& s) {
  const int A = m.A;
  m.Q.assign((size_t)A * A, 0.);
  for (int i = 0; i < A; i++)
    for (int j = 0; j < A; j++)
      if (i != j) m.Q[i * A + j] = s[i * A + j] * m.pi[j];
  finish(m);
}

// PAML .dat: lower triangle of exchangeabilities (rows 2..A) then A frequencies; '#' comments
void read_paml(const std::string& path, Model& m) {
  std::istringstream in(read_file(path));
  std::string line;
  std::vector<double> v;
  while (std::getline(in, line)) {
    size_t h = line.find('#');
    if (h != std::string::npos) line = line.substr(0, h);
    std::istringstream ls(line);
    double x;
    while (ls >> x) v.push_back(x);
  }
  const int A = 20;
  if (v.size() < (size_t)(A * (A - 1) / 2 + A)) throw Error("'" + path + "' does not hold a 20-state PAML matrix");
  std::vector<double> s((size_t)A * A, 0.);
  size_t k = 0;
  for (int i = 1; i < A; i++)
    for (int j = 0; j < i; j++) s[i * A + j] = s[j * A + i] = v[k++];
  m.A = A;
  m.pi.assign(v.begin() + k, v.begin() + k + A);
  double t = 0.;
  for (double x : m.pi) t += x;
  for (double& x : m.pi) x /= t;
  from_exchangeabilities(m, s);
}

} // namespace

Model make_model(const std::string& desc, const Alphabet& alpha, const std::string& data_dir) {
  Procedure p = parse_procedure(desc);
  Model m;
  m.name = p.name;
  const std::string n = lower(p.name);
  const bool nuc = alpha.name != "Protein";
  if (n == "jc69" || n == "k80" || n == "hky85" || n == "t92" || n == "tn93" || n == "gtr" || n == "f84") {
    if (!nuc) throw Error("model " + p.name + " needs a nucleotide alphabet");
    m.A = 4;
    m.pi.assign(4, 0.25);
    std::vector<double> s(16, 1.);
    // index: A=0 C=1 G=2 T=3; transitions A<->G, C<->T
    auto set = [&](int i, int j, double v) { s[i * 4 + j] = s[j * 4 + i] = v; };
    if (n == "k80") {
      double k = arg(p, "kappa", 1.);
      set(0, 2, k); set(1, 3, k);
    } else if (n == "hky85") {
      double k = arg(p, "kappa", 1.);
      freqs(arg(p, "theta", .5), arg(p, "theta1", .5), arg(p, "theta2", .5), m.pi.data());
      set(0, 2, k); set(1, 3, k);
    } else if (n == "t92") {
      double k = arg(p, "kappa", 1.), th = arg(p, "theta", .5);
      m.pi = {(1. - th) / 2., th / 2., th / 2., (1. - th) / 2.};
      set(0, 2, k); set(1, 3, k);
    } else if (n == "tn93") {
      double k1 = arg(p, "kappa1", 1.), k2 = arg(p, "kappa2", 1.);
      freqs(arg(p, "theta", .5), arg(p, "theta1", .5), arg(p, "theta2", .5), m.pi.data());
      set(0, 2, k1); set(1, 3, k2);
    } else if (n == "gtr") {
      // Bio++: a = C<->T, b = A<->T, c = G<->T, d = A<->C, e = C<->G, A<->G = 1
      freqs(arg(p, "theta", .5), arg(p, "theta1", .5), arg(p, "theta2", .5), m.pi.data());
      set(1, 3, arg(p, "a", 1.)); set(0, 3, arg(p, "b", 1.)); set(2, 3, arg(p, "c", 1.));
      set(0, 1, arg(p, "d", 1.)); set(1, 2, arg(p, "e", 1.)); set(0, 2, 1.);
    } else if (n == "f84") {
      throw Error("model F84 is not supported; use HKY85");
    }
    from_exchangeabilities(m, s);
    return m;
  }
  if (nuc) throw Error("model '" + p.name + "' is not supported for nucleotides (JC69, K80, HKY85, T92, TN93, GTR)");
  if (n == "jc69") {
    m.A = 20;
    m.pi.assign(20, 0.05);
    from_exchangeabilities(m, std::vector<double>(400, 1.));
    return m;
  }
  // empirical protein matrices: PAML files under comap_b200/data.  JTT92 is bundled and validated by the reference's
  // goldens; LG08 (examples/simple/*/comap.bpp:33) is bundled UNVERIFIED (re-typed from memory, no golden exists for
  // it) and says so at every use; any other Bio++ table: drop <name>.dat next to them or use Empirical(file=...)
  std::string file = n == "jtt92" ? "jtt92_dcmut.dat" : n + ".dat";
  std::string path = data_dir + "/" + file;
  if (n == "empirical") { // Bio++: model = Empirical(name=..., file=<PAML .dat>)
    path = get_string(p.args, "file", "none");
    if (path == "none") throw Error("model Empirical(...) needs file=<PAML exchangeability file>");
  }
  if (n == "lg08")
    m.warning = "WARNING!!! model=LG08 uses the bundled table " + path + ", re-typed from memory and NOT validated against Bio++ "
                "(no golden output exists for it): replace it with PAML's lg.dat or use model=Empirical(file=...) for production.";
  try {
    read_paml(path, m);
  } catch (const Error&) {
    throw Error("model '" + p.name + "': empirical matrix file '" + path + "' is not available (bundled: JTT92)");
  }
  return m;
}

// ---------------------------------------------------------------- gamma functions
double pgamma(double x, double a) { // regularised lower incomplete gamma P(a, x)
  if (x <= 0.) return 0.;
  const double lg = std::lgamma(a);
  if (x < a + 1.) { // series
    double ap = a, del = 1. / a, sum = del;
    for (int n = 0; n < 10000; n++) {
      ap += 1.;
      del *= x / ap;
      sum += del;
      if (std::fabs(del) < std::fabs(sum) * 1e-17) break;
    }
    return sum * std::exp(-x + a * std::log(x) - lg);
  }
  // continued fraction (modified Lentz)
  const double tiny = 1e-300;
  double b = x + 1. - a, c = 1. / tiny, d = 1. / b, h = d;
  for (int i = 1; i < 10000; i++) {
    double an = -(double)i * ((double)i - a);
    b += 2.;
    d = an * d + b;
    if (std::fabs(d) < tiny) d = tiny;
    c = b + an / c;
    if (std::fabs(c) < tiny) c = tiny;
    d = 1. / d;
    double del = d * c;
    h *= del;
    if (std::fabs(del - 1.) < 1e-17) break;
  }
  return 1. - std::exp(-x + a * std::log(x) - lg) * h;
}

double qgamma(double p, double a) { // x with P(a, x) = p
  if (p <= 0.) return 0.;
  if (p >= 1.) return INFINITY;
  double lo = 0., hi = a + 10. * std::sqrt(a) + 10.;
  while (pgamma(hi, a) < p) hi *= 2.;
  for (int i = 0; i < 400; i++) {
    double mid = 0.5 * (lo + hi);
    if (pgamma(mid, a) < p) lo = mid; else hi = mid;
    if (hi - lo <= 1e-16 * hi) break;
  }
  double x = 0.5 * (lo + hi);
  // Newton polish on the smooth function
  for (int i = 0; i < 4; i++) {
    double f = pgamma(x, a) - p;
    double dens = std::exp(-x + (a - 1.) * std::log(x) - std::lgamma(a));
    if (!(dens > 0.)) break;
    double nx = x - f / dens;
    if (!(nx > lo && nx < hi)) break;
    x = nx;
  }
  return x;
}

RateDist make_rate_distribution(const std::string& desc) {
  Procedure p = parse_procedure(desc);
  RateDist r;
  r.name = p.name;
  const std::string n = lower(p.name);
  if (n == "constant" || n == "uniform" || n.empty()) {
    r.rates = {1.};
    r.probs = {1.};
    return r;
  }
  if (n == "gamma") {
    int k = (int)arg(p, "n", 4);
    double alpha = arg(p, "alpha", 1.);
    if (k < 1) throw Error("Gamma: n must be positive");
    if (!(alpha > 0.)) throw Error("Gamma: alpha must be positive");
    // n equiprobable classes; class rate = conditional mean of its quantile bin (beta = alpha)
    std::vector<double> e(k + 1);
    for (int i = 0; i <= k; i++) {
      double q = i == 0 ? 0. : i == k ? INFINITY : qgamma((double)i / k, alpha); // in units of 1/beta
      e[i] = i == 0 ? 0. : i == k ? 1. : pgamma(q, alpha + 1.);
    }
    for (int i = 0; i < k; i++) {
      r.rates.push_back((double)k * (e[i + 1] - e[i]));
      r.probs.push_back(1. / k);
    }
    r.cont_kind = 2; r.alpha = alpha;
    return r;
  }
  if (n == "invariant") {
    auto it = p.args.find("dist");
    if (it == p.args.end()) throw Error("Invariant: missing dist= argument");
    RateDist d = make_rate_distribution(it->second);
    double pi = arg(p, "p", 0.);
    if (!(pi >= 0. && pi < 1.)) throw Error("Invariant: p must be in [0, 1)");
    r.rates.push_back(0.);
    r.probs.push_back(pi);
    for (size_t i = 0; i < d.rates.size(); i++) {
      r.rates.push_back(d.rates[i] / (1. - pi));
      r.probs.push_back(d.probs[i] * (1. - pi));
    }
    r.cont_kind = d.cont_kind == 2 ? 3 : (pi > 0. ? 0 : 1); // Invariant + Constant has no continuous sampler here
    r.alpha = d.alpha; r.p_inv = pi;
    return r;
  }
  throw Error("rate distribution '" + desc + "' is not supported (Constant, Gamma, Invariant)");
}

} // namespace host

// ---------------------------------------------------------------------------------------------
// Weights of the weighted substitution count (PhylogeneticsApplicationTools::getSubstitutionCount,
// reference call site CoMap.cpp:152; examples/simple/ProteinPairCompensation/comap.bpp:48).
// Diff(index1=<AlphabetIndex1>, symmetrical=yes|no) = SimpleIndexDistance: w[x][y] = index[y] -
// index[x] (absolute value when symmetrical).  Bundled indices (protein alphabet, order
// A R N D C Q E G H I L K M F P S T W Y V): Grantham (1974) volume and polarity, Klein net charge.
namespace host {
namespace {
const double kGranthamVolume[20] = {31, 124, 56, 54, 55, 85, 83, 3, 96, 111, 111, 119, 105, 132, 32.5, 32, 61, 170, 136, 84};
const double kGranthamPolarity[20] = {8.1, 10.5, 11.6, 13.0, 5.5, 10.5, 12.3, 9.0, 10.4, 5.2,
                                      4.9, 11.3, 5.7, 5.2, 8.0, 9.2, 8.6, 5.4, 6.2, 5.9};
const double kKleinCharge[20] = {0, 1, 0, -1, 0, 0, -1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0};
}

std::vector<double> make_count_weights(const std::string& desc, const Alphabet& alpha, bool* symmetric,
                                       const std::string& data_dir) {
  if (symmetric) *symmetric = true;
  if (desc.empty() || desc == "None" || desc == "none") return {};
  Procedure p = parse_procedure(desc);
  const size_t A = alpha.states.size();
  if (p.name == "AAdist") {
    // AAdist(type=grantham, sym=yes): Grantham's chemical distance (examples/Proteins/Benchmark/CoMap/analyse.sh);
    // the table in data/grantham.dat was recovered from the reference's golden vectors
    // (tests/golden/recover_grantham.py).  The signed variant (sym=no) is Bio++-specific and not available.
    if (A != 20) throw Error("weight=AAdist(...) needs the protein alphabet");
    if (lower(get_string(p.args, "type", "grantham")) != "grantham")
      throw Error("weight=AAdist(type=" + get_string(p.args, "type", "") + "): only type=grantham is bundled");
    if (!get_bool(p.args, "sym", true)) throw Error("weight=AAdist(sym=no) is not available in this build");
    std::istringstream in(read_file(data_dir + "/grantham.dat"));
    std::vector<double> w;
    std::string line;
    while (std::getline(in, line)) {
      if (line.empty() || line[0] == '#') continue;
      std::istringstream ls(line);
      double v;
      while (ls >> v) w.push_back(v);
    }
    if (w.size() != 400) throw Error("grantham.dat: expected a 20 x 20 table");
    return w;
  }
  if (p.name != "Diff")
    throw Error("weight=" + p.name + " is not available in this build (Diff(index1=Volume|Polarity|Charge, symmetrical=yes|no), "
                "AAdist(type=grantham, sym=yes))");
  if (A != 20) throw Error("weight=Diff(...) needs the protein alphabet (the bundled indices are amino-acid properties)");
  std::string idx = get_string(p.args, "index1", "None");
  const double* v = nullptr;
  if (idx == "Volume" || idx == "GranthamVolume") v = kGranthamVolume;
  else if (idx == "Polarity" || idx == "GranthamPolarity") v = kGranthamPolarity;
  else if (idx == "Charge" || idx == "KleinCharge") v = kKleinCharge;
  else throw Error("weight=Diff(index1=" + idx + "): unknown index (Volume, Polarity, Charge)");
  const bool sym = get_bool(p.args, "symmetrical", true);
  if (symmetric) *symmetric = sym;
  std::vector<double> w(A * A);
  for (size_t x = 0; x < A; x++)
    for (size_t y = 0; y < A; y++) {
      const double d = v[y] - v[x];
      w[x * A + y] = sym ? std::fabs(d) : d;
    }
  return w;
}
} // namespace host
