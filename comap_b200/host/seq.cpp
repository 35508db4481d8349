// Alphabets, alignment readers (Mase / Phylip / Fasta) and site selection -- what
// SequenceApplicationTools::{getAlphabet,getSiteContainer,getSitesToAnalyse} and
// SiteTools::{isConstant,isComplete} do for CoETools::readData (CoETools.cpp:91-93,347-360).
#include "bpp.h"
#include <algorithm>
#include <cstdlib>
#include <sstream>

namespace host {

// ---------------------------------------------------------------- alphabets
static const char* kProtein = "ARNDCQEGHILKMFPSTWYV";

Alphabet make_alphabet(const std::string& desc) {
  Procedure p = parse_procedure(desc);
  Alphabet a;
  std::string n = lower(p.name);
  if (n == "dna") { a.name = "DNA"; a.states = "ACGT"; }
  else if (n == "rna") { a.name = "RNA"; a.states = "ACGU"; }
  else if (n == "protein") { a.name = "Protein"; a.states = kProtein; }
  else throw Error("alphabet '" + desc + "' is not supported (DNA, RNA, Protein)");
  return a;
}

bool Alphabet::is_unknown(char c) const {
  c = (char)toupper((unsigned char)c);
  if (name == "Protein") return c == 'X' || c == '?' || c == 'O' || c == '0';
  return c == 'N' || c == 'X' || c == '?' || c == 'O' || c == '0';
}

uint32_t Alphabet::mask_of(char ch) const {
  char c = (char)toupper((unsigned char)ch);
  const uint32_t all = (1u << states.size()) - 1u;
  if (c == '-' || is_unknown(c)) return all;
  if (name != "Protein") {
    if (c == 'T' || c == 'U') c = states[3];
    size_t k = states.find(c);
    if (k != std::string::npos) return 1u << k;
    auto m = [&](const char* s) {
      uint32_t r = 0;
      for (; *s; s++) r |= 1u << std::string("ACGT").find(*s);
      return r;
    };
    switch (c) { // IUPAC
      case 'R': return m("AG"); case 'Y': return m("CT"); case 'S': return m("CG"); case 'W': return m("AT");
      case 'K': return m("GT"); case 'M': return m("AC"); case 'B': return m("CGT"); case 'D': return m("AGT");
      case 'H': return m("ACT"); case 'V': return m("ACG");
      default: return 0;
    }
  }
  size_t k = states.find(c);
  if (k != std::string::npos) return 1u << k;
  auto m = [&](const char* s) {
    uint32_t r = 0;
    for (; *s; s++) r |= 1u << states.find(*s);
    return r;
  };
  switch (c) {
    case 'B': return m("DN"); case 'Z': return m("EQ"); case 'J': return m("IL");
    case '*': return all;
    default: return 0;
  }
}

bool Alphabet::is_resolved(char c) const {
  uint32_t m = mask_of(c);
  return c != '-' && m != 0 && (m & (m - 1)) == 0;
}

// ---------------------------------------------------------------- readers
static std::vector<std::string> lines_of(const std::string& text) {
  std::vector<std::string> out;
  std::istringstream in(text);
  std::string l;
  while (std::getline(in, l)) {
    if (!l.empty() && l.back() == '\r') l.pop_back();
    out.push_back(l);
  }
  return out;
}
static std::string strip_spaces(const std::string& s) {
  std::string r;
  for (char c : s) if (!isspace((unsigned char)c)) r += c;
  return r;
}
static void check_aligned(const Alignment& a) {
  if (a.seqs.empty()) throw Error("alignment holds no sequence");
  for (size_t i = 1; i < a.seqs.size(); i++)
    if (a.seqs[i].size() != a.seqs[0].size())
      throw Error("sequences are not aligned: '" + a.names[i] + "' has " + std::to_string(a.seqs[i].size()) +
                  " characters, '" + a.names[0] + "' has " + std::to_string(a.seqs[0].size()));
}

Alignment parse_mase(const std::string& text) {
  Alignment a;
  auto L = lines_of(text);
  size_t i = 0;
  // header: ";;" lines; site selections look like
  //   ;;# of segments=9 SelectedSites
  //   ;; 23,31 35,66 ...
  while (i < L.size() && L[i].rfind(";;", 0) == 0) {
    const std::string& h = L[i];
    size_t k = h.find("# of segments=");
    if (k != std::string::npos) {
      std::istringstream ss(h.substr(k + 14));
      int nseg = 0;
      std::string name;
      ss >> nseg >> name;
      std::vector<std::pair<int, int>> segs;
      size_t j = i + 1;
      while ((int)segs.size() < nseg && j < L.size() && L[j].rfind(";;", 0) == 0) {
        std::istringstream rs(L[j].substr(2));
        std::string tok;
        while (rs >> tok) {
          size_t c = tok.find(',');
          if (c == std::string::npos) continue;
          segs.push_back({atoi(tok.substr(0, c).c_str()), atoi(tok.substr(c + 1).c_str())});
        }
        j++;
      }
      a.selections[name] = segs;
      i = j;
      continue;
    }
    i++;
  }
  while (i < L.size()) {
    if (L[i].empty()) { i++; continue; }
    if (L[i][0] != ';') throw Error("Mase: expected a comment line before sequence name at line " + std::to_string(i + 1));
    while (i < L.size() && !L[i].empty() && L[i][0] == ';') i++;
    if (i >= L.size()) break;
    a.names.push_back(trim(L[i++]));
    std::string s;
    while (i < L.size() && (L[i].empty() || L[i][0] != ';')) s += strip_spaces(L[i++]);
    a.seqs.push_back(s);
  }
  check_aligned(a);
  return a;
}

Alignment parse_fasta(const std::string& text) {
  Alignment a;
  for (auto& l : lines_of(text)) {
    if (l.empty()) continue;
    if (l[0] == '>') {
      std::string n = trim(l.substr(1));
      size_t sp = n.find_first_of(" \t");
      a.names.push_back(sp == std::string::npos ? n : n.substr(0, sp));
      a.seqs.push_back("");
    } else if (!a.seqs.empty()) a.seqs.back() += strip_spaces(l);
  }
  check_aligned(a);
  return a;
}

Alignment parse_phylip(const std::string& text, bool sequential, bool extended) {
  Alignment a;
  auto L = lines_of(text);
  size_t i = 0;
  while (i < L.size() && trim(L[i]).empty()) i++;
  if (i >= L.size()) throw Error("Phylip: empty file");
  std::istringstream hs(L[i++]);
  long n = 0, len = 0;
  hs >> n >> len;
  if (n <= 0 || len <= 0) throw Error("Phylip: bad header line");
  auto split_name = [&](const std::string& l, std::string& name, std::string& rest) {
    if (extended) {
      std::string t = l;
      size_t b = t.find_first_not_of(" \t");
      if (b == std::string::npos) { name = ""; rest = ""; return; }
      size_t e = t.find_first_of(" \t", b);
      name = t.substr(b, e == std::string::npos ? std::string::npos : e - b);
      rest = e == std::string::npos ? "" : t.substr(e);
    } else {
      name = trim(l.substr(0, std::min<size_t>(10, l.size())));
      rest = l.size() > 10 ? l.substr(10) : "";
    }
  };
  if (sequential) {
    while ((long)a.names.size() < n) {
      while (i < L.size() && trim(L[i]).empty()) i++;
      if (i >= L.size()) throw Error("Phylip: fewer sequences than announced");
      std::string name, rest;
      split_name(L[i++], name, rest);
      std::string s = strip_spaces(rest);
      while ((long)s.size() < len && i < L.size()) s += strip_spaces(L[i++]);
      a.names.push_back(name);
      a.seqs.push_back(s);
    }
  } else {
    for (long k = 0; k < n; k++) {
      while (i < L.size() && trim(L[i]).empty()) i++;
      if (i >= L.size()) throw Error("Phylip: fewer sequences than announced");
      std::string name, rest;
      split_name(L[i++], name, rest);
      a.names.push_back(name);
      a.seqs.push_back(strip_spaces(rest));
    }
    long k = 0;
    for (; i < L.size(); i++) {
      if (trim(L[i]).empty()) { k = 0; continue; }
      if (k < n) a.seqs[k++] += strip_spaces(L[i]);
    }
  }
  for (auto& s : a.seqs)
    if ((long)s.size() != len) throw Error("Phylip: a sequence does not have the announced length");
  check_aligned(a);
  return a;
}

Alignment read_alignment(const std::string& path, const std::string& format_desc) {
  Procedure f = parse_procedure(format_desc);
  std::string n = lower(f.name);
  std::string text = read_file(path);
  if (n == "mase") return parse_mase(text);
  if (n == "fasta") return parse_fasta(text);
  if (n == "phylip") {
    std::string order = lower(get_string(f.args, "order", "interleaved"));
    std::string type = lower(get_string(f.args, "type", "classic"));
    return parse_phylip(text, order == "sequential", type == "extended");
  }
  throw Error("sequence format '" + format_desc + "' is not supported (Mase, Fasta, Phylip)");
}

// ---------------------------------------------------------------- site selection
// SiteTools::isConstant(site, ignoreUnknown = true): gaps and the unknown character are
// skipped; any other pair of different characters makes the site variable.
bool site_is_constant(const Alignment& aln, const Alphabet& alpha, int col) {
  char first = 0;
  for (auto& s : aln.seqs) {
    char c = (char)toupper((unsigned char)s[col]);
    if (alpha.name != "Protein" && c == 'U') c = 'T';
    if (alpha.is_gap(c) || alpha.is_unknown(c)) continue;
    if (!first) first = c;
    else if (c != first) return false;
  }
  return true;
}
bool site_is_complete(const Alignment& aln, const Alphabet& alpha, int col) {
  for (auto& s : aln.seqs)
    if (!alpha.is_resolved(s[col])) return false;
  return true;
}

std::vector<int> select_sites(const Alignment& aln, const Alphabet& alpha, const Params& p,
                              const std::string& format_desc, std::vector<int>* after_selection) {
  const int L = (int)aln.length();
  std::vector<int> cols;
  // Mase site selection (input.sequence.format = Mase(site_selection=NAME))
  Procedure f = parse_procedure(format_desc);
  std::string sel = get_string(f.args, "site_selection", "");
  if (!sel.empty() && lower(f.name) == "mase") {
    auto it = aln.selections.find(sel);
    if (it == aln.selections.end()) throw Error("Mase site selection '" + sel + "' not found in the file header");
    for (auto& seg : it->second)
      for (int c = seg.first; c <= seg.second; c++)
        if (c >= 1 && c <= L) cols.push_back(c - 1);
  } else {
    for (int c = 0; c < L; c++) cols.push_back(c);
  }
  for (auto& s : aln.seqs)
    for (int c : cols)
      if (alpha.mask_of(s[c]) == 0)
        throw Error(std::string("character '") + s[c] + "' is not in the " + alpha.name + " alphabet");
  std::string use = lower(get_string(p, "input.sequence.sites_to_use", "complete"));
  std::vector<int> kept;
  if (use == "all") {
    std::string mg = get_string(p, "input.sequence.max_gap_allowed", "100%");
    double limit;
    if (!mg.empty() && mg.back() == '%') limit = atof(mg.substr(0, mg.size() - 1).c_str()) / 100. * (double)aln.seqs.size();
    else limit = atof(mg.c_str());
    for (int c : cols) {
      int gaps = 0;
      for (auto& s : aln.seqs) gaps += s[c] == '-';
      if ((double)gaps <= limit) kept.push_back(c);
    }
  } else if (use == "nogap") {
    for (int c : cols) {
      bool ok = true;
      for (auto& s : aln.seqs) ok = ok && s[c] != '-';
      if (ok) kept.push_back(c);
    }
  } else if (use == "complete") {
    for (int c : cols)
      if (site_is_complete(aln, alpha, c)) kept.push_back(c);
  } else throw Error("input.sequence.sites_to_use = '" + use + "' is not one of all, nogap, complete");
  if (after_selection) *after_selection = kept;
  if (get_bool(p, "input.remove_const", true)) {
    std::vector<int> var;
    for (int c : kept)
      if (!site_is_constant(aln, alpha, c)) var.push_back(c);
    kept = var;
  }
  return kept;
}

} // namespace host
