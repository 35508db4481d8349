// BppO option syntax (what BppApplication / AttributesTools / ApplicationTools /
// KeyvalTools provide to CoMap.cpp:120 and every get*Parameter call in CoETools.cpp).
#include "bpp.h"
#include <cstdlib>
#include <fstream>
#include <sstream>

namespace host {

std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && isspace((unsigned char)s[a])) a++;
  while (b > a && isspace((unsigned char)s[b - 1])) b--;
  return s.substr(a, b - a);
}
std::string lower(const std::string& s) {
  std::string r = s;
  for (auto& c : r) c = (char)tolower((unsigned char)c);
  return r;
}
std::string read_file(const std::string& path) {
  std::ifstream in(path, std::ios::binary);
  if (!in) throw Error("cannot open file '" + path + "'");
  std::ostringstream ss;
  ss << in.rdbuf();
  return ss.str();
}

static std::string dir_of(const std::string& path) {
  size_t k = path.find_last_of('/');
  return k == std::string::npos ? std::string(".") : path.substr(0, k);
}

// One logical line = physical lines joined on a trailing backslash, comment stripped.
void parse_option_text(const std::string& text, Params& out, const std::string& dir, int depth) {
  if (depth > 8) throw Error("option files are nested too deeply (param= loop?)");
  std::istringstream in(text);
  std::string line, logical;
  auto flush = [&]() {
    std::string l = logical;
    logical.clear();
    size_t h = l.find('#');
    if (h != std::string::npos) l = l.substr(0, h);
    l = trim(l);
    if (l.empty()) return;
    size_t eq = l.find('=');
    if (eq == std::string::npos) return; // Bio++ ignores lines without a delimiter
    std::string key = trim(l.substr(0, eq)), val = trim(l.substr(eq + 1));
    if (key.empty()) return;
    if (key == "param" || key == "params") {
      std::string path = (val.size() && val[0] != '/') ? dir + "/" + val : val;
      parse_option_text(read_file(path), out, dir_of(path), depth + 1);
      return;
    }
    out[key] = val;
  };
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    std::string t = line;
    size_t e = t.find_last_not_of(" \t");
    if (e != std::string::npos && t[e] == '\\') {
      logical += t.substr(0, e);
      continue;
    }
    logical += t;
    flush();
  }
  if (!logical.empty()) flush();
}

static void resolve_variables(Params& p) {
  for (int pass = 0; pass < 10; pass++) {
    bool changed = false;
    for (auto& kv : p) {
      std::string& v = kv.second;
      size_t a = v.find("$(");
      while (a != std::string::npos) {
        size_t b = v.find(')', a);
        if (b == std::string::npos) break;
        std::string var = v.substr(a + 2, b - a - 2);
        auto it = p.find(var);
        if (it == p.end()) throw Error("option variable $(" + var + ") is not defined");
        v = v.substr(0, a) + it->second + v.substr(b + 1);
        changed = true;
        a = v.find("$(", a + it->second.size());
      }
    }
    if (!changed) return;
  }
}

Application parse_command_line(int argc, const char* const* argv) {
  Application app;
  Params cmd;
  std::string param_file;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    if (a.rfind("--", 0) == 0) {
      if (a.rfind("--seed=", 0) == 0) {
        app.seed = std::strtoull(a.c_str() + 7, nullptr, 10);
        app.seed_given = true;
      }
      continue; // --noninteractive, --warning=N: accepted, nothing to do
    }
    size_t eq = a.find('=');
    if (eq == std::string::npos) continue;
    std::string key = trim(a.substr(0, eq)), val = trim(a.substr(eq + 1));
    if (key == "param" || key == "params") param_file = val;
    else cmd[key] = val;
  }
  app.base_dir = ".";
  if (!param_file.empty()) {
    parse_option_text(read_file(param_file), app.params, dir_of(param_file));
    app.base_dir = dir_of(param_file);
  }
  for (auto& kv : cmd) app.params[kv.first] = kv.second; // the command line wins
  resolve_variables(app.params);
  return app;
}

std::string get_string(const Params& p, const std::string& key, const std::string& def) {
  auto it = p.find(key);
  return it == p.end() ? def : it->second;
}
bool get_bool(const Params& p, const std::string& key, bool def) {
  auto it = p.find(key);
  if (it == p.end() || it->second.empty()) return def;
  const std::string& v = it->second;
  return v == "true" || v == "TRUE" || v == "t" || v == "T" || v == "yes" || v == "YES" || v == "y" || v == "Y" ||
         v == "1";
}
double get_double(const Params& p, const std::string& key, double def) {
  auto it = p.find(key);
  if (it == p.end() || it->second.empty()) return def;
  char* end = nullptr;
  double v = std::strtod(it->second.c_str(), &end);
  if (end == it->second.c_str()) throw Error("option " + key + ": '" + it->second + "' is not a number");
  return v;
}
long get_int(const Params& p, const std::string& key, long def) {
  auto it = p.find(key);
  if (it == p.end() || it->second.empty()) return def;
  char* end = nullptr;
  long v = std::strtol(it->second.c_str(), &end, 10);
  if (end == it->second.c_str()) throw Error("option " + key + ": '" + it->second + "' is not an integer");
  return v;
}
std::string get_path(const Params& p, const std::string& key, const std::string& def) {
  std::string v = get_string(p, key, def);
  if (v.empty() || v == "None" || v == "NONE") return "none";
  return v;
}

Procedure parse_procedure(const std::string& desc) {
  Procedure pr;
  std::string d = trim(desc);
  size_t open = d.find('(');
  if (open == std::string::npos) {
    pr.name = d;
    return pr;
  }
  size_t close = d.rfind(')');
  if (close == std::string::npos || close < open) throw Error("unbalanced parentheses in '" + desc + "'");
  pr.name = trim(d.substr(0, open));
  std::string body = d.substr(open + 1, close - open - 1);
  // split on commas at nesting depth 0
  std::vector<std::string> parts;
  int depth = 0;
  std::string cur;
  for (char c : body) {
    if (c == '(') depth++;
    if (c == ')') depth--;
    if (c == ',' && depth == 0) {
      parts.push_back(cur);
      cur.clear();
    } else cur += c;
  }
  if (!trim(cur).empty()) parts.push_back(cur);
  for (auto& part : parts) {
    size_t eq = std::string::npos;
    int dd = 0;
    for (size_t i = 0; i < part.size(); i++) {
      if (part[i] == '(') dd++;
      if (part[i] == ')') dd--;
      if (part[i] == '=' && dd == 0) { eq = i; break; }
    }
    if (eq == std::string::npos) throw Error("argument '" + trim(part) + "' of '" + pr.name + "' has no value");
    pr.args[trim(part.substr(0, eq))] = trim(part.substr(eq + 1));
  }
  return pr;
}

} // namespace host
