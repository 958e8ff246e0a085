// Command-line grammar of the reference's programs, re-implemented for the
// drop-in CLIs (behaviour of hclust/src/smithlab_cpp/OptionParser.cpp:156-184,
// 272-347, 381-448):
//   * an option is named `-x` (short), `-long` (single dash) or bare `long`;
//   * its value is the next token; bool options take none and toggle;
//   * `-config-file <f>` reads `name = value` lines ('#' comments);
//   * `-help`, `-?`, `-about`; argc == 1 behaves like -help;
//   * the first required option that was not given is reported as
//     "required argument missing: [-x, -long]"; callers print help / about /
//     missing messages on stderr and exit with status 0, as the reference does
//     (motif_both_points.cpp:324-335).
#pragma once
#include <algorithm>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace hscli {

struct OptionError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

class Options {
 public:
  Options(std::string prog, std::string descr) : prog_(std::move(prog)), descr_(std::move(descr)) {
    add("help", '?', "print this help message", false, help_);
    add("about", '\0', "print about message", false, about_);
  }

  void add(const std::string &l, char s, const std::string &d, bool req, std::string &v) { push(l, s, d, req, kString, &v); }
  void add(const std::string &l, char s, const std::string &d, bool req, unsigned &v) { push(l, s, d, req, kUint, &v); }
  void add(const std::string &l, char s, const std::string &d, bool req, int &v) { push(l, s, d, req, kInt, &v); }
  void add(const std::string &l, char s, const std::string &d, bool req, double &v) { push(l, s, d, req, kDouble, &v); }
  void add(const std::string &l, char s, const std::string &d, bool req, bool &v) { push(l, s, d, req, kBool, &v); }

  // Consumes recognised options from argv[1..]; what is left goes to `rest`.
  void parse(int argc, const char **argv, std::vector<std::string> &rest) {
    rest.assign(argv + 1, argv + argc);
    for (size_t i = 0; i < rest.size();) {
      if (rest[i] != "-config-file") {
        ++i;
        continue;
      }
      if (i + 1 >= rest.size()) throw OptionError("-config-line requires config filename");
      std::vector<std::string> lines = read_config(rest[i + 1]);
      for (Opt &o : opts_) {
        for (size_t j = 0; j < lines.size();) {
          const size_t eq = lines[j].find('=');
          const std::string name = strip(lines[j].substr(0, eq));
          const std::string val = strip(lines[j].substr(lines[j].rfind('=') + 1));
          if (matches(o, name)) {
            assign(o, val);
            o.given = true;
            lines.erase(lines.begin() + j);
          } else {
            ++j;
          }
        }
      }
      rest.erase(rest.begin() + i, rest.begin() + i + 2);
    }
    for (Opt &o : opts_) {
      for (size_t i = 0; i < rest.size();) {
        if (!matches(o, rest[i])) {
          ++i;
          continue;
        }
        assign(o, i + 1 < rest.size() ? rest[i + 1] : std::string());
        o.given = true;
        rest.erase(rest.begin() + i);
        if (o.type != kBool && i < rest.size()) rest.erase(rest.begin() + i);
      }
      if (o.required && !o.given && missing_.empty()) missing_ = label(o);
    }
  }

  bool help_requested() const { return help_; }
  bool about_requested() const { return about_; }
  bool option_missing() const { return !missing_.empty(); }
  std::string option_missing_message() const { return "required argument missing: [" + missing_ + "]"; }

  std::string help_message() const {
    size_t width = 0;
    for (const Opt &o : opts_) width = std::max(width, label(o).size());
    std::ostringstream ss;
    ss << "Usage: " << prog_ << " [OPTIONS]\n\n";
    if (opts_.size() > 2) {
      ss << "Options:\n";
      for (size_t i = 2; i < opts_.size(); ++i) line(ss, opts_[i], width);
    }
    ss << "\nHelp options:\n";
    for (size_t i = 0; i < 2; ++i) line(ss, opts_[i], width);
    return ss.str();
  }

  std::string about_message() const {
    std::ostringstream ss;
    ss << "PROGRAM: " << prog_ << "\n";
    wrap(ss, descr_, 0);
    return ss.str();
  }

 private:
  enum Type { kString, kUint, kInt, kDouble, kBool };
  struct Opt {
    std::string long_name, descr;
    char short_name;
    bool required, given;
    Type type;
    void *target;
  };
  static constexpr size_t kMaxLine = 72;

  void push(const std::string &l, char s, const std::string &d, bool req, Type t, void *v) {
    opts_.push_back(Opt{l, d, s, req, false, t, v});
  }
  static bool matches(const Opt &o, const std::string &tok) {
    if (tok == o.long_name) return true;
    if (tok.size() > 1 && tok[0] == '-') {
      if (tok.substr(1) == o.long_name) return true;
      if (tok.size() == 2 && tok[1] == o.short_name) return true;
    }
    return false;
  }
  static std::string label(const Opt &o) {
    if (o.short_name != '\0') return std::string("-") + o.short_name + ", -" + o.long_name;
    return "    -" + o.long_name;
  }
  static void assign(Opt &o, const std::string &val) {
    std::istringstream ss(val);
    bool ok = true;
    switch (o.type) {
      case kString: *static_cast<std::string *>(o.target) = val; break;
      case kUint: ok = static_cast<bool>(ss >> *static_cast<unsigned *>(o.target)); break;
      case kInt: ok = static_cast<bool>(ss >> *static_cast<int *>(o.target)); break;
      case kDouble: ok = static_cast<bool>(ss >> *static_cast<double *>(o.target)); break;
      case kBool: {
        bool &b = *static_cast<bool *>(o.target);
        b = !b;
        if (val == "true" || val == "on") b = true;
        if (val == "false" || val == "off") b = false;
        break;
      }
    }
    if (!ok) throw OptionError("Invalid argument [" + val + "] to option [" + label(o) + "]");
  }
  static void wrap(std::ostringstream &ss, const std::string &text, size_t offset) {
    std::istringstream words(text);
    std::string w;
    size_t used = 0;
    bool first = true;
    while (words >> w) {
      if (!first && offset + used + w.size() >= kMaxLine) {
        ss << "\n";
        used = 0;
      }
      if (!first && used == 0) ss << std::string(offset, ' ');
      ss << w << " ";
      used += w.size();
      first = false;
    }
  }
  static void line(std::ostringstream &ss, const Opt &o, size_t width) {
    ss << "  " << std::left << std::setw((int)width) << label(o) << "  ";
    wrap(ss, o.descr, width + 4);
    ss << "\n";
  }
  static std::string strip(const std::string &s) {
    const size_t a = s.find_first_not_of(" \t");
    if (a == std::string::npos) return std::string();
    const size_t b = s.find_last_not_of(" \t");
    return s.substr(a, b - a + 1);
  }
  static std::vector<std::string> read_config(const std::string &path) {
    std::ifstream in(path.c_str());
    if (!in) throw OptionError("cannot open config file " + path);
    std::vector<std::string> out;
    std::string raw;
    size_t n = 0;
    while (std::getline(in, raw)) {
      ++n;
      if (!raw.empty() && raw.back() == '\r') raw.pop_back();
      if (raw.empty() || raw[0] == '#') continue;
      const std::string s = strip(raw);
      if (s.empty() || s.find('=') == std::string::npos)
        throw OptionError("Line " + std::to_string(n) + " malformed in config file " + path);
      out.push_back(s);
    }
    return out;
  }

  std::string prog_, descr_, missing_;
  std::vector<Opt> opts_;
  bool help_ = false, about_ = false;
};

inline std::string strip_path(const std::string &full) {
  const size_t p = full.find_last_of("/\\");
  return p == std::string::npos ? full : full.substr(p + 1);
}

}  // namespace hscli
