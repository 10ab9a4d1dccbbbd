// cfg.cpp -- the transit option surface: same option names, defaults, `key value` file grammar
// and prefix matching as the reference (transit/src/argum.c:112-320 option table and 372-739
// handling; pu/src/procopt.c:649-704 file lines, 281-398 file/argv interleaving), so that the
// configuration file BART's makecfg.makeTransit writes (code/makecfg.py:23-108) is accepted
// unchanged.  Only options that affect the forward-model path have an effect; the others are
// parsed and ignored, as listed in DESIGN.md.
#include "host.hpp"
#include <cstring>
#include <cstdlib>
#include <cstdarg>
#include <fstream>
#include <sys/stat.h>

namespace bart {

namespace {
enum Arg { NOARG, REQ };
struct OptDef { const char *name; char shortc; Arg arg; const char *def; };
// Table ORDER matters: abbreviated keys resolve to the first name they prefix (procopt.c:677).
const OptDef kTable[] = {
  {"version", 'V', NOARG, nullptr}, {"help", 'h', NOARG, nullptr}, {"quiet", 'q', NOARG, nullptr},
  {"verb", 'v', REQ, "2"}, {"config_file", 'c', REQ, nullptr},
  {"atm", 0, REQ, "NULL"}, {"linedb", 0, REQ, nullptr}, {"outtoomuch", 0, REQ, nullptr},
  {"outsample", 0, REQ, nullptr}, {"outspec", 0, REQ, "outspectrum"},
  {"outintens", 0, REQ, nullptr}, {"molfile", 0, REQ, "../inputs/molecules.dat"},
  {"savefiles", 0, REQ, nullptr},
  {"raddelt", 0, REQ, "-1"}, {"radlow", 0, REQ, "0"}, {"radhigh", 0, REQ, "0"},
  {"radfct", 0, REQ, "0"},
  {"allowq", 0, REQ, "0.00001"}, {"refpress", 0, REQ, nullptr}, {"refradius", 0, REQ, nullptr},
  {"gsurf", 0, REQ, nullptr}, {"qmol", 0, REQ, nullptr}, {"qscale", 0, REQ, nullptr},
  {"wllow", 0, REQ, nullptr}, {"wlhigh", 0, REQ, nullptr}, {"wlfct", 0, REQ, "1e-4"},
  {"wnlow", 0, REQ, nullptr}, {"wnhigh", 0, REQ, nullptr}, {"wndelt", 0, REQ, "0"},
  {"wnosamp", 0, REQ, "2160"}, {"wnfct", 0, REQ, "0"},
  {"ndop", 0, REQ, "60"}, {"nlor", 0, REQ, "60"}, {"dmin", 0, REQ, "1e-3"},
  {"dmax", 0, REQ, "0.25"}, {"lmin", 0, REQ, "1e-4"}, {"lmax", 0, REQ, "10.0"},
  {"nwidth", 'a', REQ, "20"},
  {"ethreshold", 0, REQ, "1e-8"}, {"cloud", 0, REQ, nullptr}, {"cloudtop", 0, REQ, nullptr},
  {"scattering", 0, REQ, nullptr}, {"detailext", 0, REQ, nullptr},
  {"detailcia", 0, REQ, nullptr}, {"csfile", 0, REQ, nullptr}, {"saveext", 0, REQ, nullptr},
  {"opacityfile", 0, REQ, nullptr}, {"tlow", 0, REQ, "500"}, {"thigh", 0, REQ, "3000"},
  {"tempdelt", 0, REQ, "100.0"}, {"justOpacity", 0, NOARG, nullptr},
  {"shareOpacity", 0, NOARG, nullptr},
  {"solution", 's', REQ, "eclipse"}, {"toomuch", 0, REQ, "20"}, {"taulevel", 0, REQ, "1"},
  {"modlevel", 0, REQ, "1"}, {"detailtau", 0, REQ, nullptr},
  {"starrad", 0, REQ, "1.125"}, {"gorbpar", 0, REQ, nullptr}, {"gorbparfct", 0, REQ, nullptr},
  {"transparent", 0, NOARG, nullptr}, {"raygrid", 0, REQ, "0 20 40 60 80"},
};
const int kNopt = sizeof(kTable) / sizeof(kTable[0]);

std::string rstrip(const std::string &s) {
  size_t e = s.size();
  while (e > 0 && (s[e - 1] == ' ' || s[e - 1] == '\t' || s[e - 1] == '\r' || s[e - 1] == '\n')) e--;
  return s.substr(0, e);
}

void apply(Options &o, int idx, const std::string &val, std::vector<std::string> &pending_files) {
  const std::string n = kTable[idx].name;
  const char *v = val.c_str();
  if (n == "version") { printf("This is 'transit' (bart_b200) version 4.0\n\n"); exit(EXIT_SUCCESS); }
  else if (n == "help") { printf("bart_b200 transit: options follow transit/src/argum.c\n"); exit(EXIT_SUCCESS); }
  else if (n == "quiet") o.verb = 1;
  else if (n == "verb") o.verb = (int)strtol(v, nullptr, 10);
  else if (n == "config_file") pending_files.push_back(val);
  else if (n == "atm") o.atm = val;
  else if (n == "linedb") o.linedb = val;
  else if (n == "outtoomuch") o.outtoomuch = val;
  else if (n == "outsample") o.outsample = val;
  else if (n == "outspec") o.outspec = val;
  else if (n == "outintens") o.outintens = val;
  else if (n == "molfile") o.molfile = val;
  else if (n == "savefiles") {
    if (strncmp(v, "yes", 3) == 0) o.savefiles = true;
    else if (strncmp(v, "no", 2) == 0) o.savefiles = false;
    else fail("Allowed arguments for savefiles are: 'yes' or 'no'");
  }
  else if (n == "raddelt") o.raddelt = atof(v);
  else if (n == "radlow") o.radlow = atof(v);
  else if (n == "radhigh") o.radhigh = atof(v);
  else if (n == "radfct") o.radfct = atof(v);
  else if (n == "allowq") o.allowq = (float)atof(v);
  else if (n == "refpress") o.refpress = atof(v);
  else if (n == "refradius") o.refradius = atof(v);
  else if (n == "gsurf") o.gsurf = atof(v);
  else if (n == "qmol") o.qmol = val;
  else if (n == "qscale") o.qscale = val;
  else if (n == "wllow") o.wllow = atof(v);
  else if (n == "wlhigh") o.wlhigh = atof(v);
  else if (n == "wlfct") o.wlfct = atof(v);
  else if (n == "wnlow") o.wnlow = atof(v);
  else if (n == "wnhigh") o.wnhigh = atof(v);
  else if (n == "wndelt") o.wndelt = atof(v);
  else if (n == "wnosamp") o.wnosamp = (int)atof(v);
  else if (n == "wnfct") o.wnfct = atof(v);
  else if (n == "ndop") o.ndop = atoi(v);
  else if (n == "nlor") o.nlor = atoi(v);
  else if (n == "dmin") o.dmin = (float)atof(v);
  else if (n == "dmax") o.dmax = (float)atof(v);
  else if (n == "lmin") o.lmin = (float)atof(v);
  else if (n == "lmax") o.lmax = (float)atof(v);
  else if (n == "nwidth") o.nwidth = (float)atof(v);
  else if (n == "ethreshold") o.ethreshold = atof(v);
  else if (n == "cloud") {
    // cloudtype,cloudext,cloudtop,cloudbot[,...] (argum.c:637-711).  Only the constant-extinction
    // type is supported: the other types read an uninitialised array in the reference
    // (tau.c:127-131,203) and have no defined result to reproduce.
    if (val.compare(0, 3, "ext") != 0)
      fail("--cloud: only the 'ext' (constant extinction) cloud type is supported");
    double a[3] = {0, 0, 0};
    const char *p = v + 3;
    for (int k = 0; k < 3; k++) {
      if (*p != ',' || p[1] == '\0')
        fail("Syntax error in option '--cloud', parameters need to be given as "
             "cloudtype,cloudext,cloudtop,cloudbot.");
      char *e; a[k] = strtod(p + 1, &e); p = e;
    }
    o.cloud_flag = 1; o.cloudext = a[0]; o.cloudtop = a[1]; o.cloudbot = a[2];
    if (o.cloudtop > o.cloudbot)
      fail("Syntax error in '--cloud', the cloud top (%g) needs to be less than the cloud "
           "bottom (%g).", o.cloudtop, o.cloudbot);
  }
  else if (n == "cloudtop") {                       // argum.c:713-719
    o.cloudtop = atof(v); o.cloudbot = o.cloudtop + 10; o.cloudext = 100.0; o.cloud_flag = 1;
  }
  else if (n == "scattering") {                     // argum.c:721-735
    if (val == "polar") { o.scat_logext = 0.0; o.scat_flag = 2; }
    else { o.scat_logext = atof(v); o.scat_flag = 1; }
  }
  else if (n == "csfile") {
    o.csfiles.clear();
    size_t s = 0;
    while (true) {
      size_t c = val.find(',', s);
      std::string f = rstrip(val.substr(s, c == std::string::npos ? c : c - s));
      size_t b = f.find_first_not_of(" \t");
      if (b != std::string::npos) o.csfiles.push_back(f.substr(b));
      if (c == std::string::npos) break;
      s = c + 1;
    }
  }
  else if (n == "saveext") o.saveext = val;
  else if (n == "opacityfile") o.opacityfile = val;
  else if (n == "tlow") o.tlow = atof(v);
  else if (n == "thigh") o.thigh = atof(v);
  else if (n == "tempdelt") o.tempdelt = atof(v);
  else if (n == "justOpacity") o.justOpacity = true;
  else if (n == "shareOpacity") o.shareOpacity = true;
  else if (n == "solution") o.solution = val;
  else if (n == "toomuch") o.toomuch = atof(v);
  else if (n == "taulevel") o.taulevel = atoi(v);
  else if (n == "modlevel") o.modlevel = atoi(v);
  else if (n == "starrad") o.starrad = atof(v);
  else if (n == "transparent") o.transparent = true;
  else if (n == "raygrid") o.raygrid = val;
  // detailext/detailcia/detailtau/gorbpar/gorbparfct: accepted, no effect on this path
}

int find_by_prefix(const char *key, size_t len) {
  for (int i = 0; i < kNopt; i++)
    if (strncmp(kTable[i].name, key, len) == 0) return i;
  return -1;
}

void process_file(const std::string &path, Options &o, bool must_exist, int depth);

void drain(std::vector<std::string> &pending, Options &o, int depth) {
  std::vector<std::string> files;
  files.swap(pending);
  for (auto &f : files) process_file(f, o, true, depth + 1);
}

void process_file(const std::string &path, Options &o, bool must_exist, int depth) {
  if (depth > 8) fail("config files nested too deeply at '%s'", path.c_str());
  std::ifstream in(path);
  if (!in) {
    if (must_exist) fail("Unable to succesfully open parameter file '%s'", path.c_str());
    return;
  }
  std::string line;
  std::vector<std::string> pending;
  while (std::getline(in, line)) {
    if (line.empty() || line[0] == '#') continue;            // procopt.c:357-358
    size_t k = 0;
    while (k < line.size() && line[k] != ' ' && line[k] != '\t') k++;
    size_t vpos = k;
    while (vpos < line.size() && (line[vpos] == ' ' || line[vpos] == '\t')) vpos++;
    if (k == 0) continue;                                    // line of blanks
    int idx = find_by_prefix(line.c_str(), k);
    std::string val = rstrip(line.substr(vpos));
    if (idx < 0 || (kTable[idx].arg == REQ && val.empty()))
      fail("Unknown, unsupported, or missing parameter to option '%s' in '%s', use '-h' to "
           "see the available options.", line.substr(0, k).c_str(), path.c_str());
    apply(o, idx, val, pending);
    if (!pending.empty()) drain(pending, o, depth);
  }
}
}  // namespace

void parse_options(int argc, char **argv, Options &o) {
  std::vector<std::string> pending;
  // 1. defaults, in table order (procopt.c:87-170: every non-NULL default is fed through the
  //    same switch as a user value)
  for (int i = 0; i < kNopt; i++)
    if (kTable[i].def && std::string(kTable[i].name) != "atm") apply(o, i, kTable[i].def, pending);
  // 2. ./.transitrc when present (argum.c:36, procopt.c:333-347)
  struct stat st;
  if (stat("./.transitrc", &st) == 0) process_file("./.transitrc", o, false, 0);
  // 3. command line, getopt_long semantics (long options may be abbreviated; -c reads a file
  //    at that point)
  for (int i = 1; i < argc; i++) {
    const char *a = argv[i];
    if (!a) break;
    int idx = -1;
    std::string val;
    bool have_val = false;
    if (a[0] == '-' && a[1] == '-' && a[2]) {
      const char *eq = strchr(a + 2, '=');
      size_t len = eq ? (size_t)(eq - (a + 2)) : strlen(a + 2);
      // exact match first, then unique/first prefix
      for (int k = 0; k < kNopt; k++)
        if (strlen(kTable[k].name) == len && strncmp(kTable[k].name, a + 2, len) == 0) idx = k;
      if (idx < 0) idx = find_by_prefix(a + 2, len);
      if (eq) { val = eq + 1; have_val = true; }
    } else if (a[0] == '-' && a[1]) {
      for (int k = 0; k < kNopt; k++)
        if (kTable[k].shortc && kTable[k].shortc == a[1]) idx = k;
      if (idx >= 0 && a[2]) { val = a + 2; have_val = true; }
    } else {
      continue;                                              // non-option argument: ignored
    }
    if (idx < 0)
      fail("Unknown, unsupported, or missing parameter to option '%s' passed as argument, "
           "use '-h' to see the available options.", a);
    if (kTable[idx].arg == REQ && !have_val) {
      if (i + 1 >= argc || !argv[i + 1])
        fail("Missing parameter to option '%s'.", a);
      val = argv[++i];
    }
    apply(o, idx, rstrip(val), pending);
    if (!pending.empty()) drain(pending, o, 0);
  }
  g_verb = o.verb;
}

}  // namespace bart
