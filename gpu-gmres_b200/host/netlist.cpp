// netlist.cpp -- SPICE-subset power-grid netlist -> MNA matrices G, C, B and source tables.
// Pure host C++ behind the bg_netlist_* entry points; what it produces is exactly what bg_matrix_create_csr /
// bg_transient_create consume, so a netlist runs through the device solver without further host code.
//
// It plays the role of parser_sub + stampG / stampC / stampB behind parser_wrapper in the reference
// (src/etbr_wrapper.cpp:46-171, src/parser.cpp:273-636, 1904-2886) and keeps that front end's conventions, because
// they decide what every row and column of the system means:
//   * nodes are numbered in order of first appearance on element lines (R, C, L, V, I), "0" and "gnd" are ground
//     (src/element.cpp:90-129); the unknown vector is [node voltages | inductor currents | voltage-source currents]
//     (src/etbr_wrapper.cpp:112);
//   * stamps (src/parser.cpp:1981-2049, 2176-2210, 2352-2364, 2528-2541):
//       R a b r   : G(a,a) += 1/r, G(b,b) += 1/r, G(a,b) -= 1/r, G(b,a) -= 1/r         C a b c : the same into C
//       L a b l   : row q = nNodes + k:  G(q,a) = -1, G(a,q) = +1, G(q,b) = +1, G(b,q) = -1, C(q,q) = l
//       V a b ... : row q = nNodes + nL + k:  G(a,q) = +1, G(q,a) = -1, G(b,q) = -1, G(q,b) = +1, B(q,k) = -1
//       I a b ... : column nVS + k of B:  B(a,.) = -1, B(b,.) = +1
//     so that G x + C dx/dt = B u(t); entries pushed more than once add up in file order (src/matrix.cpp:86-125);
//   * numbers take the scale letters t g k m(eg) u n p f (StrToNum, src/parser.cpp:30-66);
//   * sources: a plain number is DC, PWL(t v t v ...) possibly continued on "+ t v" lines gets a point at t = 0
//     holding the first value when its first time is not 0 (src/parser.cpp:2392-2402), PULSE(v1 v2 td tr tf pw per)
//     is kept as its seven numbers (what gen_PULSEut_kernel consumes) and also expanded to the six breakpoints per
//     period the reference stores (src/parser.cpp:2497-2518); ".tran tstep tstop", ".print tran v(node) ..." name
//     the time axis and the observed ports (src/parser.cpp:514-546); ".include" is followed.
// Deliberate differences: one pass instead of re-reading the file four times; lines of any length; a malformed
// element is an error with its line number instead of a printf; the reference's read loop handles the last line of a
// file twice when it is not ".end" (feof-before-fgets, src/parser.cpp:339-350) -- not reproduced; sub-circuits
// (.subckt / X, stamp_sub) are refused as BG_ERR_UNSUPPORTED.
//
// BG_NETLIST_REDUCED (not in the reference): nodes held by a voltage source to ground are eliminated together with
// that source's current -- the rows whose zero diagonal forces the reference onto a pivoting ILU (ILU++ ILUCDP) --
// their known voltage moves into B.  What is left is the diagonally dominant conductance system the ILU(0) paths
// of this library are built for; bg_netlist_get_reduction maps solutions back.
#include <algorithm>
#include <cctype>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <cstdint>
#include <vector>

#include "../../include/b200gmres.h"

namespace bg { void set_error(const char* fmt, ...); }

namespace {
constexpr int kGround = -1;

struct Triplets {  // in push order; assembled by summing duplicates in that order
  std::vector<int> r, c;
  std::vector<double> v;
  void push(int i, int j, double x) { r.push_back(i); c.push_back(j); v.push_back(x); }
};
struct Csr {
  int n_rows = 0, n_cols = 0;
  std::vector<int> rp, ci;
  std::vector<double> v;
};
struct Source {
  int kind = BG_SRC_DC;
  bool is_v = false;
  int a = kGround, b = kGround;
  std::vector<double> t, v;  // the table the reference stores (DC: two points, PULSE: six per period)
  double pulse[7] = {0, 0, 0, 0, 0, 0, 0};  // v1 v2 td tr tf pw period
  int line = 0;
};
struct Branch { int a, b; double value; int line; };

// node name -> node number: open addressing over 64-bit hashes, the names themselves live in node_name (a netlist
// of N nodes costs 3-7 lookups per line; std::unordered_map spent more than half of the parse time here)
struct NameTable {
  struct Slot { uint64_t hash; int id; int pad; };  // hash 0 = empty; one cache line miss per probe
  std::vector<Slot> slot;
  size_t used = 0;
  static uint64_t hash_of(const char* p, size_t n) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t k = 0; k < n; ++k) h = (h ^ (unsigned char)p[k]) * 0x100000001b3ull;
    h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 32;
    return h ? h : 1;
  }
  void grow() {
    const size_t cap = slot.empty() ? 1024 : slot.size() * 2;
    std::vector<Slot> s2(cap, Slot{0, 0, 0});
    for (const Slot& s : slot)
      if (s.hash) {
        size_t q = s.hash & (cap - 1);
        while (s2[q].hash) q = (q + 1) & (cap - 1);
        s2[q] = s;
      }
    slot.swap(s2);
  }
  // number of `name`, or -1 when absent; `add` >= 0 inserts it with that number
  int find(const std::vector<std::string>& names, const char* p, size_t n, int add = -1) {
    if (used * 2 >= slot.size()) grow();
    const uint64_t h = hash_of(p, n);
    const size_t mask = slot.size() - 1;
    size_t q = h & mask;
    for (; slot[q].hash; q = (q + 1) & mask)
      if (slot[q].hash == h) {
        const std::string& nm = names[(size_t)slot[q].id];
        if (nm.size() == n && std::memcmp(nm.data(), p, n) == 0) return slot[q].id;
      }
    if (add >= 0) { slot[q] = Slot{h, add, 0}; ++used; }
    return -1;
  }
};
inline bool is_ground(const char* p, size_t n) { return (n == 1 && p[0] == '0') || (n == 3 && p[0] == 'g' && p[1] == 'n' && p[2] == 'd'); }
}  // namespace

struct bg_netlist_s {
  int formulation = BG_NETLIST_MNA;
  std::vector<std::string> node_name;  // by node number
  NameTable node_of;  // name -> number (ground is recognised by name)
  int lookup(const char* p, size_t n) {  // -1 ground, -2 unknown
    if (is_ground(p, n)) return kGround;
    const int id = node_of.find(node_name, p, n);
    return id < 0 ? -2 : id;
  }
  std::vector<Branch> res, cap, ind;
  std::vector<Source> vs, is;
  std::vector<std::string> port_name;
  double tstep = 0, tstop = 0;
  int64_t n_lines = 0;
  // assembled, active formulation
  Csr G, C, B;
  std::vector<int> port_row;
  // sources in column order of B
  std::vector<int> kind, pwl_ptr, pulse_src;
  std::vector<double> pwl_t, pwl_v, pulse_time, pulse_val;
  // reduction
  std::vector<int> kept, new_of_mna, elim_row, elim_src;
  std::vector<double> elim_sign;
  int n_nodes() const { return (int)node_name.size(); }
  int size_mna() const { return n_nodes() + (int)ind.size() + (int)vs.size(); }
};

namespace {
// StrToNum: the scale letter follows the number directly (src/parser.cpp:30-66).  Tokens point into the NUL-terminated
// text and end at a separator, which also stops strtod.
struct Tok {
  const char* p;
  size_t n;
  std::string str() const { return std::string(p, n); }
};
bool parse_number(const Tok& tk, double& out) {
  char* end = nullptr;
  const double x = std::strtod(tk.p, &end);
  if (end == tk.p || end > tk.p + tk.n) return false;
  const char c0 = end < tk.p + tk.n ? end[0] : '\0', c1 = end + 1 < tk.p + tk.n ? end[1] : '\0';
  int e = 0;
  switch (c0) {
    case 'T': case 't': e = 12; break;
    case 'G': case 'g': e = 9; break;
    case 'K': case 'k': e = 3; break;
    case 'M': case 'm': e = (c1 == 'E' || c1 == 'e') ? 6 : -3; break;
    case 'U': case 'u': e = -6; break;
    case 'N': case 'n': e = -9; break;
    case 'P': case 'p': e = -12; break;
    case 'F': case 'f': e = -15; break;
    default: break;
  }
  out = e ? x * std::pow(10.0, e) : x;  // value * pow(10.0, e) like the reference, so the doubles agree
  return true;
}
bool ieq(const char* p, size_t n, const char* lit) {  // case-insensitive prefix match of the literal
  size_t k = 0;
  for (; lit[k]; ++k)
    if (k >= n || std::tolower((unsigned char)p[k]) != lit[k]) return false;
  return true;
}
bool ieq(const Tok& t, const char* lit) { return ieq(t.p, t.n, lit); }
void tokens_of(const char* p, size_t n, bool split_punct, std::vector<Tok>& out) {
  out.clear();
  size_t k = 0;
  while (k < n) {
    auto sep = [&](char ch) { return ch == ' ' || ch == '\t' || ch == '\r' || ch == '\n' || (split_punct && (ch == '(' || ch == ')' || ch == ',' || ch == '=')); };
    while (k < n && sep(p[k])) ++k;
    const size_t b = k;
    while (k < n && !sep(p[k])) ++k;
    if (k > b) out.push_back({p + b, k - b});
  }
}

struct Parser {
  bg_netlist_s& N;
  std::string err;
  bool open_pwl = false;  // "+ t v" continuation lines extend the last source (of the list open_is_v names)
  bool open_is_v = false;
  int depth = 0;
  std::vector<Tok> tk;
  explicit Parser(bg_netlist_s& n) : N(n) {}

  int node(const Tok& t) {
    if (is_ground(t.p, t.n)) return kGround;
    const int id = N.node_of.find(N.node_name, t.p, t.n, N.n_nodes());
    if (id >= 0) return id;
    N.node_name.emplace_back(t.p, t.n);
    return N.n_nodes() - 1;
  }
  bool fail(int line, const char* what, const std::string& text) {
    char buf[512];
    std::snprintf(buf, sizeof buf, "netlist line %d: %s: %.200s", line, what, text.c_str());
    err = buf;
    return false;
  }
  void pwl_point(Source& s, double t, double v) {
    if (s.t.empty() && t != 0) { s.t.push_back(0.0); s.v.push_back(v); }  // src/parser.cpp:2392-2402
    s.t.push_back(t);
    s.v.push_back(v);
  }
  bool source_line(const char* lp, size_t ll, int ln, bool is_v) {
    tokens_of(lp, ll, true, tk);
    if (tk.size() < 4) return fail(ln, "source needs two nodes and a value", std::string(lp, ll));
    Source s;
    s.is_v = is_v;
    s.line = ln;
    s.a = node(tk[1]);
    s.b = node(tk[2]);
    size_t k = 3;
    // the waveform keyword may follow a DC value (the reference's PULSE lines carry one, src/parser.cpp:2456-2460)
    size_t kw = k;
    while (kw < tk.size() && !ieq(tk[kw], "pwl") && !ieq(tk[kw], "pu")) ++kw;
    if (kw < tk.size() && ieq(tk[kw], "pwl")) {
      s.kind = BG_SRC_PWL;
      if ((tk.size() - kw - 1) % 2) return fail(ln, "PWL needs time/value pairs", std::string(lp, ll));
      for (size_t q = kw + 1; q < tk.size(); q += 2) {
        double t, v;
        if (!parse_number(tk[q], t) || !parse_number(tk[q + 1], v)) return fail(ln, "bad number in PWL", std::string(lp, ll));
        pwl_point(s, t, v);
      }
    } else if (kw < tk.size()) {
      s.kind = BG_SRC_PULSE;
      if (tk.size() - kw - 1 != 7) return fail(ln, "PULSE needs v1 v2 td tr tf pw period", std::string(lp, ll));
      for (int q = 0; q < 7; ++q)
        if (!parse_number(tk[kw + 1 + q], s.pulse[q])) return fail(ln, "bad number in PULSE", std::string(lp, ll));
      if (!(s.pulse[6] > 0)) return fail(ln, "PULSE period must be positive", std::string(lp, ll));
    } else {
      if (ieq(tk[k], "dc") && tk[k].n == 2) ++k;
      double x;
      if (k >= tk.size() || !parse_number(tk[k], x)) return fail(ln, "bad source value", std::string(lp, ll));
      s.kind = BG_SRC_DC;
      s.pulse[0] = x;
    }
    open_pwl = s.kind == BG_SRC_PWL;
    open_is_v = is_v;
    (is_v ? N.vs : N.is).push_back(std::move(s));
    return true;
  }
  bool branch_line(const char* lp, size_t ll, int ln, std::vector<Branch>& dst, const char* what) {
    tokens_of(lp, ll, false, tk);
    double x;
    if (tk.size() < 4 || !parse_number(tk[3], x)) return fail(ln, what, std::string(lp, ll));
    if (!(x > 0) || !std::isfinite(x)) return fail(ln, "element value must be positive and finite", std::string(lp, ll));
    const int a = node(tk[1]), b = node(tk[2]);
    dst.push_back({a, b, x, ln});
    open_pwl = false;
    return true;
  }
  // `p` must be NUL-terminated at p[len]
  bool text(const char* p, size_t len, const std::string& dir) {
    if (++depth > 16) { err = "netlist: .include nested deeper than 16"; return false; }
    size_t pos = 0;
    int ln = 0;
    while (pos < len) {
      size_t e = pos;
      while (e < len && p[e] != '\n') ++e;
      const char* lp = p + pos;
      size_t ll = e - pos;
      pos = e + 1;
      ++ln;
      ++N.n_lines;
      while (ll && (*lp == ' ' || *lp == '\t')) { ++lp; --ll; }
      if (!ll) continue;
      switch (*lp) {
        case 'R': case 'r': if (!branch_line(lp, ll, ln, N.res, "resistor needs two nodes and a value")) return false; break;
        case 'C': case 'c': if (!branch_line(lp, ll, ln, N.cap, "capacitor needs two nodes and a value")) return false; break;
        case 'L': case 'l': if (!branch_line(lp, ll, ln, N.ind, "inductor needs two nodes and a value")) return false; break;
        case 'V': case 'v': if (!source_line(lp, ll, ln, true)) return false; break;
        case 'I': case 'i': if (!source_line(lp, ll, ln, false)) return false; break;
        case 'X': case 'x': return fail(ln, "sub-circuit instances are not supported", std::string(lp, ll));
        case '+': {
          if (!open_pwl) break;  // continuation of something this subset ignores
          tokens_of(lp + 1, ll - 1, true, tk);
          if (tk.size() % 2) return fail(ln, "PWL continuation needs time/value pairs", std::string(lp, ll));
          Source& s = open_is_v ? N.vs.back() : N.is.back();
          for (size_t q = 0; q < tk.size(); q += 2) {
            double t, v;
            if (!parse_number(tk[q], t) || !parse_number(tk[q + 1], v)) return fail(ln, "bad number in PWL continuation", std::string(lp, ll));
            pwl_point(s, t, v);
          }
          break;
        }
        case '.': {
          const std::string line(lp, ll);
          if (ieq(lp, ll, ".tran")) {
            tokens_of(lp, ll, false, tk);
            if (tk.size() < 3 || !parse_number(tk[1], N.tstep) || !parse_number(tk[2], N.tstop)) return fail(ln, ".tran needs tstep tstop", line);
          } else if (ieq(lp, ll, ".print")) {
            for (size_t q = 0; q < line.size(); ++q)
              if (line[q] == '(') {
                const size_t r = line.find(')', q);
                if (r == std::string::npos) return fail(ln, "unterminated v( in .print", line);
                std::string nm = line.substr(q + 1, r - q - 1);
                while (!nm.empty() && nm.back() == ' ') nm.pop_back();
                while (!nm.empty() && nm.front() == ' ') nm.erase(0, 1);
                N.port_name.push_back(nm);
                q = r;
              }
          } else if (ieq(lp, ll, ".include")) {
            tokens_of(lp, ll, false, tk);
            if (tk.size() < 2) return fail(ln, ".include needs a file name", line);
            std::string name;
            for (size_t q = 0; q < tk[1].n; ++q) if (tk[1].p[q] != '"' && tk[1].p[q] != '\'') name.push_back(tk[1].p[q]);
            const std::string path = (!name.empty() && name[0] == '/') ? name : dir + name;
            std::string sub;
            if (!slurp(path, sub)) return fail(ln, "cannot open included file", path);
            const size_t slash = path.find_last_of('/');
            if (!text(sub.c_str(), sub.size(), slash == std::string::npos ? std::string() : path.substr(0, slash + 1))) return false;
          } else if (ieq(lp, ll, ".subckt")) {
            return fail(ln, "sub-circuit definitions are not supported", line);
          } else if (ieq(lp, ll, ".end") && !ieq(lp, ll, ".ends")) {
            --depth;
            return true;
          }
          break;
        }
        default: break;  // '*' comments and everything this subset does not know
      }
    }
    --depth;
    return true;
  }
  static bool slurp(const std::string& path, std::string& out) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    if (std::fseek(f, 0, SEEK_END) == 0) {
      const long sz = std::ftell(f);
      if (sz > 0) out.reserve((size_t)sz + 1);
      std::rewind(f);
    }
    char buf[1 << 16];
    size_t got;
    while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, got);
    std::fclose(f);
    return true;
  }
};

// triplets -> CSR with sorted columns; duplicates add up in push order (src/matrix.cpp:86-125)
void assemble(const Triplets& T, int n_rows, int n_cols, Csr& M) {
  const size_t nz = T.r.size();
  std::vector<size_t> order(nz);
  for (size_t k = 0; k < nz; ++k) order[k] = k;
  std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return T.r[x] != T.r[y] ? T.r[x] < T.r[y] : T.c[x] < T.c[y]; });
  M.n_rows = n_rows;
  M.n_cols = n_cols;
  M.rp.assign((size_t)n_rows + 1, 0);
  M.ci.clear();
  M.v.clear();
  for (size_t q = 0; q < nz;) {
    const size_t k = order[q];
    double s = T.v[k];
    size_t e = q + 1;
    while (e < nz && T.r[order[e]] == T.r[k] && T.c[order[e]] == T.c[k]) s += T.v[order[e++]];
    M.ci.push_back(T.c[k]);
    M.v.push_back(s);
    M.rp[(size_t)T.r[k] + 1]++;
    q = e;
  }
  for (int i = 0; i < n_rows; ++i) M.rp[(size_t)i + 1] += M.rp[i];
}

// the table the reference keeps per source (src/parser.cpp:2497-2541)
void reference_table(Source& s, double tstop) {
  if (s.kind == BG_SRC_DC) {
    s.t = {0.0, tstop};
    s.v = {s.pulse[0], s.pulse[0]};
  } else if (s.kind == BG_SRC_PULSE) {
    const double v1 = s.pulse[0], v2 = s.pulse[1], td = s.pulse[2], tr = s.pulse[3], tf = s.pulse[4], pw = s.pulse[5], per = s.pulse[6];
    const int np = (int)std::floor(tstop / per) + 1;
    s.t.assign((size_t)6 * np, 0.0);
    s.v.assign((size_t)6 * np, 0.0);
    for (int i = 0; i < np; ++i) {
      double* t = &s.t[(size_t)6 * i];
      double* v = &s.v[(size_t)6 * i];
      t[0] = i * per;                     v[0] = v1;
      t[1] = i * per + td;                v[1] = v1;
      t[2] = i * per + td + tr;           v[2] = v2;
      t[3] = i * per + td + tr + pw;      v[3] = v2;
      t[4] = i * per + td + tr + pw + tf; v[4] = v1;
      t[5] = (i + 1) * per;               v[5] = v1;
    }
    if (np * per > tstop) s.t[(size_t)6 * np - 1] = tstop;
  }
}

int build(bg_netlist_s& N) {
  const int nn = N.n_nodes(), nl = (int)N.ind.size(), nv = (int)N.vs.size(), ni = (int)N.is.size();
  const int64_t size64 = (int64_t)nn + nl + nv;
  if (size64 > 0x7ffffff0LL) { bg::set_error("netlist: more than 2^31 unknowns"); return BG_ERR_UNSUPPORTED; }
  const int size = (int)size64, nsrc = nv + ni;
  Triplets g, c, b;
  for (const Branch& e : N.res) {
    const double y = 1.0 / e.value;
    if (e.a != kGround) g.push(e.a, e.a, y);
    if (e.b != kGround) g.push(e.b, e.b, y);
    if (e.a != kGround && e.b != kGround) { g.push(e.a, e.b, -y); g.push(e.b, e.a, -y); }
  }
  for (int k = 0; k < nl; ++k) {
    const Branch& e = N.ind[k];
    const int q = nn + k;
    if (e.a != kGround) { g.push(q, e.a, -1.0); g.push(e.a, q, 1.0); }
    if (e.b != kGround) { g.push(q, e.b, 1.0); g.push(e.b, q, -1.0); }
    c.push(q, q, e.value);
  }
  for (int k = 0; k < nv; ++k) {
    const Source& s = N.vs[k];
    const int q = nn + nl + k;
    if (s.a != kGround) { g.push(s.a, q, 1.0); g.push(q, s.a, -1.0); }
    if (s.b != kGround) { g.push(s.b, q, -1.0); g.push(q, s.b, 1.0); }
    b.push(q, k, -1.0);
  }
  for (const Branch& e : N.cap) {
    if (e.a != kGround) c.push(e.a, e.a, e.value);
    if (e.b != kGround) c.push(e.b, e.b, e.value);
    if (e.a != kGround && e.b != kGround) { c.push(e.a, e.b, -e.value); c.push(e.b, e.a, -e.value); }
  }
  for (int k = 0; k < ni; ++k) {
    const Source& s = N.is[k];
    if (s.a != kGround) b.push(s.a, nv + k, -1.0);
    if (s.b != kGround) b.push(s.b, nv + k, 1.0);
  }

  // sources in column order of B: voltage sources, then current sources
  N.kind.clear(); N.pwl_ptr.assign(1, 0); N.pwl_t.clear(); N.pwl_v.clear(); N.pulse_src.clear(); N.pulse_time.clear(); N.pulse_val.clear();
  for (int k = 0; k < nsrc; ++k) {
    Source& s = k < nv ? N.vs[k] : N.is[k - nv];
    reference_table(s, N.tstop);
    N.kind.push_back(s.kind);
    N.pwl_t.insert(N.pwl_t.end(), s.t.begin(), s.t.end());
    N.pwl_v.insert(N.pwl_v.end(), s.v.begin(), s.v.end());
    N.pwl_ptr.push_back((int)N.pwl_t.size());
    if (s.kind == BG_SRC_PULSE) {  // (td, tr, tf, tw, tp), (vlo, vhi): the record of bg_transient_set_pulse
      N.pulse_src.push_back(k);
      const double rec[5] = {s.pulse[2], s.pulse[3], s.pulse[4], s.pulse[5], s.pulse[6]};
      N.pulse_time.insert(N.pulse_time.end(), rec, rec + 5);
      N.pulse_val.push_back(s.pulse[0]);
      N.pulse_val.push_back(s.pulse[1]);
    }
  }

  N.kept.clear(); N.elim_row.clear(); N.elim_src.clear(); N.elim_sign.clear();
  N.new_of_mna.assign(size, 0);
  if (N.formulation == BG_NETLIST_REDUCED) {
    std::vector<int> fixed_by(size, -1);  // node row -> source, branch row -> source
    std::vector<double> sign_of(size, 0.0);
    for (int k = 0; k < nv; ++k) {
      const Source& s = N.vs[k];
      if ((s.a == kGround) == (s.b == kGround)) continue;  // floating (or shorted) sources keep their MNA row
      const int node = s.a != kGround ? s.a : s.b;
      if (fixed_by[node] >= 0) { bg::set_error("netlist line %d: node %s is held by two voltage sources", s.line, N.node_name[node].c_str()); return BG_ERR_ARG; }
      fixed_by[node] = k;
      sign_of[node] = s.a != kGround ? 1.0 : -1.0;  // v(a) - v(b) = u
      fixed_by[nn + nl + k] = k;
      N.elim_row.push_back(node);
      N.elim_src.push_back(k);
      N.elim_sign.push_back(sign_of[node]);
    }
    for (int i = 0; i < size; ++i) {
      if (fixed_by[i] >= 0) N.new_of_mna[i] = -1;
      else { N.new_of_mna[i] = (int)N.kept.size(); N.kept.push_back(i); }
    }
    Triplets g2, c2, b2;
    for (size_t k = 0; k < g.r.size(); ++k) {
      const int i = N.new_of_mna[g.r[k]], j = N.new_of_mna[g.c[k]];
      if (i < 0) continue;
      if (j >= 0) g2.push(i, j, g.v[k]);
      else if (g.c[k] < nn) b2.push(i, fixed_by[g.c[k]], -g.v[k] * sign_of[g.c[k]]);  // known voltage to the right-hand side
      else { bg::set_error("netlist: internal error, eliminated branch current referenced by a kept row"); return BG_ERR_ARG; }
    }
    for (size_t k = 0; k < c.r.size(); ++k) {
      const int i = N.new_of_mna[c.r[k]], j = N.new_of_mna[c.c[k]];
      if (i < 0) continue;
      if (j >= 0) c2.push(i, j, c.v[k]);
      else if (N.vs[fixed_by[c.c[k]]].kind != BG_SRC_DC) {
        bg::set_error("netlist: a capacitor couples node %s, held by a time-varying voltage source, to the grid (needs du/dt; use BG_NETLIST_MNA)", N.node_name[c.c[k]].c_str());
        return BG_ERR_UNSUPPORTED;
      }
    }
    // B: existing entries of kept rows first (file order), then the moved voltages
    Triplets bb;
    for (size_t k = 0; k < b.r.size(); ++k)
      if (N.new_of_mna[b.r[k]] >= 0) bb.push(N.new_of_mna[b.r[k]], b.c[k], b.v[k]);
    for (size_t k = 0; k < b2.r.size(); ++k) bb.push(b2.r[k], b2.c[k], b2.v[k]);
    const int m = (int)N.kept.size();
    assemble(g2, m, m, N.G);
    assemble(c2, m, m, N.C);
    assemble(bb, m, nsrc, N.B);
  } else {
    for (int i = 0; i < size; ++i) { N.new_of_mna[i] = i; N.kept.push_back(i); }
    assemble(g, size, size, N.G);
    assemble(c, size, size, N.C);
    assemble(b, size, nsrc, N.B);
  }
  N.port_row.clear();
  for (const std::string& nm : N.port_name) {  // like NodeList::pushPort: unknown names and ground are skipped (src/element.cpp:163-175)
    const int id = N.lookup(nm.data(), nm.size());
    if (id < 0) continue;
    N.port_row.push_back(N.new_of_mna[id]);  // -1: the port is a fixed node of the reduced system
  }
  return BG_OK;
}

int parse_common(const char* text, size_t len, const std::string& dir, int formulation, bg_netlist_t* out) {
  if (formulation != BG_NETLIST_MNA && formulation != BG_NETLIST_REDUCED) { bg::set_error("bg_netlist: unknown formulation %d", formulation); return BG_ERR_ARG; }
  bg_netlist_s* N = new bg_netlist_s;
  N->formulation = formulation;
  Parser P(*N);
  const char* te = std::getenv("B200GMRES_SETUP_TIMING");
  const bool timing = te && te[0] == '1';
  const auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (timing) std::fprintf(stderr, "[b200gmres] netlist %-10s %.3f s\n", what, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  };
  if (!P.text(text, len, dir)) {  // text[len] == 0 (callers pass std::string storage)
    bg::set_error("%s", P.err.c_str());
    const bool unsupported = P.err.find("not supported") != std::string::npos;
    delete N;
    return unsupported ? BG_ERR_UNSUPPORTED : BG_ERR_ARG;
  }
  lap("parsed");
  const int rc = build(*N);
  lap("assembled");
  if (rc < 0) { delete N; return rc; }
  *out = N;
  return BG_OK;
}
}  // namespace

extern "C" int bg_netlist_read(const char* path, int formulation, bg_netlist_t* out) {
  if (!path || !out) { bg::set_error("bg_netlist_read: null argument"); return BG_ERR_ARG; }
  *out = nullptr;
  std::string text;
  if (!Parser::slurp(path, text)) { bg::set_error("bg_netlist_read: cannot open %s", path); return BG_ERR_ARG; }
  const std::string p(path);
  const size_t slash = p.find_last_of('/');
  return parse_common(text.c_str(), text.size(), slash == std::string::npos ? std::string() : p.substr(0, slash + 1), formulation, out);
}
extern "C" int bg_netlist_parse(const char* text, int64_t len, const char* include_dir, int formulation, bg_netlist_t* out) {
  if (!text || len < 0 || !out) { bg::set_error("bg_netlist_parse: bad argument"); return BG_ERR_ARG; }
  *out = nullptr;
  std::string dir = include_dir ? include_dir : "";
  if (!dir.empty() && dir.back() != '/') dir.push_back('/');
  const std::string copy(text, (size_t)len);  // the tokenizer wants a terminator after the last token
  return parse_common(copy.c_str(), copy.size(), dir, formulation, out);
}
extern "C" int bg_netlist_get_info(bg_netlist_t N, bg_netlist_info* o) {
  if (!N || !o) return BG_ERR_ARG;
  std::memset(o, 0, sizeof *o);
  o->n_nodes = N->n_nodes(); o->n_inductors = (int)N->ind.size(); o->n_vsources = (int)N->vs.size(); o->n_isources = (int)N->is.size();
  o->n_resistors = (int)N->res.size(); o->n_capacitors = (int)N->cap.size();
  o->size = N->G.n_rows; o->size_mna = N->size_mna(); o->n_sources = o->n_vsources + o->n_isources;
  o->n_ports = (int)N->port_row.size(); o->formulation = N->formulation; o->n_eliminated = (int)N->elim_row.size();
  o->nnz_g = (int64_t)N->G.ci.size(); o->nnz_c = (int64_t)N->C.ci.size(); o->nnz_b = (int64_t)N->B.ci.size();
  o->n_pwl_points = (int64_t)N->pwl_t.size(); o->n_pulse = (int)N->pulse_src.size();
  o->tstep = N->tstep; o->tstop = N->tstop; o->dc = N->tstep == 0 && N->tstop == 0;  // src/etbr_wrapper.cpp:77-81
  o->n_lines = N->n_lines;
  return BG_OK;
}
extern "C" int bg_netlist_get_matrix(bg_netlist_t N, int which, int* rowptr, int* col, double* val) {
  if (!N || !rowptr) return BG_ERR_ARG;
  const Csr* M = which == 'G' || which == 'g' ? &N->G : which == 'C' || which == 'c' ? &N->C : which == 'B' || which == 'b' ? &N->B : nullptr;
  if (!M) { bg::set_error("bg_netlist_get_matrix: which must be 'G', 'C' or 'B'"); return BG_ERR_ARG; }
  if (!M->ci.empty() && (!col || !val)) return BG_ERR_ARG;
  std::copy(M->rp.begin(), M->rp.end(), rowptr);
  std::copy(M->ci.begin(), M->ci.end(), col);
  std::copy(M->v.begin(), M->v.end(), val);
  return BG_OK;
}
// the same matrix column-compressed with 64-bit indices and rows ascending inside a column: exactly the p / i / x
// arrays of the cs_dl the reference's parser_wrapper hands on (matrix::sort + mat2csdl, src/matrix.cpp:127-135, 343-362)
extern "C" int bg_netlist_get_matrix_csc(bg_netlist_t N, int which, int64_t* colptr, int64_t* row, double* val) {
  if (!N || !colptr) return BG_ERR_ARG;
  const Csr* M = which == 'G' || which == 'g' ? &N->G : which == 'C' || which == 'c' ? &N->C : which == 'B' || which == 'b' ? &N->B : nullptr;
  if (!M) { bg::set_error("bg_netlist_get_matrix_csc: which must be 'G', 'C' or 'B'"); return BG_ERR_ARG; }
  if (!M->ci.empty() && (!row || !val)) return BG_ERR_ARG;
  const int nc = M->n_cols;
  for (int j = 0; j <= nc; ++j) colptr[j] = 0;
  for (int c : M->ci) colptr[(size_t)c + 1]++;
  for (int j = 0; j < nc; ++j) colptr[(size_t)j + 1] += colptr[j];
  std::vector<int64_t> next(colptr, colptr + nc);
  for (int i = 0; i < M->n_rows; ++i)   // rows visited in ascending order: every column comes out sorted
    for (int p = M->rp[i]; p < M->rp[(size_t)i + 1]; ++p) {
      const int64_t q = next[(size_t)M->ci[p]]++;
      row[q] = i;
      val[q] = M->v[p];
    }
  return BG_OK;
}
extern "C" int bg_netlist_get_sources(bg_netlist_t N, int* kind, int* pwl_ptr, double* pwl_t, double* pwl_v, int* pulse_src,
                                      double* pulse_time, double* pulse_val) {
  if (!N) return BG_ERR_ARG;
  if (kind) std::copy(N->kind.begin(), N->kind.end(), kind);
  if (pwl_ptr) std::copy(N->pwl_ptr.begin(), N->pwl_ptr.end(), pwl_ptr);
  if (pwl_t) std::copy(N->pwl_t.begin(), N->pwl_t.end(), pwl_t);
  if (pwl_v) std::copy(N->pwl_v.begin(), N->pwl_v.end(), pwl_v);
  if (pulse_src) std::copy(N->pulse_src.begin(), N->pulse_src.end(), pulse_src);
  if (pulse_time) std::copy(N->pulse_time.begin(), N->pulse_time.end(), pulse_time);
  if (pulse_val) std::copy(N->pulse_val.begin(), N->pulse_val.end(), pulse_val);
  return BG_OK;
}
extern "C" int bg_netlist_get_ports(bg_netlist_t N, int* rows) {
  if (!N || (!rows && !N->port_row.empty())) return BG_ERR_ARG;
  std::copy(N->port_row.begin(), N->port_row.end(), rows);
  return BG_OK;
}
extern "C" int bg_netlist_node_row(bg_netlist_t N, const char* name) {
  if (!N || !name) return -2;
  const int id = N->lookup(name, std::strlen(name));
  if (id < 0) return id;  // -1 ground, -2 unknown
  const int r = N->new_of_mna[id];
  return r < 0 ? -3 : r;
}
extern "C" int bg_netlist_row_name(bg_netlist_t N, int row, char* buf, int cap) {
  if (!N || !buf || cap < 1 || row < 0 || row >= (int)N->kept.size()) return BG_ERR_ARG;
  const int q = N->kept[row], nn = N->n_nodes(), nl = (int)N->ind.size();
  if (q < nn) std::snprintf(buf, (size_t)cap, "%s", N->node_name[q].c_str());
  else if (q < nn + nl) std::snprintf(buf, (size_t)cap, "i(L#%d)", q - nn);
  else std::snprintf(buf, (size_t)cap, "i(V#%d)", q - nn - nl);
  return BG_OK;
}
extern "C" int bg_netlist_get_reduction(bg_netlist_t N, int* kept_mna_row, int* elim_mna_row, int* elim_source, double* elim_sign) {
  if (!N) return BG_ERR_ARG;
  if (kept_mna_row) std::copy(N->kept.begin(), N->kept.end(), kept_mna_row);
  if (elim_mna_row) std::copy(N->elim_row.begin(), N->elim_row.end(), elim_mna_row);
  if (elim_source) std::copy(N->elim_src.begin(), N->elim_src.end(), elim_source);
  if (elim_sign) std::copy(N->elim_sign.begin(), N->elim_sign.end(), elim_sign);
  return BG_OK;
}
extern "C" int bg_netlist_destroy(bg_netlist_t N) {
  delete N;
  return BG_OK;
}
