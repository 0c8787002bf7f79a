// oracle/ref_parser_wrap.cpp -- TEST INFRASTRUCTURE ONLY.
// C interface over the reference's own netlist front end, compiled UNCHANGED from /root/reference/src (parser.cpp,
// element.cpp, matrix.cpp, namepool.cpp, hashtable.cpp, circuit.cpp) against the stand-in headers under oracle/shim
// (IT++ and CXSparse are not in this image).  It runs the same sequence as parser_wrapper (src/etbr_wrapper.cpp:46-171:
// parser_sub, then stampB, stampC, stampG, each followed by mat2csdl) and hands the results out, so that
// tests/test_netlist.py can pin the product's bg_netlist_* against the reference's parser bit for bit.
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "parser.h"
#include "matrix.h"

cs* cs_spalloc(int m, int n, int nzmax, int values, int) {
  cs* A = (cs*)calloc(1, sizeof(cs));
  A->m = m; A->n = n; A->nzmax = nzmax; A->nz = -1;
  A->p = (int*)calloc((size_t)n + 1, sizeof(int));
  A->i = (int*)calloc((size_t)(nzmax > 0 ? nzmax : 1), sizeof(int));
  A->x = values ? (double*)calloc((size_t)(nzmax > 0 ? nzmax : 1), sizeof(double)) : nullptr;
  return A;
}
cs_dl* cs_dl_spalloc(UF_long m, UF_long n, UF_long nzmax, UF_long values, UF_long) {
  cs_dl* A = (cs_dl*)calloc(1, sizeof(cs_dl));
  A->m = m; A->n = n; A->nzmax = nzmax; A->nz = -1;
  A->p = (UF_long*)calloc((size_t)n + 1, sizeof(UF_long));
  A->i = (UF_long*)calloc((size_t)(nzmax > 0 ? nzmax : 1), sizeof(UF_long));
  A->x = values ? (double*)calloc((size_t)(nzmax > 0 ? nzmax : 1), sizeof(double)) : nullptr;
  return A;
}

namespace {
struct State {
  cs_dl *G = nullptr, *C = nullptr, *B = nullptr;
  Source *VS = nullptr, *IS = nullptr;
  int nVS = 0, nIS = 0, nL = 0, nNodes = 0, nsub = 0, size = 0;
  double tstep = 0, tstop = 0;
  std::vector<int> ports;
  std::map<int, std::string> names;
  gpuETBR etbr;
} st;
}  // namespace

extern "C" int refp_parse(const char* path) {
  st = State();
  std::memset(&st.etbr, 0, sizeof st.etbr);
  NodeList* pool = new NodeList;
  parser_sub(path, st.tstep, st.tstop, st.nsub, st.nVS, st.nIS, st.nL, pool);
  if (st.nsub != 0) return -1;  // sub-circuits: not compared
  st.nNodes = pool->numNode();
  st.size = st.nNodes + st.nL + st.nVS;
  for (int i = 0; i < pool->numPort(); ++i) st.ports.push_back(pool->getPort(i)->row_no);
  pool->get_name_table(st.names);
  matrix* G = new matrix(st.size, st.size);
  matrix* C = new matrix(st.size, st.size);
  matrix* B = new matrix(st.size, st.nVS + st.nIS);
  st.VS = new Source[st.nVS > 0 ? st.nVS : 1];
  st.IS = new Source[st.nIS > 0 ? st.nIS : 1];
  stampB(path, st.nL, st.nIS, st.nVS, st.nNodes, st.tstop, st.VS, st.IS, B, pool, &st.etbr);
  st.B = B->mat2csdl();
  stampC(path, st.nL, st.nVS, st.nNodes, C, pool);
  st.C = C->mat2csdl();
  stampG(path, st.nL, st.nVS, st.nNodes, G, pool);
  st.G = G->mat2csdl();
  return 0;
}
extern "C" void refp_info(int* counts /* nNodes nL nVS nIS nports */, double* tran /* tstep tstop */, long* nnz /* G C B */) {
  counts[0] = st.nNodes; counts[1] = st.nL; counts[2] = st.nVS; counts[3] = st.nIS; counts[4] = (int)st.ports.size();
  tran[0] = st.tstep; tran[1] = st.tstop;
  const cs_dl* M[3] = {st.G, st.C, st.B};
  for (int k = 0; k < 3; ++k) nnz[k] = M[k] ? M[k]->p[M[k]->n] : 0;
}
extern "C" void refp_matrix(int which, long* colptr, long* rowidx, double* x) {  // column-compressed, as mat2csdl leaves it
  const cs_dl* M = which == 0 ? st.G : which == 1 ? st.C : st.B;
  if (!M) return;
  const long nz = M->p[M->n];
  for (long j = 0; j <= M->n; ++j) colptr[j] = M->p[j];
  for (long k = 0; k < nz; ++k) { rowidx[k] = M->i[k]; x[k] = M->x[k]; }
}
extern "C" int refp_source_len(int k) { return k < st.nVS ? st.VS[k].time.size() : st.IS[k - st.nVS].time.size(); }
extern "C" void refp_source(int k, double* t, double* v) {
  Source& s = k < st.nVS ? st.VS[k] : st.IS[k - st.nVS];
  for (int i = 0; i < s.time.size(); ++i) { t[i] = s.time(i); v[i] = s.value(i); }
}
extern "C" void refp_ports(int* rows) { for (size_t k = 0; k < st.ports.size(); ++k) rows[k] = st.ports[k]; }
extern "C" int refp_node_name(int row, char* buf, int cap) {
  auto it = st.names.find(row);
  if (it == st.names.end()) return -1;
  std::snprintf(buf, (size_t)cap, "%s", it->second.c_str());
  return 0;
}
extern "C" int refp_pulse_current(double* time5, double* val2) {  // the PULSE records gen_PULSEut_kernel consumes
  const int n = st.etbr.PULSEcurExist;
  if (time5) for (int k = 0; k < 5 * n; ++k) time5[k] = st.etbr.PULSEtime_host[k];
  if (val2) for (int k = 0; k < 2 * n; ++k) val2[k] = st.etbr.PULSEval_host[k];
  return n;
}
