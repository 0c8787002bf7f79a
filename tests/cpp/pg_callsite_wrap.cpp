// pg_callsite_wrap.cpp -- TEST INFRASTRUCTURE.  Compiles the reference's own power-grid call sites
// (src/mna_solve_gpu_gmres.cpp:507-532 and :608-621), which tests/test_shim.py extracts VERBATIM from
// /root/reference into pg_callsite_A.inc / pg_callsite_B.inc at test time (nothing of the reference is committed),
// against the shim's gmres_interface_pg.h.  The surrounding driver (IT++ vectors, CXSparse, ILU++ set-up) is
// replaced by the minimal declarations those lines need.
#include <cstdio>
#include <iostream>
#include <vector>

#include "gmres_interface_pg.h"
using namespace std;

struct VecLike {  // itpp::vec stand-in: only _data() is used by the call sites
  std::vector<double> v;
  double* _data() { return v.data(); }
};
struct TimerLike { void start() {} void stop() {} double get_time() { return 0.0; } };

int pg_callsite(int n, int useDoubleILU, int nsteps, MySpMatrix& GmySpM, MySpMatrix& AmySpM, MySpMatrixDouble& PrLeft_GmySpMdouble,
                MySpMatrixDouble& PrRight_GmySpMdouble, MySpMatrix& PrMiddle_GmySpM, MySpMatrix& PrPermRow_GmySpM,
                MySpMatrix& PrPermCol_GmySpM, MySpMatrixDouble& PrLscale_GmySpMdouble, MySpMatrixDouble& PrRscale_GmySpMdouble,
                MySpMatrix& PrLscale_GmySpM, MySpMatrix& PrRscale_GmySpM, MySpMatrixDouble& PrLeft_AmySpMdouble,
                MySpMatrixDouble& PrRight_AmySpMdouble, MySpMatrix& PrMiddle_AmySpM, MySpMatrix& PrPermRow_AmySpM,
                MySpMatrix& PrPermCol_AmySpM, MySpMatrixDouble& PrLscale_AmySpMdouble, MySpMatrixDouble& PrRscale_AmySpMdouble,
                MySpMatrix& PrLscale_AmySpM, MySpMatrix& PrRscale_AmySpM, VecLike& w, VecLike& xn1) {
  TimerLike gmresCPUilu_time;
  int iterTotal = 0;
#include "pg_callsite_A.inc"
  for (int step = 0; step < nsteps; ++step) {
#include "pg_callsite_B.inc"
  }
  return iterTotal;
}
