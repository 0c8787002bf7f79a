// oracle/ref_stubs.cpp -- TEST INFRASTRUCTURE ONLY.  Link-time stand-ins for the legacy cuBLAS / cuSPARSE
// entry points the reference's DEVICE paths call (src/gmres.cu:2342-2363, src/preconditioner.cu:263-358,
// 1581-1637).  CUDA 12 no longer ships the cuSPARSE ones; the oracle only runs the reference's HOST twins, so
// every stub aborts loudly if it is ever reached.
#include <cstdio>
#include <cstdlib>
#define STUB(name) extern "C" int name(...) { std::fprintf(stderr, "oracle/_ref: device path %s is not available\n", #name); std::abort(); }
STUB(cublasSaxpy) STUB(cublasScopy) STUB(cublasSdot) STUB(cublasSgemv) STUB(cublasSnrm2) STUB(cublasSscal)
STUB(cusparseCreate) STUB(cusparseCreateMatDescr) STUB(cusparseCreateSolveAnalysisInfo) STUB(cusparseDcsrmv)
STUB(cusparseDcsrsv_solve) STUB(cusparseDestroyMatDescr) STUB(cusparseDestroySolveAnalysisInfo) STUB(cusparseGetVersion)
STUB(cusparseScsrmv) STUB(cusparseScsrsv_analysis) STUB(cusparseScsrsv_solve) STUB(cusparseSetMatDiagType)
STUB(cusparseSetMatFillMode) STUB(cusparseSetMatIndexBase) STUB(cusparseSetMatType) STUB(cusparseDcsrsv_analysis)
