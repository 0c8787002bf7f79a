// oracle/shim/cusparse_v2.h -- TEST INFRASTRUCTURE ONLY.  The reference calls legacy cuSPARSE entry points
// (csrsv analysis/solve, csrmv) that CUDA 12 removed.  They are only DECLARED here so that the reference's
// translation units compile unchanged; the oracle build never executes a device path (ref_stubs.cpp aborts).
#pragma once
#include_next <cusparse_v2.h>
struct cusparseSolveAnalysisInfo;
typedef struct cusparseSolveAnalysisInfo* cusparseSolveAnalysisInfo_t;
extern "C" {
cusparseStatus_t cusparseCreateSolveAnalysisInfo(cusparseSolveAnalysisInfo_t* info);
cusparseStatus_t cusparseDestroySolveAnalysisInfo(cusparseSolveAnalysisInfo_t info);
cusparseStatus_t cusparseScsrsv_analysis(cusparseHandle_t, cusparseOperation_t, int, int, const cusparseMatDescr_t, const float*, const int*, const int*, cusparseSolveAnalysisInfo_t);
cusparseStatus_t cusparseDcsrsv_analysis(cusparseHandle_t, cusparseOperation_t, int, int, const cusparseMatDescr_t, const double*, const int*, const int*, cusparseSolveAnalysisInfo_t);
cusparseStatus_t cusparseScsrsv_solve(cusparseHandle_t, cusparseOperation_t, int, const float*, const cusparseMatDescr_t, const float*, const int*, const int*, cusparseSolveAnalysisInfo_t, const float*, float*);
cusparseStatus_t cusparseDcsrsv_solve(cusparseHandle_t, cusparseOperation_t, int, const double*, const cusparseMatDescr_t, const double*, const int*, const int*, cusparseSolveAnalysisInfo_t, const double*, double*);
cusparseStatus_t cusparseScsrmv(cusparseHandle_t, cusparseOperation_t, int, int, int, const float*, const cusparseMatDescr_t, const float*, const int*, const int*, const float*, const float*, float*);
cusparseStatus_t cusparseDcsrmv(cusparseHandle_t, cusparseOperation_t, int, int, int, const double*, const cusparseMatDescr_t, const double*, const int*, const int*, const double*, const double*, double*);
}
