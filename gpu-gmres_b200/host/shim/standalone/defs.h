// defs.h (B200 shim, STANDALONE fallback) -- the constants and macros of the reference's src/defs.h:11-75 for
// building the shim itself and drivers outside the reference tree.  Inside the reference tree the tree's own defs.h
// is used instead (it is not part of the replaced layer): keep this directory LAST on the include path.
#ifndef DEFS_H_
#define DEFS_H_
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <sys/time.h>
const int restart = 32, max_iter = 60000;
const float tolerance = 1e-6;
enum PreconditionerType { NONE, DIAG, ILU0, ILUK, AINV };
#define trace
#define perr printf("<<<<<<<<<< ERROR: File: %s, Line: %d\n", __FILE__, __LINE__)
#define prval(x)
#define TIME(statement, acc_time) { timeval st, et; gettimeofday(&st, NULL); (statement); gettimeofday(&et, NULL); acc_time += difftime(st, et); }
const float eps = 1e-9;
#define Equal(a, b) (abs((a) - (b)) < eps ? true : false)
#define HandleError { cudaError_t cudaError = cudaGetLastError(); if (cudaError != cudaSuccess) { printf("Error: %s\n", cudaGetErrorString(cudaError)); exit(-1); } }
extern float difftime(timeval& st, timeval& et);
#define MAX_STR_LEN 1024
#endif
