// preconditioner.h (B200 shim) -- the reference's Preconditioner hierarchy (src/preconditioner.h:34-357,
// src_thermal/preconditioner.h) with the same class names, public data members the drivers poke, and virtual
// interface; device work is forwarded to bg_precond_* (include/b200gmres.h).  The legacy cuSPARSE/cusp members are
// gone.  This build has no CPU compute path: HostPrecond* take HOST vectors like the reference's CPU twins, copy them
// to the device, apply the same operator there and copy the result back.
#ifndef PRECONDITIONER_H_
#define PRECONDITIONER_H_
#include <assert.h>
#include <sys/time.h>
#include "SpMV.h"
#include "config.h"
#include "SpMV_kernel.h"
#include "defs.h"
#include "leftILU.h"
#ifndef MAX_STR_LEN  // src_thermal keeps this enum here (src_thermal/preconditioner.h:30), src/ in defs.h (src/defs.h:15)
enum PreconditionerType { NONE, DIAG, ILU0, AINV };
#endif

class Preconditioner {
 public:
  typedef int IndexType;
  typedef float ValueType;
  int numRows;
  // GMRES workspace the PG path used to keep inside the preconditioner (preconditioner.h:47-50).
  // Kept for source compatibility; the B200 solver owns its workspace (bg_solver_t), these stay NULL.
  float *d_r, *d_rr, *d_bb, *d_y;
  float *s, *cs, *sn, *H;
  float *d_v, *d_w, *d_ww;
  void* bg_handle;  // bg_precond_t

  Preconditioner();
  virtual void HostPrecond(const ValueType* i_data, ValueType* o_data);
  virtual void DevPrecond(const ValueType* i_data, ValueType* o_data);
  virtual void Initilize(const MySpMatrix& mySpM) = 0;
  virtual void HostPrecond_rhs(const ValueType* i, ValueType* o);
  virtual void HostPrecond_right(const ValueType* i, ValueType* o);
  virtual void HostPrecond_left(const ValueType* i, ValueType* o);
  virtual void HostPrecond_starting_value(const ValueType* i, ValueType* o);
  virtual void DevPrecond_rhs(float* i_data, float* o_data);
  virtual void DevPrecond_right(float* i_data, float* o_data);
  virtual void DevPrecond_left(float* i_data, float* o_data);
  virtual void DevPrecond_starting_value(float* i_data, float* o_data);
  virtual ~Preconditioner();
};

class MyNONE : public Preconditioner { public: void Initilize(const MySpMatrix& mySpM); };
class MyDIAG : public Preconditioner { public: void Initilize(const MySpMatrix& mySpM); };
// ILU(0) through leftILU semantics on the host, level-scheduled solves on the device
// (MyILU0::Initilize / DevPrecond, src/preconditioner.cu:253-362)
class MyILU0 : public Preconditioner {
 public:
  ValueType *l_val, *u_val;
  IndexType *l_rowIndices, *l_indices, *u_rowIndices, *u_indices;
  MyILU0();
  void Initilize(const MySpMatrix& mySpM);
  ~MyILU0();
};
// the reference's MyILUK / MyILUPP(MySpMatrix) run ITSOL ilukC and then ignore it, using leftILU
// factors: they behave like MyILU0.
class MyILUK : public MyILU0 {};
// approximate inverse applied as  o = W_t (D (Z i))  (MyAINV::HostPrecond / DevPrecond, src/preconditioner.cu:110-137).
// The reference builds the factors with CUSP's nonsym_bridson_ainv(A, 0.1), which is outside the replaced layer:
// Initilize(A) builds them with the shim's own Benzi-Tuma biconjugation (same drop tolerance) on the host, and
// Initilize(factors) takes factors produced elsewhere (e.g. by a CUSP build).
class MyAINV : public Preconditioner {
 public:
  void Initilize(const MySpMatrix& mySpM);
  void Initilize(int n, const int* z_rowIndices, const int* z_indices, const float* z_val, const float* diag_val,
                 const int* w_t_rowIndices, const int* w_t_indices, const float* w_t_val);
};
// split preconditioner from ILU++ factors, FP64 factors / FP32 vectors
// (MyILUPPfloat::Initilize, src/preconditioner.cu:1205-1334)
class MyILUPPfloat : public Preconditioner {
 public:
  void Initilize(const MySpMatrix& mySpM);
  void Initilize(const MySpMatrixDouble& PrLeft, const MySpMatrixDouble& PrRight, const MySpMatrix& PrMiddle,
                 const MySpMatrix& PrPermRow, const MySpMatrix& PrPermCol, const MySpMatrix& PrLscale,
                 const MySpMatrix& PrRscale, int restart);
};
// the same with FP64 scaling vectors (MyILUPP::Initilize, src/preconditioner.cu:872-1016); used by gmresInterfacePG
class MyILUPP : public Preconditioner {
 public:
  void Initilize(const MySpMatrix& mySpM);
  void Initilize(const MySpMatrixDouble& PrLeft, const MySpMatrixDouble& PrRight, const MySpMatrix& PrMiddle,
                 const MySpMatrix& PrPermRow, const MySpMatrix& PrPermCol, const MySpMatrixDouble& PrLscale,
                 const MySpMatrixDouble& PrRscale);
};
#endif
