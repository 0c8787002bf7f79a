// dist.cu -- row-partitioned GMRES(m) across the GPUs of one box (north star:
// "the matrix is row-partitioned ... boundary-vector halo exchange uses NCCL point-to-point over
// NVLink, Arnoldi inner products use a batched NCCL allreduce, and the preconditioner becomes a
// per-partition block-Jacobi ILU").  The reference itself is single-GPU; its row
// partitioners (src/partition.cpp, partition3.cpp) only feed a CPU Schur solver.
//
// One process per GPU.  Rank p owns the rows of its part, all Krylov vectors restricted to those
// rows and the ILU factors of its own diagonal block; there is no separator set.  Per inner
// iteration the ranks exchange
//   * one halo message per neighbour (the boundary entries of the vector entering the SpMV, packed
//     by a kernel, grouped ncclSend/ncclRecv straight into the halo slots behind the vector), and
//   * ONE ncclAllReduce of j+2 doubles: all projections V^T w and w.w of the fused CGS step.
// Triangular solves need no communication.  Both calls are enqueued on the context stream, so the
// host never waits for a collective; stopping is decided from the allreduced (rank-identical)
// values, which keeps every rank's enqueue sequence identical.
//
// NCCL is resolved with dlopen at first use (libnccl.so.2, the copy the host framework already
// loaded if any), so single-GPU users need no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "internal.h"

using namespace bg;

namespace {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { set_error("NCCL not found (libnccl.so.2): %s", dlerror()); return api; }
#define BG_SYM(name) *(void**)(&api.name) = dlsym(h, "nccl" #name)
  BG_SYM(GetUniqueId); BG_SYM(CommInitRank); BG_SYM(CommDestroy); BG_SYM(Send); BG_SYM(Recv); BG_SYM(AllReduce);
  BG_SYM(GroupStart); BG_SYM(GroupEnd); BG_SYM(GetErrorString);
#undef BG_SYM
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Send && api.Recv && api.AllReduce && api.GroupStart &&
           api.GroupEnd && api.GetErrorString;
  if (!api.ok) set_error("libnccl.so.2 lacks a required entry point");
  return api;
}
#define BG_NCCL(call)                                                                              \
  do {                                                                                             \
    ncclResult_t r__ = (call);                                                                     \
    if (r__ != ncclSuccess) { set_error("NCCL error %d (%s) in %s", (int)r__, nccl().GetErrorString(r__), #call); return BG_ERR_COMM; } \
  } while (0)

}  // namespace

struct bg_dist_plan_s {
  DistPlan<double> p;
};

struct bg_dist_s : DistHooks {
  bg_context_t ctx = nullptr;
  int nranks = 1, rank = 0, dtype = BG_F64, n_local = 0, n_halo = 0;
  int64_t n_global = 0;
  ncclComm_t comm = nullptr;
  bg_matrix_t A = nullptr;
  bg_precond_t M = nullptr;
  bg_solver_t S = nullptr;
  std::vector<int> nbr, recv_ptr, send_ptr;
  DevBuf send_idx, sendbuf;
  int64_t halo_msgs = 0, allreduces = 0;

  int exchange(void* vec, const int* skip) override {
    if (nranks == 1 || nbr.empty()) return BG_OK;
    const size_t es = dtype_size(dtype);
    const ncclDataType_t dt = dtype == BG_F32 ? ncclFloat32 : ncclFloat64;
    BG_TRY(launch_pack(ctx, dtype, send_ptr.back(), send_idx.as<int>(), vec, sendbuf.p, skip));
    NcclApi& N = nccl();
    BG_NCCL(N.GroupStart());
    for (size_t k = 0; k < nbr.size(); ++k) {
      const size_t ns = (size_t)(send_ptr[k + 1] - send_ptr[k]), nr = (size_t)(recv_ptr[k + 1] - recv_ptr[k]);
      if (ns) BG_NCCL(N.Send(sendbuf.as<char>() + (size_t)send_ptr[k] * es, ns, dt, nbr[k], comm, ctx->stream));
      if (nr) BG_NCCL(N.Recv((char*)vec + ((size_t)n_local + recv_ptr[k]) * es, nr, dt, nbr[k], comm, ctx->stream));
    }
    BG_NCCL(N.GroupEnd());
    halo_msgs += (int64_t)nbr.size();
    return BG_OK;
  }
  int allreduce(double* buf, int count) override {
    if (nranks == 1) return BG_OK;
    BG_NCCL(nccl().AllReduce(buf, buf, (size_t)count, ncclFloat64, ncclSum, comm, ctx->stream));
    ++allreduces;
    return BG_OK;
  }
};

// ------------------------------------------------------------------------------- host plan
extern "C" int bg_host_dist_plan(int n, const int* rp, const int* ci, const void* val, int val_dtype, int nranks, int rank,
                                 const int* part, bg_dist_plan_t* out) {
  if (!out || n < 0 || nranks < 1 || rank < 0 || rank >= nranks || (n && (!rp || !ci || !val || !part))) return BG_ERR_ARG;
  bg_dist_plan_t P = new bg_dist_plan_s;
  try {
    std::vector<double> v((size_t)rp[n]);
    if (val_dtype == BG_F64) std::memcpy(v.data(), val, sizeof(double) * v.size());
    else for (size_t i = 0; i < v.size(); ++i) v[i] = (double)((const float*)val)[i];
    build_dist_plan<double>(n, rp, ci, v.data(), nranks, rank, part, P->p);
  } catch (const std::exception& e) { set_error("partition plan failed: %s", e.what()); delete P; return BG_ERR_ARG; }
  *out = P;
  return BG_OK;
}
extern "C" int bg_dist_plan_get_info(bg_dist_plan_t P, bg_dist_plan_info* o) {
  if (!P || !o) return BG_ERR_ARG;
  o->nranks = P->p.nranks; o->rank = P->p.rank; o->n_global = P->p.n_global; o->n_local = P->p.n_local; o->n_halo = P->p.n_halo;
  o->nnz_local = (int64_t)P->p.A.ci.size(); o->n_neighbors = (int)P->p.nbr.size(); o->send_count = (int64_t)P->p.send_idx.size();
  return BG_OK;
}
extern "C" int bg_dist_plan_get(bg_dist_plan_t P, int* l2g, int* h2g, int* lrp, int* lci, double* lval, int* nbr, int* recv_ptr,
                                int* send_ptr, int* send_idx) {
  if (!P) return BG_ERR_ARG;
  const DistPlan<double>& p = P->p;
  if (l2g) std::copy(p.l2g.begin(), p.l2g.end(), l2g);
  if (h2g) std::copy(p.h2g.begin(), p.h2g.end(), h2g);
  if (lrp) std::copy(p.A.rp.begin(), p.A.rp.end(), lrp);
  if (lci) std::copy(p.A.ci.begin(), p.A.ci.end(), lci);
  if (lval) std::copy(p.A.v.begin(), p.A.v.end(), lval);
  if (nbr) std::copy(p.nbr.begin(), p.nbr.end(), nbr);
  if (recv_ptr) std::copy(p.recv_ptr.begin(), p.recv_ptr.end(), recv_ptr);
  if (send_ptr) std::copy(p.send_ptr.begin(), p.send_ptr.end(), send_ptr);
  if (send_idx) std::copy(p.send_idx.begin(), p.send_idx.end(), send_idx);
  return BG_OK;
}
extern "C" int bg_dist_plan_destroy(bg_dist_plan_t P) { delete P; return BG_OK; }

// ------------------------------------------------------------------------------- device side
extern "C" int bg_dist_unique_id(void* id) {
  if (!id) return BG_ERR_ARG;
  NcclApi& N = nccl();
  if (!N.ok) return BG_ERR_COMM;
  static_assert(sizeof(ncclUniqueId) == BG_DIST_ID_BYTES, "unique id size");
  BG_NCCL(N.GetUniqueId((ncclUniqueId*)id));
  return BG_OK;
}

extern "C" int bg_dist_create(bg_context_t ctx, bg_dist_plan_t P, const void* id, int dtype, int m, const bg_precond_opts* po,
                              int use_ilut, bg_dist_t* out) {
  if (!ctx || !P || !out || (P->p.nranks > 1 && !id)) return BG_ERR_ARG;
  *out = nullptr;
  BG_CUDA(cudaSetDevice(ctx->device));
  const DistPlan<double>& p = P->p;
  bg_dist_t D = new bg_dist_s;
  D->ctx = ctx; D->nranks = p.nranks; D->rank = p.rank; D->dtype = dtype; D->n_local = p.n_local; D->n_halo = p.n_halo;
  D->n_global = p.n_global; D->nbr = p.nbr; D->recv_ptr = p.recv_ptr; D->send_ptr = p.send_ptr;
  int rc = BG_OK;
  do {
    if (p.nranks > 1) {
      NcclApi& N = nccl();
      if (!N.ok) { rc = BG_ERR_COMM; break; }
      ncclUniqueId uid;
      std::memcpy(&uid, id, sizeof(uid));
      ncclResult_t r = N.CommInitRank(&D->comm, p.nranks, uid, p.rank);
      if (r != ncclSuccess) { set_error("ncclCommInitRank: %s", N.GetErrorString(r)); rc = BG_ERR_COMM; break; }
    }
    // local rows: own columns then halo slots
    rc = matrix_create_internal(ctx, p.n_local, p.n_local + p.n_halo, p.A.rp, p.A.ci, p.A.v.data(), BG_F64, dtype, BG_SPMV_AUTO, &D->A);
    if (rc < 0) break;
    // block-Jacobi ILU of the own diagonal block (no communication in the triangular solves)
    std::vector<int> drp(p.n_local + 1, 0), dci;
    std::vector<double> dv;
    for (int i = 0; i < p.n_local; ++i) {
      for (int q = p.A.rp[i]; q < p.A.rp[i + 1]; ++q)
        if (p.A.ci[q] < p.n_local) { dci.push_back(p.A.ci[q]); dv.push_back(p.A.v[q]); }
      drp[i + 1] = (int)dci.size();
    }
    bg_precond_opts o;
    bg_precond_opts_default(&o);
    if (po) o = *po;
    o.trsv_schedule = BG_TRSV_BLOCK;
    rc = use_ilut ? bg_precond_create_ilut(ctx, p.n_local, drp.data(), dci.data(), dv.data(), BG_F64, dtype, &o, &D->M)
                  : bg_precond_create_ilu0(ctx, p.n_local, drp.data(), dci.data(), dv.data(), BG_F64, dtype, &o, &D->M);
    if (rc < 0) break;
    rc = solver_create_internal(ctx, p.n_local, m, dtype, p.n_halo, &D->S);
    if (rc < 0) break;
    // the solver works in the preconditioner's block-major order: send lists follow
    std::vector<int> inv(p.n_local), sidx(p.send_idx.size());
    for (int i = 0; i < p.n_local; ++i) inv[D->M->fb.h_perm[i]] = i;
    for (size_t k = 0; k < sidx.size(); ++k) sidx[k] = inv[p.send_idx[k]];
    rc = D->send_idx.upload(sidx.data(), sizeof(int) * sidx.size(), ctx->stream);
    if (rc < 0) break;
    rc = D->sendbuf.alloc(dtype_size(dtype) * (sidx.size() + 2));
    if (rc < 0) break;
    cudaStreamSynchronize(ctx->stream);
  } while (false);
  if (rc < 0) { bg_dist_destroy(D); return rc; }
  *out = D;
  return BG_OK;
}

extern "C" int bg_dist_destroy(bg_dist_t D) {
  if (!D) return BG_OK;
  if (D->ctx) cudaStreamSynchronize(D->ctx->stream);
  if (D->S) bg_solver_destroy(D->S);
  if (D->M) bg_precond_destroy(D->M);
  if (D->A) bg_matrix_destroy(D->A);
  if (D->comm && nccl().ok) nccl().CommDestroy(D->comm);
  delete D;
  return BG_OK;
}

extern "C" int bg_dist_gmres_solve(bg_dist_t D, const void* d_b, void* d_x, const bg_gmres_opts* o, bg_gmres_result* r) {
  if (!D || !d_b || !d_x || !o) return BG_ERR_ARG;
  return gmres_solve_dist(D->ctx, D->S, D->A, D->M, d_b, d_x, o, r, D);
}

extern "C" int bg_dist_gmres_solve_host(bg_dist_t D, const void* h_b, void* h_x, const bg_gmres_opts* o, bg_gmres_result* r) {
  if (!D || !h_b || !h_x || !o) return BG_ERR_ARG;
  bg_solver_t S = D->S;
  const size_t bytes = (size_t)D->n_local * dtype_size(D->dtype);
  if (S->hx.bytes < bytes) { BG_TRY(S->hx.alloc(bytes + 16)); BG_TRY(S->hb.alloc(bytes + 16)); }
  BG_CUDA(cudaMemcpyAsync(S->hx.p, h_x, bytes, cudaMemcpyHostToDevice, D->ctx->stream));
  BG_CUDA(cudaMemcpyAsync(S->hb.p, h_b, bytes, cudaMemcpyHostToDevice, D->ctx->stream));
  int rc = bg_dist_gmres_solve(D, S->hb.p, S->hx.p, o, r);
  if (rc < 0) return rc;
  BG_CUDA(cudaMemcpyAsync(h_x, S->hx.p, bytes, cudaMemcpyDeviceToHost, D->ctx->stream));
  BG_CUDA(cudaStreamSynchronize(D->ctx->stream));
  return rc;
}

extern "C" int bg_dist_get_info(bg_dist_t D, bg_dist_info* o) {
  if (!D || !o) return BG_ERR_ARG;
  std::memset(o, 0, sizeof(*o));
  o->nranks = D->nranks; o->rank = D->rank; o->n_global = D->n_global; o->n_local = D->n_local; o->n_halo = D->n_halo;
  o->n_neighbors = (int)D->nbr.size(); o->halo_messages = D->halo_msgs; o->allreduces = D->allreduces;
  o->halo_bytes_per_exchange = (int64_t)D->send_ptr.back() * (int64_t)dtype_size(D->dtype);
  bg_precond_get_info(D->M, &o->precond);
  return BG_OK;
}
