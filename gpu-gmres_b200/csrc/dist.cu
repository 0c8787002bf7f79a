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
// Two transports carry those two exchanges:
//   BG_TRANSPORT_P2P (default when CUDA IPC between the ranks works): our own kernels over NVLink /
//     NVSwitch peer memory.  Each rank exports a small window through cudaIpcGetMemHandle (the handles
//     travel in one ncclAllGather at creation).  Messages are flag-in-data 16-byte elements (value +
//     sequence number, internal.h LL16), so no fence and no separate flag round trip is needed.  The
//     all-reduce is FUSED INTO the dots / orth kernels: the last CTA stores the rank's partial sums into
//     every rank's inbox, polls its own, sums in rank order and runs the Givens epilogue -- no collective
//     call, no extra launch.  The halo exchange is one kernel that gathers the boundary entries, stores
//     them into the neighbours' inboxes and unpacks its own arrivals (orth.cu halo_exchange_kernel).
//   BG_TRANSPORT_NCCL: pack kernel + grouped ncclSend/ncclRecv, ncclAllReduce + finalize kernel (the
//     baseline the P2P path is measured against; also the fallback when IPC is unavailable).
//
// NCCL is resolved with dlopen at first use (libnccl.so.2, the copy the host framework already
// loaded if any), so single-GPU users need no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>

#include <cstdlib>
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
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
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
  BG_SYM(GetUniqueId); BG_SYM(CommInitRank); BG_SYM(CommDestroy); BG_SYM(Send); BG_SYM(Recv); BG_SYM(AllReduce); BG_SYM(AllGather);
  BG_SYM(GroupStart); BG_SYM(GroupEnd); BG_SYM(GetErrorString);
#undef BG_SYM
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Send && api.Recv && api.AllReduce && api.AllGather && api.GroupStart &&
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
  // peer-memory transport
  int transport = BG_TRANSPORT_NCCL;
  bool p2p_ready = false;
  DevBuf window, d_par, d_ph;
  std::vector<void*> ipc_open;   // mappings of the other ranks' memory, closed in the destructor
  int send_count = 0;
  uint32_t solve_index = 0, exchange_in_solve = 0;
  bool fuse_halo = true;  // B200GMRES_FUSE_HALO=0: separate halo_exchange_kernel instead of the SpMV-carried exchange

  const PeerAR* peer_ar() override { return transport == BG_TRANSPORT_P2P && nranks > 1 ? d_par.as<PeerAR>() : nullptr; }

  // (solve, exchange within the solve): identical on every rank although ranks may enqueue different numbers of
  // (skipped) launches past the stopping iteration of a solve.  32 bits: the solve part wraps after 4096 solves, which
  // is harmless -- an inbox element always carries the number of the LAST exchange (never one from 4096 solves ago),
  // and 0, the value of untouched memory, is never produced (the exchange count starts at 1 and stays below 2^20).
  unsigned next_seq() {
    exchange_in_solve = exchange_in_solve >= 0xFFFFEu ? 1u : exchange_in_solve + 1u;
    return ((solve_index & 0xFFFu) << 20) | exchange_in_solve;
  }
  bool halo_args(HaloArgs* h) override {
    if (transport != BG_TRANSPORT_P2P || nranks == 1 || !fuse_halo) return false;
    h->ph = d_ph.as<PeerHalo>(); h->send_idx = send_idx.as<int>(); h->n_send = send_count; h->n_local = n_local;
    h->seq = next_seq();
    halo_msgs += (int64_t)nbr.size();
    return true;
  }
  int exchange(void* vec, const int* skip) override {
    if (nranks == 1 || nbr.empty()) return BG_OK;
    const size_t es = dtype_size(dtype);
    if (transport == BG_TRANSPORT_P2P) {
      // sequence number of this exchange: (solve, launch within the solve) -- identical on both sides
      // because every rank enqueues the same sequence until convergence (later launches are skipped)
      const unsigned seq = next_seq();
      halo_msgs += (int64_t)nbr.size();
      return launch_halo_exchange(ctx, dtype, d_ph.as<PeerHalo>(), send_count, n_halo, send_idx.as<int>(), vec,
                                  (char*)vec + (size_t)n_local * es, seq, skip);
    }
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

namespace {
// what every rank tells every other rank at creation (one ncclAllGather)
struct PeerRecord {
  cudaIpcMemHandle_t win;
  int n_halo, n_nbr, ok, pad;
  int nbr[kMaxPeers], recv_ptr[kMaxPeers + 1];
};
struct WindowLayout {
  size_t ar_inbox, counters, halo_inbox, bytes;
  WindowLayout(int R, int n_halo) {
    ar_inbox = 0;
    counters = ar_inbox + sizeof(LL16) * 2 * R * kArLd;
    halo_inbox = counters + 64;
    bytes = halo_inbox + sizeof(LL16) * 2 * (size_t)std::max(n_halo, 1);  // two buffers (sequence parity)
  }
};

// Map the other ranks' windows and build the device-side descriptors.  Collective.
// Returns BG_OK with D->p2p_ready = false when peer memory is unavailable on ANY rank (all ranks agree).
int setup_p2p(bg_dist_t D) {
  NcclApi& N = nccl();
  bg_context_t ctx = D->ctx;
  cudaStream_t st = ctx->stream;
  const int R = D->nranks, me = D->rank, nn = (int)D->nbr.size();
  PeerRecord mine;
  std::memset(&mine, 0, sizeof(mine));
  bool ok = R <= kMaxPeers && nn <= kMaxPeers;
  const char* env = getenv("B200GMRES_DIST_TRANSPORT");
  if (env && std::strcmp(env, "nccl") == 0) ok = false;
  const WindowLayout wl(R, D->n_halo);
  BG_TRY(D->window.alloc(wl.bytes));
  BG_CUDA(cudaMemsetAsync(D->window.p, 0, wl.bytes, st));  // sequence numbers start at 1: zero never matches
  if (ok) {
    ok = cudaIpcGetMemHandle(&mine.win, D->window.p) == cudaSuccess;
    if (!ok) cudaGetLastError();
  }
  mine.n_halo = D->n_halo; mine.n_nbr = nn; mine.ok = ok ? 1 : 0;
  for (int k = 0; k < nn && k < kMaxPeers; ++k) { mine.nbr[k] = D->nbr[k]; mine.recv_ptr[k] = D->recv_ptr[k]; }
  if (nn <= kMaxPeers) mine.recv_ptr[nn] = D->recv_ptr[nn];
  // exchange the records (the memset above precedes the all-gather in stream order on every rank, so
  // nobody can store into a window that is not cleared yet)
  DevBuf sb, rb;
  BG_TRY(sb.upload(&mine, sizeof(mine), st));
  BG_TRY(rb.alloc(sizeof(PeerRecord) * R));
  BG_NCCL(N.AllGather(sb.p, rb.p, sizeof(PeerRecord), ncclInt8, D->comm, st));
  std::vector<PeerRecord> all(R);
  BG_CUDA(cudaMemcpyAsync(all.data(), rb.p, sizeof(PeerRecord) * R, cudaMemcpyDeviceToHost, st));
  BG_CUDA(cudaStreamSynchronize(st));
  for (int r = 0; r < R; ++r) ok = ok && all[r].ok;
  std::vector<char*> win(R, nullptr);
  if (ok) {
    win[me] = D->window.as<char>();
    for (int r = 0; r < R && ok; ++r) {
      if (r == me) continue;
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[r].win, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
      D->ipc_open.push_back(p);
      win[r] = (char*)p;
    }
  }
  // agree on the outcome (a rank whose open failed must not leave the others polling its window)
  int flag = ok ? 1 : 0;
  BG_TRY(sb.upload(&flag, sizeof(int), st));
  BG_NCCL(N.AllReduce(sb.p, sb.p, 1, ncclInt32, ncclMin, D->comm, st));
  BG_CUDA(cudaMemcpyAsync(&flag, sb.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  BG_CUDA(cudaStreamSynchronize(st));
  if (!flag) {
    for (void* p : D->ipc_open) cudaIpcCloseMemHandle(p);
    D->ipc_open.clear();
    cudaGetLastError();
    return BG_OK;
  }
  PeerAR par;
  std::memset(&par, 0, sizeof(par));
  par.nranks = R; par.rank = me;
  double tmo = 120.0;
  if (const char* e = getenv("B200GMRES_PEER_TIMEOUT_S")) { const double t = atof(e); if (t >= 1.0) tmo = t; }
  par.timeout_ns = (unsigned long long)(tmo * 1e9);
  par.seq = (unsigned*)(D->window.as<char>() + wl.counters);
  for (int r = 0; r < R; ++r) par.inbox[r] = (LL16*)(win[r] + WindowLayout(R, all[r].n_halo).ar_inbox);
  PeerHalo ph;
  std::memset(&ph, 0, sizeof(ph));
  ph.n_nbr = nn; ph.n_halo = D->n_halo; ph.timeout_ns = par.timeout_ns;
  ph.my_inbox = (LL16*)(D->window.as<char>() + wl.halo_inbox);
  for (int k = 0; k <= nn; ++k) ph.send_ptr[k] = D->send_ptr[k];
  for (int k = 0; k < nn; ++k) {
    const int q = D->nbr[k];
    const PeerRecord& pr = all[q];
    int slot = -1;  // my position in q's neighbour list
    for (int t = 0; t < pr.n_nbr; ++t) if (pr.nbr[t] == me) slot = t;
    if (slot < 0) { set_error("neighbour lists of ranks %d and %d disagree", me, q); return BG_ERR_COMM; }
    if (pr.recv_ptr[slot + 1] - pr.recv_ptr[slot] != D->send_ptr[k + 1] - D->send_ptr[k]) { set_error("halo sizes of ranks %d and %d disagree", me, q); return BG_ERR_COMM; }
    ph.peer_inbox[k] = (LL16*)(win[q] + WindowLayout(R, pr.n_halo).halo_inbox) + pr.recv_ptr[slot];
    ph.peer_half[k] = std::max(pr.n_halo, 1);
  }
  BG_TRY(D->d_par.upload(&par, sizeof(par), st));
  BG_TRY(D->d_ph.upload(&ph, sizeof(ph), st));
  BG_CUDA(cudaStreamSynchronize(st));
  D->p2p_ready = true;
  D->transport = BG_TRANSPORT_P2P;
  { const char* e = getenv("B200GMRES_FUSE_HALO"); D->fuse_halo = !(e && e[0] == '0'); }
  return BG_OK;
}
}  // namespace

extern "C" int bg_dist_create(bg_context_t ctx, bg_dist_plan_t P, const void* id, int dtype, int m, const bg_precond_opts* po,
                              int use_ilut, bg_dist_t* out) {
  if (!ctx || !P || !out || (P->p.nranks > 1 && !id)) return BG_ERR_ARG;
  *out = nullptr;
  BG_CUDA(cudaSetDevice(ctx->device));
  const DistPlan<double>& p = P->p;
  // B200GMRES_SETUP_TIMING=1: wall-clock of the setup stages on stderr (development aid)
  static const bool timing = [] { const char* e = getenv("B200GMRES_SETUP_TIMING"); return e && e[0] == '1'; }();
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t_prev = now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    cudaDeviceSynchronize();
    const double t = now();
    std::fprintf(stderr, "[bg_dist_create rank %d] %-28s %.3f s\n", p.rank, what, t - t_prev);
    t_prev = t;
  };
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
    lap("ncclCommInitRank");
    // local rows: own columns then halo slots
    rc = matrix_create_internal(ctx, p.n_local, p.n_local + p.n_halo, p.A.rp, p.A.ci, p.A.v.data(), BG_F64, dtype, BG_SPMV_AUTO, &D->A);
    if (rc < 0) break;
    lap("local matrix");
    // block-Jacobi ILU of the own diagonal block (no communication in the triangular solves)
    std::vector<int> drp(p.n_local + 1, 0), dci;
    std::vector<double> dv;
    for (int i = 0; i < p.n_local; ++i) {
      for (int q = p.A.rp[i]; q < p.A.rp[i + 1]; ++q)
        if (p.A.ci[q] < p.n_local) { dci.push_back(p.A.ci[q]); dv.push_back(p.A.v[q]); }
      drp[i + 1] = (int)dci.size();
    }
    lap("diagonal block");
    bg_precond_opts o;
    bg_precond_opts_default(&o);
    if (po) o = *po;
    o.trsv_schedule = BG_TRSV_BLOCK;
    rc = use_ilut ? bg_precond_create_ilut(ctx, p.n_local, drp.data(), dci.data(), dv.data(), BG_F64, dtype, &o, &D->M)
                  : bg_precond_create_ilu0(ctx, p.n_local, drp.data(), dci.data(), dv.data(), BG_F64, dtype, &o, &D->M);
    if (rc < 0) break;
    lap("block-Jacobi ILU");
    rc = solver_create_internal(ctx, p.n_local, m, dtype, p.n_halo, &D->S);
    if (rc < 0) break;
    lap("workspace");
    // the solver works in the preconditioner's block-major order: send lists follow
    std::vector<int> inv(p.n_local), sidx(p.send_idx.size());
    for (int i = 0; i < p.n_local; ++i) inv[D->M->fb.h_perm[i]] = i;
    for (size_t k = 0; k < sidx.size(); ++k) sidx[k] = inv[p.send_idx[k]];
    rc = D->send_idx.upload(sidx.data(), sizeof(int) * sidx.size(), ctx->stream);
    if (rc < 0) break;
    rc = D->sendbuf.alloc(dtype_size(dtype) * (sidx.size() + 2));
    if (rc < 0) break;
    D->send_count = (int)sidx.size();
    cudaStreamSynchronize(ctx->stream);
    if (p.nranks > 1) { rc = setup_p2p(D); if (rc < 0) break; }
    lap("send lists + peer windows");
  } while (false);
  if (rc < 0) { bg_dist_destroy(D); return rc; }
  *out = D;
  return BG_OK;
}

extern "C" int bg_dist_destroy(bg_dist_t D) {
  if (!D) return BG_OK;
  if (D->ctx) cudaStreamSynchronize(D->ctx->stream);
  for (void* p : D->ipc_open) cudaIpcCloseMemHandle(p);
  D->ipc_open.clear();
  if (D->S) bg_solver_destroy(D->S);
  if (D->M) bg_precond_destroy(D->M);
  if (D->A) bg_matrix_destroy(D->A);
  if (D->comm && nccl().ok) nccl().CommDestroy(D->comm);
  delete D;
  return BG_OK;
}

extern "C" int bg_dist_gmres_solve(bg_dist_t D, const void* d_b, void* d_x, const bg_gmres_opts* o, bg_gmres_result* r) {
  if (!D || !d_b || !d_x || !o) return BG_ERR_ARG;
  D->solve_index++; D->exchange_in_solve = 0;
  return gmres_solve_dist(D->ctx, D->S, D->A, D->M, d_b, d_x, o, r, D->nranks > 1 ? D : nullptr);  // one rank: plain single-GPU solve
}

extern "C" int bg_dist_profile(bg_dist_t D, const void* d_b, void* d_x, const bg_gmres_opts* o, bg_gmres_result* r, bg_phase_stats* st) {
  if (!D || !d_b || !d_x || !o || !st) return BG_ERR_ARG;
  D->solve_index++; D->exchange_in_solve = 0;
  return gmres_profile_dist(D->ctx, D->S, D->A, D->M, d_b, d_x, o, r, st, D->nranks > 1 ? D : nullptr);
}

extern "C" int bg_dist_bench_comm(bg_dist_t D, int what, int count, int reps, double* us) {
  if (!D || !us || reps < 1 || count < 1 || count > 33) return BG_ERR_ARG;
  bg_solver_t S = D->S;
  bg_context_t ctx = D->ctx;
  GmresDev st = S->dev;
  D->solve_index++; D->exchange_in_solve = 0;
  BG_TRY(launch_state_reset(ctx, st, 0.0, 0.1));
  BG_CUDA(cudaMemsetAsync(S->w.p, 0, (size_t)S->ld * dtype_size(S->dtype), ctx->stream));
  BG_CUDA(cudaMemsetAsync(S->V.p, 0, (size_t)S->ld * dtype_size(S->dtype) * (S->m + 1), ctx->stream));
  double* red = S->red.as<double>();
  auto op = [&]() -> int {
    if (what == 0) return D->exchange(S->V.p, nullptr);
    // an all-reduce of `count` values carried by a dots step over 32 rows (j = count - 1 basis columns)
    const int rows = std::min(32, S->n);
    if (D->peer_ar()) return launch_dots(ctx, S->dtype, rows, S->V.p, S->ld, count - 1, S->w.p, st, FIN_NONE, 0, 0, BG_ORTH_CGS_ADAPTIVE, 0, nullptr, D->peer_ar());
    BG_TRY(launch_dots(ctx, S->dtype, rows, S->V.p, S->ld, count - 1, S->w.p, st, FIN_NONE, 0, 0, BG_ORTH_CGS_ADAPTIVE, 0, red));
    BG_TRY(D->allreduce(red, count));
    return launch_finalize(ctx, st, FIN_NONE, 0, 0, BG_ORTH_CGS_ADAPTIVE, 0, count - 1, red);
  };
  for (int i = 0; i < 20; ++i) BG_TRY(op());
  BG_CUDA(cudaEventRecord(S->t0, ctx->stream));
  for (int i = 0; i < reps; ++i) BG_TRY(op());
  BG_CUDA(cudaEventRecord(S->t1, ctx->stream));
  BG_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, S->t0, S->t1);
  *us = 1e3 * ms / reps;
  return BG_OK;
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
  o->transport = D->transport;
  o->p2p_available = D->p2p_ready ? 1 : 0;
  return BG_OK;
}

extern "C" int bg_dist_set_transport(bg_dist_t D, int transport) {
  if (!D || (transport != BG_TRANSPORT_NCCL && transport != BG_TRANSPORT_P2P)) return BG_ERR_ARG;
  if (transport == BG_TRANSPORT_P2P && D->nranks > 1 && !D->p2p_ready) { set_error("peer-memory transport is not available (CUDA IPC between the ranks failed at creation)"); return BG_ERR_UNSUPPORTED; }
  D->transport = transport;
  return BG_OK;
}
