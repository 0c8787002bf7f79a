// transient.cu -- the backward-Euler time loop around the solver, resident on the device.
//
// Replaces the host loop of the power-grid driver (src/mna_solve_gpu_gmres.cpp:564-647):
//     u = interp1(source tables, t_i)             (:574-583, CPU)      -> pwl_eval_kernel
//     w = B u ; w += (C/h) x_{i-1}                (:585-592, cs_dl_gaxpy on the CPU) -> two SpMVs on the device
//     H2D rhs ; GMRES_dev_PG() ; D2H x            (:615-619, gmres_interface_pg.cu:118-129) -> bg_gmres_solve on
//                                                 device vectors, warm-started from x_{i-1} (x is never reset, :544)
//     sim_port_value(j, i) = x(port(j))           (:628-629)           -> probe_kernel into a device table
// The ILU factor inside `M` is built once by the caller and reused by every step, as the driver does
// (:518-523).  No vector crosses the PCIe bus inside the loop; the host reads one status per step (the
// solver's own convergence poll) and the probe table once at the end.
#include <algorithm>
#include <cstring>

#include "internal.h"

using namespace bg;

struct bg_transient_s {
  bg_context_t ctx = nullptr;
  bg_solver_t S = nullptr;
  bg_matrix_t A = nullptr, B = nullptr, Ch = nullptr;
  bg_precond_t M = nullptr;
  int n = 0, nsrc = 0, dtype = BG_F64;
  DevBuf pwl_ptr, pwl_t, pwl_v, u, bu, rhs, probe_idx, probe_out;
  DevBuf pulse_of, pulse_t, pulse_v;  // per source: index of its PULSE record or -1 (piecewise linear); records: 5 times, 2 values
  bool has_pulse = false;
  int nprobe = 0, probe_cap = 0;
};

namespace {
// piecewise-linear sources, held at the end values outside the table (interp1, src/interp.cpp; gen_PWLut*, src/kernels.cu:154-230)
// PULSE sources (gen_PULSEut_kernel, src/kernels.cu:224-246): record (td, tr, tf, tw, tp), (vlo, vhi); periodic in tp
template <class VT>
__global__ void pwl_eval_kernel(int nsrc, const int* __restrict__ ptr, const double* __restrict__ tt, const double* __restrict__ vv,
                                const int* __restrict__ pulse_of, const double* __restrict__ pt, const double* __restrict__ pv,
                                double t, VT* __restrict__ u) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nsrc) return;
  const int pr = pulse_of ? pulse_of[k] : -1;
  if (pr >= 0) {
    const double td = pt[5 * pr], tr = pt[5 * pr + 1], tf = pt[5 * pr + 2], tw = pt[5 * pr + 3], tp = pt[5 * pr + 4];
    const double vlo = pv[2 * pr], vhi = pv[2 * pr + 1];
    double mt = t - floor(t / tp) * tp, value;
    if (mt < td) value = vlo;
    else if (mt < td + tr) value = vlo + (mt - td) * (vhi - vlo) / tr;
    else if (mt < td + tr + tw) value = vhi;
    else if (mt < td + tr + tw + tf) value = vhi - (mt - td - tr - tw) * (vhi - vlo) / tf;
    else value = vlo;
    u[k] = (VT)value;
    return;
  }
  const int lb = ptr[k], ub = ptr[k + 1];
  double val = 0.0;
  if (ub > lb) {
    if (t <= tt[lb]) val = vv[lb];
    else if (t >= tt[ub - 1]) val = vv[ub - 1];
    else {
      int lo = lb, hi = ub - 1;  // tt[lo] < t < tt[hi]
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (tt[mid] <= t) lo = mid; else hi = mid; }
      val = vv[lo] + (vv[hi] - vv[lo]) * (t - tt[lo]) / (tt[hi] - tt[lo]);
    }
  }
  u[k] = (VT)val;
}
template <class VT>
__global__ void probe_kernel(int nprobe, const int* __restrict__ idx, const VT* __restrict__ x, double* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nprobe) out[k] = (double)x[idx[k]];
}
}  // namespace

extern "C" int bg_transient_create(bg_context_t ctx, bg_solver_t S, bg_matrix_t A, bg_precond_t M, int n, const int* c_rowptr,
                                   const int* c_col, const void* c_val, int nsrc, const int* b_rowptr, const int* b_col,
                                   const void* b_val, int val_dtype, const int* pwl_ptr, const double* pwl_t, const double* pwl_v,
                                   int nprobe, const int* probe_idx, bg_transient_t* out) {
  if (!ctx || !S || !A || !M || !out || n < 0 || nsrc < 0 || !c_rowptr || !b_rowptr || !pwl_ptr || (nprobe > 0 && !probe_idx)) return BG_ERR_ARG;
  if (A->n != n || S->n != n || M->n != n) { set_error("bg_transient_create: A, M and the workspace must all have n = %d rows", n); return BG_ERR_ARG; }
  *out = nullptr;
  BG_CUDA(cudaSetDevice(ctx->device));
  bg_transient_t T = new bg_transient_s;
  T->ctx = ctx; T->S = S; T->A = A; T->M = M; T->n = n; T->nsrc = nsrc; T->dtype = S->dtype; T->nprobe = nprobe;
  int rc = BG_OK;
  do {
    if ((rc = prepare_solve(A, M)) < 0) break;  // keep one-time setup out of the time loop
    // C/h: square, any format; B: n x nsrc, CSR only (sliced-ELL padding points at the own row, which may exceed nsrc)
    std::vector<int> rp(c_rowptr, c_rowptr + n + 1), ci(c_col, c_col + c_rowptr[n]);
    rc = matrix_create_internal(ctx, n, n, rp, ci, c_val, val_dtype, T->dtype, BG_SPMV_AUTO, &T->Ch);
    if (rc < 0) break;
    rp.assign(b_rowptr, b_rowptr + n + 1); ci.assign(b_col, b_col + b_rowptr[n]);
    rc = matrix_create_internal(ctx, n, std::max(nsrc, 1), rp, ci, b_val, val_dtype, T->dtype, BG_SPMV_CSR_VECTOR, &T->B);
    if (rc < 0) break;
    const size_t es = dtype_size(T->dtype);
    const int npts = pwl_ptr[nsrc];
    for (int k = 0; k < nsrc; ++k)
      for (int p = pwl_ptr[k] + 1; p < pwl_ptr[k + 1]; ++p)
        // repeated times are jumps (the reference's PULSE tables repeat the period boundary, src/parser.cpp:2503-2514);
        // the lookup below lands on the later point of a pair and never divides by a zero interval
        if (!(pwl_t[p] >= pwl_t[p - 1])) { set_error("source %d: breakpoint times must not decrease", k); rc = BG_ERR_ARG; break; }
    if (rc < 0) break;
    cudaStream_t s = ctx->stream;
    if ((rc = T->pwl_ptr.upload(pwl_ptr, sizeof(int) * (nsrc + 1), s)) < 0) break;
    if ((rc = T->pwl_t.upload(pwl_t, sizeof(double) * npts, s)) < 0) break;
    if ((rc = T->pwl_v.upload(pwl_v, sizeof(double) * npts, s)) < 0) break;
    if ((rc = T->u.alloc(es * std::max(nsrc, 1))) < 0) break;
    if ((rc = T->rhs.alloc(es * (size_t)S->ld)) < 0) break;
    if ((rc = T->bu.alloc(es * (size_t)S->ld)) < 0) break;
    if (nprobe > 0 && (rc = T->probe_idx.upload(probe_idx, sizeof(int) * nprobe, s)) < 0) break;
    for (int k = 0; k < nprobe; ++k) if (probe_idx[k] < 0 || probe_idx[k] >= n) { set_error("probe index out of range"); rc = BG_ERR_ARG; break; }
    if (rc < 0) break;
    cudaStreamSynchronize(s);
  } while (false);
  if (rc < 0) { bg_transient_destroy(T); return rc; }
  *out = T;
  return BG_OK;
}

extern "C" int bg_transient_set_pulse(bg_transient_t T, int npulse, const int* src_index, const double* pulse_time, const double* pulse_val) {
  if (!T || npulse < 0 || (npulse > 0 && (!src_index || !pulse_time || !pulse_val))) return BG_ERR_ARG;
  BG_CUDA(cudaSetDevice(T->ctx->device));
  std::vector<int> of(std::max(T->nsrc, 1), -1);
  for (int k = 0; k < npulse; ++k) {
    if (src_index[k] < 0 || src_index[k] >= T->nsrc) { set_error("bg_transient_set_pulse: source %d out of range", src_index[k]); return BG_ERR_ARG; }
    const double* t = pulse_time + 5 * k;
    if (!(t[4] > 0) || !(t[1] > 0) || !(t[2] > 0) || t[0] < 0 || t[3] < 0) { set_error("bg_transient_set_pulse: record %d needs tp, tr, tf > 0 and td, tw >= 0", k); return BG_ERR_ARG; }
    of[src_index[k]] = k;
  }
  cudaStream_t s = T->ctx->stream;
  BG_TRY(T->pulse_of.upload(of.data(), sizeof(int) * of.size(), s));
  BG_TRY(T->pulse_t.upload(pulse_time, sizeof(double) * 5 * (size_t)std::max(npulse, 1), s));
  BG_TRY(T->pulse_v.upload(pulse_val, sizeof(double) * 2 * (size_t)std::max(npulse, 1), s));
  BG_CUDA(cudaStreamSynchronize(s));
  T->has_pulse = npulse > 0;
  return BG_OK;
}

extern "C" int bg_transient_destroy(bg_transient_t T) {
  if (!T) return BG_OK;
  if (T->ctx) cudaStreamSynchronize(T->ctx->stream);
  bg_matrix_destroy(T->B);
  bg_matrix_destroy(T->Ch);
  delete T;
  return BG_OK;
}

extern "C" int bg_transient_run(bg_transient_t T, void* d_x, double t0, double h, int nsteps, const bg_gmres_opts* o,
                                bg_transient_stats* st, double* h_probe) {
  if (!T || !d_x || !o || nsteps < 0 || !(h > 0)) return BG_ERR_ARG;
  bg_context_t ctx = T->ctx;
  BG_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  if (T->nprobe > 0 && h_probe && T->probe_cap < nsteps) {
    BG_TRY(T->probe_out.alloc(sizeof(double) * (size_t)T->nprobe * std::max(nsteps, 1)));
    T->probe_cap = nsteps;
  }
  bg_transient_stats acc;
  std::memset(&acc, 0, sizeof(acc));
  cudaEvent_t e0, e1;
  BG_CUDA(cudaEventCreate(&e0));
  BG_CUDA(cudaEventCreate(&e1));
  BG_CUDA(cudaEventRecord(e0, s));
  int rc = BG_OK;
  for (int i = 1; i <= nsteps; ++i) {
    const double t = t0 + h * i;
    if (T->nsrc > 0) {
      const int g = (T->nsrc + 255) / 256;
      const int* pof = T->has_pulse ? T->pulse_of.as<int>() : nullptr;
      if (T->dtype == BG_F32) pwl_eval_kernel<float><<<g, 256, 0, s>>>(T->nsrc, T->pwl_ptr.as<int>(), T->pwl_t.as<double>(), T->pwl_v.as<double>(), pof, T->pulse_t.as<double>(), T->pulse_v.as<double>(), t, T->u.as<float>());
      else pwl_eval_kernel<double><<<g, 256, 0, s>>>(T->nsrc, T->pwl_ptr.as<int>(), T->pwl_t.as<double>(), T->pwl_v.as<double>(), pof, T->pulse_t.as<double>(), T->pulse_v.as<double>(), t, T->u.as<double>());
      ctx->launches++;
    } else {
      BG_CUDA(cudaMemsetAsync(T->u.p, 0, T->u.bytes, s));
    }
    // bu = B u ; rhs = (C/h) x_{i-1} + bu
    if ((rc = launch_spmv(ctx, T->B, 1.0, T->u.p, 0.0, nullptr, T->bu.p, false, nullptr)) < 0) break;
    if ((rc = launch_spmv(ctx, T->Ch, 1.0, d_x, 1.0, T->bu.p, T->rhs.p, true, nullptr)) < 0) break;
    bg_gmres_result r;
    rc = bg_gmres_solve(ctx, T->S, T->A, T->M, T->rhs.p, d_x, o, &r);  // d_x holds x_{i-1}: warm start
    if (rc < 0) break;
    acc.iters_total += r.iters_done;
    acc.iters_max = std::max(acc.iters_max, r.iters_done);
    acc.launches += r.launches;
    if (rc == BG_NOT_CONVERGED) acc.not_converged++;
    acc.last_resid = r.resid;
    if (T->nprobe > 0 && h_probe) {
      const int g = (T->nprobe + 255) / 256;
      double* dst = T->probe_out.as<double>() + (size_t)(i - 1) * T->nprobe;
      if (T->dtype == BG_F32) probe_kernel<float><<<g, 256, 0, s>>>(T->nprobe, T->probe_idx.as<int>(), (const float*)d_x, dst);
      else probe_kernel<double><<<g, 256, 0, s>>>(T->nprobe, T->probe_idx.as<int>(), (const double*)d_x, dst);
      ctx->launches++;
    }
    rc = BG_OK;
  }
  cudaEventRecord(e1, s);
  cudaStreamSynchronize(s);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (rc < 0) return rc;
  if (T->nprobe > 0 && h_probe && nsteps > 0) BG_CUDA(cudaMemcpy(h_probe, T->probe_out.p, sizeof(double) * (size_t)T->nprobe * nsteps, cudaMemcpyDeviceToHost));
  acc.steps = nsteps;
  acc.device_ms = ms;
  if (st) *st = acc;
  return acc.not_converged ? BG_NOT_CONVERGED : BG_OK;
}
