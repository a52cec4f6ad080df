// Host-buffer entry points of the C ABI (pic_host_*): what a binding of the reference
// would call with NumPy arrays.  Each uploads, launches the pic_dev_* kernels on a private
// stream, downloads and synchronises.  Device workspaces are cached between calls.
#include <mutex>
#include <vector>
#include "host_common.h"

namespace {

struct Workspace {
    std::mutex mu;
    cudaStream_t stream = nullptr;
    std::vector<std::pair<void*, size_t>> slots;   // grow-only device buffers by slot id
    double* pinned = nullptr;                        // 8 doubles of pinned host memory

    int ensure_stream() {
        if (!stream) PIC_CHECK_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        if (!pinned) PIC_CHECK_CUDA(cudaHostAlloc((void**)&pinned, 8 * sizeof(double), cudaHostAllocDefault));
        return PIC_OK;
    }
    int get(size_t slot, size_t bytes, void** out) {
        if (slots.size() <= slot) slots.resize(slot + 1, {nullptr, 0});
        if (slots[slot].second < bytes) {
            if (slots[slot].first) PIC_CHECK_CUDA(cudaFree(slots[slot].first));
            slots[slot] = {nullptr, 0};
            void* p = nullptr;
            PIC_CHECK_CUDA(cudaMalloc(&p, bytes ? bytes : 8));
            slots[slot] = {p, bytes};
        }
        *out = slots[slot].first;
        return PIC_OK;
    }
};
Workspace g_ws;

#define WS_GET(slot, type, count, var)                                                     \
    type* var = nullptr;                                                                   \
    do {                                                                                   \
        int _rc = g_ws.get(slot, (size_t)(count) * sizeof(type), (void**)&var);            \
        if (_rc) return _rc;                                                               \
    } while (0)

int check_range(int* d_err, cudaStream_t st, const char* what) {
    int h = 0;
    PIC_CHECK_CUDA(cudaMemcpyAsync(&h, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    PIC_CHECK_CUDA(cudaStreamSynchronize(st));
    if (h) {
        pic::set_error("%s: %d particle position(s) outside the grid (reference behaviour undefined); indices were clamped",
                       what, h);
        return PIC_ERR_RANGE;
    }
    return PIC_OK;
}

}  // namespace

extern "C" {

// small utilities used by the Python host (kept here so it needs no torch calls per iteration)
int pic_dev_read(const void* dev, void* host, int64_t bytes, void* stream) {
    PIC_REQUIRE(dev && host && bytes >= 0, "dev_read: bad argument");
    PIC_CHECK_CUDA(cudaMemcpyAsync(host, dev, (size_t)bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    PIC_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return PIC_OK;
}
int pic_dev_write(void* dev, const void* host, int64_t bytes, void* stream) {
    PIC_REQUIRE(dev && host && bytes >= 0, "dev_write: bad argument");
    PIC_CHECK_CUDA(cudaMemcpyAsync(dev, host, (size_t)bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return PIC_OK;
}
int pic_dev_zero(void* dev, int64_t bytes, void* stream) {
    PIC_REQUIRE(dev && bytes >= 0, "dev_zero: bad argument");
    PIC_CHECK_CUDA(cudaMemsetAsync(dev, 0, (size_t)bytes, (cudaStream_t)stream));
    return PIC_OK;
}
// start of a sheath timestep: Es = E0 and the per-step accumulators / statistics / loop flag cleared, one call
int pic_dev_dd_step_begin(double* Es, const double* E0, int Ng, double* wall_cum, double* stats, int64_t nstats,
                          int32_t* ctl, void* stream) {
    PIC_REQUIRE(Es && E0 && Ng > 0 && wall_cum && stats && nstats > 0 && ctl, "dd_step_begin: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    PIC_CHECK_CUDA(cudaMemcpyAsync(Es, E0, (size_t)Ng * sizeof(double), cudaMemcpyDeviceToDevice, st));
    PIC_CHECK_CUDA(cudaMemsetAsync(wall_cum, 0, 4 * sizeof(double), st));
    PIC_CHECK_CUDA(cudaMemsetAsync(stats, 0, (size_t)nstats * sizeof(double), st));
    PIC_CHECK_CUDA(cudaMemsetAsync(ctl, 0, sizeof(int32_t), st));
    return PIC_OK;
}
int pic_dev_copy(void* dst, const void* src, int64_t bytes, void* stream) {
    PIC_REQUIRE(dst && src && bytes >= 0, "dev_copy: bad argument");
    PIC_CHECK_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return PIC_OK;
}
int pic_stream_sync(void* stream) {
    PIC_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return PIC_OK;
}

// ---- pypic function-level drop-ins ------------------------------------------------
int pic_host_pypic_interpolate_p(const double* F, const double* x, int Ng, int64_t N, double dx, double* out) {
    PIC_REQUIRE(F && (x || N == 0) && (out || N == 0) && Ng >= 2 && N >= 0, "interpolate_p: bad argument");
    std::lock_guard<std::mutex> lk(g_ws.mu);
    int rc = g_ws.ensure_stream();
    if (rc) return rc;
    cudaStream_t st = g_ws.stream;
    WS_GET(0, double, Ng, dF);
    WS_GET(1, double, N, dx_);
    WS_GET(2, double, N, dout);
    WS_GET(3, int, 1, derr);
    PIC_CHECK_CUDA(cudaMemsetAsync(derr, 0, sizeof(int), st));
    PIC_CHECK_CUDA(cudaMemcpyAsync(dF, F, (size_t)Ng * 8, cudaMemcpyHostToDevice, st));
    PIC_CHECK_CUDA(cudaMemcpyAsync(dx_, x, (size_t)N * 8, cudaMemcpyHostToDevice, st));
    rc = pic_dev_pypic_interpolate(dF, dx_, dout, N, Ng, dx, derr, st);
    if (rc) return rc;
    PIC_CHECK_CUDA(cudaMemcpyAsync(out, dout, (size_t)N * 8, cudaMemcpyDeviceToHost, st));
    return check_range(derr, st, "interpolate_p");
}

static int host_pypic_weight(const double* x, const double* q, const double* v, int p2c, int Ng, int64_t N, double dx,
                             double* out) {
    PIC_REQUIRE((x || N == 0) && (q || N == 0) && out && Ng >= 2 && N >= 0, "weight_*_p: bad argument");
    std::lock_guard<std::mutex> lk(g_ws.mu);
    int rc = g_ws.ensure_stream();
    if (rc) return rc;
    cudaStream_t st = g_ws.stream;
    WS_GET(0, double, Ng, dacc);
    WS_GET(1, double, N, dx_);
    WS_GET(2, double, N, dq);
    WS_GET(4, double, N, dv);
    WS_GET(3, int, 1, derr);
    PIC_CHECK_CUDA(cudaMemsetAsync(derr, 0, sizeof(int), st));
    PIC_CHECK_CUDA(cudaMemsetAsync(dacc, 0, (size_t)Ng * 8, st));
    PIC_CHECK_CUDA(cudaMemcpyAsync(dx_, x, (size_t)N * 8, cudaMemcpyHostToDevice, st));
    PIC_CHECK_CUDA(cudaMemcpyAsync(dq, q, (size_t)N * 8, cudaMemcpyHostToDevice, st));
    if (v) PIC_CHECK_CUDA(cudaMemcpyAsync(dv, v, (size_t)N * 8, cudaMemcpyHostToDevice, st));
    rc = pic_dev_pypic_weight(dx_, dq, v ? dv : nullptr, dacc, N, Ng, dx, (double)p2c, derr, st);
    if (rc) return rc;
    PIC_CHECK_CUDA(cudaMemcpyAsync(out, dacc, (size_t)Ng * 8, cudaMemcpyDeviceToHost, st));
    return check_range(derr, st, "weight_*_p");
}
int pic_host_pypic_weight_current_p(const double* x, const double* q, const double* v, int p2c, int Ng, int64_t N,
                                    double dx, double* j) {
    PIC_REQUIRE(v || N == 0, "weight_current_p: v is null");
    static const double dummy = 0.0;
    return host_pypic_weight(x, q, v ? v : &dummy, p2c, Ng, N, dx, j);
}
int pic_host_pypic_weight_density_p(const double* x, const double* q, int p2c, int Ng, int64_t N, double dx,
                                    double* rho) {
    return host_pypic_weight(x, q, nullptr, p2c, Ng, N, dx, rho);
}

// ---- PIC_L_DD function-level drop-ins -----------------------------------------------
int pic_host_dd_interpolateField(const double* F, const double* x, int Ng, int64_t N, double dx, double* out) {
    PIC_REQUIRE(F && (x || N == 0) && (out || N == 0) && Ng >= 2 && N >= 0, "interpolateField: bad argument");
    std::lock_guard<std::mutex> lk(g_ws.mu);
    int rc = g_ws.ensure_stream();
    if (rc) return rc;
    cudaStream_t st = g_ws.stream;
    WS_GET(0, double, Ng, dF);
    WS_GET(1, double, N, dx_);
    WS_GET(2, double, N, dout);
    WS_GET(3, int, 1, derr);
    PIC_CHECK_CUDA(cudaMemsetAsync(derr, 0, sizeof(int), st));
    PIC_CHECK_CUDA(cudaMemcpyAsync(dF, F, (size_t)Ng * 8, cudaMemcpyHostToDevice, st));
    PIC_CHECK_CUDA(cudaMemcpyAsync(dx_, x, (size_t)N * 8, cudaMemcpyHostToDevice, st));
    rc = pic_dev_dd_interpolate(dF, dx_, dout, N, Ng, dx, derr, st);
    if (rc) return rc;
    PIC_CHECK_CUDA(cudaMemcpyAsync(out, dout, (size_t)N * 8, cudaMemcpyDeviceToHost, st));
    return check_range(derr, st, "interpolateField");
}

static int host_dd_weight(const double* x, const double* q, const double* v, double p2c, int Ng, int64_t N, double dx,
                          double dt, const double* active, double* out) {
    PIC_REQUIRE((x || N == 0) && (q || N == 0) && (active || N == 0) && out && Ng >= 3 && N >= 0, "weight*: bad argument");
    std::lock_guard<std::mutex> lk(g_ws.mu);
    int rc = g_ws.ensure_stream();
    if (rc) return rc;
    cudaStream_t st = g_ws.stream;
    WS_GET(0, double, Ng, dout);
    WS_GET(1, double, N, dx_);
    WS_GET(2, double, N, dq);
    WS_GET(4, double, N, dv);
    WS_GET(5, double, N, da);
    WS_GET(3, int, 1, derr);
    PIC_CHECK_CUDA(cudaMemsetAsync(derr, 0, sizeof(int), st));
    PIC_CHECK_CUDA(cudaMemcpyAsync(dx_, x, (size_t)N * 8, cudaMemcpyHostToDevice, st));
    PIC_CHECK_CUDA(cudaMemcpyAsync(dq, q, (size_t)N * 8, cudaMemcpyHostToDevice, st));
    PIC_CHECK_CUDA(cudaMemcpyAsync(da, active, (size_t)N * 8, cudaMemcpyHostToDevice, st));
    if (v) PIC_CHECK_CUDA(cudaMemcpyAsync(dv, v, (size_t)N * 8, cudaMemcpyHostToDevice, st));
    rc = pic_dev_dd_weight(dx_, dq, v ? dv : nullptr, da, dout, N, Ng, dx, dt, p2c, derr, st);
    if (rc) return rc;
    PIC_CHECK_CUDA(cudaMemcpyAsync(out, dout, (size_t)Ng * 8, cudaMemcpyDeviceToHost, st));
    return check_range(derr, st, "weight*");
}
int pic_host_dd_weightCurrents(const double* x, const double* q, const double* v, double p2c, int Ng, int64_t N,
                               double dx, double dt, const double* active, double* j) {
    PIC_REQUIRE(v || N == 0, "weightCurrents: v is null");
    static const double dummy = 0.0;
    return host_dd_weight(x, q, v ? v : &dummy, p2c, Ng, N, dx, dt, active, j);
}
int pic_host_dd_weightDensities(const double* x, const double* q, double p2c, int Ng, int64_t N, double dx,
                                const double* active, double* rho) {
    return host_dd_weight(x, q, nullptr, p2c, Ng, N, dx, 1.0, active, rho);
}

// ---- whole sheath timesteps with host buffers ------------------------------------------
// Pipeline of independent batches: while batch b runs its Picard loop on the compute stream,
// the inputs of batch b+1 travel host->device and the results of batch b-1 travel device->host
// on two more streams (two device slots).  The residual of every Picard iteration reaches the
// host through MAPPED pinned memory written by a one-thread kernel, not through a copy-engine
// transfer that would queue behind the bulk copies.
namespace {
__global__ void publish_resid_k(const double* __restrict__ stats, volatile double* host_out) { host_out[0] = stats[0]; }

struct Pipe {
    cudaStream_t h2d = nullptr, cmp = nullptr, d2h = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_cmp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    double* resid_host = nullptr;   // mapped pinned
    int* err_host = nullptr;        // pinned int[2]
    int init() {
        if (h2d) return PIC_OK;
        PIC_CHECK_CUDA(cudaStreamCreateWithFlags(&h2d, cudaStreamNonBlocking));
        PIC_CHECK_CUDA(cudaStreamCreateWithFlags(&cmp, cudaStreamNonBlocking));
        PIC_CHECK_CUDA(cudaStreamCreateWithFlags(&d2h, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            PIC_CHECK_CUDA(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
            PIC_CHECK_CUDA(cudaEventCreateWithFlags(&ev_cmp[i], cudaEventDisableTiming));
            PIC_CHECK_CUDA(cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming));
        }
        PIC_CHECK_CUDA(cudaHostAlloc((void**)&resid_host, 8 * sizeof(double), cudaHostAllocMapped));
        PIC_CHECK_CUDA(cudaHostAlloc((void**)&err_host, 2 * sizeof(int), cudaHostAllocDefault));
        return PIC_OK;
    }
};
Pipe g_pipe;
}  // namespace

int pic_host_dd_step_batches(const pic_dd_params* p, int nbatch, const double* const* x0, const double* const* u0,
                             const double* const* E0, double tol, int maxiter, double* const* x1, double* const* u1,
                             int8_t* const* active, double* const* E1, double* const* j1, int* iters, double* resid) {
    PIC_REQUIRE(p && x0 && u0 && E0 && x1 && u1 && active && E1 && j1 && iters && resid, "dd_step_batches: null pointer");
    PIC_REQUIRE(nbatch >= 1 && p->N >= 1 && p->Ng >= 3, "dd_step_batches: bad sizes");
    PIC_REQUIRE(!(p->flags & 128), "dd_step_batches: the reproducible build (flags bit7) is served by the device-resident driver");
    for (int b = 0; b < nbatch; ++b)
        PIC_REQUIRE(x0[b] && u0[b] && E0[b] && x1[b] && u1[b] && active[b] && E1[b] && j1[b], "dd_step_batches: null batch pointer");
    std::lock_guard<std::mutex> lk(g_ws.mu);
    int rc = g_pipe.init();
    if (rc) return rc;
    Pipe& q = g_pipe;
    const int64_t N = p->N;
    const int Ng = p->Ng;
    const int nslot = nbatch > 1 ? 2 : 1;
    double *dx0[2], *du0[2], *dx1[2], *du1[2], *grid[2];
    int8_t* dact[2];
    int* derr[2];
    for (int s = 0; s < nslot; ++s) {
        const size_t base = 10 + 8 * (size_t)s;
        if ((rc = g_ws.get(base + 0, (size_t)N * 8, (void**)&dx0[s]))) return rc;
        if ((rc = g_ws.get(base + 1, (size_t)N * 8, (void**)&du0[s]))) return rc;
        if ((rc = g_ws.get(base + 2, (size_t)N * 8, (void**)&dx1[s]))) return rc;
        if ((rc = g_ws.get(base + 3, (size_t)N * 8, (void**)&du1[s]))) return rc;
        if ((rc = g_ws.get(base + 4, (size_t)N, (void**)&dact[s]))) return rc;
        if ((rc = g_ws.get(base + 5, (8 * (size_t)Ng + 24) * 8, (void**)&grid[s]))) return rc;
        if ((rc = g_ws.get(base + 6, sizeof(int), (void**)&derr[s]))) return rc;
    }
    double* resid_dev = nullptr;
    PIC_CHECK_CUDA(cudaHostGetDevicePointer((void**)&resid_dev, q.resid_host, 0));
    auto upload = [&](int b) -> int {
        const int s = b & (nslot - 1);
        // the slot's inputs were last read by the compute of batch b-2, which the host has waited for
        PIC_CHECK_CUDA(cudaMemcpyAsync(dx0[s], x0[b], (size_t)N * 8, cudaMemcpyHostToDevice, q.h2d));
        PIC_CHECK_CUDA(cudaMemcpyAsync(du0[s], u0[b], (size_t)N * 8, cudaMemcpyHostToDevice, q.h2d));
        PIC_CHECK_CUDA(cudaMemcpyAsync(grid[s], E0[b], (size_t)Ng * 8, cudaMemcpyHostToDevice, q.h2d));
        PIC_CHECK_CUDA(cudaEventRecord(q.ev_in[s], q.h2d));
        return PIC_OK;
    };
    q.err_host[0] = q.err_host[1] = 0;
    long long bad_total = 0;        // range errors of every batch (the two pinned words are reused per slot)
    if ((rc = upload(0))) return rc;
    for (int b = 0; b < nbatch; ++b) {
        const int s = b & (nslot - 1);
        if (b + 1 < nbatch && (rc = upload(b + 1))) return rc;
        double *dE0 = grid[s], *dEs = grid[s] + Ng, *dE1 = grid[s] + 2 * Ng, *dj1 = grid[s] + 3 * Ng,
               *dacc = grid[s] + 4 * Ng, *dwall = grid[s] + 6 * Ng + 4, *dstats = grid[s] + 6 * Ng + 8;
        PIC_CHECK_CUDA(cudaStreamWaitEvent(q.cmp, q.ev_in[s], 0));
        if (b >= 2) PIC_CHECK_CUDA(cudaStreamWaitEvent(q.cmp, q.ev_out[s], 0));   // results of b-2 have left the slot
        PIC_CHECK_CUDA(cudaMemsetAsync(derr[s], 0, sizeof(int), q.cmp));
        PIC_CHECK_CUDA(cudaMemsetAsync(dacc, 0, (size_t)(2 * Ng + 4 + 4 + 8) * 8, q.cmp));   // acc, wall_cum, stats[8]
        PIC_CHECK_CUDA(cudaMemsetAsync(dact[s], 1, (size_t)N, q.cmp));
        PIC_CHECK_CUDA(cudaMemcpyAsync(dEs, dE0, (size_t)Ng * 8, cudaMemcpyDeviceToDevice, q.cmp));
        double r = 1.0;
        int k = 0;
        while (r > tol && k < maxiter) {      // PIC_L_DD.py:458
            rc = pic_dev_dd_picard_iter(p, dx0[s], du0[s], dx1[s], du1[s], dact[s], dEs, dacc, k == 0, derr[s], q.cmp);
            if (rc) return rc;
            rc = pic_dev_dd_field_update(p, dacc, dwall, dE0, dEs, dE1, dj1, dstats, q.cmp);
            if (rc) return rc;
            publish_resid_k<<<1, 1, 0, q.cmp>>>(dstats, resid_dev);
            PIC_CHECK_LAUNCH();
            PIC_CHECK_CUDA(cudaStreamSynchronize(q.cmp));
            r = *(volatile double*)q.resid_host;
            ++k;
        }
        iters[b] = k;
        resid[b] = r;
        PIC_CHECK_CUDA(cudaEventRecord(q.ev_cmp[s], q.cmp));
        PIC_CHECK_CUDA(cudaStreamWaitEvent(q.d2h, q.ev_cmp[s], 0));
        PIC_CHECK_CUDA(cudaMemcpyAsync(x1[b], dx1[s], (size_t)N * 8, cudaMemcpyDeviceToHost, q.d2h));
        PIC_CHECK_CUDA(cudaMemcpyAsync(u1[b], du1[s], (size_t)N * 8, cudaMemcpyDeviceToHost, q.d2h));
        PIC_CHECK_CUDA(cudaMemcpyAsync(active[b], dact[s], (size_t)N, cudaMemcpyDeviceToHost, q.d2h));
        PIC_CHECK_CUDA(cudaMemcpyAsync(E1[b], dE1, (size_t)Ng * 8, cudaMemcpyDeviceToHost, q.d2h));
        PIC_CHECK_CUDA(cudaMemcpyAsync(j1[b], dj1, (size_t)Ng * 8, cudaMemcpyDeviceToHost, q.d2h));
        PIC_CHECK_CUDA(cudaMemcpyAsync(&q.err_host[s], derr[s], sizeof(int), cudaMemcpyDeviceToHost, q.d2h));
        PIC_CHECK_CUDA(cudaEventRecord(q.ev_out[s], q.d2h));
        if (b >= 1) {   // the previous batch's results (other slot) are complete: take its error word before the
                        // slot's next batch (b+1) overwrites it
            const int sp = (b - 1) & (nslot - 1);
            PIC_CHECK_CUDA(cudaEventSynchronize(q.ev_out[sp]));
            bad_total += q.err_host[sp];
            q.err_host[sp] = 0;
        }
    }
    PIC_CHECK_CUDA(cudaStreamSynchronize(q.d2h));
    bad_total += q.err_host[(nbatch - 1) & (nslot - 1)];
    const long long bad = bad_total;
    if (bad) {
        pic::set_error("dd_step: %lld particle position(s) outside the grid (reference behaviour undefined); indices were clamped", bad);
        return PIC_ERR_RANGE;
    }
    return PIC_OK;
}

int pic_host_dd_step(const pic_dd_params* p, const double* x0, const double* u0, const double* E0, double tol,
                     int maxiter, double* x1, double* u1, int8_t* active, double* E1, double* j1, int* iters,
                     double* resid) {
    PIC_REQUIRE(x0 && u0 && E0 && x1 && u1 && active && E1 && j1, "dd_step: null pointer");
    return pic_host_dd_step_batches(p, 1, &x0, &u0, &E0, tol, maxiter, &x1, &u1, &active, &E1, &j1, iters, resid);
}

int pic_host_release(void) {
    std::lock_guard<std::mutex> lk(g_ws.mu);
    for (auto& s : g_ws.slots)
        if (s.first) cudaFree(s.first);
    g_ws.slots.clear();
    return PIC_OK;
}

}  // extern "C"
