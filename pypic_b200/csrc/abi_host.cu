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

// ---- whole sheath timestep with host buffers ---------------------------------------
int pic_host_dd_step(const pic_dd_params* p, const double* x0, const double* u0, const double* E0, double tol,
                     int maxiter, double* x1, double* u1, int8_t* active, double* E1, double* j1, int* iters,
                     double* resid) {
    PIC_REQUIRE(p && x0 && u0 && E0 && x1 && u1 && active && E1 && j1 && iters && resid, "dd_step: null pointer");
    PIC_REQUIRE(p->N >= 1 && p->Ng >= 3, "dd_step: bad sizes");
    std::lock_guard<std::mutex> lk(g_ws.mu);
    int rc = g_ws.ensure_stream();
    if (rc) return rc;
    cudaStream_t st = g_ws.stream;
    const int64_t N = p->N;
    const int Ng = p->Ng;
    WS_GET(10, double, N, dx0);
    WS_GET(11, double, N, du0);
    WS_GET(12, double, N, dx1);
    WS_GET(13, double, N, du1);
    WS_GET(14, int8_t, N, dact);
    WS_GET(15, double, 8 * (size_t)Ng + 16, grid);
    WS_GET(3, int, 1, derr);
    double *dE0 = grid, *dEs = grid + Ng, *dE1 = grid + 2 * Ng, *dj1 = grid + 3 * Ng, *dacc = grid + 4 * Ng,
           *dwall = grid + 6 * Ng + 4, *dstats = grid + 6 * Ng + 8;
    PIC_CHECK_CUDA(cudaMemsetAsync(derr, 0, sizeof(int), st));
    PIC_CHECK_CUDA(cudaMemsetAsync(dacc, 0, (size_t)(2 * Ng + 4 + 4 + 4) * 8, st));   // acc, wall_cum, stats
    PIC_CHECK_CUDA(cudaMemsetAsync(dact, 1, (size_t)N, st));
    PIC_CHECK_CUDA(cudaMemcpyAsync(dx0, x0, (size_t)N * 8, cudaMemcpyHostToDevice, st));
    PIC_CHECK_CUDA(cudaMemcpyAsync(du0, u0, (size_t)N * 8, cudaMemcpyHostToDevice, st));
    PIC_CHECK_CUDA(cudaMemcpyAsync(dE0, E0, (size_t)Ng * 8, cudaMemcpyHostToDevice, st));
    PIC_CHECK_CUDA(cudaMemcpyAsync(dEs, dE0, (size_t)Ng * 8, cudaMemcpyDeviceToDevice, st));
    double r = 1.0;
    int k = 0;
    while (r > tol && k < maxiter) {      // PIC_L_DD.py:458
        rc = pic_dev_dd_picard_iter(p, dx0, du0, dx1, du1, dact, dEs, dacc, k == 0, derr, st);
        if (rc) return rc;
        rc = pic_dev_dd_field_update(p, dacc, dwall, dE0, dEs, dE1, dj1, dstats, st);
        if (rc) return rc;
        PIC_CHECK_CUDA(cudaMemcpyAsync(g_ws.pinned, dstats, sizeof(double), cudaMemcpyDeviceToHost, st));
        PIC_CHECK_CUDA(cudaStreamSynchronize(st));
        r = g_ws.pinned[0];
        ++k;
    }
    PIC_CHECK_CUDA(cudaMemcpyAsync(x1, dx1, (size_t)N * 8, cudaMemcpyDeviceToHost, st));
    PIC_CHECK_CUDA(cudaMemcpyAsync(u1, du1, (size_t)N * 8, cudaMemcpyDeviceToHost, st));
    PIC_CHECK_CUDA(cudaMemcpyAsync(active, dact, (size_t)N, cudaMemcpyDeviceToHost, st));
    PIC_CHECK_CUDA(cudaMemcpyAsync(E1, dE1, (size_t)Ng * 8, cudaMemcpyDeviceToHost, st));
    PIC_CHECK_CUDA(cudaMemcpyAsync(j1, dj1, (size_t)Ng * 8, cudaMemcpyDeviceToHost, st));
    *iters = k;
    *resid = r;
    return check_range(derr, st, "dd_step");
}

int pic_host_release(void) {
    std::lock_guard<std::mutex> lk(g_ws.mu);
    for (auto& s : g_ws.slots)
        if (s.first) cudaFree(s.first);
    g_ws.slots.clear();
    return PIC_OK;
}

}  // extern "C"
