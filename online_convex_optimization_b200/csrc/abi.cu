// abi.cu -- the extern "C" boundary declared in include/oco_b200.h.
// Validates arguments, owns scratch (stream-ordered allocations), picks the kernel family
// and, for the *_host entry points, runs the chunked double-buffered host<->device pipeline.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/oco_b200.h"
#include "oco_internal.h"

using namespace oco;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}
#define CU(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) \
    return fail(OCO_ERR_RUNTIME, "%s failed: %s", #expr, cudaGetErrorString(e_)); } while (0)

inline size_t elem_size(int dtype) { return dtype == OCO_F32 ? 4 : 8; }

int check_common(const void* z, const void* y, int dtype, int64_t B, int64_t T, int d, int A) {
    if (dtype != OCO_F32 && dtype != OCO_F64) return fail(OCO_ERR_VALUE, "dtype must be OCO_F32 or OCO_F64");
    if (B < 0 || T < 0 || d < 1) return fail(OCO_ERR_VALUE, "need B >= 0, T >= 0, d >= 1 (got B=%lld T=%lld d=%d)", (long long)B, (long long)T, d);
    if (A < 1 || A > OCO_MAX_VARIANTS) return fail(OCO_ERR_VALUE, "need 1 <= A <= %d (got %d)", OCO_MAX_VARIANTS, A);
    if (B > 0 && T > 0 && (!z || !y)) return fail(OCO_ERR_VALUE, "z / y must not be NULL");
    return OCO_OK;
}

int make_variants(const int* alg_flag, const double* eta0, int A, Variants& v, Etas& e) {
    if (!alg_flag || !eta0) return fail(OCO_ERR_VALUE, "alg_flag / eta0 must not be NULL");
    v.A = A;
    for (int a = 0; a < kMaxVariants; ++a) { v.alg_flag[a] = a < A ? alg_flag[a] : 0; e.e[a] = a < A ? eta0[a] : 0.0; }
    return OCO_OK;
}

struct DevBuf {           // stream-ordered scratch from the library's own pool, freed on the same stream
    void* p = nullptr; cudaStream_t st = nullptr;
    cudaError_t alloc(size_t n, cudaStream_t s) { st = s; return oco::scratch_alloc(&p, n ? n : 8, s); }
    ~DevBuf() { if (p) cudaFreeAsync(p, st); }
    template <typename T> T* as() { return (T*)p; }
};

struct DeviceGuard {
    int prev = -1; bool ok = false;
    int set(int device) {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
            return fail(OCO_ERR_RUNTIME, "no CUDA device: this engine has no CPU fallback");
        if (device < 0 || device >= n) return fail(OCO_ERR_VALUE, "device %d out of range (have %d)", device, n);
        CU(cudaGetDevice(&prev));
        CU(cudaSetDevice(device));
        ok = true;
        return OCO_OK;
    }
    ~DeviceGuard() { if (ok && prev >= 0) cudaSetDevice(prev); }
};

struct StreamPair {
    cudaStream_t s[2] = {nullptr, nullptr};
    int create() { CU(cudaStreamCreateWithFlags(&s[0], cudaStreamNonBlocking)); CU(cudaStreamCreateWithFlags(&s[1], cudaStreamNonBlocking)); return OCO_OK; }
    ~StreamPair() { for (auto x : s) if (x) cudaStreamDestroy(x); }
};

// Scratch (the sweep kernels' row slab can be tens of GB) comes from a PRIVATE memory pool per device, owned by the
// library (oco::scratch_pool, kernels_seeded.cu): it stays cached between calls without touching the release
// threshold of the process-wide default pool, and oco_release_scratch() hands it back to the driver.
void retain_pool_memory() {}

int upload_keys(const uint64_t* keys, int64_t B, DevBuf& buf, cudaStream_t st, const uint64_t** out) {
    *out = nullptr;
    if (!keys || B <= 0) return OCO_OK;
    CU(buf.alloc(sizeof(uint64_t) * 2 * (size_t)B, st));
    CU(cudaMemcpyAsync(buf.p, keys, sizeof(uint64_t) * 2 * (size_t)B, cudaMemcpyHostToDevice, st));
    *out = buf.as<uint64_t>();
    return OCO_OK;
}

int check_family(int family, int d, bool fused, int comparator_mode = OCO_COMPARATOR_TWO_PASS) {
    if (family < OCO_FAM_GPATH || family > OCO_FAM_SWITCH) return fail(OCO_ERR_VALUE, "unknown sequence family %d", family);
    if (fused) {
        if (family > OCO_FAM_NOISY) return fail(OCO_ERR_VALUE, "fused sweep supports the seeded families only (got %d)", family);
        const int dmax = family == OCO_FAM_GPATH ? 8 : 5;      // theta of every variant lives in registers (thread per sequence)
        if (d > dmax) return fail(OCO_ERR_VALUE, "fused sweep supports d <= %d here (got %d; g-path: d <= 8, iid families: d <= 5); generate + simulate instead", dmax, d);
    } else if (family == OCO_FAM_GPATH) {
        if (d > 128) return fail(OCO_ERR_VALUE, "g-path generator supports d <= 128 (got %d)", d);
    } else if (family <= OCO_FAM_NOISY) {
        const bool ok = (d >= 1 && d <= 5) || d == 8 || d == 16 || d == 32 || d == 64;
        if (!ok) return fail(OCO_ERR_VALUE, "iid generators support d in {1..5, 8, 16, 32, 64} (got %d)", d);
    }
    return OCO_OK;
}

}  // namespace

extern "C" {

OCO_API const char* oco_last_error(void) { return g_err.c_str(); }
OCO_API int oco_abi_version(void) { return 1; }

OCO_API int oco_device_count(int* count) {
    if (!count) return fail(OCO_ERR_VALUE, "count must not be NULL");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; }
    *count = n;
    return OCO_OK;
}

// ---------------------------------------------------------------------------------- simulate
OCO_API int oco_simulate_batch(const void* z, const void* y, int in_dtype, int compute_dtype, int64_t B, int64_t T, int d,
                       const int* alg_flag, const double* eta0, int A, const double* comparator,
                       double* regret, double* cum_loss, double* comp_loss,
                       double* theta_final, double* x_last, double* x_trace, int comparator_mode, void* stream) {
    if (int rc = check_common(z, y, in_dtype, B, T, d, A)) return rc;
    if (d > 256) return fail(OCO_ERR_VALUE, "d <= 256 supported (got %d)", d);
    if (x_trace && d > kMaxSmallD)
        return fail(OCO_ERR_VALUE, "x_trace is available for d <= %d only", kMaxSmallD);
    if (comparator_mode != OCO_COMPARATOR_TWO_PASS && comparator_mode != OCO_COMPARATOR_LINEAR)
        return fail(OCO_ERR_VALUE, "unknown comparator_mode %d", comparator_mode);
    if (compute_dtype != OCO_F32 && compute_dtype != OCO_F64) return fail(OCO_ERR_VALUE, "compute_dtype must be OCO_F32 or OCO_F64");
    if (compute_dtype == OCO_F32 && (in_dtype != OCO_F32 || d > kMaxSmallD || comparator || x_trace))
        return fail(OCO_ERR_VALUE, "float32 compute mode needs float32 inputs, d <= %d, no comparator / x_trace", kMaxSmallD);
    Variants v; Etas e;
    if (int rc = make_variants(alg_flag, eta0, A, v, e)) return rc;
    if (B == 0) return OCO_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DevBuf tab;
    CU(tab.alloc(sizeof(double) * (size_t)(T * A), st));
    CU(launch_ftrl_scale_table(tab.as<double>(), e, T, A, st));
    if (compute_dtype == OCO_F32)
        CU(launch_sim_stream_f32(z, y, B, T, d, v, tab.as<double>(), regret, cum_loss, comp_loss, theta_final, x_last, st));
    else if (d <= kMaxSmallD)
        CU(launch_sim_stream(z, y, in_dtype == OCO_F32, B, T, d, v, tab.as<double>(), comparator, regret, cum_loss, comp_loss,
                             theta_final, x_last, x_trace, comparator_mode, st));
    else if (comparator || x_last)   // exact_ftl._simulate_ftrl for d > 8: one warp per (sequence, variant)
        CU(launch_wide_fixed_comp(z, y, in_dtype == OCO_F32, B, T, d, v, tab.as<double>(), comparator, regret, cum_loss, comp_loss,
                                  theta_final, x_last, st));
    else if (wide_tma_eligible(z, y, in_dtype == OCO_F32, T, d))
        CU(launch_sim_wide_tma(z, y, in_dtype == OCO_F32, B, T, d, v, tab.as<double>(), regret, cum_loss, comp_loss, theta_final, comparator_mode, st));
    else    // unaligned shapes: plain coalesced loads
        CU(launch_sim_wide(z, y, in_dtype == OCO_F32, B, T, d, v, tab.as<double>(), regret, cum_loss, comp_loss, theta_final, comparator_mode, st));
    return OCO_OK;
}

OCO_API int oco_simulate_batch_host(const void* z, const void* y, int in_dtype, int compute_dtype, int64_t B, int64_t T, int d,
                            const int* alg_flag, const double* eta0, int A, const double* comparator,
                            double* regret, double* cum_loss, double* comp_loss,
                            double* theta_final, double* x_last, double* x_trace, int comparator_mode, int device) {
    if (int rc = check_common(z, y, in_dtype, B, T, d, A)) return rc;
    DeviceGuard g; if (int rc = g.set(device)) return rc;
    if (B == 0) return OCO_OK;
    StreamPair sp; if (int rc = sp.create()) return rc;
    const size_t es = elem_size(in_dtype);
    const size_t seq_bytes = (size_t)T * (d + 1) * es;
    const size_t out_per_seq = (size_t)A * 8 * (3 + (theta_final ? d : 0) + (x_trace ? (size_t)T * d : 0));
    // Two buffer sets are in flight (copy of chunk i+1 overlaps the kernel of chunk i).  A trajectory is a
    // serial chain of T rounds (~2 us each), so a chunk must hold enough sequences for its kernel to be
    // shorter than its own PCIe copy: up to 2 GiB per set, and at least two chunks when the batch is big.
    retain_pool_memory();
    const size_t per_seq = std::max<size_t>(1, seq_bytes + out_per_seq);
    int64_t chunk = (int64_t)std::max<size_t>(1, ((size_t)2 << 30) / per_seq);
    if ((size_t)B * per_seq > ((size_t)512 << 20)) chunk = std::min<int64_t>(chunk, (B + 1) / 2);
    chunk = std::min<int64_t>(chunk, B);
    DevBuf zd[2], yd[2], rd[2], cd[2], pd[2], td[2], xd[2], ld[2], qd[2];
    for (int i = 0; i < 2; ++i) {
        cudaStream_t s = sp.s[i];
        CU(zd[i].alloc((size_t)chunk * T * d * es, s)); CU(yd[i].alloc((size_t)chunk * T * es, s));
        CU(rd[i].alloc((size_t)chunk * A * 8, s)); CU(cd[i].alloc((size_t)chunk * A * 8, s)); CU(pd[i].alloc((size_t)chunk * A * 8, s));
        if (theta_final) CU(td[i].alloc((size_t)chunk * A * d * 8, s));
        if (x_last) CU(ld[i].alloc((size_t)chunk * A * d * 8, s));
        if (comparator) CU(qd[i].alloc((size_t)chunk * d * 8, s));
        if (x_trace) CU(xd[i].alloc((size_t)chunk * A * T * d * 8, s));
    }
    int it = 0;
    for (int64_t b0 = 0; b0 < B; b0 += chunk, ++it) {
        const int i = it & 1; cudaStream_t s = sp.s[i];
        const int64_t nb = std::min<int64_t>(chunk, B - b0);
        CU(cudaMemcpyAsync(zd[i].p, (const char*)z + (size_t)b0 * T * d * es, (size_t)nb * T * d * es, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(yd[i].p, (const char*)y + (size_t)b0 * T * es, (size_t)nb * T * es, cudaMemcpyHostToDevice, s));
        if (comparator) CU(cudaMemcpyAsync(qd[i].p, comparator + b0 * d, (size_t)nb * d * 8, cudaMemcpyHostToDevice, s));
        if (int rc = oco_simulate_batch(zd[i].p, yd[i].p, in_dtype, compute_dtype, nb, T, d, alg_flag, eta0, A,
                                        comparator ? qd[i].as<double>() : nullptr,
                                        rd[i].as<double>(), cd[i].as<double>(), pd[i].as<double>(),
                                        theta_final ? td[i].as<double>() : nullptr, x_last ? ld[i].as<double>() : nullptr,
                                        x_trace ? xd[i].as<double>() : nullptr, comparator_mode, s)) return rc;
        if (x_last) CU(cudaMemcpyAsync(x_last + b0 * A * d, ld[i].p, (size_t)nb * A * d * 8, cudaMemcpyDeviceToHost, s));
        if (regret) CU(cudaMemcpyAsync(regret + b0 * A, rd[i].p, (size_t)nb * A * 8, cudaMemcpyDeviceToHost, s));
        if (cum_loss) CU(cudaMemcpyAsync(cum_loss + b0 * A, cd[i].p, (size_t)nb * A * 8, cudaMemcpyDeviceToHost, s));
        if (comp_loss) CU(cudaMemcpyAsync(comp_loss + b0 * A, pd[i].p, (size_t)nb * A * 8, cudaMemcpyDeviceToHost, s));
        if (theta_final) CU(cudaMemcpyAsync(theta_final + b0 * A * d, td[i].p, (size_t)nb * A * d * 8, cudaMemcpyDeviceToHost, s));
        if (x_trace) CU(cudaMemcpyAsync(x_trace + b0 * A * T * d, xd[i].p, (size_t)nb * A * T * d * 8, cudaMemcpyDeviceToHost, s));
    }
    CU(cudaStreamSynchronize(sp.s[0])); CU(cudaStreamSynchronize(sp.s[1]));
    return OCO_OK;
}

// ---------------------------------------------------------------------------------- exact FTL (linear regime)
OCO_API int oco_exact_ftl_prefix_batch(const void* z, const void* y, int in_dtype, int64_t B, int64_t T, int d,
                               double* actions, double* cum_loss, double* comp_loss, double* regret,
                               int* status, void* stream) {
    if (int rc = check_common(z, y, in_dtype, B, T, d, 1)) return rc;
    if (d > 256) return fail(OCO_ERR_VALUE, "exact FTL: d <= 256 supported (got %d)", d);
    if (B > 0 && !status) return fail(OCO_ERR_VALUE, "status must not be NULL");
    if (B == 0) return OCO_OK;
    if (d <= kMaxSmallD)
        CU(launch_exact_ftl_linear(z, y, in_dtype == OCO_F32, B, T, d, actions, cum_loss, comp_loss, regret, status, (cudaStream_t)stream));
    else    // the reference has no dimension limit (its own smoke run is d = 10): warp per sequence
        CU(launch_wide_exact_prefix(z, y, in_dtype == OCO_F32, B, T, d, actions, cum_loss, comp_loss, regret, status, (cudaStream_t)stream));
    return OCO_OK;
}

OCO_API int oco_exact_ftl_prefix_batch_host(const void* z, const void* y, int in_dtype, int64_t B, int64_t T, int d,
                                    double* actions, double* cum_loss, double* comp_loss, double* regret,
                                    int* status, int device) {
    if (int rc = check_common(z, y, in_dtype, B, T, d, 1)) return rc;
    if (B > 0 && !status) return fail(OCO_ERR_VALUE, "status must not be NULL");
    DeviceGuard g; if (int rc = g.set(device)) return rc;
    if (B == 0) return OCO_OK;
    cudaStream_t st; CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const size_t es = elem_size(in_dtype);
    int rc = OCO_OK;
    {
        DevBuf zd, yd, ad, cd, pd, rd, sd;
        bool ok = zd.alloc((size_t)B * T * d * es, st) == cudaSuccess && yd.alloc((size_t)B * T * es, st) == cudaSuccess &&
                  cd.alloc((size_t)B * 8, st) == cudaSuccess && pd.alloc((size_t)B * 8, st) == cudaSuccess &&
                  rd.alloc((size_t)B * 8, st) == cudaSuccess && sd.alloc((size_t)B * 4, st) == cudaSuccess;
        if (actions) ok = ok && ad.alloc((size_t)B * (T + 1) * d * 8, st) == cudaSuccess;
        if (!ok) rc = fail(OCO_ERR_RUNTIME, "device allocation failed");
        if (!rc && (cudaMemcpyAsync(zd.p, z, (size_t)B * T * d * es, cudaMemcpyHostToDevice, st) != cudaSuccess ||
                    cudaMemcpyAsync(yd.p, y, (size_t)B * T * es, cudaMemcpyHostToDevice, st) != cudaSuccess))
            rc = fail(OCO_ERR_RUNTIME, "host-to-device copy failed");
        if (!rc) rc = oco_exact_ftl_prefix_batch(zd.p, yd.p, in_dtype, B, T, d, ad.as<double>(), cd.as<double>(), pd.as<double>(),
                                                 rd.as<double>(), sd.as<int>(), st);
        if (!rc) {
            bool c = cudaMemcpyAsync(status, sd.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st) == cudaSuccess;
            if (actions) c = c && cudaMemcpyAsync(actions, ad.p, (size_t)B * (T + 1) * d * 8, cudaMemcpyDeviceToHost, st) == cudaSuccess;
            if (cum_loss) c = c && cudaMemcpyAsync(cum_loss, cd.p, (size_t)B * 8, cudaMemcpyDeviceToHost, st) == cudaSuccess;
            if (comp_loss) c = c && cudaMemcpyAsync(comp_loss, pd.p, (size_t)B * 8, cudaMemcpyDeviceToHost, st) == cudaSuccess;
            if (regret) c = c && cudaMemcpyAsync(regret, rd.p, (size_t)B * 8, cudaMemcpyDeviceToHost, st) == cudaSuccess;
            if (!c) rc = fail(OCO_ERR_RUNTIME, "device-to-host copy failed");
        }
    }
    cudaError_t es_ = cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    if (!rc && es_ != cudaSuccess) rc = fail(OCO_ERR_RUNTIME, "exact FTL failed: %s", cudaGetErrorString(es_));
    return rc;
}

OCO_API int oco_exact_ftl_general_batch(const void* z, const void* y, int in_dtype, int64_t B, int64_t T, int d, int norm,
                                int max_iter, double tol, int64_t first_prefix, double* actions, double* cum_loss, double* comp_loss,
                                double* regret, int* status, void* stream) {
    if (int rc = check_common(z, y, in_dtype, B, T, d, 1)) return rc;
    if (d > 256) return fail(OCO_ERR_VALUE, "exact FTL: d <= 256 supported (got %d)", d);
    if (2 * d > kWideIpmMaxD && norm == 2)
        return fail(OCO_ERR_VALUE, "exact FTL on the l1 ball: d <= %d supported (got %d)", kWideIpmMaxD / 2, d);
    if (d > kWideIpmMaxD && norm != 0)
        return fail(OCO_ERR_VALUE, "exact FTL on the linf ball: d <= %d supported (got %d)", kWideIpmMaxD, d);
    if (norm < 0 || norm > 2) return fail(OCO_ERR_VALUE, "norm must be 0 (l2), 1 (linf) or 2 (l1)");
    if (max_iter < 1 || !(tol > 0.0)) return fail(OCO_ERR_VALUE, "need max_iter >= 1 and tol > 0");
    if (first_prefix < 0 || first_prefix > T + 1) return fail(OCO_ERR_VALUE, "need 0 <= first_prefix <= T + 1");
    if (B > 0 && (!status || !actions)) return fail(OCO_ERR_VALUE, "status / actions must not be NULL");
    if (B == 0) return OCO_OK;
    cudaStream_t st = (cudaStream_t)stream;
    retain_pool_memory();
    DevBuf fl;
    const int* flagged = nullptr;
    if (d > kMaxSmallD && d <= kWideIpmMaxD) {
        // 8 < d <= 32: lane-per-variable interior-point kernel (kernels_wide_ipm.cu).  l2: the closed form wherever the
        // sequence is in the linear regime (its status = 1 marks the others, which are then solved); linf: all solved
        if (norm == 0) {
            CU(launch_wide_exact_prefix(z, y, in_dtype == OCO_F32, B, T, d, actions, nullptr, nullptr, nullptr, status, st));
            CU(fl.alloc(sizeof(int) * (size_t)B, st));
            CU(cudaMemcpyAsync(fl.p, status, sizeof(int) * (size_t)B, cudaMemcpyDeviceToDevice, st));
            flagged = fl.as<int>();
        }
        CU(cudaMemsetAsync(status, 0, sizeof(int) * (size_t)B, st));
        CU(launch_exact_ftl_wide_general(z, y, in_dtype == OCO_F32, B, T, d, norm, max_iter, tol, first_prefix, flagged, actions,
                                         status, st));
        if (cum_loss || comp_loss || regret)
            CU(launch_wide_replay(z, y, in_dtype == OCO_F32, B, T, d, actions, cum_loss, comp_loss, regret, st));
        return OCO_OK;
    }
    if (d > kMaxSmallD) {
        // l2 ball, d > 32: the closed form of the linear regime only (every reference workload); a sequence outside the
        // regime gets status T (= every prefix unsolved), which the shim turns into the reference's solver failure
        CU(launch_wide_exact_prefix(z, y, in_dtype == OCO_F32, B, T, d, actions, cum_loss, comp_loss, regret, status, st));
        CU(launch_status_scale(status, B, (int)std::min<int64_t>(T > 0 ? T : 1, 0x7fffffff), st));
        return OCO_OK;
    }
    if (norm == 0) {    // closed form wherever the sequence is in the linear regime; its status = 1 marks the others
        CU(fl.alloc(sizeof(int) * (size_t)B, st));
        CU(launch_exact_ftl_linear(z, y, in_dtype == OCO_F32, B, T, d, actions, nullptr, nullptr, nullptr, fl.as<int>(), st));
        flagged = fl.as<int>();
    }
    CU(cudaMemsetAsync(status, 0, sizeof(int) * (size_t)B, st));
    CU(launch_exact_ftl_general(z, y, in_dtype == OCO_F32, B, T, d, norm, max_iter, tol, first_prefix, flagged, actions,
                                status, st));
    if (cum_loss || comp_loss || regret)
        CU(launch_replay_actions(z, y, in_dtype == OCO_F32, B, T, d, actions, cum_loss, comp_loss, regret, st));
    return OCO_OK;
}

OCO_API int oco_exact_ftl_general_batch_host(const void* z, const void* y, int in_dtype, int64_t B, int64_t T, int d, int norm,
                                     int max_iter, double tol, int64_t first_prefix, double* actions, double* cum_loss, double* comp_loss,
                                     double* regret, int* status, int device) {
    if (int rc = check_common(z, y, in_dtype, B, T, d, 1)) return rc;
    if (B > 0 && (!status || !actions)) return fail(OCO_ERR_VALUE, "status / actions must not be NULL");
    DeviceGuard g; if (int rc = g.set(device)) return rc;
    if (B == 0) return OCO_OK;
    cudaStream_t st; CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const size_t es = elem_size(in_dtype);
    int rc = OCO_OK;
    {
        DevBuf zd, yd, ad, cd, pd, rd, sd;
        const size_t na = (size_t)B * (T + 1) * d * 8;
        bool ok = zd.alloc((size_t)B * T * d * es, st) == cudaSuccess && yd.alloc((size_t)B * T * es, st) == cudaSuccess &&
                  ad.alloc(na, st) == cudaSuccess && cd.alloc((size_t)B * 8, st) == cudaSuccess &&
                  pd.alloc((size_t)B * 8, st) == cudaSuccess && rd.alloc((size_t)B * 8, st) == cudaSuccess &&
                  sd.alloc((size_t)B * 4, st) == cudaSuccess;
        if (!ok) rc = fail(OCO_ERR_RUNTIME, "device allocation failed");
        if (!rc && T > 0 && (cudaMemcpyAsync(zd.p, z, (size_t)B * T * d * es, cudaMemcpyHostToDevice, st) != cudaSuccess ||
                             cudaMemcpyAsync(yd.p, y, (size_t)B * T * es, cudaMemcpyHostToDevice, st) != cudaSuccess))
            rc = fail(OCO_ERR_RUNTIME, "host-to-device copy failed");
        if (!rc) rc = oco_exact_ftl_general_batch(zd.p, yd.p, in_dtype, B, T, d, norm, max_iter, tol, first_prefix, ad.as<double>(),
                                                  cd.as<double>(), pd.as<double>(), rd.as<double>(), sd.as<int>(), st);
        if (!rc) {
            bool c = cudaMemcpyAsync(status, sd.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st) == cudaSuccess &&
                     cudaMemcpyAsync(actions, ad.p, na, cudaMemcpyDeviceToHost, st) == cudaSuccess;
            if (cum_loss) c = c && cudaMemcpyAsync(cum_loss, cd.p, (size_t)B * 8, cudaMemcpyDeviceToHost, st) == cudaSuccess;
            if (comp_loss) c = c && cudaMemcpyAsync(comp_loss, pd.p, (size_t)B * 8, cudaMemcpyDeviceToHost, st) == cudaSuccess;
            if (regret) c = c && cudaMemcpyAsync(regret, rd.p, (size_t)B * 8, cudaMemcpyDeviceToHost, st) == cudaSuccess;
            if (!c) rc = fail(OCO_ERR_RUNTIME, "device-to-host copy failed");
        }
    }
    cudaError_t es_ = cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    if (!rc && es_ != cudaSuccess) rc = fail(OCO_ERR_RUNTIME, "exact FTL (general) failed: %s", cudaGetErrorString(es_));
    return rc;
}

OCO_API int oco_replay_actions_batch_host(const void* z, const void* y, int in_dtype, int64_t B, int64_t T, int d,
                                  const double* actions, double* cum_loss, double* comp_loss, double* regret, int device) {
    if (int rc = check_common(z, y, in_dtype, B, T, d, 1)) return rc;
    if (d > 256) return fail(OCO_ERR_VALUE, "replay: d <= 256 supported (got %d)", d);
    if (B > 0 && !actions) return fail(OCO_ERR_VALUE, "actions must not be NULL");
    DeviceGuard g; if (int rc = g.set(device)) return rc;
    if (B == 0) return OCO_OK;
    cudaStream_t st; CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const size_t es = elem_size(in_dtype);
    int rc = OCO_OK;
    {
        DevBuf zd, yd, ad, cd, pd, rd;
        const size_t na = (size_t)B * (T + 1) * d * 8;
        bool ok = zd.alloc((size_t)B * T * d * es, st) == cudaSuccess && yd.alloc((size_t)B * T * es, st) == cudaSuccess &&
                  ad.alloc(na, st) == cudaSuccess && cd.alloc((size_t)B * 8, st) == cudaSuccess &&
                  pd.alloc((size_t)B * 8, st) == cudaSuccess && rd.alloc((size_t)B * 8, st) == cudaSuccess;
        if (!ok) rc = fail(OCO_ERR_RUNTIME, "device allocation failed");
        if (!rc && ((T > 0 && (cudaMemcpyAsync(zd.p, z, (size_t)B * T * d * es, cudaMemcpyHostToDevice, st) != cudaSuccess ||
                               cudaMemcpyAsync(yd.p, y, (size_t)B * T * es, cudaMemcpyHostToDevice, st) != cudaSuccess)) ||
                    cudaMemcpyAsync(ad.p, actions, na, cudaMemcpyHostToDevice, st) != cudaSuccess))
            rc = fail(OCO_ERR_RUNTIME, "host-to-device copy failed");
        if (!rc && (d <= kMaxSmallD ? launch_replay_actions(zd.p, yd.p, in_dtype == OCO_F32, B, T, d, ad.as<double>(), cd.as<double>(),
                                                            pd.as<double>(), rd.as<double>(), st)
                                    : launch_wide_replay(zd.p, yd.p, in_dtype == OCO_F32, B, T, d, ad.as<double>(), cd.as<double>(),
                                                         pd.as<double>(), rd.as<double>(), st)) != cudaSuccess)
            rc = fail(OCO_ERR_RUNTIME, "replay launch failed");
        if (!rc) {
            bool c = true;
            if (cum_loss) c = c && cudaMemcpyAsync(cum_loss, cd.p, (size_t)B * 8, cudaMemcpyDeviceToHost, st) == cudaSuccess;
            if (comp_loss) c = c && cudaMemcpyAsync(comp_loss, pd.p, (size_t)B * 8, cudaMemcpyDeviceToHost, st) == cudaSuccess;
            if (regret) c = c && cudaMemcpyAsync(regret, rd.p, (size_t)B * 8, cudaMemcpyDeviceToHost, st) == cudaSuccess;
            if (!c) rc = fail(OCO_ERR_RUNTIME, "device-to-host copy failed");
        }
    }
    cudaError_t es_ = cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    if (!rc && es_ != cudaSuccess) rc = fail(OCO_ERR_RUNTIME, "replay failed: %s", cudaGetErrorString(es_));
    return rc;
}

// ---------------------------------------------------------------------------------- SMART
OCO_API int oco_smart_batch(const void* z, const void* y, int in_dtype, int64_t B, int64_t T, int d,
                    const double* theta_thresh, const double* eta0, int A,
                    double* regret, int64_t* switch_t, int smart_mode, void* stream) {
    if (int rc = check_common(z, y, in_dtype, B, T, d, A)) return rc;
    if (A > OCO_MAX_SMART_THRESHOLDS) return fail(OCO_ERR_VALUE, "SMART: A <= %d thresholds per call (got %d)", OCO_MAX_SMART_THRESHOLDS, A);
    if (d > 256) return fail(OCO_ERR_VALUE, "SMART: d <= 256 supported (got %d)", d);
    if (!theta_thresh || !eta0 || (B > 0 && !regret)) return fail(OCO_ERR_VALUE, "theta_thresh / eta0 / regret must not be NULL");
    if (smart_mode != 0 && smart_mode != 1) return fail(OCO_ERR_VALUE, "unknown smart_mode %d", smart_mode);
    if (B == 0) return OCO_OK;
    cudaStream_t st = (cudaStream_t)stream;
    Etas e; for (int a = 0; a < kMaxVariants; ++a) e.e[a] = a < A ? eta0[a] : 0.0;
    DevBuf tab, th;
    CU(tab.alloc(sizeof(double) * (size_t)(T * A), st));
    CU(launch_ftrl_scale_table(tab.as<double>(), e, T, A, st));
    CU(th.alloc(sizeof(double) * A, st));
    CU(cudaMemcpyAsync(th.p, theta_thresh, sizeof(double) * A, cudaMemcpyHostToDevice, st));
    if (d > kMaxSmallD) {
        // 8 < d <= 256: warp per sequence with lane-distributed states (kernels_wide_smart.cu).  Both modes run the exact
        // kernel: its pruning bound already makes sequences that stay away from their thresholds cost O(T d)
        CU(launch_smart_wide(z, y, in_dtype == OCO_F32, B, T, d, th.as<double>(), tab.as<double>(), A, regret, switch_t, st));
        return OCO_OK;
    }
    DevBuf fl;
    if (smart_mode == 1) CU(fl.alloc(sizeof(int) * (size_t)B, st));
    CU(launch_smart(z, y, in_dtype == OCO_F32, B, T, d, th.as<double>(), tab.as<double>(), A, regret, switch_t, smart_mode,
                    fl.as<int>(), st));
    return OCO_OK;
}

OCO_API int oco_smart_batch_host(const void* z, const void* y, int in_dtype, int64_t B, int64_t T, int d,
                         const double* theta_thresh, const double* eta0, int A,
                         double* regret, int64_t* switch_t, int smart_mode, int device) {
    if (int rc = check_common(z, y, in_dtype, B, T, d, A)) return rc;
    DeviceGuard g; if (int rc = g.set(device)) return rc;
    if (B == 0) return OCO_OK;
    StreamPair sp; if (int rc = sp.create()) return rc;
    const size_t es = elem_size(in_dtype);
    const size_t seq_bytes = (size_t)T * (d + 1) * es;
    retain_pool_memory();
    int64_t chunk = (int64_t)std::max<size_t>(1, ((size_t)2 << 30) / std::max<size_t>(1, seq_bytes));
    if ((size_t)B * seq_bytes > ((size_t)512 << 20)) chunk = std::min<int64_t>(chunk, (B + 1) / 2);
    chunk = std::min<int64_t>(chunk, B);
    DevBuf zd[2], yd[2], rd[2], sd[2];
    for (int i = 0; i < 2; ++i) {
        cudaStream_t s = sp.s[i];
        CU(zd[i].alloc((size_t)chunk * T * d * es, s)); CU(yd[i].alloc((size_t)chunk * T * es, s));
        CU(rd[i].alloc((size_t)chunk * A * 8, s)); CU(sd[i].alloc((size_t)chunk * A * 8, s));
    }
    int it = 0;
    for (int64_t b0 = 0; b0 < B; b0 += chunk, ++it) {
        const int i = it & 1; cudaStream_t s = sp.s[i];
        const int64_t nb = std::min<int64_t>(chunk, B - b0);
        CU(cudaMemcpyAsync(zd[i].p, (const char*)z + (size_t)b0 * T * d * es, (size_t)nb * T * d * es, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(yd[i].p, (const char*)y + (size_t)b0 * T * es, (size_t)nb * T * es, cudaMemcpyHostToDevice, s));
        if (int rc = oco_smart_batch(zd[i].p, yd[i].p, in_dtype, nb, T, d, theta_thresh, eta0, A, rd[i].as<double>(), sd[i].as<int64_t>(), smart_mode, s)) return rc;
        CU(cudaMemcpyAsync(regret + b0 * A, rd[i].p, (size_t)nb * A * 8, cudaMemcpyDeviceToHost, s));
        if (switch_t) CU(cudaMemcpyAsync(switch_t + b0 * A, sd[i].p, (size_t)nb * A * 8, cudaMemcpyDeviceToHost, s));
    }
    CU(cudaStreamSynchronize(sp.s[0])); CU(cudaStreamSynchronize(sp.s[1]));
    return OCO_OK;
}

// ---------------------------------------------------------------------------------- generators
OCO_API int oco_gen_sequences(int family, uint64_t key_a, uint64_t key_c0, const uint64_t* keys,
                      int64_t B, int64_t T, int d, double p_flip, int64_t block_len,
                      int out_dtype, void* z, void* y, void* stream) {
    if (int rc = check_common(z, y, out_dtype, B, T, d, 1)) return rc;
    if (int rc = check_family(family, d, false)) return rc;
    if (family == OCO_FAM_SWITCH && block_len < 1) return fail(OCO_ERR_VALUE, "block_len must be >= 1");
    if (T > 0x7fffffff) return fail(OCO_ERR_VALUE, "T must be < 2^31");
    if (B == 0 || T == 0) return OCO_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DevBuf kb; const uint64_t* kd = nullptr;
    if (int rc = upload_keys(keys, B, kb, st, &kd)) return rc;
    CU(launch_fill(family, key_a, key_c0, kd, B, T, d, p_flip, block_len, out_dtype == OCO_F32, z, y, st));
    return OCO_OK;
}

// Touch one byte per page of an output buffer the library is about to overwrite anyway.
static void prefault_output(void* p, size_t bytes) {
    if (!p || !bytes) return;
    volatile unsigned char* c = reinterpret_cast<volatile unsigned char*>(p);
    for (size_t off = 0; off < bytes; off += 4096) c[off] = 0;
    c[bytes - 1] = 0;
}

OCO_API int oco_gen_sequences_host(int family, uint64_t key_a, uint64_t key_c0, const uint64_t* keys,
                           int64_t B, int64_t T, int d, double p_flip, int64_t block_len,
                           int out_dtype, void* z, void* y, int device) {
    if (int rc = check_common(z, y, out_dtype, B, T, d, 1)) return rc;
    DeviceGuard g; if (int rc = g.set(device)) return rc;
    if (B == 0 || T == 0) return OCO_OK;
    cudaStream_t st; CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const size_t es = elem_size(out_dtype);
    int rc = OCO_OK;
    {
        DevBuf zd, yd;
        cudaError_t e1 = zd.alloc((size_t)B * T * d * es, st), e2 = yd.alloc((size_t)B * T * es, st);
        if (e1 != cudaSuccess || e2 != cudaSuccess) rc = fail(OCO_ERR_RUNTIME, "device allocation failed");
        if (!rc) rc = oco_gen_sequences(family, key_a, key_c0, keys, B, T, d, p_flip, block_len, out_dtype, zd.p, yd.p, st);
        if (!rc) { prefault_output(z, (size_t)B * T * d * es); prefault_output(y, (size_t)B * T * es); }   // while the fill runs
        if (!rc && (cudaMemcpyAsync(z, zd.p, (size_t)B * T * d * es, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                    cudaMemcpyAsync(y, yd.p, (size_t)B * T * es, cudaMemcpyDeviceToHost, st) != cudaSuccess))
            rc = fail(OCO_ERR_RUNTIME, "device-to-host copy failed");
    }
    cudaError_t es_ = cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    if (!rc && es_ != cudaSuccess) rc = fail(OCO_ERR_RUNTIME, "stream sync failed: %s", cudaGetErrorString(es_));
    return rc;
}

// ---------------------------------------------------------------------------------- fused sweep
OCO_API int oco_regret_sweep(int family, uint64_t key_a, uint64_t key_c0, const uint64_t* keys,
                     int64_t B, int64_t T, int d, double p_flip,
                     const int* alg_flag, const double* eta0, int A,
                     double* regret, double* cum_loss, double* comp_loss, double* stats, void* stream) {
    return oco_regret_sweep_mode(family, key_a, key_c0, keys, B, T, d, p_flip, alg_flag, eta0, A, regret, cum_loss, comp_loss,
                                 stats, OCO_COMPARATOR_TWO_PASS, stream);
}

OCO_API int oco_regret_sweep_mode(int family, uint64_t key_a, uint64_t key_c0, const uint64_t* keys,
                          int64_t B, int64_t T, int d, double p_flip,
                          const int* alg_flag, const double* eta0, int A,
                          double* regret, double* cum_loss, double* comp_loss, double* stats, int comparator_mode, void* stream) {
    if (comparator_mode != OCO_COMPARATOR_TWO_PASS && comparator_mode != OCO_COMPARATOR_LINEAR)
        return fail(OCO_ERR_VALUE, "comparator_mode must be OCO_COMPARATOR_TWO_PASS or OCO_COMPARATOR_LINEAR");
    if (B < 0 || T < 0 || d < 1) return fail(OCO_ERR_VALUE, "need B >= 0, T >= 0, d >= 1");
    if (A < 1 || A > OCO_MAX_VARIANTS) return fail(OCO_ERR_VALUE, "need 1 <= A <= %d (got %d)", OCO_MAX_VARIANTS, A);
    if (int rc = check_family(family, d, true, comparator_mode)) return rc;
    Variants v; Etas e;
    if (int rc = make_variants(alg_flag, eta0, A, v, e)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (B == 0) {
        if (stats) CU(cudaMemsetAsync(stats, 0, sizeof(double) * 4 * A, st));
        return OCO_OK;
    }
    if (T > 0x7fffffff) return fail(OCO_ERR_VALUE, "T must be < 2^31");
    retain_pool_memory();
    DevBuf tab, kb, part;
    SweepPlan plan; make_sweep_plan(v, e, &plan);
    CU(tab.alloc(sizeof(double) * (size_t)sweep_table_elems(plan, T), st));
    CU(launch_sweep_table(tab.as<double>(), plan, T, st));
    const uint64_t* kd = nullptr;
    if (int rc = upload_keys(keys, B, kb, st, &kd)) return rc;
    const int nblocks = sweep_gpath_blocks(B);
    if (stats) CU(part.alloc(sizeof(double) * 3 * (size_t)nblocks * A, st));
    CU(launch_sweep(family, key_a, key_c0, kd, T, B, d, p_flip, plan, tab.as<double>(), regret, cum_loss, comp_loss,
                    stats ? part.as<double>() : nullptr, comparator_mode, st));
    if (stats) CU(launch_stats_finalize(part.as<double>(), nblocks, A, B, stats, st));
    return OCO_OK;
}

OCO_API int oco_regret_sweep_host(int family, uint64_t key_a, uint64_t key_c0, const uint64_t* keys,
                          int64_t B, int64_t T, int d, double p_flip,
                          const int* alg_flag, const double* eta0, int A,
                          double* regret, double* cum_loss, double* comp_loss, double* stats, int device) {
    return oco_regret_sweep_mode_host(family, key_a, key_c0, keys, B, T, d, p_flip, alg_flag, eta0, A, regret, cum_loss,
                                      comp_loss, stats, OCO_COMPARATOR_TWO_PASS, device);
}

OCO_API int oco_regret_sweep_mode_host(int family, uint64_t key_a, uint64_t key_c0, const uint64_t* keys,
                               int64_t B, int64_t T, int d, double p_flip,
                               const int* alg_flag, const double* eta0, int A,
                               double* regret, double* cum_loss, double* comp_loss, double* stats, int comparator_mode, int device) {
    if (B < 0 || A < 1 || A > OCO_MAX_VARIANTS) return fail(OCO_ERR_VALUE, "need B >= 0 and 1 <= A <= %d", OCO_MAX_VARIANTS);
    DeviceGuard g; if (int rc = g.set(device)) return rc;
    cudaStream_t st; CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int rc = OCO_OK;
    {
        DevBuf rd, cd, pd, sd;
        const size_t n = (size_t)B * A * 8;
        bool ok = true;
        if (regret) ok = ok && rd.alloc(n, st) == cudaSuccess;
        if (cum_loss) ok = ok && cd.alloc(n, st) == cudaSuccess;
        if (comp_loss) ok = ok && pd.alloc(n, st) == cudaSuccess;
        if (stats) ok = ok && sd.alloc(sizeof(double) * 4 * A, st) == cudaSuccess;
        if (!ok) rc = fail(OCO_ERR_RUNTIME, "device allocation failed");
        if (!rc) rc = oco_regret_sweep_mode(family, key_a, key_c0, keys, B, T, d, p_flip, alg_flag, eta0, A,
                                            rd.as<double>(), cd.as<double>(), pd.as<double>(), sd.as<double>(), comparator_mode, st);
        if (!rc) {
            // The sweep is now running; fault in the caller's (typically fresh, pageable) output pages meanwhile.
            // First-touch faults on 72 MB cost ~13 ms when they happen inside the device-to-host copy instead.
            prefault_output(regret, n); prefault_output(cum_loss, n); prefault_output(comp_loss, n);
            bool c = true;
            if (regret && n) c = c && cudaMemcpyAsync(regret, rd.p, n, cudaMemcpyDeviceToHost, st) == cudaSuccess;
            if (cum_loss && n) c = c && cudaMemcpyAsync(cum_loss, cd.p, n, cudaMemcpyDeviceToHost, st) == cudaSuccess;
            if (comp_loss && n) c = c && cudaMemcpyAsync(comp_loss, pd.p, n, cudaMemcpyDeviceToHost, st) == cudaSuccess;
            if (stats) c = c && cudaMemcpyAsync(stats, sd.p, sizeof(double) * 4 * A, cudaMemcpyDeviceToHost, st) == cudaSuccess;
            if (!c) rc = fail(OCO_ERR_RUNTIME, "device-to-host copy failed");
        }
    }
    cudaError_t es_ = cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    if (!rc && es_ != cudaSuccess) rc = fail(OCO_ERR_RUNTIME, "sweep failed: %s", cudaGetErrorString(es_));
    return rc;
}

OCO_API int oco_g_sweep_host(const int64_t* T_grid, int nT, uint64_t run_lo, int64_t n_runs, uint64_t base_seed,
                     int d, double eta0, double* g_max, double* stats, int device) {
    return oco_g_sweep_mode_host(T_grid, nT, run_lo, n_runs, base_seed, d, eta0, g_max, stats, OCO_COMPARATOR_TWO_PASS, device);
}

// Exact maximum of one horizon at single-pass speed (OCO_COMPARATOR_EXACT_MAX, see the header): single-pass sweep with
// per-run regrets, candidate selection on the device, literal two-pass simulation of the candidates' sequences.
// Writes (max, sum, sumsq, count) to stats_dev[0..3] (device); returns OCO_OK or an error code.
static int g_sweep_exact_max_one(int64_t T, uint64_t run_lo, int64_t n_runs, uint64_t base_seed, int d, double eta0,
                                 double* stats_dev, cudaStream_t st) {
    const int flag = OCO_ALG_FTRL;
    constexpr int kCap = 4096;
    DevBuf reg, idx, cnt;
    CU(reg.alloc(sizeof(double) * (size_t)n_runs, st));
    CU(idx.alloc(sizeof(int64_t) * kCap, st));
    CU(cnt.alloc(sizeof(int), st));
    if (int rc = oco_regret_sweep_mode(OCO_FAM_GPATH, base_seed, run_lo, nullptr, n_runs, T, d, -1.0, &flag, &eta0, 1,
                                       reg.as<double>(), nullptr, nullptr, stats_dev, OCO_COMPARATOR_LINEAR, st)) return rc;
    const double Td = (double)T, margin = 1e-9 * Td + 2.3e-16 * Td * Td;    // worst-case rounding of every sum involved
    CU(launch_select_candidates(reg.as<double>(), n_runs, 1, stats_dev, margin, idx.as<int64_t>(), cnt.as<int>(), kCap, st));
    int n_cand = 0;
    std::vector<int64_t> hidx(kCap);
    CU(cudaMemcpyAsync(&n_cand, cnt.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(hidx.data(), idx.p, sizeof(int64_t) * kCap, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (n_cand > kCap) {       // a pathological pile-up at the maximum: the literal pass over every run
        return oco_regret_sweep_mode(OCO_FAM_GPATH, base_seed, run_lo, nullptr, n_runs, T, d, -1.0, &flag, &eta0, 1,
                                     nullptr, nullptr, nullptr, stats_dev, OCO_COMPARATOR_TWO_PASS, st);
    }
    // literal two-pass regrets of the candidates: generate their sequences (explicit keys) and simulate (bit-exact for d <= 8)
    std::vector<uint64_t> keys(2 * (size_t)n_cand);
    for (int k = 0; k < n_cand; ++k) { keys[2 * k] = base_seed; keys[2 * k + 1] = run_lo + (uint64_t)hidx[k]; }
    DevBuf zc, yc, rc_, mx;
    CU(zc.alloc(sizeof(double) * (size_t)n_cand * T * d, st));
    CU(yc.alloc(sizeof(double) * (size_t)n_cand * T, st));
    CU(rc_.alloc(sizeof(double) * (size_t)n_cand, st));
    if (int rc = oco_gen_sequences(OCO_FAM_GPATH, 0, 0, keys.data(), n_cand, T, d, -1.0, 20, OCO_F64, zc.p, yc.p, st)) return rc;
    if (int rc = oco_simulate_batch(zc.p, yc.p, OCO_F64, OCO_F64, n_cand, T, d, &flag, &eta0, 1, nullptr, rc_.as<double>(),
                                    nullptr, nullptr, nullptr, nullptr, nullptr, OCO_COMPARATOR_TWO_PASS, st)) return rc;
    std::vector<double> hr((size_t)n_cand);
    CU(cudaMemcpyAsync(hr.data(), rc_.p, sizeof(double) * (size_t)n_cand, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    double best = -1.0 / 0.0;
    for (double v : hr) best = v > best ? v : best;
    CU(cudaMemcpyAsync(stats_dev, &best, sizeof(double), cudaMemcpyHostToDevice, st));     // exact maximum replaces the bound's
    CU(cudaStreamSynchronize(st));
    return OCO_OK;
}

OCO_API int oco_g_sweep_mode_host(const int64_t* T_grid, int nT, uint64_t run_lo, int64_t n_runs, uint64_t base_seed,
                          int d, double eta0, double* g_max, double* stats, int comparator_mode, int device) {
    if (!T_grid || !g_max || nT < 0) return fail(OCO_ERR_VALUE, "T_grid / g_max must not be NULL");
    if (n_runs < 0) return fail(OCO_ERR_VALUE, "n_runs must be >= 0");
    if (comparator_mode < OCO_COMPARATOR_TWO_PASS || comparator_mode > OCO_COMPARATOR_EXACT_MAX)
        return fail(OCO_ERR_VALUE, "comparator_mode must be OCO_COMPARATOR_TWO_PASS, _LINEAR or _EXACT_MAX");
    DeviceGuard g; if (int rc = g.set(device)) return rc;
    cudaStream_t st; CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int rc = OCO_OK;
    std::vector<double> hs((size_t)std::max(nT, 1) * 4, 0.0);
    {
        DevBuf sd;
        if (sd.alloc(sizeof(double) * 4 * (size_t)std::max(nT, 1), st) != cudaSuccess) rc = fail(OCO_ERR_RUNTIME, "device allocation failed");
        const int flag = OCO_ALG_FTRL;
        for (int i = 0; i < nT && !rc; ++i) {
            if (T_grid[i] < 0) { rc = fail(OCO_ERR_VALUE, "T_grid[%d] < 0", i); break; }
            // the pruned exact maximum applies where the single-pass kernel does (2 <= d <= 8); d = 1 is exact in one pass anyway
            if (comparator_mode == OCO_COMPARATOR_EXACT_MAX && d >= 2 && d <= 8 && n_runs > 0 && T_grid[i] > 0)
                rc = g_sweep_exact_max_one(T_grid[i], run_lo, n_runs, base_seed, d, eta0, sd.as<double>() + 4 * i, st);
            else
                rc = oco_regret_sweep_mode(OCO_FAM_GPATH, base_seed, run_lo, nullptr, n_runs, T_grid[i], d, -1.0, &flag, &eta0, 1,
                                           nullptr, nullptr, nullptr, sd.as<double>() + 4 * i,
                                           comparator_mode == OCO_COMPARATOR_LINEAR ? OCO_COMPARATOR_LINEAR : OCO_COMPARATOR_TWO_PASS, st);
        }
        if (!rc && nT > 0 && cudaMemcpyAsync(hs.data(), sd.p, sizeof(double) * 4 * nT, cudaMemcpyDeviceToHost, st) != cudaSuccess)
            rc = fail(OCO_ERR_RUNTIME, "device-to-host copy failed");
    }
    cudaError_t es_ = cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    if (!rc && es_ != cudaSuccess) rc = fail(OCO_ERR_RUNTIME, "g sweep failed: %s", cudaGetErrorString(es_));
    if (rc) return rc;
    for (int i = 0; i < nT; ++i) {
        const double mx = n_runs > 0 ? hs[4 * i] : 0.0;
        g_max[i] = mx > 0.0 ? mx : 0.0;                 // max_regret starts at 0.0 (fast_algorithms.py:228)
        if (stats) for (int k = 0; k < 4; ++k) stats[4 * i + k] = hs[4 * i + k];
    }
    return OCO_OK;
}

OCO_API int oco_rng_probe_host(int what, uint64_t a, uint64_t b, uint64_t c, int64_t n_skip, int64_t n, void* out, int device) {
    if (!out || n < 0 || n_skip < 0 || what < 0 || what > 6) return fail(OCO_ERR_VALUE, "bad rng probe arguments");
    DeviceGuard g; if (int rc = g.set(device)) return rc;
    if (n == 0) return OCO_OK;
    void* d = nullptr;
    CU(cudaMalloc(&d, (size_t)n * 8));
    cudaError_t e = launch_rng_probe(what, a, b, c, n_skip, n, d, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(out, d, (size_t)n * 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(OCO_ERR_RUNTIME, "rng probe failed: %s", cudaGetErrorString(e));
    return OCO_OK;
}

OCO_API int oco_release_scratch(int device) {
    DeviceGuard g; if (int rc = g.set(device)) return rc;
    CU(cudaDeviceSynchronize());
    CU(oco::scratch_trim());
    return OCO_OK;
}

OCO_API int oco_fp64_peak_probe_host(int device, double* ops_per_sec) {
    if (!ops_per_sec) return fail(OCO_ERR_VALUE, "ops_per_sec must not be NULL");
    DeviceGuard g; if (int rc = g.set(device)) return rc;
    cudaDeviceProp prop; CU(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, iters = 1 << 15;
    double* d = nullptr; CU(cudaMalloc(&d, 8));
    cudaEvent_t e0, e1; CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    double best = 0.0; int rc = OCO_OK;
    for (int rep = 0; rep < 4 && !rc; ++rep) {            // rep 0 is the warm-up
        cudaEventRecord(e0, nullptr);
        cudaError_t e = launch_fp64_probe(d, blocks, iters, nullptr);
        cudaEventRecord(e1, nullptr);
        if (e == cudaSuccess) e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { rc = fail(OCO_ERR_RUNTIME, "fp64 probe failed: %s", cudaGetErrorString(e)); break; }
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
        const double ops = 16.0 * iters * 256.0 * blocks / (ms * 1e-3);
        if (rep > 0 && ops > best) best = ops;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *ops_per_sec = best;
    return rc;
}

}  // extern "C"
