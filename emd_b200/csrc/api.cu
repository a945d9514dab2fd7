// C-ABI plumbing shared by all stages: error text, version, device probe.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void emd_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* emd_last_error_string() { return g_err; }

extern "C" int emd_abi_version() { return 3; }

// 0 when the current device can run the sm_100a kernels in this library.
extern "C" int emd_device_check() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        emd_set_error("emd_device_check: %s", cudaGetErrorString(e));
        return EMD_ERR_CUDA;
    }
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (major != 10) {
        emd_set_error("emd_device_check: device is sm_%d%d; this library is built for sm_100a only", major, minor);
        return EMD_ERR_UNSUPPORTED;
    }
    return EMD_OK;
}

// ---- launch accounting + optional profiling --------------------------------------------------
#include <atomic>
#include <mutex>
#include <vector>

static std::atomic<long long> g_launches{0};
static std::atomic<int> g_prof_on{0};
struct ProfRec { int id; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof_open;   // begun, not ended (per id LIFO is enough: launches do not nest)
static std::vector<ProfRec> g_prof_done;

static const char* kKernelNames[EK_COUNT] = {
    "projection_fwd", "projection_bwd", "scan", "isect_emit", "sort_hist", "sort_scatter", "isect_offsets",
    "raster_pack", "raster_fwd", "raster_bwd", "raster_gather", "sh_fwd", "sh_bwd", "activate_fwd", "activate_bwd",
    "rigid_fwd", "rigid_bwd", "smpl_fwd", "smpl_bwd", "mlp_fwd", "mlp_bwd", "dg_preprocess_fwd", "dg_preprocess_bwd",
    "hexplane_fwd", "hexplane_bwd", "adam", "image_loss_fwd", "image_loss_bwd", "voxel_lbs_fwd", "voxel_lbs_bwd",
    "dense_fwd", "dense_bwd", "deform_input", "misc"};

void emd_prof_begin(int id, cudaStream_t stream) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    ProfRec r;
    r.id = id;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, stream);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_open.push_back(r);
}

void emd_prof_end(int id, cudaStream_t stream) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (size_t i = g_prof_open.size(); i-- > 0;) {
        if (g_prof_open[i].id == id) {
            cudaEventRecord(g_prof_open[i].b, stream);
            g_prof_done.push_back(g_prof_open[i]);
            g_prof_open.erase(g_prof_open.begin() + i);
            return;
        }
    }
}

extern "C" long long emd_launch_count() { return g_launches.load(); }
extern "C" int emd_kernel_id_count() { return EK_COUNT; }
extern "C" const char* emd_kernel_name(int id) { return (id >= 0 && id < EK_COUNT) ? kKernelNames[id] : "?"; }
// When on, every kernel launch is bracketed by CUDA events on its stream.
extern "C" void emd_profile_enable(int on) { g_prof_on.store(on ? 1 : 0); }
// Synchronises the recorded events and ADDS their durations (ms) / launch counts into the caller's
// [emd_kernel_id_count()] arrays; clears the record list.
extern "C" int emd_profile_collect(double* ms_sum, long long* counts) {
    std::vector<ProfRec> recs;
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        recs.swap(g_prof_done);
    }
    for (auto& r : recs) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            ms_sum[r.id] += ms;
            counts[r.id] += 1;
        }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    return EMD_OK;
}
