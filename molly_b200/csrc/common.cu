#include "common.h"

#include <atomic>
#include <mutex>
#include <vector>

#include "kernels.h"

namespace molly {

namespace {
thread_local std::string g_last_error;
std::atomic<int> g_launches{0};

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        // resolved through the runtime so that the library has no link-time dependency on libcuda
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}
}  // namespace

void set_last_error(const std::string& msg) { g_last_error = msg; }
const char* get_last_error() { return g_last_error.c_str(); }

namespace {
struct ProfRecord { int family; double work; cudaEvent_t a, b; int dev_slot; };
bool g_prof_on = false;
std::vector<ProfRecord> g_prof;
constexpr int PROF_DEV_SLOTS = 1 << 16;
double* g_prof_dev = nullptr;          // device-side work values (launches whose work depends on device data)
int g_prof_dev_used = 0;
int g_prof_dev_pending = -1;           // slot handed out by prof_next_device_work for the next record

__global__ void attention_work_kernel(const int32_t* __restrict__ kv_info, int n_seq, double per_key, double* out) {
    __shared__ double part[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n_seq; i += blockDim.x) acc += kv_info[2 * i] > 0 ? kv_info[2 * i] : 0;
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = part[0] * per_key;
}
thread_local int g_gemm_family = PF_GEMM_OTHER;
}  // namespace

void set_gemm_family(int f) { g_gemm_family = f; }
int gemm_family() { return g_gemm_family; }

ProfScope::ProfScope(int family, double work, cudaStream_t s) : stream(s), slot(-1) {
    if (!g_prof_on) return;
    ProfRecord r{family, work, nullptr, nullptr, g_prof_dev_pending};
    g_prof_dev_pending = -1;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, s);
    g_prof.push_back(r);
    slot = static_cast<int>(g_prof.size()) - 1;
}
ProfScope::~ProfScope() {
    if (slot >= 0) cudaEventRecord(g_prof[slot].b, stream);
}

double* prof_next_device_work() {
    if (!g_prof_on || g_prof_dev == nullptr || g_prof_dev_used >= PROF_DEV_SLOTS) return nullptr;
    g_prof_dev_pending = g_prof_dev_used++;
    return g_prof_dev + g_prof_dev_pending;
}

void prof_attention_work(const int32_t* kv_info, int n_seq, int k_tokens, int h, double coef, cudaStream_t stream) {
    if (double* slot = prof_next_device_work())
        attention_work_kernel<<<1, 256, 0, stream>>>(kv_info, n_seq, coef * k_tokens * static_cast<double>(h), slot);
}

void prof_start() {
    for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
    if (g_prof_dev == nullptr && cudaMalloc(&g_prof_dev, PROF_DEV_SLOTS * sizeof(double)) != cudaSuccess) g_prof_dev = nullptr;
    g_prof_dev_used = 0;
    g_prof_dev_pending = -1;
    g_prof_on = true;
}
// aggregates per family: launches[f], ms[f], work[f]; returns 0 on success
int prof_stop(int* launches, double* ms, double* work) {
    g_prof_on = false;
    for (int f = 0; f < PF_COUNT; ++f) { launches[f] = 0; ms[f] = 0.0; work[f] = 0.0; }
    int rc = 0;
    std::vector<double> dev_work(g_prof_dev_used > 0 ? g_prof_dev_used : 1, 0.0);
    if (g_prof_dev_used > 0 &&
        (cudaDeviceSynchronize() != cudaSuccess ||
         cudaMemcpy(dev_work.data(), g_prof_dev, g_prof_dev_used * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess))
        rc = 1;
    for (auto& r : g_prof) {
        float t = 0.f;
        if (cudaEventSynchronize(r.b) != cudaSuccess || cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) rc = 1;
        launches[r.family] += 1; ms[r.family] += t;
        work[r.family] += (r.dev_slot >= 0 && r.dev_slot < g_prof_dev_used && rc == 0) ? dev_work[r.dev_slot] : r.work;
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    g_prof.clear();
    return rc;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int launch_count() { return g_launches.load(std::memory_order_relaxed); }
void add_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int device_sm_count() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

int make_tma_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                uint32_t box_cols, uint32_t elem_bytes, bool swizzle) {
    EncodeTiledFn fn = get_encode_fn();
    MOLLY_CHECK(fn != nullptr, MOLLY_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    MOLLY_CHECK(elem_bytes == 2 || elem_bytes == 4, MOLLY_ERR_UNSUPPORTED, "tma: element size must be 2 (bf16) or 4 (fp32)");
    const uint32_t inner = box_cols * elem_bytes;
    CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
    if (swizzle) {
        if (inner == 128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
        else if (inner == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
        else if (inner == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
        else MOLLY_CHECK(false, MOLLY_ERR_UNSUPPORTED, "tma: swizzled box inner extent must be 32/64/128 B, got %u", inner);
    }
    MOLLY_CHECK((ld * elem_bytes) % 16 == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0, MOLLY_ERR_INVALID,
                "tma: base and row pitch must be 16-B aligned");
    MOLLY_CHECK(box_rows <= 256 && box_cols <= 256, MOLLY_ERR_INVALID, "tma: box dims must be <= 256");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * elem_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MOLLY_CHECK(r == CUDA_SUCCESS, MOLLY_ERR_CUDA,
                "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu box=%ux%u", static_cast<int>(r),
                static_cast<unsigned long long>(rows), static_cast<unsigned long long>(cols),
                static_cast<unsigned long long>(ld), box_rows, box_cols);
    return MOLLY_OK;
}

}  // namespace molly
