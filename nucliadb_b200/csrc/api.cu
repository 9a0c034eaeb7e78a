// nidx_b200 — C ABI (include/nidx_b200.h) over the CUDA kernels.  Host side of the hot path:
// what nidx_vector's OpenSegment / HnswBuilder and the tantivy collector call do on the CPU in the
// reference is orchestrated here on one GPU.  No CPU fallback: every entry point needs a device.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/nidx_b200.h"
#include "bm25.cuh"
#include "common.cuh"
#include "hnsw_build.cuh"
#include "hnsw_search.cuh"
#include "hnsw_rabitq.cuh"
#include "rabitq.cuh"
#include "rank_fusion.cuh"
#include "scan.cu"
#include "scan_tc.cuh"
#include "scan_tc2.cuh"
#include "segment_io.hpp"
#include "shard.cuh"
#include "topk.cuh"

using namespace nidx;

// ---------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CU(expr)                                                                                              \
    do {                                                                                                      \
        cudaError_t e__ = (expr);                                                                             \
        if (e__ != cudaSuccess) return fail(NIDX_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
#define LAUNCHED() (g_launches.fetch_add(1, std::memory_order_relaxed))

static int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }
static int ilog2(int x) { int b = 0; while ((1 << b) < x) ++b; return b; }

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
        if (e != cudaSuccess) return fail(NIDX_ECUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};
#define ENSURE(buf, bytes) do { int r__ = (buf).ensure(bytes); if (r__) return r__; } while (0)

// Per-call scratch; a segment keeps a pool so concurrent searches do not share one.
struct Workspace {
    DevBuf queries, qnorms, out_ids, out_scores, out_counts, scores, partial, filter, misc, sched;
    cudaEvent_t done = nullptr;
    cudaStream_t last_stream = nullptr;
    bool busy = false;
    ~Workspace() {
        queries.release(); qnorms.release(); out_ids.release(); out_scores.release(); out_counts.release();
        scores.release(); partial.release(); filter.release(); misc.release(); sched.release();
        if (done) cudaEventDestroy(done);
    }
};

struct WorkspacePool {
    std::mutex mu;
    std::vector<Workspace*> all;
    // A free workspace last used on this stream (stream order protects it), else one whose last call has completed, else a new one
    // (up to MAX_POOL: calls issued from one host thread on alternating streams then overlap on the device instead of the second
    // waiting for the first one's buffers), else the host waits for a free one's last call.
    static constexpr size_t MAX_POOL = 8;
    Workspace* acquire(cudaStream_t stream) {
        std::lock_guard<std::mutex> g(mu);
        Workspace* pick = nullptr;
        for (Workspace* w : all)
            if (!w->busy && w->last_stream == stream) { pick = w; break; }
        if (!pick) {
            for (Workspace* w : all)
                if (!w->busy && (!w->done || cudaEventQuery(w->done) == cudaSuccess)) { pick = w; break; }
            (void)cudaGetLastError();   // cudaErrorNotReady from a query is not an error: do not leave it for the next cudaGetLastError()
        }
        if (!pick && all.size() >= MAX_POOL)
            for (Workspace* w : all)
                if (!w->busy) { cudaEventSynchronize(w->done); pick = w; break; }
        if (!pick) {
            pick = new Workspace();
            cudaEventCreateWithFlags(&pick->done, cudaEventDisableTiming);
            all.push_back(pick);
        }
        pick->busy = true;
        return pick;
    }
    void release(Workspace* w, cudaStream_t stream) {
        cudaEventRecord(w->done, stream);
        std::lock_guard<std::mutex> g(mu);
        w->last_stream = stream;
        w->busy = false;
    }
    ~WorkspacePool() { for (Workspace* w : all) delete w; }
};
struct WsGuard {
    WorkspacePool& pool; Workspace* w; cudaStream_t s;
    WsGuard(WorkspacePool& p, cudaStream_t st) : pool(p), w(p.acquire(st)), s(st) {}
    ~WsGuard() { pool.release(w, s); }
};

struct nidx_vec_segment {
    nidx_vec_config cfg;
    uint64_t n = 0;
    int d = 0, ld = 0;
    int sm_count = 0;
    float* d_vecs = nullptr;
    float* d_norms = nullptr;
    uint32_t* d_par_of = nullptr;
    uint32_t* d_par_first = nullptr;
    uint32_t n_par = 0;
    uint64_t* d_alive = nullptr;
    uint64_t alive_count = 0;         // set bits of d_alive (counted in nidx_vec_set_alive): `matching` of an unfiltered search
    uint64_t* d_par_keys = nullptr;   // [n_par] caller-supplied paragraph keys for the cross-segment de-duplication (shard.cuh)
    struct InvIndex {                 // one inverted index (inverted_index/fst_index.rs + map.rs): sorted keys on the host, postings in HBM
        std::vector<unsigned char> key_bytes;
        std::vector<uint64_t> key_off, post_off;
        uint32_t* d_post = nullptr;
        uint32_t n_keys = 0;
    } inv[2];
    // graph
    bool has_graph = false;
    std::vector<uint8_t> h_level;
    uint8_t* d_level = nullptr;
    uint32_t entry_node = 0, entry_layer = 0;
    int s0 = 0, su = 0;
    uint64_t upper_rows = 0;
    uint32_t* d_adj0 = nullptr; float* d_w0 = nullptr;
    uint64_t* d_upper_off = nullptr;
    uint32_t* d_adjU = nullptr; float* d_wU = nullptr;
    float max_norm = 0.0f;              // max |v| (error bound of the tensor-core filter for Dot)
    CUtensorMap map_v;                  // TMA descriptor of the vector block (scan_tc2.cuh), built on first use
    bool map_v_ready = false;
    std::mutex map_mu;
    unsigned char* d_quant = nullptr;   // RaBitQ codes [n][quant_stride] (vectors.quant records, padded)
    int quant_stride = 0;
    unsigned long long* d_counters = nullptr;  // [8] the build's counters
    std::atomic<unsigned long long*> last_counters{nullptr};   // counters of the LAST search call (they live in that call's workspace: concurrent
                                                               // searches never add into each other's), read by nidx_vec_counters*
    unsigned int* d_work_counter = nullptr;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;  // around the dominant kernel of the last search (bench roofline)
    WorkspacePool pool;

    VecDev vdev() const {
        VecDev v;
        v.vecs = d_vecs; v.norms = d_norms; v.paragraph_of = d_par_of; v.n = (uint32_t)n; v.d = d; v.ld = ld; v.sim = cfg.similarity;
        return v;
    }
    GraphDev gdev() const {
        GraphDev g;
        g.n = (uint32_t)n; g.M = cfg.m; g.M0 = cfg.m0; g.s0 = s0; g.su = su; g.level = d_level;
        g.entry_node = entry_node; g.entry_layer = entry_layer;
        g.adj0 = d_adj0; g.w0 = d_w0; g.upper_off = d_upper_off; g.adjU = d_adjU; g.wU = d_wU;
        return g;
    }
};

static int stride0_for(int M0) { return (M0 + 31) / 32 * 32; }
static int strideU_for(int M) { return (M + 15) / 16 * 16; }

static void free_graph(nidx_vec_segment* s) {
    cudaFree(s->d_level); cudaFree(s->d_adj0); cudaFree(s->d_w0); cudaFree(s->d_upper_off); cudaFree(s->d_adjU); cudaFree(s->d_wU);
    s->d_level = nullptr; s->d_adj0 = nullptr; s->d_w0 = nullptr; s->d_upper_off = nullptr; s->d_adjU = nullptr; s->d_wU = nullptr;
    s->has_graph = false;
}

// Allocate graph storage for the given levels (ram_hnsw.rs:88-107: every node in layers 0..=level,
// entry point in the top layer -- lowest id there).
static int alloc_graph(nidx_vec_segment* s, const uint8_t* level) {
    free_graph(s);
    uint64_t n = s->n;
    s->h_level.assign(level, level + n);
    std::vector<uint64_t> off(n ? n : 1);
    uint64_t rows = 0;
    uint32_t top = 0;
    for (uint64_t i = 0; i < n; ++i) { off[i] = rows; rows += level[i]; top = std::max<uint32_t>(top, level[i]); }
    s->upper_rows = rows;
    s->entry_layer = top;
    s->entry_node = 0;
    for (uint64_t i = 0; i < n; ++i) if (level[i] == top) { s->entry_node = (uint32_t)i; break; }
    s->s0 = stride0_for(s->cfg.m0);
    s->su = strideU_for(s->cfg.m);
    size_t n0 = (size_t)n * s->s0, nu = (size_t)std::max<uint64_t>(rows, 1) * s->su;
    CU(cudaMalloc(&s->d_level, std::max<uint64_t>(n, 1)));
    CU(cudaMalloc(&s->d_adj0, n0 * 4 + 16));
    CU(cudaMalloc(&s->d_w0, n0 * 4 + 16));
    CU(cudaMalloc(&s->d_upper_off, std::max<uint64_t>(n, 1) * 8));
    CU(cudaMalloc(&s->d_adjU, nu * 4));
    CU(cudaMalloc(&s->d_wU, nu * 4));
    CU(cudaMemcpy(s->d_level, level, n, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(s->d_upper_off, off.data(), n * 8, cudaMemcpyHostToDevice));
    CU(cudaMemset(s->d_adj0, 0xFF, n0 * 4));
    CU(cudaMemset(s->d_w0, 0, n0 * 4));
    CU(cudaMemset(s->d_adjU, 0xFF, nu * 4));
    CU(cudaMemset(s->d_wU, 0, nu * 4));
    return 0;
}

__global__ void max_norm_kernel(const float* __restrict__ norms, uint64_t n, unsigned int* __restrict__ out_bits) {
    float m = 0.0f;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) m = fmaxf(m, norms[i]);
    for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, off));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));   // non-negative floats order like their bit patterns
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_row_tensor_map(CUtensorMap* map, const float* base, uint64_t rows, int ld, int box_rows) {
    static tmap_encode_fn encode = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            encode = reinterpret_cast<tmap_encode_fn>(fn);
    });
    if (!encode) return fail(NIDX_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)TC2_KB, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(NIDX_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

extern "C" {

const char* nidx_last_error(void) { return g_err.c_str(); }
int nidx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
uint64_t nidx_launch_count(void) { return g_launches.load(); }

static int check_device(int device) {
    int n = nidx_device_count();
    if (n <= 0) return fail(NIDX_ENODEVICE, "no CUDA device available (nidx_b200 has no CPU fallback)");
    if (device < 0 || device >= n) return fail(NIDX_EINVAL, "device %d out of range (have %d)", device, n);
    CU(cudaSetDevice(device));
    return 0;
}

static int fill_defaults(nidx_vec_config* c) {
    if (c->dimension <= 0) return fail(NIDX_EINVAL, "dimension must be positive");
    if (c->similarity != NIDX_SIM_DOT && c->similarity != NIDX_SIM_COSINE && c->similarity != NIDX_SIM_L2) return fail(NIDX_EINVAL, "unknown similarity %d", c->similarity);
    if (c->m <= 0) c->m = 30;                              // params.rs:40
    if (c->m0 <= 0) c->m0 = 60;                            // params.rs:34
    if (c->ef_construction <= 0) c->ef_construction = 100; // params.rs:43
    if (c->ef_search <= 0) c->ef_search = 30;              // params.rs:46
    if (c->m0 > HS_MAX_ROW || c->m > HS_MAX_ROW) return fail(NIDX_EINVAL, "M / M0 above %d not supported", HS_MAX_ROW);
    if (c->ef_construction > HB_MAX_CAND) return fail(NIDX_EINVAL, "ef_construction above %d not supported", HB_MAX_CAND);
    return 0;
}

// Common tail of create/open: vectors are in d_vecs; compute norms, paragraph CSR.
static int finish_create(nidx_vec_segment* s, const uint32_t* paragraph_of_host) {
    uint64_t n = s->n;
    CU(cudaMalloc(&s->d_norms, std::max<uint64_t>(n, 1) * 4));
    if (n) {
        int blocks = (int)std::min<uint64_t>((n + 7) / 8, (uint64_t)s->sm_count * 16);
        row_norms_kernel<<<blocks, 256>>>(s->d_vecs, s->ld, n, s->d_norms);
        LAUNCHED();
        CU(cudaGetLastError());
    }
    s->n_par = (uint32_t)n;
    if (paragraph_of_host) {
        // paragraphs own contiguous vector ranges (data_store/v2: first_vector / num_vectors)
        std::vector<uint32_t> first;
        uint32_t prev = NIDX_NIL;
        for (uint64_t i = 0; i < n; ++i) {
            uint32_t p = paragraph_of_host[i];
            if (i == 0 || p != prev) {
                if (p != (uint32_t)first.size()) return fail(NIDX_EINVAL, "paragraph_of must be contiguous and ascending from 0 (vector %llu -> %u)", (unsigned long long)i, p);
                first.push_back((uint32_t)i);
                prev = p;
            }
        }
        first.push_back((uint32_t)n);
        s->n_par = (uint32_t)first.size() - 1;
        if (s->n_par != n) {  // only materialise when some paragraph has several vectors
            CU(cudaMalloc(&s->d_par_of, n * 4));
            CU(cudaMemcpy(s->d_par_of, paragraph_of_host, n * 4, cudaMemcpyHostToDevice));
            CU(cudaMalloc(&s->d_par_first, first.size() * 4));
            CU(cudaMemcpy(s->d_par_first, first.data(), first.size() * 4, cudaMemcpyHostToDevice));
        }
    }
    if (n) {   // max |v| for the Dot error bound of the tensor-core filter
        unsigned int* d_bits = nullptr;
        CU(cudaMalloc(&d_bits, 4));
        CU(cudaMemset(d_bits, 0, 4));
        max_norm_kernel<<<std::min<uint64_t>((n + 255) / 256, (uint64_t)s->sm_count * 8), 256>>>(s->d_norms, n, d_bits);
        LAUNCHED();
        unsigned int hb = 0;
        CU(cudaMemcpy(&hb, d_bits, 4, cudaMemcpyDeviceToHost));
        cudaFree(d_bits);
        memcpy(&s->max_norm, &hb, 4);
    }
    CU(cudaMalloc(&s->d_counters, 8 * sizeof(unsigned long long)));
    CU(cudaMemset(s->d_counters, 0, 8 * sizeof(unsigned long long)));
    CU(cudaMalloc(&s->d_work_counter, 64));
    CU(cudaEventCreate(&s->ev_k0));
    CU(cudaEventCreate(&s->ev_k1));
    CU(cudaDeviceSynchronize());
    return 0;
}

static int new_segment(const nidx_vec_config* cfg, nidx_vec_segment** out) {
    if (!cfg || !out) return fail(NIDX_EINVAL, "null argument");
    nidx_vec_config c = *cfg;
    int r = fill_defaults(&c);
    if (r) return r;
    r = check_device(c.device);
    if (r) return r;
    nidx_vec_segment* s = new nidx_vec_segment();
    s->cfg = c;
    s->d = c.dimension;
    s->ld = (c.dimension + 3) / 4 * 4;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, c.device);
    s->sm_count = prop.multiProcessorCount;
    *out = s;
    return 0;
}

int nidx_vec_create(const nidx_vec_config* cfg, const float* vectors, uint64_t n, int32_t ld, int mem, const uint32_t* paragraph_of,
                    nidx_vec_segment** out) {
    nidx_vec_segment* s = nullptr;
    int r = new_segment(cfg, &s);
    if (r) return r;
    if (n >= (1ull << 31)) { delete s; return fail(NIDX_EINVAL, "at most 2^31-1 vectors per segment"); }
    if (ld < s->d) { delete s; return fail(NIDX_EINVAL, "ld %d < dimension %d (VectorErr::InconsistentDimensions)", ld, s->d); }
    s->n = n;
    r = [&]() -> int {
        CU(cudaMalloc(&s->d_vecs, std::max<size_t>((size_t)n * s->ld * 4, 16)));
        if (n == 0) return 0;
        if (mem == NIDX_MEM_HOST && ld == s->ld) {
            CU(cudaMemcpy(s->d_vecs, vectors, (size_t)n * ld * 4, cudaMemcpyHostToDevice));
        } else if (mem == NIDX_MEM_DEVICE && ld == s->ld) {
            CU(cudaMemcpy(s->d_vecs, vectors, (size_t)n * ld * 4, cudaMemcpyDeviceToDevice));
        } else {
            const unsigned char* src = reinterpret_cast<const unsigned char*>(vectors);
            void* staged = nullptr;
            if (mem == NIDX_MEM_HOST) {
                CU(cudaMalloc(&staged, (size_t)n * ld * 4));
                CU(cudaMemcpy(staged, vectors, (size_t)n * ld * 4, cudaMemcpyHostToDevice));
                src = reinterpret_cast<const unsigned char*>(staged);
            }
            pad_rows_kernel<<<s->sm_count * 8, 256>>>(src, (size_t)ld * 4, s->d, s->d_vecs, s->ld, n);
            LAUNCHED();
            CU(cudaGetLastError());
            CU(cudaDeviceSynchronize());
            if (staged) cudaFree(staged);
        }
        return 0;
    }();
    if (!r) r = finish_create(s, paragraph_of);
    if (r) { nidx_vec_close(s); return r; }
    *out = s;
    return 0;
}

// ---- utils::normalize_vector (nidx_vector/src/utils.rs:20-23) ----------------------------------------------
// magnitude = sqrt(fold(0.0, |acc, x| acc + x.powi(2))) -- a SEQUENTIAL f32 fold (powi(2) = one rounded multiply, no FMA) --
// then x / magnitude per element.  One warp per vector: the row is staged in shared memory with coalesced loads, lane 0 replays
// the fold in the reference's order (bit-identical to it), all lanes divide.  A zero vector divides by zero like the reference
// (NaN components).  Called at index time (indexer.rs:94-146) and per query (searcher.rs:246-252): negligible next to a search.
constexpr int NRM_WARPS = 4;
__global__ void __launch_bounds__(NRM_WARPS * 32) normalize_rows_kernel(float* __restrict__ v, uint64_t n, int d, int ld) {
    extern __shared__ float nrm_row[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* row = nrm_row + (size_t)warp * d;
    for (uint64_t i = (uint64_t)blockIdx.x * NRM_WARPS + warp; i < n; i += (uint64_t)gridDim.x * NRM_WARPS) {
        float* src = v + i * (uint64_t)ld;
        for (int j = lane; j < d; j += 32) row[j] = src[j];
        __syncwarp();
        float acc = 0.0f;
        if (lane == 0)
            for (int j = 0; j < d; ++j) acc = __fadd_rn(acc, __fmul_rn(row[j], row[j]));
        float mag = __fsqrt_rn(__shfl_sync(0xFFFFFFFFu, acc, 0));
        for (int j = lane; j < d; j += 32) src[j] = __fdiv_rn(row[j], mag);
        __syncwarp();
    }
}

int nidx_normalize_vectors(int32_t device, float* vectors, uint64_t n, int32_t d, int32_t ld, int mem, void* stream_) {
    int dc = 0;
    if (cudaGetDeviceCount(&dc) != cudaSuccess || dc == 0) return fail(NIDX_ENODEVICE, "no CUDA device (there is no CPU fallback)");
    if (d <= 0 || ld < d || !vectors) return fail(NIDX_EINVAL, "normalize: d %d, ld %d", d, ld);
    if ((size_t)d * 4 * NRM_WARPS > 200 * 1024) return fail(NIDX_EINVAL, "normalize: dimension %d too large", d);
    if (n == 0) return 0;
    CU(cudaSetDevice(device));
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    float* dv = vectors;
    void* staged = nullptr;
    size_t bytes = (size_t)n * ld * 4;
    if (mem == NIDX_MEM_HOST) {
        CU(cudaMalloc(&staged, bytes));
        dv = static_cast<float*>(staged);
        cudaError_t e = cudaMemcpyAsync(dv, vectors, bytes, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) { cudaFree(staged); return fail(NIDX_ECUDA, "normalize: H2D failed: %s", cudaGetErrorString(e)); }
    }
    int sm = 0;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device);
    size_t smem = (size_t)d * 4 * NRM_WARPS;
    cudaFuncSetAttribute(normalize_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    uint64_t want = (n + NRM_WARPS - 1) / NRM_WARPS;
    int grid = (int)std::min<uint64_t>(want, (uint64_t)sm * 16);
    normalize_rows_kernel<<<grid, NRM_WARPS * 32, smem, st>>>(dv, n, d, ld);
    LAUNCHED();
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && staged) e = cudaMemcpyAsync(vectors, dv, bytes, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && staged) e = cudaStreamSynchronize(st);
    if (staged) cudaFree(staged);
    if (e != cudaSuccess) return fail(NIDX_ECUDA, "normalize failed: %s", cudaGetErrorString(e));
    return 0;
}

void nidx_vec_close(nidx_vec_segment* s) {
    if (!s) return;
    cudaSetDevice(s->cfg.device);
    cudaDeviceSynchronize();
    free_graph(s);
    cudaFree(s->d_vecs); cudaFree(s->d_norms); cudaFree(s->d_par_of); cudaFree(s->d_par_first); cudaFree(s->d_alive);
    cudaFree(s->d_counters); cudaFree(s->d_work_counter); cudaFree(s->d_quant); cudaFree(s->d_par_keys); cudaFree(s->inv[0].d_post); cudaFree(s->inv[1].d_post);
    if (s->ev_k0) cudaEventDestroy(s->ev_k0);
    if (s->ev_k1) cudaEventDestroy(s->ev_k1);
    delete s;
}

uint64_t nidx_vec_len(const nidx_vec_segment* s) { return s ? s->n : 0; }
const float* nidx_vec_device_vectors(const nidx_vec_segment* s, int32_t* ld_out) {
    if (!s) return nullptr;
    if (ld_out) *ld_out = s->ld;
    return s->d_vecs;
}

int nidx_vec_set_alive(nidx_vec_segment* s, const uint64_t* alive_bits, int mem) {
    if (!s) return fail(NIDX_EINVAL, "null segment");
    CU(cudaSetDevice(s->cfg.device));
    if (!alive_bits) { cudaFree(s->d_alive); s->d_alive = nullptr; return 0; }
    size_t words = ((size_t)s->n_par + 63) / 64;
    if (!s->d_alive) CU(cudaMalloc(&s->d_alive, std::max<size_t>(words, 1) * 8 + 8));
    CU(cudaMemcpy(s->d_alive, alive_bits, words * 8, mem == NIDX_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    // the number of alive paragraphs: what the reference counts per request (segment.rs:531) when there is no filter
    std::vector<uint64_t> h(words);
    if (words) CU(cudaMemcpy(h.data(), s->d_alive, words * 8, cudaMemcpyDeviceToHost));
    uint64_t cnt = 0;
    for (size_t i = 0; i < words; ++i) {
        uint64_t v = h[i];
        if ((i + 1) * 64 > s->n_par) v &= s->n_par > i * 64 ? (~0ull >> (64 - (s->n_par - i * 64))) : 0ull;
        cnt += (uint64_t)__builtin_popcountll(v);
    }
    s->alive_count = cnt;
    return 0;
}

static bool use_hnsw_cost(size_t total_nodes, size_t matching_nodes, size_t top_k, size_t M, bool has_rabitq);
int nidx_use_hnsw(uint64_t total_nodes, uint64_t matching_nodes, uint64_t top_k, int has_rabitq, int m) {
    if (matching_nodes == 0 || top_k == 0 || m <= 0) return 0;
    return use_hnsw_cost((size_t)total_nodes, (size_t)matching_nodes, (size_t)top_k, (size_t)m, has_rabitq != 0) ? 1 : 0;
}

int nidx_vec_graph_dims(const nidx_vec_segment* s, int32_t* s0, int32_t* su, uint64_t* upper_rows, uint32_t* entry_node, uint32_t* entry_layer) {
    if (!s) return fail(NIDX_EINVAL, "null segment");
    if (!s->has_graph) return fail(NIDX_ESTATE, "segment has no HNSW graph");
    if (s0) *s0 = s->s0;
    if (su) *su = s->su;
    if (upper_rows) *upper_rows = s->upper_rows;
    if (entry_node) *entry_node = s->entry_node;
    if (entry_layer) *entry_layer = s->entry_layer;
    return 0;
}

int nidx_vec_set_graph(nidx_vec_segment* s, const uint8_t* level, const uint32_t* adj0, const float* w0, const uint32_t* adjU, const float* wU) {
    if (!s || !level || !adj0) return fail(NIDX_EINVAL, "null argument");
    CU(cudaSetDevice(s->cfg.device));
    for (uint64_t i = 0; i < s->n; ++i)
        if (level[i] >= HS_MAX_LAYERS) return fail(NIDX_EINVAL, "node %llu has level %d >= %d", (unsigned long long)i, level[i], HS_MAX_LAYERS);
    int r = alloc_graph(s, level);
    if (r) return r;
    size_t n0 = (size_t)s->n * s->s0, nu = (size_t)s->upper_rows * s->su;
    CU(cudaMemcpy(s->d_adj0, adj0, n0 * 4, cudaMemcpyHostToDevice));
    if (w0) CU(cudaMemcpy(s->d_w0, w0, n0 * 4, cudaMemcpyHostToDevice));
    if (nu && adjU) CU(cudaMemcpy(s->d_adjU, adjU, nu * 4, cudaMemcpyHostToDevice));
    if (nu && wU) CU(cudaMemcpy(s->d_wU, wU, nu * 4, cudaMemcpyHostToDevice));
    s->has_graph = true;
    return 0;
}

int nidx_vec_get_graph(const nidx_vec_segment* s, uint8_t* level, uint32_t* adj0, float* w0, uint32_t* adjU, float* wU) {
    if (!s) return fail(NIDX_EINVAL, "null segment");
    if (!s->has_graph) return fail(NIDX_ESTATE, "segment has no HNSW graph");
    CU(cudaSetDevice(s->cfg.device));
    CU(cudaDeviceSynchronize());
    size_t n0 = (size_t)s->n * s->s0, nu = (size_t)s->upper_rows * s->su;
    if (level) memcpy(level, s->h_level.data(), s->n);
    if (adj0) CU(cudaMemcpy(adj0, s->d_adj0, n0 * 4, cudaMemcpyDeviceToHost));
    if (w0) CU(cudaMemcpy(w0, s->d_w0, n0 * 4, cudaMemcpyDeviceToHost));
    if (adjU && nu) CU(cudaMemcpy(adjU, s->d_adjU, nu * 4, cudaMemcpyDeviceToHost));
    if (wU && nu) CU(cudaMemcpy(wU, s->d_wU, nu * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int nidx_vec_counters_ex(nidx_vec_segment* s, uint64_t out[6]) {
    if (!s || !out) return fail(NIDX_EINVAL, "null argument");
    CU(cudaSetDevice(s->cfg.device));
    unsigned long long h[8];
    unsigned long long* src = s->last_counters.load();
    CU(cudaMemcpy(h, src ? src : s->d_counters, sizeof(h), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 6; ++i) out[i] = h[i];
    return 0;
}

int nidx_vec_counters(nidx_vec_segment* s, uint64_t out[3]) {
    if (!s) return fail(NIDX_EINVAL, "null segment");
    CU(cudaSetDevice(s->cfg.device));
    unsigned long long h[4];
    unsigned long long* src = s->last_counters.load();
    CU(cudaMemcpy(h, src ? src : s->d_counters, sizeof(h), cudaMemcpyDeviceToHost));
    out[0] = h[0]; out[1] = h[1]; out[2] = h[2] + h[3];
    return 0;
}

// ---- RaBitQ --------------------------------------------------------------------------------------
static int rabitq_check(const nidx_vec_segment* s) {
    if (s->cfg.similarity != NIDX_SIM_DOT || s->d % 64 != 0) return fail(NIDX_EINVAL, "RaBitQ needs Dot similarity and dimension %% 64 == 0 (config.rs:170-173)");
    if (s->d / 32 > RQ_MAX_WORDS32) return fail(NIDX_EINVAL, "RaBitQ: dimension above %d not supported", RQ_MAX_WORDS32 * 32);
    return 0;
}

int nidx_vec_rabitq_encode(nidx_vec_segment* s, void* stream_) {
    if (!s) return fail(NIDX_EINVAL, "null segment");
    int r = rabitq_check(s);
    if (r) return r;
    CU(cudaSetDevice(s->cfg.device));
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    s->quant_stride = rabitq_stride(s->d);
    if (!s->d_quant) CU(cudaMalloc(&s->d_quant, std::max<size_t>((size_t)s->n * s->quant_stride, 16)));
    if (s->n) {
        rabitq_encode_kernel<<<(unsigned)((s->n + 7) / 8), 256, 0, stream>>>(s->vdev(), s->d_quant, s->quant_stride);
        LAUNCHED();
        CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(stream));
    return 0;
}

int nidx_vec_rabitq_codes(const nidx_vec_segment* s, uint8_t* out) {
    if (!s || !out) return fail(NIDX_EINVAL, "null argument");
    if (!s->d_quant) return fail(NIDX_ESTATE, "segment has no RaBitQ codes (call nidx_vec_rabitq_encode)");
    CU(cudaSetDevice(s->cfg.device));
    size_t rec = (size_t)s->d / 8 + 8;
    CU(cudaMemcpy2D(out, rec, s->d_quant, (size_t)s->quant_stride, rec, (size_t)s->n, cudaMemcpyDeviceToHost));
    return 0;
}

// queries -> padded device copy + RaBitQ planes/params in the workspace; returns device pointers
static int rabitq_prepare_queries(nidx_vec_segment* s, Workspace& w, const float* queries, int nq, int ldq, bool host, cudaStream_t stream, float** dq_out,
                                  uint32_t** planes_out, RabitqQueryParams** params_out) {
    ENSURE(w.queries, (size_t)nq * s->ld * 4);
    float* dq = w.queries.as<float>();
    if (ldq == s->ld) {
        CU(cudaMemcpyAsync(dq, queries, (size_t)nq * ldq * 4, host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, stream));
    } else {
        const unsigned char* src = reinterpret_cast<const unsigned char*>(queries);
        if (host) {
            ENSURE(w.misc, (size_t)nq * ldq * 4);
            CU(cudaMemcpyAsync(w.misc.p, queries, (size_t)nq * ldq * 4, cudaMemcpyHostToDevice, stream));
            src = w.misc.as<unsigned char>();
        }
        pad_rows_kernel<<<std::min(nq, 1024), 256, 0, stream>>>(src, (size_t)ldq * 4, s->d, dq, s->ld, (uint64_t)nq);
        LAUNCHED();
    }
    size_t plane_bytes = (size_t)nq * 4 * (s->d / 32) * 4;
    ENSURE(w.qnorms, plane_bytes + (size_t)nq * sizeof(RabitqQueryParams) + 64);
    uint32_t* planes = w.qnorms.as<uint32_t>();
    RabitqQueryParams* params = reinterpret_cast<RabitqQueryParams*>(w.qnorms.as<unsigned char>() + ((plane_bytes + 15) / 16) * 16);
    rabitq_query_kernel<<<(nq + 7) / 8, 256, 0, stream>>>(dq, s->ld, s->d, nq, planes, params);
    LAUNCHED();
    *dq_out = dq; *planes_out = planes; *params_out = params;
    return 0;
}

int nidx_vec_rabitq_estimate(nidx_vec_segment* s, const float* queries, int32_t nq, int32_t ldq, int mem, float* out_est, float* out_err, void* stream_) {
    if (!s || !queries || !out_est || !out_err || nq <= 0) return fail(NIDX_EINVAL, "bad argument");
    if (!s->d_quant) return fail(NIDX_ESTATE, "segment has no RaBitQ codes (call nidx_vec_rabitq_encode)");
    if (ldq < s->d) return fail(NIDX_EINVAL, "query dimension %d != index dimension %d", ldq, s->d);
    CU(cudaSetDevice(s->cfg.device));
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    WsGuard g(s->pool, stream);
    Workspace& w = *g.w;
    bool host = mem == NIDX_MEM_HOST;
    float* dq; uint32_t* planes; RabitqQueryParams* params;
    int r = rabitq_prepare_queries(s, w, queries, nq, ldq, host, stream, &dq, &planes, &params);
    if (r) return r;
    float *d_est = out_est, *d_err = out_err;
    if (host) {
        ENSURE(w.scores, (size_t)nq * s->n * 8);
        d_est = w.scores.as<float>(); d_err = d_est + (size_t)nq * s->n;
    }
    if (s->n) {
        rabitq_estimate_kernel<<<dim3((unsigned)((s->n + 255) / 256), nq), 256, 0, stream>>>(s->d_quant, s->quant_stride, (uint32_t)s->n, s->d, planes, params, d_est, d_err);
        LAUNCHED();
        CU(cudaGetLastError());
    }
    if (host) {
        CU(cudaMemcpyAsync(out_est, d_est, (size_t)nq * s->n * 4, cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(out_err, d_err, (size_t)nq * s->n * 4, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
    }
    return 0;
}

int nidx_vec_last_kernel_ms(nidx_vec_segment* s, float* ms) {
    if (!s || !ms) return fail(NIDX_EINVAL, "null argument");
    CU(cudaSetDevice(s->cfg.device));
    CU(cudaEventSynchronize(s->ev_k1));
    CU(cudaEventElapsedTime(ms, s->ev_k0, s->ev_k1));
    return 0;
}

// ---- search -----------------------------------------------------------------------------------
// segment.rs:626-660: estimated vector evaluations of the HNSW walk vs the exhaustive scan; with RaBitQ codes a raw
// vector costs 16 quantised ones, layer 0 is searched for RERANKING_FACTOR * 3/4 times more nodes and
// RERANKING_FACTOR / 2 candidates per result are reranked (rabitq.rs:34).
static bool use_hnsw_cost(size_t total_nodes, size_t matching_nodes, size_t top_k, size_t M, bool has_rabitq) {
    const size_t RERANKING_FACTOR = 100;
    size_t full_cost = has_rabitq ? 16 : 1, search_mult = has_rabitq ? RERANKING_FACTOR * 3 / 4 : 1, rerank_mult = has_rabitq ? RERANKING_FACTOR / 2 : 0;
    float l = logf((float)total_nodes) - 2.0f;
    float hnsw_rq = l * l * logf((float)top_k) * (float)search_mult;
    size_t hnsw_full = top_k * rerank_mult + top_k * M * total_nodes / std::max<size_t>(matching_nodes, 1);
    size_t hnsw_cost = (size_t)(hnsw_rq < 0 ? 0 : hnsw_rq) + hnsw_full * full_cost;   // `as usize` saturates a negative estimate to 0
    size_t bf_cost = matching_nodes + top_k * rerank_mult * full_cost;
    return hnsw_cost < bf_cost;
}

__global__ void and_bits_kernel(const uint64_t* a, const uint64_t* b, uint64_t* out, size_t words, unsigned long long* count) {
    unsigned long long local = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t v = a[i] & (b ? b[i] : ~0ull);
        out[i] = v;
        local += __popcll(v);
    }
    for (int off = 16; off >= 1; off >>= 1) local += __shfl_xor_sync(0xFFFFFFFFu, local, off);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}

// hnsw_search_kernel is instantiated for the common row lengths (ld = NG * 128 floats); other dimensions use
// the run-time loop (NG = 0).
typedef void (*hs_kernel_t)(VecDev, GraphDev, SearchArgs);
static hs_kernel_t pick_search_kernel(int ld) {
    if (ld % 128 == 0) switch (ld / 128) {
        case 1: return hnsw_search_kernel<1>;
        case 2: return hnsw_search_kernel<2>;
        case 3: return hnsw_search_kernel<3>;
        case 4: return hnsw_search_kernel<4>;
        case 6: return hnsw_search_kernel<6>;
        case 8: return hnsw_search_kernel<8>;
        default: break;
    }
    return hnsw_search_kernel<0>;
}

// the 4-warp shape (two rows per warp in flight, 7 CTAs per SM): see hnsw_search_kernel
static hs_kernel_t pick_search_kernel_w4(int ld) {
    if (ld % 128 == 0) switch (ld / 128) {
        case 1: return hnsw_search_kernel<1, 4, true>;
        case 2: return hnsw_search_kernel<2, 4, true>;
        case 3: return hnsw_search_kernel<3, 4, true>;
        case 4: return hnsw_search_kernel<4, 4, true>;
        case 6: return hnsw_search_kernel<6, 4, true>;
        case 8: return hnsw_search_kernel<8, 4, true>;
        default: break;
    }
    return nullptr;
}

static hs_kernel_t pick_rabitq_walk_kernel_w4(int ld) {
    if (ld % 128 == 0) switch (ld / 128) {
        case 2: return hnsw_rabitq_kernel<2, 4>;
        case 3: return hnsw_rabitq_kernel<3, 4>;
        case 4: return hnsw_rabitq_kernel<4, 4>;
        case 6: return hnsw_rabitq_kernel<6, 4>;
        case 8: return hnsw_rabitq_kernel<8, 4>;
        default: break;
    }
    return hnsw_rabitq_kernel<0, 4>;
}

static hs_kernel_t pick_rabitq_walk_kernel(int ld) {
    if (ld % 128 == 0) switch (ld / 128) {
        case 2: return hnsw_rabitq_kernel<2>;
        case 3: return hnsw_rabitq_kernel<3>;
        case 4: return hnsw_rabitq_kernel<4>;
        case 6: return hnsw_rabitq_kernel<6>;
        case 8: return hnsw_rabitq_kernel<8>;
        default: break;
    }
    return hnsw_rabitq_kernel<0>;
}

typedef void (*scan_kernel_t)(VecDev, const float*, const float*, int, int, float*);
static scan_kernel_t pick_scan_kernel(int ld) {
    if (ld % 128 == 0) switch (ld / 128) {
        case 1: return scan_scores_kernel_t<1, 4>;
        case 2: return scan_scores_kernel_t<2, 4>;
        case 3: return scan_scores_kernel_t<3, 4>;
        case 4: return scan_scores_kernel_t<4, 2>;
        case 6: return scan_scores_kernel_t<6, 2>;
        case 8: return scan_scores_kernel_t<8, 2>;
        default: break;
    }
    return scan_scores_kernel;
}

static int hnsw_search_smem(const nidx_vec_segment* s, int ef0, int k, int* list_cap, int* cu_cap, int* hash_bits, size_t* bytes) {
    // closest_up_nodes pops at most k-1 candidates before it has k results when nothing is filtered
    // (search.rs:205-216), each adding at most one adjacency row of pending candidates.
    int cu = std::min(std::max(ef0 + k * s->s0, 2 * ef0), 4096);
    int lc = std::max(ef0, cu);
    // visited-set slots: a walk visits about ef0 * s0 * 0.6 nodes on easy data and up to ~1.3x that on clustered data at small ef
    // (10 M x 768, 4096 centres, ef = 30: 13 of 1024 queries overflowed 2048 slots): at least 4096, 1.5 x ef0 x s0 above that
    int slots = next_pow2(std::max(4096, (ef0 * s->s0 * 3) / 2));
    slots = std::max(slots, next_pow2(4 * lc));
    *list_cap = lc; *cu_cap = cu; *hash_bits = ilog2(slots);
    *bytes = hs_smem_bytes(s->ld, lc, *hash_bits);
    if (*bytes > 200 * 1024) return fail(NIDX_EINVAL, "HNSW search needs %zu bytes of shared memory (ef=%d, k=%d, dim=%d): too large", *bytes, ef0, k, s->d);
    return 0;
}

}  // extern "C"

// ---- filter formulas on the device (inverted_index/paragraph.rs:124-186) ---------------------------------------------
// ranges[2r], ranges[2r + 1] = [begin, end) into the postings of one inverted index; one block per range sets the bits
__global__ void bits_scatter_kernel(const uint32_t* __restrict__ postings, const uint64_t* __restrict__ ranges, uint64_t* __restrict__ out) {
    uint64_t b = ranges[2 * blockIdx.x], e = ranges[2 * blockIdx.x + 1];
    for (uint64_t i = b + threadIdx.x; i < e; i += blockDim.x) {
        uint32_t p = postings[i];
        atomicOr(reinterpret_cast<unsigned long long*>(out) + (p >> 6), 1ull << (p & 63));
    }
}
// op 0: a &= b, 1: a |= b, 2: a = ~a (bits beyond n_bits stay clear)
__global__ void bits_combine_kernel(uint64_t* __restrict__ a, const uint64_t* __restrict__ b, size_t words, uint64_t n_bits, int op) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t v = a[i];
        if (op == 0) v &= b[i];
        else if (op == 1) v |= b[i];
        else {
            v = ~v;
            if ((i + 1) * 64 > n_bits) v &= n_bits > i * 64 ? (~0ull >> (64 - (n_bits - i * 64))) : 0ull;
        }
        a[i] = v;
    }
}

struct FilterEval {
    nidx_vec_segment* s;
    const nidx_filter_node* nodes;
    int n_nodes;
    uint64_t* bufs;          // [n_nodes + 1][words] device
    size_t words;
    cudaStream_t stream;
    std::vector<uint64_t> ranges;     // all atoms' ranges, uploaded once
    std::vector<std::pair<size_t, size_t>> atom_ranges;   // per atom node: [first, count) in `ranges`
    uint64_t* d_ranges = nullptr;
    int next_buf = 0;

    static int cmp_key(const unsigned char* a, size_t la, const unsigned char* b, size_t lb) {
        int c = memcmp(a, b, std::min(la, lb));
        return c ? c : (la < lb ? -1 : (la > lb ? 1 : 0));
    }
    // first key >= q
    static uint32_t lower_bound(const nidx_vec_segment::InvIndex& ix, const unsigned char* q, size_t lq) {
        uint32_t lo = 0, hi = ix.n_keys;
        while (lo < hi) {
            uint32_t mid = (lo + hi) / 2;
            if (cmp_key(ix.key_bytes.data() + ix.key_off[mid], ix.key_off[mid + 1] - ix.key_off[mid], q, lq) < 0) lo = mid + 1; else hi = mid;
        }
        return lo;
    }
    // host pass: the key lookups (the fst's job in the reference), in node order
    int collect(int i) {
        const nidx_filter_node& nd = nodes[i];
        if (nd.kind == NIDX_F_LABEL || nd.kind == NIDX_F_KEYS) {
            const nidx_vec_segment::InvIndex& ix = s->inv[nd.kind == NIDX_F_LABEL ? NIDX_INV_LABELS : NIDX_INV_FIELDS];
            size_t first = ranges.size() / 2;
            for (int j = 0; j < nd.n; ++j) {
                const unsigned char* q = nd.keys[j];
                size_t lq = nd.key_len[j];
                uint32_t lo = lower_bound(ix, q, lq), hi = lo;
                if (nd.kind == NIDX_F_LABEL) {   // get_prefix: every key that starts with q
                    uint32_t a = lo, b = ix.n_keys;
                    while (a < b) {
                        uint32_t mid = (a + b) / 2;
                        size_t lk = ix.key_off[mid + 1] - ix.key_off[mid];
                        bool starts = lk >= lq && memcmp(ix.key_bytes.data() + ix.key_off[mid], q, lq) == 0;
                        if (starts) a = mid + 1; else b = mid;
                    }
                    hi = a;
                } else if (lo < ix.n_keys && cmp_key(ix.key_bytes.data() + ix.key_off[lo], ix.key_off[lo + 1] - ix.key_off[lo], q, lq) == 0) {
                    hi = lo + 1;                  // get: the exact key
                }
                if (hi > lo && ix.post_off[hi] > ix.post_off[lo]) { ranges.push_back(ix.post_off[lo]); ranges.push_back(ix.post_off[hi]); }
            }
            atom_ranges[i] = {first, ranges.size() / 2 - first};
            return i + 1;
        }
        int j = i + 1;
        for (int c = 0; c < nd.n; ++c) { if (j >= n_nodes) return -1; j = collect(j); if (j < 0) return -1; }
        return j;
    }
    // device pass: returns the buffer holding node i's bitset, *next = the node after its subtree
    uint64_t* eval(int i, int* next) {
        const nidx_filter_node& nd = nodes[i];
        uint64_t* out = bufs + (size_t)(next_buf++) * words;
        if (nd.kind == NIDX_F_LABEL || nd.kind == NIDX_F_KEYS) {
            cudaMemsetAsync(out, 0, words * 8, stream);
            auto ar = atom_ranges[i];
            if (ar.second) {
                const nidx_vec_segment::InvIndex& ix = s->inv[nd.kind == NIDX_F_LABEL ? NIDX_INV_LABELS : NIDX_INV_FIELDS];
                bits_scatter_kernel<<<(unsigned)ar.second, 128, 0, stream>>>(ix.d_post, d_ranges + 2 * ar.first, out);
                LAUNCHED();
            }
            *next = i + 1;
            return out;
        }
        int j = i + 1;
        uint64_t* acc = nullptr;
        unsigned blocks = (unsigned)std::min<size_t>((words + 255) / 256, 1024);
        for (int c = 0; c < nd.n; ++c) {
            int nx;
            uint64_t* child = eval(j, &nx);
            j = nx;
            if (!acc) { cudaMemcpyAsync(out, child, words * 8, cudaMemcpyDeviceToDevice, stream); acc = out; }
            else {
                bits_combine_kernel<<<blocks, 256, 0, stream>>>(acc, child, words, s->n_par, nd.kind == NIDX_F_OR ? 1 : 0);   // Not | And => intersect (paragraph.rs:160-164)
                LAUNCHED();
            }
        }
        if (nd.kind == NIDX_F_NOT) { bits_combine_kernel<<<blocks, 256, 0, stream>>>(acc, nullptr, words, s->n_par, 2); LAUNCHED(); }
        *next = j;
        return acc;
    }
};

// ParagraphInvertedIndexes::filter on the device: nodes (pre-order; several top-level nodes are not allowed: wrap them in an AND /
// OR node) -> bitset in `w.filter` (first `words` words), returns the device pointer in *out
static int filter_formula_device(nidx_vec_segment* s, Workspace& w, const nidx_filter_node* nodes, int n_nodes, cudaStream_t stream, uint64_t** out) {
    if (!nodes || n_nodes <= 0) return fail(NIDX_EINVAL, "empty filter formula");
    size_t words = ((size_t)s->n_par + 63) / 64;
    for (int i = 0; i < n_nodes; ++i) {
        int kd = nodes[i].kind;
        if (kd < NIDX_F_LABEL || kd > NIDX_F_NOT || nodes[i].n < 0) return fail(NIDX_EINVAL, "filter node %d: bad kind / count", i);
        if ((kd == NIDX_F_LABEL || kd == NIDX_F_KEYS) && nodes[i].n > 0 && (!nodes[i].keys || !nodes[i].key_len)) return fail(NIDX_EINVAL, "filter node %d: null keys", i);
        if ((kd == NIDX_F_AND || kd == NIDX_F_OR || kd == NIDX_F_NOT) && nodes[i].n < 1) return fail(NIDX_EINVAL, "filter node %d: a compound clause needs operands", i);
    }
    FilterEval ev;
    ev.s = s; ev.nodes = nodes; ev.n_nodes = n_nodes; ev.words = words; ev.stream = stream;
    ev.atom_ranges.assign(n_nodes, {0, 0});
    if (ev.collect(0) != n_nodes) return fail(NIDX_EINVAL, "malformed filter formula (operand counts do not add up to %d nodes)", n_nodes);
    size_t range_bytes = ev.ranges.size() * 8;
    ENSURE(w.filter, ((size_t)n_nodes + 3) * words * 8 + 64 + range_bytes + 64);
    uint64_t* base = w.filter.as<uint64_t>();
    ev.bufs = base + 2 * words + 8;          // the first 2 * words + 8 words are vec_search_impl's (filter AND alive, count)
    ev.d_ranges = ev.bufs + ((size_t)n_nodes + 1) * words;
    if (range_bytes) CU(cudaMemcpyAsync(ev.d_ranges, ev.ranges.data(), range_bytes, cudaMemcpyHostToDevice, stream));
    int next = 0;
    uint64_t* res = ev.eval(0, &next);
    CU(cudaGetLastError());
    if (range_bytes) CU(cudaStreamSynchronize(stream));   // `ranges` is a host temporary
    *out = res;
    return 0;
}

// The tensor-core filter + refine path serves large batches of small-k queries on single-vector segments whose rows are whole
// 128-byte swizzle rows.  NIDX_B200_SCAN=exact forces the CUDA-core kernels (same results, bit for bit), =tensor forces the filter.
static bool use_tc_filter(const nidx_vec_segment* s, int nq, int k) {
    if (k > TC2_KMAX || s->ld % TC2_KB != 0 || s->d_par_first || s->n < (uint64_t)TC2_N || s->cfg.similarity == NIDX_SIM_L2) return false;
    const char* e = getenv("NIDX_B200_SCAN");
    if (e && !strcmp(e, "exact")) return false;
    if (e && !strcmp(e, "tensor")) return true;
    return nq >= 64;
}

// The body of nidx_vec_search.  qhost: the queries (and filter bits) are host pointers; ohost: the outputs are host pointers
// (copied back and the stream synchronised before returning).  The sharded entry point (shard.cuh) passes host queries with
// device outputs: the partial results go straight into the exchange buffer.
static int vec_search_impl(nidx_vec_segment* s, const float* queries, int32_t nq, int32_t ldq, bool qhost, bool ohost, const nidx_vec_search_params* p,
                           uint32_t* out_ids, float* out_scores, int32_t* out_counts, cudaStream_t stream, const nidx_filter_node* formula = nullptr,
                           int32_t n_formula = 0) {
    if (!s || !p || (!queries && nq > 0) || !out_ids || !out_scores) return fail(NIDX_EINVAL, "null argument");
    if (nq <= 0) return 0;
    if (ldq < s->d) return fail(NIDX_EINVAL, "query dimension %d != index dimension %d (VectorErr::InconsistentDimensions)", ldq, s->d);
    int k = p->k;
    if (k <= 0) return fail(NIDX_EINVAL, "k must be positive");
    CU(cudaSetDevice(s->cfg.device));
    WsGuard g(s->pool, stream);
    Workspace& w = *g.w;
    bool host = qhost;
    VecDev V = s->vdev();

    // queries -> [nq][ld] zero padded on device, norms
    ENSURE(w.queries, (size_t)nq * s->ld * 4);
    ENSURE(w.qnorms, (size_t)nq * 4);
    float* dq = w.queries.as<float>();
    if (ldq == s->ld) {
        CU(cudaMemcpyAsync(dq, queries, (size_t)nq * ldq * 4, host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, stream));
    } else {
        const unsigned char* src = reinterpret_cast<const unsigned char*>(queries);
        if (host) {
            ENSURE(w.misc, (size_t)nq * ldq * 4);
            CU(cudaMemcpyAsync(w.misc.p, queries, (size_t)nq * ldq * 4, cudaMemcpyHostToDevice, stream));
            src = w.misc.as<unsigned char>();
        }
        pad_rows_kernel<<<std::min(nq, 1024), 256, 0, stream>>>(src, (size_t)ldq * 4, s->d, dq, s->ld, (uint64_t)nq);
        LAUNCHED();
    }
    if (s->cfg.similarity != NIDX_SIM_DOT) {
        row_norms_kernel<<<(nq + 7) / 8, 256, 0, stream>>>(dq, s->ld, (uint64_t)nq, w.qnorms.as<float>());
        LAUNCHED();
    }

    // filter ∧ alive (segment.rs:516-534)
    const uint64_t* bits = s->d_alive;
    size_t words = ((size_t)s->n_par + 63) / 64;
    uint64_t matching = s->d_alive ? s->alive_count : s->n_par;
    const bool filtered = p->filter_bits || formula;
    if (filtered) {
        const uint64_t* fsrc = p->filter_bits;
        bool fhost = host;
        if (formula) {   // the formula is evaluated on the device (inverted_index/paragraph.rs:124-186): no host bitset, no copy
            uint64_t* fdev = nullptr;
            int fr = filter_formula_device(s, w, formula, n_formula, stream, &fdev);
            if (fr) return fr;
            fsrc = fdev; fhost = false;
        } else ENSURE(w.filter, words * 8 * 2 + 64);
        uint64_t* d_in = w.filter.as<uint64_t>();
        uint64_t* d_out = d_in + words;
        unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(d_out + words);
        if (fhost) { CU(cudaMemcpyAsync(d_in, p->filter_bits, words * 8, cudaMemcpyHostToDevice, stream)); fsrc = d_in; }
        CU(cudaMemsetAsync(d_cnt, 0, 8, stream));
        and_bits_kernel<<<std::min<size_t>((words + 255) / 256, 1024), 256, 0, stream>>>(fsrc, s->d_alive, d_out, words, d_cnt);
        LAUNCHED();
        bits = d_out;
        matching = formula ? 0 : p->filter_matching;
        if (matching == 0) {   // segment.rs:531: the reference counts the matches of every filtered request (8 bytes back, one sync)
            unsigned long long h = 0;
            CU(cudaMemcpyAsync(&h, d_cnt, 8, cudaMemcpyDeviceToHost, stream));
            CU(cudaStreamSynchronize(stream));
            matching = h;
        }
    }

    if (matching == 0) {   // segment.rs:532-534: nothing can match (everything deleted / filtered out) -> empty, whatever the method
        uint32_t* e_ids = out_ids; float* e_sc = out_scores; int* e_cnt = out_counts;
        if (ohost) {
            for (size_t i = 0; i < (size_t)nq * k; ++i) { e_ids[i] = NIDX_NIL; e_sc[i] = 0.0f; }
            if (e_cnt) for (int i = 0; i < nq; ++i) e_cnt[i] = 0;
            CU(cudaStreamSynchronize(stream));
        } else {
            CU(cudaMemsetAsync(e_ids, 0xFF, (size_t)nq * k * 4, stream));
            CU(cudaMemsetAsync(e_sc, 0, (size_t)nq * k * 4, stream));
            if (e_cnt) CU(cudaMemsetAsync(e_cnt, 0, (size_t)nq * 4, stream));
        }
        return 0;
    }
    int method = p->method;
    if (method == NIDX_METHOD_AUTO) {
        if (!s->has_graph) method = NIDX_METHOD_BRUTE;
        else if (matching == 0 && filtered) method = NIDX_METHOD_BRUTE;
        else {
            // a segment that carries codes is searched with a RaBitQ query on either path (segment.rs:506-513: `rabitq` =
            // has_quantized): the quantised walk (hnsw/search.rs:332-366) or the quantised scan (segment.rs:581-608)
            bool walk_ok = s->d_quant && s->cfg.similarity == NIDX_SIM_DOT;
            bool scan_ok = walk_ok && !s->d_par_first && k <= 1024;
            if (use_hnsw_cost(s->n_par, matching, (size_t)k, (size_t)s->cfg.m, walk_ok)) method = walk_ok ? NIDX_METHOD_HNSW_RABITQ : NIDX_METHOD_HNSW;
            else method = scan_ok ? NIDX_METHOD_BRUTE_RABITQ : NIDX_METHOD_BRUTE;
            // A walk keeps its list and visited set in shared memory: a very large top_k (the reference has no limit on it) does not
            // fit one CTA.  AUTO then takes the exhaustive scan -- exact results -- instead of failing the request.
            if (method == NIDX_METHOD_HNSW || method == NIDX_METHOD_HNSW_RABITQ) {
                bool fits;
                if (method == NIDX_METHOD_HNSW) {
                    int ef = p->ef > 0 ? p->ef : s->cfg.ef_search;
                    int cu = std::min(std::max(std::max(k, ef) + k * s->s0, 2 * std::max(k, ef)), 4096);
                    int lcap = std::max(std::max(k, ef), cu);
                    int slots = next_pow2(std::max(4096, (std::max(k, ef) * s->s0 * 3) / 2));
                    slots = std::max(slots, next_pow2(4 * lcap));
                    fits = hs_smem_bytes(s->ld, lcap, ilog2(slots)) <= 200 * 1024;
                } else {
                    int last_k = (int)std::min<size_t>((size_t)k * 100, 2000);
                    int cu_cap = std::min(std::max(k + k * s->s0, 2 * k), 4096);
                    fits = rq_smem_bytes(s->ld, s->d, std::max(last_k, cu_cap), ilog2(next_pow2(std::max(2048, 4 * cu_cap))), k) <= 200 * 1024;
                }
                if (!fits && k <= 1024) method = NIDX_METHOD_BRUTE;
            }
        }
    }
    if ((method == NIDX_METHOD_HNSW || method == NIDX_METHOD_HNSW_RABITQ) && !s->has_graph) return fail(NIDX_ESTATE, "HNSW search requested but the segment has no graph");

    // outputs
    uint32_t* d_ids = out_ids; float* d_sc = out_scores; int* d_cnt = out_counts;
    if (ohost || !out_counts) {
        ENSURE(w.out_ids, (size_t)nq * k * 4);
        ENSURE(w.out_scores, (size_t)nq * k * 4);
        ENSURE(w.out_counts, (size_t)nq * 4);
        if (ohost) { d_ids = w.out_ids.as<uint32_t>(); d_sc = w.out_scores.as<float>(); }
        if (ohost || !out_counts) d_cnt = w.out_counts.as<int>();
    }

    if (method == NIDX_METHOD_BRUTE_RABITQ && s->n != 0) {
        // segment.rs:581-608 with SearchVector::RabitQ: estimate every vector from its 1-bit code, keep upper_bound >= min_score,
        // rerank_top with the raw vectors (sequential semantics preserved, see rabitq_rerank_kernel)
        int rr = rabitq_check(s);
        if (rr) return rr;
        if (!s->d_quant) return fail(NIDX_ESTATE, "segment has no RaBitQ codes (call nidx_vec_rabitq_encode)");
        if (s->d_par_first) return fail(NIDX_EINVAL, "RaBitQ scan of multi-vector paragraphs is not implemented");
        if (k > 1024) return fail(NIDX_EINVAL, "k above 1024 not supported");
        uint32_t* planes; RabitqQueryParams* params; float* dq2;
        rr = rabitq_prepare_queries(s, w, queries, nq, ldq, host, stream, &dq2, &planes, &params);
        if (rr) return rr;
        int qgroup = (int)std::max<size_t>(1, std::min<size_t>((size_t)nq, ((size_t)4 << 30) / ((size_t)s->n * 8)));
        ENSURE(w.scores, (size_t)qgroup * s->n * 8);
        size_t smem_rr = rr_smem_bytes(s->ld, k);
        CU(cudaFuncSetAttribute(rabitq_rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rr));
        for (int q0 = 0; q0 < nq; q0 += qgroup) {
            int nqg = std::min(qgroup, nq - q0);
            float* d_est = w.scores.as<float>();
            float* d_err = d_est + (size_t)nqg * s->n;
            if (q0 == 0) CU(cudaEventRecord(s->ev_k0, stream));
            rabitq_estimate_kernel<<<dim3((unsigned)((s->n + 255) / 256), nqg), 256, 0, stream>>>(s->d_quant, s->quant_stride, (uint32_t)s->n, s->d,
                                                                                                   planes + (size_t)q0 * 4 * (s->d / 32), params + q0, d_est, d_err);
            if (q0 == 0) CU(cudaEventRecord(s->ev_k1, stream));
            LAUNCHED();
            rabitq_rerank_kernel<<<nqg, RR_THREADS, smem_rr, stream>>>(V, dq2 + (size_t)q0 * s->ld, d_est, d_err, bits, p->min_score, k, d_ids + (size_t)q0 * k,
                                                                       d_sc + (size_t)q0 * k, d_cnt + q0, nullptr);
            LAUNCHED();
        }
        CU(cudaGetLastError());
    } else if (s->n == 0) {
        CU(cudaMemsetAsync(d_ids, 0xFF, (size_t)nq * k * 4, stream));
        CU(cudaMemsetAsync(d_sc, 0, (size_t)nq * k * 4, stream));
        CU(cudaMemsetAsync(d_cnt, 0, (size_t)nq * 4, stream));
    } else if (method == NIDX_METHOD_BRUTE && use_tc_filter(s, nq, k)) {
        // large batch: TF32 tensor-core filter + bit-exact refine (scan_tc2.cuh); nothing of size [Q x N] touches HBM
        int n_chunks = (int)((s->n + TC2_CHUNK - 1) / TC2_CHUNK), n_qblocks = (nq + TC2_M - 1) / TC2_M;
        {
            std::lock_guard<std::mutex> lk(s->map_mu);
            if (!s->map_v_ready) {
                int mr = make_row_tensor_map(&s->map_v, s->d_vecs, s->n, s->ld, TC2_N);
                if (mr) return mr;
                s->map_v_ready = true;
            }
        }
        CUtensorMap map_q;
        int mr = make_row_tensor_map(&map_q, dq, (uint64_t)nq, s->ld, TC2_M);
        if (mr) return mr;
        int grid = std::min(n_chunks * n_qblocks, s->sm_count);
        int slots = std::max(1, std::min(grid / n_qblocks, n_chunks));   // CTAs per query block; 1 when there are more blocks than CTAs
        size_t cand_n = (size_t)nq * slots * TC2_LISTS * TC2_L;
        ENSURE(w.scores, cand_n * 8 + 64);
        ENSURE(w.sched, 128);
        unsigned long long* call_counters = reinterpret_cast<unsigned long long*>(w.sched.as<unsigned char>() + 64);
        Tc2Args ta;
        ta.nq = nq; ta.n_qblocks = n_qblocks; ta.n_chunks = n_chunks; ta.slots = slots; ta.qnorms = w.qnorms.as<float>(); ta.bits = bits;
        ta.cand_score = w.scores.as<float>(); ta.cand_id = reinterpret_cast<uint32_t*>(w.scores.as<float>() + cand_n);
        ta.work_counter = w.sched.as<unsigned int>();
        CU(cudaFuncSetAttribute(scan_tc_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC2_SMEM_BYTES));
        grid = n_qblocks <= grid ? n_qblocks * slots : grid;
        CU(cudaMemsetAsync(w.sched.p, 0, 128, stream));
        s->last_counters.store(call_counters);
        CU(cudaEventRecord(s->ev_k0, stream));
        scan_tc_filter_kernel<<<grid, TC2_THREADS, TC2_SMEM_BYTES, stream>>>(map_q, s->map_v, V, ta);
        CU(cudaEventRecord(s->ev_k1, stream));
        LAUNCHED();
        int cap = topk_cap(k, 256);
        size_t smem_rf = tc2_refine_smem(s->ld, cap);
        if (smem_rf > 48 * 1024) CU(cudaFuncSetAttribute(scan_tc_refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rf));
        scan_tc_refine_kernel<<<nq, 256, smem_rf, stream>>>(V, dq, w.qnorms.as<float>(), slots * TC2_LISTS, ta.cand_score, ta.cand_id, bits, s->max_norm, p->min_score, k, cap,
                                                            d_ids, d_sc, d_cnt, call_counters + 6);
        LAUNCHED();
        CU(cudaGetLastError());
    } else if (method == NIDX_METHOD_BRUTE) {
        if (k > 1024) return fail(NIDX_EINVAL, "brute-force k above 1024 not supported");
        int cap = topk_cap(k, 256);
        int n_chunks = (int)std::min<uint64_t>(std::max<uint64_t>(1, ((uint64_t)s->sm_count * 4 + nq - 1) / nq), (s->n_par + 4095) / 4096);
        n_chunks = std::max(n_chunks, 1);
        // bound the score matrix: process queries in groups
        size_t max_score_bytes = (size_t)4 << 30;
        int qgroup = (int)std::max<size_t>(1, std::min<size_t>((size_t)nq, max_score_bytes / ((size_t)s->n * 4)));
        qgroup = std::max(SCAN_QT, qgroup / SCAN_QT * SCAN_QT);
        qgroup = std::min(qgroup, (nq + SCAN_QT - 1) / SCAN_QT * SCAN_QT);
        ENSURE(w.scores, (size_t)qgroup * s->n * 4);
        ENSURE(w.partial, (size_t)qgroup * n_chunks * k * 8);
        size_t smem_scan = (size_t)SCAN_QT * s->ld * 4;
        scan_kernel_t scan_kern = pick_scan_kernel(s->ld);
        // Batches of >= 128 queries go to the tensor cores (scores within ~3e-7 of the lane-blocked order);
        // NIDX_B200_SCAN=exact forces the bit-exact CUDA-core kernel, =tensor forces the tensor path.
        const char* scan_env = getenv("NIDX_B200_SCAN");
        bool tensor_scan = s->ld % TC_KB == 0 && scan_env && !strcmp(scan_env, "tensor3x");   // round 1's 3xTF32 score-matrix kernel: on request only
        if (tensor_scan) CU(cudaFuncSetAttribute(scan_scores_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
        if (smem_scan > 48 * 1024) CU(cudaFuncSetAttribute(scan_kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_scan));
        if ((size_t)cap * 8 > 48 * 1024) {
            CU(cudaFuncSetAttribute(scan_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap * 8));
            CU(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap * 8));
        }
        for (int q0 = 0; q0 < nq; q0 += qgroup) {
            int nqg = std::min(qgroup, nq - q0);
            int n_qtiles = (nqg + SCAN_QT - 1) / SCAN_QT;
            uint64_t n_vchunks = (s->n + SCAN_WARPS * SCAN_VPW - 1) / (SCAN_WARPS * SCAN_VPW);
            uint64_t grid = n_vchunks * n_qtiles;
            if (grid > 0x7FFFFFFFull) return fail(NIDX_EINVAL, "scan grid too large");
            if (q0 == 0) CU(cudaEventRecord(s->ev_k0, stream));
            if (tensor_scan) {
                // large batch: the score matrix is a dense GEMM -> tcgen05 (3xTF32, f32 accumulate in TMEM)
                dim3 tgrid((unsigned)((s->n + TC_N - 1) / TC_N), (unsigned)((nqg + TC_M - 1) / TC_M));
                scan_scores_tc_kernel<<<tgrid, TC_THREADS, TC_SMEM_BYTES, stream>>>(V, dq + (size_t)q0 * s->ld, w.qnorms.as<float>() + q0, nqg, w.scores.as<float>());
            } else {
                scan_kern<<<(unsigned)grid, SCAN_WARPS * 32, smem_scan, stream>>>(V, dq + (size_t)q0 * s->ld, w.qnorms.as<float>() + q0, nqg, n_qtiles,
                                                                                     w.scores.as<float>());
            }
            if (q0 == 0) CU(cudaEventRecord(s->ev_k1, stream));
            LAUNCHED();
            scan_select_kernel<<<dim3(n_chunks, nqg), 256, (size_t)cap * 8, stream>>>(w.scores.as<float>(), (uint32_t)s->n, s->n_par, s->d_par_first, nullptr, bits,
                                                                                      p->min_score, k, cap, n_chunks, w.partial.as<uint64_t>());
            LAUNCHED();
            topk_merge_kernel<<<nqg, 256, (size_t)cap * 8, stream>>>(w.partial.as<uint64_t>(), n_chunks * k, k, cap, d_ids + (size_t)q0 * k, d_sc + (size_t)q0 * k,
                                                                    d_cnt + q0);
            LAUNCHED();
        }
        CU(cudaGetLastError());
    } else if (method == NIDX_METHOD_HNSW_RABITQ) {
        // hnsw/search.rs:306-383 with SearchVector::RabitQ: estimate-ranked walk, k * 100 layer-0 results, exact rerank + closest_up
        int rr = rabitq_check(s);
        if (rr) return rr;
        if (!s->d_quant) return fail(NIDX_ESTATE, "segment has no RaBitQ codes (call nidx_vec_rabitq_encode)");
        int nw = s->d / 32;
        size_t plane_bytes = (size_t)nq * 4 * nw * 4;
        ENSURE(w.misc, ((plane_bytes + 15) / 16) * 16 + (size_t)nq * sizeof(RabitqQueryParams) + 64);
        uint32_t* planes = w.misc.as<uint32_t>();
        RabitqQueryParams* params = reinterpret_cast<RabitqQueryParams*>(w.misc.as<unsigned char>() + ((plane_bytes + 15) / 16) * 16);
        rabitq_query_kernel<<<(nq + 7) / 8, 256, 0, stream>>>(dq, s->ld, s->d, nq, planes, params);
        LAUNCHED();
        int last_k = (int)std::min<size_t>((size_t)k * 100, 2000);              // rabitq.rs:34-36 RERANKING_FACTOR / RERANKING_LIMIT
        int cu_cap = std::min(std::max(k + k * s->s0, 2 * k), 4096);
        int list_cap = std::max(last_k, cu_cap);
        int slots = next_pow2(std::max(2048, 4 * cu_cap));
        int hash_bits = ilog2(slots);
        int gv_bits = 16;                                                       // 64 k slots: layer 0 visits ~10-20 k nodes at k = 10
        while ((1 << gv_bits) < 24 * last_k) ++gv_bits;
        if (const char* e = getenv("NIDX_B200_RQ_VISITED_BITS")) { int b = atoi(e); if (b >= 12 && b <= 22) gv_bits = b; }
        size_t smem = rq_smem_bytes(s->ld, s->d, list_cap, hash_bits, k);
        if (smem > 200 * 1024) return fail(NIDX_EINVAL, "quantised HNSW search needs %zu bytes of shared memory (k=%d, dim=%d): too large", smem, k, s->d);
        // CTA shape: the walk is bound by the ~1 000 dependent hops of a query, so what counts is how many queries are resident.
        // 8 warps per query: 4 CTAs per SM; 4 warps: 7 per SM.  The 4-warp shape is taken when the batch does not fit one wave of the
        // 8-warp shape (NIDX_B200_RQ_W = 4 / 8 forces one).
        hs_kernel_t kern = pick_rabitq_walk_kernel(s->ld);
        int threads = HS_THREADS;
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
        {
            const char* ew = getenv("NIDX_B200_RQ_W");
            int force = ew ? atoi(ew) : 0;
            if (force == 4 || (force != 8 && nq > std::max(1, occ) * s->sm_count)) {
                kern = pick_rabitq_walk_kernel_w4(s->ld);
                threads = 128;
                CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
            }
        }
        int grid = std::min(nq, std::max(1, occ) * s->sm_count);
        ENSURE(w.scores, ((size_t)grid << gv_bits) * 4);
        SearchArgs a;
        memset(&a, 0, sizeof(a));
        a.mode = 0; a.nq = nq; a.queries = dq; a.qnorms = w.qnorms.as<float>(); a.k = k; a.ef0 = last_k; a.min_score = p->min_score;
        a.with_duplicates = p->with_duplicates; a.multi_vector = s->cfg.multi_vector; a.filter = bits;
        a.out_ids = d_ids; a.out_scores = d_sc; a.out_counts = d_cnt;
        a.hash_bits = hash_bits; a.list_cap = list_cap; a.cu_cap = cu_cap;
        a.codes = s->d_quant; a.code_stride = s->quant_stride; a.planes = planes; a.qparams = params;
        a.gvisited = w.scores.as<uint32_t>(); a.gv_bits = gv_bits; a.last_k = last_k;
        { const char* ep = getenv("NIDX_B200_RQ_PREFETCH"); a.rq_prefetch = ep ? atoi(ep) : 1; }
        ENSURE(w.sched, 128);
        a.work_counter = w.sched.as<unsigned int>();
        a.counters = reinterpret_cast<unsigned long long*>(w.sched.as<unsigned char>() + 64);   // per call, in the call's workspace
        CU(cudaMemsetAsync(a.work_counter, 0, 128, stream));
        s->last_counters.store(a.counters);
        CU(cudaEventRecord(s->ev_k0, stream));
        kern<<<grid, threads, smem, stream>>>(V, s->gdev(), a);
        CU(cudaEventRecord(s->ev_k1, stream));
        LAUNCHED();
        CU(cudaGetLastError());
    } else {
        int ef = p->ef > 0 ? p->ef : s->cfg.ef_search;
        int ef0 = std::max(k, ef);  // search.rs:338-345
        int list_cap, cu_cap, hash_bits;
        size_t smem;
        int r = hnsw_search_smem(s, ef0, k, &list_cap, &cu_cap, &hash_bits, &smem);
        if (r) return r;
        SearchArgs a;
        memset(&a, 0, sizeof(a));
        a.mode = 0; a.nq = nq; a.queries = dq; a.qnorms = w.qnorms.as<float>(); a.k = k; a.ef0 = ef0; a.min_score = p->min_score;
        a.with_duplicates = p->with_duplicates; a.multi_vector = s->cfg.multi_vector; a.filter = bits;
        a.out_ids = d_ids; a.out_scores = d_sc; a.out_counts = d_cnt;
        a.hash_bits = hash_bits; a.list_cap = list_cap; a.cu_cap = cu_cap;
        // the scheduler counter lives in the workspace: concurrent calls must not share it
        ENSURE(w.sched, 128);
        a.work_counter = w.sched.as<unsigned int>();
        a.counters = reinterpret_cast<unsigned long long*>(w.sched.as<unsigned char>() + 64);   // per call, in the call's workspace
        CU(cudaMemsetAsync(a.work_counter, 0, 128, stream));
        s->last_counters.store(a.counters);
        hs_kernel_t kern = pick_search_kernel(s->ld);
        int threads = HS_THREADS;
        // Shape: 8 warps per query, 4 CTAs per SM -- or 4 warps with two rows in flight each, 7 CTAs per SM, when that lets the
        // whole batch run as one wave (NIDX_B200_HS_W=4 / 8 forces a shape; the experiment's visited table is hs_w4_bits slots).
        {
            const char* e1 = getenv("NIDX_B200_HS_W");
            const char* e2 = getenv("NIDX_B200_HS_W4_BITS");
            const int force_w = e1 ? atoi(e1) : 0, w4_bits = e2 ? atoi(e2) : 0;
            hs_kernel_t k4 = pick_search_kernel_w4(s->ld);
            if (k4 && force_w == 4) {
                int hb = w4_bits > 0 ? w4_bits : hash_bits;
                size_t smem4 = hs_smem_bytes(s->ld, list_cap, hb);
                kern = k4; threads = 128; smem = smem4; a.hash_bits = hb;
            }
        }
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
        int grid = std::min(nq, std::max(1, occ) * s->sm_count);
        CU(cudaEventRecord(s->ev_k0, stream));
        kern<<<grid, threads, smem, stream>>>(V, s->gdev(), a);
        CU(cudaEventRecord(s->ev_k1, stream));
        LAUNCHED();
        CU(cudaGetLastError());
    }

    if (ohost) {
        CU(cudaMemcpyAsync(out_ids, d_ids, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(out_scores, d_sc, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, stream));
        if (out_counts) CU(cudaMemcpyAsync(out_counts, d_cnt, (size_t)nq * 4, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
    }
    return 0;
}

extern "C" {

int nidx_vec_search(nidx_vec_segment* s, const float* queries, int32_t nq, int32_t ldq, int mem, const nidx_vec_search_params* p, uint32_t* out_ids,
                    float* out_scores, int32_t* out_counts, void* stream_) {
    bool host = mem == NIDX_MEM_HOST;
    return vec_search_impl(s, queries, nq, ldq, host, host, p, out_ids, out_scores, out_counts, reinterpret_cast<cudaStream_t>(stream_));
}

int nidx_vec_search_formula(nidx_vec_segment* s, const float* queries, int32_t nq, int32_t ldq, int mem, const nidx_vec_search_params* p,
                            const nidx_filter_node* nodes, int32_t n_nodes, uint32_t* out_ids, float* out_scores, int32_t* out_counts, void* stream_) {
    bool host = mem == NIDX_MEM_HOST;
    if (p && p->filter_bits) return fail(NIDX_EINVAL, "give either filter_bits or a formula");
    return vec_search_impl(s, queries, nq, ldq, host, host, p, out_ids, out_scores, out_counts, reinterpret_cast<cudaStream_t>(stream_), nodes, n_nodes);
}

int nidx_vec_set_inverted_index(nidx_vec_segment* s, int32_t which, uint32_t n_keys, const uint8_t* key_bytes, const uint64_t* key_off, const uint64_t* post_off,
                                const uint32_t* postings) {
    if (!s || (which != NIDX_INV_LABELS && which != NIDX_INV_FIELDS)) return fail(NIDX_EINVAL, "bad argument");
    if (n_keys && (!key_off || !post_off || (!key_bytes && key_off[n_keys]) || (!postings && post_off[n_keys]))) return fail(NIDX_EINVAL, "null argument");
    CU(cudaSetDevice(s->cfg.device));
    nidx_vec_segment::InvIndex& ix = s->inv[which];
    cudaFree(ix.d_post); ix.d_post = nullptr; ix.n_keys = 0;
    if (!n_keys) { ix.key_bytes.clear(); ix.key_off.assign(1, 0); ix.post_off.assign(1, 0); return 0; }
    for (uint32_t i = 0; i + 1 < n_keys; ++i)
        if (FilterEval::cmp_key(key_bytes + key_off[i], key_off[i + 1] - key_off[i], key_bytes + key_off[i + 1], key_off[i + 2] - key_off[i + 1]) >= 0)
            return fail(NIDX_EINVAL, "inverted index keys must be strictly ascending (key %u)", i + 1);
    uint64_t np = post_off[n_keys];
    for (uint64_t i = 0; i < np; ++i) if (postings[i] >= s->n_par) return fail(NIDX_EINVAL, "posting %llu: paragraph %u out of range", (unsigned long long)i, postings[i]);
    ix.key_bytes.assign(key_bytes, key_bytes + key_off[n_keys]);
    ix.key_off.assign(key_off, key_off + n_keys + 1);
    ix.post_off.assign(post_off, post_off + n_keys + 1);
    CU(cudaMalloc(&ix.d_post, std::max<uint64_t>(np, 1) * 4));
    if (np) CU(cudaMemcpy(ix.d_post, postings, np * 4, cudaMemcpyHostToDevice));
    ix.n_keys = n_keys;
    return 0;
}

int nidx_vec_filter(nidx_vec_segment* s, const nidx_filter_node* nodes, int32_t n_nodes, uint64_t* out_bits, int mem, uint64_t* out_matching, void* stream_) {
    if (!s) return fail(NIDX_EINVAL, "null segment");
    CU(cudaSetDevice(s->cfg.device));
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    WsGuard g(s->pool, stream);
    Workspace& w = *g.w;
    uint64_t* fdev = nullptr;
    int r = filter_formula_device(s, w, nodes, n_nodes, stream, &fdev);
    if (r) return r;
    size_t words = ((size_t)s->n_par + 63) / 64;
    uint64_t* d_out = w.filter.as<uint64_t>() + words;
    unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(d_out + words);
    CU(cudaMemsetAsync(d_cnt, 0, 8, stream));
    and_bits_kernel<<<std::min<size_t>((words + 255) / 256, 1024), 256, 0, stream>>>(fdev, s->d_alive, d_out, words, d_cnt);   // segment.rs:523-526: intersect with the alive set
    LAUNCHED();
    if (out_bits) CU(cudaMemcpyAsync(out_bits, d_out, words * 8, mem == NIDX_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, stream));
    unsigned long long h = 0;
    CU(cudaMemcpyAsync(&h, d_cnt, 8, cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    if (out_matching) *out_matching = h;
    return 0;
}

int nidx_merge_topk(int32_t device, const uint32_t* ids, const float* scores, int32_t n_parts, int64_t part_stride, int32_t nq, int32_t k,
                    uint32_t* out_ids, float* out_scores, int32_t* out_part, void* stream_) {
    int r = check_device(device);
    if (r) return r;
    if (!ids || !scores || !out_ids || !out_scores || n_parts <= 0 || nq <= 0 || k <= 0) return fail(NIDX_EINVAL, "bad argument");
    if (k > 1024 || (long long)n_parts * k >= (1ll << 31)) return fail(NIDX_EINVAL, "k above 1024 not supported");
    int cap = topk_cap(k, 256);
    if ((size_t)cap * 8 > 48 * 1024) CU(cudaFuncSetAttribute(parts_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap * 8));
    parts_merge_kernel<<<nq, 256, (size_t)cap * 8, reinterpret_cast<cudaStream_t>(stream_)>>>(ids, scores, n_parts, part_stride > 0 ? (size_t)part_stride : (size_t)nq * k, nq, k, cap, out_ids, out_scores, out_part);
    LAUNCHED();
    CU(cudaGetLastError());
    return 0;
}

// ---- build ------------------------------------------------------------------------------------
// Level RNG: build.rs:40,97-101.  rand 0.10 SmallRng = xoshiro256++ seeded by SplitMix64 [recalled].
static void host_assign_levels(uint64_t n, int M, uint64_t seed, uint8_t* level) {
    uint64_t st[4], state = seed;
    for (int i = 0; i < 4; ++i) {
        state += 0x9e3779b97f4a7c15ull;
        uint64_t z = state;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        st[i] = z ^ (z >> 31);
    }
    auto rotl = [](uint64_t x, int k) { return (x << k) | (x >> (64 - k)); };
    double level_factor = 1.0 / std::log((double)M);
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t r = rotl(st[0] + st[3], 23) + st[0];
        uint64_t t = st[1] << 17;
        st[2] ^= st[0]; st[3] ^= st[1]; st[1] ^= st[2]; st[0] ^= st[3];
        st[2] ^= t;
        st[3] = rotl(st[3], 45);
        double u = (double)(r >> 12) * (1.0 / 4503599627370496.0);
        double lv = std::round(-std::log(u) * level_factor);
        if (!(lv < (double)(HS_MAX_LAYERS - 1))) lv = (double)(HS_MAX_LAYERS - 1);
        level[i] = (uint8_t)lv;
    }
}

}  // extern "C"

// Batch-synchronous insertion of order[0 .. n) into the segment's graph (which may already hold other nodes):
// the loop of build.rs:123-166 as search / select / sort / reverse-link kernels per batch (hnsw_build.cuh).
// entry_after_first (node, layer), if given, becomes the entry point once the first batch has been inserted.
static int run_insertions(nidx_vec_segment* s, const std::vector<uint8_t>& level, const std::vector<uint32_t>& order, const std::vector<uint32_t>& ends,
                          cudaStream_t stream, const uint32_t* entry_after_first = nullptr) {
    uint64_t n = order.size();
    int r = 0;
    // work items (node position, layer), insertion order, layer ascending
    std::vector<uint64_t> wstart(n + 1);
    uint64_t W = 0;
    for (uint64_t i = 0; i < n; ++i) { wstart[i] = W; W += (uint64_t)level[order[i]] + 1; }
    wstart[n] = W;
    std::vector<uint32_t> w_pos(W);
    std::vector<unsigned char> w_layer(W);
    for (uint64_t i = 0; i < n; ++i)
        for (int l = 0; l <= level[order[i]]; ++l) { w_pos[wstart[i] + l] = (uint32_t)i; w_layer[wstart[i] + l] = (unsigned char)l; }
    uint64_t max_b = 0, max_w = 0;
    for (size_t b = 0, begin = 0; b < ends.size(); begin = ends[b], ++b) {
        max_b = std::max<uint64_t>(max_b, ends[b] - begin);
        max_w = std::max<uint64_t>(max_w, wstart[ends[b]] - wstart[begin]);
    }
    int efC = s->cfg.ef_construction, M = s->cfg.m;
    uint32_t *d_order = nullptr, *d_wpos = nullptr, *d_rev_x = nullptr, *d_idx = nullptr, *d_idx_sorted = nullptr, *d_heads = nullptr;
    unsigned int* d_head_ctr = nullptr;  // [0] number of heads, [1] work counter
    unsigned char* d_wlayer = nullptr;
    uint64_t *d_found = nullptr, *d_rev_key = nullptr, *d_key_sorted = nullptr;
    int* d_found_count = nullptr;
    float* d_rev_sim = nullptr;
    void* d_cub = nullptr;
    size_t cub_bytes = 0;
    size_t max_rev = (size_t)max_w * M;
    auto cleanup = [&]() {
        cudaFree(d_order); cudaFree(d_wpos); cudaFree(d_wlayer); cudaFree(d_found); cudaFree(d_found_count); cudaFree(d_rev_key); cudaFree(d_rev_x);
        cudaFree(d_rev_sim); cudaFree(d_key_sorted); cudaFree(d_idx); cudaFree(d_idx_sorted); cudaFree(d_cub); cudaFree(d_heads); cudaFree(d_head_ctr);
    };
    r = [&]() -> int {
        CU(cudaMalloc(&d_order, n * 4));
        CU(cudaMalloc(&d_wpos, W * 4));
        CU(cudaMalloc(&d_wlayer, W));
        CU(cudaMalloc(&d_found, (size_t)max_b * HS_MAX_LAYERS * efC * 8));
        CU(cudaMalloc(&d_found_count, (size_t)max_b * HS_MAX_LAYERS * 4));
        CU(cudaMalloc(&d_rev_key, max_rev * 8));
        CU(cudaMalloc(&d_key_sorted, max_rev * 8));
        CU(cudaMalloc(&d_rev_x, max_rev * 4));
        CU(cudaMalloc(&d_rev_sim, max_rev * 4));
        CU(cudaMalloc(&d_idx, max_rev * 4));
        CU(cudaMalloc(&d_idx_sorted, max_rev * 4));
        CU(cudaMalloc(&d_heads, max_rev * 4));
        CU(cudaMalloc(&d_head_ctr, 64));
        CU(cudaMemcpyAsync(d_order, order.data(), n * 4, cudaMemcpyHostToDevice, stream));
        CU(cudaMemcpyAsync(d_wpos, w_pos.data(), W * 4, cudaMemcpyHostToDevice, stream));
        CU(cudaMemcpyAsync(d_wlayer, w_layer.data(), W, cudaMemcpyHostToDevice, stream));
        {
            std::vector<uint32_t> iota(max_rev);
            for (size_t i = 0; i < max_rev; ++i) iota[i] = (uint32_t)i;
            CU(cudaMemcpyAsync(d_idx, iota.data(), max_rev * 4, cudaMemcpyHostToDevice, stream));
            CU(cudaStreamSynchronize(stream));
        }
        CU(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, d_rev_key, d_key_sorted, d_idx, d_idx_sorted, (int)max_rev, 0, 40, stream));
        CU(cudaMalloc(&d_cub, cub_bytes));

        // shared-memory plans

        int list_cap = efC, hash_bits;
        int slots = next_pow2(std::max(2048, (efC * s->s0 * 3) / 2));
        slots = std::max(slots, next_pow2(4 * list_cap));
        hash_bits = ilog2(slots);
        size_t smem_search = hs_smem_bytes(s->ld, list_cap, hash_bits);
        if (smem_search > 200 * 1024) return fail(NIDX_EINVAL, "HNSW build search needs %zu bytes of shared memory", smem_search);
        hs_kernel_t kern = pick_search_kernel(s->ld);
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_search));
        int occ = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, HS_THREADS, smem_search));
        size_t row_bytes = (size_t)s->ld * 4;
        size_t budget = 96 * 1024;
        int cache_sel = (int)std::min<size_t>(M, budget / row_bytes);
        int prune_max = std::max(s->cfg.m0, M) * 95 / 100;
        int cache_rev = (int)std::min<size_t>(prune_max, budget / row_bytes);
        // prune of a full list: stage all mmax + 1 vectors and the pairwise table when they fit (<= ~200 KB)
        int full_rows = std::max(s->cfg.m0, M) + 1;
        // NIDX_B200_PRUNE=table switches the prune to the staged pairwise-table variant (same result; measured
        // slightly slower than the candidate-at-a-time loop at d = 768, M0 = 32: 4.94 s vs 4.71 s per 1M vectors)
        const char* prune_env = getenv("NIDX_B200_PRUNE");
        bool preload = hb_smem_bytes(s->ld, full_rows, true) <= 200 * 1024 && prune_env && !strcmp(prune_env, "table");
        if (preload) cache_rev = full_rows;
        size_t smem_sel = hb_smem_bytes(s->ld, cache_sel), smem_rev = hb_smem_bytes(s->ld, cache_rev, preload);
        CU(cudaFuncSetAttribute(select_link_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sel));
        CU(cudaFuncSetAttribute(reverse_link_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rev));
        int occ_rev = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_rev, reverse_link_kernel, HB_THREADS, smem_rev));
        int rev_grid = std::max(1, occ_rev) * s->sm_count;
        CU(cudaMemsetAsync(s->d_counters, 0, 8 * sizeof(unsigned long long), stream));
        s->last_counters.store(nullptr);   // the getters report the build's counters until the next search


        VecDev V = s->vdev();
        GraphDev G = s->gdev();
        uint32_t begin = 0;
        for (size_t b = 0; b < ends.size(); ++b) {
            uint32_t end = ends[b];
            int nb = (int)(end - begin);
            int nw = (int)(wstart[end] - wstart[begin]);
            SearchArgs a;
            memset(&a, 0, sizeof(a));
            a.mode = 1; a.nq = nb; a.nodes = d_order + begin; a.efC = efC; a.found = d_found; a.found_count = d_found_count;
            a.hash_bits = hash_bits; a.list_cap = list_cap; a.cu_cap = 0; a.work_counter = s->d_work_counter; a.counters = s->d_counters;
            CU(cudaMemsetAsync(s->d_work_counter, 0, 4, stream));
            int grid = std::min(nb, std::max(1, occ) * s->sm_count);
            kern<<<grid, HS_THREADS, smem_search, stream>>>(V, G, a);
            LAUNCHED();
            BuildArgs ba;
            ba.n_work = nw; ba.w_pos = d_wpos + wstart[begin]; ba.w_layer = d_wlayer + wstart[begin]; ba.order = d_order; ba.batch_begin = begin;
            ba.efC = efC; ba.M = M; ba.found = d_found; ba.found_count = d_found_count; ba.rev_key = d_rev_key; ba.rev_x = d_rev_x; ba.rev_sim = d_rev_sim;
            ba.cache_cap = cache_sel;
            select_link_kernel<<<nw, HB_THREADS, smem_sel, stream>>>(V, G, ba);
            LAUNCHED();
            int n_rev = nw * M;
            size_t tmp = cub_bytes;
            CU(cub::DeviceRadixSort::SortPairs(d_cub, tmp, d_rev_key, d_key_sorted, d_idx, d_idx_sorted, n_rev, 0, 40, stream));
            LAUNCHED();
            ReverseArgs ra;
            ra.n_rev = n_rev; ra.key_sorted = d_key_sorted; ra.idx_sorted = d_idx_sorted; ra.rev_x = d_rev_x; ra.rev_sim = d_rev_sim; ra.cache_cap = cache_rev; ra.preload = preload ? 1 : 0;
            ra.heads = d_heads; ra.n_heads = d_head_ctr; ra.work_counter = d_head_ctr + 1;
            CU(cudaMemsetAsync(d_head_ctr, 0, 8, stream));
            collect_heads_kernel<<<(n_rev + 255) / 256, 256, 0, stream>>>(d_key_sorted, n_rev, d_heads, d_head_ctr);
            LAUNCHED();
            reverse_link_kernel<<<std::min(n_rev, rev_grid), HB_THREADS, smem_rev, stream>>>(V, G, ra);
            LAUNCHED();
            begin = end;
            if (b == 0 && entry_after_first) {   // kernel arguments travel by value: later batches start from the new entry point
                s->entry_node = entry_after_first[0];
                s->entry_layer = entry_after_first[1];
                G = s->gdev();
            }
        }
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(stream));
        return 0;
    }();
    cleanup();
    return r;
}

extern "C" {

int nidx_hnsw_levels(uint64_t n, int32_t m, uint64_t seed, uint8_t* out_level) {
    if (!out_level && n) return fail(NIDX_EINVAL, "null argument");
    if (m < 2) return fail(NIDX_EINVAL, "M must be at least 2");
    host_assign_levels(n, m, seed, out_level);
    return 0;
}

int nidx_vec_build_hnsw(nidx_vec_segment* s, uint64_t seed, int32_t max_batch, void* stream_) {
    if (!s) return fail(NIDX_EINVAL, "null segment");
    CU(cudaSetDevice(s->cfg.device));
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    uint64_t n = s->n;
    if (max_batch <= 0) max_batch = 4096;
    std::vector<uint8_t> level(n ? n : 1);
    host_assign_levels(n, s->cfg.m, seed, level.data());
    int r = alloc_graph(s, level.data());
    if (r) return r;
    s->has_graph = true;
    if (n == 0) return 0;

    // insertion order: entry point first, then ascending id; batch b = min(max_batch, max(1, done/16))
    std::vector<uint32_t> order(n);
    order[0] = s->entry_node;
    for (uint64_t i = 0, j = 1; i < n; ++i) if (i != s->entry_node) order[j++] = (uint32_t)i;
    std::vector<uint32_t> ends;
    for (uint64_t done = 0; done < n;) {
        uint64_t b = std::min<uint64_t>({(uint64_t)max_batch, std::max<uint64_t>(1, done / 16), n - done});
        done += b;
        ends.push_back((uint32_t)done);
    }
    r = run_insertions(s, level, order, ends, stream);
    if (r) { free_graph(s); return r; }
    return 0;
}

// merge_indexes' fast path (segment.rs:143-167): the first n_existing vectors of this (merged) segment are the largest
// input segment, which had no deletions, so its graph is reused and only the remaining vectors are inserted.
// HnswBuilder::new seeds a fresh level RNG and initialize_graph(skip_nodes = n_existing, total) draws the new nodes'
// levels (build.rs:36-55); the entry point moves only if a higher layer appears (ram_hnsw.rs:99-107).
int nidx_vec_extend_hnsw(nidx_vec_segment* s, uint64_t n_existing, const uint8_t* level_existing, const uint32_t* adj0, const float* w0,
                         const uint32_t* adjU, const float* wU, uint32_t entry_node, uint32_t entry_layer, uint64_t seed, int32_t max_batch, void* stream_) {
    if (!s || !level_existing || !adj0) return fail(NIDX_EINVAL, "null argument");
    if (!w0 || (adjU && !wU)) return fail(NIDX_EINVAL, "the existing edges' similarities are required (hnsw.edges): the reverse-link prune ranks by them");
    if (n_existing == 0 || n_existing > s->n) return fail(NIDX_EINVAL, "n_existing must be in 1..len");
    if (entry_node >= n_existing || level_existing[entry_node] < entry_layer) return fail(NIDX_EINVAL, "bad entry point");
    CU(cudaSetDevice(s->cfg.device));
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    uint64_t n = s->n;
    if (max_batch <= 0) max_batch = 4096;
    std::vector<uint8_t> level(n);
    memcpy(level.data(), level_existing, n_existing);
    for (uint64_t i = 0; i < n_existing; ++i)
        if (level[i] >= HS_MAX_LAYERS) return fail(NIDX_EINVAL, "node %llu has level %d >= %d", (unsigned long long)i, level[i], HS_MAX_LAYERS);
    uint64_t rows_existing = 0;
    for (uint64_t i = 0; i < n_existing; ++i) rows_existing += level[i];
    if (rows_existing && !adjU) return fail(NIDX_EINVAL, "existing nodes have upper layers but adjU is null");
    host_assign_levels(n - n_existing, s->cfg.m, seed, level.data() + n_existing);
    int r = alloc_graph(s, level.data());
    if (r) return r;
    // alloc_graph left the entry point on the lowest id of the global top layer.  If that is a NEW node (the merge raises the
    // top layer) the reference moves the entry point there before the node has a single link (update_entry_point in
    // initialize_graph, build.rs:49-55), so every later search starts on an island and the reused graph becomes unreachable.
    // Deliberate deviation: that node is inserted first, from the old entry point, and the entry point moves afterwards.
    uint32_t raised[2] = {s->entry_node, s->entry_layer};
    bool raises = raised[1] > entry_layer && raised[0] >= n_existing;
    if (raised[1] > entry_layer && !raises) { entry_node = raised[0]; entry_layer = raised[1]; }   // the caller's entry point was below its own top layer
    s->entry_layer = entry_layer;
    s->entry_node = entry_node;
    // the existing nodes' rows: layer 0 rows are a prefix, and so are their upper-pool rows (offsets depend on earlier nodes only)
    CU(cudaMemcpy(s->d_adj0, adj0, (size_t)n_existing * s->s0 * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(s->d_w0, w0, (size_t)n_existing * s->s0 * 4, cudaMemcpyHostToDevice));
    if (rows_existing) {
        // fix_broken_graph (ram_hnsw.rs:52-64,118-123): a link in layer L > 0 to a node that is not in layer L is dropped
        // (graphs written by an old version can hold them); the patched copy is made only if one is found
        std::vector<uint32_t> fixed_adj;
        std::vector<float> fixed_w;
        uint64_t row = 0;
        for (uint64_t i = 0; i < n_existing; ++i)
            for (int layer = 1; layer <= level[i]; ++layer, ++row) {
                const uint32_t* r0 = adjU + row * s->su;
                bool broken = false;
                for (int j = 0; j < s->su && r0[j] != NIL; ++j)
                    if (r0[j] >= n_existing || level[r0[j]] < layer) { broken = true; break; }
                if (!broken) continue;
                if (fixed_adj.empty()) {
                    fixed_adj.assign(adjU, adjU + rows_existing * s->su);
                    if (wU) fixed_w.assign(wU, wU + rows_existing * s->su);
                }
                uint32_t* dst = fixed_adj.data() + row * s->su;
                float* dw = wU ? fixed_w.data() + row * s->su : nullptr;
                int kept = 0;
                for (int j = 0; j < s->su && r0[j] != NIL; ++j)
                    if (r0[j] < n_existing && level[r0[j]] >= layer) {
                        dst[kept] = r0[j];
                        if (dw) dw[kept] = wU[row * s->su + j];
                        ++kept;
                    }
                for (int j = kept; j < s->su; ++j) { dst[j] = NIL; if (dw) dw[j] = 0.0f; }
            }
        const uint32_t* srcA = fixed_adj.empty() ? adjU : fixed_adj.data();
        const float* srcW = fixed_adj.empty() ? wU : fixed_w.data();
        CU(cudaMemcpy(s->d_adjU, srcA, (size_t)rows_existing * s->su * 4, cudaMemcpyHostToDevice));
        if (wU) CU(cudaMemcpy(s->d_wU, srcW, (size_t)rows_existing * s->su * 4, cudaMemcpyHostToDevice));
    }
    s->has_graph = true;
    if (n == n_existing) return 0;
    std::vector<uint32_t> order;
    order.reserve(n - n_existing);
    if (raises) order.push_back(raised[0]);
    for (uint64_t i = n_existing; i < n; ++i)
        if (!raises || i != raised[0]) order.push_back((uint32_t)i);
    std::vector<uint32_t> ends;
    if (raises) ends.push_back(1);
    for (uint64_t done = n_existing + (raises ? 1 : 0); done < n;) {
        uint64_t b = std::min<uint64_t>({(uint64_t)max_batch, std::max<uint64_t>(1, done / 16), n - done});
        done += b;
        ends.push_back((uint32_t)(done - n_existing));
    }
    r = run_insertions(s, level, order, ends, stream, raises ? raised : nullptr);
    if (r) { free_graph(s); return r; }
    return 0;
}

// ---- segment files ----------------------------------------------------------------------------
int nidx_vec_open(const nidx_vec_config* cfg, const char* dir, nidx_vec_segment** out) {
    if (!dir) return fail(NIDX_EINVAL, "null dir");
    nidx_vec_segment* s = nullptr;
    int r = new_segment(cfg, &s);
    if (r) return r;
    std::string err;
    std::vector<unsigned char> raw;
    if (!segio::read_file(std::string(dir) + "/vectors.bin", raw, err)) { delete s; return fail(NIDX_EIO, "%s", err.c_str()); }
    size_t rec = (size_t)s->d * 4 + 4;  // vector_store.rs:33-68: [dim x f32 LE][paragraph_addr u32]
    if (raw.size() % rec) { delete s; return fail(NIDX_EIO, "vectors.bin size %zu is not a multiple of the record length %zu", raw.size(), rec); }
    uint64_t n = raw.size() / rec;
    s->n = n;
    std::vector<uint32_t> par(n);
    for (uint64_t i = 0; i < n; ++i) memcpy(&par[i], raw.data() + i * rec + (size_t)s->d * 4, 4);
    r = [&]() -> int {
        CU(cudaMalloc(&s->d_vecs, std::max<size_t>((size_t)n * s->ld * 4, 16)));
        if (n) {
            void* staged = nullptr;
            CU(cudaMalloc(&staged, raw.size()));
            CU(cudaMemcpy(staged, raw.data(), raw.size(), cudaMemcpyHostToDevice));
            pad_rows_kernel<<<s->sm_count * 8, 256>>>(reinterpret_cast<unsigned char*>(staged), rec, s->d, s->d_vecs, s->ld, n);
            LAUNCHED();
            CU(cudaGetLastError());
            CU(cudaDeviceSynchronize());
            cudaFree(staged);
        }
        return 0;
    }();
    if (!r) r = finish_create(s, n ? par.data() : nullptr);
    if (r) { nidx_vec_close(s); return r; }
    // hnsw.graph (+ hnsw.edges) if present
    std::vector<unsigned char> graph, edges;
    if (segio::read_file(std::string(dir) + "/hnsw.graph", graph, err) && !graph.empty()) {
        segio::read_file(std::string(dir) + "/hnsw.edges", edges, err);
        segio::FlatGraph fg;
        if (!segio::parse_graph_v2(graph, edges, n, stride0_for(s->cfg.m0), strideU_for(s->cfg.m), HS_MAX_LAYERS, fg, err)) {
            nidx_vec_close(s);
            return fail(NIDX_EIO, "hnsw.graph: %s", err.c_str());
        }
        r = nidx_vec_set_graph(s, fg.level.data(), fg.adj0.data(), fg.w0.empty() ? nullptr : fg.w0.data(), fg.adjU.data(), fg.wU.empty() ? nullptr : fg.wU.data());
        if (r) { nidx_vec_close(s); return r; }
        s->entry_node = fg.entry_node;   // the file's entry point (ram_hnsw.rs: hash-order dependent in the reference)
        s->entry_layer = fg.entry_layer;
    }
    // vectors.quant (data_store/v2/quant_vector_store.rs:29-62): the RaBitQ codes, one record of dim / 8 + 8 bytes per vector, loaded as they
    // are (re-laid out to the 16-byte stride in HBM) instead of re-encoding
    std::vector<unsigned char> quant;
    if (s->cfg.similarity == NIDX_SIM_DOT && s->d % 64 == 0 && s->d / 32 <= RQ_MAX_WORDS32 && segio::read_file(std::string(dir) + "/vectors.quant", quant, err) && !quant.empty()) {
        size_t rec_q = (size_t)s->d / 8 + 8;
        if (quant.size() != rec_q * n) { nidx_vec_close(s); return fail(NIDX_EIO, "vectors.quant holds %zu bytes, expected %zu records of %zu", quant.size(), (size_t)n, rec_q); }
        s->quant_stride = rabitq_stride(s->d);
        r = [&]() -> int {
            CU(cudaMalloc(&s->d_quant, std::max<size_t>((size_t)n * s->quant_stride, 16)));
            CU(cudaMemset(s->d_quant, 0, std::max<size_t>((size_t)n * s->quant_stride, 16)));
            CU(cudaMemcpy2D(s->d_quant, (size_t)s->quant_stride, quant.data(), rec_q, rec_q, (size_t)n, cudaMemcpyHostToDevice));
            return 0;
        }();
        if (r) { nidx_vec_close(s); return r; }
    }
    *out = s;
    return 0;
}

int nidx_vec_save(nidx_vec_segment* s, const char* dir) {
    if (!s || !dir) return fail(NIDX_EINVAL, "null argument");
    CU(cudaSetDevice(s->cfg.device));
    CU(cudaDeviceSynchronize());
    uint64_t n = s->n;
    std::vector<float> vecs((size_t)n * s->ld);
    CU(cudaMemcpy(vecs.data(), s->d_vecs, vecs.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> par(n);
    if (s->d_par_of) CU(cudaMemcpy(par.data(), s->d_par_of, n * 4, cudaMemcpyDeviceToHost));
    else for (uint64_t i = 0; i < n; ++i) par[i] = (uint32_t)i;
    std::string err;
    if (!segio::write_vectors_bin(std::string(dir) + "/vectors.bin", vecs.data(), n, s->d, s->ld, par.data(), err)) return fail(NIDX_EIO, "%s", err.c_str());
    if (s->has_graph) {
        segio::FlatGraph fg;
        fg.n = n; fg.s0 = s->s0; fg.su = s->su; fg.entry_node = s->entry_node; fg.entry_layer = s->entry_layer;
        fg.level = s->h_level;
        fg.upper_rows = s->upper_rows;
        fg.adj0.resize((size_t)n * s->s0); fg.w0.resize((size_t)n * s->s0);
        fg.adjU.resize((size_t)s->upper_rows * s->su); fg.wU.resize((size_t)s->upper_rows * s->su);
        int r = nidx_vec_get_graph(s, nullptr, fg.adj0.data(), fg.w0.data(), fg.adjU.data(), fg.wU.data());
        if (r) return r;
        if (!segio::write_graph_v2(std::string(dir) + "/hnsw.graph", std::string(dir) + "/hnsw.edges", fg, err)) return fail(NIDX_EIO, "%s", err.c_str());
    }
    if (s->d_quant) {   // vectors.quant: QuantVectorStoreWriter (quant_vector_store.rs), records back to back
        size_t rec_q = (size_t)s->d / 8 + 8;
        std::vector<unsigned char> quant(rec_q * n);
        if (n) CU(cudaMemcpy2D(quant.data(), rec_q, s->d_quant, (size_t)s->quant_stride, rec_q, (size_t)n, cudaMemcpyDeviceToHost));
        FILE* f = fopen((std::string(dir) + "/vectors.quant").c_str(), "wb");
        if (!f) return fail(NIDX_EIO, "cannot write %s/vectors.quant", dir);
        size_t wr = quant.empty() ? 0 : fwrite(quant.data(), 1, quant.size(), f);
        fclose(f);
        if (wr != quant.size()) return fail(NIDX_EIO, "short write to %s/vectors.quant", dir);
    }
    return 0;
}

// ---- text ---------------------------------------------------------------------------------------
struct nidx_txt_segment {
    int device = 0, sm_count = 0;
    uint32_t n_docs = 0, n_terms = 0;
    uint64_t n_post = 0;
    uint64_t* d_term_off = nullptr;
    uint2* d_post = nullptr;          // (doc, tf << 8 | fieldnorm id)
    uint32_t* d_skip_row = nullptr;   // [n_terms]
    uint32_t* d_skip = nullptr;       // [rows][n_fine + 1]
    uint32_t n_fine = 0;
    uint64_t* d_alive = nullptr;
    float* d_weight = nullptr;   // [n_terms]
    float* d_norm_cache = nullptr;  // [256]
    unsigned int* d_error = nullptr;
    uint64_t* d_doc_keys = nullptr;   // [n_docs] caller keys of the documents (paragraph ids) for rank fusion (nidx_txt_set_doc_keys)
    std::vector<uint64_t> own_df;
    uint64_t own_tokens = 0;
    float max_weight = 0.0f;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;  // around bm25_kernel of the last search (bench roofline)
    WorkspacePool pool;
};

// tantivy fieldnorm code -> token count (Lucene SmallFloat.byte4ToInt) [recalled]
static uint32_t fieldnorm_id_to_value(uint32_t id) {
    if (id < 24) return id;
    uint32_t j = id - 24, bits = j & 7, shift = j >> 3;
    return 24 + (shift == 0 ? bits : ((bits | 8u) << (shift - 1)));
}

static int txt_upload_stats(nidx_txt_segment* t, uint64_t total_docs, uint64_t total_tokens, const uint64_t* df) {
    const float K1 = 1.2f, B = 0.75f;
    float avg = (float)total_tokens / (float)total_docs;
    float cache[256];
    for (int i = 0; i < 256; ++i) cache[i] = K1 * (1.0f - B + B * (float)fieldnorm_id_to_value(i) / avg);
    std::vector<float> weight(t->n_terms);
    float wmax = 0.0f;
    for (uint32_t i = 0; i < t->n_terms; ++i) {
        float x = ((float)(total_docs - df[i]) + 0.5f) / ((float)df[i] + 0.5f);
        weight[i] = logf(1.0f + x) * (1.0f + K1);
        wmax = std::max(wmax, weight[i]);
    }
    t->max_weight = wmax;
    CU(cudaMemcpy(t->d_norm_cache, cache, sizeof(cache), cudaMemcpyHostToDevice));
    if (t->n_terms) CU(cudaMemcpy(t->d_weight, weight.data(), (size_t)t->n_terms * 4, cudaMemcpyHostToDevice));
    return 0;
}

int nidx_txt_create(int32_t device, uint32_t n_docs, uint32_t n_terms, const uint64_t* term_off, const uint32_t* post_doc, const uint32_t* post_tf,
                    const uint8_t* fieldnorm_id, nidx_txt_segment** out) {
    if (!term_off || !fieldnorm_id || !out) return fail(NIDX_EINVAL, "null argument");
    int r = check_device(device);
    if (r) return r;
    if (n_docs >= (1u << 31)) return fail(NIDX_EINVAL, "at most 2^31-1 documents per segment");
    nidx_txt_segment* t = new nidx_txt_segment();
    t->device = device; t->n_docs = n_docs; t->n_terms = n_terms; t->n_post = term_off[n_terms];
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    t->sm_count = prop.multiProcessorCount;
    r = [&]() -> int {
        t->n_fine = (n_docs + BM_FINE - 1) / BM_FINE;
        CU(cudaMalloc(&t->d_term_off, ((size_t)n_terms + 1) * 8));
        CU(cudaMalloc(&t->d_post, std::max<uint64_t>(t->n_post, 1) * 8));
        CU(cudaMalloc(&t->d_weight, std::max<uint32_t>(n_terms, 1) * 4));
        CU(cudaMalloc(&t->d_norm_cache, 1024));
        CU(cudaMalloc(&t->d_skip_row, std::max<uint32_t>(n_terms, 1) * 4));
        CU(cudaMalloc(&t->d_error, 4));
        CU(cudaMemset(t->d_error, 0, 4));
        CU(cudaEventCreate(&t->ev_k0));
        CU(cudaEventCreate(&t->ev_k1));
        CU(cudaMemcpy(t->d_term_off, term_off, ((size_t)n_terms + 1) * 8, cudaMemcpyHostToDevice));
        if (t->n_post) {
            // staged only to be packed into the 8-byte posting records
            uint32_t *d_doc = nullptr, *d_tf = nullptr;
            unsigned char* d_fn = nullptr;
            CU(cudaMalloc(&d_doc, t->n_post * 4));
            CU(cudaMalloc(&d_fn, std::max<uint32_t>(n_docs, 1)));
            CU(cudaMemcpy(d_doc, post_doc, t->n_post * 4, cudaMemcpyHostToDevice));
            CU(cudaMemcpy(d_fn, fieldnorm_id, n_docs, cudaMemcpyHostToDevice));
            if (post_tf) {
                CU(cudaMalloc(&d_tf, t->n_post * 4));
                CU(cudaMemcpy(d_tf, post_tf, t->n_post * 4, cudaMemcpyHostToDevice));
            }
            bm25_pack_kernel<<<t->sm_count * 8, 256>>>(d_doc, d_tf, d_fn, t->n_post, t->d_post);
            LAUNCHED();
            CU(cudaGetLastError());
            CU(cudaDeviceSynchronize());
            cudaFree(d_doc); cudaFree(d_tf); cudaFree(d_fn);
        }
        // skip rows for the terms with enough postings
        std::vector<uint32_t> skip_row(n_terms, NIDX_NIL), row_term;
        for (uint32_t i = 0; i < n_terms; ++i)
            if (term_off[i + 1] - term_off[i] >= (uint64_t)BM_SKIP_DF) { skip_row[i] = (uint32_t)row_term.size(); row_term.push_back(i); }
        if (n_terms) CU(cudaMemcpy(t->d_skip_row, skip_row.data(), (size_t)n_terms * 4, cudaMemcpyHostToDevice));
        size_t skip_words = std::max<size_t>(row_term.size(), 1) * ((size_t)t->n_fine + 1);
        CU(cudaMalloc(&t->d_skip, skip_words * 4));
        if (!row_term.empty()) {
            uint32_t* d_row_term = nullptr;
            CU(cudaMalloc(&d_row_term, row_term.size() * 4));
            CU(cudaMemcpy(d_row_term, row_term.data(), row_term.size() * 4, cudaMemcpyHostToDevice));
            bm25_build_skip_kernel<<<t->sm_count * 8, 256>>>(t->d_term_off, t->d_post, d_row_term, (uint32_t)row_term.size(), t->n_fine, t->d_skip);
            LAUNCHED();
            CU(cudaGetLastError());
            CU(cudaDeviceSynchronize());
            cudaFree(d_row_term);
        }
        t->own_df.resize(n_terms);
        for (uint32_t i = 0; i < n_terms; ++i) t->own_df[i] = term_off[i + 1] - term_off[i];
        // a segment alone only knows the quantised lengths; the exact token total comes with set_stats
        uint64_t tokens = 0;
        for (uint32_t i = 0; i < n_docs; ++i) tokens += fieldnorm_id_to_value(fieldnorm_id[i]);
        t->own_tokens = tokens;
        return txt_upload_stats(t, std::max<uint32_t>(n_docs, 1), std::max<uint64_t>(tokens, 1), t->own_df.data());
    }();
    if (r) { nidx_txt_close(t); return r; }
    *out = t;
    return 0;
}

int nidx_txt_set_stats(nidx_txt_segment* t, uint64_t total_docs, uint64_t total_tokens, const uint64_t* doc_freq) {
    if (!t || !total_docs) return fail(NIDX_EINVAL, "bad argument");
    CU(cudaSetDevice(t->device));
    return txt_upload_stats(t, total_docs, total_tokens, doc_freq ? doc_freq : t->own_df.data());
}

int nidx_txt_set_alive(nidx_txt_segment* t, const uint64_t* alive_bits) {
    if (!t) return fail(NIDX_EINVAL, "null segment");
    CU(cudaSetDevice(t->device));
    if (!alive_bits) { cudaFree(t->d_alive); t->d_alive = nullptr; return 0; }
    size_t words = ((size_t)t->n_docs + 63) / 64;
    if (!t->d_alive) CU(cudaMalloc(&t->d_alive, std::max<size_t>(words, 1) * 8));
    CU(cudaMemcpy(t->d_alive, alive_bits, words * 8, cudaMemcpyHostToDevice));
    return 0;
}

void nidx_txt_close(nidx_txt_segment* t) {
    if (!t) return;
    cudaSetDevice(t->device);
    cudaDeviceSynchronize();
    cudaFree(t->d_term_off); cudaFree(t->d_post); cudaFree(t->d_skip_row); cudaFree(t->d_skip); cudaFree(t->d_alive); cudaFree(t->d_weight);
    cudaFree(t->d_norm_cache); cudaFree(t->d_error); cudaFree(t->d_doc_keys);
    if (t->ev_k0) cudaEventDestroy(t->ev_k0);
    if (t->ev_k1) cudaEventDestroy(t->ev_k1);
    delete t;
}

int nidx_txt_last_kernel_ms(nidx_txt_segment* t, float* ms) {
    if (!t || !ms) return fail(NIDX_EINVAL, "null argument");
    CU(cudaSetDevice(t->device));
    CU(cudaEventSynchronize(t->ev_k1));
    CU(cudaEventElapsedTime(ms, t->ev_k0, t->ev_k1));
    return 0;
}

}  // extern "C"

typedef void (*bm_kernel_t)(TxtDev, Bm25Args);

// The body of nidx_txt_search.  qhost / ohost as in vec_search_impl.
static int txt_search_impl(nidx_txt_segment* t, const uint32_t* query_terms, const uint32_t* query_off, int32_t nq, bool qhost, bool ohost,
                           const nidx_txt_search_params* p, uint32_t* out_docs, float* out_scores, int32_t* out_counts, uint64_t* out_total, cudaStream_t stream) {
    if (!t || !p || !query_off || !out_docs || !out_scores || !out_counts) return fail(NIDX_EINVAL, "null argument");
    if (nq <= 0) return 0;
    int k = p->k;
    if (k <= 0 || k > 1024) return fail(NIDX_EINVAL, "k must be in 1..1024");
    CU(cudaSetDevice(t->device));
    WsGuard g(t->pool, stream);
    Workspace& w = *g.w;
    // query offsets are needed on the host to size things
    std::vector<uint32_t> h_off(nq + 1);
    if (qhost) memcpy(h_off.data(), query_off, ((size_t)nq + 1) * 4);
    else { CU(cudaMemcpyAsync(h_off.data(), query_off, ((size_t)nq + 1) * 4, cudaMemcpyDeviceToHost, stream)); CU(cudaStreamSynchronize(stream)); }
    uint32_t n_qt = h_off[nq];
    int max_terms = 0;
    for (int i = 0; i < nq; ++i) max_terms = std::max<int>(max_terms, h_off[i + 1] - h_off[i]);
    if (max_terms > BM_MAX_TERMS) return fail(NIDX_EINVAL, "queries with more than %d terms are not supported", BM_MAX_TERMS);
    const uint32_t *d_qt = query_terms, *d_qo = query_off;
    if (qhost) {
        ENSURE(w.queries, ((size_t)n_qt + nq + 1) * 4 + 16);
        uint32_t* base = w.queries.as<uint32_t>();
        CU(cudaMemcpyAsync(base, query_off, ((size_t)nq + 1) * 4, cudaMemcpyHostToDevice, stream));
        if (n_qt) CU(cudaMemcpyAsync(base + nq + 1, query_terms, (size_t)n_qt * 4, cudaMemcpyHostToDevice, stream));
        d_qo = base; d_qt = base + nq + 1;
    }
    // fixed-point scale: the largest possible sum is max_terms * max term weight (tf factor < 1)
    float bound = std::max(1.0f, t->max_weight * (float)std::max(max_terms, 1));
    int shift = 24;
    while (shift > 4 && bound * (float)(1u << shift) >= 4.0e9f) --shift;

    bool conj = p->mode == NIDX_BM25_AND;
    int cap = topk_cap(k, BM_THREADS);
    size_t smem = bm_smem_bytes(cap, conj);
    if (smem > 220 * 1024) return fail(NIDX_EINVAL, "BM25 needs %zu bytes of shared memory (k=%d): too large", smem, k);
    ENSURE(w.partial, (size_t)nq * k * 8);
    ENSURE(w.misc, (size_t)nq * 8);
    uint32_t* d_docs = out_docs; float* d_sc = out_scores; int* d_cnt = out_counts;
    unsigned long long* d_total = reinterpret_cast<unsigned long long*>(out_total);
    if (ohost) {
        ENSURE(w.out_ids, (size_t)nq * k * 4);
        ENSURE(w.out_scores, (size_t)nq * k * 4);
        ENSURE(w.out_counts, (size_t)nq * 4);
        d_docs = w.out_ids.as<uint32_t>(); d_sc = w.out_scores.as<float>(); d_cnt = w.out_counts.as<int>();
        d_total = w.misc.as<unsigned long long>();
    }
    TxtDev T;
    T.n_docs = t->n_docs; T.n_terms = t->n_terms; T.n_fine = t->n_fine; T.term_off = t->d_term_off; T.post = t->d_post;
    T.skip_row = t->d_skip_row; T.skip = t->d_skip; T.alive = t->d_alive;
    Bm25Args a;
    a.query_terms = d_qt; a.query_off = d_qo; a.nq = nq; a.k = k; a.cap = cap;
    a.term_weight = t->d_weight; a.norm_cache = t->d_norm_cache; a.shift = shift;
    a.after_mode = p->after_mode; a.after_score = p->after_score; a.after_docaddr = p->after_docaddr; a.docaddr_base = p->docaddr_base;
    a.out_keys = w.partial.as<uint64_t>(); a.out_total = d_total; a.error_flag = t->d_error;
    bm_kernel_t kern = conj ? (p->use_tf ? bm25_kernel<true, true> : bm25_kernel<true, false>) : (p->use_tf ? bm25_kernel<false, true> : bm25_kernel<false, false>);
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU(cudaEventRecord(t->ev_k0, stream));
    kern<<<nq, BM_THREADS, smem, stream>>>(T, a);
    CU(cudaEventRecord(t->ev_k1, stream));
    LAUNCHED();
    bm25_finish_kernel<<<nq, 128, 0, stream>>>(w.partial.as<uint64_t>(), nq, k, p->min_score, d_docs, d_sc, d_cnt);
    LAUNCHED();
    CU(cudaGetLastError());
    if (ohost) {
        CU(cudaMemcpyAsync(out_docs, d_docs, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(out_scores, d_sc, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(out_counts, d_cnt, (size_t)nq * 4, cudaMemcpyDeviceToHost, stream));
        if (out_total) CU(cudaMemcpyAsync(out_total, d_total, (size_t)nq * 8, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        unsigned int err = 0;   // an accumulator table that filled up would mean incomplete sums: report, never return them silently
        CU(cudaMemcpy(&err, t->d_error, 4, cudaMemcpyDeviceToHost));
        if (err) { cudaMemset(t->d_error, 0, 4); return fail(NIDX_EOVERFLOW, "BM25 accumulator table overflow"); }
    }
    return 0;
}

extern "C" {

int nidx_txt_search(nidx_txt_segment* t, const uint32_t* query_terms, const uint32_t* query_off, int32_t nq, int mem, const nidx_txt_search_params* p,
                    uint32_t* out_docs, float* out_scores, int32_t* out_counts, uint64_t* out_total, void* stream_) {
    bool host = mem == NIDX_MEM_HOST;
    return txt_search_impl(t, query_terms, query_off, nq, host, host, p, out_docs, out_scores, out_counts, out_total, reinterpret_cast<cudaStream_t>(stream_));
}

// ---- sharded search (shard.cuh) -------------------------------------------------------------------
struct nidx_shard_comm {
    int rank = 0, world = 1, device = 0;
    nccl_comm_t comm = nullptr;
    std::mutex mu;          // collectives on one communicator must be issued in the same order by every rank: one call at a time
    DevBuf local, gathered, total;
};

#define NC(expr)                                                                                                   \
    do {                                                                                                           \
        int e__ = (expr);                                                                                          \
        if (e__ != NCCL_SUCCESS) return fail(NIDX_ECUDA, "%s failed: %s", #expr, nccl_api().GetErrorString(e__)); \
    } while (0)

int nidx_shard_unique_id(uint8_t out[128]) {
    if (!out) return fail(NIDX_EINVAL, "null argument");
    NcclApi& N = nccl_api();
    if (!N.ok) return fail(NIDX_ESTATE, "NCCL (libnccl.so.2) is not available in this process");
    nccl_unique_id id;
    NC(N.GetUniqueId(&id));
    memcpy(out, id.internal, 128);
    return 0;
}

int nidx_shard_init(const uint8_t unique_id[128], int32_t rank, int32_t world, int32_t device, nidx_shard_comm** out) {
    if (!unique_id || !out || world <= 0 || rank < 0 || rank >= world) return fail(NIDX_EINVAL, "bad argument");
    int r = check_device(device);
    if (r) return r;
    NcclApi& N = nccl_api();
    if (!N.ok) return fail(NIDX_ESTATE, "NCCL (libnccl.so.2) is not available in this process");
    nidx_shard_comm* c = new nidx_shard_comm();
    c->rank = rank; c->world = world; c->device = device;
    nccl_unique_id id;
    memcpy(id.internal, unique_id, 128);
    int e = N.CommInitRank(&c->comm, world, id, rank);
    if (e != NCCL_SUCCESS) { delete c; return fail(NIDX_ECUDA, "ncclCommInitRank failed: %s", N.GetErrorString(e)); }
    *out = c;
    return 0;
}

void nidx_shard_destroy(nidx_shard_comm* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->comm) nccl_api().CommDestroy(c->comm);
    c->local.release(); c->gathered.release(); c->total.release();
    delete c;
}

int nidx_vec_set_paragraph_keys(nidx_vec_segment* s, const uint64_t* keys) {
    if (!s) return fail(NIDX_EINVAL, "null segment");
    CU(cudaSetDevice(s->cfg.device));
    if (!keys) { cudaFree(s->d_par_keys); s->d_par_keys = nullptr; return 0; }
    if (!s->d_par_keys) CU(cudaMalloc(&s->d_par_keys, std::max<size_t>(s->n_par, 1) * 8));
    CU(cudaMemcpy(s->d_par_keys, keys, (size_t)s->n_par * 8, cudaMemcpyHostToDevice));
    return 0;
}

// common tail: all-gather the local part, merge, deliver
static int shard_exchange_and_merge(nidx_shard_comm* c, int nq, int k, bool dedup, int with_duplicates, bool ohost, uint32_t* out_ids, float* out_scores,
                                    int32_t* out_part, int32_t* out_counts, cudaStream_t stream) {
    NcclApi& N = nccl_api();
    size_t words = shard_part_words(nq, k, dedup);
    uint32_t* local = c->local.as<uint32_t>();
    uint32_t* gathered = c->gathered.as<uint32_t>();
    NC(N.AllGather(local, gathered, words * 4, NCCL_INT8, c->comm, stream));
    LAUNCHED();
    // outputs: device pointers, or staged behind the gathered parts when the caller's are host pointers
    uint32_t* d_ids = out_ids; float* d_sc = out_scores; int* d_part = out_part; int* d_cnt = out_counts;
    if (ohost) {
        uint32_t* stage = gathered + (size_t)c->world * words;
        d_ids = stage; d_sc = reinterpret_cast<float*>(stage + (size_t)nq * k); d_part = reinterpret_cast<int*>(stage + 2 * (size_t)nq * k);
        d_cnt = reinterpret_cast<int*>(stage + 3 * (size_t)nq * k);
    }
    if (dedup) {
        size_t per = (size_t)k * 16 + (size_t)c->world * k * 8;
        int threads = (int)std::max<size_t>(1, std::min<size_t>(64, (size_t)(96 * 1024) / per));
        if (per > 96 * 1024) return fail(NIDX_EINVAL, "k = %d is too large for the de-duplicating merge over %d parts", k, c->world);
        threads = std::min(threads, nq);
        size_t smem = per * threads;
        if (smem > 48 * 1024) CU(cudaFuncSetAttribute(shard_fssc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        shard_fssc_kernel<<<(nq + threads - 1) / threads, threads, smem, stream>>>(gathered, c->world, words, nq, k, with_duplicates, d_ids, d_sc, d_part, d_cnt);
        LAUNCHED();
    } else {
        if (k > 1024 || (long long)c->world * k >= (1ll << 31)) return fail(NIDX_EINVAL, "k above 1024 not supported");
        int cap = topk_cap(k, 256);
        if ((size_t)cap * 8 > 48 * 1024) CU(cudaFuncSetAttribute(parts_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap * 8));
        parts_merge_kernel<<<nq, 256, (size_t)cap * 8, stream>>>(gathered, reinterpret_cast<const float*>(gathered + (size_t)nq * k), c->world, words, nq, k, cap, d_ids,
                                                                  d_sc, d_part);
        LAUNCHED();
        if (d_cnt) {
            shard_count_kernel<<<(nq + 255) / 256, 256, 0, stream>>>(d_ids, nq, k, d_cnt);
            LAUNCHED();
        }
    }
    CU(cudaGetLastError());
    if (ohost) {
        CU(cudaMemcpyAsync(out_ids, d_ids, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(out_scores, d_sc, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, stream));
        if (out_part) CU(cudaMemcpyAsync(out_part, d_part, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, stream));
        if (out_counts) CU(cudaMemcpyAsync(out_counts, d_cnt, (size_t)nq * 4, cudaMemcpyDeviceToHost, stream));
    }
    return 0;
}

int nidx_vec_search_sharded(nidx_shard_comm* c, nidx_vec_segment* seg, const float* queries, int32_t nq, int32_t ldq, int mem, const nidx_vec_search_params* p,
                            int32_t dedup, uint32_t* out_ids, float* out_scores, int32_t* out_part, int32_t* out_counts, void* stream_) {
    if (!c || !seg || !p || !out_ids || !out_scores) return fail(NIDX_EINVAL, "null argument");
    if (nq <= 0) return 0;
    if (seg->cfg.device != c->device) return fail(NIDX_EINVAL, "segment on device %d, communicator on device %d", seg->cfg.device, c->device);
    int k = p->k;
    if (k <= 0) return fail(NIDX_EINVAL, "k must be positive");
    CU(cudaSetDevice(c->device));
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    bool host = mem == NIDX_MEM_HOST;
    std::lock_guard<std::mutex> lock(c->mu);
    size_t words = shard_part_words(nq, k, dedup != 0);
    ENSURE(c->local, words * 4 + (size_t)nq * 4);
    ENSURE(c->gathered, (size_t)c->world * words * 4 + (size_t)nq * k * 12 + (size_t)nq * 4 + 64);
    uint32_t* local = c->local.as<uint32_t>();
    int* local_cnt = reinterpret_cast<int*>(local + words);
    // 1. this rank's segment: queries from the caller's memory, results straight into the exchange record
    int r = vec_search_impl(seg, queries, nq, ldq, host, false, p, local, reinterpret_cast<float*>(local + (size_t)nq * k), local_cnt, stream);
    if (r) return r;
    if (dedup) {
        int n_res = nq * k;
        shard_keys_kernel<<<(n_res * 32 + 255) / 256, 256, 0, stream>>>(seg->vdev(), local, n_res, seg->d_par_keys, (uint32_t)c->rank, p->with_duplicates ? 0 : 1,
                                                                       reinterpret_cast<uint64_t*>(local + 2 * (size_t)nq * k),
                                                                       reinterpret_cast<uint64_t*>(local + 4 * (size_t)nq * k));
        LAUNCHED();
    }
    // 2. + 3. exchange and merge
    r = shard_exchange_and_merge(c, nq, k, dedup != 0, p->with_duplicates, host, out_ids, out_scores, out_part, out_counts, stream);
    if (r) return r;
    if (host) CU(cudaStreamSynchronize(stream));
    return 0;
}

int nidx_txt_search_sharded(nidx_shard_comm* c, nidx_txt_segment* seg, const uint32_t* query_terms, const uint32_t* query_off, int32_t nq, int mem,
                            const nidx_txt_search_params* p, uint32_t* out_docs, float* out_scores, int32_t* out_part, int32_t* out_counts, uint64_t* out_total,
                            void* stream_) {
    if (!c || !seg || !p || !out_docs || !out_scores) return fail(NIDX_EINVAL, "null argument");
    if (nq <= 0) return 0;
    if (seg->device != c->device) return fail(NIDX_EINVAL, "segment on device %d, communicator on device %d", seg->device, c->device);
    int k = p->k;
    CU(cudaSetDevice(c->device));
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    bool host = mem == NIDX_MEM_HOST;
    std::lock_guard<std::mutex> lock(c->mu);
    NcclApi& N = nccl_api();
    size_t words = shard_part_words(nq, k, false);
    ENSURE(c->local, words * 4 + (size_t)nq * 4);
    ENSURE(c->gathered, (size_t)c->world * words * 4 + (size_t)nq * k * 12 + (size_t)nq * 4 + 64);
    ENSURE(c->total, (size_t)nq * 8);
    uint32_t* local = c->local.as<uint32_t>();
    int* local_cnt = reinterpret_cast<int*>(local + words);
    uint64_t* d_total = host ? c->total.as<uint64_t>() : (out_total ? out_total : c->total.as<uint64_t>());
    // the min_score cut is applied to the merged list by the caller's convention (reader.rs:302-305 drops below min_score after top-k):
    // every part applies it locally, which commutes with the merge.
    int r = txt_search_impl(seg, query_terms, query_off, nq, host, false, p, local, reinterpret_cast<float*>(local + (size_t)nq * k), local_cnt, d_total, stream);
    if (r) return r;
    NC(N.AllReduce(d_total, d_total, (size_t)nq, NCCL_UINT64, NCCL_SUM, c->comm, stream));   // Count collector over all parts
    LAUNCHED();
    r = shard_exchange_and_merge(c, nq, k, false, 1, host, out_docs, out_scores, out_part, out_counts, stream);
    if (r) return r;
    if (host) {
        if (out_total) CU(cudaMemcpyAsync(out_total, d_total, (size_t)nq * 8, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
    }
    return 0;
}

// ---- rank fusion + the fused shard search --------------------------------------------------------------
int nidx_txt_set_doc_keys(nidx_txt_segment* t, const uint64_t* keys) {
    if (!t) return fail(NIDX_EINVAL, "null segment");
    CU(cudaSetDevice(t->device));
    if (!keys) { cudaFree(t->d_doc_keys); t->d_doc_keys = nullptr; return 0; }
    if (!t->d_doc_keys) CU(cudaMalloc(&t->d_doc_keys, std::max<size_t>(t->n_docs, 1) * 8));
    CU(cudaMemcpy(t->d_doc_keys, keys, (size_t)t->n_docs * 8, cudaMemcpyHostToDevice));
    return 0;
}

}  // extern "C"

static WorkspacePool* plan_pool(int device) {
    static std::mutex mu;
    static WorkspacePool* pools[64] = {nullptr};
    std::lock_guard<std::mutex> g(mu);
    if (device < 0 || device >= 64) return nullptr;
    if (!pools[device]) pools[device] = new WorkspacePool();
    return pools[device];
}

// sources already on the device; outputs on the device
static int rrf_launch(const RrfSourceDev* src, int n_sources, int nq, double k, uint64_t* out_keys, double* out_scores, uint32_t* out_refs, int32_t* out_counts,
                      cudaStream_t stream) {
    RrfArgs a;
    memset(&a, 0, sizeof(a));
    int cap = 0;
    for (int i = 0; i < n_sources; ++i) { a.src[i] = src[i]; cap += src[i].k; }
    a.n_sources = n_sources; a.nq = nq; a.cap = cap; a.k = k;
    a.out_keys = out_keys; a.out_scores = out_scores; a.out_refs = out_refs; a.out_counts = out_counts;
    size_t smem = rf_smem_bytes(cap);
    if (smem > 200 * 1024) return fail(NIDX_EINVAL, "rank fusion of %d items per query needs %zu bytes of shared memory: too many", cap, smem);
    CU(cudaFuncSetAttribute(rrf_fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rrf_fuse_kernel<<<nq, RF_THREADS, smem, stream>>>(a);
    LAUNCHED();
    CU(cudaGetLastError());
    return 0;
}

// side streams of the fused shard search: one set per host thread and device (the reference runs the index searches of one request on
// scoped threads, shard_search.rs:215-239; here they are streams forked from, and joined back into, the caller's stream)
struct PlanStreams {
    cudaStream_t s[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t fork = nullptr, join[3] = {nullptr, nullptr, nullptr};
    int device = -1;
    int init(int dev) {
        if (device == dev) return 0;
        if (device >= 0) return fail(NIDX_EINVAL, "a host thread uses nidx_shard_search on one device only (first %d, now %d)", device, dev);
        for (int i = 0; i < 3; ++i) {
            CU(cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming));
        }
        CU(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
        device = dev;
        return 0;
    }
};
static thread_local PlanStreams g_plan_streams;

extern "C" {

int nidx_rank_fusion_rrf(int32_t device, const nidx_rrf_source* sources, int32_t n_sources, int32_t nq, double k, int mem, uint64_t* out_keys,
                         double* out_scores, uint32_t* out_refs, int32_t* out_counts, void* stream_) {
    int r = check_device(device);
    if (r) return r;
    if (!sources || n_sources <= 0 || n_sources > RF_MAX_SOURCES) return fail(NIDX_EINVAL, "rank fusion takes 1..%d sources", RF_MAX_SOURCES);
    if (!out_keys || !out_scores || !out_refs || !out_counts) return fail(NIDX_EINVAL, "null output");
    if (nq <= 0) return 0;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    bool host = mem == NIDX_MEM_HOST;
    int cap = 0;
    for (int i = 0; i < n_sources; ++i) {
        if (sources[i].k <= 0 || !sources[i].keys || !sources[i].scores) return fail(NIDX_EINVAL, "rank fusion source %d: keys, scores and k > 0 are required", i);
        if (sources[i].k >= (1 << 24)) return fail(NIDX_EINVAL, "rank fusion source %d: k too large", i);
        cap += sources[i].k;
    }
    RrfSourceDev dev[RF_MAX_SOURCES];
    WorkspacePool* pool = plan_pool(device);
    if (!pool) return fail(NIDX_EINVAL, "device %d", device);
    WsGuard g(*pool, stream);
    Workspace& w = *g.w;
    uint64_t* d_keys = out_keys; double* d_sc = out_scores; uint32_t* d_refs = out_refs; int32_t* d_cnt = out_counts;
    if (host) {
        size_t in_bytes = 0;
        auto al16 = [](size_t b) { return (b + 15) / 16 * 16; };
        for (int i = 0; i < n_sources; ++i) in_bytes += al16((size_t)nq * sources[i].k * 8) + al16((size_t)nq * sources[i].k * 4) + al16((size_t)nq * 4);
        ENSURE(w.queries, in_bytes);
        ENSURE(w.scores, (size_t)nq * cap * 20 + (size_t)nq * 4 + 128);
        unsigned char* p = w.queries.as<unsigned char>();
        for (int i = 0; i < n_sources; ++i) {
            size_t nk = (size_t)nq * sources[i].k;
            uint64_t* dk = reinterpret_cast<uint64_t*>(p); p += al16(nk * 8);
            float* ds = reinterpret_cast<float*>(p); p += al16(nk * 4);
            int32_t* dc = reinterpret_cast<int32_t*>(p); p += al16((size_t)nq * 4);
            CU(cudaMemcpyAsync(dk, sources[i].keys, nk * 8, cudaMemcpyHostToDevice, stream));
            CU(cudaMemcpyAsync(ds, sources[i].scores, nk * 4, cudaMemcpyHostToDevice, stream));
            if (sources[i].counts) CU(cudaMemcpyAsync(dc, sources[i].counts, (size_t)nq * 4, cudaMemcpyHostToDevice, stream));
            dev[i] = RrfSourceDev{dk, ds, sources[i].counts ? dc : nullptr, sources[i].k, sources[i].weight};
        }
        unsigned char* o = w.scores.as<unsigned char>();
        d_keys = reinterpret_cast<uint64_t*>(o); o += (size_t)nq * cap * 8;
        d_sc = reinterpret_cast<double*>(o); o += (size_t)nq * cap * 8;
        d_refs = reinterpret_cast<uint32_t*>(o); o += al16((size_t)nq * cap * 4);
        d_cnt = reinterpret_cast<int32_t*>(o);
    } else {
        for (int i = 0; i < n_sources; ++i) dev[i] = RrfSourceDev{sources[i].keys, sources[i].scores, sources[i].counts, sources[i].k, sources[i].weight};
    }
    r = rrf_launch(dev, n_sources, nq, k, d_keys, d_sc, d_refs, d_cnt, stream);
    if (r) return r;
    if (host) {
        CU(cudaMemcpyAsync(out_keys, d_keys, (size_t)nq * cap * 8, cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(out_scores, d_sc, (size_t)nq * cap * 8, cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(out_refs, d_refs, (size_t)nq * cap * 4, cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(out_counts, d_cnt, (size_t)nq * 4, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
    }
    return 0;
}

int nidx_shard_search(const nidx_shard_search_request* rq, nidx_shard_search_response* rs, int mem, void* stream_) {
    if (!rq || !rs) return fail(NIDX_EINVAL, "null argument");
    const int nq = rq->nq;
    if (nq <= 0) return 0;
    if (!rq->vec && !rq->par && !rq->doc) return fail(NIDX_EINVAL, "shard search without any index request");
    int device = rq->vec ? rq->vec->cfg.device : (rq->par ? rq->par->device : rq->doc->device);
    if ((rq->vec && rq->vec->cfg.device != device) || (rq->par && rq->par->device != device) || (rq->doc && rq->doc->device != device))
        return fail(NIDX_EINVAL, "the indexes of one shard search must live on one device");
    if (rq->vec && (!rq->vec_params || !rs->vec_ids || !rs->vec_scores || !rs->vec_counts)) return fail(NIDX_EINVAL, "vector request: params and outputs are required");
    if (rq->par && (!rq->par_params || !rs->par_docs || !rs->par_scores || !rs->par_counts)) return fail(NIDX_EINVAL, "paragraph request: params and outputs are required");
    if (rq->doc && (!rq->doc_params || !rs->doc_docs || !rs->doc_scores || !rs->doc_counts)) return fail(NIDX_EINVAL, "document request: params and outputs are required");
    const bool fuse = rq->rrf_k > 0.0 && rq->vec && rq->par;
    if (fuse && (!rs->fused_keys || !rs->fused_scores || !rs->fused_refs || !rs->fused_counts)) return fail(NIDX_EINVAL, "rank fusion: outputs are required");
    int r = check_device(device);
    if (r) return r;
    CU(cudaSetDevice(device));
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    const bool host = mem == NIDX_MEM_HOST;
    PlanStreams& ps = g_plan_streams;
    r = ps.init(device);
    if (r) return r;
    const int kv = rq->vec ? rq->vec_params->k : 0, kp = rq->par ? rq->par_params->k : 0, kd = rq->doc ? rq->doc_params->k : 0;
    if ((rq->vec && kv <= 0) || (rq->par && kp <= 0) || (rq->doc && kd <= 0)) return fail(NIDX_EINVAL, "k must be positive");
    WorkspacePool* pool = plan_pool(device);
    WsGuard g(*pool, stream);
    Workspace& w = *g.w;
    // device-side results (the callers' buffers, or staging when they are host buffers) + fusion scratch
    size_t nv = (size_t)nq * kv, np = (size_t)nq * kp, nd = (size_t)nq * kd, nf = (size_t)nq * (kv + kp), cw = ((size_t)nq * 4 + 15) / 16 * 16;
    size_t stage_bytes = host ? nv * 8 + np * 8 + nd * 8 + 3 * cw + 2 * (size_t)nq * 8 + nf * 20 + cw : 0;
    size_t key_bytes = fuse ? (nv + np) * 8 : 0;
    ENSURE(w.scores, stage_bytes + key_bytes + 512);
    unsigned char* p = w.scores.as<unsigned char>();
    auto take = [&](size_t bytes) { unsigned char* q = p; p += (bytes + 15) / 16 * 16; return q; };
    uint32_t *d_vid = rs->vec_ids, *d_pdoc = rs->par_docs, *d_ddoc = rs->doc_docs, *d_frf = rs->fused_refs;
    float *d_vsc = rs->vec_scores, *d_psc = rs->par_scores, *d_dsc = rs->doc_scores;
    int32_t *d_vcnt = rs->vec_counts, *d_pcnt = rs->par_counts, *d_dcnt = rs->doc_counts, *d_fcnt = rs->fused_counts;
    uint64_t *d_ptot = rs->par_total, *d_dtot = rs->doc_total, *d_fkey = rs->fused_keys;
    double* d_fsc = rs->fused_scores;
    if (host) {
        if (rq->vec) { d_vid = (uint32_t*)take(nv * 4); d_vsc = (float*)take(nv * 4); d_vcnt = (int32_t*)take(cw); }
        if (rq->par) { d_pdoc = (uint32_t*)take(np * 4); d_psc = (float*)take(np * 4); d_pcnt = (int32_t*)take(cw); d_ptot = (uint64_t*)take((size_t)nq * 8); }
        if (rq->doc) { d_ddoc = (uint32_t*)take(nd * 4); d_dsc = (float*)take(nd * 4); d_dcnt = (int32_t*)take(cw); d_dtot = (uint64_t*)take((size_t)nq * 8); }
        if (fuse) { d_fkey = (uint64_t*)take(nf * 8); d_fsc = (double*)take(nf * 8); d_frf = (uint32_t*)take(nf * 4); d_fcnt = (int32_t*)take(cw); }
    }
    uint64_t *d_vkey = nullptr, *d_pkey = nullptr;
    if (fuse) { d_vkey = (uint64_t*)take(nv * 8); d_pkey = (uint64_t*)take(np * 8); }

    // fork: the three index searches run on their own streams, each ordered after whatever the caller enqueued before this call
    CU(cudaEventRecord(ps.fork, stream));
    for (int i = 0; i < 3; ++i) CU(cudaStreamWaitEvent(ps.s[i], ps.fork, 0));
    // (text searches first: with host queries they enqueue without waiting; the vector search may wait for a filter count)
    if (rq->par) {
        r = txt_search_impl(rq->par, rq->par_terms, rq->par_off, nq, host, false, rq->par_params, d_pdoc, d_psc, d_pcnt, d_ptot, ps.s[1]);
        if (r) return r;
        CU(cudaEventRecord(ps.join[1], ps.s[1]));
        CU(cudaStreamWaitEvent(stream, ps.join[1], 0));
    }
    if (rq->doc) {
        r = txt_search_impl(rq->doc, rq->doc_terms, rq->doc_off, nq, host, false, rq->doc_params, d_ddoc, d_dsc, d_dcnt, d_dtot, ps.s[2]);
        if (r) return r;
        CU(cudaEventRecord(ps.join[2], ps.s[2]));
        CU(cudaStreamWaitEvent(stream, ps.join[2], 0));
    }
    if (rq->vec) {
        r = vec_search_impl(rq->vec, rq->queries, nq, rq->ldq, host, false, rq->vec_params, d_vid, d_vsc, d_vcnt, ps.s[0], rq->formula, rq->n_formula);
        if (r) return r;
        CU(cudaEventRecord(ps.join[0], ps.s[0]));
        CU(cudaStreamWaitEvent(stream, ps.join[0], 0));
    }
    // join + rank fusion of the paragraph (keyword) and vector (semantic) results on the caller's stream
    if (fuse) {
        ids_to_keys_kernel<<<std::min<size_t>((nv + 255) / 256, 1024), 256, 0, stream>>>(d_vid, nv, rq->vec->d_par_of, rq->vec->d_par_keys, d_vkey);
        LAUNCHED();
        ids_to_keys_kernel<<<std::min<size_t>((np + 255) / 256, 1024), 256, 0, stream>>>(d_pdoc, np, nullptr, rq->par->d_doc_keys, d_pkey);
        LAUNCHED();
        RrfSourceDev src[2];
        RrfSourceDev kw{d_pkey, d_psc, d_pcnt, kp, rq->weight_keyword}, sem{d_vkey, d_vsc, d_vcnt, kv, rq->weight_semantic};
        src[0] = rq->semantic_first ? sem : kw;
        src[1] = rq->semantic_first ? kw : sem;
        r = rrf_launch(src, 2, nq, rq->rrf_k, d_fkey, d_fsc, d_frf, d_fcnt, stream);
        if (r) return r;
    }
    if (host) {
        auto back = [&](void* dst, const void* src, size_t bytes) { return dst ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream) : cudaSuccess; };
        if (rq->vec) { CU(back(rs->vec_ids, d_vid, nv * 4)); CU(back(rs->vec_scores, d_vsc, nv * 4)); CU(back(rs->vec_counts, d_vcnt, (size_t)nq * 4)); }
        if (rq->par) { CU(back(rs->par_docs, d_pdoc, np * 4)); CU(back(rs->par_scores, d_psc, np * 4)); CU(back(rs->par_counts, d_pcnt, (size_t)nq * 4)); CU(back(rs->par_total, d_ptot, (size_t)nq * 8)); }
        if (rq->doc) { CU(back(rs->doc_docs, d_ddoc, nd * 4)); CU(back(rs->doc_scores, d_dsc, nd * 4)); CU(back(rs->doc_counts, d_dcnt, (size_t)nq * 4)); CU(back(rs->doc_total, d_dtot, (size_t)nq * 8)); }
        if (fuse) { CU(back(rs->fused_keys, d_fkey, nf * 8)); CU(back(rs->fused_scores, d_fsc, nf * 8)); CU(back(rs->fused_refs, d_frf, nf * 4)); CU(back(rs->fused_counts, d_fcnt, (size_t)nq * 4)); }
        CU(cudaStreamSynchronize(stream));
    }
    return 0;
}

}  // extern "C"
