// libnrldpc: handle management, error reporting, host-side table queries (see include/nrldpc.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "nr_bg_tables.h"
#include "nrldpc_internal.cuh"

static thread_local char g_err[512] = "";

void nr_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* nrldpc_last_error(void) { return g_err; }
extern "C" int nrldpc_version(void) { return NRLDPC_VERSION; }

extern "C" int nrldpc_lifting_set_index(int zc)
{
    for (int i = 0; i < NR_NUM_LIFTING_SETS; i++)
        for (int j = 0; j < 8; j++)
            if (NR_LIFTING_SETS[i][j] == zc && zc > 0) return i;
    return -1;
}

extern "C" int nrldpc_graph_info(int bg, int* rows, int* cols, int* sys_cols, int* edges)
{
    if (bg != 1 && bg != 2) { nr_set_error("'baseGraphNo' must be 1 or 2!"); return NRLDPC_ERR_ARG; }
    if (rows) *rows = bg == 1 ? NR_BG1_ROWS : NR_BG2_ROWS;
    if (cols) *cols = bg == 1 ? NR_BG1_COLS : NR_BG2_COLS;
    if (sys_cols) *sys_cols = bg == 1 ? 22 : 10;
    if (edges) *edges = bg == 1 ? NR_BG1_EDGES : NR_BG2_EDGES;
    return NRLDPC_OK;
}

int nr_build_graph(int bg, int zc, NrGraph* g)
{
    if (bg != 1 && bg != 2) { nr_set_error("'baseGraphNo' must be 1 or 2!"); return NRLDPC_ERR_ARG; }
    const int ils = nrldpc_lifting_set_index(zc);
    if (ils < 0) { nr_set_error("illegal lifting size %d", zc); return NRLDPC_ERR_ARG; }
    memset(g, 0, sizeof(*g));
    g->P = bg == 1 ? NR_BG1_ROWS : NR_BG2_ROWS;
    g->ncols = bg == 1 ? NR_BG1_COLS : NR_BG2_COLS;
    g->ksys = bg == 1 ? 22 : 10;
    g->ncore = g->ksys + 4;
    g->Z = zc;
    const uint8_t* deg = bg == 1 ? NR_BG1_ROW_DEG : NR_BG2_ROW_DEG;
    const uint8_t* col = bg == 1 ? NR_BG1_COL : NR_BG2_COL;
    const uint16_t* sh = bg == 1 ? NR_BG1_SHIFT[ils] : NR_BG2_SHIFT[ils];
    int e = 0;
    for (int i = 0; i < g->P; i++) {
        g->rowEdge0[i] = (uint16_t)e;
        for (int j = 0; j < deg[i]; j++, e++) g->edge[e] = ((uint32_t)col[e] << 16) | (uint32_t)(sh[e] % zc);
    }
    g->rowEdge0[g->P] = (uint16_t)e;
    g->rowEdge0[g->P + 1] = (uint16_t)e;
    return 0;
}

extern "C" int nrldpc_base_graph(int bg, int set_index, int zc, int16_t* out)
{
    if (bg != 1 && bg != 2) { nr_set_error("'baseGraphNo' must be 1 or 2!"); return NRLDPC_ERR_ARG; }
    if (zc <= 0 || !out) { nr_set_error("base_graph: bad argument"); return NRLDPC_ERR_ARG; }
    if (set_index < 0) set_index = nrldpc_lifting_set_index(zc);
    if (set_index < 0 || set_index > 7) { nr_set_error("illegal lifting size %d", zc); return NRLDPC_ERR_ARG; }
    const int P = bg == 1 ? NR_BG1_ROWS : NR_BG2_ROWS, n = bg == 1 ? NR_BG1_COLS : NR_BG2_COLS;
    const uint8_t* deg = bg == 1 ? NR_BG1_ROW_DEG : NR_BG2_ROW_DEG;
    const uint8_t* col = bg == 1 ? NR_BG1_COL : NR_BG2_COL;
    const uint16_t* sh = bg == 1 ? NR_BG1_SHIFT[set_index] : NR_BG2_SHIFT[set_index];
    for (int i = 0; i < P * n; i++) out[i] = -1;
    int e = 0;
    for (int i = 0; i < P; i++)
        for (int j = 0; j < deg[i]; j++, e++) out[i * n + col[e]] = (int16_t)(sh[e] % zc);
    return NRLDPC_OK;
}

extern "C" int nrldpc_create(int device, nrldpc_handle** out)
{
    if (!out) { nr_set_error("create: null out"); return NRLDPC_ERR_ARG; }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        nr_set_error("no CUDA device available (%s); libnrldpc has no CPU fallback",
                     e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return NRLDPC_ERR_CUDA;
    }
    if (device < 0 || device >= count) { nr_set_error("create: device %d out of range", device); return NRLDPC_ERR_ARG; }
    NR_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    NR_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        nr_set_error("device %d is sm_%d%d; libnrldpc is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return NRLDPC_ERR_CUDA;
    }
    nrldpc_handle* h = (nrldpc_handle*)calloc(1, sizeof(nrldpc_handle));
    if (!h) return NRLDPC_ERR_NOMEM;
    h->device = device;
    h->numSMs = prop.multiProcessorCount;
    h->maxSmemOptin = (int)prop.sharedMemPerBlockOptin;
    h->smemPerSM = (int)prop.sharedMemPerMultiprocessor;
    const char* occ = getenv("NRLDPC_DEC_OCC");
    h->decOcc = occ ? atoi(occ) : 0;
    const char* nt = getenv("NRLDPC_NO_TMEM");
    h->noTmem = nt ? atoi(nt) : 0;
    const char* ns = getenv("NRLDPC_NO_STATIC_ROWS");
    h->noStaticRows = ns ? atoi(ns) : 0;
    const char* nst = getenv("NRLDPC_NO_STAGE");
    h->noStage = nst ? atoi(nst) : 0;
    e = cudaMalloc(&h->workCounter, 16 * sizeof(unsigned int));
    if (e != cudaSuccess) { free(h); nr_set_error("cudaMalloc failed: %s", cudaGetErrorString(e)); return NRLDPC_ERR_CUDA; }
    cudaMemset(h->workCounter, 0, 16 * sizeof(unsigned int));
    {   // [5]: running minimum of the iteration counts (NRLDPC_DEC_ES_AUTO)
        const unsigned int big = 0x7fffffffu;
        cudaMemcpy(h->workCounter + 5, &big, sizeof(big), cudaMemcpyHostToDevice);
    }
    *out = h;
    return NRLDPC_OK;
}

extern "C" int nrldpc_destroy(nrldpc_handle* h)
{
    if (!h) return NRLDPC_OK;
    cudaSetDevice(h->device);
    if (h->scratch) cudaFree(h->scratch);
    if (h->tmp) cudaFree(h->tmp);
    if (h->tmp2) cudaFree(h->tmp2);
    if (h->goldTables) cudaFree(h->goldTables);
    for (int i = 0; i < 4; i++) {
        if (h->sub[i]) nrldpc_destroy(h->sub[i]);
        if (h->subStream[i]) cudaStreamDestroy(h->subStream[i]);
        if (h->subJoin[i]) cudaEventDestroy(h->subJoin[i]);
    }
    if (h->subFork) cudaEventDestroy(h->subFork);
    if (h->crcFacDev) cudaFree(h->crcFacDev);
    if (h->tbAcc) cudaFree(h->tbAcc);
    if (h->tbFacDev) cudaFree(h->tbFacDev);
    if (h->workCounter) cudaFree(h->workCounter);
    if (h->symLlr) cudaFree(h->symLlr);
    free(h);
    return NRLDPC_OK;
}

int nr_reserve_tmp2(nrldpc_handle* h, size_t bytes, void** out)
{
    if (bytes > h->tmp2Bytes) {
        if (h->tmp2) NR_CUDA_CHECK(cudaFree(h->tmp2));
        h->tmp2 = nullptr;
        h->tmp2Bytes = 0;
        NR_CUDA_CHECK(cudaMalloc(&h->tmp2, bytes));
        h->tmp2Bytes = bytes;
    }
    *out = h->tmp2;
    return NRLDPC_OK;
}

int nr_reserve_tmp(nrldpc_handle* h, size_t bytes, void** out)
{
    if (bytes > h->tmpBytes) {
        if (h->tmp) NR_CUDA_CHECK(cudaFree(h->tmp));
        h->tmp = nullptr;
        h->tmpBytes = 0;
        NR_CUDA_CHECK(cudaMalloc(&h->tmp, bytes));
        h->tmpBytes = bytes;
    }
    *out = h->tmp;
    return NRLDPC_OK;
}

// ---- unified-memory buffers (include/nrldpc.h, "Unified-memory buffers") ----------------------------------------------
extern "C" int nrldpc_managed_supported(nrldpc_handle* h)
{
    if (!h) return 0;
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrConcurrentManagedAccess, h->device) != cudaSuccess) return 0;
    return v ? 1 : 0;
}

extern "C" int nrldpc_managed_alloc(nrldpc_handle* h, uint64_t bytes, int zero, void** out, nrldpc_stream stream)
{
    if (!h || !out || bytes == 0) { nr_set_error("managed_alloc: bad argument"); return NRLDPC_ERR_ARG; }
    *out = nullptr;
    NR_CUDA_CHECK(cudaSetDevice(h->device));
    void* p = nullptr;
    NR_CUDA_CHECK(cudaMallocManaged(&p, (size_t)bytes, cudaMemAttachGlobal));
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemPrefetchAsync(p, (size_t)bytes, h->device, s);   // first touch on the device: no page faults in the kernels
    if (e == cudaSuccess && zero) e = cudaMemsetAsync(p, 0, (size_t)bytes, s);
    if (e != cudaSuccess) {
        cudaFree(p);
        nr_set_error("managed_alloc: %s", cudaGetErrorString(e));
        return NRLDPC_ERR_CUDA;
    }
    *out = p;
    return NRLDPC_OK;
}

extern "C" int nrldpc_managed_free(nrldpc_handle* h, void* p)
{
    if (!p) return NRLDPC_OK;
    if (h) cudaSetDevice(h->device);
    NR_CUDA_CHECK(cudaFree(p));
    return NRLDPC_OK;
}

extern "C" int nrldpc_managed_prefetch(nrldpc_handle* h, void* p, uint64_t bytes, int to_device, nrldpc_stream stream)
{
    if (!h || !p) { nr_set_error("managed_prefetch: bad argument"); return NRLDPC_ERR_ARG; }
    NR_CUDA_CHECK(cudaSetDevice(h->device));
    NR_CUDA_CHECK(cudaMemPrefetchAsync(p, (size_t)bytes, to_device ? h->device : cudaCpuDeviceId, (cudaStream_t)stream));
    return NRLDPC_OK;
}

extern "C" int nrldpc_managed_clear(nrldpc_handle* h, void* p, uint64_t bytes, nrldpc_stream stream)
{
    if (!h || !p) { nr_set_error("managed_clear: bad argument"); return NRLDPC_ERR_ARG; }
    NR_CUDA_CHECK(cudaSetDevice(h->device));
    NR_CUDA_CHECK(cudaMemsetAsync(p, 0, (size_t)bytes, (cudaStream_t)stream));
    return NRLDPC_OK;
}
