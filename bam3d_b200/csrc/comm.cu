// Multi-GPU exchange of the path: one process per GPU, pair lists shard with no data-path collective, and the
// one exchange step is the gather of the ranks' compacted contact lists (bam3d_contact40, hits only). NCCL over NVLink /
// NVSwitch, bound at run time: libnccl.so.2 is dlopen'ed on first use, so a single-GPU host needs no NCCL, and inside a
// process that already loaded a libnccl (PyTorch's) the same copy is used. Nothing here needs PyTorch.
//
//   rank 0:      bam3d_comm_unique_id(id)            -> 128 bytes, handed to the other ranks by any host-side means
//   every rank:  bam3d_comm_create(device, id, rank, n_ranks, &comm)
//   per step:    bam3d_compact_contacts_dev(...)     (hits only, tagged with the global pair index)
//                bam3d_gather_contacts_begin(ctx, comm, d_count, &ticket)
//                  ... enqueue the next step's kernels (other ctx / stream): the exchange overlaps them ...
//                bam3d_gather_contacts_finish(ctx, comm, ticket, d_compact, d_gathered, capacity, counts, &total)
//                bam3d_comm_join(ctx, comm)          (before the ctx stream reads d_gathered)
// `begin` all-gathers the ranks' record counts (ordered after the ctx stream) and copies them to pinned host memory;
// `finish` waits for that copy on the host — by then the GPU is busy with the next step — and enqueues one grouped
// ncclBroadcast per rank with the exact counts, so that d_gathered is the concatenation of the ranks' lists in rank order
// with no padding. Ranks must call begin / finish in the same order.
// Counts and payloads travel on TWO communicators, each with its own stream: in a pipelined loop `begin(k+1)` is issued
// before `finish(k)`, and on one in-order stream the payload of step k would queue behind the count exchange of step k+1,
// i.e. behind step k+1's kernels — the exchange would then serialise the steps it is meant to overlap. No compute kernel
// ever waits for an NCCL kernel while resident, so the two communicators cannot deadlock each other.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/bam3d_gpu.h"
#include "narrow.h"
#include "ctx.h"

namespace {

// the slice of nccl.h this file uses (ABI-stable since NCCL 2.0)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;   // 0 = ncclSuccess
typedef int ncclDataType_t;  // ncclUint8 = 1, ncclUint64 = 5
constexpr ncclDataType_t NCCL_UINT8 = 1, NCCL_UINT64 = 5;
static_assert(2 * sizeof(ncclUniqueId) == BAM3D_UNIQUE_ID_BYTES, "unique id size: one NCCL id per communicator");

// ncclConfig_t as NCCL 2.18 defined it (nccl.h: ncclConfig_v21700 with splitShare); later versions read it by its size / version fields
struct NcclConfig217 {
    size_t size;
    unsigned int magic;
    unsigned int version;
    int blocking, cgaClusterSize, minCTAs, maxCTAs;
    const char* netName;
    int splitShare;
};
constexpr int NCCL_UNDEF_INT = (int)0x80000000;  // NCCL_CONFIG_UNDEF_INT (INT_MIN)

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, NcclConfig217*) = nullptr;  // optional (NCCL >= 2.14)
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
    char why[256] = "";
};

NcclApi& nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    void* h = nullptr;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    if (const char* e = getenv("BAM3D_NCCL_LIB")) h = dlopen(e, RTLD_NOW | RTLD_GLOBAL);
    for (const char* nme : names) if (!h) h = dlopen(nme, RTLD_NOW | RTLD_GLOBAL);
    if (!h) { std::snprintf(api.why, sizeof api.why, "libnccl.so.2 not found (%s)", dlerror()); return api; }
    bool all = true;
    auto sym = [&](const char* nme) { void* p = dlsym(h, nme); if (!p) { all = false; std::snprintf(api.why, sizeof api.why, "%s missing from libnccl", nme); } return p; };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.ok = all;
    api.CommInitRankConfig = (decltype(api.CommInitRankConfig))dlsym(h, "ncclCommInitRankConfig");
    return api;
}

// A communicator limited to `max_ctas` thread blocks per collective. The exchange kernels mostly WAIT — for the slowest rank to
// reach the same collective — and a waiting NCCL block holds a multiprocessor the compute lanes could use: measured on the C5
// step (2 GPUs, payload kernels alive 9.4 ms per step for < 1 ms of transfer), 43.7 ms per step with NCCL's default block count,
// 41.9 ms with 2 blocks; the volume (104 MB per rank and step) needs no more. Falls back to the default configuration when
// the library has no ncclCommInitRankConfig.
ncclResult_t comm_init(NcclApi& N, ncclComm_t* out, int n_ranks, const ncclUniqueId& id, int rank, int max_ctas) {
    if (N.CommInitRankConfig && max_ctas > 0) {
        NcclConfig217 cfg;
        cfg.size = sizeof cfg; cfg.magic = 0xcafebeefu; cfg.version = 21800;
        cfg.blocking = NCCL_UNDEF_INT; cfg.cgaClusterSize = NCCL_UNDEF_INT;
        cfg.minCTAs = 1; cfg.maxCTAs = max_ctas;
        cfg.netName = nullptr; cfg.splitShare = NCCL_UNDEF_INT;
        return N.CommInitRankConfig(out, n_ranks, id, rank, &cfg);
    }
    return N.CommInitRank(out, n_ranks, id, rank);
}

constexpr int TICKETS = 16;  // count exchanges in flight

}  // namespace

struct bam3d_comm {
    int device = 0, rank = 0, n_ranks = 1;
    ncclComm_t comm_counts = nullptr, comm = nullptr;  // record counts | payload
    cudaStream_t s_counts = nullptr, stream = nullptr;
    cudaEvent_t ordered = nullptr, done = nullptr;  // ctx stream -> comm streams, comm streams -> ctx stream
    cudaEvent_t exchanged[16] = {};  // per ticket: its payload has arrived
    uint64_t* d_counts = nullptr;                   // TICKETS x n_ranks
    uint64_t* h_counts = nullptr;                   // pinned, TICKETS x n_ranks
    cudaEvent_t counted[TICKETS] = {};
    uint64_t next_ticket = 0;
    // payload as ONE all-gather of max-count records per rank into a staging buffer + exact-size local copies (default from 3
    // ranks up; BAM3D_COMM_ALLGATHER=0/1 forces it): measured on 8 GPUs, 7 x 104 MB received per rank and step, 44.6 ms per C5 step
    // against 46.1 ms with eight grouped broadcasts (same block count)
    bool allgather = false;
    unsigned char* stage = nullptr;  // one buffer for all tickets: the payloads run one after the other on `stream`
    size_t stage_bytes = 0;
    // BAM3D_COMM_TRACE=1 (development): host time blocked on the counts and device time of the payload launches, printed at destroy
    bool trace = false;
    cudaEvent_t t0[TICKETS] = {}, t1[TICKETS] = {}, tc0[TICKETS] = {};
    bool t_open[TICKETS] = {};
    double host_wait_ms = 0, payload_ms = 0, payload_max_ms = 0, counts_ms = 0, counts_max_ms = 0;
    uint64_t traced = 0;
    void collect(int slot) {
        if (!trace || !t_open[slot]) return;
        float ms = 0, mc = 0;
        if (cudaEventSynchronize(t1[slot]) == cudaSuccess && cudaEventElapsedTime(&ms, t0[slot], t1[slot]) == cudaSuccess) {
            payload_ms += ms; if (ms > payload_max_ms) payload_max_ms = ms;
        }
        if (cudaEventElapsedTime(&mc, tc0[slot], counted[slot]) == cudaSuccess) { counts_ms += mc; if (mc > counts_max_ms) counts_max_ms = mc; }
        ++traced;
        t_open[slot] = false;
    }
};

static int32_t nccl_fail(bam3d_ctx* ctx, const char* what, ncclResult_t r) {
    char buf[384];
    std::snprintf(buf, sizeof buf, "%s: %s", what, nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error");
    return b3_fail(ctx, BAM3D_ERR_COMM, buf, cudaSuccess);
}

extern "C" {

int32_t bam3d_comm_unique_id(uint8_t* out_id) {
    if (!out_id) return BAM3D_ERR_ARG;
    NcclApi& N = nccl();
    if (!N.ok) return BAM3D_ERR_COMM;
    for (int k = 0; k < 2; ++k) {
        ncclUniqueId id;
        if (N.GetUniqueId(&id) != 0) return BAM3D_ERR_COMM;
        std::memcpy(out_id + k * sizeof(ncclUniqueId), id.internal, sizeof(ncclUniqueId));
    }
    return BAM3D_OK;
}

void bam3d_comm_destroy(bam3d_comm* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->s_counts) cudaStreamSynchronize(c->s_counts);
    if (c->trace) {
        for (int t = 0; t < TICKETS; ++t) c->collect(t);
        std::fprintf(stderr, "[bam3d comm rank %d] %llu exchanges: host blocked on counts %.3f ms avg; count exchange %.3f ms avg (max %.3f) on the device; payload %.3f ms avg (max %.3f)\n",
                     c->rank, (unsigned long long)c->traced, c->traced ? c->host_wait_ms / c->traced : 0.0, c->traced ? c->counts_ms / c->traced : 0.0, c->counts_max_ms,
                     c->traced ? c->payload_ms / c->traced : 0.0, c->payload_max_ms);
        for (int t = 0; t < TICKETS; ++t) { if (c->t0[t]) cudaEventDestroy(c->t0[t]); if (c->t1[t]) cudaEventDestroy(c->t1[t]); if (c->tc0[t]) cudaEventDestroy(c->tc0[t]); }
    }
    if (c->comm) nccl().CommDestroy(c->comm);
    if (c->comm_counts) nccl().CommDestroy(c->comm_counts);
    for (cudaEvent_t e : c->counted) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : c->exchanged) if (e) cudaEventDestroy(e);
    if (c->s_counts) cudaStreamDestroy(c->s_counts);
    if (c->ordered) cudaEventDestroy(c->ordered);
    if (c->done) cudaEventDestroy(c->done);
    if (c->d_counts) cudaFree(c->d_counts);
    if (c->stage) cudaFree(c->stage);
    if (c->h_counts) cudaFreeHost(c->h_counts);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int32_t bam3d_comm_create(int32_t device, const uint8_t* id, int32_t rank, int32_t n_ranks, bam3d_comm** out) {
    if (!out || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return BAM3D_ERR_ARG;
    *out = nullptr;
    NcclApi& N = nccl();
    if (!N.ok) return BAM3D_ERR_COMM;
    if (cudaSetDevice(device) != cudaSuccess) return BAM3D_ERR_CUDA;
    bam3d_comm* c = new (std::nothrow) bam3d_comm();
    if (!c) return BAM3D_ERR_OOM;
    c->device = device; c->rank = rank; c->n_ranks = n_ranks;
    // Exchange streams at the default priority. (BAM3D_COMM_PRIORITY=1 puts them at the highest: measured on 2 GPUs it makes the
    // C5 step SLOWER, 51.7 vs 44.9 ms — the NCCL kernels then displace blocks of the persistent compute kernels, whose
    // remaining blocks carry the tail.)
    int prio_least = 0, prio = 0;
    bool high = false;
    if (const char* e = std::getenv("BAM3D_COMM_PRIORITY")) high = std::atoi(e) != 0;
    if (!high || cudaDeviceGetStreamPriorityRange(&prio_least, &prio) != cudaSuccess) prio = 0;
    bool ok = cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio) == cudaSuccess &&
              cudaStreamCreateWithPriority(&c->s_counts, cudaStreamNonBlocking, prio) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ordered, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming) == cudaSuccess &&
              cudaMalloc(&c->d_counts, sizeof(uint64_t) * TICKETS * n_ranks) == cudaSuccess &&
              cudaHostAlloc(&c->h_counts, sizeof(uint64_t) * TICKETS * n_ranks, cudaHostAllocDefault) == cudaSuccess;
    if (const char* e = std::getenv("BAM3D_COMM_TRACE")) c->trace = std::atoi(e) != 0;
    c->allgather = n_ranks >= 3;
    if (const char* e = std::getenv("BAM3D_COMM_ALLGATHER")) c->allgather = std::atoi(e) != 0;
    for (int t = 0; ok && c->trace && t < TICKETS; ++t)
        ok = cudaEventCreate(&c->t0[t]) == cudaSuccess && cudaEventCreate(&c->t1[t]) == cudaSuccess && cudaEventCreate(&c->tc0[t]) == cudaSuccess;
    for (int t = 0; ok && t < TICKETS; ++t)
        ok = cudaEventCreateWithFlags(&c->counted[t], c->trace ? cudaEventDefault : cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->exchanged[t], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { bam3d_comm_destroy(c); return BAM3D_ERR_CUDA; }
    ncclUniqueId uid;
    // thread blocks per payload collective (BAM3D_NCCL_MAX_CTAS; 0: NCCL's default). Measured on the C5 step: 2 ranks 41.9 ms with 4
    // (44.5 with NCCL's default); 8 ranks 44.6 ms with 4, 43.1 ms with 8 (41.8 ms with the exchange switched off).
    int payload_ctas = n_ranks <= 2 ? 4 : 8;
    if (const char* e = std::getenv("BAM3D_NCCL_MAX_CTAS")) payload_ctas = std::atoi(e);
    std::memcpy(uid.internal, id, sizeof uid);
    if (comm_init(N, &c->comm_counts, n_ranks, uid, rank, 1) != 0) { c->comm_counts = nullptr; bam3d_comm_destroy(c); return BAM3D_ERR_COMM; }
    std::memcpy(uid.internal, id + sizeof uid, sizeof uid);
    if (comm_init(N, &c->comm, n_ranks, uid, rank, payload_ctas) != 0) { c->comm = nullptr; bam3d_comm_destroy(c); return BAM3D_ERR_COMM; }
    *out = c;
    return BAM3D_OK;
}

int32_t bam3d_comm_rank(const bam3d_comm* c) { return c ? c->rank : -1; }
int32_t bam3d_comm_size(const bam3d_comm* c) { return c ? c->n_ranks : 0; }
void* bam3d_comm_stream(bam3d_comm* c) { return c ? (void*)c->stream : nullptr; }
const char* bam3d_comm_unavailable_reason(void) { NcclApi& N = nccl(); return N.ok ? "" : N.why; }

int32_t bam3d_gather_contacts_begin(bam3d_ctx* ctx, bam3d_comm* c, const uint64_t* d_count, uint64_t* out_ticket) {
    if (!ctx || !c || !d_count || !out_ticket) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_gather_contacts_begin: null argument", cudaSuccess);
    NcclApi& N = nccl();
    cudaError_t e;
    if ((e = cudaSetDevice(c->device)) != cudaSuccess) return b3_fail(ctx, BAM3D_ERR_CUDA, "cudaSetDevice", e);
    const uint64_t ticket = c->next_ticket++;
    const int slot = (int)(ticket % TICKETS);
    uint64_t* d = c->d_counts + (size_t)slot * c->n_ranks;
    uint64_t* h = c->h_counts + (size_t)slot * c->n_ranks;
    // the count exchange continues after everything enqueued on the ctx stream so far (the compaction that wrote d_count)
    if ((e = cudaEventRecord(c->ordered, ctx->stream)) != cudaSuccess || (e = cudaStreamWaitEvent(c->s_counts, c->ordered, 0)) != cudaSuccess)
        return b3_fail(ctx, BAM3D_ERR_CUDA, "bam3d_gather_contacts_begin: stream ordering", e);
    if (c->trace) { c->collect(slot); cudaEventRecord(c->tc0[slot], c->s_counts); }
    ncclResult_t r = N.AllGather(d_count, d, 1, NCCL_UINT64, c->comm_counts, c->s_counts);
    if (r != 0) return nccl_fail(ctx, "ncclAllGather(counts)", r);
    if ((e = cudaMemcpyAsync(h, d, sizeof(uint64_t) * c->n_ranks, cudaMemcpyDeviceToHost, c->s_counts)) != cudaSuccess ||
        (e = cudaEventRecord(c->counted[slot], c->s_counts)) != cudaSuccess)
        return b3_fail(ctx, BAM3D_ERR_CUDA, "bam3d_gather_contacts_begin: counts copy", e);
    *out_ticket = ticket;
    return BAM3D_OK;
}

int32_t bam3d_gather_contacts_finish(bam3d_ctx* ctx, bam3d_comm* c, uint64_t ticket, const bam3d_contact40* d_compact,
                                     bam3d_contact40* d_gathered, uint64_t capacity, uint64_t* out_counts, uint64_t* out_total) {
    if (!ctx || !c || !out_total) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_gather_contacts_finish: null argument", cudaSuccess);
    if (ticket >= c->next_ticket || ticket + TICKETS < c->next_ticket)
        return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_gather_contacts_finish: unknown ticket (at most 16 exchanges may be in flight)", cudaSuccess);
    NcclApi& N = nccl();
    cudaError_t e;
    if ((e = cudaSetDevice(c->device)) != cudaSuccess) return b3_fail(ctx, BAM3D_ERR_CUDA, "cudaSetDevice", e);
    const int slot = (int)(ticket % TICKETS);
    const auto hw0 = std::chrono::steady_clock::now();
    if ((e = cudaEventSynchronize(c->counted[slot])) != cudaSuccess) return b3_fail(ctx, BAM3D_ERR_CUDA, "bam3d_gather_contacts_finish: counts", e);
    if (c->trace) c->host_wait_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - hw0).count();
    const uint64_t* h = c->h_counts + (size_t)slot * c->n_ranks;
    uint64_t total = 0;
    for (int r = 0; r < c->n_ranks; ++r) { if (out_counts) out_counts[r] = h[r]; total += h[r]; }
    *out_total = total;
    if (total > capacity) {
        char msg[160];
        std::snprintf(msg, sizeof msg, "bam3d_gather_contacts: %llu records over all ranks, capacity %llu", (unsigned long long)total, (unsigned long long)capacity);
        return b3_fail(ctx, BAM3D_ERR_CAPACITY, msg, cudaSuccess);
    }
    if (total == 0) { cudaEventRecord(c->exchanged[slot], c->stream); return BAM3D_OK; }
    if (!d_gathered || (h[c->rank] && !d_compact)) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_gather_contacts_finish: null buffer", cudaSuccess);
    if (c->trace) { cudaEventRecord(c->t0[slot], c->stream); c->t_open[slot] = true; }
    if (c->allgather) {
        uint64_t maxc = 0;
        for (int r = 0; r < c->n_ranks; ++r) maxc = h[r] > maxc ? h[r] : maxc;
        const size_t per_rank = (size_t)maxc * sizeof(bam3d_contact40);
        const size_t need = per_rank * (size_t)c->n_ranks;
        if (c->stage_bytes < need) {  // grows during the first steps only (12 % headroom)
            if (c->stage) { cudaStreamSynchronize(c->stream); cudaFree(c->stage); c->stage = nullptr; c->stage_bytes = 0; }
            const size_t want = need + need / 8;
            if ((e = cudaMalloc(&c->stage, want)) != cudaSuccess) return b3_fail(ctx, BAM3D_ERR_OOM, "bam3d_gather_contacts_finish: staging buffer", e);
            c->stage_bytes = want;
        }
        unsigned char* st = c->stage;
        unsigned char* mine = st + per_rank * (size_t)c->rank;
        if (h[c->rank] && (e = cudaMemcpyAsync(mine, d_compact, h[c->rank] * sizeof(bam3d_contact40), cudaMemcpyDeviceToDevice, c->stream)) != cudaSuccess)
            return b3_fail(ctx, BAM3D_ERR_CUDA, "bam3d_gather_contacts_finish: stage copy", e);
        ncclResult_t ra = N.AllGather(mine, st, per_rank, NCCL_UINT8, c->comm, c->stream);  // in place: every rank's slot is maxc records wide
        if (ra != 0) return nccl_fail(ctx, "ncclAllGather(contacts)", ra);
        uint64_t off2 = 0;
        for (int r = 0; r < c->n_ranks; ++r) {  // exact sizes, rank order: the same contiguous list the broadcast form produces
            if (h[r] && (e = cudaMemcpyAsync(reinterpret_cast<unsigned char*>(d_gathered) + off2 * sizeof(bam3d_contact40), st + per_rank * (size_t)r,
                                             h[r] * sizeof(bam3d_contact40), cudaMemcpyDeviceToDevice, c->stream)) != cudaSuccess)
                return b3_fail(ctx, BAM3D_ERR_CUDA, "bam3d_gather_contacts_finish: unstage copy", e);
            off2 += h[r];
        }
        if (c->trace) cudaEventRecord(c->t1[slot], c->stream);
        if ((e = cudaEventRecord(c->exchanged[slot], c->stream)) != cudaSuccess) return b3_fail(ctx, BAM3D_ERR_CUDA, "cudaEventRecord", e);
        return BAM3D_OK;
    }
    // exact sizes: one broadcast per rank, grouped into a single NCCL launch; rank r's records land behind those of ranks < r
    ncclResult_t r0 = N.GroupStart();
    if (r0 != 0) return nccl_fail(ctx, "ncclGroupStart", r0);
    uint64_t off = 0;
    ncclResult_t bad = 0;
    for (int r = 0; r < c->n_ranks; ++r) {
        if (h[r]) {
            ncclResult_t rr = N.Broadcast(d_compact, reinterpret_cast<unsigned char*>(d_gathered) + off * sizeof(bam3d_contact40),
                                          h[r] * sizeof(bam3d_contact40), NCCL_UINT8, r, c->comm, c->stream);
            if (rr != 0 && bad == 0) bad = rr;
        }
        off += h[r];
    }
    r0 = N.GroupEnd();
    if (bad != 0) return nccl_fail(ctx, "ncclBroadcast(contacts)", bad);
    if (r0 != 0) return nccl_fail(ctx, "ncclGroupEnd", r0);
    if (c->trace) cudaEventRecord(c->t1[slot], c->stream);
    if ((e = cudaEventRecord(c->exchanged[slot], c->stream)) != cudaSuccess) return b3_fail(ctx, BAM3D_ERR_CUDA, "cudaEventRecord", e);
    return BAM3D_OK;
}

int32_t bam3d_gather_contacts_wait(bam3d_ctx* ctx, bam3d_comm* c, uint64_t ticket) {
    if (!ctx || !c) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_gather_contacts_wait: null argument", cudaSuccess);
    if (ticket >= c->next_ticket || ticket + TICKETS < c->next_ticket) return BAM3D_OK;  // long gone
    cudaError_t e = cudaStreamWaitEvent(ctx->stream, c->exchanged[ticket % TICKETS], 0);
    if (e != cudaSuccess) return b3_fail(ctx, BAM3D_ERR_CUDA, "bam3d_gather_contacts_wait", e);
    return BAM3D_OK;
}

int32_t bam3d_comm_join(bam3d_ctx* ctx, bam3d_comm* c) {
    if (!ctx || !c) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_comm_join: null argument", cudaSuccess);
    cudaError_t e;
    if ((e = cudaEventRecord(c->done, c->stream)) != cudaSuccess || (e = cudaStreamWaitEvent(ctx->stream, c->done, 0)) != cudaSuccess ||
        (e = cudaEventRecord(c->done, c->s_counts)) != cudaSuccess || (e = cudaStreamWaitEvent(ctx->stream, c->done, 0)) != cudaSuccess)
        return b3_fail(ctx, BAM3D_ERR_CUDA, "bam3d_comm_join", e);
    return BAM3D_OK;
}

int32_t bam3d_gather_contacts_dev(bam3d_ctx* ctx, bam3d_comm* c, const bam3d_contact40* d_compact, const uint64_t* d_count,
                                  bam3d_contact40* d_gathered, uint64_t capacity, uint64_t* out_counts, uint64_t* out_total) {
    uint64_t ticket = 0;
    int32_t rc;
    if ((rc = bam3d_gather_contacts_begin(ctx, c, d_count, &ticket))) return rc;
    if ((rc = bam3d_gather_contacts_finish(ctx, c, ticket, d_compact, d_gathered, capacity, out_counts, out_total))) return rc;
    return bam3d_comm_join(ctx, c);
}

// Host-buffer form: this rank's dense results (as bam3d_gjk_contact returned them) in, the contact list of ALL ranks out.
int32_t bam3d_gather_contacts(bam3d_ctx* ctx, bam3d_comm* c, const bam3d_contact* contacts, const uint8_t* status, uint64_t n, uint64_t base,
                              bam3d_contact40* out, uint64_t capacity, uint64_t* out_counts, uint64_t* out_total) {
    if (!ctx || !c || !out_total || (n && (!contacts || !status)) || (capacity && !out))
        return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_gather_contacts: null argument", cudaSuccess);
    cudaError_t e;
    if ((e = cudaSetDevice(ctx->device)) != cudaSuccess) return b3_fail(ctx, BAM3D_ERR_CUDA, "cudaSetDevice", e);
    int32_t rc;
    const size_t nn = n ? n : 1, cap = capacity ? capacity : 1;
    if ((rc = ctx->contacts.reserve(ctx, nn * 32)) || (rc = ctx->status.reserve(ctx, nn)) || (rc = ctx->f0.reserve(ctx, nn * 40)) ||
        (rc = ctx->f1.reserve(ctx, cap * 40)) || (rc = ctx->small.reserve(ctx, 1024))) return rc;
    unsigned long long* d_count = reinterpret_cast<unsigned long long*>(static_cast<char*>(ctx->small.ptr) + 960);
    if (n) {
        cudaMemcpyAsync(ctx->contacts.ptr, contacts, n * 32, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(ctx->status.ptr, status, n, cudaMemcpyHostToDevice, ctx->stream);
    }
    ctx->launches += b3::launch_compact_contacts(ctx->stream, (const float4*)ctx->contacts.ptr, (const uint8_t*)ctx->status.ptr, n, base, ctx->f0.ptr, n, d_count);
    if ((e = cudaGetLastError()) != cudaSuccess) return b3_fail(ctx, BAM3D_ERR_CUDA, "compact_contacts_kernel", e);
    if ((rc = bam3d_gather_contacts_dev(ctx, c, (const bam3d_contact40*)ctx->f0.ptr, (const uint64_t*)d_count, (bam3d_contact40*)ctx->f1.ptr, capacity,
                                        out_counts, out_total))) return rc;
    if (*out_total) cudaMemcpyAsync(out, ctx->f1.ptr, *out_total * 40, cudaMemcpyDeviceToHost, ctx->stream);
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return b3_fail(ctx, BAM3D_ERR_CUDA, "bam3d_gather_contacts", e);
    return BAM3D_OK;
}

}  // extern "C"
