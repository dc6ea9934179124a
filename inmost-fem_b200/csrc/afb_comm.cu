// Multi-GPU interface exchange inside the library: NCCL point-to-point over NVLink, driven from C.
//
// Reference context: under MPI the reference assembles every owned row completely on its owner by recomputing the ghost cells
// (inmost_interface/assembler.inl:162-183) and exchanges only the numbering (global_enumerator.cpp:594-604, :698, :754).  Here
// every rank assembles ITS elements once into (owned rows ++ interface rows of other ranks) and the interface contributions
// travel to their owners (proposal of SURVEY 8b: a communicator in the context + afb_halo_exchange):
//     send:  the tail of the extended value / rhs arrays IS the send buffer (the foreign rows are sorted by owner),
//     recv:  one contiguous segment per peer, added with precomputed slots in RANK ORDER (distinct slots per peer => no atomics,
//            bit-reproducible), on the context's stream.
// The exchange runs on its own stream between two events, so it overlaps the second phase of a phased assembly
// (afb_assemble_distributed).  NCCL is loaded with dlopen("libnccl.so.2") on first use: the library has no link-time dependency
// on it, and inside a process that already uses NCCL (torch.distributed) the same copy is picked up.
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "afb_internal.h"

using namespace afb;

namespace {

// the part of nccl.h this file needs (stable since NCCL 2.7: point-to-point API)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;      // 0 = ncclSuccess
constexpr int kNcclFloat64 = 8;   // ncclDataType_t: ncclFloat64 / ncclDouble

struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string why;
};

NcclApi* nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {getenv("AFB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n) continue;
            api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.h) break;
        }
        if (!api.h) { api.why = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : ""); return; }
#define AFB_SYM(field, name)                                                            \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.h, name));             \
    if (!api.field) { api.why = std::string("NCCL symbol missing: ") + name; api.h = nullptr; return; }
        AFB_SYM(GetUniqueId, "ncclGetUniqueId")
        AFB_SYM(CommInitRank, "ncclCommInitRank")
        AFB_SYM(CommDestroy, "ncclCommDestroy")
        AFB_SYM(Send, "ncclSend")
        AFB_SYM(Recv, "ncclRecv")
        AFB_SYM(GroupStart, "ncclGroupStart")
        AFB_SYM(GroupEnd, "ncclGroupEnd")
        AFB_SYM(GetErrorString, "ncclGetErrorString")
#undef AFB_SYM
    });
    return api.h ? &api : nullptr;
}

int nccl_fail(afb_ctx* ctx, ncclResult_t r, const char* what) {
    NcclApi* a = nccl();
    set_error(ctx, (std::string(what) + ": " + (a ? a->GetErrorString(r) : "NCCL error")).c_str());
    return -8;
}
#define AFB_NCCL(ctx, call)                                      \
    do {                                                         \
        ncclResult_t _r = (call);                                \
        if (_r != 0) return nccl_fail(ctx, _r, #call);           \
    } while (0)

inline unsigned grid_for(long long n) { return (unsigned)std::max<long long>(1, std::min<long long>((n + 255) / 256, 148LL * 32)); }

// dst[slot[k]] += src[k]; the slots of one peer are distinct
__global__ void k_halo_add_peer(long long n, const long long* __restrict__ slot, const double* __restrict__ src, double* __restrict__ dst) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) dst[slot[k]] += src[k];
}

int ensure_streams(afb_ctx* ctx) {
    if (!ctx->comm_stream) AFB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k)
        if (!ctx->comm_ev[k]) AFB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->comm_ev[k], cudaEventDisableTiming));
    return 0;
}

int64_t total(const std::vector<int64_t>& v) { int64_t s = 0; for (int64_t x : v) s += x; return s; }

}  // namespace

namespace afb {
void comm_release(afb_ctx* ctx) {
    if (ctx->comm && ctx->own_comm) { if (NcclApi* a = nccl()) a->CommDestroy(static_cast<ncclComm_t>(ctx->comm)); }
    ctx->comm = nullptr; ctx->own_comm = false;
    if (ctx->comm_stream) { cudaStreamDestroy(ctx->comm_stream); ctx->comm_stream = nullptr; }
    for (int k = 0; k < 2; ++k) if (ctx->comm_ev[k]) { cudaEventDestroy(ctx->comm_ev[k]); ctx->comm_ev[k] = nullptr; }
    ctx->halo_val_slots.release(); ctx->halo_rhs_slots.release(); ctx->halo_val_recv.release(); ctx->halo_rhs_recv.release();
    ctx->has_halo_plan = false;
}
}  // namespace afb

extern "C" {

int afb_comm_unique_id(void* id128) {
    if (!id128) return -7;
    NcclApi* a = nccl();
    if (!a) { set_error(nullptr, "NCCL not available (libnccl.so.2)"); return -8; }
    ncclUniqueId id;
    if (a->GetUniqueId(&id) != 0) { set_error(nullptr, "ncclGetUniqueId failed"); return -8; }
    std::memcpy(id128, &id, sizeof(id));
    return 0;
}

int afb_comm_init(afb_ctx* ctx, const void* id128, int rank, int nranks) {
    if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) { if (ctx) set_error(ctx, "afb_comm_init: bad arguments"); return -7; }
    NcclApi* a = nccl();
    if (!a) { set_error(ctx, "NCCL not available (libnccl.so.2)"); return -8; }
    if (cudaSetDevice(ctx->device) != cudaSuccess) { set_error(ctx, "no CUDA device (the library has no CPU fallback)"); return -4; }
    if (ctx->comm && ctx->own_comm) a->CommDestroy(static_cast<ncclComm_t>(ctx->comm));
    ctx->comm = nullptr;
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    ncclComm_t c = nullptr;
    AFB_NCCL(ctx, a->CommInitRank(&c, nranks, id, rank));
    ctx->comm = c; ctx->own_comm = true; ctx->comm_rank = rank; ctx->comm_size = nranks;
    return ensure_streams(ctx);
}

int afb_comm_set(afb_ctx* ctx, void* nccl_comm, int rank, int nranks) {
    if (!ctx || !nccl_comm || nranks < 1 || rank < 0 || rank >= nranks) { if (ctx) set_error(ctx, "afb_comm_set: bad arguments"); return -7; }
    if (!nccl()) { set_error(ctx, "NCCL not available (libnccl.so.2)"); return -8; }
    if (cudaSetDevice(ctx->device) != cudaSuccess) { set_error(ctx, "no CUDA device (the library has no CPU fallback)"); return -4; }
    if (ctx->comm && ctx->own_comm) nccl()->CommDestroy(static_cast<ncclComm_t>(ctx->comm));
    ctx->comm = nccl_comm; ctx->own_comm = false; ctx->comm_rank = rank; ctx->comm_size = nranks;
    return ensure_streams(ctx);
}

int afb_halo_plan_set(afb_ctx* ctx, int nranks, int64_t n_own, int64_t nnz_own, const int64_t* send_val, const int64_t* send_rhs,
                      const int64_t* recv_val, const int64_t* recv_rhs, const int64_t* val_slots, const int64_t* rhs_slots, int mem_space) {
    if (!ctx) return -7;
    if (nranks < 1 || n_own < 0 || nnz_own < 0 || !send_val || !send_rhs || !recv_val || !recv_rhs) { set_error(ctx, "afb_halo_plan_set: bad arguments"); return -7; }
    if (ctx->comm && nranks != ctx->comm_size) { set_error(ctx, "afb_halo_plan_set: nranks differs from the communicator"); return -7; }
    if (cudaSetDevice(ctx->device) != cudaSuccess) { set_error(ctx, "no CUDA device (the library has no CPU fallback)"); return -4; }
    ctx->has_halo_plan = false;
    ctx->halo_send_val.assign(send_val, send_val + nranks); ctx->halo_send_rhs.assign(send_rhs, send_rhs + nranks);
    ctx->halo_recv_val.assign(recv_val, recv_val + nranks); ctx->halo_recv_rhs.assign(recv_rhs, recv_rhs + nranks);
    for (int p = 0; p < nranks; ++p)
        if (send_val[p] < 0 || send_rhs[p] < 0 || recv_val[p] < 0 || recv_rhs[p] < 0) { set_error(ctx, "afb_halo_plan_set: negative count"); return -7; }
    const int64_t nv = total(ctx->halo_recv_val), nr = total(ctx->halo_recv_rhs);
    if ((nv > 0 && !val_slots) || (nr > 0 && !rhs_slots)) { set_error(ctx, "afb_halo_plan_set: slot arrays missing"); return -7; }
    const cudaMemcpyKind kind = mem_space == AFB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    AFB_CUDA(ctx, ctx->halo_val_slots.reserve(std::max<int64_t>(1, nv) * sizeof(int64_t)));
    AFB_CUDA(ctx, ctx->halo_rhs_slots.reserve(std::max<int64_t>(1, nr) * sizeof(int64_t)));
    AFB_CUDA(ctx, ctx->halo_val_recv.reserve(std::max<int64_t>(1, nv) * sizeof(double)));
    AFB_CUDA(ctx, ctx->halo_rhs_recv.reserve(std::max<int64_t>(1, nr) * sizeof(double)));
    if (nv) AFB_CUDA(ctx, cudaMemcpyAsync(ctx->halo_val_slots.p, val_slots, nv * sizeof(int64_t), kind, ctx->stream));
    if (nr) AFB_CUDA(ctx, cudaMemcpyAsync(ctx->halo_rhs_slots.p, rhs_slots, nr * sizeof(int64_t), kind, ctx->stream));
    AFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->halo_n_own = n_own; ctx->halo_nnz_own = nnz_own;
    ctx->comm_size = nranks;
    ctx->has_halo_plan = true;
    return 0;
}

int afb_halo_exchange_start(afb_ctx* ctx, double* val_ext, double* rhs_ext) {
    if (!ctx) return -7;
    if (!ctx->has_halo_plan) { set_error(ctx, "afb_halo_exchange: no halo plan (afb_halo_plan_set)"); return -6; }
    if (ctx->comm_size > 1 && !ctx->comm) { set_error(ctx, "afb_halo_exchange: no communicator (afb_comm_init / afb_comm_set)"); return -6; }
    if (ctx->halo_in_flight) { set_error(ctx, "afb_halo_exchange_start: an exchange is already in flight"); return -6; }
    if (cudaSetDevice(ctx->device) != cudaSuccess) { set_error(ctx, "no CUDA device (the library has no CPU fallback)"); return -4; }
    ctx->halo_in_flight = true;
    if (ctx->comm_size <= 1) return 0;
    NcclApi* a = nccl();
    ncclComm_t comm = static_cast<ncclComm_t>(ctx->comm);
    int rc = ensure_streams(ctx);
    if (rc) return rc;
    // the send buffers are produced on the context stream
    AFB_CUDA(ctx, cudaEventRecord(ctx->comm_ev[0], ctx->stream));
    AFB_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->comm_ev[0], 0));
    AFB_NCCL(ctx, a->GroupStart());
    int64_t so = 0, ro = 0, sro = 0, rro = 0;
    for (int p = 0; p < ctx->comm_size; ++p) {
        const int64_t sv = ctx->halo_send_val[p], rv = ctx->halo_recv_val[p], sr = ctx->halo_send_rhs[p], rr = ctx->halo_recv_rhs[p];
        if (p != ctx->comm_rank) {
            if (val_ext && sv) AFB_NCCL(ctx, a->Send(val_ext + ctx->halo_nnz_own + so, (size_t)sv, kNcclFloat64, p, comm, ctx->comm_stream));
            if (val_ext && rv) AFB_NCCL(ctx, a->Recv(ctx->halo_val_recv.as<double>() + ro, (size_t)rv, kNcclFloat64, p, comm, ctx->comm_stream));
            if (rhs_ext && sr) AFB_NCCL(ctx, a->Send(rhs_ext + ctx->halo_n_own + sro, (size_t)sr, kNcclFloat64, p, comm, ctx->comm_stream));
            if (rhs_ext && rr) AFB_NCCL(ctx, a->Recv(ctx->halo_rhs_recv.as<double>() + rro, (size_t)rr, kNcclFloat64, p, comm, ctx->comm_stream));
        }
        so += sv; ro += rv; sro += sr; rro += rr;
    }
    AFB_NCCL(ctx, a->GroupEnd());
    AFB_CUDA(ctx, cudaEventRecord(ctx->comm_ev[1], ctx->comm_stream));
    return 0;
}

int afb_halo_exchange_finish(afb_ctx* ctx, double* val_ext, double* rhs_ext) {
    if (!ctx) return -7;
    if (!ctx->halo_in_flight) { set_error(ctx, "afb_halo_exchange_finish: no exchange in flight"); return -6; }
    ctx->halo_in_flight = false;
    if (ctx->comm_size <= 1) return 0;
    if (cudaSetDevice(ctx->device) != cudaSuccess) { set_error(ctx, "no CUDA device (the library has no CPU fallback)"); return -4; }
    AFB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->comm_ev[1], 0));
    int64_t ro = 0, rro = 0;
    for (int p = 0; p < ctx->comm_size; ++p) {   // rank order: the sum of every interface entry has a fixed order
        const int64_t rv = ctx->halo_recv_val[p], rr = ctx->halo_recv_rhs[p];
        if (p != ctx->comm_rank) {
            if (val_ext && rv) {
                k_halo_add_peer<<<grid_for(rv), 256, 0, ctx->stream>>>(rv, ctx->halo_val_slots.as<long long>() + ro, ctx->halo_val_recv.as<double>() + ro, val_ext);
                ctx->launches++;
            }
            if (rhs_ext && rr) {
                k_halo_add_peer<<<grid_for(rr), 256, 0, ctx->stream>>>(rr, ctx->halo_rhs_slots.as<long long>() + rro, ctx->halo_rhs_recv.as<double>() + rro, rhs_ext);
                ctx->launches++;
            }
        }
        ro += rv; rro += rr;
    }
    AFB_CUDA(ctx, cudaGetLastError());
    return 0;
}

int afb_halo_exchange(afb_ctx* ctx, double* val_ext, double* rhs_ext) {
    const int rc = afb_halo_exchange_start(ctx, val_ext, rhs_ext);
    return rc ? rc : afb_halo_exchange_finish(ctx, val_ext, rhs_ext);
}

// One distributed assembly = what parallel.DistributedAssembler._assemble sequences: interface rows first (phase 1), their
// exchange on the communication stream while the remaining clusters run (phase 2), then the additions in rank order.  Problems
// the phased cluster gather does not cover run in one piece followed by the exchange.  Returns 0 / -1 like afb_assemble;
// the results are complete on the context's stream (afb_sync, or stream order for device consumers).
int afb_assemble_distributed(afb_ctx* ctx, int nforms, const afb_form* forms, int nrhs_forms, const afb_form* rhs_forms, double* val_ext,
                             double* rhs_ext, double drop_val) {
    if (!ctx) return -7;
    if (!ctx->has_halo_plan) { set_error(ctx, "afb_assemble_distributed: no halo plan (afb_halo_plan_set)"); return -6; }
    int st = 0, rc = 0;
    const bool phased = ctx->comm_size > 1 && ctx->priority_row >= 0;
    if (phased) {
        rc = afb_assemble_phase(ctx, nforms, forms, nrhs_forms, rhs_forms, val_ext, rhs_ext, drop_val, 1);
        if (rc < -1) return rc;
        rc = afb_halo_exchange_start(ctx, val_ext, rhs_ext);
        if (rc) return rc;
        st = afb_assemble_phase(ctx, nforms, forms, nrhs_forms, rhs_forms, val_ext, rhs_ext, drop_val, 2);
        if (st < -1) { ctx->halo_in_flight = false; return st; }
    } else {
        st = afb_assemble(ctx, nforms, forms, nrhs_forms, rhs_forms, val_ext, rhs_ext, 0, drop_val, AFB_DEVICE);
        if (st < -1) return st;
        rc = afb_halo_exchange_start(ctx, val_ext, rhs_ext);
        if (rc) return rc;
    }
    rc = afb_halo_exchange_finish(ctx, val_ext, rhs_ext);
    return rc ? rc : st;
}

}  // extern "C"
