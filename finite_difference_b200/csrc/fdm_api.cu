// fdm_api.cu -- the C ABI of libfdm_b200 (include/fdm_b200.h): validation, the host-buffer entry points with their
// device workspace, the dispatch to the per-class kernel launchers (fdm_internal.h), the book fan-out / epilogue
// kernels and the FP64 micro-benchmark.
//
// No CPU fallback exists: every entry point fails with FDM_ERR_NO_DEVICE / FDM_ERR_CUDA if the GPU path cannot run.

#include <mutex>

#include "fdm_common.cuh"
#include "fdm_lean_classes.cuh"
#include "fdm_book.cuh"

namespace {

thread_local char g_err[512] = "";

}  // namespace

int fdm_fail(int code, const char *fmt, const char *detail)
{
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}


namespace {

inline int fail(int code, const char *fmt, const char *detail = "") { return fdm_fail(code, fmt, detail); }

// Device workspaces of the host-buffer entry points: grow-only, ONE PER DEVICE, reused across calls (cudaMalloc /
// cudaFree cost milliseconds and synchronise the device).  Each has its own lock, held for the whole (synchronous)
// call: host threads driving different devices run concurrently, threads sharing a device serialise.
constexpr int MAX_DEVICES = 64;
struct Arena {
    std::mutex mu;
    char *p = nullptr;
    size_t bytes = 0;
    // the book host entries: a pinned mirror of the arena's input / output prefix (one H2D and one D2H per call instead of
    // one staged copy from pageable memory per array) and a stream + event pair kept across calls
    char *hp = nullptr;
    size_t hbytes = 0;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};
Arena g_arenas[MAX_DEVICES];

// caller holds a.mu and has `dev` current
cudaError_t arena_get(Arena &a, size_t bytes, char **out)
{
    if (a.p && bytes > a.bytes) {
        cudaFree(a.p);
        a.p = nullptr;
        a.bytes = 0;
    }
    if (!a.p) {
        const cudaError_t e = cudaMalloc((void **)&a.p, bytes);
        if (e != cudaSuccess) { a.p = nullptr; return e; }
        a.bytes = bytes;
    }
    *out = a.p;
    return cudaSuccess;
}

inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

// caller holds a.mu and has the arena's device current
cudaError_t arena_host(Arena &a, size_t bytes, char **out)
{
    if (a.hp && bytes > a.hbytes) {
        cudaFreeHost(a.hp);
        a.hp = nullptr;
        a.hbytes = 0;
    }
    if (!a.hp) {
        const cudaError_t e = cudaHostAlloc((void **)&a.hp, bytes, cudaHostAllocDefault);
        if (e != cudaSuccess) { a.hp = nullptr; return e; }
        a.hbytes = bytes;
    }
    *out = a.hp;
    return cudaSuccess;
}

cudaError_t arena_stream(Arena &a)
{
    cudaError_t e = cudaSuccess;
    if (!a.st) e = cudaStreamCreateWithFlags(&a.st, cudaStreamNonBlocking);
    if (e == cudaSuccess && !a.e0) e = cudaEventCreate(&a.e0);
    if (e == cudaSuccess && !a.e1) e = cudaEventCreate(&a.e1);
    return e;
}

// drains the arena's stream before the arena lock is released, on every exit path
struct StreamDrain {
    cudaStream_t st;
    ~StreamDrain() { if (st) cudaStreamSynchronize(st); }
};

// stream + two events of one host-entry call; on every exit path the stream is drained before it is destroyed, so
// nothing queued on it can still touch the arena when the arena's lock is released
struct CallCtx {
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~CallCtx()
    {
        if (st) cudaStreamSynchronize(st);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        if (st) cudaStreamDestroy(st);
    }
};

// ---------------------------------------------------------------------------------------------
// FP64 pipe micro-benchmark
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) fp64_peak_kernel(double *sink, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) sink[blockIdx.x] = s;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

// the register-lean kernels of the size classes present (fdm_lean.cuh); solves of other classes exit at once
int launch_lean_classes(const FdmBatch &b, const FdmOut &o, cudaStream_t st, bool am, bool lin, bool dv, int class_mask,
                        int32_t *n_launches)
{
    typedef int (*Fn)(const FdmBatch &, const FdmOut &, cudaStream_t, bool, bool, bool, int32_t *);
    constexpr lean::ClassList cl = lean::class_list();
    for (int i = 0; i < cl.n; ++i) {
        if (!(class_mask & (1 << i))) continue;
        Fn fn = nullptr;
        switch (cl.c[i]) {
#ifdef FDM_LEAN_ALL_CLASSES
        case 4: fn = fdm_launch_lean_4; break;
        case 7: fn = fdm_launch_lean_7; break;
        case 8: fn = fdm_launch_lean_8; break;
#endif
        case 10: fn = fdm_launch_lean_10; break;
        case 13: fn = fdm_launch_lean_13; break;
        case 16: fn = fdm_launch_lean_16; break;
        default: break;
        }
        if (!fn) continue;
        const int rc = fn(b, o, st, am, lin, dv, n_launches);
        if (rc != FDM_OK) return rc;
    }
    return FDM_OK;
}

// launch hints of the size classes present in a call (FdmBatch.reserved bits 20..29):
// bit 20+i: class i of lean::class_list() is present; bit 28: a warp-sized grid outside every class is present;
// bit 29: these hints are valid
int lean_hint_bits(int n_space)
{
    if (n_space > FDM_WARP_MAX_N_SPACE) return 0;
    constexpr lean::ClassList cl = lean::class_list();
    const int c = lean::class_of(n_space - 1);
    if (c == 0) return 1 << 28;
    for (int i = 0; i < cl.n; ++i) if (cl.c[i] == c) return 1 << (20 + i);
    return 1 << 28;
}
constexpr int LEAN_HINTS_VALID = 1 << 29;


int validate(const FdmBatch *b, const FdmOut *o)
{
    if (!b || !o) return fail(FDM_ERR_INVALID, "batch / out is NULL%s");
    if (b->n_solves < 0) return fail(FDM_ERR_INVALID, "n_solves < 0%s");
    if (b->n_solves == 0) return FDM_OK;
    if (!b->x_min || !b->dx || !b->sigma || !b->r || !b->b || !b->spot || !b->strike || !b->n_space ||
        !b->flags || !b->rannacher || !b->seg_begin || !b->seg_steps || !b->seg_dt || !b->seg_tau0 ||
        !b->seg_restart || !b->seg_div)
        return fail(FDM_ERR_INVALID, "a required FdmBatch array is NULL%s");
    if (!o->price) return fail(FDM_ERR_INVALID, "FdmOut.price is NULL%s");
    if (b->n_space_max < 3) return fail(FDM_ERR_INVALID, "n_space_max < 3%s");
    if (b->n_space_max > FDM_MAX_N_SPACE)
        return fail(FDM_ERR_UNSUPPORTED, "n_space_max exceeds FDM_MAX_N_SPACE (8193)%s");
    if (b->n_mon_total > 0 && (!b->mon_begin || !b->mon_step))
        return fail(FDM_ERR_INVALID, "monitor tables missing%s");
    return FDM_OK;
}

// what only a host copy of the arrays can tell (fdm_solve_batch_host): every kernel skips a solve outside the size
// classes it was launched for, so a bad n_space would otherwise come back as an untouched result
int validate_host_arrays(const FdmBatch *b)
{
    const int64_t n = b->n_solves;
    for (int64_t i = 0; i < n; ++i) {
        if (b->n_space[i] < 3 || b->n_space[i] > b->n_space_max)
            return fail(FDM_ERR_INVALID, "n_space[i] outside [3, n_space_max]%s");
        if (b->seg_begin[i + 1] <= b->seg_begin[i]) return fail(FDM_ERR_INVALID, "a solve has no time segment%s");
    }
    if (b->seg_begin[0] != 0 || b->seg_begin[n] > b->n_seg_total)
        return fail(FDM_ERR_INVALID, "seg_begin does not match n_seg_total%s");
    const int64_t nseg = b->seg_begin[n];
    for (int64_t j = 0; j < nseg; ++j) {
        if (b->seg_steps[j] < 1) return fail(FDM_ERR_INVALID, "seg_steps < 1%s");
        if (!(b->seg_dt[j] > 0.0)) return fail(FDM_ERR_INVALID, "seg_dt must be positive%s");
    }
    return FDM_OK;
}


// flags of the whole call (FdmBatch.reserved): bit0 any american, bit1 any linear, bit2 any dividend,
// bit3 any solve with n_space <= FDM_WARP_MAX_N_SPACE, bit4 any larger solve,
// bits 8..19 the largest n_space of the small class (0 = unknown)
int launch(const FdmBatch &b, const FdmOut &o, cudaStream_t st, int family_mask, int32_t *n_launches)
{
    const bool am = family_mask & 1, lin = family_mask & 2, dv = family_mask & 4;
    bool small = family_mask & 8, large = family_mask & 16;
    if (!small && !large) {
        large = b.n_space_max > FDM_WARP_MAX_N_SPACE;
        small = true;                                   // unknown mix: the warp kernel skips what it cannot hold
    }
    if (b.n_space_max <= FDM_WARP_MAX_N_SPACE) large = false;
    int rc = FDM_OK;
    if (small) {
        int small_max = (family_mask >> 8) & 0xfff;
        if (small_max == 0 || small_max > FDM_WARP_MAX_N_SPACE)
            small_max = b.n_space_max < FDM_WARP_MAX_N_SPACE ? b.n_space_max : FDM_WARP_MAX_N_SPACE;
        const int rows = small_max - 1;
        // register-lean kernels first (one launch per size class present), then fdm_warp_kernel for the rest:
        // sizes outside the classes and the solves the lean kernel handed back
        static const bool no_lean = getenv("FDM_B200_NO_LEAN") != nullptr;
        int only_rest = 0;
        if (!no_lean) {
            const bool hinted = family_mask & LEAN_HINTS_VALID;
            const int class_mask = hinted ? (family_mask >> 20) & 0xff : 0xff;
            rc = launch_lean_classes(b, o, st, am, lin, dv, class_mask, n_launches);
            if (rc != FDM_OK) return rc;
            only_rest = 1;
        }
        if (rows <= 32 * 4) rc = fdm_launch_warp_4(b, o, st, am, lin, dv, only_rest, n_launches);
        else if (rows <= 32 * 8) rc = fdm_launch_warp_8(b, o, st, am, lin, dv, only_rest, n_launches);
        else if (rows <= 32 * 13) rc = fdm_launch_warp_13(b, o, st, am, lin, dv, only_rest, n_launches);
        else rc = fdm_launch_warp_16(b, o, st, am, lin, dv, only_rest, n_launches);
        if (rc != FDM_OK) return rc;
    }
    if (large) {
        // one launch per CTA size class that can be populated;
        // size classes in ascending capacity (32 W C interior rows); a solve runs in the smallest class that holds it
        // (n_lo = largest n_space of the class below), other classes' CTAs exit at once
        const int rows_max = b.n_space_max - 1;
        int n_lo = FDM_WARP_MAX_N_SPACE;
#define FDM_RUN_CTA(W, C)                                                                                          \
        if (rc == FDM_OK && rows_max + 1 > n_lo) rc = fdm_launch_cta_##W##x##C(b, o, st, am, lin, dv, n_lo, n_launches); \
        n_lo = 32 * W * C + 1;
        FDM_CTA_CLASSES(FDM_RUN_CTA)
#undef FDM_RUN_CTA
        // what the CTA kernels flagged FDM_STATUS_STIFF is solved again by fdm_exact_kernel (one CTA screens 256 solves)
        if (rc == FDM_OK) rc = fdm_launch_cta_exact(b, o, st, n_launches);
    }
    return rc;
}

}  // namespace
namespace {

// ---------------------------------------------------------------------------------------------
// books: workspace layouts and the queued fan-out -> solves -> epilogue sequence
// ---------------------------------------------------------------------------------------------

struct SolveWs {                 // expanded descriptors and per-solve results of a book, offsets into the workspace
    size_t d[7], ns, fl, rn, sb, st, dt, t0, rs, dv;          // FdmBatch arrays (x_min dx sigma r b spot strike ...)
    size_t bar, reb, bup, mb, me;                             // barrier books only
    size_t p, dl, g, idx, status;                             // FdmOut arrays
    size_t total;
};

SolveWs solve_ws(int64_t m, int64_t nseg, bool barrier)
{
    SolveWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += up256(bytes); return o; };
    for (int i = 0; i < 7; ++i) w.d[i] = take(m * 8);
    w.ns = take(m * 4); w.fl = take(m * 4); w.rn = take(m * 4); w.sb = take((m + 1) * 4);
    w.st = take(nseg * 4); w.dt = take(nseg * 8); w.t0 = take(nseg * 8); w.rs = take(nseg * 4); w.dv = take(nseg * 8);
    w.bar = take(barrier ? m * 8 : 0); w.reb = take(barrier ? m * 8 : 0); w.bup = take(barrier ? m * 8 : 0);
    w.mb = take(barrier ? m * 4 : 0); w.me = take(barrier ? m * 4 : 0);
    w.p = take(m * 8); w.dl = take(m * 8); w.g = take(m * 8); w.idx = take(m * 16); w.status = take(m * 4);
    w.total = off;
    return w;
}

void bind_solve_ws(const SolveWs &w, char *ws, int64_t m, int64_t nseg, FdmBatch &sb, FdmOut &so)
{
    memset(&sb, 0, sizeof(sb));
    memset(&so, 0, sizeof(so));
    sb.n_solves = m; sb.n_seg_total = (int32_t)nseg;
    sb.x_min = (const double *)(ws + w.d[0]); sb.dx = (const double *)(ws + w.d[1]);
    sb.sigma = (const double *)(ws + w.d[2]); sb.r = (const double *)(ws + w.d[3]);
    sb.b = (const double *)(ws + w.d[4]); sb.spot = (const double *)(ws + w.d[5]);
    sb.strike = (const double *)(ws + w.d[6]);
    sb.n_space = (const int32_t *)(ws + w.ns); sb.flags = (const int32_t *)(ws + w.fl);
    sb.rannacher = (const int32_t *)(ws + w.rn); sb.seg_begin = (const int32_t *)(ws + w.sb);
    sb.seg_steps = (const int32_t *)(ws + w.st); sb.seg_dt = (const double *)(ws + w.dt);
    sb.seg_tau0 = (const double *)(ws + w.t0); sb.seg_restart = (const int32_t *)(ws + w.rs);
    sb.seg_div = (const double *)(ws + w.dv);
    so.price = (double *)(ws + w.p); so.delta = (double *)(ws + w.dl); so.gamma = (double *)(ws + w.g);
    so.idx = (int32_t *)(ws + w.idx); so.status = (int32_t *)(ws + w.status);
}

// queue fan-out, solves and epilogue of an American book on `st`; db / dbo hold device pointers
int american_book_queue(const FdmAmericanBook &db, const FdmBookOut &dbo, char *ws, cudaStream_t st, int mask,
                        int n_space_max, const book::RoleTable &roles, int32_t *n_launches)
{
    const int64_t n = db.n_options, m = n * roles.n;
    const int64_t ndiv = db.div_begin ? db.n_div_total : 0;
    const int64_t nseg = (n + ndiv) * roles.n;
    if (m + 1 > 0x7fffffff || nseg > 0x7fffffff) return fail(FDM_ERR_UNSUPPORTED, "book too large for 32-bit CSR tables%s");
    const SolveWs w = solve_ws(m, nseg, false);
    FdmBatch sb;
    FdmOut so;
    bind_solve_ws(w, ws, m, nseg, sb, so);
    sb.n_space_max = n_space_max;
    sb.reserved = mask;
    FDM_CUDA(cudaMemsetAsync(ws + w.status, 0xff, m * 4, st));            // FDM_STATUS_NOT_SOLVED until a kernel takes it
    const int threads = 256;
    book::expand_kernel<<<(unsigned)((m + threads - 1) / threads), threads, 0, st>>>(db, sb, roles);
    const int rc = launch(sb, so, st, mask, n_launches);
    if (rc != FDM_OK) return rc;
    book::combine_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, st>>>(db, so, dbo, roles.n);
    FDM_CUDA(cudaGetLastError());
    if (n_launches) *n_launches += 2;
    return FDM_OK;
}

int barrier_roles(int what) { return what == FDM_BARRIER_BOOK_GREEKS ? 3 : 1; }

int barrier_book_queue(const FdmBarrierBook &db, const FdmBookOut &dbo, char *ws, cudaStream_t st, int mask,
                       int n_space_max, int32_t *n_launches)
{
    const int n_roles = barrier_roles(db.what);
    const int64_t n = db.n_options, m = n * n_roles;
    if (m + 1 > 0x7fffffff) return fail(FDM_ERR_UNSUPPORTED, "book too large for 32-bit CSR tables%s");
    const SolveWs w = solve_ws(m, m, true);
    FdmBatch sb;
    FdmOut so;
    bind_solve_ws(w, ws, m, m, sb, so);
    sb.n_space_max = n_space_max;
    sb.reserved = mask;
    sb.barrier = (const double *)(ws + w.bar);
    sb.rebate = (const double *)(ws + w.reb);
    sb.barrier_up = db.barrier_up ? (const double *)(ws + w.bup) : nullptr;
    if (db.mon_begin && db.mon_step && db.n_mon_total > 0) {
        sb.n_mon_total = (int32_t)db.n_mon_total;
        sb.mon_begin = (const int32_t *)(ws + w.mb);
        sb.mon_end = (const int32_t *)(ws + w.me);
        sb.mon_step = db.mon_step;                                         // shared by the option's solves
    }
    FDM_CUDA(cudaMemsetAsync(ws + w.status, 0xff, m * 4, st));
    const int threads = 256;
    book::barrier_expand_kernel<<<(unsigned)((m + threads - 1) / threads), threads, 0, st>>>(db, sb, n_roles);
    const int rc = launch(sb, so, st, mask, n_launches);
    if (rc != FDM_OK) return rc;
    book::barrier_combine_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, st>>>(db, so, dbo, n_roles);
    FDM_CUDA(cudaGetLastError());
    if (n_launches) *n_launches += 2;
    return FDM_OK;
}

int no_device()
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(FDM_ERR_NO_DEVICE, "no CUDA device; libfdm_b200 has no CPU fallback%s");
    return FDM_OK;
}

#define FDM_TRY(call)                                                                      \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess && rc == FDM_OK)                                             \
            rc = fail(FDM_ERR_CUDA, #call ": %s", cudaGetErrorString(e_));                 \
    } while (0)

}  // namespace

extern "C" {

int fdm_version(void) { return FDM_ABI_VERSION; }

int fdm_sizeof(int which)
{
    switch (which) {
    case 0: return (int)sizeof(FdmBatch);
    case 1: return (int)sizeof(FdmOut);
    case 2: return (int)sizeof(FdmAmericanBook);
    case 3: return (int)sizeof(FdmBookOut);
    case 4: return (int)sizeof(FdmBarrierBook);
    case 5: return (int)sizeof(FdmRowVarBatch);
    case 6: return (int)sizeof(FdmRowVarOut);
    default: return -1;
    }
}

const char *fdm_last_error(void) { return g_err; }

int fdm_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(FDM_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return n;
}

// Device-pointer entry.  The family mask cannot be derived without reading device memory, so
// `reserved` carries it: bit0 any American solve, bit1 any linear (European / barrier) solve,
// bit2 any dividend jump.  0 means "unknown": launch both families with dividend scratch.
int fdm_solve_batch(const FdmBatch *batch, const FdmOut *out, void *stream, int32_t *n_launches)
{
    g_err[0] = 0;
    if (n_launches) *n_launches = 0;
    int rc = validate(batch, out);
    if (rc != FDM_OK || batch->n_solves == 0) return rc;
    if ((rc = no_device()) != FDM_OK) return rc;
    int mask = batch->reserved;
    if ((mask & 3) == 0) mask |= 7;
    // a solve that no launched kernel takes (n_space beyond n_space_max, hints that miss its class) stays visible
    if (out->status) FDM_CUDA(cudaMemsetAsync(out->status, 0xff, batch->n_solves * 4, (cudaStream_t)stream));
    return launch(*batch, *out, (cudaStream_t)stream, mask, n_launches);
}

int fdm_solve_batch_host(const FdmBatch *hb, const FdmOut *ho, int32_t *n_launches, float *kernel_ms)
{
    g_err[0] = 0;
    if (n_launches) *n_launches = 0;
    if (kernel_ms) *kernel_ms = 0.f;
    int rc = validate(hb, ho);
    if (rc != FDM_OK || hb->n_solves == 0) return rc;
    if ((rc = validate_host_arrays(hb)) != FDM_OK) return rc;
    if ((rc = no_device()) != FDM_OK) return rc;

    const int64_t n = hb->n_solves;
    const int64_t nseg = hb->seg_begin[n];
    const int64_t nmon = (hb->mon_begin && hb->n_mon_total > 0) ? hb->mon_begin[n] : 0;
    const bool has_ko = hb->barrier && hb->rebate;

    const int mask = fdm_launch_hints(hb->n_space, hb->flags, n, hb->seg_div, nseg);

    // one device arena for descriptors and results
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += up256(bytes); return o; };
    const size_t o_d[9] = {take(n * 8), take(n * 8), take(n * 8), take(n * 8), take(n * 8), take(n * 8),
                           take(n * 8), take(has_ko ? n * 8 : 0), take(has_ko ? n * 8 : 0)};
    const size_t o_nspace = take(n * 4), o_flags = take(n * 4), o_ran = take(n * 4);
    const size_t o_segb = take((n + 1) * 4), o_monb = take(nmon ? (n + 1) * 4 : 0);
    const size_t o_mone = take((nmon && hb->mon_end) ? n * 4 : 0);
    const size_t o_ssteps = take(nseg * 4), o_sdt = take(nseg * 8), o_stau = take(nseg * 8);
    const size_t o_srst = take(nseg * 4), o_sdiv = take(nseg * 8), o_mon = take(nmon * 4);
    const size_t o_bb = take(hb->b_bdry ? n * 8 : 0), o_bup = take(hb->barrier_up ? n * 8 : 0);
    const size_t o_mreb = take((hb->mon_rebate && nmon) ? nmon * 8 : 0);
    const size_t r_price = take(n * 8), r_delta = take(n * 8), r_gamma = take(n * 8);
    const size_t r_sv = take(n * 32), r_ss = take(n * 32), r_idx = take(n * 16), r_status = take(n * 4);
    const size_t r_dump = take(ho->v_dump ? (size_t)n * ho->v_dump_stride * 8 : 0);
    const size_t r_trace = take(ho->ex_trace ? (size_t)n * ho->ex_trace_stride * 4 : 0);

    int dev = 0;
    FDM_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEVICES) return fail(FDM_ERR_UNSUPPORTED, "device index beyond the workspace table%s");
    Arena &ar = g_arenas[dev];
    std::lock_guard<std::mutex> arena_lock(ar.mu);
    char *arena = nullptr;
    FDM_CUDA(arena_get(ar, off, &arena));
    CallCtx cx;                                      // destroyed (stream drained) before the lock is released
    rc = FDM_OK;
    FDM_TRY(cudaStreamCreateWithFlags(&cx.st, cudaStreamNonBlocking));
    FDM_TRY(cudaEventCreate(&cx.e0));
    FDM_TRY(cudaEventCreate(&cx.e1));
    if (rc != FDM_OK) return rc;
    cudaStream_t st = cx.st;
    auto h2d = [&](size_t o, const void *src, size_t bytes) {
        if (bytes && src) FDM_TRY(cudaMemcpyAsync(arena + o, src, bytes, cudaMemcpyHostToDevice, st));
    };
    const double *hd[9] = {hb->x_min, hb->dx, hb->sigma, hb->r, hb->b, hb->spot, hb->strike,
                           has_ko ? hb->barrier : nullptr, has_ko ? hb->rebate : nullptr};
    for (int i = 0; i < 9; ++i) h2d(o_d[i], hd[i], n * 8);
    h2d(o_nspace, hb->n_space, n * 4);
    h2d(o_flags, hb->flags, n * 4);
    h2d(o_ran, hb->rannacher, n * 4);
    h2d(o_segb, hb->seg_begin, (n + 1) * 4);
    if (nmon) h2d(o_monb, hb->mon_begin, (n + 1) * 4);
    if (nmon && hb->mon_end) h2d(o_mone, hb->mon_end, n * 4);
    h2d(o_ssteps, hb->seg_steps, nseg * 4);
    h2d(o_sdt, hb->seg_dt, nseg * 8);
    h2d(o_stau, hb->seg_tau0, nseg * 8);
    h2d(o_srst, hb->seg_restart, nseg * 4);
    h2d(o_sdiv, hb->seg_div, nseg * 8);
    if (nmon) h2d(o_mon, hb->mon_step, nmon * 4);
    if (hb->b_bdry) h2d(o_bb, hb->b_bdry, n * 8);
    if (hb->barrier_up) h2d(o_bup, hb->barrier_up, n * 8);
    if (hb->mon_rebate && nmon) h2d(o_mreb, hb->mon_rebate, nmon * 8);

    FdmBatch db = *hb;
    db.x_min = (const double *)(arena + o_d[0]);
    db.dx = (const double *)(arena + o_d[1]);
    db.sigma = (const double *)(arena + o_d[2]);
    db.r = (const double *)(arena + o_d[3]);
    db.b = (const double *)(arena + o_d[4]);
    db.spot = (const double *)(arena + o_d[5]);
    db.strike = (const double *)(arena + o_d[6]);
    db.barrier = has_ko ? (const double *)(arena + o_d[7]) : nullptr;
    db.rebate = has_ko ? (const double *)(arena + o_d[8]) : nullptr;
    db.n_space = (const int32_t *)(arena + o_nspace);
    db.flags = (const int32_t *)(arena + o_flags);
    db.rannacher = (const int32_t *)(arena + o_ran);
    db.seg_begin = (const int32_t *)(arena + o_segb);
    db.mon_begin = nmon ? (const int32_t *)(arena + o_monb) : nullptr;
    db.mon_end = (nmon && hb->mon_end) ? (const int32_t *)(arena + o_mone) : nullptr;
    db.seg_steps = (const int32_t *)(arena + o_ssteps);
    db.seg_dt = (const double *)(arena + o_sdt);
    db.seg_tau0 = (const double *)(arena + o_stau);
    db.seg_restart = (const int32_t *)(arena + o_srst);
    db.seg_div = (const double *)(arena + o_sdiv);
    db.mon_step = nmon ? (const int32_t *)(arena + o_mon) : nullptr;
    db.b_bdry = hb->b_bdry ? (const double *)(arena + o_bb) : nullptr;
    db.barrier_up = hb->barrier_up ? (const double *)(arena + o_bup) : nullptr;
    db.mon_rebate = (hb->mon_rebate && nmon) ? (const double *)(arena + o_mreb) : nullptr;

    FdmOut dout = *ho;
    dout.price = (double *)(arena + r_price);
    dout.delta = (double *)(arena + r_delta);
    dout.gamma = (double *)(arena + r_gamma);
    dout.stencil_v = (double *)(arena + r_sv);
    dout.stencil_s = (double *)(arena + r_ss);
    dout.idx = (int32_t *)(arena + r_idx);
    dout.status = (int32_t *)(arena + r_status);
    dout.v_dump = ho->v_dump ? (double *)(arena + r_dump) : nullptr;
    dout.ex_trace = ho->ex_trace ? (int32_t *)(arena + r_trace) : nullptr;
    if (ho->ex_trace) FDM_TRY(cudaMemsetAsync(arena + r_trace, 0xff, (size_t)n * ho->ex_trace_stride * 4, st));
    if (ho->v_dump) FDM_TRY(cudaMemsetAsync(arena + r_dump, 0, (size_t)n * ho->v_dump_stride * 8, st));
    FDM_TRY(cudaMemsetAsync(arena + r_status, 0xff, n * 4, st));          // FDM_STATUS_NOT_SOLVED

    if (rc == FDM_OK) {
        FDM_TRY(cudaEventRecord(cx.e0, st));
        int lrc = launch(db, dout, st, mask, n_launches);
        if (lrc != FDM_OK) rc = lrc;
        FDM_TRY(cudaEventRecord(cx.e1, st));
    }
    auto d2h = [&](void *dst, size_t o, size_t bytes) {
        if (dst && bytes) FDM_TRY(cudaMemcpyAsync(dst, arena + o, bytes, cudaMemcpyDeviceToHost, st));
    };
    if (rc == FDM_OK) {
        d2h(ho->price, r_price, n * 8);
        d2h(ho->delta, r_delta, n * 8);
        d2h(ho->gamma, r_gamma, n * 8);
        d2h(ho->stencil_v, r_sv, n * 32);
        d2h(ho->stencil_s, r_ss, n * 32);
        d2h(ho->idx, r_idx, n * 16);
        d2h(ho->status, r_status, n * 4);
        if (ho->v_dump) d2h(ho->v_dump, r_dump, (size_t)n * ho->v_dump_stride * 8);
        if (ho->ex_trace) d2h(ho->ex_trace, r_trace, (size_t)n * ho->ex_trace_stride * 4);
        FDM_TRY(cudaStreamSynchronize(st));
        if (kernel_ms && rc == FDM_OK) FDM_TRY(cudaEventElapsedTime(kernel_ms, cx.e0, cx.e1));
    }
    return rc;
}

int fdm_book_solves_per_option(int32_t what, int32_t use_richardson)
{
    book::Role r[8];
    if (what < 0 || what > FDM_BOOK_SCENARIO) return fail(FDM_ERR_INVALID, "unknown book fan-out%s");
    return book::roles_of(what, use_richardson != 0, r);
}

size_t fdm_american_book_workspace(int64_t n_options, int64_t n_div_total, int32_t what, int32_t use_richardson)
{
    book::Role r[8];
    if (n_options <= 0 || what < 0 || what > FDM_BOOK_SCENARIO) return 0;
    const int nr = book::roles_of(what, use_richardson != 0, r);
    if (n_div_total < 0) n_div_total = 0;
    return solve_ws(n_options * nr, (n_options + n_div_total) * nr, false).total;
}

int fdm_american_book(const FdmAmericanBook *book_, const FdmBookOut *out, void *workspace, size_t workspace_bytes,
                      void *stream, int32_t *n_launches)
{
    g_err[0] = 0;
    if (n_launches) *n_launches = 0;
    if (!book_ || !out) return fail(FDM_ERR_INVALID, "book / out is NULL%s");
    const FdmAmericanBook &bk = *book_;
    if (bk.n_options < 0) return fail(FDM_ERR_INVALID, "n_options < 0%s");
    if (bk.what < 0 || bk.what > FDM_BOOK_SCENARIO) return fail(FDM_ERR_INVALID, "unknown book fan-out%s");
    if (bk.n_options == 0) return FDM_OK;
    if (!bk.spot || !bk.strike || !bk.T || !bk.sigma || !bk.r || !bk.b || !bk.flags || !bk.n_space || !bk.n_time ||
        !bk.rannacher)
        return fail(FDM_ERR_INVALID, "a required FdmAmericanBook array is NULL%s");
    if (bk.div_begin && bk.n_div_total > 0 && (!bk.div_tau || !bk.div_cash))
        return fail(FDM_ERR_INVALID, "dividend tables missing%s");
    if (bk.n_space_max < 3) return fail(FDM_ERR_INVALID, "n_space_max must be set (>= 3) for the device entry%s");
    if (bk.n_space_max > FDM_MAX_N_SPACE) return fail(FDM_ERR_UNSUPPORTED, "n_space exceeds FDM_MAX_N_SPACE (8193)%s");
    const size_t need = fdm_american_book_workspace(bk.n_options, bk.div_begin ? bk.n_div_total : 0, bk.what, bk.use_richardson);
    if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 255))
        return fail(FDM_ERR_INVALID, "workspace missing, misaligned or smaller than fdm_american_book_workspace()%s");
    int rc = no_device();
    if (rc != FDM_OK) return rc;
    book::RoleTable roles;
    roles.n = book::roles_of(bk.what, bk.use_richardson != 0, roles.r);
    int mask = (bk.hints & ~6) | 1;                                  // American solves only
    if (bk.div_begin && bk.n_div_total > 0) mask |= 4;
    return american_book_queue(bk, *out, (char *)workspace, (cudaStream_t)stream, mask, bk.n_space_max, roles, n_launches);
}

int fdm_american_book_host(const FdmAmericanBook *hb, const FdmBookOut *ho, int64_t *node_steps,
                           int32_t *n_launches, float *kernel_ms)
{
    g_err[0] = 0;
    if (n_launches) *n_launches = 0;
    if (kernel_ms) *kernel_ms = 0.f;
    if (node_steps) *node_steps = 0;
    if (!hb || !ho) return fail(FDM_ERR_INVALID, "book / out is NULL%s");
    if (hb->n_options < 0) return fail(FDM_ERR_INVALID, "n_options < 0%s");
    if (hb->what < 0 || hb->what > FDM_BOOK_SCENARIO) return fail(FDM_ERR_INVALID, "unknown book fan-out%s");
    const int64_t n = hb->n_options;
    if (n == 0) return FDM_OK;
    if (!hb->spot || !hb->strike || !hb->T || !hb->sigma || !hb->r || !hb->b || !hb->flags || !hb->n_space ||
        !hb->n_time || !hb->rannacher)
        return fail(FDM_ERR_INVALID, "a required FdmAmericanBook array is NULL%s");
    const int64_t ndiv = hb->div_begin ? hb->div_begin[n] : 0;
    if (ndiv > 0 && (!hb->div_tau || !hb->div_cash)) return fail(FDM_ERR_INVALID, "dividend tables missing%s");
    book::RoleTable roles;
    roles.n = book::roles_of(hb->what, hb->use_richardson != 0, roles.r);
    // host-side validation (the reference raises ValueError, fd_american_option_pricer.py:68-69, :123-124)
    int n_space_max = 3, small_max = 0, mask = 1;
    int64_t ns_total = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (!(hb->spot[i] > 0.0) || !(hb->strike[i] > 0.0) || !(hb->T[i] > 0.0))
            return fail(FDM_ERR_INVALID, "spot, strike and time_to_expiry must be positive%s");
        const int ns = hb->n_space[i] < 3 ? 3 : hb->n_space[i];
        const int nt = hb->n_time[i] < 4 ? 4 : hb->n_time[i];
        if (ns > FDM_MAX_N_SPACE) return fail(FDM_ERR_UNSUPPORTED, "n_space exceeds FDM_MAX_N_SPACE (8193)%s");
        if (ns > n_space_max) n_space_max = ns;
        if (ns <= FDM_WARP_MAX_N_SPACE) {
            mask |= 8 | LEAN_HINTS_VALID | lean_hint_bits(ns);
            if (ns > small_max) small_max = ns;
        } else mask |= 16;
        if (ndiv > 0) {
            // dividends strictly inside (0, T), ascending, no two on the same tau: the reference's segment lengths
            // must be positive (fd_american_option_pricer.py:498 divides by dt)
            const int d0 = hb->div_begin[i], d1 = hb->div_begin[i + 1];
            if (d1 < d0) return fail(FDM_ERR_INVALID, "div_begin must be ascending%s");
            if (d1 > d0 && (hb->flags[i] & FDM_FLAG_BLACK76))
                return fail(FDM_ERR_INVALID, "Black-76 options take no dividends (fd_american_black76.py:89)%s");
            double prev = 0.0;
            for (int d = d0; d < d1; ++d) {
                if (!(hb->div_tau[d] > prev) || !(hb->div_tau[d] < hb->T[i]))
                    return fail(FDM_ERR_INVALID, "zero-length time segment (dividend taus must be strictly ascending inside (0, T))%s");
                prev = hb->div_tau[d];
            }
        }
        for (int k = 0; k < roles.n; ++k) {
            const double sg = hb->sigma[i] + roles.r[k].bump * hb->dv_sigma;
            if (!(sg > 0.0)) return fail(FDM_ERR_INVALID, "sigma (after bumps) must be positive%s");
            int steps = nt;
            if (roles.r[k].fine == 1) steps = (hb->flags[i] & FDM_FLAG_BLACK76) ? 2 * nt : 2 * ns;
            if (roles.r[k].fine == 2) steps = 2 * nt;
            ns_total += (int64_t)ns * steps;
        }
    }
    mask |= small_max << 8;
    if (ndiv > 0) mask |= 4;
    if (node_steps) *node_steps = ns_total;
    int rc = no_device();
    if (rc != FDM_OK) return rc;

    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += up256(bytes); return o; };
    // option inputs
    const size_t i_d[6] = {take(n * 8), take(n * 8), take(n * 8), take(n * 8), take(n * 8), take(n * 8)};
    const size_t i_fl = take(n * 4), i_ns = take(n * 4), i_nt = take(n * 4), i_rn = take(n * 4);
    const size_t i_db = take(ndiv ? (n + 1) * 4 : 0), i_dt = take(ndiv * 8), i_dc = take(ndiv * 8);
    // per-option outputs
    const size_t o_d[6] = {take(n * 8), take(n * 8), take(n * 8), take(n * 8), take(n * 8), take(n * 8)};
    const size_t o_lo = take(n * 4), o_hi = take(n * 4), o_st = take(n * 4);
    const size_t o_end = off;
    const size_t i_oi = take(ho->out_index ? n * 4 : 0);
    const size_t stage_end = off;                       // [0, stage_end) is mirrored in pinned host memory
    // expanded descriptors and per-solve results
    const size_t o_ws = take(fdm_american_book_workspace(n, ndiv, hb->what, hb->use_richardson));

    int dev = 0;
    FDM_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEVICES) return fail(FDM_ERR_UNSUPPORTED, "device index beyond the workspace table%s");
    Arena &ar = g_arenas[dev];
    std::lock_guard<std::mutex> arena_lock(ar.mu);
    char *arena = nullptr;
    FDM_CUDA(arena_get(ar, off, &arena));
    char *hp = nullptr;
    FDM_CUDA(arena_host(ar, stage_end, &hp));
    FDM_CUDA(arena_stream(ar));
    cudaStream_t st = ar.st;
    StreamDrain drain{st};
    struct { cudaEvent_t e0, e1; } cx = {ar.e0, ar.e1};
    auto put = [&](size_t o, const void *src, size_t bytes) { if (src && bytes) memcpy(hp + o, src, bytes); };
    const double *hd[6] = {hb->spot, hb->strike, hb->T, hb->sigma, hb->r, hb->b};
    for (int i = 0; i < 6; ++i) put(i_d[i], hd[i], n * 8);
    put(i_fl, hb->flags, n * 4); put(i_ns, hb->n_space, n * 4); put(i_nt, hb->n_time, n * 4); put(i_rn, hb->rannacher, n * 4);
    if (ndiv) { put(i_db, hb->div_begin, (n + 1) * 4); put(i_dt, hb->div_tau, ndiv * 8); put(i_dc, hb->div_cash, ndiv * 8); }
    FDM_TRY(cudaMemcpyAsync(arena, hp, o_d[0], cudaMemcpyHostToDevice, st));       // every input array in one copy

    FdmAmericanBook db = *hb;
    db.spot = (const double *)(arena + i_d[0]); db.strike = (const double *)(arena + i_d[1]);
    db.T = (const double *)(arena + i_d[2]); db.sigma = (const double *)(arena + i_d[3]);
    db.r = (const double *)(arena + i_d[4]); db.b = (const double *)(arena + i_d[5]);
    db.flags = (const int32_t *)(arena + i_fl); db.n_space = (const int32_t *)(arena + i_ns);
    db.n_time = (const int32_t *)(arena + i_nt); db.rannacher = (const int32_t *)(arena + i_rn);
    db.div_begin = ndiv ? (const int32_t *)(arena + i_db) : nullptr;
    db.div_tau = ndiv ? (const double *)(arena + i_dt) : nullptr;
    db.div_cash = ndiv ? (const double *)(arena + i_dc) : nullptr;
    db.n_div_total = ndiv;

    FdmBookOut dbo;
    dbo.price_log2 = (double *)(arena + o_d[0]); dbo.price = (double *)(arena + o_d[1]);
    dbo.delta = (double *)(arena + o_d[2]); dbo.gamma = (double *)(arena + o_d[3]);
    dbo.vega = (double *)(arena + o_d[4]); dbo.theta = (double *)(arena + o_d[5]);
    dbo.ex_lo = (int32_t *)(arena + o_lo); dbo.ex_hi = (int32_t *)(arena + o_hi); dbo.status = (int32_t *)(arena + o_st);
    dbo.out_index = nullptr;
    if (ho->out_index) {
        put(i_oi, ho->out_index, n * 4);
        FDM_TRY(cudaMemcpyAsync(arena + i_oi, hp + i_oi, n * 4, cudaMemcpyHostToDevice, st));
        dbo.out_index = (const int32_t *)(arena + i_oi);
    }

    if (rc == FDM_OK) {
        FDM_TRY(cudaMemsetAsync(arena + o_d[0], 0, o_st + up256(n * 4) - o_d[0], st));
        FDM_TRY(cudaEventRecord(cx.e0, st));
        const int lrc = american_book_queue(db, dbo, arena + o_ws, st, mask, n_space_max, roles, n_launches);
        if (lrc != FDM_OK) rc = lrc;
        FDM_TRY(cudaEventRecord(cx.e1, st));
    }
    auto get = [&](void *dst, size_t o, size_t bytes) { if (dst && bytes) memcpy(dst, hp + o, bytes); };
    if (rc == FDM_OK) {
        FDM_TRY(cudaMemcpyAsync(hp + o_d[0], arena + o_d[0], o_end - o_d[0], cudaMemcpyDeviceToHost, st));   // every result column
        FDM_TRY(cudaStreamSynchronize(st));
        if (rc == FDM_OK) {
            get(ho->price_log2, o_d[0], n * 8); get(ho->price, o_d[1], n * 8); get(ho->delta, o_d[2], n * 8);
            get(ho->gamma, o_d[3], n * 8); get(ho->vega, o_d[4], n * 8); get(ho->theta, o_d[5], n * 8);
            get(ho->ex_lo, o_lo, n * 4); get(ho->ex_hi, o_hi, n * 4); get(ho->status, o_st, n * 4);
        }
        if (kernel_ms && rc == FDM_OK) FDM_TRY(cudaEventElapsedTime(kernel_ms, cx.e0, cx.e1));
    }
    return rc;
}

int fdm_barrier_book_solves_per_option(int32_t what)
{
    if (what != FDM_BARRIER_BOOK_SOLVE && what != FDM_BARRIER_BOOK_GREEKS) return fail(FDM_ERR_INVALID, "unknown barrier book fan-out%s");
    return barrier_roles(what);
}

size_t fdm_barrier_book_workspace(int64_t n_options, int32_t what)
{
    if (n_options <= 0 || (what != FDM_BARRIER_BOOK_SOLVE && what != FDM_BARRIER_BOOK_GREEKS)) return 0;
    const int64_t m = n_options * barrier_roles(what);
    return solve_ws(m, m, true).total;
}

static int barrier_book_check(const FdmBarrierBook &bk)
{
    if (bk.n_options < 0) return fail(FDM_ERR_INVALID, "n_options < 0%s");
    if (bk.what != FDM_BARRIER_BOOK_SOLVE && bk.what != FDM_BARRIER_BOOK_GREEKS)
        return fail(FDM_ERR_INVALID, "unknown barrier book fan-out%s");
    if (bk.n_options == 0) return FDM_OK;
    if (!bk.S0 || !bk.K || !bk.T || !bk.sigma || !bk.r_disc || !bk.b_carry || !bk.x_min || !bk.dx || !bk.barrier ||
        !bk.rebate || !bk.flags || !bk.n_space || !bk.n_time)
        return fail(FDM_ERR_INVALID, "a required FdmBarrierBook array is NULL%s");
    if (bk.n_mon_total > 0 && (!bk.mon_begin || !bk.mon_step)) return fail(FDM_ERR_INVALID, "monitor tables missing%s");
    return FDM_OK;
}

int fdm_barrier_book(const FdmBarrierBook *book_, const FdmBookOut *out, void *workspace, size_t workspace_bytes,
                     void *stream, int32_t *n_launches)
{
    g_err[0] = 0;
    if (n_launches) *n_launches = 0;
    if (!book_ || !out) return fail(FDM_ERR_INVALID, "book / out is NULL%s");
    const FdmBarrierBook &bk = *book_;
    int rc = barrier_book_check(bk);
    if (rc != FDM_OK || bk.n_options == 0) return rc;
    if (bk.n_space_max < 3) return fail(FDM_ERR_INVALID, "n_space_max must be set (>= 3) for the device entry%s");
    if (bk.n_space_max > FDM_MAX_N_SPACE) return fail(FDM_ERR_UNSUPPORTED, "n_space exceeds FDM_MAX_N_SPACE (8193)%s");
    const size_t need = fdm_barrier_book_workspace(bk.n_options, bk.what);
    if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 255))
        return fail(FDM_ERR_INVALID, "workspace missing, misaligned or smaller than fdm_barrier_book_workspace()%s");
    if ((rc = no_device()) != FDM_OK) return rc;
    const int mask = (bk.hints & ~7) | 2;                            // linear solves, no dividends
    return barrier_book_queue(bk, *out, (char *)workspace, (cudaStream_t)stream, mask, bk.n_space_max, n_launches);
}

int fdm_barrier_book_host(const FdmBarrierBook *hb, const FdmBookOut *ho, int64_t *node_steps, int32_t *n_launches,
                          float *kernel_ms)
{
    g_err[0] = 0;
    if (n_launches) *n_launches = 0;
    if (kernel_ms) *kernel_ms = 0.f;
    if (node_steps) *node_steps = 0;
    if (!hb || !ho) return fail(FDM_ERR_INVALID, "book / out is NULL%s");
    int rc = barrier_book_check(*hb);
    const int64_t n = hb->n_options;
    if (rc != FDM_OK || n == 0) return rc;
    const int n_roles = barrier_roles(hb->what);
    int n_space_max = 3, small_max = 0, mask = 2;
    int64_t ns_total = 0;
    const int64_t nmon = (hb->mon_begin && hb->n_mon_total > 0) ? hb->mon_begin[n] : 0;
    for (int64_t i = 0; i < n; ++i) {
        // discrete_barrier_fdm_pricer_cn.py:61-66
        if (!(hb->T[i] > 0.0)) return fail(FDM_ERR_INVALID, "T must be positive%s");
        if (!(hb->sigma[i] - hb->dv_sigma * (n_roles > 1 ? 1.0 : 0.0) > 0.0)) return fail(FDM_ERR_INVALID, "sigma must be positive%s");
        if (!(hb->S0[i] > 0.0)) return fail(FDM_ERR_INVALID, "S0 must be positive%s");
        const int ns = hb->n_space[i];
        if (ns < 3 || hb->n_time[i] < 1) return fail(FDM_ERR_INVALID, "N_space must be >= 3 and N_time >= 1%s");
        if (ns > FDM_MAX_N_SPACE) return fail(FDM_ERR_UNSUPPORTED, "n_space exceeds FDM_MAX_N_SPACE (8193)%s");
        if (ns > n_space_max) n_space_max = ns;
        if (ns <= FDM_WARP_MAX_N_SPACE) {
            mask |= 8 | LEAN_HINTS_VALID | lean_hint_bits(ns);
            if (ns > small_max) small_max = ns;
        } else mask |= 16;
        if (nmon && hb->mon_begin[i + 1] < hb->mon_begin[i]) return fail(FDM_ERR_INVALID, "mon_begin must be ascending%s");
        ns_total += (int64_t)ns * hb->n_time[i] * n_roles;
    }
    mask |= small_max << 8;
    if (node_steps) *node_steps = ns_total;
    if ((rc = no_device()) != FDM_OK) return rc;

    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += up256(bytes); return o; };
    const int ND = 11;
    const double *hd[ND] = {hb->S0, hb->K, hb->T, hb->sigma, hb->r_disc, hb->b_carry, hb->x_min, hb->dx, hb->barrier,
                            hb->rebate, hb->barrier_up};
    size_t i_d[ND];
    for (int i = 0; i < ND; ++i) i_d[i] = take(hd[i] ? n * 8 : 0);
    const size_t i_fl = take(n * 4), i_ns = take(n * 4), i_nt = take(n * 4);
    const size_t i_mb = take(nmon ? (n + 1) * 4 : 0), i_ms = take(nmon * 4);
    const size_t o_d[5] = {take(n * 8), take(n * 8), take(n * 8), take(n * 8), take(n * 8)};
    const size_t o_st = take(n * 4);
    const size_t o_end = off;
    const size_t i_oi = take(ho->out_index ? n * 4 : 0);
    const size_t stage_end = off;                       // [0, stage_end) is mirrored in pinned host memory
    const size_t o_ws = take(fdm_barrier_book_workspace(n, hb->what));

    int dev = 0;
    FDM_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEVICES) return fail(FDM_ERR_UNSUPPORTED, "device index beyond the workspace table%s");
    Arena &ar = g_arenas[dev];
    std::lock_guard<std::mutex> arena_lock(ar.mu);
    char *arena = nullptr;
    FDM_CUDA(arena_get(ar, off, &arena));
    char *hp = nullptr;
    FDM_CUDA(arena_host(ar, stage_end, &hp));
    FDM_CUDA(arena_stream(ar));
    cudaStream_t st = ar.st;
    StreamDrain drain{st};
    struct { cudaEvent_t e0, e1; } cx = {ar.e0, ar.e1};
    auto put = [&](size_t o, const void *src, size_t bytes) { if (src && bytes) memcpy(hp + o, src, bytes); };
    for (int i = 0; i < ND; ++i) put(i_d[i], hd[i], n * 8);
    put(i_fl, hb->flags, n * 4); put(i_ns, hb->n_space, n * 4); put(i_nt, hb->n_time, n * 4);
    if (nmon) { put(i_mb, hb->mon_begin, (n + 1) * 4); put(i_ms, hb->mon_step, nmon * 4); }
    FDM_TRY(cudaMemcpyAsync(arena, hp, o_d[0], cudaMemcpyHostToDevice, st));       // every input array in one copy
    FdmBarrierBook db = *hb;
    const double **dp[ND] = {&db.S0, &db.K, &db.T, &db.sigma, &db.r_disc, &db.b_carry, &db.x_min, &db.dx, &db.barrier,
                             &db.rebate, &db.barrier_up};
    for (int i = 0; i < ND; ++i) *dp[i] = hd[i] ? (const double *)(arena + i_d[i]) : nullptr;
    db.flags = (const int32_t *)(arena + i_fl); db.n_space = (const int32_t *)(arena + i_ns);
    db.n_time = (const int32_t *)(arena + i_nt);
    db.mon_begin = nmon ? (const int32_t *)(arena + i_mb) : nullptr;
    db.mon_step = nmon ? (const int32_t *)(arena + i_ms) : nullptr;
    db.n_mon_total = nmon;

    FdmBookOut dbo;
    memset(&dbo, 0, sizeof(dbo));
    dbo.price = (double *)(arena + o_d[0]); dbo.delta = (double *)(arena + o_d[1]); dbo.gamma = (double *)(arena + o_d[2]);
    dbo.vega = (double *)(arena + o_d[3]); dbo.theta = (double *)(arena + o_d[4]); dbo.status = (int32_t *)(arena + o_st);
    if (ho->out_index) {
        put(i_oi, ho->out_index, n * 4);
        FDM_TRY(cudaMemcpyAsync(arena + i_oi, hp + i_oi, n * 4, cudaMemcpyHostToDevice, st));
        dbo.out_index = (const int32_t *)(arena + i_oi);
    }

    if (rc == FDM_OK) {
        FDM_TRY(cudaEventRecord(cx.e0, st));
        const int lrc = barrier_book_queue(db, dbo, arena + o_ws, st, mask, n_space_max, n_launches);
        if (lrc != FDM_OK) rc = lrc;
        FDM_TRY(cudaEventRecord(cx.e1, st));
    }
    auto get = [&](void *dst, size_t o, size_t bytes) { if (dst && bytes) memcpy(dst, hp + o, bytes); };
    if (rc == FDM_OK) {
        FDM_TRY(cudaMemcpyAsync(hp + o_d[0], arena + o_d[0], o_end - o_d[0], cudaMemcpyDeviceToHost, st));   // every result column
        FDM_TRY(cudaStreamSynchronize(st));
        if (rc == FDM_OK) {
            get(ho->price, o_d[0], n * 8); get(ho->delta, o_d[1], n * 8); get(ho->gamma, o_d[2], n * 8);
            get(ho->vega, o_d[3], n * 8); get(ho->theta, o_d[4], n * 8); get(ho->status, o_st, n * 4);
        }
        if (kernel_ms && rc == FDM_OK) FDM_TRY(cudaEventElapsedTime(kernel_ms, cx.e0, cx.e1));
    }
    return rc;
}

// ---- row-varying solves ---------------------------------------------------------------------------------------------

static int rowvar_check(const FdmRowVarBatch *b, const FdmRowVarOut *o)
{
    if (!b || !o) return fail(FDM_ERR_INVALID, "batch / out is NULL%s");
    if (b->n_solves < 0) return fail(FDM_ERR_INVALID, "n_solves < 0%s");
    if (b->n_solves == 0) return FDM_OK;
    if (!b->node_begin || !b->s_nodes || !b->v_terminal || !b->sigma || !b->r || !b->strike || !b->tenor || !b->dt ||
        !b->lo_bar || !b->up_bar || !b->h_minus || !b->h_plus || !b->s_eff || !b->n_steps || !b->rannacher || !b->flags ||
        !b->bar_j)
        return fail(FDM_ERR_INVALID, "a required FdmRowVarBatch array is NULL%s");
    if (!o->price) return fail(FDM_ERR_INVALID, "FdmRowVarOut.price is NULL%s");
    if (b->n_nodes_max < 3) return fail(FDM_ERR_INVALID, "n_nodes_max < 3%s");
    if (b->n_nodes_max > FDM_ROWVAR_MAX_NODES) return fail(FDM_ERR_UNSUPPORTED, "grid larger than FDM_ROWVAR_MAX_NODES (1024)%s");
    if (b->n_mon_total > 0 && (!b->mon_begin || !b->mon_step)) return fail(FDM_ERR_INVALID, "monitor tables missing%s");
    return FDM_OK;
}

static int rowvar_launch(const FdmRowVarBatch &b, const FdmRowVarOut &o, cudaStream_t st, int32_t *n_launches)
{
    if (o.status) FDM_CUDA(cudaMemsetAsync(o.status, 0xff, b.n_solves * 4, st));      // FDM_STATUS_NOT_SOLVED
    // one launch per size class that can be populated; a solve runs in the smallest class that holds it
    typedef int (*Fn)(const FdmRowVarBatch &, const FdmRowVarOut &, cudaStream_t, int, int32_t *);
    static const struct { int c; Fn fn; } cls[] = {{7, fdm_launch_rowvar_7},   {10, fdm_launch_rowvar_10}, {13, fdm_launch_rowvar_13},
                                                   {19, fdm_launch_rowvar_19}, {26, fdm_launch_rowvar_26}, {32, fdm_launch_rowvar_32}};
    int n_lo = 0;
    for (const auto &k : cls) {
        const int rc = k.fn(b, o, st, n_lo, n_launches);
        if (rc != FDM_OK) return rc;
        n_lo = 32 * k.c;
        if (b.n_nodes_max <= n_lo) break;
    }
    return FDM_OK;
}

int fdm_rowvar_solve(const FdmRowVarBatch *batch, const FdmRowVarOut *out, void *stream, int32_t *n_launches)
{
    g_err[0] = 0;
    if (n_launches) *n_launches = 0;
    int rc = rowvar_check(batch, out);
    if (rc != FDM_OK || batch->n_solves == 0) return rc;
    if ((rc = no_device()) != FDM_OK) return rc;
    return rowvar_launch(*batch, *out, (cudaStream_t)stream, n_launches);
}

int fdm_rowvar_solve_host(const FdmRowVarBatch *hb, const FdmRowVarOut *ho, int32_t *n_launches, float *kernel_ms)
{
    g_err[0] = 0;
    if (n_launches) *n_launches = 0;
    if (kernel_ms) *kernel_ms = 0.f;
    int rc = rowvar_check(hb, ho);
    if (rc != FDM_OK || hb->n_solves == 0) return rc;
    const int64_t n = hb->n_solves;
    if (hb->node_begin[0] != 0) return fail(FDM_ERR_INVALID, "node_begin[0] must be 0%s");
    for (int64_t i = 0; i < n; ++i) {
        const int64_t k = hb->node_begin[i + 1] - hb->node_begin[i];
        if (k < 3 || k > hb->n_nodes_max) return fail(FDM_ERR_INVALID, "node count outside [3, n_nodes_max]%s");
        if (hb->n_steps[i] < 1) return fail(FDM_ERR_INVALID, "n_steps < 1%s");
    }
    if ((rc = no_device()) != FDM_OK) return rc;
    const int64_t nn = hb->node_begin[n];
    const int64_t nmon = (hb->mon_begin && hb->n_mon_total > 0) ? hb->mon_begin[n] : 0;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += up256(bytes); return o; };
    const size_t o_nb = take((n + 1) * 8), o_s = take(nn * 8), o_v = take(nn * 8);
    const int ND = 10;
    const double *hd[ND] = {hb->sigma, hb->r, hb->strike, hb->tenor, hb->dt, hb->lo_bar, hb->up_bar, hb->h_minus, hb->h_plus, hb->s_eff};
    size_t o_d[ND];
    for (int i = 0; i < ND; ++i) o_d[i] = take(n * 8);
    const int NI = 4;
    const int32_t *hi[NI] = {hb->n_steps, hb->rannacher, hb->flags, hb->bar_j};
    size_t o_i[NI];
    for (int i = 0; i < NI; ++i) o_i[i] = take(n * 4);
    const size_t o_mb = take(nmon ? (n + 1) * 4 : 0), o_ms = take(nmon * 4);
    const size_t r_p = take(n * 8), r_st = take(n * 4);
    const size_t r_dump = take(ho->v_dump ? (size_t)n * ho->v_dump_stride * 8 : 0);

    int dev = 0;
    FDM_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEVICES) return fail(FDM_ERR_UNSUPPORTED, "device index beyond the workspace table%s");
    Arena &ar = g_arenas[dev];
    std::lock_guard<std::mutex> arena_lock(ar.mu);
    char *arena = nullptr;
    FDM_CUDA(arena_get(ar, off, &arena));
    CallCtx cx;
    FDM_TRY(cudaStreamCreateWithFlags(&cx.st, cudaStreamNonBlocking));
    FDM_TRY(cudaEventCreate(&cx.e0));
    FDM_TRY(cudaEventCreate(&cx.e1));
    if (rc != FDM_OK) return rc;
    cudaStream_t st = cx.st;
    auto h2d = [&](size_t o, const void *src, size_t bytes) {
        if (bytes && src) FDM_TRY(cudaMemcpyAsync(arena + o, src, bytes, cudaMemcpyHostToDevice, st));
    };
    h2d(o_nb, hb->node_begin, (n + 1) * 8);
    h2d(o_s, hb->s_nodes, nn * 8);
    h2d(o_v, hb->v_terminal, nn * 8);
    for (int i = 0; i < ND; ++i) h2d(o_d[i], hd[i], n * 8);
    for (int i = 0; i < NI; ++i) h2d(o_i[i], hi[i], n * 4);
    if (nmon) { h2d(o_mb, hb->mon_begin, (n + 1) * 4); h2d(o_ms, hb->mon_step, nmon * 4); }
    FdmRowVarBatch db = *hb;
    db.node_begin = (const int64_t *)(arena + o_nb);
    db.s_nodes = (const double *)(arena + o_s);
    db.v_terminal = (const double *)(arena + o_v);
    const double **dp[ND] = {&db.sigma, &db.r, &db.strike, &db.tenor, &db.dt, &db.lo_bar, &db.up_bar, &db.h_minus, &db.h_plus, &db.s_eff};
    for (int i = 0; i < ND; ++i) *dp[i] = (const double *)(arena + o_d[i]);
    const int32_t **ip[NI] = {&db.n_steps, &db.rannacher, &db.flags, &db.bar_j};
    for (int i = 0; i < NI; ++i) *ip[i] = (const int32_t *)(arena + o_i[i]);
    db.mon_begin = nmon ? (const int32_t *)(arena + o_mb) : nullptr;
    db.mon_step = nmon ? (const int32_t *)(arena + o_ms) : nullptr;
    FdmRowVarOut dout = *ho;
    dout.price = (double *)(arena + r_p);
    dout.status = (int32_t *)(arena + r_st);
    dout.v_dump = ho->v_dump ? (double *)(arena + r_dump) : nullptr;
    if (ho->v_dump) FDM_TRY(cudaMemsetAsync(arena + r_dump, 0, (size_t)n * ho->v_dump_stride * 8, st));
    if (rc == FDM_OK) {
        FDM_TRY(cudaEventRecord(cx.e0, st));
        const int lrc = rowvar_launch(db, dout, st, n_launches);
        if (lrc != FDM_OK) rc = lrc;
        FDM_TRY(cudaEventRecord(cx.e1, st));
    }
    if (rc == FDM_OK) {
        FDM_TRY(cudaMemcpyAsync(ho->price, arena + r_p, n * 8, cudaMemcpyDeviceToHost, st));
        if (ho->status) FDM_TRY(cudaMemcpyAsync(ho->status, arena + r_st, n * 4, cudaMemcpyDeviceToHost, st));
        if (ho->v_dump) FDM_TRY(cudaMemcpyAsync(ho->v_dump, arena + r_dump, (size_t)n * ho->v_dump_stride * 8, cudaMemcpyDeviceToHost, st));
        FDM_TRY(cudaStreamSynchronize(st));
        if (kernel_ms && rc == FDM_OK) FDM_TRY(cudaEventElapsedTime(kernel_ms, cx.e0, cx.e1));
    }
    return rc;
}

int fdm_launch_hints(const int32_t *n_space, const int32_t *flags, int64_t n_solves, const double *seg_div,
                     int64_t n_seg)
{
    if (!n_space || !flags || n_solves < 0) return 0;
    int mask = 0, small_max = 0;
    for (int64_t i = 0; i < n_solves; ++i) {
        mask |= (flags[i] & FDM_FLAG_AMERICAN) ? 1 : 2;
        if (n_space[i] <= FDM_WARP_MAX_N_SPACE) {
            mask |= 8 | LEAN_HINTS_VALID | lean_hint_bits(n_space[i]);
            if (n_space[i] > small_max) small_max = n_space[i];
        } else {
            mask |= 16;
        }
    }
    mask |= small_max << 8;
    if (seg_div) {
        for (int64_t i = 0; i < n_seg; ++i) if (seg_div[i] != 0.0) { mask |= 4; break; }
    } else {
        mask |= 4;
    }
    return mask;
}

int fdm_release_workspace(void)
{
    int cur = 0;
    const bool have_cur = cudaGetDevice(&cur) == cudaSuccess;
    for (int d = 0; d < MAX_DEVICES; ++d) {
        Arena &a = g_arenas[d];
        std::lock_guard<std::mutex> arena_lock(a.mu);
        if (a.p || a.hp || a.st) {
            cudaSetDevice(d);
            if (a.st) cudaStreamSynchronize(a.st);
            if (a.p) cudaFree(a.p);
            if (a.hp) cudaFreeHost(a.hp);
            if (a.e0) cudaEventDestroy(a.e0);
            if (a.e1) cudaEventDestroy(a.e1);
            if (a.st) cudaStreamDestroy(a.st);
        }
        a.p = nullptr; a.bytes = 0; a.hp = nullptr; a.hbytes = 0; a.st = nullptr; a.e0 = a.e1 = nullptr;
    }
    if (have_cur) cudaSetDevice(cur);
    return FDM_OK;
}

int fdm_fp64_peak(double *tflops_out, float *ms_out)
{
    g_err[0] = 0;
    if (!tflops_out) return fail(FDM_ERR_INVALID, "tflops_out is NULL%s");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(FDM_ERR_NO_DEVICE, "no CUDA device%s");
    int dev = 0, sms = 0;
    FDM_CUDA(cudaGetDevice(&dev));
    FDM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double *sink = nullptr;
    const int blocks = sms * 8, threads = 256, iters = 4096;
    FDM_CUDA(cudaMalloc((void **)&sink, blocks * sizeof(double)));
    cudaEvent_t e0, e1;
    FDM_CUDA(cudaEventCreate(&e0));
    FDM_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        FDM_CUDA(cudaEventRecord(e0));
        fp64_peak_kernel<<<blocks, threads>>>(sink, iters, 1.0 + rep);
        FDM_CUDA(cudaEventRecord(e1));
        FDM_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        FDM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    const double flops = 2.0 * 64.0 * (double)iters * (double)threads * (double)blocks;
    *tflops_out = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return FDM_OK;
}

}  // extern "C"
