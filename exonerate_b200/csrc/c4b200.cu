// c4b200.cu -- the C ABI of include/c4b200.h: engine, staging, kernel dispatch.
//
// Host side of the drop-in boundary.  Mirrors, batched, what Optimal_find_score /
// Optimal_find_path (src/c4/optimal.c:123-133,368-413) do for one lattice:
// choose the fill variant, run it, trace back, hand an operation list to the
// caller.  No CPU fallback exists: every entry point needs a CUDA device.
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cstring>
#include <functional>
#include <map>
#include <set>
#include <string>
#include <thread>
#include <vector>
#include <unistd.h>

#include "affine_systolic.cuh"
#include "affine_packed16.cuh"
#include "affine_traceback.cuh"
#include "generic_wavefront.cuh"
#include "e2g_systolic.cuh"
#include "e2g_packed16.cuh"
#include "hsp_extend.cuh"
#include "span_integrate.cuh"

namespace c4b {
static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }
}  // namespace c4b

using namespace c4b;

struct c4b_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    int sm_count = 0;
    int64_t launches = 0;
    // pinned bounce buffer for host->device staging, grow-only, reused by every batch
    uint8_t *h_stage = nullptr;
    size_t h_stage_cap = 0;
    cudaEvent_t stage_free = nullptr;  // last H2D that read h_stage
    // staging pipeline of the affine path: sequence slices are copied and encoded on
    // copy_stream while the score pass of the lattices that are already resident
    // runs on the aux streams (concurrent launches fill each other's tails)
    cudaStream_t copy_stream = nullptr;
    cudaStream_t aux[4] = {nullptr, nullptr, nullptr, nullptr};
    // device copies of host buffers the caller declared stable (C4B_PAIR_BUFFERS_STABLE):
    // (host address, bytes) -> device address; dropped by c4b_engine_forget_buffers
    c4b::ResidentBuffers resident;
};

namespace {

struct EventPair {
    cudaEvent_t a = nullptr, b = nullptr;
};

struct Chunk {
    int begin, end;  // range in the ordered lattice list
};

// Device buffers come from the stream-ordered pool of the device (cudaMallocAsync
// with an unbounded release threshold, set in c4b_engine_create): after the first
// batch, allocation and free are bookkeeping, not driver calls that serialise.
static thread_local cudaStream_t tl_pool_stream = nullptr;

// C4B_TIMING=1: host-side timeline of one batch on stderr (tuning aid)
struct HostTimeline {
    bool on;
    std::chrono::steady_clock::time_point t0;
    HostTimeline() : on(getenv("C4B_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(const char *what) {
        if (!on) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "[c4b timing] %8.2f ms  %s\n", ms, what);
    }
};
static thread_local HostTimeline *tl_timeline = nullptr;
static void tmark(const char *what) { if (tl_timeline) tl_timeline->mark(what); }

// worker threads for host-side staging: min(16, cores), or C4B_HOST_THREADS (one process
// per GPU on a shared host should divide the cores: bench.py sets cores / world size)
static unsigned host_threads() {
    unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (const char *env = getenv("C4B_HOST_THREADS")) hw = (unsigned)std::max(1, std::min(64, atoi(env)));
    return hw;
}

// Grow-only pinned staging memory of this host thread (pinning is far too slow to
// repeat per batch); `busy` is the last copy that read it.
struct PinnedScratch {
    uint8_t *p = nullptr;
    size_t cap = 0;
    cudaEvent_t busy = nullptr;
    uint8_t *get(size_t bytes) {
        if (busy) cudaEventSynchronize(busy);
        if (cap < bytes) {
            if (p) cudaFreeHost(p);
            p = nullptr;
            cap = 0;
            if (cudaMallocHost(&p, bytes + bytes / 8) != cudaSuccess) return nullptr;
            cap = bytes + bytes / 8;
        }
        return p;
    }
    void mark(cudaStream_t st) {
        if (!busy) cudaEventCreateWithFlags(&busy, cudaEventDisableTiming);
        cudaEventRecord(busy, st);
    }
};

template <typename F>
static void parallel_for(int n, F f) {
    const unsigned nt = std::max(1u, std::min<unsigned>(host_threads(), (unsigned)n));
    if (nt <= 1) {
        for (int k = 0; k < n; ++k) f(k);
        return;
    }
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; ++t)
        th.emplace_back([&, t] { for (int k = (int)t; k < n; k += (int)nt) f(k); });
    for (int k = 0; k < n; k += (int)nt) f(k);
    for (auto &x : th) x.join();
}

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaStream_t stream = nullptr;
    int alloc(size_t count) {
        n = count;
        if (!count) return 0;
        stream = tl_pool_stream;
        C4B_CUDA(cudaMallocAsync(&p, count * sizeof(T), stream));
        return 0;
    }
    void release() {
        if (p) cudaFreeAsync(p, stream);
        p = nullptr;
    }
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- model analysis ---------------------------------------------------------
// Is the closed model the affine template of SURVEY.md §8a?  (Checked, not
// assumed: the order is read from the tables the caller passes.)
bool analyze_affine(const c4b_model &m, AffModel *am, int *match_kind) {
    if (m.n_states != 5 || m.n_transitions != 9 || m.n_shadow_slots != 0) return false;
    if (m.max_query_advance != 1 || m.max_target_advance != 1) return false;
    const int S = m.start_state, E = m.end_state;
    const c4b_transition *t = m.transitions;
    auto is_const = [&](int k) { return t[k].calc >= 0 && m.calcs[t[k].calc].kind == C4B_CALC_CONST &&
                                        m.calcs[t[k].calc].protect == 0; };
    auto adv = [&](int k, int aq, int at) { return t[k].advance_query == aq && t[k].advance_target == at; };
    // T4 fixes the match state, T2/T3 the delete/insert states
    if (!adv(4, 1, 1) || t[4].input != t[4].output || t[4].calc < 0) return false;
    const int M = t[4].input;
    const int mk = m.calcs[t[4].calc].kind;
    if (mk != C4B_CALC_MATCH_DNA && mk != C4B_CALC_MATCH_PROTEIN) return false;
    if (m.calcs[t[4].calc].protect != 0) return false;
    if (!adv(2, 0, 1) || t[2].input != M || !is_const(2)) return false;
    const int D = t[2].output;
    if (!adv(3, 1, 0) || t[3].input != M || !is_const(3)) return false;
    const int I = t[3].output;
    if (M == S || M == E || D == M || I == M || D == I || D == S || D == E || I == S || I == E) return false;
    if (!adv(0, 0, 1) || t[0].input != D || t[0].output != D || !is_const(0)) return false;
    if (!adv(1, 1, 0) || t[1].input != I || t[1].output != I || !is_const(1)) return false;
    if (!adv(5, 0, 0) || t[5].input != S || t[5].output != M || t[5].calc >= 0) return false;
    if (!adv(6, 0, 0) || t[6].input != D || t[6].output != M || t[6].calc >= 0) return false;
    if (!adv(7, 0, 0) || t[7].input != I || t[7].output != M || t[7].calc >= 0) return false;
    if (!adv(8, 0, 0) || t[8].input != M || t[8].output != E || t[8].calc >= 0) return false;
    if (t[4].label != C4B_LABEL_MATCH) return false;
    am->extD = m.calcs[t[0].calc].param[0];
    am->extI = m.calcs[t[1].calc].param[0];
    am->openD = m.calcs[t[2].calc].param[0];
    am->openI = m.calcs[t[3].calc].param[0];
    // the pad-row argument of the systolic kernel needs strictly negative gaps
    if (am->extD >= 0 || am->extI >= 0 || am->openD >= 0 || am->openI >= 0) return false;
    // the local instantiation assumes START and END are both ANYWHERE (Affine_create
    // always configures the two scopes alike, src/model/affine.c:189-190)
    if ((m.start_scope == C4B_SCOPE_ANYWHERE) != (m.end_scope == C4B_SCOPE_ANYWHERE)) return false;
    // the kernel carries M + open (one shared open penalty, as Affine_create builds it)
    if (am->openD != am->openI) return false;
    am->one = 1;
    am->start_scope = m.start_scope;
    am->end_scope = m.end_scope;
    am->tDD = 0; am->tII = 1; am->tMD = 2; am->tMI = 3; am->tMM = 4;
    am->tSM = 5; am->tDM = 6; am->tIM = 7; am->tME = 8;
    *match_kind = mk;
    return true;
}

__global__ void encode_kernel(uint8_t *buf, size_t n, const uint8_t *__restrict__ lut, int *bad,
                              const uint32_t safe = 0xFFu) {
    // raw symbol bytes -> matrix codes, 16 bytes per thread; a symbol outside the
    // substitution matrix alphabet (LUT entry 0xFF; the reference reads out of
    // bounds there, src/sequence/submat.c:26-55) raises *bad and is stored as
    // `safe`, so a fill that was queued before the host saw the flag stays in bounds
    const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t base = v * 16;
    if (base >= n) return;
    if (base + 16 <= n) {
        uint4 w = *reinterpret_cast<uint4 *>(buf + base);
        uint32_t a[4] = {w.x, w.y, w.z, w.w};
        bool any_bad = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t o = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                uint32_t c = lut[(a[k] >> (8 * b)) & 255u];
                if (c == 0xFFu) { any_bad = true; c = safe; }
                o |= c << (8 * b);
            }
            a[k] = o;
        }
        *reinterpret_cast<uint4 *>(buf + base) = make_uint4(a[0], a[1], a[2], a[3]);
        if (any_bad) atomicOr(bad, 1);
    } else {
        for (size_t k = base; k < n; ++k) {
            uint8_t c = lut[buf[k]];
            if (c == 0xFF) { atomicOr(bad, 1); c = (uint8_t)safe; }
            buf[k] = c;
        }
    }
}

__global__ void apply_threshold_kernel(c4b_result *results, int n, int threshold) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    if (results[p].status == 0 && results[p].score < threshold) {
        results[p].status = 1;  // Optimal_find_path returns NULL (optimal.c:408-411)
        results[p].n_ops = 0;
    }
}

// exclusive scan of n_ops in result order (single block; n is a batch size)
__global__ void ops_scan_kernel(const c4b_result *results, int n, int64_t *new_off, int64_t *total) {
    __shared__ int64_t part[1024];
    const int tid = threadIdx.x;
    const int per = (n + 1023) / 1024;
    const int b = tid * per, e = min(n, b + per);
    int64_t s = 0;
    for (int k = b; k < e; ++k) s += results[k].n_ops;
    part[tid] = s;
    __syncthreads();
    if (tid == 0) {
        int64_t run = 0;
        for (int k = 0; k < 1024; ++k) {
            const int64_t v = part[k];
            part[k] = run;
            run += v;
        }
        *total = run;
    }
    __syncthreads();
    int64_t off = part[tid];
    for (int k = b; k < e; ++k) {
        new_off[k] = off;
        off += results[k].n_ops;
    }
}

__global__ void ops_compact_kernel(c4b_result *results, int n, const int64_t *new_off,
                                   const int32_t *__restrict__ slots, int32_t *__restrict__ packed) {
    const int p = blockIdx.x;
    if (p >= n) return;
    const int64_t src = results[p].ops_offset, dst = new_off[p];
    const int cnt = results[p].n_ops * 2;
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) packed[2 * dst + k] = slots[2 * src + k];
    __syncthreads();
    if (threadIdx.x == 0) results[p].ops_offset = dst;
}

}  // namespace

#include "generic_jit.inl"
#include "generic_host.inl"
#include "e2g_host.inl"

// =============================================================================
struct c4b_batch {
    c4b_engine *e = nullptr;
    int n = 0;
    bool want_path = false;
    bool affine = false;
    c4b_model model;
    c4b_scoring scoring;
    int64_t cells = 0;
    const char *kernel_name = "none";
    std::string description;   // which lattices took which kernels (c4b_batch_description)
    bool ran = false;

    // ---- affine path ----
    AffModel aff;
    int R = 32, score_mode = SCORE_PRMT, max_sub = 0, gap_min = 0;
    int fill_warps = 1;  // warps per lattice of the int32 fill (concurrent sweeps of long queries)
    int R32s = 32;       // rows per lane of the int32 SCORE pass (fitted to the longest query; the recording pass keeps R)
    int n16 = 0;  // leading lattices of score_list that take the packed 16-bit score pass
    bool p16_unsigned = false;  // offset-binary variant (affine_fill16u_kernel) is applicable
    bool p16_multi = false;     // some packed lattice needs more than one sweep
    // rows per lane / warps per lattice PAIR of the packed score pass.  Normally R16 = R and one
    // sweep; a batch too small to fill the GPU with one warp per pair takes a smaller R16, so
    // that every query becomes several sweeps that run as pipelined warps of one CTA
    int R16 = 32, fill_warps16 = 1;
    // small batches: ONE lattice per warp, the two register halves on the two halves of its rows
    // (affine_fill16f_kernel), so that a batch of n lattices is n warps instead of n / 2
    bool p16_fold = false;
    int Rf = 16;                // folded score pass: rows per lane and half
    bool tb16_band = false, tb16_direct = false;  // traceback pass on affine_fill16tb_kernel
    std::vector<int> score_list, direct_list;  // original pair indices, cost-descending
    std::vector<Chunk> band_chunks, direct_chunks;
    DevBuf<uint8_t> d_seq;
    DevBuf<uint8_t> d_lut;          // [0,256) query LUT, [256,512) target LUT
    DevBuf<uint8_t> d_score_table;
    DevBuf<int> d_bad;
    DevBuf<AffPair> d_full, d_band, d_direct;
    DevBuf<AffOut> d_out1, d_out2, d_outd;
    DevBuf<int32_t> d_band_j0, d_qorg, d_torg;
    DevBuf<TbJob> d_jobs_band, d_jobs_direct;
    DevBuf<uint32_t> d_tb;
    DevBuf<int2> d_top;
    DevBuf<int2> d_blk;             // SubOpt blocked cells as {column, row mask} per lane strip
    DevBuf<int32_t> d_blk_off;
    bool any_blocked = false;       // some lattice has blocked cells: int32 launches use the BLK kernels
    DevBuf<c4b_result> d_results;
    DevBuf<int32_t> d_ops_slots, d_ops_packed;
    DevBuf<int64_t> d_new_off;  // n + 1 (last = total)
    size_t q_bytes = 0, t_bytes = 0;
    std::vector<EventPair> fill_events;
    int fill_events_used = 0;
    // score pass (pass 1) launch groups: a contiguous range of score_list that becomes
    // runnable when staging slice `slice` is resident
    struct P1Group { int begin, end, slice; cudaEvent_t done; };
    std::vector<P1Group> groups;
    std::vector<cudaEvent_t> slice_events;
    cudaEvent_t ev_base = nullptr, ev_p1a = nullptr, ev_p1b = nullptr;
    bool pass1_inflight = false;  // launched by affine_create while staging was still running
    bool bad_checked = false;

    // ---- generic path ----
    GenericBatch *generic = nullptr;

    // ---- est2genome systolic path ----
    E2gBatch *e2g = nullptr;

    ~c4b_batch() {
        if (e && affine) {  // frees are ordered on e->stream: drain the side streams first
            if (e->copy_stream) cudaStreamSynchronize(e->copy_stream);
            for (cudaStream_t a : e->aux)
                if (a) cudaStreamSynchronize(a);
        }
        d_seq.release(); d_lut.release(); d_score_table.release(); d_bad.release();
        d_full.release(); d_band.release(); d_direct.release();
        d_out1.release(); d_out2.release(); d_outd.release();
        d_band_j0.release(); d_qorg.release(); d_torg.release();
        d_jobs_band.release(); d_jobs_direct.release();
        d_tb.release(); d_top.release(); d_blk.release(); d_blk_off.release(); d_results.release();
        d_ops_slots.release(); d_ops_packed.release(); d_new_off.release();
        for (auto &ev : fill_events) {
            if (ev.a) cudaEventDestroy(ev.a);
            if (ev.b) cudaEventDestroy(ev.b);
        }
        for (auto &g : groups)
            if (g.done) cudaEventDestroy(g.done);
        for (auto ev : slice_events)
            if (ev) cudaEventDestroy(ev);
        for (cudaEvent_t ev : {ev_base, ev_p1a, ev_p1b})
            if (ev) cudaEventDestroy(ev);
        if (generic) generic_batch_destroy(generic);
        delete e2g;
    }
};

namespace {

template <int R, bool TB, int ENDMODE>
void launch_fill_sm(c4b_batch *b, const AffPair *pairs, AffOut *outs, int count, cudaStream_t s) {
    const int threads = 32 * b->fill_warps;  // warps per lattice = concurrent sweeps (affine_systolic.cuh)
    if (b->any_blocked) {   // SubOpt: the variant that masks T4 at blocked cells
        if (b->score_mode == SCORE_PRMT)
            affine_fill_kernel<R, TB, ENDMODE, SCORE_PRMT, true><<<count, threads, 0, s>>>(pairs, outs, b->aff, b->d_score_table.p);
        else
            affine_fill_kernel<R, TB, ENDMODE, SCORE_SMEM, true><<<count, threads, 0, s>>>(pairs, outs, b->aff, b->d_score_table.p);
        return;
    }
    if (b->score_mode == SCORE_PRMT)
        affine_fill_kernel<R, TB, ENDMODE, SCORE_PRMT><<<count, threads, 0, s>>>(pairs, outs, b->aff, b->d_score_table.p);
    else
        affine_fill_kernel<R, TB, ENDMODE, SCORE_SMEM><<<count, threads, 0, s>>>(pairs, outs, b->aff, b->d_score_table.p);
}

template <int R>
void launch_fill_r(c4b_batch *b, const AffPair *pairs, AffOut *outs, int count, bool tb, cudaStream_t s) {
    const bool any = (b->aff.end_scope == C4B_SCOPE_ANYWHERE);
    if (tb) {
        if (any) launch_fill_sm<R, true, END_ANYWHERE>(b, pairs, outs, count, s);
        else launch_fill_sm<R, true, END_RESTRICTED>(b, pairs, outs, count, s);
    } else {
        if (any) launch_fill_sm<R, false, END_ANYWHERE>(b, pairs, outs, count, s);
        else launch_fill_sm<R, false, END_RESTRICTED>(b, pairs, outs, count, s);
    }
}

// score pass only, no SubOpt lists: the rows-per-lane values that exist for fitting a query (R32s)
template <int R>
void launch_fill_score(c4b_batch *b, const AffPair *pairs, AffOut *outs, int count, cudaStream_t s) {
    const int threads = 32 * b->fill_warps;
    const bool any = (b->aff.end_scope == C4B_SCOPE_ANYWHERE), prmt = (b->score_mode == SCORE_PRMT);
    if (any && prmt) affine_fill_kernel<R, false, END_ANYWHERE, SCORE_PRMT><<<count, threads, 0, s>>>(pairs, outs, b->aff, b->d_score_table.p);
    else if (any) affine_fill_kernel<R, false, END_ANYWHERE, SCORE_SMEM><<<count, threads, 0, s>>>(pairs, outs, b->aff, b->d_score_table.p);
    else if (prmt) affine_fill_kernel<R, false, END_RESTRICTED, SCORE_PRMT><<<count, threads, 0, s>>>(pairs, outs, b->aff, b->d_score_table.p);
    else affine_fill_kernel<R, false, END_RESTRICTED, SCORE_SMEM><<<count, threads, 0, s>>>(pairs, outs, b->aff, b->d_score_table.p);
}

// timed = bracket the launch with an event pair on its stream (c4b_batch_last_fill_ms)
int launch_fill(c4b_batch *b, const AffPair *pairs, AffOut *outs, int count, bool tb, cudaStream_t s,
                bool timed) {
    if (!count) return 0;
    EventPair *ev = nullptr;
    if (timed) {
        if ((size_t)b->fill_events_used >= b->fill_events.size()) {
            EventPair nev;
            C4B_CUDA(cudaEventCreate(&nev.a));
            C4B_CUDA(cudaEventCreate(&nev.b));
            b->fill_events.push_back(nev);
        }
        ev = &b->fill_events[b->fill_events_used++];
        C4B_CUDA(cudaEventRecord(ev->a, s));
    }
    switch (tb ? b->R : b->R32s) {
    case 8: launch_fill_r<8>(b, pairs, outs, count, tb, s); break;
    case 12: launch_fill_score<12>(b, pairs, outs, count, s); break;
    case 16: launch_fill_r<16>(b, pairs, outs, count, tb, s); break;
    case 20: launch_fill_score<20>(b, pairs, outs, count, s); break;
    case 24: launch_fill_score<24>(b, pairs, outs, count, s); break;
    case 28: launch_fill_score<28>(b, pairs, outs, count, s); break;
    default: launch_fill_r<32>(b, pairs, outs, count, tb, s); break;
    }
    C4B_CUDA(cudaGetLastError());
    if (timed) C4B_CUDA(cudaEventRecord(ev->b, s));
    b->e->launches++;
    return 0;
}

// score pass of the first `count` lattices of d_full, two per warp (affine_packed16.cuh)
int launch_fill16(c4b_batch *b, const AffPair *pairs, AffOut *outs, int count, cudaStream_t s) {
    if (!count) return 0;
    const int blocks = (count + 1) / 2;
    // the offset-binary variant sweeps long queries with up to 8 pipelined warps per CTA
    const int threads = (b->p16_unsigned && b->p16_multi) ? 32 * b->fill_warps16 : 32;
    auto go = [&](auto kernel) -> int {
        kernel<<<blocks, threads, 0, s>>>(pairs, outs, count, b->aff, b->d_score_table.p);
        return 0;
    };
    int rc;
    if (b->p16_fold) {
        auto gof = [&](auto kernel) -> int {
            kernel<<<count, 32, 0, s>>>(pairs, outs, count, b->aff, b->d_score_table.p);
            return 0;
        };
        switch (b->Rf) {   // rows per lane and half, fitted to the longest query (even numbers)
        case 6: rc = gof(affine_fill16f_kernel<6>); break;
        case 8: rc = gof(affine_fill16f_kernel<8>); break;
        case 10: rc = gof(affine_fill16f_kernel<10>); break;
        case 12: rc = gof(affine_fill16f_kernel<12>); break;
        case 14: rc = gof(affine_fill16f_kernel<14>); break;
        default: rc = gof(affine_fill16f_kernel<16>); break;
        }
    } else if (b->p16_unsigned && b->p16_multi) {   // some query is longer than one sweep of 32 R16 rows
        switch (b->R16) {
        case 8: rc = go(affine_fill16u_multi_kernel<8>); break;
        case 16: rc = go(affine_fill16u_multi_kernel<16>); break;
        case 20: rc = go(affine_fill16u_multi_kernel<20>); break;
        case 24: rc = go(affine_fill16u_multi_kernel<24>); break;
        case 28: rc = go(affine_fill16u_multi_kernel<28>); break;
        default: rc = go(affine_fill16u_multi_kernel<32>); break;
        }
    } else if (b->p16_unsigned) {
        switch (b->R16) {   // rows per lane fitted to the longest query (multiples of 4)
        case 4: rc = go(affine_fill16u_kernel<4>); break;
        case 6: rc = go(affine_fill16u_kernel<6>); break;
        case 8: rc = go(affine_fill16u_kernel<8>); break;
        case 12: rc = go(affine_fill16u_kernel<12>); break;
        case 16: rc = go(affine_fill16u_kernel<16>); break;
        case 20: rc = go(affine_fill16u_kernel<20>); break;
        case 24: rc = go(affine_fill16u_kernel<24>); break;
        case 28: rc = go(affine_fill16u_kernel<28>); break;
        default: rc = go(affine_fill16u_kernel<32>); break;
        }
    } else {
        switch (b->R) {
        case 8: rc = go(affine_fill16_kernel<8>); break;
        case 16: rc = go(affine_fill16_kernel<16>); break;
        default: rc = go(affine_fill16_kernel<32>); break;
        }
    }
    if (rc) return rc;
    C4B_CUDA(cudaGetLastError());
    b->e->launches++;
    return 0;
}

// traceback pass of `count` lattices, two per warp (affine_fill16tb_kernel), timed like launch_fill
int launch_fill16tb(c4b_batch *b, const AffPair *pairs, AffOut *outs, int count, cudaStream_t s) {
    if (!count) return 0;
    if ((size_t)b->fill_events_used >= b->fill_events.size()) {
        EventPair nev;
        C4B_CUDA(cudaEventCreate(&nev.a));
        C4B_CUDA(cudaEventCreate(&nev.b));
        b->fill_events.push_back(nev);
    }
    EventPair *ev = &b->fill_events[b->fill_events_used++];
    C4B_CUDA(cudaEventRecord(ev->a, s));
    const int blocks = (count + 1) / 2;
    switch (b->R) {
    case 8: affine_fill16tb_kernel<8><<<blocks, 32, 0, s>>>(pairs, outs, count, b->aff, b->d_score_table.p); break;
    case 16: affine_fill16tb_kernel<16><<<blocks, 32, 0, s>>>(pairs, outs, count, b->aff, b->d_score_table.p); break;
    default: affine_fill16tb_kernel<32><<<blocks, 32, 0, s>>>(pairs, outs, count, b->aff, b->d_score_table.p); break;
    }
    C4B_CUDA(cudaGetLastError());
    C4B_CUDA(cudaEventRecord(ev->b, s));
    b->e->launches++;
    return 0;
}

// Pass 1 (score + END cell, nothing written per cell) of one launch group, on an aux
// stream, as soon as the group's sequences are resident.
int launch_pass1_group(c4b_batch *b, int g) {
    c4b_batch::P1Group &G = b->groups[g];
    cudaStream_t s = b->e->aux[g % 4];
    C4B_CUDA(cudaStreamWaitEvent(s, b->ev_base, 0));
    C4B_CUDA(cudaStreamWaitEvent(s, b->slice_events[G.slice], 0));
    const int a16 = G.begin, b16 = std::min(G.end, b->n16);
    const int a32 = std::max(G.begin, b->n16), b32 = G.end;
    if (b16 > a16 && launch_fill16(b, b->d_full.p + a16, b->d_out1.p, b16 - a16, s)) return -1;
    if (b32 > a32 && launch_fill(b, b->d_full.p + a32, b->d_out1.p, b32 - a32, false, s, false)) return -1;
    C4B_CUDA(cudaEventRecord(G.done, s));
    return 0;
}

int launch_traceback(c4b_batch *b, const AffPair *pairs, const AffOut *outs, const AffOut *score_outs,
                     const int32_t *band_j0, const TbJob *jobs, int count) {
    if (!count) return 0;
    const int threads = 64, blocks = (count + threads - 1) / threads;
    switch (b->R) {
    case 8:
        affine_traceback_kernel<8><<<blocks, threads, 0, b->e->stream>>>(
            pairs, outs, score_outs, band_j0, jobs, count, b->aff, b->d_results.p, b->d_ops_slots.p);
        break;
    case 16:
        affine_traceback_kernel<16><<<blocks, threads, 0, b->e->stream>>>(
            pairs, outs, score_outs, band_j0, jobs, count, b->aff, b->d_results.p, b->d_ops_slots.p);
        break;
    default:
        affine_traceback_kernel<32><<<blocks, threads, 0, b->e->stream>>>(
            pairs, outs, score_outs, band_j0, jobs, count, b->aff, b->d_results.p, b->d_ops_slots.p);
        break;
    }
    C4B_CUDA(cudaGetLastError());
    b->e->launches++;
    return 0;
}

size_t tb_words(int R, int Q, int T) {
    const int rows_per_sweep = 32 * R;
    const size_t nsweeps = (size_t)(Q + 1 + rows_per_sweep - 1) / rows_per_sweep;
    return nsweeps * (size_t)(T + 1 + 31) * 32 * (R / 8);
}

int affine_create(c4b_batch *b, const c4b_pair *pairs, int match_kind) {
    c4b_engine *e = b->e;
    const int n = b->n;
    const c4b_scoring &sc = b->scoring;
    const int32_t *matrix = (match_kind == C4B_CALC_MATCH_DNA) ? sc.dna_matrix : sc.protein_matrix;
    const uint8_t *index = (match_kind == C4B_CALC_MATCH_DNA) ? sc.dna_index : sc.protein_index;
    typedef std::pair<const uint8_t *, int> SeqKey;

    // ---- validation, alphabet use, lattice geometry
    int maxQ = 0;
    bool used[24] = {false};
    std::vector<uint8_t> query_wide(n, 0);  // the query holds a symbol outside the primary four
    {
        std::set<SeqKey> seen;
        std::vector<SeqKey> distinct_q;
        for (int p = 0; p < n; ++p) {
            const c4b_pair &pp = pairs[p];
            if (pp.query_length < 0 || pp.target_length < 0 || pp.query_start < 0 || pp.target_start < 0 ||
                (int64_t)pp.query_start + pp.query_length > pp.query_len ||
                (int64_t)pp.target_start + pp.target_length > pp.target_len) {
                set_error("pair " + std::to_string(p) + ": region outside the sequences");
                return -1;
            }
            if (pp.n_blocked < 0 || (pp.n_blocked && (!pp.blocked_query_pos || !pp.blocked_target_pos))) {
                set_error("pair " + std::to_string(p) + ": bad SubOpt blocked-cell list");
                return -1;
            }
            b->any_blocked = b->any_blocked || pp.n_blocked > 0;
            const SeqKey qk(pp.query + pp.query_start, pp.query_length);
            if (seen.insert(qk).second) distinct_q.push_back(qk);
            maxQ = std::max(maxQ, pp.query_length);
            b->cells += (int64_t)pp.query_length * pp.target_length;
        }
        // which matrix rows do the queries use (decides PRMT classes / packed16); the
        // scan is the only per-byte host work before staging starts, so it is threaded
        size_t qtotal = 0;
        for (const SeqKey &k : distinct_q) qtotal += (size_t)k.second;
        const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>(
            host_threads(), qtotal >> 20));
        std::vector<std::array<bool, 256>> seen_byte(nt);
        // "narrow" queries hold only the four primary symbols (A, C, G, T of a DNA matrix):
        // those get PRMT classes 0..3, which is what the packed kernels' 4-byte pools hold
        bool prim[256];
        for (int c = 0; c < 256; ++c)
            prim[c] = match_kind == C4B_CALC_MATCH_DNA &&
                      (index[c] == index['A'] || index[c] == index['C'] || index[c] == index['G'] || index[c] == index['T']);
        std::vector<uint8_t> wide(distinct_q.size(), 0);
        auto scan = [&](unsigned t) {
            std::array<bool, 256> &sb = seen_byte[t];
            sb.fill(false);
            for (size_t k = t; k < distinct_q.size(); k += nt) {
                const uint8_t *q = distinct_q[k].first;
                bool w = false;
                for (int i = 0; i < distinct_q[k].second; ++i) {
                    sb[q[i]] = true;
                    w |= !prim[q[i]];
                }
                wide[k] = w;
            }
        };
        {
            std::vector<std::thread> th;
            for (unsigned t = 1; t < nt; ++t) th.emplace_back(scan, t);
            scan(0);
            for (auto &x : th) x.join();
        }
        for (unsigned t = 0; t < nt; ++t)
            for (int c = 0; c < 256; ++c)
                if (seen_byte[t][c]) {
                    if (index[c] >= 24) {
                        set_error("a query holds a symbol outside the substitution matrix");
                        return -1;
                    }
                    used[index[c]] = true;
                }
        {
            std::map<SeqKey, int> slot;
            for (size_t k = 0; k < distinct_q.size(); ++k) slot[distinct_q[k]] = (int)k;
            for (int p = 0; p < n; ++p)
                query_wide[p] = wide[slot[SeqKey(pairs[p].query + pairs[p].query_start, pairs[p].query_length)]];
        }
    }
    tmark("create: validated, alphabet scanned");
    b->R = (maxQ + 1 > 512) ? 32 : (maxQ + 1 > 256 ? 16 : 8);
    if (const char *env = getenv("C4B_AFFINE_R")) {  // tuning override: rows per lane
        const int r = atoi(env);
        if (r == 8 || r == 16 || r == 32) b->R = r;
    }
    // The int32 score pass (proteins, N-rich queries, global / bestfit / overlap scopes) takes the smallest
    // multiple of 4 rows per lane that holds the longest query in ONE sweep, like the packed score pass
    // (a 300-residue protein: 12 rows per lane instead of 16); the recording pass keeps R (records are
    // groups of 8 rows), and so do batches with SubOpt lists (their entries are laid out per R rows).
    b->R32s = b->R;
    if (!b->any_blocked && !getenv("C4B_AFFINE_R") && maxQ + 1 <= 32 * b->R)
        b->R32s = std::min(b->R, std::max(8, ((maxQ + 1 + 31) / 32 + 3) / 4 * 4));
    b->fill_warps = std::max(1, std::min(kAffMaxWarps, (maxQ + 1 + 32 * b->R - 1) / (32 * b->R)));
    if (const char *env = getenv("C4B_AFFINE_WARPS")) b->fill_warps = std::max(1, std::min(kAffMaxWarps, atoi(env)));

    // ---- scoring tables
    int n_used = 0, cls_of[24], code_of[8];
    bool fits8 = true;
    int max_sub = INT32_MIN;
    for (int a = 0; a < 24; ++a) {
        cls_of[a] = -1;
        for (int c = 0; c < 24; ++c) {
            const int v = matrix[a * 24 + c];
            max_sub = std::max(max_sub, v);
            if (used[a] && (v - b->aff.openD < -127 || v - b->aff.openD > 127)) fits8 = false;
        }
    }
    // PRMT classes: the primary four first (when used), then the other used rows
    {
        bool primary_row[24] = {false};
        if (match_kind == C4B_CALC_MATCH_DNA)
            for (char ch : {'A', 'C', 'G', 'T'})
                if (index[(uint8_t)ch] < 24) primary_row[index[(uint8_t)ch]] = true;
        for (int pass = 0; pass < 2; ++pass)
            for (int a = 0; a < 24; ++a)
                if (used[a] && primary_row[a] == (pass == 0)) {
                    if (n_used < 7) code_of[n_used] = a;
                    cls_of[a] = n_used++;
                }
    }

    b->max_sub = max_sub;
    b->gap_min = std::min(-b->aff.openD, -b->aff.extD);
    b->score_mode = (n_used <= 7 && fits8) ? SCORE_PRMT : SCORE_SMEM;
    b->aff.score_mode = b->score_mode;
    std::vector<uint8_t> lut(512, 0xFF);
    std::vector<uint8_t> table;
    if (b->score_mode == SCORE_PRMT) {
        table.assign(25 * 8, 0);
        for (int tc = 0; tc < 25; ++tc) {
            int8_t *x = reinterpret_cast<int8_t *>(&table[tc * 8]);
            // entries are s - gap_open (the kernel carries M + gap_open)
            for (int k = 0; k < n_used; ++k)
                x[k] = (tc < 24) ? (int8_t)(matrix[code_of[k] * 24 + tc] - b->aff.openD) : 0;
            x[kPadClass] = (int8_t)(-100 - b->aff.openD);
        }
        for (int c = 0; c < 256; ++c)
            if (index[c] < 24 && cls_of[index[c]] >= 0) lut[c] = (uint8_t)cls_of[index[c]];
    } else {
        std::vector<int32_t> t32(25 * 25, 0);
        for (int a = 0; a < 25; ++a)
            for (int c = 0; c < 25; ++c)
                t32[a * 25 + c] = ((a < 24 && c < 24) ? matrix[a * 24 + c] : (a == 24 ? -100 : 0)) - b->aff.openD;
        table.resize(t32.size() * 4);
        memcpy(table.data(), t32.data(), table.size());
        for (int c = 0; c < 256; ++c)
            if (index[c] < 24) lut[c] = index[c];
    }
    for (int c = 0; c < 256; ++c)
        if (index[c] < 24) lut[256 + c] = index[c];

    // ---- which lattices take which route
    const bool local = (b->aff.start_scope == C4B_SCOPE_ANYWHERE && b->aff.end_scope == C4B_SCOPE_ANYWHERE);
    std::vector<int64_t> band_cols(n, 0);
    for (int p = 0; p < n; ++p) {
        const int Q = pairs[p].query_length, T = pairs[p].target_length;
        bool two_pass = !b->want_path;
        if (b->want_path && local) {
            const int64_t W = std::min<int64_t>(T, affine_band_width(Q, b->max_sub, b->gap_min));
            band_cols[p] = W;
            two_pass = (2 * (W + 33) < 3 * ((int64_t)T + 33) / 2);  // band < 75% of the lattice
        }
        if (two_pass) b->score_list.push_back(p);
        else b->direct_list.push_back(p);
    }
    auto by_cost = [&](int a, int c) {
        const int64_t ca = (int64_t)pairs[a].query_length * pairs[a].target_length;
        const int64_t cc = (int64_t)pairs[c].query_length * pairs[c].target_length;
        return ca != cc ? ca > cc : a < c;
    };
    std::sort(b->score_list.begin(), b->score_list.end(), by_cost);
    std::sort(b->direct_list.begin(), b->direct_list.end(), by_cost);
    // Packed 16-bit score pass (affine_packed16.cuh): exact whenever no halfword add
    // can wrap.  Eligible lattices go first in score_list, still cost-descending.
    {
        const char *env = getenv("C4B_AFFINE_PACK16");
        const bool allow = !(env && atoi(env) == 0);
        const bool model_ok = allow && local && b->score_mode == SCORE_PRMT && max_sub > 0 &&
                              b->aff.openD < 0 && b->aff.openD > -1000 && b->aff.extD < 0 && b->aff.extD > -1000 &&
                              b->aff.extI < 0 && b->aff.extI > -1000;
        // the offset-binary variant adds score' = s - open as an unsigned halfword and
        // shortens the I chain with open <= extend; it also sweeps queries longer than 32 R rows
        bool nonneg = true;
        for (int a = 0; a < 24; ++a)
            for (int c = 0; c < 24 && used[a]; ++c) nonneg = nonneg && matrix[a * 24 + c] >= b->aff.openD;
        const char *v = getenv("C4B_P16_VARIANT");
        b->p16_unsigned = nonneg && b->aff.openI <= b->aff.extI && !(v && v[0] == 's');
        auto fits16 = [&](int p) {
            const int64_t Q = pairs[p].query_length, T = pairs[p].target_length;
            // per lattice: a query of primary symbols only (classes 0..3), values in 15 bits
            // (the signed variant: one sweep only)
            // (lattices with SubOpt blocked cells take the int32 kernel, which has the masked variant)
            return model_ok && !query_wide[p] && !pairs[p].n_blocked && (b->p16_unsigned || Q + 1 <= 32 * b->R) &&
                   (int64_t)max_sub * (std::min(Q, T) + 1) <= 32000;
        };
        auto mid = std::stable_partition(b->score_list.begin(), b->score_list.end(), fits16);
        b->n16 = (int)(mid - b->score_list.begin());
        // Rows per lane of the packed score pass.  Fewer rows per lane (a 1 kbp query as 2 or 4
        // pipelined sweeps on the warps of affine_fill16u_multi_kernel) would put more warps on a
        // GPU that a small batch leaves under-filled, but measured on the B200 it never pays:
        // 1250 pairs 2301 / 2299 / 1663 GCUPS at 32 / 16 / 8 rows per lane, 2500 pairs 3226 /
        // 2833 / 2135 (profiles/r02_strong_sweep.md) -- the per-step hand-off through L2 costs more
        // than the extra warps bring.  So R16 = R unless C4B_P16_R says otherwise (tuning aid).
        b->R16 = b->R;
        int maxQ16 = 0;
        for (int k = 0; k < b->n16; ++k) maxQ16 = std::max(maxQ16, pairs[b->score_list[k]].query_length);
        if (b->p16_unsigned && b->n16 > 0) {
            if (const char *env = getenv("C4B_P16_R")) {
                const int r = atoi(env);
                if ((r == 8 || r == 16 || r == 32) && r <= b->R) b->R16 = r;
            }
        }
        // The one-sweep score pass takes the smallest multiple of 4 rows per lane that holds the longest
        // packed query (the traceback keeps R in {8, 16, 32}: its records are groups of 8 rows): a
        // 600-symbol query runs 20 rows per lane instead of 32 -- rows beyond the query are padding that
        // is computed like any other row.
        if (b->p16_unsigned && b->n16 > 0 && !getenv("C4B_P16_R") && maxQ16 + 1 <= 32 * b->R16)
        {
            const int need = (maxQ16 + 1 + 31) / 32;
            b->R16 = need <= 8 ? std::max(4, (need + 1) / 2 * 2) : (need + 3) / 4 * 4;   // 4, 6, 8, 12, 16, .. 32
        }
        // ... and a query of several sweeps keeps its sweep count but takes the fewest rows per lane that
        // still cover it (2049 rows: 3 sweeps of 24 rows per lane instead of 32)
        if (b->p16_unsigned && b->n16 > 0 && !getenv("C4B_P16_R") && b->R16 == 32 && maxQ16 + 1 > 32 * 32) {
            const int sweeps = (maxQ16 + 1 + 1023) / 1024;
            const int need = (maxQ16 + 1 + 32 * sweeps - 1) / (32 * sweeps);
            b->R16 = std::min(32, std::max(16, (need + 3) / 4 * 4));
        }
        b->fill_warps16 = std::max(1, std::min(kAffMaxWarps, (maxQ16 + 1 + 32 * b->R16 - 1) / (32 * b->R16)));
        b->p16_multi = b->p16_unsigned && maxQ16 + 1 > 32 * b->R16;
        // fold: one warp per lattice instead of one per PAIR while the batch is small enough that
        // the finer grain wins (fewer empty warp slots, a smaller tail wave).  Measured on the B200
        // (profiles/r02_strong_sweep.md, 1 kbp x 100 kbp, score pass GCUPS packed -> folded):
        // 1250 lattices 2302 -> 2847, 2500 3227 -> 3525, 3552 3646 -> 4135, 5000 3352 -> 3860,
        // 7104 3678 -> 4020, 10000 4417 -> 4282: folded up to 2.4 full waves of packed CTAs.
        // (queries of up to 511 rows fold to 8 rows per lane, where the per-step overhead weighs more:
        // folded only below one packed wave -- 512 x 100 kbp, 7812 lattices: 2293 packed vs 2084 folded)
        b->p16_fold = b->p16_unsigned && !b->p16_multi && b->R >= 16 && b->n16 > 0 &&
                      (int64_t)(b->n16 + 1) / 2 * 10 < (int64_t)e->sm_count * 12 * (b->R == 32 ? 24 : 10);
        if (const char *env = getenv("C4B_P16_FOLD"))
            b->p16_fold = b->p16_unsigned && !b->p16_multi && b->R >= 16 && b->n16 > 0 && atoi(env) != 0;
        // the folded kernel's halves hold 32 Rf rows each: the smallest even Rf that takes the longest query
        b->Rf = std::min(b->R / 2, std::max(6, ((maxQ16 + 1 + 63) / 64 + 1) / 2 * 2));
        if (getenv("C4B_P16_R")) b->Rf = b->R / 2;
        // packed traceback pass (tagged unsigned halfwords, 8 * value + 1024): whole lists only
        const char *tv = getenv("C4B_AFFINE_TB16");
        const bool tb_ok = model_ok && nonneg && b->aff.openI <= b->aff.extI && b->aff.openD >= -24 &&
                           b->aff.extD >= -24 && b->aff.extI >= -24 && b->aff.extD == b->aff.extI &&
                           !(tv && atoi(tv) == 0);
        auto fits_tb16 = [&](int p) {
            const int64_t Q = pairs[p].query_length, T = pairs[p].target_length;
            return !query_wide[p] && !pairs[p].n_blocked && Q + 1 <= 32 * b->R &&
                   8 * (int64_t)max_sub * (std::min(Q, T) + 1) + 2048 < 65000;
        };
        b->tb16_band = tb_ok && b->want_path && !b->score_list.empty() && b->n16 == (int)b->score_list.size() &&
                       std::all_of(b->score_list.begin(), b->score_list.end(), fits_tb16);
        b->tb16_direct = tb_ok && b->want_path && !b->direct_list.empty() &&
                         std::all_of(b->direct_list.begin(), b->direct_list.end(), fits_tb16);
    }
    const int ns = (int)b->score_list.size(), nd = (int)b->direct_list.size();

    // ---- sequence placement: every distinct host buffer once (dedupe by pointer +
    // length), IN LAUNCH ORDER (score_list, then direct_list), so that the lattices
    // at the head of the launch order are the first to be resident
    std::map<SeqKey, size_t> qmap, tmap;
    std::vector<size_t> qoff(n), toff(n);
    size_t qbytes = 0, tbytes = 0;
    struct CopyJob { size_t dst; const uint8_t *src; size_t len, slot; uint8_t fill; };
    std::vector<CopyJob> qjobs, tjobs;
    auto place = [&](int p) {
        const c4b_pair &pp = pairs[p];
        const SeqKey qk(pp.query + pp.query_start, pp.query_length);
        auto it = qmap.find(qk);
        if (it == qmap.end()) {
            qmap[qk] = qbytes;
            qoff[p] = qbytes;
            const size_t slot = align_up((size_t)pp.query_length, 16) + 16;
            qjobs.push_back({qbytes, qk.first, (size_t)qk.second, slot, 0});
            qbytes += slot;
        } else {
            qoff[p] = it->second;
        }
        const SeqKey tk(pp.target + pp.target_start, pp.target_length);
        auto jt = tmap.find(tk);
        if (jt == tmap.end()) {
            tmap[tk] = tbytes;
            toff[p] = tbytes;
            const size_t slot = align_up((size_t)pp.target_length, 16) + 16;
            tjobs.push_back({tbytes, tk.first, (size_t)tk.second, slot, 0});
            tbytes += slot;
        } else {
            toff[p] = jt->second;
        }
    };
    for (int k = 0; k < ns; ++k) place(b->score_list[k]);
    for (int k = 0; k < nd; ++k) place(b->direct_list[k]);
    b->q_bytes = qbytes;
    b->t_bytes = tbytes;
    // alignment gaps between sequences are encoded too: fill them with a symbol
    // of the alphabet so only real sequence bytes can raise the "bad symbol" flag
    uint8_t qfill = 0, tfill = 0;
    for (int c = 255; c >= 0; --c) {
        if (lut[c] != 0xFF) qfill = (uint8_t)c;
        if (lut[256 + c] != 0xFF) tfill = (uint8_t)c;
    }
    std::vector<CopyJob> jobs;
    jobs.reserve(qjobs.size() + tjobs.size());
    for (CopyJob j : qjobs) { j.fill = qfill; jobs.push_back(j); }
    for (CopyJob j : tjobs) { j.dst += qbytes; j.fill = tfill; jobs.push_back(j); }  // already in dst order
    // staging slices: [jobs j0, j1) -> device bytes [lo, hi)
    struct Slice { size_t j0, j1, lo, hi; };
    std::vector<Slice> slices;
    const size_t stage_bytes = qbytes + tbytes + 64;
    {
        // slice sizes grow geometrically (1/64, 1/32, 1/16 of the batch, then 1/8 each):
        // the fill starts after the first small slice, later slices amortise the launch
        size_t unit = std::max<size_t>(stage_bytes / 64, 4u << 20), cap = std::max<size_t>(stage_bytes / 8, 32u << 20);
        if (const char *env = getenv("C4B_STAGE_SLICE_KB")) unit = cap = std::max(1, atoi(env)) * (size_t)1024;
        size_t j0 = 0, slice_target = unit;
        while (j0 < jobs.size()) {
            size_t j1 = j0, bytes = 0;
            while (j1 < jobs.size() && bytes < slice_target) bytes += jobs[j1++].slot;
            slices.push_back({j0, j1, jobs[j0].dst, (j1 < jobs.size()) ? jobs[j1].dst : stage_bytes});
            j0 = j1;
            slice_target = std::min(cap, slice_target * 2);
        }
        if (slices.empty()) slices.push_back({0, 0, 0, stage_bytes});
    }
    auto slice_of = [&](size_t dst) {
        int lo = 0, hi = (int)slices.size() - 1;
        while (lo < hi) {
            const int mid = (lo + hi) / 2;
            if (dst < slices[mid].hi) hi = mid; else lo = mid + 1;
        }
        return lo;
    };
    // pass-1 launch groups: maximal runs of score_list with the same (running max)
    // resident slice; boundaries even, because the packed kernel pairs neighbours
    {
        int begin = 0, cur = 0, min_group = 256;
        if (const char *env = getenv("C4B_P1_MIN_GROUP")) min_group = std::max(2, atoi(env));
        for (int k = 0; k < ns; ++k) {
            const int p = b->score_list[k];
            const int ready = std::max(slice_of(qoff[p]), slice_of(qbytes + toff[p]));
            if (ready > cur && (k & 1) == 0 && k - begin >= min_group) {
                b->groups.push_back({begin, k, cur, nullptr});
                begin = k;
            }
            cur = std::max(cur, ready);
        }
        if (begin < ns) b->groups.push_back({begin, ns, cur, nullptr});
    }

    // ---- traceback arena, chunked to a memory budget
    size_t free_b = 0, total_b = 0;
    C4B_CUDA(cudaMemGetInfo(&free_b, &total_b));
    size_t fixed = qbytes + tbytes;
    const size_t tb_budget_words = (free_b > fixed + (1ull << 30) ? (free_b - fixed - (1ull << 30)) : (1ull << 28)) / 4 / 2;
    std::vector<size_t> tb_off_s(ns, 0), tb_off_d(nd, 0);
    size_t arena_words = 0;
    if (b->want_path) {
        auto chunkify = [&](const std::vector<int> &list, std::vector<size_t> &offs,
                            std::vector<Chunk> &chunks, bool band) -> int {
            size_t cur = 0;
            int begin = 0;
            for (int k = 0; k < (int)list.size(); ++k) {
                const int p = list[k];
                const int Tt = band ? (int)band_cols[p] : pairs[p].target_length;
                const size_t w = align_up(tb_words(b->R, pairs[p].query_length, Tt), 4);
                if (w > tb_budget_words) {
                    set_error("traceback of pair " + std::to_string(p) + " exceeds the device memory budget");
                    return -1;
                }
                if (cur + w > tb_budget_words) {
                    chunks.push_back({begin, k});
                    begin = k;
                    cur = 0;
                }
                offs[k] = cur;
                cur += w;
                arena_words = std::max(arena_words, cur);
            }
            if (begin < (int)list.size()) chunks.push_back({begin, (int)list.size()});
            return 0;
        };
        if (chunkify(b->score_list, tb_off_s, b->band_chunks, true)) return -1;
        if (chunkify(b->direct_list, tb_off_d, b->direct_chunks, false)) return -1;
    }

    tmark("create: lists, placement, chunks planned");
    // ---- device allocations (stream-ordered on the engine stream)
    if (b->d_seq.alloc(stage_bytes)) return -1;
    if (b->d_lut.alloc(512)) return -1;
    if (b->d_score_table.alloc(table.size())) return -1;
    if (b->d_bad.alloc(1)) return -1;
    if (b->d_full.alloc(ns) || b->d_out1.alloc(ns)) return -1;
    if (b->d_direct.alloc(nd) || b->d_outd.alloc(nd)) return -1;
    if (b->d_results.alloc(n)) return -1;
    if (b->d_qorg.alloc(ns) || b->d_torg.alloc(ns)) return -1;
    if (b->want_path) {
        if (b->d_band.alloc(ns) || b->d_out2.alloc(ns) || b->d_band_j0.alloc(ns)) return -1;
        if (b->d_jobs_band.alloc(ns) || b->d_jobs_direct.alloc(nd)) return -1;
        if (b->d_tb.alloc(arena_words + 4)) return -1;
        if (b->d_new_off.alloc((size_t)n + 1)) return -1;
    }
    // sweep hand-off rows for queries longer than one sweep
    // (the packed score pass keeps the hand-off row of a PAIR of lattices in the first one's
    // buffers: they are sized for the longer target / needed if either query is long)
    std::vector<size_t> top_off(n, (size_t)-1), top_len(n, 0);
    for (int p = 0; p < n; ++p)
        if (pairs[p].query_length + 1 > 32 * b->R) top_len[p] = (size_t)pairs[p].target_length + 1;
    for (int k = 0; k < b->n16; k += 2) {
        const int pa = b->score_list[k], pb = b->score_list[std::min(k + 1, b->n16 - 1)];
        if (std::max(pairs[pa].query_length, pairs[pb].query_length) + 1 > 32 * b->R16)
            top_len[pa] = std::max(top_len[pa], (size_t)std::max(pairs[pa].target_length, pairs[pb].target_length) + 1);
    }
    size_t top_elems = 0;
    for (int p = 0; p < n; ++p)
        if (top_len[p]) {
            top_off[p] = top_elems;
            top_elems += 2 * top_len[p];
        }
    if (b->d_top.alloc(top_elems)) return -1;

    // ---- SubOpt blocked cells (src/c4/subopt.c:250-338: region coordinates of the DESTINATION
    // cell, sorted by target then query position) -> per lane strip {column, row mask}, by column
    std::vector<int2> h_blk;
    std::vector<int32_t> h_blk_off;
    std::vector<size_t> blk_seg(n, (size_t)-1);
    if (b->any_blocked) {
        std::vector<std::array<int32_t, 3>> cells;   // (strip, column, row in strip)
        for (int p = 0; p < n; ++p) {
            const c4b_pair &pp = pairs[p];
            if (!pp.n_blocked) continue;
            const int nstrips = 32 * ((pp.query_length + 1 + 32 * b->R - 1) / (32 * b->R));
            cells.clear();
            for (int k = 0; k < pp.n_blocked; ++k) {
                const int i = pp.blocked_query_pos[k], j = pp.blocked_target_pos[k];
                if (i < 0 || i > pp.query_length || j < 0 || j > pp.target_length) continue;  // never looked at
                cells.push_back({i / b->R, j, i % b->R});
            }
            std::sort(cells.begin(), cells.end());
            blk_seg[p] = h_blk_off.size();
            size_t c = 0;
            for (int strip = 0; strip < nstrips; ++strip) {
                h_blk_off.push_back((int32_t)h_blk.size());
                while (c < cells.size() && cells[c][0] == strip) {
                    const int j = cells[c][1];
                    uint32_t mask = 0;
                    for (; c < cells.size() && cells[c][0] == strip && cells[c][1] == j; ++c) mask |= 1u << cells[c][2];
                    h_blk.push_back(make_int2(j, (int)mask));
                }
            }
            h_blk_off.push_back((int32_t)h_blk.size());
            if (h_blk.size() > (size_t)INT32_MAX / 2) {
                set_error("too many SubOpt blocked cells in one batch");
                return -1;
            }
        }
        if (b->d_blk.alloc(h_blk.size() + 1) || b->d_blk_off.alloc(h_blk_off.size() + 1)) return -1;
    }

    // ---- tables, lattice descriptors and traceback jobs go up first (small)
    cudaStream_t st = e->stream;
    if (!h_blk_off.empty()) {
        if (!h_blk.empty())
            C4B_CUDA(cudaMemcpyAsync(b->d_blk.p, h_blk.data(), h_blk.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
        C4B_CUDA(cudaMemcpyAsync(b->d_blk_off.p, h_blk_off.data(), h_blk_off.size() * sizeof(int32_t),
                                 cudaMemcpyHostToDevice, st));
    }
    C4B_CUDA(cudaMemcpyAsync(b->d_lut.p, lut.data(), 512, cudaMemcpyHostToDevice, st));
    C4B_CUDA(cudaMemcpyAsync(b->d_score_table.p, table.data(), table.size(), cudaMemcpyHostToDevice, st));
    C4B_CUDA(cudaMemsetAsync(b->d_bad.p, 0, sizeof(int), st));
    std::vector<AffPair> h_full(ns), h_direct(nd);
    std::vector<int32_t> h_qorg(ns), h_torg(ns);
    std::vector<TbJob> h_jb(ns), h_jd(nd);
    int64_t ops_cursor = 0;
    auto make_pair_desc = [&](int p, int slot, size_t tbo) {
        AffPair a;
        a.q = b->d_seq.p + qoff[p];
        a.t = b->d_seq.p + qbytes + toff[p];
        a.Q = pairs[p].query_length;
        a.T = pairs[p].target_length;
        a.tb = b->want_path ? b->d_tb.p + tbo : nullptr;
        a.top0 = a.top1 = nullptr;
        if (top_off[p] != (size_t)-1) {
            a.top0 = b->d_top.p + top_off[p];
            a.top1 = a.top0 + top_len[p];
        }
        a.out_index = slot;
        a.blk = b->d_blk.p;
        a.blk_off = (blk_seg[p] != (size_t)-1) ? b->d_blk_off.p + blk_seg[p] : nullptr;
        a.blk_j0 = 0;
        a.blk_pad = 0;
        return a;
    };
    auto make_job = [&](int p, int slot, bool band) {
        TbJob j;
        j.pair = slot;
        j.result = p;
        j.q_origin = pairs[p].query_start;
        j.t_origin = pairs[p].target_start;
        j.expect = band ? 1 : 0;
        j.score_slot = band ? slot : -1;
        const int64_t span = band ? band_cols[p] : pairs[p].target_length;
        j.ops_cap = (int32_t)std::min<int64_t>(pairs[p].query_length + span + 4, INT32_MAX);
        j.ops_off = ops_cursor;
        j.reserved = 1;   // every affine traceback pass records nibbles in tag format
        ops_cursor += j.ops_cap;
        return j;
    };
    for (int k = 0; k < ns; ++k) {
        const int p = b->score_list[k];
        h_full[k] = make_pair_desc(p, k, tb_off_s[k]);
        h_qorg[k] = pairs[p].query_start;
        h_torg[k] = pairs[p].target_start;
        if (b->want_path) h_jb[k] = make_job(p, k, true);
    }
    for (int k = 0; k < nd; ++k) {
        const int p = b->direct_list[k];
        h_direct[k] = make_pair_desc(p, k, tb_off_d[k]);
        if (b->want_path) h_jd[k] = make_job(p, k, false);
    }
    if (b->want_path) {
        if (b->d_ops_slots.alloc(2 * (size_t)ops_cursor + 2)) return -1;
        if (b->d_ops_packed.alloc(2 * (size_t)ops_cursor + 2)) return -1;
    }
    if (ns) {
        C4B_CUDA(cudaMemcpyAsync(b->d_full.p, h_full.data(), ns * sizeof(AffPair), cudaMemcpyHostToDevice, st));
        C4B_CUDA(cudaMemcpyAsync(b->d_qorg.p, h_qorg.data(), ns * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        C4B_CUDA(cudaMemcpyAsync(b->d_torg.p, h_torg.data(), ns * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        if (b->want_path)
            C4B_CUDA(cudaMemcpyAsync(b->d_jobs_band.p, h_jb.data(), ns * sizeof(TbJob), cudaMemcpyHostToDevice, st));
    }
    if (nd) {
        C4B_CUDA(cudaMemcpyAsync(b->d_direct.p, h_direct.data(), nd * sizeof(AffPair), cudaMemcpyHostToDevice, st));
        if (b->want_path)
            C4B_CUDA(cudaMemcpyAsync(b->d_jobs_direct.p, h_jd.data(), nd * sizeof(TbJob), cudaMemcpyHostToDevice, st));
    }
    // the pageable sources above are consumed before cudaMemcpyAsync returns
    for (cudaEvent_t *ev : {&b->ev_base, &b->ev_p1a, &b->ev_p1b}) C4B_CUDA(cudaEventCreate(ev));
    for (auto &g : b->groups) C4B_CUDA(cudaEventCreateWithFlags(&g.done, cudaEventDisableTiming));
    b->slice_events.assign(slices.size(), nullptr);
    for (auto &ev : b->slice_events) C4B_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    C4B_CUDA(cudaEventRecord(b->ev_p1a, st));
    C4B_CUDA(cudaEventRecord(b->ev_base, st));  // allocations + tables + descriptors are in place

    tmark("create: allocations + descriptors uploaded");
    // ---- stage sequences: pinned bounce buffer -> HBM -> encode in place, slice by
    // slice on the copy stream.  Worker threads fill the bounce buffer (grow-only,
    // owned by the engine); as soon as a slice is issued, the score pass of every
    // launch group whose sequences it completes is queued on an aux stream, so host
    // copy, DMA and the fill overlap.
    cudaStream_t cs = e->copy_stream;
    C4B_CUDA(cudaStreamWaitEvent(cs, b->ev_base, 0));
    // Callers whose sequence buffers are page-locked (C4B_PAIR_BUFFERS_PINNED on every pair) skip
    // the bounce buffer: the DMA reads their memory directly.  Runs of equally long sequences at a
    // constant stride (rows of one array, the usual case) go as ONE 2-D copy each.
    bool pinned = n > 0;
    for (int p = 0; p < n && pinned; ++p) pinned = (pairs[p].reserved & C4B_PAIR_BUFFERS_PINNED) != 0;
    uint8_t *h_seq = nullptr;
    if (!pinned) {
        if (e->stage_free) C4B_CUDA(cudaEventSynchronize(e->stage_free));  // previous batch done reading
        if (e->h_stage_cap < stage_bytes) {
            if (e->h_stage) cudaFreeHost(e->h_stage);
            e->h_stage = nullptr;
            e->h_stage_cap = 0;
            C4B_CUDA(cudaMallocHost(&e->h_stage, stage_bytes + stage_bytes / 8));
            e->h_stage_cap = stage_bytes + stage_bytes / 8;
        }
        h_seq = e->h_stage;
        memset(h_seq + qbytes + tbytes, tfill, 64);
    } else {
        C4B_CUDA(cudaMemsetAsync(b->d_seq.p + qbytes + tbytes, tfill, 64, cs));
    }
    const unsigned hw = host_threads();
    size_t next_group = 0;
    for (size_t si = 0; si < slices.size(); ++si) {
        const Slice &S = slices[si];
        const size_t bytes = S.hi - S.lo;
        if (pinned) {
            // slot padding first (the encode kernels read whole slots), then the sequences
            const size_t qhi_ = std::min(S.hi, qbytes), tlo_ = std::max(S.lo, qbytes);
            if (S.lo < qhi_) C4B_CUDA(cudaMemsetAsync(b->d_seq.p + S.lo, qfill, qhi_ - S.lo, cs));
            if (tlo_ < S.hi) C4B_CUDA(cudaMemsetAsync(b->d_seq.p + tlo_, tfill, S.hi - tlo_, cs));
            for (size_t k = S.j0; k < S.j1;) {
                size_t run = 1;
                if (k + 1 < S.j1 && jobs[k + 1].len == jobs[k].len && jobs[k + 1].slot == jobs[k].slot &&
                    jobs[k + 1].src > jobs[k].src) {
                    const size_t stride = (size_t)(jobs[k + 1].src - jobs[k].src);
                    while (k + run < S.j1 && jobs[k + run].len == jobs[k].len && jobs[k + run].slot == jobs[k].slot &&
                           jobs[k + run].src == jobs[k].src + run * stride && jobs[k + run].dst == jobs[k].dst + run * jobs[k].slot)
                        ++run;
                    if (run > 1 && stride >= jobs[k].len && jobs[k].len > 0)
                        C4B_CUDA(cudaMemcpy2DAsync(b->d_seq.p + jobs[k].dst, jobs[k].slot, jobs[k].src, stride, jobs[k].len,
                                                   run, cudaMemcpyHostToDevice, cs));
                    else
                        run = 1;
                }
                if (run == 1 && jobs[k].len)
                    C4B_CUDA(cudaMemcpyAsync(b->d_seq.p + jobs[k].dst, jobs[k].src, jobs[k].len, cudaMemcpyHostToDevice, cs));
                k += run;
            }
        } else {
            const unsigned nt = (unsigned)std::min<size_t>(hw, std::max<size_t>(1, bytes >> 22));
            auto work = [&](unsigned t) {
                for (size_t k = S.j0 + t; k < S.j1; k += nt) {
                    const CopyJob &c = jobs[k];
                    memcpy(h_seq + c.dst, c.src, c.len);
                    memset(h_seq + c.dst + c.len, c.fill, c.slot - c.len);
                }
            };
            if (nt <= 1) {
                work(0);
            } else {
                std::vector<std::thread> th;
                for (unsigned t = 1; t < nt; ++t) th.emplace_back(work, t);
                work(0);
                for (auto &x : th) x.join();
            }
            C4B_CUDA(cudaMemcpyAsync(b->d_seq.p + S.lo, h_seq + S.lo, bytes, cudaMemcpyHostToDevice, cs));
        }
        // encode [lo, hi): the part inside the query region with the query LUT, the rest
        // with the target LUT (slots are multiples of 16 bytes, so is every boundary)
        const size_t qhi = std::min(S.hi, qbytes);
        if (S.lo < qhi) {
            encode_kernel<<<(unsigned)(((qhi - S.lo) / 16 + 255) / 256 + 1), 256, 0, cs>>>(
                b->d_seq.p + S.lo, qhi - S.lo, b->d_lut.p, b->d_bad.p, 0u);
            e->launches++;
        }
        const size_t tlo = std::max(S.lo, qbytes);
        if (tlo < S.hi) {
            encode_kernel<<<(unsigned)(((S.hi - tlo) / 16 + 255) / 256 + 1), 256, 0, cs>>>(
                b->d_seq.p + tlo, S.hi - tlo, b->d_lut.p + 256, b->d_bad.p, (uint32_t)kTargetNone);
            e->launches++;
        }
        C4B_CUDA(cudaGetLastError());
        C4B_CUDA(cudaEventRecord(b->slice_events[si], cs));
        while (next_group < b->groups.size() && b->groups[next_group].slice <= (int)si) {
            if (launch_pass1_group(b, (int)next_group)) return -1;
            ++next_group;
        }
    }
    while (next_group < b->groups.size()) {
        if (launch_pass1_group(b, (int)next_group)) return -1;
        ++next_group;
    }
    b->pass1_inflight = true;
    tmark("create: all slices staged, pass 1 queued");
    if (!pinned) {
        if (!e->stage_free) C4B_CUDA(cudaEventCreateWithFlags(&e->stage_free, cudaEventDisableTiming));
        C4B_CUDA(cudaEventRecord(e->stage_free, cs));
    }
    b->kernel_name = "affine_systolic";
    {
        char buf[512];
        int nblk = 0;
        for (int p = 0; p < n; ++p) nblk += pairs[p].n_blocked > 0;
        snprintf(buf, sizeof buf,
                 "affine: %d lattices; score pass: %d packed 16-bit (%s, %d rows/lane, %d warp(s) per lattice pair), "
                 "%d int32 (%d rows/lane, %d warp(s) per lattice%s); traceback: %d banded (%s), %d single-pass (%s); "
                 "%d with SubOpt blocked cells",
                 n, b->n16, b->p16_fold ? "offset-binary, folded: one lattice per warp" : b->p16_unsigned ? "offset-binary" : "signed",
                 b->p16_fold ? b->Rf : b->R16,
                 b->p16_multi ? b->fill_warps16 : 1, ns - b->n16, b->R32s, b->fill_warps,
                 b->any_blocked ? ", BLK variant" : "", b->want_path ? ns : 0, b->tb16_band ? "packed 16-bit" : "int32",
                 b->want_path ? nd : 0, b->tb16_direct ? "packed 16-bit" : "int32", nblk);
        b->description = buf;
    }
    return 0;
}

// The encode kernels flag symbols outside the matrix alphabet; the host looks at
// the flag once, at the first synchronisation point of the batch.
int affine_check_symbols(c4b_batch *b) {
    if (b->bad_checked) return 0;
    int bad = 0;
    C4B_CUDA(cudaStreamSynchronize(b->e->copy_stream));
    C4B_CUDA(cudaMemcpy(&bad, b->d_bad.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (bad) {
        set_error("a sequence holds a symbol outside the substitution matrix alphabet");
        return -1;
    }
    b->bad_checked = true;
    return 0;
}

int affine_run(c4b_batch *b, c4b_score threshold) {
    c4b_engine *e = b->e;
    cudaStream_t st = e->stream;
    const int ns = (int)b->score_list.size(), nd = (int)b->direct_list.size();
    b->fill_events_used = 0;
    // pass 1: score + END cell over the full lattices, nothing written per cell.
    // The first run of a batch finds it already queued by affine_create (it started
    // while the sequences were still being staged); later runs queue it again.
    if (!b->pass1_inflight) {
        C4B_CUDA(cudaEventRecord(b->ev_p1a, st));
        C4B_CUDA(cudaEventRecord(b->ev_base, st));
        for (int g = 0; g < (int)b->groups.size(); ++g)
            if (launch_pass1_group(b, g)) return -1;
    }
    b->pass1_inflight = false;
    for (auto &g : b->groups) C4B_CUDA(cudaStreamWaitEvent(st, g.done, 0));
    C4B_CUDA(cudaStreamWaitEvent(st, b->slice_events.back(), 0));  // direct lattices read every slice
    C4B_CUDA(cudaEventRecord(b->ev_p1b, st));
    if (!b->want_path) {
        if (ns) {
            // results are indexed by ordered slot here; the host un-permutes on fetch
            affine_score_results_kernel<<<(ns + 127) / 128, 128, 0, st>>>(b->d_full.p, b->d_out1.p, b->d_qorg.p,
                                                                         b->d_torg.p, ns, b->d_results.p);
            e->launches++;
        }
        C4B_CUDA(cudaGetLastError());
        return 0;
    }
    if (ns) {
        affine_plan_band_kernel<<<(ns + 127) / 128, 128, 0, st>>>(b->d_full.p, b->d_out1.p, b->d_band.p,
                                                                  b->d_band_j0.p, ns, b->max_sub, b->gap_min);
        e->launches++;
        C4B_CUDA(cudaGetLastError());
    }
    // pass 2: refill only the band with the traceback record, then walk it
    for (const Chunk &c : b->band_chunks) {
        const int cnt = c.end - c.begin;
        if (b->tb16_band ? launch_fill16tb(b, b->d_band.p + c.begin, b->d_out2.p, cnt, st)
                         : launch_fill(b, b->d_band.p + c.begin, b->d_out2.p, cnt, true, st, true)) return -1;
        if (launch_traceback(b, b->d_band.p, b->d_out2.p, b->d_out1.p, b->d_band_j0.p,
                             b->d_jobs_band.p + c.begin, cnt)) return -1;
    }
    for (const Chunk &c : b->direct_chunks) {
        const int cnt = c.end - c.begin;
        if (b->tb16_direct ? launch_fill16tb(b, b->d_direct.p + c.begin, b->d_outd.p, cnt, st)
                           : launch_fill(b, b->d_direct.p + c.begin, b->d_outd.p, cnt, true, st, true)) return -1;
        if (launch_traceback(b, b->d_direct.p, b->d_outd.p, nullptr, nullptr,
                             b->d_jobs_direct.p + c.begin, cnt)) return -1;
    }
    (void)nd;
    apply_threshold_kernel<<<(b->n + 127) / 128, 128, 0, st>>>(b->d_results.p, b->n, threshold);
    ops_scan_kernel<<<1, 1024, 0, st>>>(b->d_results.p, b->n, b->d_new_off.p, b->d_new_off.p + b->n);
    ops_compact_kernel<<<b->n, 64, 0, st>>>(b->d_results.p, b->n, b->d_new_off.p, b->d_ops_slots.p,
                                            b->d_ops_packed.p);
    e->launches += 3;
    C4B_CUDA(cudaGetLastError());
    return 0;
}

int affine_fetch(c4b_batch *b, c4b_result *results, int32_t *ops, int64_t ops_capacity) {
    cudaStream_t st = b->e->stream;
    const int n = b->n;
    if (affine_check_symbols(b)) return -1;
    if (!b->want_path) {
        std::vector<c4b_result> tmp(n);
        C4B_CUDA(cudaMemcpyAsync(tmp.data(), b->d_results.p, n * sizeof(c4b_result), cudaMemcpyDeviceToHost, st));
        C4B_CUDA(cudaStreamSynchronize(st));
        for (int k = 0; k < n; ++k) results[b->score_list[k]] = tmp[k];
        return 0;
    }
    int64_t total = 0;
    C4B_CUDA(cudaMemcpyAsync(results, b->d_results.p, n * sizeof(c4b_result), cudaMemcpyDeviceToHost, st));
    C4B_CUDA(cudaMemcpyAsync(&total, b->d_new_off.p + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    C4B_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < n; ++k)
        if (results[k].status >= 2) {
            set_error("internal: traceback of pair " + std::to_string(k) + " failed with status " +
                      std::to_string(results[k].status));
            return -1;
        }
    if (total > ops_capacity) {
        set_error("ops buffer too small: need capacity " + std::to_string(total));
        return -3;
    }
    if (total)
        C4B_CUDA(cudaMemcpy(ops, b->d_ops_packed.p, 2 * (size_t)total * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return 0;
}

}  // namespace

// =============================================================================
extern "C" {

int c4b_abi_version(void) { return C4B_ABI_VERSION; }
const char *c4b_last_error(void) { return c4b::g_error.c_str(); }

int c4b_engine_create(int device, c4b_engine **out) {
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0) {
        set_error(std::string("no CUDA device: ") + cudaGetErrorString(err) +
                  " (libc4b200 has no CPU fallback)");
        return -1;
    }
    if (device < 0 || device >= count) {
        set_error("device ordinal out of range");
        return -1;
    }
    C4B_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    C4B_CUDA(cudaGetDeviceProperties(&prop, device));
    c4b_engine *e = new c4b_engine();
    e->device = device;
    e->sm_count = prop.multiProcessorCount;
    {   // keep freed device memory cached in the stream-ordered pool
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    bool ok = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (cudaStream_t &a : e->aux) ok = ok && cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking) == cudaSuccess;
    if (!ok) {
        set_error("cudaStreamCreate failed");
        delete e;
        return -1;
    }
    *out = e;
    return 0;
}

void c4b_engine_destroy(c4b_engine *e) {
    if (!e) return;
    if (e->stream) cudaStreamSynchronize(e->stream);
    {   // give the cached pool memory back to the driver (other processes may share the GPU)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, e->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    }
    for (auto &kv : e->resident.map) cudaFree(kv.second);
    if (e->h_stage) cudaFreeHost(e->h_stage);
    if (e->stage_free) cudaEventDestroy(e->stage_free);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    for (cudaStream_t a : e->aux)
        if (a) cudaStreamDestroy(a);
    if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

int c4b_engine_set_stream(c4b_engine *e, void *cuda_stream) {
    if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
    e->stream = (cudaStream_t)cuda_stream;
    e->own_stream = false;
    return 0;
}

int64_t c4b_engine_kernel_launches(const c4b_engine *e) { return e->launches; }

int c4b_batch_create(c4b_engine *e, const c4b_model *model, const c4b_scoring *scoring, int32_t n,
                     const c4b_pair *pairs, int want_path, c4b_batch **out) {
    if (!e || !model || !scoring || n < 0 || (n && !pairs)) {
        set_error("c4b_batch_create: bad arguments");
        return -1;
    }
    C4B_CUDA(cudaSetDevice(e->device));
    tl_pool_stream = e->stream;
    c4b_batch *b = new c4b_batch();
    b->e = e;
    b->n = n;
    b->want_path = want_path != 0;
    b->model = *model;
    b->scoring = *scoring;
    int match_kind = 0;
    int rc;
    if (check_model(*model)) {   // every path: the specialised kernels index the same tables
        delete b;
        return -1;
    }
    if (n == 0) {
        *out = b;  // an empty batch runs and fetches nothing
        return 0;
    }
    const bool force_generic = getenv("C4B_FORCE_GENERIC") != nullptr;  // testing: every model through the table-driven path
    if (force_generic) {
        rc = 1;
    } else if (analyze_affine(*model, &b->aff, &match_kind)) {
        b->affine = true;
        rc = affine_create(b, pairs, match_kind);
    } else {
        rc = getenv("C4B_NO_E2G") ? 1
                                  : e2g_batch_create(e->stream, &e->launches, model, scoring, n, pairs,
                                                     want_path != 0, &b->e2g);
        if (rc == 0) {
            b->kernel_name = b->e2g->packed ? "e2g_packed16" : "e2g_systolic";
            b->cells = b->e2g->cells;
        }
    }
    if (rc == 1) {  // neither specialised template applies: table-driven wavefront
        rc = generic_batch_create(e->stream, &e->launches, model, scoring, n, pairs, want_path != 0,
                                  &b->generic, nullptr, false, e->sm_count, &e->resident);
        if (!rc) {
            b->kernel_name = "generic_wavefront";
            b->cells = generic_batch_cells(b->generic);
        }
    }
    if (rc) {
        delete b;
        return rc;
    }
    *out = b;
    return 0;
}

int c4b_batch_run(c4b_batch *b, c4b_score threshold) {
    C4B_CUDA(cudaSetDevice(b->e->device));
    tl_pool_stream = b->e->stream;
    b->ran = true;
    if (b->n == 0) return 0;
    if (b->affine) return affine_run(b, threshold);
    if (b->e2g) return e2g_batch_run(b->e2g, threshold);
    return generic_batch_run(b->generic, threshold);
}

int c4b_batch_fetch(c4b_batch *b, c4b_result *results, int32_t *ops, int64_t ops_capacity) {
    if (!b->ran) {
        set_error("c4b_batch_fetch before c4b_batch_run");
        return -1;
    }
    if (b->n == 0) return 0;
    if (b->affine) return affine_fetch(b, results, ops, ops_capacity);
    if (b->e2g) return e2g_batch_fetch(b->e2g, results, ops, ops_capacity);
    return generic_batch_fetch(b->generic, results, ops, ops_capacity);
}

int64_t c4b_batch_ops_needed(c4b_batch *b) {
    if (!b || !b->ran) {
        set_error("c4b_batch_ops_needed before c4b_batch_run");
        return -1;
    }
    if (b->n == 0 || !b->want_path) return 0;
    cudaSetDevice(b->e->device);
    const int64_t *d_total = b->affine ? b->d_new_off.p + b->n
                             : b->e2g  ? b->e2g->d_new_off.p + b->n
                                       : b->generic->d_new_off.p + b->n;
    cudaStream_t st = b->affine ? b->e->stream : b->e2g ? b->e2g->stream : b->generic->stream;
    int64_t total = 0;
    if (cudaMemcpyAsync(&total, d_total, sizeof(int64_t), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) {
        set_error(std::string("c4b_batch_ops_needed: ") + cudaGetErrorString(cudaGetLastError()));
        return -1;
    }
    return total;
}

const void *c4b_batch_device_results(const c4b_batch *b) {
    if (!b->ran || b->n == 0) return nullptr;
    if (b->affine) return b->want_path ? b->d_results.p : nullptr;  // score-only is in slot order
    if (b->e2g) return b->want_path ? b->e2g->d_results.p : nullptr;
    return generic_batch_device_results(b->generic);
}

int64_t c4b_batch_cells(const c4b_batch *b) { return b->cells; }

double c4b_batch_last_fill_ms(c4b_batch *b) {
    if (!b->ran || b->n == 0) return -1.0;
    if (b->e2g) return e2g_batch_fill_ms(b->e2g);
    if (!b->affine) return generic_batch_fill_ms(b->generic);
    cudaStreamSynchronize(b->e->stream);
    double ms = 0;
    {
        float f = 0;  // pass 1: fork to the aux streams .. join, measured on the engine stream
        if (cudaEventElapsedTime(&f, b->ev_p1a, b->ev_p1b) == cudaSuccess) ms += f;
    }
    for (int k = 0; k < b->fill_events_used; ++k) {
        float f = 0;
        if (cudaEventElapsedTime(&f, b->fill_events[k].a, b->fill_events[k].b) == cudaSuccess) ms += f;
    }
    return ms;
}

int c4b_model_specialise(const c4b_model *model, int32_t mode, int32_t cta_threads, int64_t *cubin_bytes) {
    if (!model || mode < 0 || mode > 2 || (cta_threads != 128 && cta_threads != 256 && cta_threads != 512)) {
        set_error("c4b_model_specialise: bad arguments");
        return -1;
    }
    if (check_model(*model)) return -1;
    std::vector<char> cubin;
    std::string log;
    // every variant the launcher can ask for: lattice ring in shared memory or L2, and for
    // FIND_REGION the one-slot / two-slot start cell; with C4B_JIT_CACHE_DIR set they are left
    // in the disk cache under the names the batch entry points look up
    for (int smem_ring = 0; smem_ring < 2; ++smem_ring)
        for (int pack_start = 0; pack_start < (mode == GEN_REGION ? 2 : 1); ++pack_start) {
            const std::string src = jit_program_source(*model, mode, cta_threads, smem_ring != 0, pack_start != 0);
            if (!jit_compile(src, &cubin, &log)) {
                set_error("model specialisation failed: " + log);
                return -1;
            }
            jit_store(jit_cache_path(src), cubin);
        }
    // ... and the systolic specialisation the launcher prefers for lattices without SubOpt
    // blocked cells / cell-callback tables (cta_threads does not apply: one warp per strip)
    {
        const char *env = getenv("C4B_JIT_SYSTOLIC");
        const SysLayout L = jit_sys_layout(*model, mode, mode == GEN_REGION);
        if (L.ok && !(env && atoi(env) == 0)) {
            // (for the strip counts of queries that fill cta_threads rows of the thread-per-row kernel)
            // + the column-window variants: checkpointing score pass / PATH over one window
            // + the SubOpt form of the whole-lattice pass (blocked cells as per-strip entries): what every
            // iteration after the first of a `--subopt yes` run (the CLI default) launches
            for (int v = 0; v < 4; ++v) {
                const int win = v < 3 ? v : 0;
                const bool blk = v == 3;
                if ((win == 1 && mode != GEN_SCORE) || (win == 2 && mode != GEN_PATH)) continue;
                const std::string src = jit_sys_program_source(*model, mode, mode == GEN_REGION, L,
                                                               jit_sys_warps(L, cta_threads - 1), win, blk);
                if (!jit_compile(src, &cubin, &log)) {
                    set_error("systolic model specialisation failed: " + log);
                    return -1;
                }
                jit_store(jit_cache_path(src), cubin);
            }
        }
    }
    if (cubin_bytes) *cubin_bytes = (int64_t)cubin.size();
    return 0;
}

const char *c4b_batch_kernel_name(const c4b_batch *b) {
    return b->generic ? b->generic->kernel_used : b->kernel_name;
}

const char *c4b_batch_description(const c4b_batch *b) {
    if (!b->description.empty()) return b->description.c_str();
    return c4b_batch_kernel_name(b);
}

void c4b_batch_destroy(c4b_batch *b) { delete b; }

int c4b_find_score_batch(c4b_engine *e, const c4b_model *model, const c4b_scoring *scoring, int32_t n,
                         const c4b_pair *pairs, c4b_score *scores) {
    c4b_batch *b = nullptr;
    int rc = c4b_batch_create(e, model, scoring, n, pairs, 0, &b);
    if (rc) return rc;
    std::vector<c4b_result> res(n);
    rc = c4b_batch_run(b, C4B_IMPOSSIBLY_LOW_SCORE);
    if (!rc) rc = c4b_batch_fetch(b, res.data(), nullptr, 0);
    if (!rc)
        for (int k = 0; k < n; ++k) scores[k] = res[k].score;
    c4b_batch_destroy(b);
    return rc;
}

int c4b_find_path_batch(c4b_engine *e, const c4b_model *model, const c4b_scoring *scoring, int32_t n,
                        const c4b_pair *pairs, c4b_score threshold, c4b_result *results, int32_t *ops,
                        int64_t ops_capacity) {
    HostTimeline tl;
    tl_timeline = &tl;
    c4b_batch *b = nullptr;
    int rc = c4b_batch_create(e, model, scoring, n, pairs, 1, &b);
    if (rc) { tl_timeline = nullptr; return rc; }
    tmark("batch created");
    rc = c4b_batch_run(b, threshold);
    tmark("run: everything queued");
    if (!rc) rc = c4b_batch_fetch(b, results, ops, ops_capacity);
    tmark("fetch: results and ops on the host");
    c4b_batch_destroy(b);
    tmark("batch destroyed");
    tl_timeline = nullptr;
    return rc;
}

// ---- device groups -----------------------------------------------------------------------
}  // extern "C"

struct c4b_group {
    std::vector<c4b_engine *> engines;
};

namespace {

// LPT: lattices by cost, largest first, each to the device with the least work so far
std::vector<std::vector<int>> group_shards(int n_dev, int n, const c4b_pair *pairs) {
    std::vector<int> order(n);
    for (int k = 0; k < n; ++k) order[k] = k;
    auto cost = [&](int k) { return (int64_t)pairs[k].query_length * pairs[k].target_length; };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost(a) > cost(b); });
    std::vector<std::vector<int>> shard(n_dev);
    std::vector<int64_t> load(n_dev, 0);
    for (int k : order) {
        const int d = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        shard[d].push_back(k);
        load[d] += cost(k) + 1;
    }
    for (auto &s : shard) std::sort(s.begin(), s.end());   // pair order inside a shard
    return shard;
}

// one worker per device; the first error (if any) is handed to the calling thread
template <typename Work>
int group_run(c4b_group *g, const std::vector<std::vector<int>> &shard, Work work) {
    const int nd = (int)g->engines.size();
    std::vector<int> rc(nd, 0);
    std::vector<std::string> err(nd);
    std::vector<std::thread> th;
    for (int d = 0; d < nd; ++d)
        th.emplace_back([&, d] {
            if (shard[d].empty()) return;
            rc[d] = work(d);
            if (rc[d]) err[d] = c4b::g_error;   // thread-local in the worker
        });
    for (auto &t : th) t.join();
    for (int d = 0; d < nd; ++d)
        if (rc[d]) {
            set_error("device " + std::to_string(g->engines[d]->device) + ": " + err[d]);
            return rc[d];
        }
    return 0;
}

}  // namespace

extern "C" {

int c4b_group_create(int n_devices, const int *devices, c4b_group **out) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        set_error("no CUDA device (libc4b200 has no CPU fallback)");
        return -1;
    }
    std::vector<int> dev;
    if (n_devices <= 0)
        for (int d = 0; d < count; ++d) dev.push_back(d);
    else
        dev.assign(devices, devices + n_devices);
    c4b_group *g = new c4b_group();
    for (int d : dev) {
        c4b_engine *e = nullptr;
        if (c4b_engine_create(d, &e)) {
            c4b_group_destroy(g);
            return -1;
        }
        g->engines.push_back(e);
    }
    *out = g;
    return 0;
}

void c4b_group_destroy(c4b_group *g) {
    if (!g) return;
    for (c4b_engine *e : g->engines) c4b_engine_destroy(e);
    delete g;
}

int c4b_group_size(const c4b_group *g) { return g ? (int)g->engines.size() : 0; }

int64_t c4b_group_kernel_launches(const c4b_group *g) {
    int64_t n = 0;
    for (c4b_engine *e : g->engines) n += e->launches;
    return n;
}

void c4b_free(void *p) { free(p); }

int c4b_group_find_score_batch(c4b_group *g, const c4b_model *model, const c4b_scoring *scoring, int32_t n,
                               const c4b_pair *pairs, c4b_score *scores) {
    if (!g || g->engines.empty() || n < 0 || (n && (!pairs || !scores))) {
        set_error("c4b_group_find_score_batch: bad arguments");
        return -1;
    }
    const auto shard = group_shards((int)g->engines.size(), n, pairs);
    return group_run(g, shard, [&](int d) -> int {
        const std::vector<int> &idx = shard[d];
        std::vector<c4b_pair> mine(idx.size());
        for (size_t k = 0; k < idx.size(); ++k) mine[k] = pairs[idx[k]];
        std::vector<c4b_score> sc(idx.size());
        const int rc = c4b_find_score_batch(g->engines[d], model, scoring, (int32_t)mine.size(), mine.data(), sc.data());
        if (!rc)
            for (size_t k = 0; k < idx.size(); ++k) scores[idx[k]] = sc[k];
        return rc;
    });
}

int c4b_group_find_path_batch(c4b_group *g, const c4b_model *model, const c4b_scoring *scoring, int32_t n,
                              const c4b_pair *pairs, c4b_score threshold, c4b_result *results, int32_t **ops_out,
                              int64_t *n_ops_out) {
    if (!g || g->engines.empty() || n < 0 || (n && (!pairs || !results)) || !ops_out || !n_ops_out) {
        set_error("c4b_group_find_path_batch: bad arguments");
        return -1;
    }
    *ops_out = nullptr;
    *n_ops_out = 0;
    const int nd = (int)g->engines.size();
    const auto shard = group_shards(nd, n, pairs);
    std::vector<std::vector<c4b_result>> res(nd);
    std::vector<std::vector<int32_t>> ops(nd);
    int rc = group_run(g, shard, [&](int d) -> int {
        const std::vector<int> &idx = shard[d];
        std::vector<c4b_pair> mine(idx.size());
        for (size_t k = 0; k < idx.size(); ++k) mine[k] = pairs[idx[k]];
        c4b_batch *b = nullptr;
        int r = c4b_batch_create(g->engines[d], model, scoring, (int32_t)mine.size(), mine.data(), 1, &b);
        if (r) return r;
        r = c4b_batch_run(b, threshold);
        int64_t need = r ? -1 : c4b_batch_ops_needed(b);
        if (!r && need < 0) r = -1;
        if (!r) {
            res[d].resize(idx.size());
            ops[d].resize(2 * (size_t)need + 2);
            r = c4b_batch_fetch(b, res[d].data(), ops[d].data(), need);
            ops[d].resize(2 * (size_t)need);
        }
        c4b_batch_destroy(b);
        return r;
    });
    if (rc) return rc;
    int64_t total = 0;
    for (int d = 0; d < nd; ++d) total += (int64_t)ops[d].size() / 2;
    int32_t *all = (int32_t *)malloc(sizeof(int32_t) * (2 * (size_t)total + 2));
    if (!all) {
        set_error("c4b_group_find_path_batch: out of host memory");
        return -1;
    }
    int64_t base = 0;
    for (int d = 0; d < nd; ++d) {   // shard op lists one after the other; offsets rebased
        if (!ops[d].empty()) memcpy(all + 2 * base, ops[d].data(), ops[d].size() * sizeof(int32_t));
        for (size_t k = 0; k < shard[d].size(); ++k) {
            c4b_result r = res[d][k];
            r.ops_offset += base;
            results[shard[d][k]] = r;
        }
        base += (int64_t)ops[d].size() / 2;
    }
    *ops_out = all;
    *n_ops_out = total;
    return 0;
}

int c4b_viterbi_calculate(c4b_engine *e, const c4b_model *model, const c4b_scoring *scoring,
                          const c4b_pair *pair, int mode, c4b_result *result, int32_t *ops,
                          int64_t ops_capacity) {
    // FIND_SCORE and FIND_REGION need no operation list; FIND_REGION's bounding
    // box is the path's (Viterbi_Data_finalise, viterbi.c:633-653), so it is
    // served by the path machinery and the ops are dropped.
    if (mode == 0) {
        c4b_batch *b = nullptr;
        int rc = c4b_batch_create(e, model, scoring, 1, pair, 0, &b);
        if (rc) return rc;
        rc = c4b_batch_run(b, C4B_IMPOSSIBLY_LOW_SCORE);
        if (!rc) rc = c4b_batch_fetch(b, result, nullptr, 0);
        c4b_batch_destroy(b);
        return rc;
    }
    if (mode == 1)
        return c4b_find_path_batch(e, model, scoring, 1, pair, C4B_IMPOSSIBLY_LOW_SCORE, result, ops,
                                   ops_capacity);
    if (mode == 2) {
        const int64_t cap = (int64_t)pair->query_length + pair->target_length + 8;
        std::vector<int32_t> tmp(2 * (size_t)cap);
        int rc = c4b_find_path_batch(e, model, scoring, 1, pair, C4B_IMPOSSIBLY_LOW_SCORE, result,
                                     tmp.data(), cap);
        if (!rc) result->n_ops = 0;
        return rc;
    }
    set_error("c4b_viterbi_calculate: unknown mode");
    return -1;
}

int c4b_viterbi_calculate_cells(c4b_engine *e, const c4b_model *model, const c4b_scoring *scoring,
                                const c4b_pair *pair, int mode, const c4b_score *start_cells,
                                c4b_score *end_cells, c4b_result *result, int32_t *ops, int64_t ops_capacity) {
    if (!e || !model || !scoring || !pair || !result || (mode != 0 && mode != 1)) {
        set_error("c4b_viterbi_calculate_cells: bad arguments");
        return -1;
    }
    C4B_CUDA(cudaSetDevice(e->device));
    tl_pool_stream = e->stream;
    GenericBatch *g = nullptr;
    int rc = generic_batch_create(e->stream, &e->launches, model, scoring, 1, pair, mode == 1, &g, start_cells,
                                  end_cells != nullptr, e->sm_count, &e->resident);
    if (rc) return rc;
    rc = generic_batch_run(g, C4B_IMPOSSIBLY_LOW_SCORE);
    if (!rc) rc = generic_batch_fetch(g, result, ops, ops_capacity);
    if (!rc && end_cells) {
        const size_t C = 1 + (size_t)model->n_shadow_slots;
        const size_t cells = ((size_t)pair->query_length + 1) * ((size_t)pair->target_length + 1);
        std::vector<int32_t> tmp(cells * C);
        if (cudaMemcpy(tmp.data(), g->d_endm.p, cells * C * sizeof(int32_t), cudaMemcpyDeviceToHost) != cudaSuccess) {
            set_error("copying the END cells back failed");
            rc = -1;
        } else {
            for (size_t k = 0; k < cells; ++k)
                if (tmp[k * C] != kEndMatrixUnset)
                    for (size_t l = 0; l < C; ++l) end_cells[k * C + l] = tmp[k * C + l];
        }
    }
    generic_batch_destroy(g);
    return rc;
}

int c4b_span_integrate(c4b_engine *e, const c4b_score *src_scores, const int32_t *src_region,
                       const int32_t *dst_region, const int32_t *span, int32_t *positions) {
    if (!e || !src_scores || !src_region || !dst_region || !span || !positions || src_region[2] < 0 ||
        src_region[3] < 0 || dst_region[2] < 0 || dst_region[3] < 0) {
        set_error("c4b_span_integrate: bad arguments");
        return -1;
    }
    C4B_CUDA(cudaSetDevice(e->device));
    tl_pool_stream = e->stream;
    const size_t n_src = ((size_t)src_region[2] + 1) * ((size_t)src_region[3] + 1);
    const size_t n_dst = ((size_t)dst_region[2] + 1) * ((size_t)dst_region[3] + 1);
    if (n_src > ((size_t)1 << 30) || n_dst > ((size_t)1 << 29)) {
        set_error("c4b_span_integrate: region too large");
        return -1;
    }
    DevBuf<int32_t> d_src, d_pos;
    int rc = 0;
    if (d_src.alloc(n_src) || d_pos.alloc(2 * n_dst)) rc = -1;
    if (!rc) {
        SpanArgs a;
        a.sqs = src_region[0]; a.sts = src_region[1]; a.sql = src_region[2]; a.stl = src_region[3];
        a.dqs = dst_region[0]; a.dts = dst_region[1]; a.dql = dst_region[2]; a.dtl = dst_region[3];
        a.min_q = span[0]; a.max_q = span[1]; a.min_t = span[2]; a.max_t = span[3];
        const int threads = 128;
        const int grid = (int)std::min<size_t>((n_dst + threads - 1) / threads, (size_t)e->sm_count * 16);
        if (cudaMemcpyAsync(d_src.p, src_scores, n_src * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream) !=
            cudaSuccess)
            rc = -1;
        if (!rc) {
            span_integrate_kernel<<<std::max(grid, 1), threads, 0, e->stream>>>(d_src.p, a, d_pos.p);
            e->launches++;
            if (cudaGetLastError() != cudaSuccess ||
                cudaMemcpyAsync(positions, d_pos.p, 2 * n_dst * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream) !=
                    cudaSuccess ||
                cudaStreamSynchronize(e->stream) != cudaSuccess)
                rc = -1;
        }
        if (rc) set_error(std::string("c4b_span_integrate: ") + cudaGetErrorString(cudaGetLastError()));
    }
    d_src.release();
    d_pos.release();
    return rc;
}

int c4b_span_score_batch(c4b_engine *e, const c4b_model *src_model, const c4b_model *dst_model,
                         const c4b_scoring *scoring, int32_t n, const c4b_span_job *jobs, c4b_score *scores) {
    if (!e || !src_model || !dst_model || !scoring || n < 0 || (n && (!jobs || !scores))) {
        set_error("c4b_span_score_batch: bad arguments");
        return -1;
    }
    if (n == 0) return 0;
    C4B_CUDA(cudaSetDevice(e->device));
    tl_pool_stream = e->stream;
    cudaStream_t st = e->stream;
    const int c_src = 1 + src_model->n_shadow_slots, c_dst = 1 + dst_model->n_shadow_slots;
    std::vector<c4b_pair> src(n), dst(n);
    std::vector<size_t> end_off(n), start_off(n);
    size_t end_ints = 0, start_ints = 0;
    int max_dst_cells = 1;
    for (int k = 0; k < n; ++k) {
        src[k] = jobs[k].src;
        dst[k] = jobs[k].dst;
        if (src[k].query_length < 0 || src[k].target_length < 0 || dst[k].query_length < 0 || dst[k].target_length < 0) {
            set_error("c4b_span_score_batch: job " + std::to_string(k) + " has a negative region extent");
            return -1;
        }
        end_off[k] = end_ints;
        start_off[k] = start_ints;
        end_ints += ((size_t)src[k].query_length + 1) * ((size_t)src[k].target_length + 1) * c_src;
        const size_t dc = ((size_t)dst[k].query_length + 1) * ((size_t)dst[k].target_length + 1);
        start_ints += dc * c_dst;
        max_dst_cells = (int)std::max<size_t>(max_dst_cells, std::min<size_t>(dc, 1u << 30));
    }
    DevBuf<int32_t> d_end, d_start;
    DevBuf<SpanJob> d_jobs;
    GenericBatch *gs = nullptr, *gd = nullptr;
    int rc = 0;
    auto cleanup = [&]() {
        if (gs) generic_batch_destroy(gs);
        if (gd) generic_batch_destroy(gd);
        d_end.release(); d_start.release(); d_jobs.release();
    };
    if (d_end.alloc(end_ints + 4) || d_start.alloc(start_ints + 4) || d_jobs.alloc(n)) { cleanup(); return -1; }
    // "END not reached": the byte pattern reads as a score far below C4_IMPOSSIBLY_LOW_SCORE
    if (cudaMemsetAsync(d_end.p, 0x80, end_ints * sizeof(int32_t), st) != cudaSuccess) rc = -1;
    std::vector<int32_t *> end_ptr(n);
    std::vector<const int32_t *> start_ptr(n);
    std::vector<SpanJob> hj(n);
    for (int k = 0; k < n; ++k) {
        end_ptr[k] = d_end.p + end_off[k];
        start_ptr[k] = d_start.p + start_off[k];
        SpanJob &J = hj[k];
        J.a.sqs = src[k].query_start; J.a.sts = src[k].target_start;
        J.a.sql = src[k].query_length; J.a.stl = src[k].target_length;
        J.a.dqs = dst[k].query_start; J.a.dts = dst[k].target_start;
        J.a.dql = dst[k].query_length; J.a.dtl = dst[k].target_length;
        J.a.min_q = jobs[k].span[0]; J.a.max_q = jobs[k].span[1];
        J.a.min_t = jobs[k].span[2]; J.a.max_t = jobs[k].span[3];
        J.src_end = end_ptr[k];
        J.dst_start = d_start.p + start_off[k];
        J.c_src = c_src;
        J.c_dst = c_dst;
    }
    GenDevTables ts, td;
    ts.end = end_ptr.data();
    td.start = start_ptr.data();
    // 1) src fills: END's cell of every cell that reaches END
    if (!rc) rc = generic_batch_create(st, &e->launches, src_model, scoring, n, src.data(), false, &gs, nullptr, false,
                                       e->sm_count, &e->resident, &ts);
    if (!rc) rc = generic_batch_run(gs, C4B_IMPOSSIBLY_LOW_SCORE);
    // 2) integrate + START tables of the dst fills
    if (!rc) {
        if (cudaMemcpyAsync(d_jobs.p, hj.data(), n * sizeof(SpanJob), cudaMemcpyHostToDevice, st) != cudaSuccess) rc = -1;
        const int threads = 128;
        const dim3 grid((unsigned)std::max(1, std::min((max_dst_cells + threads - 1) / threads, 64)), (unsigned)n);
        span_start_table_kernel<<<grid, threads, 0, st>>>(d_jobs.p);
        e->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = -1;
        if (rc) set_error(std::string("c4b_span_score_batch: ") + cudaGetErrorString(cudaGetLastError()));
    }
    // 3) dst fills from the tables
    if (!rc) rc = generic_batch_create(st, &e->launches, dst_model, scoring, n, dst.data(), false, &gd, nullptr, false,
                                       e->sm_count, &e->resident, &td);
    if (!rc) rc = generic_batch_run(gd, C4B_IMPOSSIBLY_LOW_SCORE);
    if (!rc) {
        std::vector<c4b_result> res(n);
        rc = generic_batch_fetch(gd, res.data(), nullptr, 0);
        if (!rc)
            for (int k = 0; k < n; ++k) scores[k] = res[k].score;
    }
    if (rc) cudaStreamSynchronize(st);   // the staging vectors above must outlive the copies
    cleanup();
    return rc;
}

void c4b_engine_forget_buffers(c4b_engine *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    for (auto &kv : e->resident.map) cudaFree(kv.second);
    e->resident.map.clear();
    e->resident.bytes = 0;
}

void c4b_engine_forget_buffer(c4b_engine *e, const void *host) {
    if (!e || !host) return;
    cudaSetDevice(e->device);
    bool synced = false;
    for (auto it = e->resident.map.begin(); it != e->resident.map.end();) {
        if (it->first.first == host) {
            if (!synced) { cudaStreamSynchronize(e->stream); synced = true; }
            cudaFree(it->second);
            e->resident.bytes -= std::min(e->resident.bytes, it->first.second);
            it = e->resident.map.erase(it);
        } else {
            ++it;
        }
    }
}

int c4b_hsp_extend_batch(c4b_engine *e, const c4b_scoring *scoring, const c4b_hsp_param *param,
                         const uint8_t *query, int32_t query_len, const uint8_t *query_mask,
                         const uint8_t *target, int32_t target_len, const uint8_t *target_mask,
                         int32_t n_seeds, const c4b_hsp_seed *seeds, c4b_hsp *out) {
    if (!e || !scoring || !param || !query || !target || query_len < 0 || target_len < 0 || n_seeds < 0 ||
        (n_seeds && (!seeds || !out))) {
        set_error("c4b_hsp_extend_batch: bad arguments");
        return -1;
    }
    const int kind = param->match_kind;
    if (kind != C4B_CALC_MATCH_DNA && kind != C4B_CALC_MATCH_PROTEIN && kind != C4B_CALC_MATCH_1_3 &&
        kind != C4B_CALC_MATCH_3_1 && kind != C4B_CALC_MATCH_3_3) {
        set_error("c4b_hsp_extend_batch: match kind has no device form");
        return -1;
    }
    if (param->seedlen <= 0) {
        set_error("c4b_hsp_extend_batch: seed length must be positive");
        return -1;
    }
    if (n_seeds == 0) return 0;
    const int tadv = (kind == C4B_CALC_MATCH_1_3 || kind == C4B_CALC_MATCH_3_3) ? 3 : 1;
    const int qadv = (kind == C4B_CALC_MATCH_3_1 || kind == C4B_CALC_MATCH_3_3) ? 3 : 1;
    for (int k = 0; k < n_seeds; ++k)   // HSP_check (hspset.c): the seed lies inside both sequences
        if (seeds[k].query_start < 0 || seeds[k].target_start < 0 ||
            (int64_t)seeds[k].query_start + (int64_t)param->seedlen * qadv > query_len ||
            (int64_t)seeds[k].target_start + (int64_t)param->seedlen * tadv > target_len) {
            set_error("seed " + std::to_string(k) + " outside the sequences");
            return -1;
        }
    C4B_CUDA(cudaSetDevice(e->device));
    tl_pool_stream = e->stream;
    cudaStream_t st = e->stream;
    const bool dna = (kind == C4B_CALC_MATCH_DNA);
    DevBuf<uint8_t> d_q, d_t, d_qm, d_tm, d_qc, d_tc, d_qmo, d_tmo, d_tab;
    DevBuf<int32_t> d_matrix;
    DevBuf<int> d_bad;
    DevBuf<c4b_hsp_seed> d_seeds;
    DevBuf<c4b_hsp> d_out;
    struct Releaser {
        std::vector<std::function<void()>> f;
        ~Releaser() { for (auto &x : f) x(); }
    } rel;
    auto own = [&](auto &buf) { rel.f.push_back([&buf] { buf.release(); }); };
    own(d_q); own(d_t); own(d_qm); own(d_tm); own(d_qc); own(d_tc); own(d_qmo); own(d_tmo); own(d_tab);
    own(d_matrix); own(d_bad); own(d_seeds); own(d_out);
    const size_t ql = (size_t)query_len, tl = (size_t)target_len;
    if (d_q.alloc(ql + 4) || d_t.alloc(tl + 4) || d_qc.alloc(ql + 4) || d_tc.alloc(tl + 4) ||
        d_tab.alloc(256 * 3 + 4096) || d_matrix.alloc(24 * 24) || d_bad.alloc(1) || d_seeds.alloc(n_seeds) ||
        d_out.alloc(n_seeds))
        return -1;
    if (query_mask && (d_qm.alloc(ql + 4) || d_qmo.alloc(ql + 4))) return -1;
    if (target_mask && (d_tm.alloc(tl + 4) || d_tmo.alloc(tl + 4))) return -1;
    // tables: [0,256) query index, [256,512) target index, [512,768) nt2d, [768,..) codon_aa
    std::vector<uint8_t> tab(256 * 3 + 4096);
    memcpy(&tab[0], dna ? scoring->dna_index : scoring->protein_index, 256);
    memcpy(&tab[256], dna ? scoring->dna_index : scoring->protein_index, 256);
    memcpy(&tab[512], scoring->nt2d, 256);
    memcpy(&tab[768], scoring->codon_aa, 4096);
    C4B_CUDA(cudaMemcpyAsync(d_tab.p, tab.data(), tab.size(), cudaMemcpyHostToDevice, st));
    C4B_CUDA(cudaMemcpyAsync(d_matrix.p, dna ? scoring->dna_matrix : scoring->protein_matrix, 24 * 24 * sizeof(int32_t),
                             cudaMemcpyHostToDevice, st));
    C4B_CUDA(cudaMemcpyAsync(d_q.p, query, ql, cudaMemcpyHostToDevice, st));
    C4B_CUDA(cudaMemcpyAsync(d_t.p, target, tl, cudaMemcpyHostToDevice, st));
    if (query_mask) C4B_CUDA(cudaMemcpyAsync(d_qm.p, query_mask, ql, cudaMemcpyHostToDevice, st));
    if (target_mask) C4B_CUDA(cudaMemcpyAsync(d_tm.p, target_mask, tl, cudaMemcpyHostToDevice, st));
    C4B_CUDA(cudaMemcpyAsync(d_seeds.p, seeds, (size_t)n_seeds * sizeof(c4b_hsp_seed), cudaMemcpyHostToDevice, st));
    C4B_CUDA(cudaMemsetAsync(d_bad.p, 0, sizeof(int), st));
    if (query_len)
        hsp_encode_kernel<<<(query_len + 255) / 256, 256, 0, st>>>(d_q.p, query_len, d_tab.p, d_tab.p + 512,
                                                                   d_tab.p + 768, qadv == 3, d_qc.p, d_qm.p, d_qmo.p,
                                                                   d_bad.p);
    if (target_len)
        hsp_encode_kernel<<<(target_len + 255) / 256, 256, 0, st>>>(d_t.p, target_len, d_tab.p + 256, d_tab.p + 512,
                                                                    d_tab.p + 768, tadv == 3, d_tc.p, d_tm.p,
                                                                    d_tmo.p, d_bad.p);
    HspArgs A;
    A.qc = d_qc.p; A.tc = d_tc.p; A.qm = query_mask ? d_qmo.p : nullptr; A.tm = target_mask ? d_tmo.p : nullptr;
    A.ql = query_len; A.tl = target_len; A.qadv = qadv; A.tadv = tadv;
    A.seedlen = param->seedlen; A.dropoff = param->dropoff; A.threshold = param->threshold;
    hsp_extend_kernel<<<(n_seeds + 127) / 128, 128, 0, st>>>(A, d_matrix.p, n_seeds, d_seeds.p, d_out.p);
    e->launches += 3;
    C4B_CUDA(cudaGetLastError());
    int bad = 0;
    C4B_CUDA(cudaMemcpyAsync(out, d_out.p, (size_t)n_seeds * sizeof(c4b_hsp), cudaMemcpyDeviceToHost, st));
    C4B_CUDA(cudaMemcpyAsync(&bad, d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    C4B_CUDA(cudaStreamSynchronize(st));
    if (bad) {
        set_error("a sequence holds a symbol outside the substitution matrix alphabet");
        return -1;
    }
    return 0;
}

}  // extern "C"
