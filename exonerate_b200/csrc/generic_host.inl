// generic_host.inl -- host side of the table-driven path (included by c4b200.cu
// after DevBuf / the ops packing kernels are defined).
namespace c4b {

struct GenericBatch {
    cudaStream_t stream = nullptr;
    int64_t *launches = nullptr;
    int n = 0;
    bool want_path = false;
    bool use_region = false;
    GenTables tables;
    int64_t cells = 0;
    int sm_count = 0, grid = 0, cmax = 1;
    int threads = 256;  // per CTA: one anti-diagonal of the longest query in one pass if it fits 1024
    bool use_jit = false;  // model-specialised kernel (generic_jit.inl) instead of the interpreter
    int jit_threads = 128;
    int ring_ctas = 0;     // CTAs the lattice ring was sized for
    int max_q = 0, max_t = 0;
    const char *kernel_used = "generic_wavefront";
    // the systolic specialisation (generic_jit_systolic.cuh): lattices without SubOpt blocked cells
    // or cell-callback tables; layouts per fill mode, hand-off rows of the strip sweeps
    bool plain = true;
    SysLayout sys_score, sys_region, sys_path;
    DevBuf<int32_t> d_top;
    size_t top_stride = 0;
    std::vector<c4b_pair> host_pairs;
    std::vector<GenPair> h_full;
    DevBuf<GenTables> d_tables;
    DevBuf<uint8_t> d_seq;
    DevBuf<int32_t> d_ints;  // splice arrays + blocked lists
    DevBuf<GenPair> d_full, d_box;
    DevBuf<GenOut> d_out_a, d_out_b;
    DevBuf<GenJob> d_jobs;
    DevBuf<int32_t> d_ring;
    size_t ring_stride = 0;
    DevBuf<int> d_cursor;
    DevBuf<uint8_t> d_tb;
    DevBuf<c4b_result> d_results;
    DevBuf<int32_t> d_ops_slots, d_ops_packed;
    DevBuf<int64_t> d_new_off;
    DevBuf<int32_t> d_endm, d_startc;  // END-cell / START-cell tables of lattice 0 (BSDP cell callbacks), else empty
    // lattices whose PATH record exceeds the budget: column checkpoints + window refills (JIT_SYS_WIN)
    DevBuf<GenPair> d_wpairs;
    DevBuf<GenWin> d_wins;
    DevBuf<GenWalk> d_wwalk;
    DevBuf<int32_t> d_wbig, d_wck;
    DevBuf<uint8_t> d_wtb;
    int windowed_lattices = 0, window_cols = 0, window_rounds = 0;
    // SubOpt blocked cells on the systolic kernel (JIT_SYS_BLK): the callers' lists (kept: the entries
    // are rebuilt per pass, for the full lattices and for the alignment boxes) and their device form
    bool sys_blk = false;
    std::vector<int32_t> h_blk_q, h_blk_t;
    std::vector<size_t> h_blk_at;
    DevBuf<int2> d_sblk_a, d_sblk_b;
    DevBuf<int32_t> d_sblk_off_a, d_sblk_off_b;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    double fill_ms = -1;
    ~GenericBatch() {
        d_endm.release(); d_startc.release(); d_top.release();
        d_wpairs.release(); d_wins.release(); d_wwalk.release(); d_wbig.release(); d_wck.release(); d_wtb.release();
        d_sblk_a.release(); d_sblk_b.release(); d_sblk_off_a.release(); d_sblk_off_b.release();
        d_tables.release(); d_seq.release(); d_ints.release(); d_full.release(); d_box.release();
        d_out_a.release(); d_out_b.release(); d_jobs.release(); d_ring.release(); d_cursor.release();
        d_tb.release(); d_results.release(); d_ops_slots.release(); d_ops_packed.release();
        d_new_off.release();
        if (ev_a) cudaEventDestroy(ev_a);
        if (ev_b) cudaEventDestroy(ev_b);
    }
};

static thread_local PinnedScratch tl_gen_scratch;

static bool model_needs_splice(const c4b_model &m) {
    for (int k = 0; k < m.n_calcs; ++k)
        if (m.calcs[k].kind == C4B_CALC_SPLICE_PRE || m.calcs[k].kind == C4B_CALC_SPLICE_POST) return true;
    return false;
}

static int check_model(const c4b_model &m) {
    if (m.n_states < 2 || m.n_states > C4B_MAX_STATES || m.n_transitions < 1 ||
        m.n_transitions > C4B_MAX_TRANSITIONS || m.n_calcs < 0 || m.n_calcs > C4B_MAX_CALCS ||
        m.n_shadow_slots < 0 || m.n_shadow_slots > C4B_MAX_SHADOW_SLOTS) {
        set_error("model tables out of range");
        return -1;
    }
    if (m.start_state < 0 || m.start_state >= m.n_states || m.end_state < 0 || m.end_state >= m.n_states ||
        m.start_scope < C4B_SCOPE_ANYWHERE || m.start_scope > C4B_SCOPE_CORNER ||
        m.end_scope < C4B_SCOPE_ANYWHERE || m.end_scope > C4B_SCOPE_CORNER) {
        set_error("model START / END state or scope out of range");
        return -1;
    }
    for (int k = 0; k < m.n_calcs; ++k) {
        const c4b_calc &c = m.calcs[k];
        if ((c.kind == C4B_CALC_SPLICE_PRE || c.kind == C4B_CALC_SPLICE_POST) &&
            (c.param[1] < 0 || c.param[1] >= C4B_SPLICE_TOTAL)) {
            set_error("calc " + std::to_string(k) + ": splice array index out of range");
            return -1;
        }
        if (c.kind >= C4B_CALC_SPLICE_POST && c.kind < C4B_CALC_KIND_TOTAL &&
            (c.param[2] < 0 || c.param[2] >= m.n_shadow_slots)) {
            set_error("calc " + std::to_string(k) + ": shadow slot out of range");
            return -1;
        }
    }
    for (int k = 0; k < m.n_transitions; ++k) {
        const c4b_transition &t = m.transitions[k];
        if (t.input < 0 || t.input >= m.n_states || t.output < 0 || t.output >= m.n_states ||
            t.advance_query < 0 || t.advance_target < 0 || t.calc >= m.n_calcs || t.calc < -1 ||
            t.advance_query > m.max_query_advance || t.advance_target > m.max_target_advance) {
            set_error("transition " + std::to_string(k) + " is malformed");
            return -1;
        }
        if (t.calc >= 0) {
            const c4b_calc &c = m.calcs[t.calc];
            if (c.kind < 0 || c.kind >= C4B_CALC_KIND_TOTAL) {
                // the reference would fall back to a host callback here; we must not
                set_error("calc kind of transition " + std::to_string(k) + " has no device form");
                return -1;
            }
            // a calc that reads a shadow slot must not leave a state that stamps the
            // same slot (then stamp-on-copy == stamp-on-source, DESIGN.md "shadows")
            if (c.kind >= C4B_CALC_SPLICE_POST && m.shadow_start[t.input][c.param[2]] != 0) {
                set_error("shadow read from its own source state is not supported");
                return -1;
            }
        }
    }
    return 0;
}

constexpr int32_t kEndMatrixUnset = (int32_t)0x80808080;  // what cudaMemset(0x80) leaves

// Device copy of a host buffer the caller declared stable (C4B_PAIR_BUFFERS_STABLE): uploaded
// once, found again by (address, bytes) until c4b_engine_forget_buffers.  `pad` zero bytes follow
// the copy (codon reads may run 2 past the end).  nullptr = not cached (cap reached / failure):
// the caller stages the buffer with the batch as usual.
static void *resident_copy(ResidentBuffers *rb, const void *host, size_t bytes, size_t pad) {
    constexpr size_t kCap = (size_t)16 << 30;
    auto key = std::make_pair(host, bytes);
    auto it = rb->map.find(key);
    if (it != rb->map.end()) return it->second;
    if (rb->bytes + bytes + pad > kCap) return nullptr;
    void *dev = nullptr;
    if (cudaMalloc(&dev, bytes + pad + 16) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cudaMemset(dev, 0, bytes + pad + 16) != cudaSuccess ||
        cudaMemcpy(dev, host, bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(dev);
        return nullptr;
    }
    rb->map[key] = dev;
    rb->bytes += bytes + pad + 16;
    return dev;
}

int generic_batch_create(cudaStream_t stream, int64_t *launch_counter, const c4b_model *model,
                         const c4b_scoring *scoring, int n, const c4b_pair *pairs, bool want_path,
                         GenericBatch **out, const int32_t *start_cells, bool end_cells, int sm_count,
                         ResidentBuffers *resident, const GenDevTables *dev) {
    if (check_model(*model)) return -1;
    GenericBatch *g = new GenericBatch();
    g->stream = stream;
    g->launches = launch_counter;
    g->n = n;
    g->want_path = want_path;
    g->tables.model = *model;
    g->tables.scoring = *scoring;
    g->host_pairs.assign(pairs, pairs + n);
    const c4b_model &m = *model;
    const bool splice = model_needs_splice(m);
    // REGION-then-box is only self-consistent for ANYWHERE starts (see tests/test_oracle_golden.py)
    // (cell-callback tables are indexed by region cell: those lattices take the direct PATH pass)
    g->use_region = want_path && m.start_scope == C4B_SCOPE_ANYWHERE && m.end_scope == C4B_SCOPE_ANYWHERE &&
                    !start_cells && !end_cells && !dev;
    g->cmax = 1 + m.n_shadow_slots + (g->use_region ? 2 : 0);
    if (sm_count <= 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
            delete g;
            set_error("no device");
            return -1;
        }
    }
    g->sm_count = sm_count;

    // ---- stage sequences (dedupe by host pointer), splice arrays, blocked lists
    std::map<std::pair<const uint8_t *, int>, size_t> smap;
    std::map<std::pair<const int32_t *, size_t>, size_t> imap;   // (address, length): one pointer may back lists of several lengths
    size_t sbytes = 0, ints = 0;
    int maxQ = 0;
    std::vector<size_t> qo(n), to(n);
    std::vector<size_t> sp(4 * (size_t)n, 0), bq(n, 0), bt(n, 0);
    // buffers found in (or added to) the engine's resident set: device addresses, not offsets
    std::vector<const uint8_t *> qdev(n, nullptr), tdev(n, nullptr);
    std::vector<const int32_t *> spdev(4 * (size_t)n, nullptr);
    auto place_seq = [&](const uint8_t *p, int len) {
        auto key = std::make_pair(p, len);
        auto it = smap.find(key);
        if (it != smap.end()) return it->second;
        const size_t off = sbytes;
        smap[key] = off;
        sbytes += align_up((size_t)len + 4, 16);  // +4: codon reads stay in-bounds
        return off;
    };
    auto place_ints = [&](const int32_t *p, size_t len) {
        const auto key = std::make_pair(p, len);
        auto it = imap.find(key);
        if (it != imap.end()) return it->second;
        const size_t off = ints;
        imap[key] = off;
        ints += align_up(len, 4);
        return off;
    };
    for (int p = 0; p < n; ++p) {
        const c4b_pair &pp = pairs[p];
        if (pp.query_length < 0 || pp.target_length < 0 || pp.query_start < 0 || pp.target_start < 0 ||
            (int64_t)pp.query_start + pp.query_length > pp.query_len ||
            (int64_t)pp.target_start + pp.target_length > pp.target_len || pp.n_blocked < 0 ||
            (pp.n_blocked && (!pp.blocked_query_pos || !pp.blocked_target_pos))) {
            set_error("pair " + std::to_string(p) + ": region outside the sequences");
            delete g;
            return -1;
        }
        const bool stable = resident && (pp.reserved & C4B_PAIR_BUFFERS_STABLE);
        if (stable) {
            qdev[p] = (const uint8_t *)resident_copy(resident, pp.query, (size_t)pp.query_len, 4);
            tdev[p] = (const uint8_t *)resident_copy(resident, pp.target, (size_t)pp.target_len, 4);
        }
        if (!qdev[p]) qo[p] = place_seq(pp.query, pp.query_len);
        if (!tdev[p]) to[p] = place_seq(pp.target, pp.target_len);
        if (splice)
            for (int k = 0; k < 4; ++k) {
                if (!pp.splice[k]) {
                    set_error("model has splice calcs but pair " + std::to_string(p) + " has no splice arrays");
                    delete g;
                    return -1;
                }
                if (stable)
                    spdev[4 * p + k] = (const int32_t *)resident_copy(resident, pp.splice[k],
                                                                      (size_t)pp.target_len * sizeof(int32_t), 16);
                if (spdev[4 * p + k]) continue;
                sp[4 * p + k] = place_ints(pp.splice[k], (size_t)pp.target_len);
            }
        if (pp.n_blocked) {
            bq[p] = place_ints(pp.blocked_query_pos, (size_t)pp.n_blocked);
            bt[p] = place_ints(pp.blocked_target_pos, (size_t)pp.n_blocked);
        }
        maxQ = std::max(maxQ, pp.query_length);
        g->max_t = std::max(g->max_t, pp.target_length);
        g->cells += (int64_t)pp.query_length * pp.target_length;
    }
    int rc = 0;
    rc |= g->d_tables.alloc(1);
    rc |= g->d_seq.alloc(sbytes + 64);
    rc |= g->d_ints.alloc(ints + 4);
    rc |= g->d_full.alloc(n);
    rc |= g->d_box.alloc(n);
    rc |= g->d_out_a.alloc(n);
    rc |= g->d_out_b.alloc(n);
    rc |= g->d_results.alloc(n);
    rc |= g->d_cursor.alloc(1);
    g->threads = (maxQ + 1 > 512) ? 1024 : (maxQ + 1 > 256 ? 512 : 256);
    g->grid = std::max(1, std::min(n, g->sm_count * (1024 / g->threads)));
    const int policy = jit_policy();
    g->use_jit = policy == 1 || (policy == 2 && g->cells >= ((int64_t)1 << 31));
    g->jit_threads = jit_threads_for(maxQ);
    g->max_q = maxQ;
    {
        const char *env = getenv("C4B_JIT_SYSTOLIC");
        g->plain = !start_cells && !end_cells && !dev;   // (set here: the descriptors are filled in below)
        if (g->use_jit && g->plain && !(env && atoi(env) == 0)) {
            // SubOpt blocked cells ride along as per-strip {column, row mask} entries (JIT_SYS_BLK)
            g->h_blk_at.assign((size_t)n + 1, 0);
            for (int p = 0; p < n; ++p) {
                g->sys_blk = g->sys_blk || pairs[p].n_blocked > 0;
                g->h_blk_at[p + 1] = g->h_blk_at[p] + (size_t)pairs[p].n_blocked;
            }
            if (g->sys_blk) {
                g->h_blk_q.resize(g->h_blk_at[n]);
                g->h_blk_t.resize(g->h_blk_at[n]);
                for (int p = 0; p < n; ++p)
                    if (pairs[p].n_blocked) {
                        memcpy(g->h_blk_q.data() + g->h_blk_at[p], pairs[p].blocked_query_pos, (size_t)pairs[p].n_blocked * 4);
                        memcpy(g->h_blk_t.data() + g->h_blk_at[p], pairs[p].blocked_target_pos, (size_t)pairs[p].n_blocked * 4);
                    }
            }
            const bool pack_start = ((int64_t)maxQ + 1) * ((int64_t)g->max_t + 1) < ((int64_t)1 << 31);
            g->sys_score = jit_sys_layout(m, GEN_SCORE, false);
            g->sys_region = jit_sys_layout(m, GEN_REGION, pack_start);
            g->sys_path = jit_sys_layout(m, GEN_PATH, false);
            // The REGION pass exists to keep the PATH records small (optimal.c:368-413).  When the
            // records of the FULL lattices fit comfortably (a third of free memory), one PATH pass
            // over everything is cheaper than REGION + PATH-in-the-box, and gives the same path.
            if (g->use_region && g->sys_path.ok) {
                size_t all = 0, free_b = 0, total_b = 0;
                for (int p = 0; p < n; ++p)
                    all += align_up(GEN_TBS_BYTES(pairs[p].query_length, pairs[p].target_length, g->sys_path.R,
                                                  g->sys_path.chunk), 16);
                const char *d = getenv("C4B_GENERIC_DIRECT_PATH");
                if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && (d ? atoi(d) != 0 : all < free_b / 3))
                    g->use_region = false;
            }
        }
    }
    g->ring_ctas = g->grid;
    if (g->use_jit) g->ring_ctas = std::max(g->ring_ctas, std::min(n, g->sm_count * (2048 / g->jit_threads)));
    const int depth = m.max_target_advance + m.max_query_advance + 1;
    g->ring_stride = align_up((size_t)depth * (maxQ + 1) * m.n_states * g->cmax, 4);
    rc |= g->d_ring.alloc(g->ring_stride * g->ring_ctas);
    if (rc) { delete g; return -1; }
    // small jobs: gather into one host block, one copy each; large batches: clear on the device and
    // copy every buffer straight from where the caller holds it (no second pass over 100s of MB)
    const bool direct = sbytes + ints * sizeof(int32_t) > ((size_t)1 << 20);
    std::vector<uint8_t> hs;
    std::vector<int32_t> hi;
    if (!direct) {
        hs.assign(sbytes + 64, 0);
        for (auto &kv : smap) memcpy(hs.data() + kv.second, kv.first.first, (size_t)kv.first.second);
        hi.assign(ints + 4, 0);
        for (auto &kv : imap) memcpy(hi.data() + kv.second, kv.first.first, kv.first.second * sizeof(int32_t));
    }
    g->h_full.resize(n);
    for (int p = 0; p < n; ++p) {
        const c4b_pair &pp = pairs[p];
        GenPair &G = g->h_full[p];
        G.q = qdev[p] ? qdev[p] : g->d_seq.p + qo[p];
        G.t = tdev[p] ? tdev[p] : g->d_seq.p + to[p];
        for (int k = 0; k < 4; ++k)
            G.splice[k] = !splice ? nullptr : (spdev[4 * p + k] ? spdev[4 * p + k] : g->d_ints.p + sp[4 * p + k]);
        G.n_blocked = pp.n_blocked;
        G.blk_q = pp.n_blocked ? g->d_ints.p + bq[p] : nullptr;
        G.blk_t = pp.n_blocked ? g->d_ints.p + bt[p] : nullptr;
        G.q_start = pp.query_start; G.t_start = pp.target_start;
        G.Q = pp.query_length; G.T = pp.target_length;
        G.blk_dq = 0; G.blk_dt = 0;
        G.tb = nullptr;
        G.start_cells = (dev && dev->start) ? dev->start[p] : nullptr;
        G.end_cells = (dev && dev->end) ? dev->end[p] : nullptr;
        G.out_index = p;
        G.tb_rows = 0;
        G.tb_chunk = 0;
        G.blk = nullptr; G.blk_off = nullptr;
    }
    if ((start_cells || end_cells) && n == 1) {
        const size_t cells = ((size_t)pairs[0].query_length + 1) * ((size_t)pairs[0].target_length + 1) *
                             (1 + (size_t)m.n_shadow_slots);
        bool ok = true;
        if (end_cells) {
            ok = ok && !g->d_endm.alloc(cells);
            ok = ok && cudaMemsetAsync(g->d_endm.p, 0x80, cells * sizeof(int32_t), stream) == cudaSuccess;
            g->h_full[0].end_cells = g->d_endm.p;
        }
        if (start_cells) {
            ok = ok && !g->d_startc.alloc(cells);
            ok = ok && cudaMemcpyAsync(g->d_startc.p, start_cells, cells * sizeof(int32_t), cudaMemcpyHostToDevice,
                                       stream) == cudaSuccess;
            g->h_full[0].start_cells = g->d_startc.p;
        }
        if (!ok) {
            set_error("staging the cell-callback tables failed");
            delete g;
            return -1;
        }
    }
    bool staged = true;
    if (direct) {
        // gathered by the host threads into this thread's pinned scratch, then two DMAs: thousands of
        // copies from pageable caller buffers (one per sequence and splice array: ~15 us each, staged by the
        // driver one after the other) were a third of an end-to-end protein2genome step
        const size_t seq_bytes = align_up(sbytes + 64, 256);
        uint8_t *scratch = tl_gen_scratch.get(seq_bytes + (ints + 4) * sizeof(int32_t));
        if (!scratch) {
            set_error("pinned staging allocation failed");
            delete g;
            return -1;
        }
        int32_t *hints = reinterpret_cast<int32_t *>(scratch + seq_bytes);
        typedef std::pair<std::pair<const uint8_t *, int>, size_t> SeqItem;
        typedef std::pair<std::pair<const int32_t *, size_t>, size_t> IntItem;
        std::vector<SeqItem> slist(smap.begin(), smap.end());
        std::vector<IntItem> ilist(imap.begin(), imap.end());
        parallel_for((int)(slist.size() + ilist.size()), [&](int k) {
            if (k < (int)slist.size()) {   // slot = len + 4 rounded up to 16: the tail reads as zero
                const size_t len = (size_t)slist[k].first.second, slot = align_up(len + 4, 16);
                memcpy(scratch + slist[k].second, slist[k].first.first, len);
                memset(scratch + slist[k].second + len, 0, slot - len);
            } else {
                const IntItem &it = ilist[k - slist.size()];
                const size_t len = it.first.second, slot = align_up(len, 4);
                memcpy(hints + it.second, it.first.first, len * sizeof(int32_t));
                memset(hints + it.second + len, 0, (slot - len) * sizeof(int32_t));
            }
        });
        memset(scratch + sbytes, 0, 64);
        memset(hints + ints, 0, 4 * sizeof(int32_t));
        staged = cudaMemcpyAsync(g->d_seq.p, scratch, sbytes + 64, cudaMemcpyHostToDevice, stream) == cudaSuccess &&
                 cudaMemcpyAsync(g->d_ints.p, hints, (ints + 4) * sizeof(int32_t), cudaMemcpyHostToDevice, stream) ==
                     cudaSuccess;
        tl_gen_scratch.mark(stream);
    } else {
        staged = cudaMemcpyAsync(g->d_seq.p, hs.data(), sbytes + 64, cudaMemcpyHostToDevice, stream) == cudaSuccess &&
                 cudaMemcpyAsync(g->d_ints.p, hi.data(), (ints + 4) * 4, cudaMemcpyHostToDevice, stream) == cudaSuccess;
    }
    if (!staged ||
        cudaMemcpyAsync(g->d_tables.p, &g->tables, sizeof(GenTables), cudaMemcpyHostToDevice, stream) != cudaSuccess ||
        cudaMemcpyAsync(g->d_full.p, g->h_full.data(), n * sizeof(GenPair), cudaMemcpyHostToDevice, stream) != cudaSuccess ||
        cudaStreamSynchronize(stream) != cudaSuccess) {
        set_error("staging the generic batch failed");
        delete g;
        return -1;
    }
    cudaEventCreate(&g->ev_a);
    cudaEventCreate(&g->ev_b);
    tmark("generic: staged");
    *out = g;
    return 0;
}

// SubOpt blocked cells (src/c4/subopt.c:250-338: region coordinates of the DESTINATION cell) of every
// lattice in `lat` -> per lane strip {column, row mask} sorted by column, for the systolic kernel with R
// rows per lane; lat[p].blk / blk_off are set (null for a lattice without blocked cells).  A lattice may
// be a box inside the caller's region: list coordinates are lattice coordinates + (blk_dq, blk_dt).
static int generic_build_sys_blk(GenericBatch *g, std::vector<GenPair> &lat, int R, DevBuf<int2> &d_blk,
                                 DevBuf<int32_t> &d_off) {
    const int n = g->n;
    std::vector<int2> h_blk;
    std::vector<int32_t> h_off;
    std::vector<size_t> seg(n, (size_t)-1);
    std::vector<std::array<int32_t, 3>> cells;   // (lane strip, column, row in strip)
    for (int p = 0; p < n; ++p) {
        GenPair &L = lat[p];
        L.blk = nullptr; L.blk_off = nullptr;
        const size_t nb = g->h_blk_at[p + 1] - g->h_blk_at[p];
        if (!nb || (L.Q == 0 && L.T == 0)) continue;
        const int32_t *bq = g->h_blk_q.data() + g->h_blk_at[p], *bt = g->h_blk_t.data() + g->h_blk_at[p];
        const int nstrips = 32 * ((L.Q + 32 * R) / (32 * R));
        cells.clear();
        for (size_t k = 0; k < nb; ++k) {
            const int64_t i = (int64_t)bq[k] - L.blk_dq, j = (int64_t)bt[k] - L.blk_dt;
            if (i < 0 || i > L.Q || j < 0 || j > L.T) continue;   // never looked at
            cells.push_back({(int32_t)(i / R), (int32_t)j, (int32_t)(i % R)});
        }
        std::sort(cells.begin(), cells.end());
        seg[p] = h_off.size();
        size_t c = 0;
        for (int strip = 0; strip < nstrips; ++strip) {
            h_off.push_back((int32_t)h_blk.size());
            while (c < cells.size() && cells[c][0] == strip) {
                const int j = cells[c][1];
                uint32_t mask = 0;
                for (; c < cells.size() && cells[c][0] == strip && cells[c][1] == j; ++c) mask |= 1u << cells[c][2];
                h_blk.push_back(make_int2(j, (int)mask));
            }
        }
        h_off.push_back((int32_t)h_blk.size());
        if (h_blk.size() > (size_t)INT32_MAX / 2) {
            set_error("too many SubOpt blocked cells in one batch");
            return -1;
        }
    }
    d_blk.release(); d_off.release();
    if (d_blk.alloc(h_blk.size() + 1) || d_off.alloc(h_off.size() + 1)) return -1;
    if (!h_blk.empty())
        C4B_CUDA(cudaMemcpyAsync(d_blk.p, h_blk.data(), h_blk.size() * sizeof(int2), cudaMemcpyHostToDevice, g->stream));
    if (!h_off.empty())
        C4B_CUDA(cudaMemcpyAsync(d_off.p, h_off.data(), h_off.size() * sizeof(int32_t), cudaMemcpyHostToDevice, g->stream));
    C4B_CUDA(cudaStreamSynchronize(g->stream));   // (the host vectors die here)
    for (int p = 0; p < n; ++p)
        if (seg[p] != (size_t)-1) { lat[p].blk = d_blk.p; lat[p].blk_off = d_off.p + seg[p]; }
    return 0;
}

// win / wins: the column-window variants of the systolic kernel (JIT_SYS_WIN), one GenWin per lattice
static int generic_launch_fill(GenericBatch *g, const GenPair *pairs, int count, GenOut *outs, int mode,
                               int win = 0, const GenWin *wins = nullptr) {
    if (!count) return 0;
    C4B_CUDA(cudaMemsetAsync(g->d_cursor.p, 0, sizeof(int), g->stream));
    const SysLayout &SL = mode == GEN_SCORE ? g->sys_score : (mode == GEN_REGION ? g->sys_region : g->sys_path);
    if (win && !(g->use_jit && SL.ok)) {
        set_error("internal: windowed PATH pass without the systolic specialisation");
        return -1;
    }
    if (g->use_jit && SL.ok) {
        const bool pack_start = mode == GEN_REGION;
        // one CTA per lattice, one warp per strip of 32 R rows (up to 8, then round-robin);
        // strips hand their bottom rows over through L2: 2 x (T + 1) x NSEND words per lattice
        int maxQ = 0, maxT = 0;
        for (int p = 0; p < g->n; ++p) { maxQ = std::max(maxQ, g->h_full[p].Q); maxT = std::max(maxT, g->h_full[p].T); }
        const int nsweeps = (maxQ + 32 * SL.R) / (32 * SL.R);
        const int warps = jit_sys_warps(SL, maxQ);
        if (JitKernel *jk = jit_get_sys(g->tables.model, mode, pack_start, SL, warps, win, g->sys_blk)) {
            const size_t nsend = std::max<size_t>(1, SL.sendD.size());
            const size_t stride = nsweeps > 1 ? align_up(2 * ((size_t)maxT + 1) * nsend, 4) : 0;
            if (stride * (size_t)count > g->d_top.n) {
                g->d_top.release();
                if (g->d_top.alloc(stride * (size_t)count + 4)) return -1;
            }
            int32_t *top = g->d_top.p;
            void *args[] = {(void *)&pairs, (void *)&count, (void *)&outs, (void *)&g->d_tables.p, (void *)&top,
                            (void *)&stride, (void *)&wins};
            C4B_CUDA(cudaLaunchKernel((const void *)jk->kern, dim3(count), dim3(32 * warps), args, 0, g->stream));
            (*g->launches)++;
            g->kernel_used = "generic_jit_systolic";
            return 0;
        }
        if (win || !getenv("C4B_JIT_FALLBACK")) {
            set_error("systolic model specialisation failed (reason on stderr); set C4B_JIT_SYSTOLIC=0 to use the "
                      "thread-per-row specialisation");
            return -1;
        }
    }
    if (g->use_jit) {
        // the lattice ring goes to shared memory when DEPTH columns of the longest query fit
        const bool pack_start = mode == GEN_REGION && ((int64_t)g->max_q + 1) * ((int64_t)g->max_t + 1) < ((int64_t)1 << 31);
        const size_t ring_bytes =
            (size_t)jit_ring_words_per_row(g->tables.model, mode, pack_start) * (g->max_q + 1) * 4;
        const bool smem_ring = ring_bytes <= (size_t)kJitSmemRingBytes && !getenv("C4B_JIT_GLOBAL_RING");
        if (JitKernel *jk = jit_get(g->tables.model, mode, g->jit_threads, smem_ring, pack_start)) {
            const int grid = std::max(1, std::min(std::min(count, g->ring_ctas), g->sm_count * jk->blocks_per_sm));
            const GenTables *tables = g->d_tables.p;
            int32_t *ring = g->d_ring.p;
            size_t stride = g->ring_stride;
            int *cursor = g->d_cursor.p;
            void *args[] = {(void *)&pairs, (void *)&count, (void *)&outs, (void *)&tables,
                            (void *)&ring, (void *)&stride, (void *)&cursor};
            C4B_CUDA(cudaLaunchKernel((const void *)jk->kern, dim3(grid), dim3(jk->threads), args,
                                      smem_ring ? ring_bytes : 0, g->stream));
            (*g->launches)++;
            g->kernel_used = "generic_jit";
            return 0;
        }
        // The specialised kernel was asked for and is not available (no NVRTC, compile or load
        // failure): that is a 10x performance cliff, so it is an error unless the caller opted in
        // to the interpreter kernel (still a device kernel; C4B_JIT_FALLBACK=1).
        if (!getenv("C4B_JIT_FALLBACK")) {
            set_error("model specialisation failed (reason on stderr); set C4B_JIT_FALLBACK=1 to run the "
                      "interpreter kernel instead, or C4B_GENERIC_JIT=0 to never specialise");
            return -1;
        }
        g->use_jit = false;
    }
    const int grid = std::min(g->grid, count);
    // small lattices (BSDP's region fills): the ring fits the CTA's shared memory, and the fill is
    // a chain of dependent ring accesses per cell -- shared-memory latency instead of L2's
    constexpr size_t kSmemRingMax = 160 * 1024;
    const size_t ring_bytes = g->ring_stride * sizeof(int32_t);
    // (taken when it does not cost residency: few lattices, or a ring small enough for 4 CTAs per SM)
    const bool smem_ring = ring_bytes <= kSmemRingMax && (count <= g->sm_count || ring_bytes <= 20 * 1024);
    if (smem_ring)   // per device and cheap: set on every launch that needs it (engines on several devices)
        C4B_CUDA(cudaFuncSetAttribute(generic_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kSmemRingMax));
    generic_fill_kernel<<<grid, g->threads, smem_ring ? ring_bytes : 0, g->stream>>>(
        pairs, count, outs, g->d_tables.p, mode, g->d_ring.p, g->ring_stride, g->d_cursor.p, smem_ring ? 1 : 0);
    C4B_CUDA(cudaGetLastError());
    (*g->launches)++;
    return 0;
}

// PATH for lattices whose whole record does not fit `budget`: pass 1 (score) leaves a checkpoint of
// the systolic register lattice after every `wcols` columns; then rounds of {refill the window under
// each traceback cursor from the checkpoint to its left, walk it} until every cursor reached START.
// Memory: checkpoints (T / wcols per strip) + one window of records per lattice.  d_jobs / d_box /
// d_ops_slots are those of generic_batch_run; results land where the ordinary traceback puts them.
static int generic_batch_run_windows(GenericBatch *g, const std::vector<GenPair> &box, const std::vector<int32_t> &big,
                                     size_t budget, c4b_score threshold) {
    cudaStream_t st = g->stream;
    const int nb = (int)big.size();
    const SysLayout &LS = g->sys_score, &LP = g->sys_path;
    if (LS.R != LP.R || LS.VW != LP.VW || LS.AQ != LP.AQ || LS.VOff != LP.VOff) {
        set_error("internal: score and PATH specialisations disagree on the register lattice");
        return -1;
    }
    const size_t R = (size_t)LP.R, ckw = (size_t)(LP.AQ + LP.R) * LP.VW;
    auto sweeps_of = [&](int p) { return ((size_t)box[p].Q + 32 * R) / (32 * R); };
    auto win_bytes = [&](int p, size_t wc) { return align_up(sweeps_of(p) * (wc + 31) * 32 * (size_t)LP.chunk, 16); };
    auto ck_words = [&](int p, size_t wc) { return ((size_t)box[p].T / wc) * sweeps_of(p) * ckw * 32; };
    size_t wcols = 4096;
    if (const char *env = getenv("C4B_GENERIC_WINDOW_COLS")) wcols = std::max<size_t>(32, (size_t)atoll(env));
    while (wcols & (wcols - 1)) wcols &= wcols - 1;   // a power of two (the kernel masks with it)
    auto totals = [&](size_t wc, size_t *wb, size_t *cb) {
        *wb = 0; *cb = 0;
        for (int p : big) { *wb += win_bytes(p, wc); *cb += ck_words(p, wc) * 4; }
    };
    size_t wb = 0, cb = 0;
    totals(wcols, &wb, &cb);
    while (wb > budget / 2 && wcols > 32) { wcols >>= 1; totals(wcols, &wb, &cb); }
    while (cb > budget / 2 && wb * 2 <= budget / 2) { wcols <<= 1; totals(wcols, &wb, &cb); }
    if (wb > budget / 2 || cb > budget / 2) {
        set_error("traceback of pair " + std::to_string(big[0]) + " exceeds the device memory budget even as column "
                  "windows (" + std::to_string(wb >> 20) + " MB of records + " + std::to_string(cb >> 20) +
                  " MB of checkpoints)");
        return -1;
    }
    g->windowed_lattices = nb;
    g->window_cols = (int)wcols;
    g->d_wpairs.release(); g->d_wins.release(); g->d_wwalk.release(); g->d_wbig.release(); g->d_wck.release();
    g->d_wtb.release();
    if (g->d_wpairs.alloc(nb) || g->d_wins.alloc(nb) || g->d_wwalk.alloc(nb) || g->d_wbig.alloc(nb) ||
        g->d_wck.alloc(cb / 4 + 32) || g->d_wtb.alloc(wb + 16))
        return -1;
    std::vector<GenPair> wp(nb);
    std::vector<GenWin> wins(nb);
    {
        size_t tb_cur = 0, ck_cur = 0;
        for (int k = 0; k < nb; ++k) {
            const int p = big[k];
            wp[k] = box[p];
            wp[k].tb = g->d_wtb.p + tb_cur;
            wp[k].tb_rows = LP.R; wp[k].tb_chunk = LP.chunk;
            tb_cur += win_bytes(p, wcols);
            wins[k].ck = g->d_wck.p + ck_cur;
            ck_cur += ck_words(p, wcols);
            wins[k].wcols = (int32_t)wcols; wins[k].c0 = 0; wins[k].c1 = box[p].T; wins[k].nsweeps = 0; wins[k].reserved = 0;
        }
    }
    C4B_CUDA(cudaMemcpyAsync(g->d_wpairs.p, wp.data(), nb * sizeof(GenPair), cudaMemcpyHostToDevice, st));
    C4B_CUDA(cudaMemcpyAsync(g->d_wins.p, wins.data(), nb * sizeof(GenWin), cudaMemcpyHostToDevice, st));
    C4B_CUDA(cudaMemcpyAsync(g->d_wbig.p, big.data(), nb * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    // pass 1 inside the boxes: END cell (checked against the REGION pass like any box refill) + checkpoints
    if (generic_launch_fill(g, g->d_wpairs.p, nb, g->d_out_b.p, GEN_SCORE, 1, g->d_wins.p)) return -1;
    generic_window_walk_init_kernel<<<(nb + 63) / 64, 64, 0, st>>>(g->d_wpairs.p, g->d_out_b.p, g->d_out_a.p, g->d_jobs.p,
                                                                   g->d_wbig.p, nb, g->d_tables.p, threshold,
                                                                   g->d_wwalk.p, g->d_results.p);
    (*g->launches)++;
    C4B_CUDA(cudaGetLastError());
    std::vector<GenWalk> walk(nb);
    int64_t round_cap = 16;
    for (int p : big) round_cap += 2 * ((int64_t)box[p].T / (int64_t)wcols + 1) + 2 * (int64_t)sweeps_of(p);
    for (int64_t round = 0;; ++round) {
        C4B_CUDA(cudaMemcpyAsync(walk.data(), g->d_wwalk.p, nb * sizeof(GenWalk), cudaMemcpyDeviceToHost, st));
        C4B_CUDA(cudaStreamSynchronize(st));
        int active = 0;
        for (int k = 0; k < nb; ++k) {
            GenWin &w = wins[k];
            w.nsweeps = 0;
            if (walk[k].done) continue;
            ++active;
            w.c0 = (int32_t)((size_t)walk[k].j / wcols * wcols);
            w.c1 = walk[k].j;                                         // the path never moves right or down
            w.nsweeps = (int32_t)((size_t)walk[k].i / (32 * R) + 1);
        }
        if (!active) break;
        if (round >= round_cap) {
            set_error("internal: windowed traceback did not terminate");
            return -1;
        }
        C4B_CUDA(cudaMemcpyAsync(g->d_wins.p, wins.data(), nb * sizeof(GenWin), cudaMemcpyHostToDevice, st));
        if (generic_launch_fill(g, g->d_wpairs.p, nb, g->d_out_b.p, GEN_PATH, 2, g->d_wins.p)) return -1;
        generic_window_walk_kernel<<<(nb + 63) / 64, 64, 0, st>>>(g->d_wpairs.p, g->d_out_b.p, g->d_jobs.p, g->d_wbig.p,
                                                                  g->d_wins.p, nb, g->d_tables.p, g->d_wwalk.p,
                                                                  g->d_results.p, g->d_ops_slots.p);
        (*g->launches)++;
        C4B_CUDA(cudaGetLastError());
        g->window_rounds++;
    }
    return 0;
}

int generic_batch_run(GenericBatch *g, c4b_score threshold) {
    cudaStream_t st = g->stream;
    const int n = g->n;
    const int S = g->tables.model.n_states;
    if (!n) return 0;
    // SubOpt blocked cells in the systolic kernel's form, for the pass over the full lattices
    if (g->sys_blk) {
        const SysLayout &first = !g->want_path ? g->sys_score : (g->use_region ? g->sys_region : g->sys_path);
        if (first.ok) {
            if (generic_build_sys_blk(g, g->h_full, first.R, g->d_sblk_a, g->d_sblk_off_a)) return -1;
            C4B_CUDA(cudaMemcpyAsync(g->d_full.p, g->h_full.data(), n * sizeof(GenPair), cudaMemcpyHostToDevice, st));
        }
    }
    C4B_CUDA(cudaEventRecord(g->ev_a, st));
    if (!g->want_path) {
        if (generic_launch_fill(g, g->d_full.p, n, g->d_out_a.p, GEN_SCORE)) return -1;
        C4B_CUDA(cudaEventRecord(g->ev_b, st));
        generic_score_results_kernel<<<(n + 127) / 128, 128, 0, st>>>(g->d_full.p, g->d_out_a.p, n, g->d_results.p);
        (*g->launches)++;
        C4B_CUDA(cudaGetLastError());
        return 0;
    }
    // 1) where is the alignment?  (FIND_REGION shadows)
    std::vector<GenPair> box = g->h_full;
    std::vector<GenOut> reg(n);
    if (g->use_region) {
        if (generic_launch_fill(g, g->d_full.p, n, g->d_out_a.p, GEN_REGION)) return -1;
        C4B_CUDA(cudaMemcpyAsync(reg.data(), g->d_out_a.p, n * sizeof(GenOut), cudaMemcpyDeviceToHost, st));
        C4B_CUDA(cudaStreamSynchronize(st));
        tmark("generic: region pass done");
        for (int p = 0; p < n; ++p) {
            GenPair &B = box[p];
            const GenOut &o = reg[p];
            if (o.flags || o.score < threshold) { B.Q = 0; B.T = 0; continue; }
            B.blk_dq += o.start_i; B.blk_dt += o.start_j;
            B.q_start += o.start_i; B.t_start += o.start_j;
            B.Q = o.end_i - o.start_i; B.T = o.end_j - o.start_j;
        }
    }
    // 2) PATH fill inside the boxes, chunked to the traceback arena
    const bool sys_tb = g->use_jit && g->sys_path.ok;   // the systolic kernel's record layout
    auto tb_bytes = [&](int p) {
        return sys_tb ? GEN_TBS_BYTES(box[p].Q, box[p].T, g->sys_path.R, g->sys_path.chunk)
                      : GEN_TB_BYTES(box[p].Q, box[p].T, S);
    };
    if (sys_tb)
        for (int p = 0; p < n; ++p) { box[p].tb_rows = g->sys_path.R; box[p].tb_chunk = g->sys_path.chunk; }
    size_t budget = 256ull << 20;  // small jobs (BSDP region fills) never need to ask the driver
    size_t window_budget = budget; // what the windowed route may use (the real budget, also when testing)
    {
        size_t all = 0;
        for (int p = 0; p < n; ++p) all += align_up(tb_bytes(p), 16);
        const char *env = getenv("C4B_GENERIC_TB_BUDGET_KB");
        if (all > budget || env) {
            size_t free_b = 0, total_b = 0;
            C4B_CUDA(cudaMemGetInfo(&free_b, &total_b));
            if (free_b > (3ull << 30)) budget = (free_b - (2ull << 30)) / 2;
        }
        window_budget = budget;
        if (env)   // testing: force the chunked / windowed routes (the budget only decides the route)
            budget = std::max<size_t>(256, (size_t)atoll(env) << 10);
    }
    std::vector<size_t> tb_off(n);
    std::vector<Chunk> chunks;
    std::vector<GenJob> jobs(n);
    std::vector<int32_t> big;   // lattices whose record alone exceeds the budget: column windows
    size_t cur = 0, arena = 0;
    int64_t ops_cursor = 0;
    int begin = 0;
    for (int p = 0; p < n; ++p) {
        const size_t need = align_up(tb_bytes(p), 16);
        GenJob &J = jobs[p];
        J.pair = p; J.result = p; J.expect = g->use_region ? 1 : 0; J.score_slot = p;
        J.ops_cap = (int32_t)std::min<int64_t>((int64_t)box[p].Q + box[p].T + 4, INT32_MAX);
        J.ops_off = ops_cursor; J.reserved = 0;
        ops_cursor += J.ops_cap;
        if (need > budget) {
            // The reference recurses through checkpoint rows here (optimal.c:183-345).  The systolic
            // kernel checkpoints its register lattice by column instead and refills one window at a
            // time under the traceback cursor (generic_batch_run_windows).
            if (!(sys_tb && g->plain && g->sys_score.ok)) {
                set_error("traceback box of pair " + std::to_string(p) + " exceeds the device memory budget "
                          "(column windows need the systolic specialisation: a batch of >= 2^31 cells or C4B_GENERIC_JIT=1, no cell tables)");
                return -1;
            }
            if (begin < p) chunks.push_back({begin, p});
            begin = p + 1;
            cur = 0;
            tb_off[p] = 0;
            big.push_back(p);
            continue;
        }
        if (cur + need > budget) { chunks.push_back({begin, p}); begin = p; cur = 0; }
        tb_off[p] = cur;
        cur += need;
        arena = std::max(arena, cur);
    }
    if (begin < n) chunks.push_back({begin, n});
    g->d_tb.release(); g->d_jobs.release(); g->d_ops_slots.release(); g->d_ops_packed.release();
    g->d_new_off.release();
    if (g->d_tb.alloc(arena + 16) || g->d_jobs.alloc(n) || g->d_ops_slots.alloc(2 * (size_t)ops_cursor + 2) ||
        g->d_ops_packed.alloc(2 * (size_t)ops_cursor + 2) || g->d_new_off.alloc((size_t)n + 1))
        return -1;
    for (int p = 0; p < n; ++p) box[p].tb = g->d_tb.p + tb_off[p];
    // (the boxes moved the lattice origins: their blocked-cell entries are built afresh)
    if (g->sys_blk && sys_tb && generic_build_sys_blk(g, box, g->sys_path.R, g->d_sblk_b, g->d_sblk_off_b)) return -1;
    C4B_CUDA(cudaMemcpyAsync(g->d_box.p, box.data(), n * sizeof(GenPair), cudaMemcpyHostToDevice, st));
    C4B_CUDA(cudaMemcpyAsync(g->d_jobs.p, jobs.data(), n * sizeof(GenJob), cudaMemcpyHostToDevice, st));
    for (const Chunk &c : chunks) {
        const int cnt = c.end - c.begin;
        if (generic_launch_fill(g, g->d_box.p + c.begin, cnt, g->d_out_b.p, GEN_PATH)) return -1;
        generic_traceback_kernel<<<(cnt + 63) / 64, 64, 0, st>>>(g->d_box.p, g->d_out_b.p, g->d_out_a.p,
                                                                g->d_jobs.p + c.begin, cnt, g->d_tables.p,
                                                                threshold, g->d_results.p, g->d_ops_slots.p);
        (*g->launches)++;
        C4B_CUDA(cudaGetLastError());
    }
    if (!big.empty() && generic_batch_run_windows(g, box, big, window_budget, threshold)) return -1;
    C4B_CUDA(cudaEventRecord(g->ev_b, st));
    apply_threshold_kernel<<<(n + 127) / 128, 128, 0, st>>>(g->d_results.p, n, threshold);
    ops_scan_kernel<<<1, 1024, 0, st>>>(g->d_results.p, n, g->d_new_off.p, g->d_new_off.p + n);
    ops_compact_kernel<<<n, 64, 0, st>>>(g->d_results.p, n, g->d_new_off.p, g->d_ops_slots.p, g->d_ops_packed.p);
    (*g->launches) += 3;
    C4B_CUDA(cudaGetLastError());
    // host_pairs/box vectors die here; the copies above were from pageable memory
    tmark("generic: path pass queued");
    C4B_CUDA(cudaStreamSynchronize(st));
    tmark("generic: path pass done");
    return 0;
}

int generic_batch_fetch(GenericBatch *g, c4b_result *results, int32_t *ops, int64_t ops_capacity) {
    cudaStream_t st = g->stream;
    const int n = g->n;
    if (!n) return 0;
    int64_t total = 0;
    C4B_CUDA(cudaMemcpyAsync(results, g->d_results.p, n * sizeof(c4b_result), cudaMemcpyDeviceToHost, st));
    if (g->want_path)
        C4B_CUDA(cudaMemcpyAsync(&total, g->d_new_off.p + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    C4B_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < n; ++k)
        if (results[k].status >= 2) {
            set_error("internal: lattice " + std::to_string(k) + " failed with status " +
                      std::to_string(results[k].status));
            return -1;
        }
    if (!g->want_path) return 0;
    if (total > ops_capacity) {
        set_error("ops buffer too small: need capacity " + std::to_string(total));
        return -3;
    }
    if (total)
        C4B_CUDA(cudaMemcpy(ops, g->d_ops_packed.p, 2 * (size_t)total * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return 0;
}

int64_t generic_batch_cells(const GenericBatch *g) { return g->cells; }
const void *generic_batch_device_results(const GenericBatch *g) { return g->d_results.p; }

double generic_batch_fill_ms(GenericBatch *g) {
    float f = 0;
    cudaStreamSynchronize(g->stream);
    if (cudaEventElapsedTime(&f, g->ev_a, g->ev_b) != cudaSuccess) return -1;
    return f;
}

void generic_batch_destroy(GenericBatch *g) { delete g; }

}  // namespace c4b
