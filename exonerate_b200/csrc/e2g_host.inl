// e2g_host.inl -- host side of the est2genome systolic path (included by
// c4b200.cu after DevBuf / the ops packing kernels are defined).
namespace c4b {

// Is the closed model the est2genome template of SURVEY.md §8a?  Every transition
// is checked (state roles are derived from the tables, the ORDER is required).
static bool analyze_est2genome(const c4b_model &m, const c4b_scoring &sc, E2gModel *em) {
    if (m.n_states != 10 || m.n_transitions != 24 || m.n_shadow_slots != 1) return false;
    if (m.max_query_advance != 1 || m.max_target_advance != 2) return false;
    if (m.start_scope != C4B_SCOPE_ANYWHERE || m.end_scope != C4B_SCOPE_ANYWHERE) return false;
    const c4b_transition *t = m.transitions;
    const int S = m.start_state, E = m.end_state;
    const int MR = t[6].input, MF = t[11].input, IR = t[7].output, DR = t[8].output;
    const int IF = t[12].output, DF = t[13].output, NR = t[0].output, NF = t[3].output;
    enum { cNONE, cMATCH, cOPEN, cEXT, cPRE, cPOST };
    struct Want { int in, out, aq, at, calc, site; };
    const Want want[24] = {
        {MR, NR, 0, 2, cPRE, C4B_SPLICE_3_REVERSE}, {NR, NR, 0, 1, cNONE, 0}, {NR, MR, 0, 2, cPOST, C4B_SPLICE_5_REVERSE},
        {MF, NF, 0, 2, cPRE, C4B_SPLICE_5_FORWARD}, {NF, NF, 0, 1, cNONE, 0}, {NF, MF, 0, 2, cPOST, C4B_SPLICE_3_FORWARD},
        {MR, MR, 1, 1, cMATCH, 0}, {MR, IR, 1, 0, cOPEN, 0}, {MR, DR, 0, 1, cOPEN, 0}, {IR, IR, 1, 0, cEXT, 0},
        {DR, DR, 0, 1, cEXT, 0}, {MF, MF, 1, 1, cMATCH, 0}, {MF, IF, 1, 0, cOPEN, 0}, {MF, DF, 0, 1, cOPEN, 0},
        {IF, IF, 1, 0, cEXT, 0}, {DF, DF, 0, 1, cEXT, 0}, {IR, MR, 0, 0, cNONE, 0}, {DR, MR, 0, 0, cNONE, 0},
        {S, MR, 0, 0, cNONE, 0}, {IF, MF, 0, 0, cNONE, 0}, {DF, MF, 0, 0, cNONE, 0}, {S, MF, 0, 0, cNONE, 0},
        {MR, E, 0, 0, cNONE, 0}, {MF, E, 0, 0, cNONE, 0}};
    int open = 0, ext = 0, intron_open = 0;
    bool have_open = false, have_ext = false, have_pre = false;
    for (int k = 0; k < 24; ++k) {
        const Want &w = want[k];
        if (t[k].input != w.in || t[k].output != w.out || t[k].advance_query != w.aq ||
            t[k].advance_target != w.at)
            return false;
        if (w.calc == cNONE) {
            if (t[k].calc >= 0) return false;
            continue;
        }
        if (t[k].calc < 0) return false;
        const c4b_calc &c = m.calcs[t[k].calc];
        switch (w.calc) {
        case cMATCH:
            if (c.kind != C4B_CALC_MATCH_DNA || c.protect != 0 || t[k].label != C4B_LABEL_MATCH) return false;
            break;
        case cOPEN:
            if (c.kind != C4B_CALC_CONST || c.protect != 0) return false;
            if (have_open && c.param[0] != open) return false;
            open = c.param[0]; have_open = true;
            break;
        case cEXT:
            if (c.kind != C4B_CALC_CONST || c.protect != 0) return false;
            if (have_ext && c.param[0] != ext) return false;
            ext = c.param[0]; have_ext = true;
            break;
        case cPRE:
            if (c.kind != C4B_CALC_SPLICE_PRE || c.param[1] != w.site || c.protect != C4B_PROTECT_UNDERFLOW) return false;
            if (have_pre && c.param[0] != intron_open) return false;
            intron_open = c.param[0]; have_pre = true;
            break;
        case cPOST:
            if (c.kind != C4B_CALC_SPLICE_POST || c.param[1] != w.site || c.param[2] != 0 ||
                c.protect != C4B_PROTECT_UNDERFLOW)
                return false;
            break;
        }
    }
    // the one shadow slot is stamped with the target position when leaving a match state
    for (int s = 0; s < m.n_states; ++s) {
        const int want_stamp = (s == MF || s == MR) ? 1 : 0;
        if (m.shadow_start[s][0] != want_stamp) return false;
    }
    if (open >= 0 || ext >= 0) return false;
    em->open = open; em->ext = ext; em->intron_open = intron_open;
    em->min_intron = sc.min_intron; em->max_intron = sc.max_intron; em->one = 1;
    // x = 0 forward, 1 reverse
    em->tNopen[0] = 3; em->tNloop[0] = 4; em->tNclose[0] = 5; em->tMatch[0] = 11; em->tIopen[0] = 12;
    em->tDopen[0] = 13; em->tIext[0] = 14; em->tDext[0] = 15; em->tI2M[0] = 19; em->tD2M[0] = 20;
    em->tS2M[0] = 21; em->tM2E[0] = 23;
    em->tNopen[1] = 0; em->tNloop[1] = 1; em->tNclose[1] = 2; em->tMatch[1] = 6; em->tIopen[1] = 7;
    em->tDopen[1] = 8; em->tIext[1] = 9; em->tDext[1] = 10; em->tI2M[1] = 16; em->tD2M[1] = 17;
    em->tS2M[1] = 18; em->tM2E[1] = 22;
    return true;
}

static thread_local PinnedScratch tl_e2g_scratch;

struct E2gBatch {
    cudaStream_t stream = nullptr;
    int64_t *launches = nullptr;
    int n = 0, warps = 1, max_target = 0;
    bool want_path = false;
    E2gModel mdl;
    int64_t cells = 0;
    std::vector<int> order;           // pair indices, cost-descending
    std::vector<Chunk> chunks;
    DevBuf<uint8_t> d_seq;
    DevBuf<uint32_t> d_sp;
    DevBuf<uint8_t> d_lut;
    DevBuf<uint2> d_xtab;
    DevBuf<int> d_bad;
    DevBuf<E2gPair> d_pairs;
    bool packed = false;              // e2g_packed16.cuh: both strands per register, one warp per lattice
    DevBuf<E2pPair> d_pairs16;
    DevBuf<uint2> d_top;              // sweep hand-off rows (packed path, queries longer than 511)
    int max_query = 0;
    int rows16 = kE2pR;               // rows per lane of the packed kernels (8 for small batches), all layouts
    bool windowed = false;            // find_path by checkpoints + window refills (e2g_packed16.cuh)
    DevBuf<uint32_t> d_ck;
    DevBuf<E2pWalk> d_walk;
    DevBuf<int32_t> d_active, d_count;
    DevBuf<uint16_t> d_win;
    size_t win_stride = 0;
    DevBuf<E2gOut> d_outs;
    DevBuf<E2gJob> d_jobs;
    DevBuf<int32_t> d_qorg, d_torg;
    DevBuf<uint16_t> d_tb;
    DevBuf<c4b_result> d_results;
    DevBuf<int32_t> d_ops_slots, d_ops_packed;
    DevBuf<int64_t> d_new_off;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    ~E2gBatch() {
        d_seq.release(); d_sp.release(); d_lut.release(); d_xtab.release(); d_bad.release();
        d_pairs.release(); d_pairs16.release(); d_top.release(); d_ck.release(); d_walk.release(); d_active.release();
        d_count.release(); d_win.release(); d_outs.release(); d_jobs.release(); d_qorg.release(); d_torg.release();
        d_tb.release(); d_results.release(); d_ops_slots.release(); d_ops_packed.release();
        d_new_off.release();
        if (ev_a) cudaEventDestroy(ev_a);
        if (ev_b) cudaEventDestroy(ev_b);
    }
};

// returns 0 ok, 1 "not eligible, use the generic path", <0 error
static int e2g_batch_create(cudaStream_t stream, int64_t *launch_counter, const c4b_model *model,
                            const c4b_scoring *scoring, int n, const c4b_pair *pairs, bool want_path,
                            E2gBatch **out, bool allow_packed = true) {
    E2gModel mdl;
    if (!analyze_est2genome(*model, *scoring, &mdl)) return 1;
    int maxQ = 0;
    bool used[24] = {false};
    for (int p = 0; p < n; ++p) {
        const c4b_pair &pp = pairs[p];
        if (pp.n_blocked) return 1;
        if (pp.query_length < 0 || pp.target_length < 0 || pp.query_start < 0 || pp.target_start < 0 ||
            (int64_t)pp.query_start + pp.query_length > pp.query_len ||
            (int64_t)pp.target_start + pp.target_length > pp.target_len) {
            set_error("pair " + std::to_string(p) + ": region outside the sequences");
            return -1;
        }
        for (int k = 0; k < 4; ++k)
            if (!pp.splice[k]) {
                set_error("est2genome pair " + std::to_string(p) + " has no splice arrays");
                return -1;
            }
        maxQ = std::max(maxQ, pp.query_length);
        for (int k = 0; k < pp.query_length; ++k) {
            const int c = scoring->dna_index[pp.query[pp.query_start + k]];
            if (c >= 24) {
                set_error("query " + std::to_string(p) + ": symbol outside the substitution matrix");
                return -1;
            }
            used[c] = true;
        }
    }
    int maxT = 0;
    for (int p = 0; p < n; ++p) maxT = std::max(maxT, pairs[p].target_length);
    // PRMT classes of the query alphabet; s - open must fit int8
    int n_used = 0, cls_of[24], code_of[8];
    for (int a = 0; a < 24; ++a) {
        cls_of[a] = -1;
        if (!used[a]) continue;
        if (n_used >= 7) return 1;
        for (int c = 0; c < 24; ++c) {
            const int v = scoring->dna_matrix[a * 24 + c] - mdl.open;
            if (v < -127 || v > 127) return 1;
        }
        code_of[n_used] = a;
        cls_of[a] = n_used++;
    }
    // packed 16-bit path (e2g_packed16.cuh): exact when every reachable value fits a
    // signed halfword and the intron-length upper bound can never fire
    bool packed = false;
    {
        int max_sub = 0;
        for (int a = 0; a < 24; ++a)
            for (int c = 0; c < 24 && used[a]; ++c) max_sub = std::max(max_sub, scoring->dna_matrix[a * 24 + c]);
        const char *env = getenv("C4B_E2G_PACK16");
        const int64_t top_score = (int64_t)max_sub * (std::min(maxQ, maxT) + 1) + 400;
        packed = allow_packed && !(env && atoi(env) == 0) && top_score <= 30000 && mdl.open > -1000 && mdl.ext > -1000 &&
                 mdl.intron_open > -4000 && mdl.intron_open < 1000 && mdl.min_intron - 2 <= 32000 &&
                 (int64_t)mdl.max_intron >= (int64_t)maxT + 2;
    }
    if (!packed && maxQ + 1 > kE2gMaxWarps * 32 * kE2gR) return 1;

    E2gBatch *b = new E2gBatch();
    b->stream = stream; b->launches = launch_counter; b->n = n; b->want_path = want_path; b->mdl = mdl;
    b->packed = packed;
    b->max_target = maxT;
    b->max_query = maxQ;
    b->warps = std::max(1, (maxQ + 1 + 32 * kE2gR - 1) / (32 * kE2gR));
    // Rows per lane of the packed kernels.  A batch that cannot give every scheduler a warp at 16 rows
    // per lane (one warp per 512-row sweep: 125 lattices of a 1 kbp cDNA are 250 warps for 592
    // schedulers) runs 8 rows per lane instead: twice the sweeps, each on its own pipelined warp.
    {
        const int sweeps16 = (maxQ + 1 + 32 * kE2pR - 1) / (32 * kE2pR);
        const int sweeps8 = (maxQ + 1 + 32 * 8 - 1) / (32 * 8);
        // (measured, 1 kbp x 100 kbp find_path GCUPS, 16 rows one warp -> 8 rows four warps: 125 lattices
        // 107 -> 183+, 250 210 -> 315+, 500 413 -> 447+, 1000 559 -> 540: profiles/r02_e2g_small.md)
        b->rows16 = ((int64_t)n * sweeps16 <= 1400 && sweeps8 > sweeps16) ? 8 : kE2pR;
        // (4 rows on eight warps: 125 lattices find_score 280 -> 259 but find_path 196 -> 214 -- the window
        // refills like the extra warps; equal at 250 lattices)
        if (b->rows16 == 8 && want_path && (int64_t)n * sweeps16 <= 300 && (maxQ + 1 + 127) / 128 > sweeps8) b->rows16 = 4;
        if (const char *env = getenv("C4B_E2G_ROWS")) b->rows16 = (atoi(env) == 8) ? 8 : (atoi(env) == 4 ? 4 : kE2pR);
    }
    const int RW = b->rows16;
    // ---- staging: region slices of query / target, packed splice words ---------------
    std::vector<size_t> qoff(n), toff(n);
    size_t qbytes = 0, tbytes = 0;
    std::map<std::pair<const uint8_t *, int>, size_t> qmap, tmap;
    for (int p = 0; p < n; ++p) {
        const c4b_pair &pp = pairs[p];
        auto qk = std::make_pair(pp.query + pp.query_start, pp.query_length);
        auto tk = std::make_pair(pp.target + pp.target_start, pp.target_length);
        auto qi = qmap.find(qk);
        if (qi == qmap.end()) { qmap[qk] = qbytes; qoff[p] = qbytes; qbytes += align_up((size_t)pp.query_length, 16) + 16; }
        else qoff[p] = qi->second;
        auto ti = tmap.find(tk);
        if (ti == tmap.end()) { tmap[tk] = tbytes; toff[p] = tbytes; tbytes += align_up((size_t)pp.target_length, 16) + 16; }
        else toff[p] = ti->second;
        b->cells += (int64_t)pp.query_length * pp.target_length;
    }
    std::vector<uint8_t> lut(512, 0xFF);
    for (int c = 0; c < 256; ++c) {
        const int idx = scoring->dna_index[c];
        if (idx < 24 && cls_of[idx] >= 0) lut[c] = (uint8_t)cls_of[idx];
        if (idx < 24) lut[256 + c] = (uint8_t)idx;
    }
    uint8_t qfill = 0, tfill = 0;
    for (int c = 255; c >= 0; --c) {
        if (lut[c] != 0xFF) qfill = (uint8_t)c;
        if (lut[256 + c] != 0xFF) tfill = (uint8_t)c;
    }
    // staged through pinned memory, filled by worker threads: sequences, slot padding, and
    // the four int32 splice arrays of every distinct target packed to one int8 x 4 word per
    // position (a score outside int8 sends the batch to the table-driven kernel instead)
    const size_t seq_bytes = align_up(qbytes + tbytes + 64, 256);
    uint8_t *scratch = tl_e2g_scratch.get(seq_bytes + (tbytes + 16) * 4);
    if (!scratch) {
        set_error("pinned staging allocation failed");
        delete b;
        return -1;
    }
    uint8_t *hseq = scratch;
    uint32_t *hsp = reinterpret_cast<uint32_t *>(scratch + seq_bytes);
    {
        typedef std::pair<const uint8_t *, int> SeqKey;
        std::vector<std::pair<SeqKey, size_t>> qlist(qmap.begin(), qmap.end()), tlist(tmap.begin(), tmap.end());
        std::map<SeqKey, int> owner;  // target slice -> a pair that carries its splice arrays
        for (int p = 0; p < n; ++p) owner[SeqKey(pairs[p].target + pairs[p].target_start, pairs[p].target_length)] = p;
        // targets in slot order: a run of them is one contiguous range of the staging block, so the DMA
        // of a packed run goes out while the host threads pack the next one
        std::sort(tlist.begin(), tlist.end(), [](const std::pair<SeqKey, size_t> &a, const std::pair<SeqKey, size_t> &c) {
            return a.second < c.second;
        });
        std::vector<const c4b_pair *> towner(tlist.size());
        for (size_t k = 0; k < tlist.size(); ++k) towner[k] = &pairs[owner[tlist[k].first]];
        for (auto &kv : qlist) {
            const size_t len = (size_t)kv.first.second, slot = align_up(len, 16) + 16;
            memcpy(hseq + kv.second, kv.first.first, len);
            memset(hseq + kv.second + len, qfill, slot - len);
        }
        memset(hseq + qbytes + tbytes, tfill, 64);
        memset(hsp + tbytes, 0, 16 * sizeof(uint32_t));
        if (b->d_seq.alloc(qbytes + tbytes + 64) || b->d_sp.alloc(tbytes + 16)) {
            delete b;
            return -1;
        }
        bool copied = true;
        if (qbytes) copied = cudaMemcpyAsync(b->d_seq.p, hseq, qbytes, cudaMemcpyHostToDevice, stream) == cudaSuccess;
        copied = copied && cudaMemcpyAsync(b->d_seq.p + qbytes + tbytes, hseq + qbytes + tbytes, 64,
                                           cudaMemcpyHostToDevice, stream) == cudaSuccess;
        copied = copied && cudaMemcpyAsync(b->d_sp.p + tbytes, hsp + tbytes, 16 * sizeof(uint32_t),
                                           cudaMemcpyHostToDevice, stream) == cudaSuccess;
        std::atomic<bool> out_of_range(false);
        // the 16-bit value bound assumes an intron never GAINS score: intron_open + 5'ss + 3'ss <= 0
        // on either strand (true for the default --intronpenalty -30; with e.g. 0, introns chain
        // without consuming query and the score grows with their number: halfword adds would wrap)
        std::atomic<int> max_gain(INT32_MIN);
        const int n_t = (int)tlist.size(), run = std::max(1, (n_t + 7) / 8);
        for (int first = 0; first < n_t; first += run) {
        const int last = std::min(n_t, first + run);
        parallel_for(last - first, [&](int kk) {
            const int k = first + kk;
            const size_t off = tlist[k].second, len = (size_t)tlist[k].first.second, slot = align_up(len, 16) + 16;
            memcpy(hseq + qbytes + off, tlist[k].first.first, len);
            memset(hseq + qbytes + off + len, tfill, slot - len);
            const c4b_pair &pp = *towner[k];
            const int32_t *s0 = pp.splice[0] + pp.target_start, *s1 = pp.splice[1] + pp.target_start;
            const int32_t *s2 = pp.splice[2] + pp.target_start, *s3 = pp.splice[3] + pp.target_start;
            uint32_t *dst = hsp + off;
            bool bad = false;
            int32_t m0 = INT32_MIN, m1 = INT32_MIN, m2 = INT32_MIN, m3 = INT32_MIN;
            for (size_t j = 0; j < len; ++j) {
                const int32_t a = s0[j], c = s1[j], d = s2[j], e = s3[j];
                m0 = std::max(m0, a); m1 = std::max(m1, c); m2 = std::max(m2, d); m3 = std::max(m3, e);
                bad |= (uint32_t)(a + 127) > 254u || (uint32_t)(c + 127) > 254u || (uint32_t)(d + 127) > 254u ||
                       (uint32_t)(e + 127) > 254u;
                dst[j] = (uint32_t)(uint8_t)a | ((uint32_t)(uint8_t)c << 8) | ((uint32_t)(uint8_t)d << 16) |
                         ((uint32_t)(uint8_t)e << 24);
            }
            if (bad) out_of_range = true;
            if (len) {
                const int g = std::max(m0 + m1, m2 + m3);
                int cur = max_gain.load();
                while (g > cur && !max_gain.compare_exchange_weak(cur, g)) {}
            }
        });
        const size_t lo = tlist[first].second;
        const size_t hi = tlist[last - 1].second + align_up((size_t)tlist[last - 1].first.second, 16) + 16;
        copied = copied && cudaMemcpyAsync(b->d_seq.p + qbytes + lo, hseq + qbytes + lo, hi - lo, cudaMemcpyHostToDevice,
                                           stream) == cudaSuccess;
        copied = copied && cudaMemcpyAsync(b->d_sp.p + lo, hsp + lo, (hi - lo) * sizeof(uint32_t), cudaMemcpyHostToDevice,
                                           stream) == cudaSuccess;
        }
        tl_e2g_scratch.mark(stream);   // (a retry on another kernel below reuses the scratch)
        if (!copied) {
            set_error("staging the est2genome batch failed");
            delete b;
            return -1;
        }
        if (out_of_range) {
            delete b;
            return 1;
        }
        if (packed && max_gain.load() != INT32_MIN && (int64_t)mdl.intron_open + max_gain.load() > 0) {
            delete b;   // values are unbounded: the int32 kernel (or the table-driven one) takes the batch
            return e2g_batch_create(stream, launch_counter, model, scoring, n, pairs, want_path, out, false);
        }
    }
    std::vector<uint2> xt(25);
    for (int tc = 0; tc < 25; ++tc) {
        int8_t x[8] = {0};
        for (int k = 0; k < n_used; ++k) x[k] = (tc < 24) ? (int8_t)(scoring->dna_matrix[code_of[k] * 24 + tc] - mdl.open) : 0;
        x[kPadClass] = (int8_t)(-100 - mdl.open);
        memcpy(&xt[tc], x, 8);
    }
    // ---- traceback arena + chunks ------------------------------------------------------
    b->order.resize(n);
    for (int p = 0; p < n; ++p) b->order[p] = p;
    std::sort(b->order.begin(), b->order.end(), [&](int a, int c) {
        const int64_t ca = (int64_t)pairs[a].query_length * pairs[a].target_length;
        const int64_t cc = (int64_t)pairs[c].query_length * pairs[c].target_length;
        return ca != cc ? ca > cc : a < c;
    });
    size_t free_b = 0, total_b = 0;
    C4B_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const size_t fixed = qbytes + tbytes * 5 + (1ull << 30);
    const size_t budget_hw = (free_b > fixed + (1ull << 30) ? (free_b - fixed) / 2 : (256ull << 20)) / 2;
    std::vector<size_t> tb_off(n, 0);
    size_t arena = 0;
    int64_t ops_cursor = 0;
    std::vector<E2gJob> jobs(n);
    // windowed traceback: per lattice checkpoints + one window of records, all resident
    std::vector<size_t> ck_off(n, 0);
    size_t ck_words = 0;
    b->mdl.win_cols = kE2pWinMax;
    if (packed && want_path) {
        const char *env = getenv("C4B_E2G_WINDOWS");
        // window width: the narrowest whose checkpoints fit a quarter of the record budget (a refill
        // covers at most one window per cursor and round: narrow windows refill fewer cells per exon)
        int wc = 512;   // (measured at 1000 lattices of 1 kbp x 100 kbp: 256 / 512 / 1024 columns 585 / 595 / 573 GCUPS)
        if (const char *wenv = getenv("C4B_E2G_WINDOW_COLS")) {
            wc = std::max(64, std::min(kE2pWinMax, atoi(wenv)));
            while (wc & (wc - 1)) wc &= wc - 1;
        }
        for (;; wc <<= 1) {
            size_t max_sweeps = 1;
            ck_words = 0;
            for (int p = 0; p < n; ++p) {
                const size_t sweeps = ((size_t)pairs[p].query_length + 1 + 32 * RW - 1) / (32 * RW);
                const size_t nwin = (size_t)pairs[p].target_length / wc + 1;
                max_sweeps = std::max(max_sweeps, sweeps);
                ck_off[p] = ck_words;
                ck_words += (nwin - 1) * sweeps * 32 * RW * kE2pCkWords;
            }
            b->win_stride = max_sweeps * (size_t)(wc + 31) * 32 * RW;
            if (ck_words * 4 <= budget_hw / 4 || wc >= kE2pWinMax) break;
        }
        b->mdl.win_cols = wc;
        const size_t need = ck_words * 4 + (size_t)n * b->win_stride * 2;
        b->windowed = !(env && atoi(env) == 0) && need / 2 < budget_hw;
    }
    if (want_path) {
        size_t cur = 0;
        int begin = 0;
        for (int k = 0; k < n; ++k) {
            const c4b_pair &pp = pairs[b->order[k]];
            const size_t sweeps = ((size_t)pp.query_length + 1 + 32 * RW - 1) / (32 * RW);
            const size_t hw = b->windowed ? 0
                              : packed ? align_up(sweeps * (pp.target_length + 32) * 32 * RW, 16)
                                       : align_up((size_t)b->warps * (pp.target_length + 32) * 32 * kE2gR, 16);
            if (hw > budget_hw) {
                set_error("traceback of pair " + std::to_string(b->order[k]) + " exceeds the device memory budget");
                delete b;
                return -1;
            }
            if (cur + hw > budget_hw) { b->chunks.push_back({begin, k}); begin = k; cur = 0; }
            tb_off[k] = cur;
            cur += hw;
            arena = std::max(arena, cur);
            E2gJob &J = jobs[k];
            J.pair = k; J.result = b->order[k]; J.q_origin = pp.query_start; J.t_origin = pp.target_start;
            J.ops_cap = (int32_t)std::min<int64_t>((int64_t)pp.query_length + pp.target_length + 4, INT32_MAX);
            J.ops_off = ops_cursor; J.reserved = 0;
            ops_cursor += J.ops_cap;
        }
        if (begin < n) b->chunks.push_back({begin, n});
    } else {
        b->chunks.push_back({0, n});
    }
    int rc = 0;
    rc |= b->d_lut.alloc(512);
    rc |= b->d_xtab.alloc(25);
    rc |= b->d_bad.alloc(1);
    rc |= b->d_pairs.alloc(n);
    std::vector<size_t> top_off(n, (size_t)-1);
    size_t top_elems = 0;
    if (packed) {
        for (int p = 0; p < n; ++p)
            if (pairs[p].query_length + 1 > 32 * RW) {
                const size_t sweeps = ((size_t)pairs[p].query_length + 1 + 32 * RW - 1) / (32 * RW);
                top_off[p] = top_elems;
                top_elems += (sweeps - 1) * ((size_t)pairs[p].target_length + 1);
            }
        rc |= b->d_pairs16.alloc(n);
        rc |= b->d_top.alloc(top_elems);
        if (b->windowed) {
            rc |= b->d_ck.alloc(ck_words + 8);
            rc |= b->d_walk.alloc(n);
            rc |= b->d_active.alloc(n);
            rc |= b->d_count.alloc(1);
            rc |= b->d_win.alloc((size_t)n * b->win_stride + 16);
        }
    }
    rc |= b->d_outs.alloc(n);
    rc |= b->d_results.alloc(n);
    rc |= b->d_qorg.alloc(n);
    rc |= b->d_torg.alloc(n);
    if (want_path) {
        rc |= b->d_jobs.alloc(n);
        rc |= b->d_tb.alloc(arena + 16);
        rc |= b->d_ops_slots.alloc(2 * (size_t)ops_cursor + 2);
        rc |= b->d_ops_packed.alloc(2 * (size_t)ops_cursor + 2);
        rc |= b->d_new_off.alloc((size_t)n + 1);
    }
    if (rc) { delete b; return -1; }
    std::vector<E2gPair> hp(n);
    std::vector<int32_t> hq(n), ht(n);
    for (int k = 0; k < n; ++k) {
        const int p = b->order[k];
        E2gPair &e = hp[k];
        e.q = b->d_seq.p + qoff[p];
        e.t = b->d_seq.p + qbytes + toff[p];
        e.sp = b->d_sp.p + toff[p];
        e.Q = pairs[p].query_length;
        e.T = pairs[p].target_length;
        e.tb = want_path ? b->d_tb.p + tb_off[k] : nullptr;
        e.out_index = k;
        hq[k] = pairs[p].query_start;
        ht[k] = pairs[p].target_start;
    }
    bool ok = true;
    std::vector<E2pPair> hp16(packed ? n : 0);
    if (packed) {
        for (int k = 0; k < n; ++k) {
            const int p = b->order[k];
            E2pPair &e = hp16[k];
            e.q = hp[k].q; e.t = hp[k].t; e.sp = hp[k].sp; e.Q = hp[k].Q; e.T = hp[k].T;
            e.tb = hp[k].tb;
            e.top = (top_off[p] != (size_t)-1) ? b->d_top.p + top_off[p] : nullptr;
            e.ck = b->windowed ? b->d_ck.p + ck_off[p] : nullptr;
            e.out_index = k;
        }
        ok &= cudaMemcpyAsync(b->d_pairs16.p, hp16.data(), n * sizeof(E2pPair), cudaMemcpyHostToDevice, stream) == cudaSuccess;
    }
    tl_e2g_scratch.mark(stream);   // (sequences and splice words went out run by run while they were packed)
    ok &= cudaMemcpyAsync(b->d_lut.p, lut.data(), 512, cudaMemcpyHostToDevice, stream) == cudaSuccess;
    ok &= cudaMemcpyAsync(b->d_xtab.p, xt.data(), 25 * sizeof(uint2), cudaMemcpyHostToDevice, stream) == cudaSuccess;
    ok &= cudaMemsetAsync(b->d_bad.p, 0, sizeof(int), stream) == cudaSuccess;
    ok &= cudaMemcpyAsync(b->d_pairs.p, hp.data(), n * sizeof(E2gPair), cudaMemcpyHostToDevice, stream) == cudaSuccess;
    ok &= cudaMemcpyAsync(b->d_qorg.p, hq.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, stream) == cudaSuccess;
    ok &= cudaMemcpyAsync(b->d_torg.p, ht.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, stream) == cudaSuccess;
    if (want_path)
        ok &= cudaMemcpyAsync(b->d_jobs.p, jobs.data(), n * sizeof(E2gJob), cudaMemcpyHostToDevice, stream) == cudaSuccess;
    if (qbytes) encode_kernel<<<(unsigned)((qbytes / 16 + 255) / 256 + 1), 256, 0, stream>>>(b->d_seq.p, qbytes, b->d_lut.p, b->d_bad.p);
    if (tbytes) encode_kernel<<<(unsigned)((tbytes / 16 + 255) / 256 + 1), 256, 0, stream>>>(b->d_seq.p + qbytes, tbytes, b->d_lut.p + 256, b->d_bad.p);
    (*launch_counter) += 2;
    ok &= cudaStreamSynchronize(stream) == cudaSuccess;
    int bad = 0;
    ok &= cudaMemcpy(&bad, b->d_bad.p, sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess;
    if (!ok || bad) {
        set_error(bad ? "a sequence holds a symbol outside the substitution matrix alphabet"
                      : "staging the est2genome batch failed");
        delete b;
        return -1;
    }
    cudaEventCreate(&b->ev_a);
    cudaEventCreate(&b->ev_b);
    *out = b;
    return 0;
}

// the packed fill kernel for a batch's rows per lane / warps per lattice (the whole-lattice record pass
// E2P_FULL_TB is always one warp per lattice: its pipelined forms are not instantiated)
template <int MODE, int ROWS>
static void e2p_launch_rows(int warps, int grid, cudaStream_t st, const E2pPair *pairs, E2gOut *outs,
                            const E2gModel &mdl, const uint2 *xtab, const int32_t *active, const E2pWalk *walk,
                            uint16_t *winbuf, size_t win_stride) {
    if constexpr (MODE != E2P_FULL_TB) {
        if (warps > 1) {
            e2g_fill16_kernel<MODE, true, ROWS><<<grid, 32 * warps, 0, st>>>(pairs, outs, mdl, xtab, active, walk, winbuf,
                                                                            win_stride);
            return;
        }
    }
    e2g_fill16_kernel<MODE, false, ROWS><<<grid, 32, 0, st>>>(pairs, outs, mdl, xtab, active, walk, winbuf, win_stride);
}

template <int MODE>
static void e2p_launch(int rows, int warps, int grid, cudaStream_t st, const E2pPair *pairs, E2gOut *outs,
                       const E2gModel &mdl, const uint2 *xtab, const int32_t *active, const E2pWalk *walk,
                       uint16_t *winbuf, size_t win_stride) {
    if (rows == 4) e2p_launch_rows<MODE, 4>(warps, grid, st, pairs, outs, mdl, xtab, active, walk, winbuf, win_stride);
    else if (rows == 8) e2p_launch_rows<MODE, 8>(warps, grid, st, pairs, outs, mdl, xtab, active, walk, winbuf, win_stride);
    else e2p_launch_rows<MODE, kE2pR>(warps, grid, st, pairs, outs, mdl, xtab, active, walk, winbuf, win_stride);
}

static int e2g_batch_run(E2gBatch *b, c4b_score threshold) {
    cudaStream_t st = b->stream;
    const int n = b->n;
    const int threads = 32 * b->warps;
    // packed kernel: small batches run the sweeps of a lattice on pipelined warps (the one-warp kernel
    // otherwise: see e2g_packed16.cuh for the measurements; C4B_E2G_WARPS / C4B_E2G_ROWS override)
    const int RW = b->rows16;
    const int sweeps16 = (b->max_query + 1 + 32 * RW - 1) / (32 * RW);
    const int max_warps = RW == 4 ? 8 : 4;   // (the kernels' launch bounds)
    int warps16 = (RW <= 8) ? std::max(1, std::min(max_warps, sweeps16)) : 1;
    if (const char *env = getenv("C4B_E2G_WARPS")) warps16 = std::max(1, std::min(std::min(max_warps, sweeps16), atoi(env)));
    // window refills: the sweeps of a refill are independent (hand-off rows kept from pass 1)
    const int warps_win = (RW <= 8) ? warps16 : 1;
    C4B_CUDA(cudaEventRecord(b->ev_a, st));
    if (b->packed) {
        if (!b->want_path) {
            e2p_launch<E2P_SCORE>(RW, warps16, n, st, b->d_pairs16.p, b->d_outs.p, b->mdl, b->d_xtab.p, nullptr, nullptr,
                                  nullptr, 0);
            C4B_CUDA(cudaGetLastError());
            C4B_CUDA(cudaEventRecord(b->ev_b, st));
            e2g16_score_results_kernel<<<(n + 127) / 128, 128, 0, st>>>(b->d_pairs16.p, b->d_outs.p, b->d_qorg.p,
                                                                       b->d_torg.p, n, b->d_results.p);
            (*b->launches) += 2;
            C4B_CUDA(cudaGetLastError());
            return 0;
        }
        if (b->windowed) {
            // pass 1: END cell + column checkpoints; then rounds of (refill the window under
            // each traceback cursor, walk it) until every cursor has reached START
            e2p_launch<E2P_SCORE_CK>(RW, warps16, n, st, b->d_pairs16.p, b->d_outs.p, b->mdl, b->d_xtab.p, nullptr,
                                     nullptr, nullptr, 0);
            e2g16_walk_init_kernel<<<(n + 127) / 128, 128, 0, st>>>(b->d_pairs16.p, b->d_outs.p, b->d_jobs.p, n, b->mdl,
                                                                   threshold, b->d_walk.p, b->d_ops_slots.p);
            e2g16_walk_rejected_kernel<<<(n + 127) / 128, 128, 0, st>>>(b->d_pairs16.p, b->d_outs.p, b->d_jobs.p, n,
                                                                       b->d_walk.p, b->d_results.p);
            (*b->launches) += 3;
            C4B_CUDA(cudaGetLastError());
            const int round_cap = 4 * (b->max_target / b->mdl.win_cols + 1) + 16;
            for (int round = 0;; ++round) {
                int cnt = 0;
                C4B_CUDA(cudaMemsetAsync(b->d_count.p, 0, sizeof(int32_t), st));
                e2g16_active_kernel<<<(n + 127) / 128, 128, 0, st>>>(b->d_walk.p, n, b->d_active.p, b->d_count.p);
                C4B_CUDA(cudaMemcpyAsync(&cnt, b->d_count.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
                C4B_CUDA(cudaStreamSynchronize(st));
                (*b->launches) += 1;
                if (cnt == 0) break;
                if (round >= round_cap) {
                    set_error("internal: est2genome windowed traceback did not terminate");
                    return -1;
                }
                e2p_launch<E2P_WINDOW_TB>(RW, warps_win, cnt, st, b->d_pairs16.p, b->d_outs.p, b->mdl, b->d_xtab.p,
                                          b->d_active.p, b->d_walk.p, b->d_win.p, b->win_stride);
                e2g16_walk_kernel<<<(cnt + 63) / 64, 64, 0, st>>>(b->d_pairs16.p, b->d_outs.p, b->d_jobs.p, b->d_active.p,
                                                                 cnt, b->mdl, b->d_walk.p, b->d_win.p, b->win_stride,
                                                                 b->d_results.p, b->d_ops_slots.p, RW);
                (*b->launches) += 2;
                C4B_CUDA(cudaGetLastError());
            }
            C4B_CUDA(cudaEventRecord(b->ev_b, st));
            apply_threshold_kernel<<<(n + 127) / 128, 128, 0, st>>>(b->d_results.p, n, threshold);
            ops_scan_kernel<<<1, 1024, 0, st>>>(b->d_results.p, n, b->d_new_off.p, b->d_new_off.p + n);
            ops_compact_kernel<<<n, 64, 0, st>>>(b->d_results.p, n, b->d_new_off.p, b->d_ops_slots.p, b->d_ops_packed.p);
            (*b->launches) += 3;
            C4B_CUDA(cudaGetLastError());
            return 0;
        }
        for (const Chunk &c : b->chunks) {
            const int cnt = c.end - c.begin;
            e2p_launch<E2P_FULL_TB>(RW, 1, cnt, st, b->d_pairs16.p + c.begin, b->d_outs.p, b->mdl, b->d_xtab.p, nullptr,
                                    nullptr, nullptr, 0);
            C4B_CUDA(cudaGetLastError());
            e2g16_traceback_kernel<<<(cnt + 63) / 64, 64, 0, st>>>(b->d_pairs16.p, b->d_outs.p, b->d_jobs.p + c.begin,
                                                                  cnt, b->mdl, threshold, b->d_results.p,
                                                                  b->d_ops_slots.p, RW);
            C4B_CUDA(cudaGetLastError());
            (*b->launches) += 2;
        }
        C4B_CUDA(cudaEventRecord(b->ev_b, st));
        apply_threshold_kernel<<<(n + 127) / 128, 128, 0, st>>>(b->d_results.p, n, threshold);
        ops_scan_kernel<<<1, 1024, 0, st>>>(b->d_results.p, n, b->d_new_off.p, b->d_new_off.p + n);
        ops_compact_kernel<<<n, 64, 0, st>>>(b->d_results.p, n, b->d_new_off.p, b->d_ops_slots.p, b->d_ops_packed.p);
        (*b->launches) += 3;
        C4B_CUDA(cudaGetLastError());
        return 0;
    }
    if (!b->want_path) {
        e2g_fill_kernel<false><<<n, threads, 0, st>>>(b->d_pairs.p, b->d_outs.p, b->mdl, b->d_xtab.p);
        C4B_CUDA(cudaGetLastError());
        C4B_CUDA(cudaEventRecord(b->ev_b, st));
        e2g_score_results_kernel<<<(n + 127) / 128, 128, 0, st>>>(b->d_pairs.p, b->d_outs.p, b->d_qorg.p,
                                                                 b->d_torg.p, n, b->d_results.p);
        (*b->launches) += 2;
        C4B_CUDA(cudaGetLastError());
        return 0;
    }
    for (const Chunk &c : b->chunks) {
        const int cnt = c.end - c.begin;
        e2g_fill_kernel<true><<<cnt, threads, 0, st>>>(b->d_pairs.p + c.begin, b->d_outs.p, b->mdl, b->d_xtab.p);
        C4B_CUDA(cudaGetLastError());
        e2g_traceback_kernel<<<(cnt + 63) / 64, 64, 0, st>>>(b->d_pairs.p, b->d_outs.p, b->d_jobs.p + c.begin, cnt,
                                                            b->mdl, threshold, b->d_results.p, b->d_ops_slots.p);
        C4B_CUDA(cudaGetLastError());
        (*b->launches) += 2;
    }
    C4B_CUDA(cudaEventRecord(b->ev_b, st));
    apply_threshold_kernel<<<(n + 127) / 128, 128, 0, st>>>(b->d_results.p, n, threshold);
    ops_scan_kernel<<<1, 1024, 0, st>>>(b->d_results.p, n, b->d_new_off.p, b->d_new_off.p + n);
    ops_compact_kernel<<<n, 64, 0, st>>>(b->d_results.p, n, b->d_new_off.p, b->d_ops_slots.p, b->d_ops_packed.p);
    (*b->launches) += 3;
    C4B_CUDA(cudaGetLastError());
    return 0;
}

static int e2g_batch_fetch(E2gBatch *b, c4b_result *results, int32_t *ops, int64_t ops_capacity) {
    cudaStream_t st = b->stream;
    const int n = b->n;
    if (!b->want_path) {
        std::vector<c4b_result> tmp(n);
        C4B_CUDA(cudaMemcpyAsync(tmp.data(), b->d_results.p, n * sizeof(c4b_result), cudaMemcpyDeviceToHost, st));
        C4B_CUDA(cudaStreamSynchronize(st));
        for (int k = 0; k < n; ++k) results[b->order[k]] = tmp[k];
        return 0;
    }
    int64_t total = 0;
    C4B_CUDA(cudaMemcpyAsync(results, b->d_results.p, n * sizeof(c4b_result), cudaMemcpyDeviceToHost, st));
    C4B_CUDA(cudaMemcpyAsync(&total, b->d_new_off.p + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    C4B_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < n; ++k)
        if (results[k].status >= 2) {
            set_error("internal: est2genome traceback of pair " + std::to_string(k) + " failed with status " +
                      std::to_string(results[k].status));
            return -1;
        }
    if (total > ops_capacity) {
        set_error("ops buffer too small: need capacity " + std::to_string(total));
        return -3;
    }
    if (total) C4B_CUDA(cudaMemcpy(ops, b->d_ops_packed.p, 2 * (size_t)total * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return 0;
}

static double e2g_batch_fill_ms(E2gBatch *b) {
    float f = 0;
    cudaStreamSynchronize(b->stream);
    if (cudaEventElapsedTime(&f, b->ev_a, b->ev_b) != cudaSuccess) return -1;
    return f;
}

}  // namespace c4b
