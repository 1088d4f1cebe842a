// generic_jit.inl -- run-time specialisation of the table-driven fill (host side).
//
// The reference compiles each model to C at build time (codegen + bootstrapper,
// src/c4/codegen.c, src/c4/viterbi.c:1638-1727).  Here the closed model is
// written out as constexpr tables in front of generic_jit_kernel.cuh and
// compiled for sm_100a with NVRTC the first time a (model, mode, CTA size) is
// needed; the cubin is kept for the life of the process (and on disk when
// C4B_JIT_CACHE_DIR is set).  libnvrtc is dlopen'ed: without it, or when the
// compile fails, the caller keeps the interpreter kernel (generic_wavefront.cuh)
// -- both are device kernels, there is no CPU path.
//
// C4B_GENERIC_JIT=0 never, =1 always, unset: batches of >= 2^31 lattice cells
// (below that the ~1 s compile per fill mode costs more than it saves in a
// one-shot process; BSDP's many small region fills stay on the interpreter) --
// unless C4B_JIT_CACHE_DIR is set: then the compile is paid once per model for
// good and every batch specialises (BSDP region fills: 0.5 ms instead of 1.5 ms).
#include <dlfcn.h>
#include <nvrtc.h>

#include <mutex>
#include <sstream>

#include "jit_embedded.inc"

namespace c4b {

// dynamic shared memory a specialised kernel may use for its lattice ring: 227 KB per
// CTA less its static tables (scoring 9.5 KB, 20 B per thread of reduction scratch)
constexpr int kJitSmemRingBytes = 232448 - 1024 - 9600 - 20 * 512;

struct JitKernel {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kern = nullptr;
    int threads = 0;
    int blocks_per_sm = 1;
};

struct NvrtcApi {
    void *handle = nullptr;
    decltype(&nvrtcCreateProgram) create = nullptr;
    decltype(&nvrtcCompileProgram) compile = nullptr;
    decltype(&nvrtcGetProgramLogSize) log_size = nullptr;
    decltype(&nvrtcGetProgramLog) log = nullptr;
    decltype(&nvrtcGetCUBINSize) cubin_size = nullptr;
    decltype(&nvrtcGetCUBIN) cubin = nullptr;
    decltype(&nvrtcDestroyProgram) destroy = nullptr;
    bool ok = false;
};

static NvrtcApi *nvrtc_api() {
    static NvrtcApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                               "/usr/local/cuda/lib64/libnvrtc.so"};
        for (const char *nm : names)
            if ((api.handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL))) break;
        if (!api.handle) return;
        api.create = (decltype(api.create))dlsym(api.handle, "nvrtcCreateProgram");
        api.compile = (decltype(api.compile))dlsym(api.handle, "nvrtcCompileProgram");
        api.log_size = (decltype(api.log_size))dlsym(api.handle, "nvrtcGetProgramLogSize");
        api.log = (decltype(api.log))dlsym(api.handle, "nvrtcGetProgramLog");
        api.cubin_size = (decltype(api.cubin_size))dlsym(api.handle, "nvrtcGetCUBINSize");
        api.cubin = (decltype(api.cubin))dlsym(api.handle, "nvrtcGetCUBIN");
        api.destroy = (decltype(api.destroy))dlsym(api.handle, "nvrtcDestroyProgram");
        api.ok = api.create && api.compile && api.log_size && api.log && api.cubin_size && api.cubin && api.destroy;
    });
    return api.ok ? &api : nullptr;
}

// The closed model as compile-time tables (what generic_jit_kernel.cuh expects).
// Columns of state s a reader can still ask for: the largest advance_query +
// advance_target of the advancing transitions that leave s, plus one (0: never read
// from the ring).  generic_jit_kernel.cuh lays the lattice ring out by these.
static std::vector<int> jit_state_depths(const c4b_model &m) {
    std::vector<int> depth(m.n_states, 0);
    for (int k = 0; k < m.n_transitions; ++k) {
        const c4b_transition &t = m.transitions[k];
        const int adv = t.advance_query + t.advance_target;
        if (t.input != m.start_state && adv > 0) depth[t.input] = std::max(depth[t.input], adv + 1);
    }
    return depth;
}

// ring words one lattice row needs (as the kernel lays it out)
static int jit_ring_words_per_row(const c4b_model &m, int mode, bool pack_start) {
    int rows = 0;
    for (int d : jit_state_depths(m)) rows += d;
    int C = 1 + m.n_shadow_slots;
    if (mode == GEN_REGION && m.start_scope != C4B_SCOPE_CORNER) {
        const int extra = (m.start_scope != C4B_SCOPE_QUERY) + (m.start_scope != C4B_SCOPE_TARGET);
        C += (extra == 2 && pack_start) ? 1 : extra;
    }
    return std::max(1, rows) * C;
}

static std::string jit_program_source(const c4b_model &m, int mode, int threads, bool smem_ring,
                                      bool pack_start) {
    std::ostringstream o;
    o << "typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t;\n"
         "typedef unsigned short uint16_t; typedef int int32_t; typedef unsigned int uint32_t;\n"
         "typedef long long int64_t; typedef unsigned long long uint64_t;\n"
         "#define INT_MIN (-2147483647 - 1)\n"
         "#define C4B200_TYPES_ONLY 1\n#include \"c4b200.h\"\n#include \"generic_types.h\"\n";
    int min_ctas = 1;  // resident CTAs per SM the register allocation must allow
    if (const char *env = getenv("C4B_JIT_MIN_CTAS")) min_ctas = std::max(1, std::min(16, atoi(env)));
    o << "#define JIT_MODE " << mode << "\n#define JIT_THREADS " << threads << "\n#define JIT_MIN_CTAS " << min_ctas
      << "\n#define JIT_SMEM_RING " << (smem_ring ? 1 : 0) << "\n#define JIT_PACK_START " << (pack_start ? 1 : 0)
      << "\n";
    o << "namespace c4bjit {\n";
    const int S = m.n_states, TN = m.n_transitions, NC = std::max(1, m.n_calcs);
    const std::vector<int> depth = jit_state_depths(m);
    o << "constexpr int S = " << S << ", TN = " << TN << ", NSH = " << m.n_shadow_slots << ", START = "
      << m.start_state << ", END = " << m.end_state << ", START_SCOPE = " << m.start_scope
      << ", END_SCOPE = " << m.end_scope << ";\n";
    auto tr_array = [&](const char *name, auto field) {
        o << "constexpr int " << name << "[TN] = {";
        for (int k = 0; k < TN; ++k) o << (k ? "," : "") << field(m.transitions[k]);
        o << "};\n";
    };
    tr_array("kTrIn", [](const c4b_transition &t) { return t.input; });
    tr_array("kTrOut", [](const c4b_transition &t) { return t.output; });
    tr_array("kTrAq", [](const c4b_transition &t) { return t.advance_query; });
    tr_array("kTrAt", [](const c4b_transition &t) { return t.advance_target; });
    tr_array("kTrCalc", [](const c4b_transition &t) { return t.calc; });
    tr_array("kTrLabel", [](const c4b_transition &t) { return t.label; });
    auto calc_array = [&](const char *name, auto field) {
        o << "constexpr int " << name << "[" << NC << "] = {";
        for (int k = 0; k < NC; ++k) o << (k ? "," : "") << (k < m.n_calcs ? field(m.calcs[k]) : 0);
        o << "};\n";
    };
    calc_array("kCalcKind", [](const c4b_calc &c) { return c.kind; });
    calc_array("kCalcProt", [](const c4b_calc &c) { return c.protect; });
    calc_array("kCalcP0", [](const c4b_calc &c) { return c.param[0]; });
    calc_array("kCalcP1", [](const c4b_calc &c) { return c.param[1]; });
    calc_array("kCalcP2", [](const c4b_calc &c) { return c.param[2]; });
    o << "constexpr int kShadow[S * C4B_MAX_SHADOW_SLOTS] = {";
    for (int s = 0; s < S; ++s)
        for (int l = 0; l < C4B_MAX_SHADOW_SLOTS; ++l)
            o << ((s || l) ? "," : "") << (l < m.n_shadow_slots ? (int)m.shadow_start[s][l] : 0);
    o << "};\nconstexpr int kDepth[S] = {";
    for (int s = 0; s < S; ++s) o << (s ? "," : "") << depth[s];
    o << "};\nconstexpr int kRingOff[S] = {";
    for (int s = 0, off = 0; s < S; ++s) { o << (s ? "," : "") << off; off += depth[s]; }
    o << "};\n}  // namespace c4bjit\n";
    o << kJitSrc_generic_jit_kernel_cuh;
    return o.str();
}

// ---- the systolic specialisation (generic_jit_systolic.cuh) -----------------------------
// Register layout of one lattice row, derived from the closed model.
struct SysLayout {
    bool ok = false;
    int R = 1;            // lattice rows per lane
    int AQ = 0;           // largest advance_query of a transition that reads the lattice
    int VW = 0;           // words per row
    int chunk = 16;       // PATH bytes per lane per step
    std::vector<int> NW, VD, VOff;          // per state
    std::vector<int> need;                  // [S * C4B_MAX_SHADOW_SLOTS]
    std::vector<int> sendD, sendOff;        // hand-off words
    // PATH record: per cell, per state, WHICH of the transitions that enter the state won, as its
    // 1-based rank among them in closed-model order (0 = state unset): ceil(log2(n_in + 1)) bits
    // instead of the reference's pointer (viterbi.c:220-227) or a byte per state
    std::vector<int> tbCode, tbBits, tbBitOff;   // per transition / per state / per state
    int rowBits = 0;
};

static void jit_tb_bits(const c4b_model &m, std::vector<int> *code, std::vector<int> *bits, std::vector<int> *off,
                        int *row_bits) {
    const int S = m.n_states, TN = m.n_transitions;
    std::vector<int> n_in(S, 0);
    code->assign(TN, 0);
    for (int k = 0; k < TN; ++k) (*code)[k] = ++n_in[m.transitions[k].output];
    bits->assign(S, 0);
    off->assign(S, 0);
    *row_bits = 0;
    for (int s = 0; s < S; ++s) {
        int b = 0;
        while ((1 << b) < n_in[s] + 1) ++b;
        (*bits)[s] = b;
        (*off)[s] = *row_bits;
        *row_bits += b;
    }
}

static SysLayout jit_sys_layout(const c4b_model &m, int mode, bool pack_start) {
    SysLayout L;
    const int S = m.n_states, TN = m.n_transitions, NSH = m.n_shadow_slots;
    const bool region = mode == GEN_REGION && m.start_scope != C4B_SCOPE_CORNER;
    // REGION carries the START cell as ONE packed word: needs both coordinates free and < 2^31 cells
    if (region && !(pack_start && m.start_scope != C4B_SCOPE_QUERY && m.start_scope != C4B_SCOPE_TARGET)) return L;
    L.need.assign((size_t)S * C4B_MAX_SHADOW_SLOTS, 0);
    for (int l = 0; l < NSH; ++l) {
        // F: states a stamped value can reach; B: states whose slot value can still be read.
        // Every transition LEAVING a stamping state overwrites the slot (viterbi.c:413-422).
        std::vector<char> F(S, 0), B(S, 0);
        for (bool grow = true; grow;) {
            grow = false;
            for (int k = 0; k < TN; ++k) {
                const c4b_transition &t = m.transitions[k];
                if ((m.shadow_start[t.input][l] || F[t.input]) && !F[t.output]) { F[t.output] = 1; grow = true; }
            }
        }
        for (int k = 0; k < TN; ++k) {
            const c4b_transition &t = m.transitions[k];
            if (t.calc >= 0 && m.calcs[t.calc].kind >= C4B_CALC_SPLICE_POST && m.calcs[t.calc].param[2] == l)
                B[t.input] = 1;
        }
        for (bool grow = true; grow;) {
            grow = false;
            for (int k = 0; k < TN; ++k) {
                const c4b_transition &t = m.transitions[k];
                if (B[t.output] && !m.shadow_start[t.input][l] && !B[t.input]) { B[t.input] = 1; grow = true; }
            }
        }
        for (int s = 0; s < S; ++s) L.need[(size_t)s * C4B_MAX_SHADOW_SLOTS + l] = F[s] && B[s];
    }
    L.NW.assign(S, 1);
    L.VD.assign(S, -1);
    L.VOff.assign(S, 0);
    for (int s = 0; s < S; ++s) {
        for (int l = 0; l < NSH; ++l) L.NW[s] += L.need[(size_t)s * C4B_MAX_SHADOW_SLOTS + l];
        if (region) L.NW[s] += 1;
    }
    for (int k = 0; k < TN; ++k) {
        const c4b_transition &t = m.transitions[k];
        if (t.input == m.start_state || t.advance_query + t.advance_target == 0) continue;
        L.VD[t.input] = std::max(L.VD[t.input], t.advance_target);
        L.AQ = std::max(L.AQ, t.advance_query);
    }
    for (int s = 0; s < S; ++s) {
        L.VOff[s] = L.VW;
        if (L.VD[s] >= 0) L.VW += L.NW[s] * (L.VD[s] + 1);
    }
    if (L.AQ < 1 || L.VW < 1) return L;   // nothing advances the query: not a lattice this mapping helps
    for (int d = 1; d <= L.AQ; ++d)
        for (int s = 0; s < S; ++s) {
            bool sent = false;
            for (int k = 0; k < TN; ++k)
                sent = sent || (m.transitions[k].input == s && s != m.start_state && m.transitions[k].advance_query >= d);
            if (!sent || L.VD[s] < 0) continue;
            for (int w = 0; w < L.NW[s]; ++w) { L.sendD.push_back(d); L.sendOff.push_back(L.VOff[s] + w); }
        }
    // rows per lane from the register budget: (AQ + R) rows of VW words next to ~70 registers of
    // working set; more rows amortise the per-step hand-off, fewer keep more warps resident
    int budget = 150;
    if (const char *env = getenv("C4B_JIT_SYS_REGS")) budget = std::max(32, atoi(env));
    L.R = std::max(1, std::min(8, budget / L.VW - L.AQ));
    if (const char *env = getenv("C4B_JIT_SYS_R")) L.R = std::max(1, std::min(16, atoi(env)));
    jit_tb_bits(m, &L.tbCode, &L.tbBits, &L.tbBitOff, &L.rowBits);
    L.chunk = std::max(4, (L.R * L.rowBits + 31) / 32 * 4);
    L.ok = (L.AQ + L.R) * L.VW <= 230 && (int)L.sendD.size() <= 64;
    return L;
}

// warps per CTA (= strips of one lattice in flight) for queries of up to max_q symbols
static int jit_sys_warps(const SysLayout &L, int max_q) {
    int max_warps = 8;
    if (const char *env = getenv("C4B_JIT_SYS_WARPS")) max_warps = std::max(1, std::min(8, atoi(env)));
    return std::max(1, std::min(max_warps, (max_q + 32 * L.R) / (32 * L.R)));
}

// win: 0 whole-lattice pass, 1 score pass that leaves column checkpoints, 2 PATH over one column
// window started from a checkpoint (JIT_SYS_WIN in generic_jit_systolic.cuh)
// blk: lattices carry SubOpt blocked cells as {column, row mask} entries (JIT_SYS_BLK)
static std::string jit_sys_program_source(const c4b_model &m, int mode, bool pack_start, const SysLayout &L,
                                          int warps, int win = 0, bool blk = false) {
    // the tables of the thread-per-row kernel first (its calc / scope code is shared), then ours
    std::string base = jit_program_source(m, mode, 128, false, pack_start);
    const size_t cut = base.find(kJitSrc_generic_jit_kernel_cuh);
    std::ostringstream o;
    o << base.substr(0, cut);
    // resident CTAs per SM the register allocation must allow: 16 warps per SM (128 registers)
    int minb = std::max(1, 16 / warps);
    if (const char *env = getenv("C4B_JIT_SYS_MINB")) minb = std::max(1, std::min(16, atoi(env)));
    o << "#define JIT_SYSTOLIC 1\n#define JIT_SYS_R " << L.R << "\n#define JIT_SYS_MINB " << minb
      << "\n#define JIT_SYS_WARPS " << warps << "\n";
    if (win) o << "#define JIT_SYS_WIN " << win << "\n";
    if (blk) o << "#define JIT_SYS_BLK 1\n";
    o << "namespace c4bjit {\n";
    o << "constexpr int AQ = " << L.AQ << ", VW = " << L.VW << ", NSEND = " << L.sendD.size() << ";\n";
    auto arr = [&](const char *name, const std::vector<int> &v, size_t n) {
        o << "constexpr int " << name << "[" << std::max<size_t>(1, n) << "] = {";
        for (size_t k = 0; k < std::max<size_t>(1, n); ++k) o << (k ? "," : "") << (k < v.size() ? v[k] : 0);
        o << "};\n";
    };
    arr("kNW", L.NW, L.NW.size());
    arr("kVD", L.VD, L.VD.size());
    arr("kVOff", L.VOff, L.VOff.size());
    arr("kNeed", L.need, L.need.size());
    arr("kSendD", L.sendD, L.sendD.size());
    arr("kSendOff", L.sendOff, L.sendOff.size());
    arr("kTbCode", L.tbCode, L.tbCode.size());
    arr("kTbBits", L.tbBits, L.tbBits.size());
    arr("kTbBitOff", L.tbBitOff, L.tbBitOff.size());
    o << "constexpr int TB_ROW_BITS = " << L.rowBits << ", TB_CHUNK = " << L.chunk << ";\n";
    o << "}  // namespace c4bjit\n";
    o << kJitSrc_generic_jit_kernel_cuh << "\n" << kJitSrc_generic_jit_systolic_cuh;
    return o.str();
}

static uint64_t fnv1a(const std::string &s) {
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
    return h;
}

static bool jit_compile(const std::string &src, std::vector<char> *cubin, std::string *log) {
    NvrtcApi *rt = nvrtc_api();
    if (!rt) { *log = "libnvrtc not found"; return false; }
    const char *headers[] = {kJitSrc_c4b200_h, kJitSrc_generic_types_h, "", ""};
    const char *names[] = {"c4b200.h", "generic_types.h", "stdint.h", "stddef.h"};
    nvrtcProgram prog = nullptr;
    if (rt->create(&prog, src.c_str(), "c4b_jit_fill.cu", 4, headers, names) != NVRTC_SUCCESS) {
        *log = "nvrtcCreateProgram failed";
        return false;
    }
    const bool verbose = getenv("C4B_JIT_VERBOSE") != nullptr;  // ptxas resource usage on stderr
    const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "--ptxas-options=-v"};
    const nvrtcResult res = rt->compile(prog, verbose ? 4 : 3, opts);
    size_t ls = 0;
    rt->log_size(prog, &ls);
    if (ls > 1) { log->resize(ls); rt->log(prog, &(*log)[0]); }
    if (verbose && ls > 1) fprintf(stderr, "%s\n", log->c_str());
    bool ok = res == NVRTC_SUCCESS;
    if (ok) {
        size_t cs = 0;
        ok = rt->cubin_size(prog, &cs) == NVRTC_SUCCESS && cs > 0;
        if (ok) { cubin->resize(cs); ok = rt->cubin(prog, cubin->data()) == NVRTC_SUCCESS; }
    }
    rt->destroy(&prog);
    if (ok)
        if (const char *dump = getenv("C4B_JIT_DUMP"))  // tuning aid: cuobjdump -sass on the result
            if (FILE *f = fopen(dump, "wb")) { fwrite(cubin->data(), 1, cubin->size(), f); fclose(f); }
    return ok;
}

static std::string jit_cache_path(const std::string &src) {
    const char *dir = getenv("C4B_JIT_CACHE_DIR");
    if (!dir || !*dir) return std::string();
    char name[64];
    snprintf(name, sizeof name, "/c4bjit_%016llx_%zu.cubin", (unsigned long long)fnv1a(src), src.size());
    return std::string(dir) + name;
}

// cubin -> loaded kernel with its launch limits; nullptr + log on failure
static JitKernel *jit_load(const std::vector<char> &cubin, int threads, bool smem_ring, std::string *log,
                           const char *entry = "c4b_jit_fill") {
    JitKernel *jk = new JitKernel();
    jk->threads = threads;
    bool ok = cudaLibraryLoadData(&jk->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) == cudaSuccess &&
              cudaLibraryGetKernel(&jk->kern, jk->lib, entry) == cudaSuccess;
    if (!ok) *log = std::string("loading the specialised kernel failed: ") + cudaGetErrorString(cudaGetLastError());
    if (ok && smem_ring) {
        jk->blocks_per_sm = 1;  // the ring takes the SM's shared memory
        ok = cudaFuncSetAttribute((const void *)jk->kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kJitSmemRingBytes) == cudaSuccess;
        if (!ok)
            *log = std::string("opting in to the shared-memory ring failed: ") +
                   cudaGetErrorString(cudaGetLastError());
    } else if (ok) {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void *)jk->kern, threads, 0) != cudaSuccess ||
            nb < 1) {
            cudaGetLastError();
            nb = 1;
        }
        jk->blocks_per_sm = nb;
    }
    if (!ok) {
        if (jk->lib) cudaLibraryUnload(jk->lib);
        delete jk;
        return nullptr;
    }
    return jk;
}

static bool read_file(const std::string &path, std::vector<char> *out) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    bool ok = sz > 0;
    if (ok) {
        out->resize((size_t)sz);
        ok = fread(out->data(), 1, (size_t)sz, f) == (size_t)sz;
    }
    fclose(f);
    if (!ok) out->clear();
    return ok;
}

// write-then-rename, so a concurrent reader (another rank, another process) never sees half a cubin
static void jit_store(const std::string &path, const std::vector<char> &cubin) {
    if (path.empty()) return;
    const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
    if (FILE *f = fopen(tmp.c_str(), "wb")) {
        const bool w = fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
        fclose(f);
        if (!w || rename(tmp.c_str(), path.c_str()) != 0) remove(tmp.c_str());
    }
}

// nullptr = no specialised kernel (reason on stderr once per program)
static JitKernel *jit_get(const c4b_model &m, int mode, int threads, bool smem_ring, bool pack_start) {
    static std::mutex mu;
    static std::map<std::string, JitKernel *> cache;
    const std::string src = jit_program_source(m, mode, threads, smem_ring, pack_start);
    // loaded kernels carry per-device state (the shared-memory opt-in, the occupancy answer):
    // one entry per (device, program)
    int device = 0;
    cudaGetDevice(&device);
    const std::string key = std::to_string(device) + ":" + src;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    JitKernel *jk = nullptr;
    std::vector<char> cubin;
    std::string log;
    const std::string path = jit_cache_path(src);
    // a cubin on disk that no longer loads (another toolkit / a torn write) is replaced, not trusted
    if (!path.empty() && read_file(path, &cubin) && !(jk = jit_load(cubin, threads, smem_ring, &log))) {
        remove(path.c_str());
        cubin.clear();
    }
    if (!jk && jit_compile(src, &cubin, &log)) {
        jit_store(path, cubin);
        jk = jit_load(cubin, threads, smem_ring, &log);
    }
    if (!jk)
        fprintf(stderr, "libc4b200: model specialisation unavailable: %s\n", log.c_str());
    cache[key] = jk;
    return jk;
}

// the systolic specialisation of (model, mode); nullptr = not available (reason on stderr once)
static JitKernel *jit_get_sys(const c4b_model &m, int mode, bool pack_start, const SysLayout &L, int warps,
                              int win = 0, bool blk = false) {
    static std::mutex mu;
    static std::map<std::string, JitKernel *> cache;
    const std::string src = jit_sys_program_source(m, mode, pack_start, L, warps, win, blk);
    int device = 0;
    cudaGetDevice(&device);
    const std::string key = std::to_string(device) + ":" + src;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    JitKernel *jk = nullptr;
    std::vector<char> cubin;
    std::string log;
    const std::string path = jit_cache_path(src);
    if (!path.empty() && read_file(path, &cubin) && !(jk = jit_load(cubin, 32, false, &log, "c4b_jit_sys"))) {
        remove(path.c_str());
        cubin.clear();
    }
    if (!jk && jit_compile(src, &cubin, &log)) {
        jit_store(path, cubin);
        jk = jit_load(cubin, 32, false, &log, "c4b_jit_sys");
    }
    if (!jk) fprintf(stderr, "libc4b200: systolic model specialisation unavailable: %s\n", log.c_str());
    cache[key] = jk;
    return jk;
}

// 0 = never, 1 = always, 2 = by batch size.  With a disk cache the compile is paid once per
// model for good (the reference's "bootstrapper" trade), so every batch specialises.
static int jit_policy() {
    const char *env = getenv("C4B_GENERIC_JIT");
    if (!env || !*env) {
        const char *dir = getenv("C4B_JIT_CACHE_DIR");
        return (dir && *dir) ? 1 : 2;
    }
    return atoi(env) ? 1 : 0;
}

constexpr int kJitMaxThreads = 512;
static int jit_threads_for(int maxQ) {
    const int rows = maxQ + 1;
    return rows > 256 ? 512 : (rows > 128 ? 256 : 128);
}

}  // namespace c4b
