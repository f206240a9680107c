// jit.hpp -- the panel-specialised scan kernel.
//
// CompilePanel in the reference builds an Aho-Corasick automaton from the panel
// (/root/reference/core/engine/compiled.go:119-122).  Here the panel is compiled
// into CODE: NVRTC turns the filter plan (plan.hpp) into an sm_100a kernel in
// which every primer column is a fixed register operand, so the hot loop is
// nothing but LOP3 over bit-planes:
//   * each lane keeps the last RS rows of its 32 stripes in REGISTERS (a ring of
//     RS uint4 = {A,C,G,T} plane words); the row loop is unrolled RS times so
//     every ring slot has a static register name -- no shifts, no shared memory,
//     no indexing;
//   * one fully coalesced 512-byte LDG.128 per warp per row streams the genome
//     (0.5 B/bp), issued RS - Lmax rows ahead of first use;
//   * per orientation: AND of the protected 3' columns, AND per pigeonhole
//     block, "at least `need` blocks" threshold -> 32 candidate bits per word;
//   * candidates are rare; one predicated branch per row leaves the fast path.
// Large panels are split into orientation GROUPS; a warp walks a chunk of rows
// once per group so the loop body stays instruction-cache resident while the
// chunk is re-read from L2.
#pragma once

#include <nvrtc.h>

#include <atomic>
#include <cstdlib>
#include <sstream>
#include <thread>
#include <string>
#include <vector>

#include "plan.hpp"

namespace ipcr {

struct JitConfig {
    int stage_rows = 8;       // SR: rows per TMA stage (SR * 512 bytes)
    int stages = 4;           // NB: stages in flight per warp
    int warps = 4;            // warps per CTA
    int group_size = 4;       // orientations per kernel (one pass over the genome each)
    int min_blocks = 0;       // 0 = derive from shared memory
    int list_cap = 64;        // K: candidate words a thread can stage per launch
};

inline JitConfig jit_config_from_env() {
    JitConfig c;
    if (const char *e = getenv("IPCR_JIT_STAGE_ROWS")) c.stage_rows = atoi(e);
    if (const char *e = getenv("IPCR_JIT_STAGES")) c.stages = atoi(e);
    if (const char *e = getenv("IPCR_JIT_WARPS")) c.warps = atoi(e);
    if (const char *e = getenv("IPCR_JIT_GROUP")) c.group_size = atoi(e);
    if (const char *e = getenv("IPCR_JIT_MIN_BLOCKS")) c.min_blocks = atoi(e);
    if (const char *e = getenv("IPCR_JIT_LIST_CAP")) c.list_cap = atoi(e);
    if (c.stage_rows != 2 && c.stage_rows != 4 && c.stage_rows != 8 && c.stage_rows != 16) c.stage_rows = 8;
    c.stages = std::max(2, std::min(c.stages, 8));
    c.warps = std::max(1, std::min(c.warps, 8));
    c.list_cap = std::max(4, c.list_cap);
    c.group_size = std::max(1, std::min(c.group_size, 32));
    return c;
}

// One NVRTC program (= one kernel) per orientation group.  A large panel is scanned as a
// sequence of small-panel kernels: every pass re-streams the packed genome, which stays
// cheaper than the LOP3 work of the pass as long as a group holds >= ~4 orientations.
struct JitProgram {
    std::vector<std::string> sources;       // per group
    std::vector<std::vector<char>> cubins;  // per group
    int n_groups = 0;
    int group_size = 0;
    int ring_slots = 0;
    int stage_rows = 8, stages = 4, warps = 4, list_cap = 64;
    size_t smem_bytes = 0;
};

inline std::string plane_term(int slot, uint8_t cls) {
    static const char *comp = "xyzw";
    std::ostringstream o;
    int n = __builtin_popcount(cls);
    if (n > 1) o << "(";
    bool first = true;
    for (int b = 0; b < 4; b++) {
        if (!((cls >> b) & 1)) continue;
        if (!first) o << "|";
        o << "R" << slot << "." << comp[b];
        first = false;
    }
    if (n > 1) o << ")";
    return o.str();
}

inline std::string and_expr(const std::vector<ColRef> &cols, int k, int RS) {
    if (cols.empty()) return "0xffffffffu";
    std::ostringstream o;
    for (size_t i = 0; i < cols.size(); i++) {
        if (i) o << " & ";
        o << plane_term((k + cols[i].row) % RS, cols[i].cls);
    }
    return o.str();
}

// code computing `c<name>` for one orientation at unrolled step k
inline void emit_orientation(std::ostringstream &o, const OrientPlan &pl, int k, int RS, const std::string &name) {
    const std::string ind = "          ";
    if (pl.need <= 0) {
        o << ind << "const u32 " << name << " = " << and_expr(pl.prot, k, RS) << ";\n";
        return;
    }
    const int B = (int)pl.blocks.size();
    if (pl.need == 1) {
        o << ind << "const u32 " << name << " = ";
        if (!pl.prot.empty()) o << and_expr(pl.prot, k, RS) << " & ";
        o << "(";
        for (int b = 0; b < B; b++) {
            if (b) o << " | ";
            o << "(" << and_expr(pl.blocks[b], k, RS) << ")";
        }
        o << ");\n";
        return;
    }
    // at least `need` of B blocks: running "at least t so far" words
    o << ind << "u32 " << name << ";\n" << ind << "{\n";
    for (int b = 0; b < B; b++)
        o << ind << "  const u32 b" << b << " = " << and_expr(pl.blocks[b], k, RS) << ";\n";
    for (int t = 1; t <= pl.need; t++) o << ind << "  u32 g" << t << " = " << (t == 1 ? "b0" : "0u") << ";\n";
    for (int b = 1; b < B; b++) {
        for (int t = std::min(pl.need, b + 1); t >= 2; t--) o << ind << "  g" << t << " |= g" << (t - 1) << " & b" << b << ";\n";
        o << ind << "  g1 |= b" << b << ";\n";
    }
    o << ind << "  " << name << " = ";
    if (!pl.prot.empty()) o << and_expr(pl.prot, k, RS) << " & ";
    o << "g" << pl.need << ";\n" << ind << "}\n";
}

inline int plan_max_row(const OrientPlan &pl) {
    int m = 0;
    for (const ColRef &c : pl.prot) m = std::max(m, (int)c.row);
    for (const auto &bl : pl.blocks)
        for (const ColRef &c : bl) m = std::max(m, (int)c.row);
    return m;
}

inline std::string jit_generate_group(const ScanPlan &sp, int q0, int q1, int RS, const JitConfig &cfg, int minb) {
    const int SR = cfg.stage_rows, NB = cfg.stages, WARPS = cfg.warps;
    std::ostringstream o;
    o << "// generated by ipcr_b200 jit.hpp: orientations " << q0 << ".." << (q1 - 1) << ", ring " << RS << "\n";
    o << "typedef unsigned int u32;\ntypedef unsigned long long u64;\n";
    o << "#define RS " << RS << "\n#define SR " << SR << "\n#define NB " << NB << "\n#define WARPS " << WARPS << "\n";
    o << "#define STAGE_BYTES (SR * 512)\n";
    o << R"(
__device__ __forceinline__ u32 s2u(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared, completion signalled on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(u32 dst, const void* src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" :: "r"(bar), "r"(parity) : "memory");
}
#define STAGE_BEGIN() { const u32 b_ = gs % NB; mbar_wait(bar_u + 8u * b_, (gs / NB) & 1u); \
                        sp = sbuf + b_ * STAGE_BYTES + lane * 16u; }
#define STAGE_END() { __syncwarp(); \
                      if (issued < nstages) { \
                          if (lane == 0) { const u32 b_ = gs % NB; mbar_expect_tx(bar_u + 8u * b_, STAGE_BYTES); \
                              bulk_g2s(sbuf_u + b_ * STAGE_BYTES, gsrc + (u64)issued * STAGE_BYTES, STAGE_BYTES, bar_u + 8u * b_); } \
                          issued++; } \
                      gs++; }
#define LDROW(t_) (*(const uint4*)(sp + ((t_) % SR) * 512u))
#define STAGE_WORD(c_, q_, w_) { if (cnt < K) my_list[cnt] = make_uint4((c_), (q_), (w_), (u32)tile); cnt++; }
)";
    o << "\nextern \"C\" __global__ void __launch_bounds__(WARPS * 32, " << minb << ")\n";
    o << "ipcr_scan_jit(const uint4* __restrict__ planes, u32 W, u32 rows, u64 n_tiles, u32 chunk_rows,\n"
         "              u64* work_counter, uint4* __restrict__ lists, u32* __restrict__ counts, u32 K) {\n";
    o << "    extern __shared__ __align__(128) unsigned char smem[];\n";
    o << "    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;\n";
    o << "    const unsigned char* sbuf = smem + warp * (NB * STAGE_BYTES);\n";
    o << "    const u32 sbuf_u = s2u(sbuf);\n";
    o << "    const u32 bar_u = s2u(smem + WARPS * NB * STAGE_BYTES) + warp * NB * 8u;\n";
    o << "    if (lane == 0) {\n";
    o << "        for (u32 s = 0; s < NB; s++) mbar_init(bar_u + 8u * s, 1u);\n";
    o << "        asm volatile(\"fence.mbarrier_init.release.cluster;\" ::: \"memory\");\n";
    o << "    }\n    __syncwarp();\n";
    o << "    const u64 gtid = (u64)blockIdx.x * blockDim.x + threadIdx.x;\n";
    o << "    uint4* my_list = lists + gtid * K;\n";
    o << "    u32 cnt = counts[gtid], gs = 0;  // lists accumulate across the group kernels of one scan\n";
    // work item = `chunk_rows` consecutive rows of the VIRTUAL row space [0, n_tiles * W): a warp
    // walks its rows tile segment by tile segment, so it pays the ring prologue only when it
    // enters a new tile, whatever the item size (items need not divide W)
    o << "    const u64 total_rows = n_tiles * (u64)W;\n";
    o << "    const u64 n_items = (total_rows + chunk_rows - 1ull) / chunk_rows;\n";
    o << "    for (;;) {\n";
    o << "      u64 item = 0;\n";
    o << "      if (lane == 0) item = atomicAdd(work_counter, 1ull);\n";
    o << "      item = __shfl_sync(0xffffffffu, item, 0);\n";
    o << "      if (item >= n_items) break;\n";
    o << "      u64 v0 = item * chunk_rows;\n";
    o << "      const u64 v1 = min(total_rows, v0 + chunk_rows);\n";
    o << "     while (v0 < v1) {\n";
    o << "      const u64 tile = v0 / W;\n";
    o << "      const u32 r0 = (u32)(v0 - tile * W);\n";
    o << "      const u32 r1 = (u32)min((u64)W, (u64)r0 + (v1 - v0));\n";
    o << "      v0 += r1 - r0;\n";
    o << "      const u32 nsteps = ((r1 - r0 + RS - 1u) / RS) * RS;\n";
    o << "      const u32 nstages = (nsteps + RS - 1u + SR - 1u) / SR;\n";
    o << "      const unsigned char* gsrc = (const unsigned char*)(planes + (tile * rows + r0) * 32ull);\n";
    o << "      const u32 gs_end = gs + nstages;\n";
    o << "      u32 issued = 0;\n";
    o << "      { const u32 n0 = min((u32)NB, nstages);\n";
    o << "        if (lane == 0) for (u32 i = 0; i < n0; i++) { const u32 b_ = (gs + i) % NB;\n";
    o << "            mbar_expect_tx(bar_u + 8u * b_, STAGE_BYTES);\n";
    o << "            bulk_g2s(sbuf_u + b_ * STAGE_BYTES, gsrc + (u64)i * STAGE_BYTES, STAGE_BYTES, bar_u + 8u * b_); }\n";
    o << "        issued = n0; }\n";
    o << "      const unsigned char* sp = sbuf;\n";
    o << "      uint4 ";
    for (int s = 0; s < RS; s++) o << (s ? ", " : "") << "R" << s;
    o << ";\n";
    for (int t = 0; t + 1 < RS; t++) {
        if (t % SR == 0) o << "      STAGE_BEGIN();\n";
        o << "      R" << t << " = LDROW(" << t << ");\n";
        if (t % SR == SR - 1) o << "      STAGE_END();\n";
    }
    o << "      for (u32 w = r0; w < r0 + nsteps; w += RS) {\n";
    for (int k = 0; k < RS; k++) {
        const int slot = (k + RS - 1) % RS;
        const int t = RS - 1 + k;
        o << "          // step " << k << "\n";
        if (t % SR == 0) o << "          STAGE_BEGIN();\n";
        o << "          R" << slot << " = LDROW(" << t << ");\n";
        if (t % SR == SR - 1) o << "          STAGE_END();\n";
        o << "          {\n";
        for (int q = q0; q < q1; q++) {
            std::ostringstream nm;
            nm << "c" << q;
            emit_orientation(o, sp.o[q], k, RS, nm.str());
        }
        o << "          if ((";
        for (int q = q0; q < q1; q++) o << (q > q0 ? " | " : "") << "c" << q;
        o << ") != 0u) {\n";
        o << "            if (w + " << k << "u < r1) {\n";
        for (int q = q0; q < q1; q++)
            o << "              if (c" << q << ") STAGE_WORD(c" << q << ", " << q << "u, w + " << k << "u);\n";
        o << "            }\n          }\n";
        o << "          }\n";
    }
    o << "      }\n";
    // the last stage is only partially read (no STAGE_END): count it as consumed so that the
    // next tile starts on a fresh buffer / barrier phase
    o << "      gs = gs_end;\n";
    o << "     }\n";
    o << "    }\n";
    o << "    counts[gtid] = cnt;\n";
    o << "}\n";
    return o.str();
}

inline JitProgram jit_generate(const Panel &panel, int variant, const JitConfig &cfg) {
    const ScanPlan &sp = panel.plan[variant];
    const int U = (int)sp.o.size();
    JitProgram prog;
    int lmax = 1;
    for (const OrientPlan &pl : sp.o) lmax = std::max(lmax, plan_max_row(pl) + 1);
    const int SR = cfg.stage_rows, NB = cfg.stages, WARPS = cfg.warps;
    const int RS = ((lmax + 1 + SR - 1) / SR) * SR;  // ring: window + 1 row of LDS look-ahead
    prog.ring_slots = RS;
    prog.stage_rows = SR;
    prog.stages = NB;
    prog.warps = WARPS;
    prog.list_cap = cfg.list_cap;
    prog.smem_bytes = (size_t)WARPS * NB * SR * 512 + (size_t)WARPS * NB * 8;
    const int G = cfg.group_size;
    prog.group_size = G;
    prog.n_groups = U ? (U + G - 1) / G : 0;
    int minb = cfg.min_blocks;
    if (minb <= 0) minb = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)(224 * 1024) / (prog.smem_bytes + 1024)));
    for (int g = 0; g < prog.n_groups; g++)
        prog.sources.push_back(jit_generate_group(sp, g * G, std::min(U, (g + 1) * G), RS, cfg, minb));
    return prog;
}

inline std::string jit_generate_source(const Panel &panel, int variant) {
    JitProgram p = jit_generate(panel, variant, jit_config_from_env());
    std::string all;
    for (const std::string &s : p.sources) all += s + "\n";
    return all;
}

// NVRTC: source -> sm_100a cubin.  Returns "" on success, else the log.
inline std::string jit_compile(const std::string &src, std::vector<char> &cubin) {
    nvrtcProgram prog;
    if (nvrtcCreateProgram(&prog, src.c_str(), "ipcr_scan_jit.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS)
        return "nvrtcCreateProgram failed";
    const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--extra-device-vectorization"};
    nvrtcResult r = nvrtcCompileProgram(prog, 3, opts);
    std::string log;
    size_t ls = 0;
    nvrtcGetProgramLogSize(prog, &ls);
    if (ls > 1) {
        log.resize(ls);
        nvrtcGetProgramLog(prog, &log[0]);
    }
    if (r != NVRTC_SUCCESS) {
        nvrtcDestroyProgram(&prog);
        return "nvrtc: " + std::string(nvrtcGetErrorString(r)) + "\n" + log;
    }
    size_t cs = 0;
    if (nvrtcGetCUBINSize(prog, &cs) != NVRTC_SUCCESS || cs == 0) {
        nvrtcDestroyProgram(&prog);
        return "nvrtc produced no cubin\n" + log;
    }
    cubin.resize(cs);
    nvrtcGetCUBIN(prog, cubin.data());
    nvrtcDestroyProgram(&prog);
    return "";
}

// compiles every group (NVRTC is thread-safe; one program per worker at a time)
inline std::string jit_compile_all(JitProgram &prog) {
    const size_t n = prog.sources.size();
    prog.cubins.assign(n, {});
    std::vector<std::string> logs(n);
    std::atomic<size_t> next{0};
    unsigned nt = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), 32u));
    if (const char *e = getenv("IPCR_JIT_COMPILE_THREADS")) nt = (unsigned)std::max(1, atoi(e));
    nt = (unsigned)std::min<size_t>(nt, n);
    auto work = [&]() {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= n) break;
            logs[i] = jit_compile(prog.sources[i], prog.cubins[i]);
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
    for (size_t i = 0; i < n; i++)
        if (!logs[i].empty()) return logs[i];
    return "";
}

struct JitKernel {
    std::vector<cudaLibrary_t> libraries;
    std::vector<cudaKernel_t> kernels;  // one per orientation group
    bool ok = false;
    std::string log;
};

}  // namespace ipcr
