// jit_launch.hpp -- compile-once / load-per-device / launch of the NVRTC kernel, and the
// bookkeeping that turns its per-thread candidate-word lists into verify work.
#pragma once

namespace ipcr {

// cumulative position x orientation work after which compiling pays for itself
constexpr double kJitWorkThreshold = 2.0e11;

struct JitRun {
    uint64_t n_entries = 0;  // staged candidate words
    uint64_t n_bits = 0;     // candidate positions in them
    uint32_t n_threads = 0;
    uint32_t K = 0;
};

// returns the loaded kernel for this device/variant, compiling on first use; nullptr = not available
static JitKernel *jit_get(ipcr_scratch *s, int variant, double work) {
    ipcr_panel *p = s->panel;
    const int mode = p->host.jit_mode == 3 ? 1 : p->host.jit_mode;  // 3: seed index first, then automatic
    if (mode <= 0) return nullptr;
    JitKernel *k = &s->tables->jit[variant];
    std::lock_guard<std::mutex> lock(p->mu);
    p->jit_work += work;
    if (mode == 1 && !p->jit_done[variant] && p->jit_work < kJitWorkThreshold) return nullptr;
    if (!p->jit_done[variant]) {
        p->jit_done[variant] = true;
        p->jit_prog[variant] = jit_generate(p->host, variant, jit_config_from_env());
        p->jit_log[variant] = jit_compile_all(p->jit_prog[variant]);
        if (!p->jit_log[variant].empty()) p->jit_prog[variant].cubins.clear();
    }
    if (p->jit_prog[variant].cubins.empty()) {
        if (mode >= 2) throw Error("JIT compile of the scan kernel failed: " + p->jit_log[variant]);
        return nullptr;
    }
    if (!s->tables->jit_tried[variant]) {
        s->tables->jit_tried[variant] = true;
        cudaError_t e = cudaSuccess;
        for (const std::vector<char> &bin : p->jit_prog[variant].cubins) {
            cudaLibrary_t lib = nullptr;
            cudaKernel_t kern = nullptr;
            e = cudaLibraryLoadData(&lib, bin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
            if (e == cudaSuccess) e = cudaLibraryGetKernel(&kern, lib, "ipcr_scan_jit");
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)p->jit_prog[variant].smem_bytes);
            if (e != cudaSuccess) break;
            k->libraries.push_back(lib);
            k->kernels.push_back(kern);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            k->log = std::string("loading the JIT kernel failed: ") + cudaGetErrorString(e);
            k->ok = false;
        } else {
            k->ok = true;
        }
    }
    if (!k->ok) {
        if (mode >= 2) throw Error(k->log);
        return nullptr;
    }
    return k;
}

// Runs the specialised filter over the whole genome.  On success the candidate words are
// in s->lists with offsets in s->list_offs; returns false if the table-driven kernel
// should be used instead.
static bool jit_filter(ipcr_scratch *s, const GenomeData &g, int variant, JitRun &run) {
    ipcr_panel *p = s->panel;
    JitKernel *k = jit_get(s, variant, (double)g.lay.total_len * (double)p->host.orients.size());
    if (!k) return false;
    const JitProgram &prog = p->jit_prog[variant];
    cudaStream_t st = s->stream;
    const int threads = prog.warps * 32;
    int blocks_per_sm = (int)std::max<size_t>(1, (size_t)(227 * 1024) / (prog.smem_bytes + 1024));
    int max_regs = 0;
    for (cudaKernel_t kern : k->kernels) {
        cudaFuncAttributes fa{};
        if (cudaFuncGetAttributes(&fa, (const void *)kern) == cudaSuccess) max_regs = std::max(max_regs, fa.numRegs);
        else cudaGetLastError();
    }
    if (max_regs > 0) blocks_per_sm = std::min(blocks_per_sm, std::max(1, 65536 / (max_regs * threads)));
    blocks_per_sm = std::min(blocks_per_sm, 2048 / threads);
    if (const char *e = getenv("IPCR_JIT_BLOCKS_PER_SM")) blocks_per_sm = std::max(1, atoi(e));
    const size_t n_groups = k->kernels.size();
    // rows per work item: the largest ring-multiple divisor of W that still yields >= 6 items per
    // resident warp (measured best on 3 Gbp: 240 of W = 480), never below two ring lengths.  Items
    // that divide W never straddle a tile, so no unrolled step and no ring prologue is wasted.
    const uint64_t resident_warps = (uint64_t)s->sm_count * blocks_per_sm * prog.warps;
    const uint32_t RSn = (uint32_t)prog.ring_slots;
    const uint64_t total_rows = g.lay.n_tiles * (uint64_t)g.lay.W;
    uint32_t chunk_rows = g.lay.W;
    for (uint32_t m = std::max(2u, g.lay.W / RSn); m >= 2; m--) {
        const uint32_t c = m * RSn;
        if (c > g.lay.W || g.lay.W % c != 0) continue;
        chunk_rows = c;
        if (g.lay.n_tiles * (g.lay.W / c) >= 6 * resident_warps) break;
    }
    if (const char *e = getenv("IPCR_JIT_CHUNK")) chunk_rows = (uint32_t)std::max(16, atoi(e));
    const uint64_t n_items = (total_rows + chunk_rows - 1) / chunk_rows;
    const unsigned grid = (unsigned)std::max<uint64_t>(
        1, std::min<uint64_t>((n_items + prog.warps - 1) / prog.warps, (uint64_t)s->sm_count * blocks_per_sm));
    // Consecutive group kernels are independent except for their candidate lists, so they are
    // issued round-robin on a few streams, each with its own set of lists: while one kernel
    // drains (tail, launch gap) the next one already fills the idle SMs.
    // (measured: +20 % on 375 Mbp launches, -7 % on 3 Gbp launches where two co-resident kernels
    // fight over the instruction cache, hence only below ~1.2 Gbp per launch)
    int n_sets = g.lay.total_len < 1200000000ull ? 2 : 1;
    if (const char *e = getenv("IPCR_JIT_STREAMS")) n_sets = std::max(1, std::min(atoi(e), 4));
    n_sets = (int)std::min<size_t>((size_t)n_sets, n_groups);
    while ((int)s->aux_streams.size() < n_sets - 1) {
        cudaStream_t x;
        CU(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
        s->aux_streams.push_back(x);
        cudaEvent_t e1;
        CU(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
        s->aux_events.push_back(e1);
    }
    const uint32_t grid_threads = grid * (uint32_t)threads;
    const uint32_t n_threads = grid_threads * (uint32_t)n_sets;  // one list per (set, thread)
    // list capacity per thread: twice the modelled candidate-word count + slack
    double expect = 0.0;
    for (const OrientPlan &op : p->host.plan[variant].o) expect += op.cand_prob;
    uint32_t K = (uint32_t)std::min<double>(1 << 20, std::max<double>(prog.list_cap,
                                            2.0 * expect * (double)g.lay.total_len / n_threads + prog.list_cap));

    for (int attempt = 0; attempt < 2; attempt++) {
        s->lists.ensure((size_t)n_threads * K * sizeof(uint4));
        s->list_counts.ensure((size_t)n_threads * 4);
        s->list_clamped.ensure((size_t)n_threads * 4);
        s->list_offs.ensure(((size_t)n_threads + 1) * 4);
        s->tmp.ensure(scan_tmp_words(n_threads) * 4);
        s->work_counters.ensure(n_groups * sizeof(unsigned long long));
        CU(cudaMemsetAsync(s->counters.p, 0, 8 * sizeof(unsigned long long), st));
        CU(cudaMemsetAsync(s->work_counters.p, 0, n_groups * sizeof(unsigned long long), st));
        CU(cudaMemsetAsync(s->list_counts.p, 0, (size_t)n_threads * 4, st));
        CU(cudaEventRecord(s->ev[3], st));
        for (int x = 0; x + 1 < n_sets; x++) CU(cudaStreamWaitEvent(s->aux_streams[x], s->ev[3], 0));
        const uint4 *planes = g.planes.as<uint4>();
        uint32_t W = g.lay.W, rows = g.lay.rows;
        uint64_t n_tiles = g.lay.n_tiles;
        for (size_t gi = 0; gi < n_groups; gi++) {
            const int set = (int)(gi % (size_t)n_sets);
            cudaStream_t ls = set == 0 ? st : s->aux_streams[set - 1];
            uint4 *lists = s->lists.as<uint4>() + (size_t)set * grid_threads * K;
            uint32_t *counts = s->list_counts.as<uint32_t>() + (size_t)set * grid_threads;
            unsigned long long *work_counter = s->work_counters.as<unsigned long long>() + gi;
            void *args[] = {&planes, &W, &rows, &n_tiles, &chunk_rows, &work_counter, &lists, &counts, &K};
            cudaError_t e = cudaLaunchKernel((const void *)k->kernels[gi], dim3(grid), dim3(threads), args,
                                             prog.smem_bytes, ls);
            if (e != cudaSuccess) {
                cudaGetLastError();
                throw Error(std::string("launch of ipcr_scan_jit failed: ") + cudaGetErrorString(e));
            }
            s->launches++;
        }
        for (int x = 0; x + 1 < n_sets; x++) {  // join
            CU(cudaEventRecord(s->aux_events[x], s->aux_streams[x]));
            CU(cudaStreamWaitEvent(st, s->aux_events[x], 0));
        }
        CU(cudaEventRecord(s->ev[4], st));  // end of the filter proper
        uint4 *lists = s->lists.as<uint4>();
        uint32_t *counts = s->list_counts.as<uint32_t>();
        WordListArgs w{};
        w.lists = lists;
        w.counts = counts;
        w.clamped = s->list_clamped.as<uint32_t>();
        w.offs = s->list_offs.as<uint32_t>();
        w.n_threads = n_threads;
        w.K = K;
        w.stats = s->counters.as<unsigned long long>() + 4;
        w.lay = g.lay;
        w.variant = (uint32_t)variant;
        k_list_prepare<<<(n_threads + 255) / 256, 256, 0, st>>>(w);
        check_launch("k_list_prepare");
        s->launches++;
        exclusive_scan(st, w.clamped, s->list_offs.as<uint32_t>(), n_threads, s->tmp.as<uint32_t>(), &s->launches);
        CU(cudaMemcpyAsync(s->h_counts.as<unsigned long long>() + 4, s->counters.as<unsigned long long>() + 4,
                           2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(s->h_counts.as<unsigned long long>() + 6, s->list_offs.as<uint32_t>() + n_threads,
                           sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        const uint64_t max_count = s->h_counts.as<unsigned long long>()[4];
        run.n_bits = s->h_counts.as<unsigned long long>()[5];
        run.n_entries = *reinterpret_cast<uint32_t *>(s->h_counts.as<unsigned long long>() + 6);
        run.n_threads = n_threads;
        run.K = K;
        if (max_count <= K) return true;
        // some thread overflowed its list: retry once with room for the worst thread
        if (attempt == 1 || 2 * max_count * (uint64_t)n_threads * sizeof(uint4) > (8ull << 30)) return false;
        // work is handed out dynamically, so per-thread counts vary from run to run: leave headroom
        K = (uint32_t)(max_count + max_count / 2 + 64);
    }
    return false;
}

}  // namespace ipcr
