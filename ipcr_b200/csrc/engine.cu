// engine.cu -- host orchestration + C ABI (include/ipcr_gpu.h) of the B200 scan.
//
// Mirrors the call protocol of the reference engine
// (/root/reference/core/engine/compiled.go:96-267, internal/pipeline/pipeline.go:52-125):
//   panel_create   == CompilePanel          (once, shared read-only)
//   scratch_create == NewSimulationScratch  (one per worker: stream + buffers)
//   scan           == ForEachCompiledProduct (per record / chunk)
// Stage 1 (find sites): pack -> filter -> verify.  Stage 2 (pair): keys -> sort ->
// hit cap / compaction -> join count -> scan -> join fill -> D2H.
// There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "device_types.cuh"
#include "fasta_host.hpp"
#include "pack_host.hpp"
#include "jit.hpp"
#include "kernels.cuh"
#include "plan.hpp"
#include "seed.hpp"
#include "seed_kernels.cuh"
#include "sortscan.cuh"
#include "pair_small.cuh"
#include "sort_products.cuh"

namespace ipcr {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

static std::string fmt(const char *f, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, f);
    vsnprintf(buf, sizeof buf, f, ap);
    va_end(ap);
    return buf;
}

#define CU(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e__ = (x);                                                                  \
        if (e__ != cudaSuccess)                                                                 \
            throw ipcr::Error(ipcr::fmt("CUDA error %s at %s:%d: %s", cudaGetErrorName(e__),    \
                                        __FILE__, __LINE__, cudaGetErrorString(e__)));          \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes) {
        if (bytes <= cap) return;
        if (p) CU(cudaFree(p));
        p = nullptr;
        cap = 0;
        size_t want = std::max(bytes, (size_t)256);
        want = (want + 255) & ~(size_t)255;
        CU(cudaMalloc(&p, want));
        cap = want;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes) {
        if (bytes <= cap) return;
        if (p) CU(cudaFreeHost(p));
        p = nullptr;
        cap = 0;
        size_t want = std::max(bytes, (size_t)4096);
        CU(cudaMallocHost(&p, want));
        cap = want;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
    ~PinBuf() {
        if (p) cudaFreeHost(p);
    }
    PinBuf() = default;
    PinBuf(const PinBuf &) = delete;
    PinBuf &operator=(const PinBuf &) = delete;
};

// Pageable host memory -> device at PCIe speed.  cudaMemcpyAsync from pageable memory is staged by the driver on ONE
// thread (~11 GB/s measured for the feeder's malloc'ed packed records), and page-locking the source costs ~0.5 s
// per GB -- more than a one-shot run's whole copy.  Here kThreads host threads copy 2 MB chunks into a small ring of
// pinned slots (allocated once per scratch) and queue the H2D of each slot on their own stream; `mark`s (slab
// numbers) let the consumer stream wait for exactly the chunks a slab needs.
struct HostStager {
    static constexpr int kThreads = 6, kSlots = 2;
    static constexpr size_t kChunk = 2u << 20;
    struct Copy { void *dst; const void *src; size_t bytes; int mark; };
    PinBuf ring;
    cudaStream_t stream[kThreads] = {};
    cudaEvent_t slot_ev[kThreads][kSlots] = {};
    std::vector<cudaEvent_t> mark_ev;  // [mark * kThreads + t]
    std::vector<Copy> jobs;
    std::vector<std::thread> th;
    std::atomic<int> progress[kThreads];  // marks thread t has recorded so far
    std::atomic<bool> failed{false};
    std::string error;
    std::mutex mu;
    int device = 0, n_marks = 0;

    void add(void *dst, const void *src, size_t bytes, int mark) {  // marks must not decrease
        const uint8_t *s8 = static_cast<const uint8_t *>(src);
        uint8_t *d8 = static_cast<uint8_t *>(dst);
        for (size_t o = 0; o < bytes; o += kChunk) jobs.push_back(Copy{d8 + o, s8 + o, std::min(kChunk, bytes - o), mark});
    }
    void start(int dev, int marks) {
        device = dev;
        n_marks = marks;
        ring.ensure((size_t)kThreads * kSlots * kChunk);
        for (int t = 0; t < kThreads; t++) {
            if (!stream[t]) CU(cudaStreamCreateWithFlags(&stream[t], cudaStreamNonBlocking));
            for (int k = 0; k < kSlots; k++)
                if (!slot_ev[t][k]) CU(cudaEventCreateWithFlags(&slot_ev[t][k], cudaEventDisableTiming));
            progress[t].store(0);
        }
        while (mark_ev.size() < (size_t)marks * kThreads) {
            cudaEvent_t e;
            CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            mark_ev.push_back(e);
        }
        failed.store(false);
        error.clear();
        for (int t = 0; t < kThreads; t++) th.emplace_back([this, t] { run(t); });
    }
    void run(int t) {
        int done = 0;  // marks recorded
        try {
            CU(cudaSetDevice(device));
            uint8_t *slots = ring.as<uint8_t>() + (size_t)t * kSlots * kChunk;
            size_t mine = 0;
            for (size_t c = (size_t)t; c < jobs.size(); c += kThreads, mine++) {
                const Copy &j = jobs[c];
                for (; done < j.mark; done++) {  // everything of the earlier marks is queued
                    CU(cudaEventRecord(mark_ev[(size_t)done * kThreads + t], stream[t]));
                    progress[t].store(done + 1, std::memory_order_release);
                }
                const int k = (int)(mine % kSlots);
                if (mine >= (size_t)kSlots) CU(cudaEventSynchronize(slot_ev[t][k]));  // the slot's last copy has left
                memcpy(slots + (size_t)k * kChunk, j.src, j.bytes);
                CU(cudaMemcpyAsync(j.dst, slots + (size_t)k * kChunk, j.bytes, cudaMemcpyHostToDevice, stream[t]));
                CU(cudaEventRecord(slot_ev[t][k], stream[t]));
            }
            for (; done < n_marks; done++) {
                CU(cudaEventRecord(mark_ev[(size_t)done * kThreads + t], stream[t]));
                progress[t].store(done + 1, std::memory_order_release);
            }
        } catch (const std::exception &e) {
            std::lock_guard<std::mutex> lock(mu);
            error = e.what();
            failed.store(true);
            progress[t].store(1 << 30, std::memory_order_release);
        }
    }
    // makes `consumer` wait for every chunk of marks <= mark (blocks the host until those chunks are QUEUED)
    void wait_mark(int mark, cudaStream_t consumer) {
        for (int t = 0; t < kThreads; t++) {
            while (progress[t].load(std::memory_order_acquire) <= mark) std::this_thread::yield();
            if (failed.load()) break;
            CU(cudaStreamWaitEvent(consumer, mark_ev[(size_t)mark * kThreads + t], 0));
        }
        if (failed.load()) {
            finish();
            throw Error("staged host-to-device copy failed: " + error);
        }
    }
    void finish() {  // joins the threads (their copies may still be in flight on their streams)
        for (auto &x : th) x.join();
        th.clear();
        jobs.clear();
    }
    ~HostStager() {
        for (auto &x : th) x.join();
        for (int t = 0; t < kThreads; t++) {
            if (stream[t]) cudaStreamSynchronize(stream[t]);
            for (int k = 0; k < kSlots; k++)
                if (slot_ev[t][k]) cudaEventDestroy(slot_ev[t][k]);
            if (stream[t]) cudaStreamDestroy(stream[t]);
        }
        for (auto &e : mark_ev) cudaEventDestroy(e);
    }
};

// ---- panel tables on one device ---------------------------------------------
struct PlanTables {
    DevBuf progs, cols, blk_len, slot_cls;
    uint32_t n_slots = 0;
};

struct DevTables {
    DevBuf orients, cls, pairs, oplanes, use_off, use_list;
    PlanTables plan[2];
    JitKernel jit[2];
    bool jit_tried[2] = {false, false};
    DevBuf seed_table[2], seed_bloom[2], seed_bitmap2[2], seed_coloff[2];  // seed-index filter, per plan variant
    bool seed_ready[2] = {false, false};
};

}  // namespace ipcr

struct ipcr_panel {
    ipcr::Panel host;
    std::mutex mu;
    std::map<int, std::unique_ptr<ipcr::DevTables>> dev;
    // NVRTC output, shared by every device (compiled at most once per plan variant)
    ipcr::JitProgram jit_prog[2];
    std::string jit_log[2];
    bool jit_done[2] = {false, false};
    double jit_work = 0.0;
    // k-mer index of the whole panel (seed.hpp), built on first use per plan variant
    ipcr::SeedIndex seed[2];
    bool seed_done[2] = {false, false};
};

namespace ipcr {

static void upload(DevBuf &b, const void *src, size_t bytes, cudaStream_t st) {
    b.ensure(std::max(bytes, (size_t)16));
    if (bytes) CU(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, st));
}

static DevTables *tables_for(ipcr_panel *p, int device, cudaStream_t st) {
    std::lock_guard<std::mutex> lock(p->mu);
    auto it = p->dev.find(device);
    if (it != p->dev.end()) return it->second.get();
    auto t = std::make_unique<DevTables>();
    const Panel &h = p->host;
    std::vector<OrientDev> od;
    std::vector<uint8_t> cls;
    const int tw = h.cfg.terminal_window;
    for (const Orient &o : h.orients) {
        OrientDev d{};
        d.L = (uint16_t)o.L;
        d.seed_off = (uint16_t)o.seed_off;
        d.seed_len = (uint16_t)o.seed_len;
        d.rc_type = o.rc_type;
        d.seeded = o.seeded;
        d.cls_off = (uint32_t)cls.size();
        d.j0[0] = (uint16_t)h.plan[0].o[od.size()].j0;
        d.j0[1] = (uint16_t)h.plan[1].o[od.size()].j0;
        uint64_t pm = 0;
        if (tw > 0) {
            for (int j = 0; j < o.L; j++) {
                const bool prot = o.rc_type ? (j < tw) : (j >= o.L - tw);
                if (prot) pm |= 1ull << j;
            }
        }
        d.prot_mask = pm;
        cls.insert(cls.end(), o.cls.begin(), o.cls.end());
        od.push_back(d);
    }
    std::vector<PairDev> pd;
    for (const PairRoles &r : h.roles) {
        PairDev d{};
        d.fwdA = (uint16_t)r.fwdA;
        d.fwdB = (uint16_t)r.fwdB;
        d.rcA = (uint16_t)r.rcA;
        d.rcB = (uint16_t)r.rcB;
        d.alen = (uint16_t)r.alen;
        d.blen = (uint16_t)r.blen;
        d.min_len = r.min_len;
        d.max_len = r.max_len;
        pd.push_back(d);
    }
    std::vector<uint64_t> oplanes;  // [orientation][A, C, G, T]: columns whose class accepts the base
    for (const Orient &o : h.orients) {
        uint64_t pl[4] = {0, 0, 0, 0};
        for (int j = 0; j < o.L && j < 64; j++) {
            const uint8_t c = o.cls[(size_t)j];
            if (c & 1u) pl[0] |= 1ull << j;
            if (c & 2u) pl[1] |= 1ull << j;
            if (c & 4u) pl[2] |= 1ull << j;
            if (c & 8u) pl[3] |= 1ull << j;
        }
        oplanes.insert(oplanes.end(), pl, pl + 4);
    }
    upload(t->oplanes, oplanes.data(), oplanes.size() * sizeof(uint64_t), st);
    // (pair, direction) slots per forward orientation, pair-major inside an orientation
    std::vector<uint32_t> use_off(h.orients.size() + 1, 0), use_list;
    {
        std::vector<std::vector<uint32_t>> uses(h.orients.size());
        for (size_t pi = 0; pi < h.roles.size(); pi++) {
            uses[(size_t)h.roles[pi].fwdA].push_back((uint32_t)(pi << 1));
            uses[(size_t)h.roles[pi].fwdB].push_back((uint32_t)(pi << 1) | 1u);
        }
        for (size_t q = 0; q < uses.size(); q++) {
            use_off[q + 1] = use_off[q] + (uint32_t)uses[q].size();
            use_list.insert(use_list.end(), uses[q].begin(), uses[q].end());
        }
    }
    upload(t->use_off, use_off.data(), use_off.size() * sizeof(uint32_t), st);
    upload(t->use_list, use_list.data(), use_list.size() * sizeof(uint32_t), st);
    upload(t->orients, od.data(), od.size() * sizeof(OrientDev), st);
    upload(t->cls, cls.data(), cls.size(), st);
    upload(t->pairs, pd.data(), pd.size() * sizeof(PairDev), st);
    for (int v = 0; v < 2; v++) {
        const ScanPlan &sp = h.plan[v];
        std::vector<OrientProg> progs;
        std::vector<ColDev> cols;
        std::vector<uint8_t> blk_len, slot_cls;
        int slot_of[16];
        for (int i = 0; i < 16; i++) slot_of[i] = -1;
        auto slot = [&](uint8_t c) {
            if (slot_of[c] < 0) {
                slot_of[c] = (int)slot_cls.size();
                slot_cls.push_back(c);
            }
            return (uint8_t)slot_of[c];
        };
        for (const OrientPlan &op : sp.o) {
            OrientProg g{};
            g.col_begin = (uint32_t)cols.size();
            g.blk_begin = (uint32_t)blk_len.size();
            g.n_prot = (uint16_t)op.prot.size();
            g.n_blocks = (uint16_t)op.blocks.size();
            g.need = (uint16_t)std::max(op.need, 0);
            g.j0 = (uint16_t)op.j0;
            for (const ColRef &c : op.prot) cols.push_back(ColDev{c.row, slot(c.cls)});
            for (const auto &bl : op.blocks) {
                blk_len.push_back((uint8_t)bl.size());
                for (const ColRef &c : bl) cols.push_back(ColDev{c.row, slot(c.cls)});
            }
            progs.push_back(g);
        }
        if (slot_cls.empty()) slot_cls.push_back(1);
        PlanTables &pt = t->plan[v];
        pt.n_slots = (uint32_t)slot_cls.size();
        upload(pt.progs, progs.data(), progs.size() * sizeof(OrientProg), st);
        upload(pt.cols, cols.data(), cols.size() * sizeof(ColDev), st);
        upload(pt.blk_len, blk_len.data(), blk_len.size(), st);
        upload(pt.slot_cls, slot_cls.data(), slot_cls.size(), st);
    }
    CU(cudaStreamSynchronize(st));  // host vectors die here
    DevTables *raw = t.get();
    p->dev.emplace(device, std::move(t));
    return raw;
}

// ---- genome --------------------------------------------------------------------
struct GenomeData {
    int device = 0;
    DevBuf ascii_owned;
    const uint8_t *ascii = nullptr;  // device pointer to the concatenated buffer
    GenomeLayout lay{};
    DevBuf planes, segs_dev, seg_reset;
    DevBuf codes, inval;  // seed-index filter: 2-bit codes + hard-mismatch bitmap
    uint64_t n_groups = 0;  // 32-base groups in codes/inval (incl. padding)
    bool planes_ready = false, codes_ready = false;  // packed on demand, whichever filter runs
    std::mutex pack_mu;
    std::vector<SegmentDev> segs;
    std::vector<ipcr_record_info> recs;
    bool any_reset = false;
    uint64_t own_bases = 0;
    const ipcr_record_info *recs_dev_ptr = nullptr;  // chain: record table uploaded together with the segments
    std::vector<uint8_t> last_tab;                   // what segs_dev holds (skip the copy when nothing changed)
    // packed input (host-packed codes + bitmap already in device memory): nothing to pack, no ASCII
    const uint32_t *codes_ext = nullptr, *inval_ext = nullptr;
};

}  // namespace ipcr

struct ipcr_genome {
    ipcr::GenomeData g;
};

struct ipcr_scratch {
    ipcr_panel *panel = nullptr;
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // H2D of later slabs overlaps the scan of earlier ones
    std::vector<cudaStream_t> aux_streams;  // group kernels of the specialised filter run round-robin
    std::vector<cudaEvent_t> aux_events;
    std::vector<cudaEvent_t> slab_ev;
    cudaEvent_t copy_ev[2] = {};
    cudaEvent_t ev[8] = {};
    ipcr::DevTables *tables = nullptr;
    ipcr::GenomeData own;  // genome of ipcr_gpu_scan / scan_batch calls (buffers are reused)
    ipcr::DevBuf lists, list_counts, list_clamped, list_offs, work_counters;  // specialised filter
    ipcr::DevBuf dirty, dcand, dirty_over, ocand;  // seed-index filter: dirty windows, their candidates, overflow
    ipcr::DevBuf cand, counters, sites, keys0, keys1, vals0, vals1, keep, keep_off, tmp, fkeys, fs, recs_dev,
        entry_tasks, task_off, task_cnt, prod_off, products;
    ipcr::PinBuf h_counts, h_products, h_sites, h_stage, h_reset;
    uint64_t n_cand = 0;
    uint64_t seed_cap_hint = 0, seed_dirty_hint = 0;  // capacities that sufficed last time (no re-scan)
    // asynchronous chain (pair_small.cuh): stage 1 writes a site block, one CTA pairs, ONE readback
    ipcr::DevBuf block, rec_reset_dev, gather;  // own site block; merged per-record flags; blocks of all shards / slabs
    ipcr::PinBuf h_args;                        // pinned staging of the small per-call tables (segments, records)
    size_t args_used = 0;
    bool inflight = false;                      // work was enqueued and not yet waited for
    uint64_t site_cap_hint = 0, task_cap_hint = 0, prod_cap_hint = 0;
    uint64_t last_copied = 0;                   // products the chain's single readback carried
    bool raw_hint = false;                      // a record with a reset byte was seen: start with the raw-rc index
    const ipcr_product *out_ptr = nullptr;      // products of the last call (pinned, scratch-owned)
    std::vector<ipcr_scratch *> peers;          // one scratch per further device (multi-GPU inside one process)
    cudaEvent_t done_ev = nullptr;              // peer: its block has arrived on the primary device
    bool own_laid = false;                      // multi-device retry: the slice is already resident
    ipcr::DevBuf seg_reset_dummy, chunk_ctr, verify_done, pair_dbg, runs_dev;
    uint64_t pending_publish = 0;  // peer exchange: step number the next stage 1 publishes from inside k_verify
    // peer exchange (one process per GPU): own double-buffered block, exported through CUDA IPC; the peers' mappings
    ipcr::DevBuf peer_buf, peer_tab[2];
    std::vector<void *> peer_map;  // cudaIpcOpenMemHandle results (NULL for this rank)
    uint32_t peer_n = 0, peer_rank = 0;
    uint64_t peer_step = 0, peer_cap = 0, peer_bb = 0;
    size_t peer_recs = 0;
    ipcr::DevBuf sp_in, sp_out, sp_k[4], sp_ka, sp_kb, sp_ia, sp_ib, sp_tie, sp_rr, sp_ro, sp_pr;  // ipcr_gpu_sort_products
    std::vector<std::unique_ptr<ipcr::GenomeData>> slab_g;  // streamed slabs: packed codes per slab (all in flight at once)
    std::vector<cudaEvent_t> slab_t;
    std::unique_ptr<ipcr::HostStager> stager;  // pageable packed records -> device (scan_packed_chain)
    // records of the last scan that are still resident (for the device-side visitors)
    std::vector<uint64_t> last_rec_base, last_rec_len;
    const uint8_t *last_ascii = nullptr;
    bool last_valid = false;
    ipcr::DevBuf v_rec_base, v_rec_len, v_products, v_cls, v_hits, v_amp, v_offs;
    ipcr::PinBuf h_hits, h_amp;
    ipcr_timing timing{};
    uint32_t launches = 0;
    ~ipcr_scratch() {
        for (auto &e : ev)
            if (e) cudaEventDestroy(e);
        for (auto &e : slab_ev)
            if (e) cudaEventDestroy(e);
        for (auto &e : slab_t)
            if (e) cudaEventDestroy(e);
        if (done_ev) cudaEventDestroy(done_ev);
        for (auto &e : copy_ev)
            if (e) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        for (auto &x : aux_streams) cudaStreamDestroy(x);
        for (auto &e : aux_events) cudaEventDestroy(e);
    }
};

namespace ipcr {

static uint32_t choose_W(uint64_t total, int sm_count) {
    // rows per stripe: large W keeps the repeated rows (31/W) cheap; the scan kernels split tiles
    // into row chunks for parallelism, so ~2 tiles per SM are enough.  480 = lcm-friendly: a
    // multiple of 16 (pack loads), 24 and 32 (register-ring lengths of the specialised kernel).
    const uint64_t want_tiles = (uint64_t)sm_count * 2ull;
    static const uint32_t cand[] = {480, 240, 96, 32};
    for (uint32_t W : cand)
        if ((total + (uint64_t)kStripesPerTile * W - 1) / ((uint64_t)kStripesPerTile * W) >= want_tiles) return W;
    return 32;
}

static void check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw Error(fmt("launch of %s failed: %s", what, cudaGetErrorString(e)));
}

// which filter serves this panel: 0 table-driven, 1 NVRTC-specialised, 2 seed index
enum { kPathGeneric = 0, kPathJit = 1, kPathSeed = 2 };

static const SeedIndex &seed_host(ipcr_panel *p, int variant) {
    std::lock_guard<std::mutex> lock(p->mu);
    if (!p->seed_done[variant]) {
        p->seed[variant] = seed_build(p->host, variant);
        p->seed_done[variant] = true;
    }
    return p->seed[variant];
}

// jit_mode: 0 table-driven only, 1 automatic, 2 force the specialised kernel, 3 force the seed index.
// Automatic = the seed index: measured on B200 it beats the bit-parallel filter from 4 to 400
// orientations and at mm 0..3 (profiles/), because its cost does not grow with the panel.  The
// bit-parallel filter takes over per call when the input is too dirty (seed_dirty_max()).
static bool prefer_seed(ipcr_panel *p) {
    const int mode = p->host.jit_mode;
    return (mode == 3 || mode == 1) && !p->host.orients.empty();
}

static void launch_pack_planes(ipcr_scratch *s, GenomeData &g, uint8_t *rec_reset = nullptr) {
    cudaStream_t st = s->stream;
    g.lay.W = choose_W(g.lay.total_len, s->sm_count);
    g.lay.rows = g.lay.W + kExtRows;
    const uint64_t per_tile = (uint64_t)kStripesPerTile * g.lay.W;
    g.lay.n_tiles = std::max<uint64_t>(1, (g.lay.total_len + per_tile - 1) / per_tile);
    // + 80 rows of slack: the unrolled JIT loop reads up to 2*ring rows past a tile's end
    g.planes.ensure(((size_t)g.lay.n_tiles * g.lay.rows + 80) * kLanes * sizeof(uint4));
    if ((reinterpret_cast<uintptr_t>(g.ascii) & 15u) == 0 && g.lay.W % 16 == 0) {
        k_pack<<<(unsigned)g.lay.n_tiles, 1024, 0, st>>>(g.ascii, g.lay, g.planes.as<uint4>(),
                                                         g.segs_dev.as<SegmentDev>(), (uint32_t)g.segs.size(),
                                                         g.seg_reset.as<uint32_t>(), rec_reset);
    } else {  // unaligned adopted buffer: byte-wise variant
        dim3 grid((unsigned)g.lay.n_tiles, (g.lay.rows + 31) / 32), block(32, 32);
        k_pack_bytes<<<grid, block, 0, st>>>(g.ascii, g.lay, g.planes.as<uint4>(), g.segs_dev.as<SegmentDev>(),
                                             (uint32_t)g.segs.size(), g.seg_reset.as<uint32_t>(), rec_reset);
    }
    check_launch("k_pack");
    s->launches++;
    g.planes_ready = true;
}

static void launch_pack_codes(ipcr_scratch *s, GenomeData &g, uint8_t *rec_reset = nullptr) {
    cudaStream_t st = s->stream;
    const uint64_t n_chunks = (g.lay.total_len + kSeedChunk - 1) / kSeedChunk;
    g.n_groups = n_chunks * (kSeedChunk / 32u) + 1;  // + one padding group (look-ahead of the last lane)
    g.codes.ensure((size_t)g.n_groups * 8 + 64);
    g.inval.ensure((size_t)g.n_groups * 4 + 64);
    Pack2Args a{};
    a.ascii = g.ascii;
    a.total_len = g.lay.total_len;
    a.n_groups = g.n_groups;
    a.codes = g.codes.as<uint32_t>();
    a.inval = g.inval.as<uint32_t>();
    a.segs = g.segs_dev.as<SegmentDev>();
    a.n_segs = (uint32_t)g.segs.size();
    a.seg_reset = g.seg_reset.as<uint32_t>();
    a.rec_reset = rec_reset;
    a.aligned = (reinterpret_cast<uintptr_t>(g.ascii) & 15u) == 0;
    const unsigned grid = (unsigned)std::min<uint64_t>((g.n_groups + 255) / 256, (uint64_t)s->sm_count * 32);
    k_pack2<<<grid, 256, 0, st>>>(a);
    check_launch("k_pack2");
    s->launches++;
    g.codes_ready = true;
}

// K1 on the genome's concatenated ASCII buffer, in the format the panel's filter reads
static void pack_genome(ipcr_scratch *s, GenomeData &g) {
    cudaStream_t st = s->stream;
    g.last_tab.clear();  // segs_dev is rewritten below in the multi-kernel path's own layout
    g.planes_ready = g.codes_ready = false;
    upload(g.segs_dev, g.segs.data(), g.segs.size() * sizeof(SegmentDev), st);
    g.seg_reset.ensure(g.segs.size() * sizeof(uint32_t));
    CU(cudaMemsetAsync(g.seg_reset.p, 0, g.segs.size() * sizeof(uint32_t), st));
    if (prefer_seed(s->panel)) launch_pack_codes(s, g);
    else launch_pack_planes(s, g);
    s->h_reset.ensure(g.segs.size() * sizeof(uint32_t));
    CU(cudaMemcpyAsync(s->h_reset.p, g.seg_reset.p, g.segs.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    g.any_reset = false;
    for (auto &r : g.recs) r.has_reset_byte = 0;
    for (size_t i = 0; i < g.segs.size(); i++) {
        if (s->h_reset.as<uint32_t>()[i]) {
            g.recs[g.segs[i].rec_idx].has_reset_byte = 1;
            g.any_reset = true;
        }
    }
}

// a genome handle may meet panels that want the other format: pack it then (once)
static void ensure_planes(ipcr_scratch *s, GenomeData &g) {
    std::lock_guard<std::mutex> lock(g.pack_mu);
    if (!g.planes_ready) {
        launch_pack_planes(s, g);
        CU(cudaStreamSynchronize(s->stream));
    }
}

static void ensure_codes(ipcr_scratch *s, GenomeData &g) {
    std::lock_guard<std::mutex> lock(g.pack_mu);
    if (!g.codes_ready) {
        launch_pack_codes(s, g);
        CU(cudaStreamSynchronize(s->stream));
    }
}

// builds segment / record tables and queues the H2D copy of host segments; `device_src` =
// segment pointers are device memory (adopt).  Nothing is packed yet.
static void layout_genome(ipcr_scratch *s, GenomeData &g, const ipcr_segment *segs, size_t n_segs, size_t n_recs,
                          bool device_src) {
    g.device = s->device;
    g.segs.clear();
    g.recs.assign(n_recs, ipcr_record_info{});
    g.own_bases = 0;
    g.planes_ready = g.codes_ready = false;
    g.any_reset = false;
    uint64_t off = 0;
    for (size_t i = 0; i < n_segs; i++) {
        const ipcr_segment &in = segs[i];
        if (in.rec_idx >= n_recs) throw Error(fmt("segment %zu: rec_idx %u out of range", i, in.rec_idx));
        if (in.own_len > in.len) throw Error(fmt("segment %zu: own_len > len", i));
        SegmentDev d{};
        if (device_src) {
            if (i > 0 && in.seq < segs[i - 1].seq + segs[i - 1].len)
                throw Error("adopted segments must be ascending and non-overlapping in one device buffer");
            d.global_off = (uint64_t)(in.seq - segs[0].seq);
        } else {
            d.global_off = off;
        }
        d.len = in.len;
        d.own_len = in.own_len;
        d.rec_off = in.rec_off;
        d.rec_idx = in.rec_idx;
        off = d.global_off + in.len;
        g.segs.push_back(d);
        g.recs[in.rec_idx].len = std::max<uint64_t>(g.recs[in.rec_idx].len, in.rec_off + in.len);
        g.own_bases += in.own_len;
    }
    g.lay.total_len = off;
    if (g.lay.total_len >= (1ull << kCandPosBits)) throw Error("genome too large for one device buffer");
    if (n_segs == 0 || g.lay.total_len == 0) {
        g.ascii = nullptr;
        s->last_valid = false;
        return;
    }
    cudaStream_t st = s->stream;
    if (device_src) {
        g.ascii = segs[0].seq;
    } else {
        g.ascii_owned.ensure((size_t)g.lay.total_len + 64);
        g.ascii = g.ascii_owned.as<uint8_t>();
        CU(cudaEventRecord(s->ev[0], st));
        for (size_t i = 0; i < n_segs; i++)
            if (segs[i].len)
                CU(cudaMemcpyAsync(g.ascii_owned.as<uint8_t>() + g.segs[i].global_off, segs[i].seq, segs[i].len,
                                   cudaMemcpyHostToDevice, st));
    }
    CU(cudaEventRecord(s->ev[1], st));
    // whole records stay addressable for the device-side visitors
    s->last_valid = true;
    s->last_ascii = g.ascii;
    s->last_rec_base.assign(n_recs, 0);
    s->last_rec_len.assign(n_recs, 0);
    std::vector<char> seen(n_recs, 0);
    for (const SegmentDev &d : g.segs) {
        if (d.rec_off != 0 || d.own_len != d.len || seen[d.rec_idx]) s->last_valid = false;
        seen[d.rec_idx] = 1;
        s->last_rec_base[d.rec_idx] = d.global_off;
        s->last_rec_len[d.rec_idx] = d.len;
    }
}

// layout + K1 (synchronous: the per-record reset flags are back on the host when it returns)
static void build_genome(ipcr_scratch *s, GenomeData &g, const ipcr_segment *segs, size_t n_segs, size_t n_recs,
                         bool device_src) {
    layout_genome(s, g, segs, n_segs, n_recs, device_src);
    if (g.lay.total_len == 0) return;
    pack_genome(s, g);
    CU(cudaEventRecord(s->ev[2], s->stream));
}

static uint64_t read_counter(ipcr_scratch *s, int idx) {
    CU(cudaMemcpyAsync(s->h_counts.as<unsigned long long>() + idx, s->counters.as<unsigned long long>() + idx,
                       sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return s->h_counts.as<unsigned long long>()[idx];
}

struct JitRun;
static bool jit_filter(ipcr_scratch *s, const GenomeData &g, int variant, JitRun &run);
struct VerifyArgs;
static bool seed_filter(ipcr_scratch *s, GenomeData &g, int variant, VerifyArgs &v, uint64_t &n_sites);

static void launch_filter_generic(ipcr_scratch *s, const GenomeData &g, int variant, uint64_t cand_cap) {
    const PlanTables &pt = s->tables->plan[variant];
    FilterArgs a{};
    a.planes = g.planes.as<uint4>();
    a.lay = g.lay;
    a.progs = pt.progs.as<OrientProg>();
    a.cols = pt.cols.as<ColDev>();
    a.blk_len = pt.blk_len.as<uint8_t>();
    a.slot_cls = pt.slot_cls.as<uint8_t>();
    a.n_orient = (uint32_t)s->panel->host.orients.size();
    a.n_slots = pt.n_slots;
    a.cand = s->cand.as<unsigned long long>();
    a.cand_count = s->counters.as<unsigned long long>();
    a.cand_cap = cand_cap;
    a.tile_begin = 0;
    a.tile_end = g.lay.n_tiles;
    const size_t per_warp = (size_t)32 * pt.n_slots * 32 * sizeof(uint32_t);
    int wpb = (int)std::min<size_t>(4, (size_t)(160 * 1024) / per_warp);
    if (wpb < 1) wpb = 1;
    const size_t smem = per_warp * wpb;
    static std::mutex attr_mu;
    {
        std::lock_guard<std::mutex> lock(attr_mu);
        CU(cudaFuncSetAttribute(k_filter_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    const uint64_t want_blocks = (g.lay.n_tiles + wpb - 1) / wpb;
    const unsigned grid = (unsigned)std::min<uint64_t>(want_blocks, (uint64_t)s->sm_count * 8);
    k_filter_generic<<<grid, wpb * 32, smem, s->stream>>>(a);
    check_launch("k_filter_generic");
    s->launches++;
}

// Stage 1.  Leaves verified sites in s->sites (device) and returns their count.
static uint64_t find_sites(ipcr_scratch *s, GenomeData &g, bool force_raw);

static uint32_t bit_length(uint64_t v) {
    uint32_t b = 0;
    while (v) {
        b++;
        v >>= 1;
    }
    return b;
}

static KeyLayout key_layout(const Panel &h, const ipcr_record_info *recs, size_t n_recs) {
    uint64_t max_len = 1;
    for (size_t r = 0; r < n_recs; r++) max_len = std::max<uint64_t>(max_len, recs[r].len);
    KeyLayout kl{};
    kl.pos_bits = std::max(1u, bit_length(max_len));
    kl.n_orient = (uint32_t)h.orients.size();
    const uint32_t prefix_bits = std::max(1u, bit_length((uint64_t)n_recs * kl.n_orient));
    kl.key_bits = kl.pos_bits + 1 + prefix_bits;
    return kl;
}

// Stage 2.  d_sites: device site list (any order).  Products end up in s->h_products.
static uint64_t pair_sites(ipcr_scratch *s, const ipcr_site *d_sites, uint64_t n_sites,
                           const ipcr_record_info *recs, size_t n_recs) {
    const Panel &h = s->panel->host;
    cudaStream_t st = s->stream;
    const size_t n_pairs = h.roles.size();
    if (n_sites == 0 || n_pairs == 0 || n_recs == 0) return 0;
    if (n_sites >= (1ull << 32)) throw Error("too many primer sites for one pairing pass");
    const KeyLayout kl = key_layout(h, recs, n_recs);
    if (kl.key_bits > 62) throw Error("record count x orientations x record length exceeds the 62-bit sort key");
    const uint64_t n_entries = (uint64_t)n_recs * n_pairs * 2ull;
    if (n_entries >= (1ull << 31)) throw Error("records x pairs too large for one pairing pass; use smaller batches");

    ScanParams prm{};
    prm.max_mm = h.cfg.max_mm;
    prm.terminal_window = h.cfg.terminal_window;
    prm.hit_cap = h.cfg.hit_cap;
    prm.circular = h.cfg.circular;

    upload(s->recs_dev, recs, n_recs * sizeof(ipcr_record_info), st);
    s->keys0.ensure(n_sites * 8);
    s->keys1.ensure(n_sites * 8);
    s->vals0.ensure(n_sites * 4);
    s->vals1.ensure(n_sites * 4);
    s->keep.ensure(n_sites * 4);
    s->keep_off.ensure((n_sites + 1) * 4);
    s->fkeys.ensure(n_sites * 8);
    s->fs.ensure(n_sites * sizeof(FinalSite));
    uint64_t tmp_words = std::max(sort_tmp_words(n_sites), scan_tmp_words(std::max<uint64_t>(n_sites, n_entries)));
    s->tmp.ensure(tmp_words * 4);

    const unsigned gs = (unsigned)((n_sites + 255) / 256);
    k_site_keys<<<gs, 256, 0, st>>>(d_sites, n_sites, s->tables->orients.as<OrientDev>(),
                                    s->recs_dev.as<ipcr_record_info>(), nullptr, prm, kl, s->keys0.as<uint64_t>(),
                                    s->vals0.as<uint32_t>());
    check_launch("k_site_keys");
    s->launches++;
    const int where = radix_sort(st, s->keys0.as<uint64_t>(), s->vals0.as<uint32_t>(), s->keys1.as<uint64_t>(),
                                 s->vals1.as<uint32_t>(), n_sites, kl.key_bits + 1, s->tmp.as<uint32_t>(),
                                 &s->launches);
    check_launch("radix_sort");
    const uint64_t *skeys = where ? s->keys1.as<uint64_t>() : s->keys0.as<uint64_t>();
    const uint32_t *svals = where ? s->vals1.as<uint32_t>() : s->vals0.as<uint32_t>();
    k_site_keep<<<gs, 256, 0, st>>>(skeys, svals, n_sites, d_sites, prm, kl, s->keep.as<uint32_t>());
    check_launch("k_site_keep");
    exclusive_scan(st, s->keep.as<uint32_t>(), s->keep_off.as<uint32_t>(), n_sites, s->tmp.as<uint32_t>(),
                   &s->launches);
    k_site_compact<<<gs, 256, 0, st>>>(skeys, svals, n_sites, d_sites, s->keep.as<uint32_t>(),
                                       s->keep_off.as<uint32_t>(), s->fkeys.as<uint64_t>(), s->fs.as<FinalSite>());
    check_launch("k_site_compact");
    s->launches += 2;
    uint32_t n_final = 0;
    CU(cudaMemcpyAsync(s->h_counts.as<unsigned long long>() + 2, s->keep_off.as<uint32_t>() + n_sites,
                       sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    n_final = *reinterpret_cast<uint32_t *>(s->h_counts.as<unsigned long long>() + 2);
    if (n_final == 0) return 0;

    JoinArgs ja{};
    ja.fkeys = s->fkeys.as<uint64_t>();
    ja.fs = s->fs.as<FinalSite>();
    ja.n_sites = n_final;
    ja.pairs = s->tables->pairs.as<PairDev>();
    ja.n_pairs = (uint32_t)n_pairs;
    ja.n_recs = (uint32_t)n_recs;
    ja.recs = s->recs_dev.as<ipcr_record_info>();
    ja.prm = prm;
    ja.kl = kl;
    s->entry_tasks.ensure(n_entries * 4);
    s->task_off.ensure((n_entries + 1) * 4);
    ja.entry_tasks = s->entry_tasks.as<uint32_t>();
    ja.task_off = s->task_off.as<uint32_t>();
    k_join_entries<<<(unsigned)((n_entries + 255) / 256), 256, 0, st>>>(ja);
    check_launch("k_join_entries");
    s->launches++;
    exclusive_scan(st, ja.entry_tasks, s->task_off.as<uint32_t>(), n_entries, s->tmp.as<uint32_t>(), &s->launches);
    CU(cudaMemcpyAsync(s->h_counts.as<unsigned long long>() + 3, s->task_off.as<uint32_t>() + n_entries,
                       sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const uint64_t n_tasks = *reinterpret_cast<uint32_t *>(s->h_counts.as<unsigned long long>() + 3);
    {  // the scans are 32-bit: one orientation can serve many pairs, so tasks can outnumber the sites
        uint64_t max_uses = 1;
        std::vector<uint32_t> uses(h.orients.size(), 0);
        for (const PairRoles &r : h.roles) {
            max_uses = std::max<uint64_t>(max_uses, ++uses[(size_t)r.fwdA]);
            max_uses = std::max<uint64_t>(max_uses, ++uses[(size_t)r.fwdB]);
        }
        if ((uint64_t)n_final * max_uses >= (1ull << 32))
            throw Error("too many forward sites x pairs for one pairing pass; use smaller batches");
    }
    if (n_tasks == 0) return 0;
    ja.n_tasks = n_tasks;
    s->task_cnt.ensure(n_tasks * 4);
    s->prod_off.ensure((n_tasks + 1) * 4);
    s->tmp.ensure(scan_tmp_words(n_tasks) * 4);
    ja.task_cnt = s->task_cnt.as<uint32_t>();
    ja.prod_off = s->prod_off.as<uint32_t>();
    const unsigned gt = (unsigned)((n_tasks + 127) / 128);
    k_join<false><<<gt, 128, 0, st>>>(ja);
    check_launch("k_join<count>");
    s->launches++;
    exclusive_scan(st, ja.task_cnt, s->prod_off.as<uint32_t>(), n_tasks, s->tmp.as<uint32_t>(), &s->launches);
    // the true total in 64 bits beside the 32-bit scan: a wrapped scan must fail, not truncate
    CU(cudaMemsetAsync(s->counters.as<unsigned long long>() + 6, 0, sizeof(unsigned long long), st));
    k_sum_u32<<<(unsigned)std::min<uint64_t>((n_tasks + 255) / 256, 1024), 256, 0, st>>>(
        ja.task_cnt, n_tasks, s->counters.as<unsigned long long>() + 6);
    check_launch("k_sum_u32");
    s->launches++;
    CU(cudaMemcpyAsync(s->h_counts.as<unsigned long long>() + 4, s->prod_off.as<uint32_t>() + n_tasks,
                       sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(s->h_counts.as<unsigned long long>() + 6, s->counters.as<unsigned long long>() + 6,
                       sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const uint64_t n_prod = *reinterpret_cast<uint32_t *>(s->h_counts.as<unsigned long long>() + 4);
    if (s->h_counts.as<unsigned long long>()[6] != n_prod)
        throw Error(fmt("%llu products do not fit one pairing pass; use smaller batches or a hit cap",
                        (unsigned long long)s->h_counts.as<unsigned long long>()[6]));
    if (n_prod == 0) return 0;
    s->products.ensure(n_prod * sizeof(ipcr_product));
    ja.products = s->products.as<ipcr_product>();
    ja.prod_cap = n_prod;
    k_join<true><<<gt, 128, 0, st>>>(ja);
    check_launch("k_join<fill>");
    s->launches++;
    CU(cudaEventRecord(s->ev[6], st));
    s->h_products.ensure(n_prod * sizeof(ipcr_product));
    CU(cudaMemcpyAsync(s->h_products.p, s->products.p, n_prod * sizeof(ipcr_product), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    s->out_ptr = s->h_products.as<ipcr_product>();
    return n_prod;
}

static float ev_ms(cudaEvent_t a, cudaEvent_t b) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) {
        cudaGetLastError();
        return 0.f;
    }
    return ms;
}

static void bind(ipcr_scratch *s) { CU(cudaSetDevice(s->device)); }

static uint64_t scan_genome_impl(ipcr_scratch *s, GenomeData &g, bool had_upload) {
    cudaStream_t st = s->stream;
    if (!had_upload) {
        CU(cudaEventRecord(s->ev[0], st));
        CU(cudaEventRecord(s->ev[1], st));
        CU(cudaEventRecord(s->ev[2], st));
    }
    CU(cudaEventRecord(s->ev[3], st));
    CU(cudaEventRecord(s->ev[4], st));
    CU(cudaEventRecord(s->ev[5], st));
    const uint64_t n_sites = find_sites(s, g, false);
    CU(cudaEventRecord(s->ev[6], st));
    const uint64_t n_prod = pair_sites(s, s->sites.as<ipcr_site>(), n_sites, g.recs.data(), g.recs.size());
    if (n_prod == 0) CU(cudaEventRecord(s->ev[6], st));
    CU(cudaEventRecord(s->ev[7], st));
    CU(cudaStreamSynchronize(st));
    ipcr_timing &t = s->timing;
    t.h2d_ms = ev_ms(s->ev[0], s->ev[1]);
    t.pack_ms = ev_ms(s->ev[1], s->ev[2]);
    t.filter_ms = ev_ms(s->ev[3], s->ev[4]);
    t.verify_ms = ev_ms(s->ev[4], s->ev[5]);
    t.pair_ms = ev_ms(s->ev[5], s->ev[6]);
    t.d2h_ms = ev_ms(s->ev[6], s->ev[7]);
    t.total_ms = ev_ms(s->ev[0], s->ev[7]);
    t.products = n_prod;
    t.kernel_launches = s->launches;
    return n_prod;
}

// Host records -> products with the H2D copy of slab k+1.. running under the scan of slab k
// (the end-to-end path is PCIe-bound once the scan is fast).  The
// concatenated records are cut into slabs exactly like multi-GPU shards (64-base halo, a site
// belongs to the slab holding its start); stage 1 runs per slab, stage 2 once.
static uint64_t scan_batch_streamed(ipcr_scratch *s, const ipcr_seq *recs, size_t n_recs, int n_slabs) {
    const Panel &h = s->panel->host;
    cudaStream_t st = s->stream, cs = s->copy_stream;
    constexpr uint64_t kHalo = 64;
    std::vector<uint64_t> roff(n_recs + 1, 0);
    for (size_t r = 0; r < n_recs; r++) roff[r + 1] = roff[r] + recs[r].len;
    const uint64_t total = roff[n_recs];
    std::vector<ipcr_record_info> rinfo(n_recs, ipcr_record_info{});
    for (size_t r = 0; r < n_recs; r++) rinfo[r].len = recs[r].len;
    GenomeData &g = s->own;
    g.device = s->device;
    g.ascii_owned.ensure((size_t)total + 64);
    uint8_t *base = g.ascii_owned.as<uint8_t>();
    std::vector<uint64_t> cut((size_t)n_slabs + 1);
    for (int k = 0; k <= n_slabs; k++) cut[k] = k == n_slabs ? total : ((total * (uint64_t)k / n_slabs) & ~15ull);
    while ((int)s->slab_ev.size() < n_slabs) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        s->slab_ev.push_back(e);
    }
    // 1. every copy is queued now; slab k's event fires once [cut[k], cut[k+1] + halo) is resident
    CU(cudaEventRecord(s->ev[0], st));
    CU(cudaEventRecord(s->copy_ev[0], cs));
    uint64_t copied_to = 0;
    for (int k = 0; k < n_slabs; k++) {
        const uint64_t want = std::min(total, cut[k + 1] + kHalo);
        for (size_t r = 0; r < n_recs && copied_to < want; r++) {
            const uint64_t a = std::max(copied_to, roff[r]), b = std::min(want, roff[r + 1]);
            if (a < b)
                CU(cudaMemcpyAsync(base + a, recs[r].seq + (a - roff[r]), b - a, cudaMemcpyHostToDevice, cs));
        }
        copied_to = std::max(copied_to, want);
        CU(cudaEventRecord(s->slab_ev[k], cs));
    }
    CU(cudaEventRecord(s->copy_ev[1], cs));
    // 2. stage 1 per slab
    std::vector<ipcr_site> all_sites;
    bool used_plain_variant = false;
    uint64_t cand_total = 0;
    float pack_ms = 0.f, filter_ms = 0.f, verify_ms = 0.f;
    uint32_t used_jit = 0;
    for (int k = 0; k < n_slabs; k++) {
        const uint64_t lo = cut[k], hi = cut[k + 1], end = std::min(total, hi + kHalo);
        if (lo >= hi) continue;
        g.segs.clear();
        g.recs.assign(n_recs, ipcr_record_info{});
        for (size_t r = 0; r < n_recs; r++) {
            g.recs[r].len = recs[r].len;
            const uint64_t s0 = std::max(roff[r], lo), s1 = std::min(roff[r + 1], end), own1 = std::min(roff[r + 1], hi);
            if (s0 >= s1 || own1 <= s0) continue;
            SegmentDev d{};
            d.global_off = s0 - lo;
            d.len = s1 - s0;
            d.own_len = own1 - s0;
            d.rec_off = s0 - roff[r];
            d.rec_idx = (uint32_t)r;
            g.segs.push_back(d);
        }
        if (g.segs.empty()) continue;
        g.lay.total_len = end - lo;
        g.ascii = base + lo;
        g.own_bases = hi - lo;
        CU(cudaStreamWaitEvent(st, s->slab_ev[k], 0));
        CU(cudaEventRecord(s->ev[1], st));
        pack_genome(s, g);
        CU(cudaEventRecord(s->ev[2], st));
        for (size_t r = 0; r < n_recs; r++) rinfo[r].has_reset_byte |= g.recs[r].has_reset_byte;
        const uint64_t n = find_sites(s, g, false);
        if (s->timing.plan_variant == 0) used_plain_variant = true;
        used_jit |= s->timing.used_jit;
        cand_total += s->timing.candidates;
        if (n) {
            s->h_sites.ensure(n * sizeof(ipcr_site));
            CU(cudaMemcpyAsync(s->h_sites.p, s->sites.p, n * sizeof(ipcr_site), cudaMemcpyDeviceToHost, st));
        }
        CU(cudaStreamSynchronize(st));
        pack_ms += ev_ms(s->ev[1], s->ev[2]);
        filter_ms += ev_ms(s->ev[3], s->ev[4]);
        verify_ms += ev_ms(s->ev[4], s->ev[5]);
        if (n) all_sites.insert(all_sites.end(), s->h_sites.as<ipcr_site>(), s->h_sites.as<ipcr_site>() + n);
    }
    // hit-cap parity: a record with a reset byte anywhere needs raw rc matches from EVERY slab
    bool any_reset = false;
    for (const auto &ri : rinfo) any_reset |= ri.has_reset_byte != 0;
    if (any_reset && used_plain_variant && h.cfg.hit_cap > 0 && h.cfg.max_mm > 0 && h.cfg.terminal_window > 0) {
        all_sites.clear();
        g.segs.clear();
        g.recs = rinfo;
        for (size_t r = 0; r < n_recs; r++) {
            if (!recs[r].len) continue;
            SegmentDev d{};
            d.global_off = roff[r];
            d.len = d.own_len = recs[r].len;
            d.rec_idx = (uint32_t)r;
            g.segs.push_back(d);
        }
        g.lay.total_len = total;
        g.ascii = base;
        CU(cudaStreamWaitEvent(st, s->slab_ev[n_slabs - 1], 0));
        pack_genome(s, g);
        const uint64_t n = find_sites(s, g, true);
        cand_total = s->timing.candidates;
        if (n) {
            s->h_sites.ensure(n * sizeof(ipcr_site));
            CU(cudaMemcpyAsync(s->h_sites.p, s->sites.p, n * sizeof(ipcr_site), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            all_sites.assign(s->h_sites.as<ipcr_site>(), s->h_sites.as<ipcr_site>() + n);
        }
    }
    // 3. stage 2 once
    CU(cudaEventRecord(s->ev[5], st));
    const uint64_t n_sites = all_sites.size();
    if (n_sites) {
        s->sites.ensure(n_sites * sizeof(ipcr_site));
        CU(cudaMemcpyAsync(s->sites.p, all_sites.data(), n_sites * sizeof(ipcr_site), cudaMemcpyHostToDevice, st));
    }
    CU(cudaEventRecord(s->ev[6], st));
    const uint64_t n_prod = pair_sites(s, s->sites.as<ipcr_site>(), n_sites, rinfo.data(), n_recs);
    if (n_prod == 0) CU(cudaEventRecord(s->ev[6], st));
    CU(cudaEventRecord(s->ev[7], st));
    CU(cudaStreamSynchronize(st));
    CU(cudaStreamSynchronize(cs));
    g.recs = rinfo;
    s->last_valid = true;
    s->last_ascii = base;
    s->last_rec_base.assign(roff.begin(), roff.begin() + n_recs);
    s->last_rec_len.resize(n_recs);
    for (size_t r = 0; r < n_recs; r++) s->last_rec_len[r] = recs[r].len;
    ipcr_timing &t = s->timing;
    t.h2d_ms = ev_ms(s->copy_ev[0], s->copy_ev[1]);
    t.pack_ms = pack_ms;
    t.filter_ms = filter_ms;
    t.verify_ms = verify_ms;
    t.pair_ms = ev_ms(s->ev[5], s->ev[6]);
    t.d2h_ms = ev_ms(s->ev[6], s->ev[7]);
    t.total_ms = ev_ms(s->ev[0], s->ev[7]);
    t.candidates = cand_total;
    t.sites = n_sites;
    t.products = n_prod;
    t.used_jit = used_jit;
    t.kernel_launches = s->launches;
    return n_prod;
}

static void set_err(char *err, size_t errlen, const std::string &m) {
    if (!err || errlen == 0) return;
    snprintf(err, errlen, "%s", m.c_str());
}

}  // namespace ipcr

#include "jit_launch.hpp"

namespace ipcr {

// more dirty windows (or unresolved dirty-window candidates) than this and the bit-parallel filter,
// which handles non-ACGT bytes natively, is the better tool
static uint64_t seed_dirty_max() {
    const char *e = getenv("IPCR_SEED_DIRTY_MAX");  // read per call: tests lower it
    return e ? (uint64_t)atoll(e) : (1ull << 26);
}

static const SeedIndex &seed_get(ipcr_scratch *s, int variant) {
    ipcr_panel *p = s->panel;
    const SeedIndex &ix = seed_host(p, variant);
    DevTables *t = s->tables;
    std::lock_guard<std::mutex> lock(p->mu);
    if (!t->seed_ready[variant]) {
        upload(t->seed_table[variant], ix.slots.data(), ix.slots.size() * sizeof(uint64_t), s->stream);
        upload(t->seed_bloom[variant], ix.bloom.data(), ix.bloom.size() * sizeof(uint32_t), s->stream);
        upload(t->seed_bitmap2[variant], ix.bitmap2.data(), ix.bitmap2.size() * sizeof(uint32_t), s->stream);
        upload(t->seed_coloff[variant], ix.col_off.data(), ix.col_off.size() * sizeof(uint16_t), s->stream);
        CU(cudaStreamSynchronize(s->stream));
        static std::mutex attr_mu;
        std::lock_guard<std::mutex> l2(attr_mu);
        for (auto fn : {k_seed_scan<1>, k_seed_scan<2>, k_seed_scan<2, 1>, k_seed_scan<2, 4>, k_seed_scan<2, 5>,
                        k_seed_scan<2, 2>, k_seed_scan<2, 8>,
                        k_seed_scan<1, 0, 2>, k_seed_scan<2, 0, 2>, k_seed_scan<1, 0, 4>, k_seed_scan<2, 0, 4>,
                        k_seed_scan<1, 1, 2>, k_seed_scan<2, 1, 2>, k_seed_scan<1, 1, 4>, k_seed_scan<2, 1, 4>})
            CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSeedBloomWords * 4)));
        t->seed_ready[variant] = true;
    }
    return ix;
}

static void launch_seed_scan(ipcr_scratch *s, const SeedIndex &ix, const SeedScanArgs &a0) {
    SeedScanArgs a = a0;
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((a.n_chunks + 31) / 32, (uint64_t)s->sm_count));
    const size_t smem = kSeedBloomWords * 4;
    int var = 0;
    if (const char *e = getenv("IPCR_SEED_VAR")) var = atoi(e) & 15;
    if (var & 4) {
        s->chunk_ctr.ensure(64);
        CU(cudaMemsetAsync(s->chunk_ctr.p, 0, 8, s->stream));
        a.chunk_counter = s->chunk_ctr.as<unsigned long long>();
    }
    a.prefetch = ix.stride > 1 ? 1u : 0u;
    if (const char *e = getenv("IPCR_SEED_PREFETCH")) a.prefetch = atoi(e) != 0 ? 1u : 0u;
    if (ix.pigeon) k_seed_scan<3><<<grid, 1024, kSeedHalfWords * 2 * 4, s->stream>>>(a);
    else if (!ix.use_bloom) k_seed_scan<0><<<grid, 1024, 0, s->stream>>>(a);
    else if (ix.stride == 4 && var == 1 && ix.bloom_k == 1) k_seed_scan<1, 1, 4><<<grid, 1024, smem, s->stream>>>(a);
    else if (ix.stride == 4 && var == 1) k_seed_scan<2, 1, 4><<<grid, 1024, smem, s->stream>>>(a);
    else if (ix.stride == 2 && var == 1 && ix.bloom_k == 1) k_seed_scan<1, 1, 2><<<grid, 1024, smem, s->stream>>>(a);
    else if (ix.stride == 2 && var == 1) k_seed_scan<2, 1, 2><<<grid, 1024, smem, s->stream>>>(a);
    else if (ix.stride == 4 && ix.bloom_k == 1) k_seed_scan<1, 0, 4><<<grid, 1024, smem, s->stream>>>(a);
    else if (ix.stride == 4) k_seed_scan<2, 0, 4><<<grid, 1024, smem, s->stream>>>(a);
    else if (ix.stride == 2 && ix.bloom_k == 1) k_seed_scan<1, 0, 2><<<grid, 1024, smem, s->stream>>>(a);
    else if (ix.stride == 2) k_seed_scan<2, 0, 2><<<grid, 1024, smem, s->stream>>>(a);
    else if (ix.bloom_k == 1) k_seed_scan<1><<<grid, 1024, smem, s->stream>>>(a);
    else if (var == 1) k_seed_scan<2, 1><<<grid, 1024, smem, s->stream>>>(a);
    else if (var == 4) k_seed_scan<2, 4><<<grid, 1024, smem, s->stream>>>(a);
    else if (var == 5) k_seed_scan<2, 5><<<grid, 1024, smem, s->stream>>>(a);
    else if (var == 2) k_seed_scan<2, 2><<<grid, 1024, smem, s->stream>>>(a);
    else if (var == 8) k_seed_scan<2, 8><<<grid, 1024, smem, s->stream>>>(a);
    else k_seed_scan<2><<<grid, 1024, smem, s->stream>>>(a);
    check_launch("k_seed_scan");
    s->launches++;
}

// Seed-index filter + verify.  Returns false (nothing changed) when the input is too dirty for
// it; on success the verified sites are in s->sites.
static bool seed_filter(ipcr_scratch *s, GenomeData &g, int variant, VerifyArgs &v, uint64_t &n_sites) {
    const Panel &h = s->panel->host;
    const SeedIndex &ix = seed_get(s, variant);
    if (ix.S <= 0) return false;
    ensure_codes(s, g);
    cudaStream_t st = s->stream;
    DevTables *t = s->tables;
    const uint64_t U = h.orients.size();
    const uint64_t n_chunks = (g.lay.total_len + kSeedChunk - 1) / kSeedChunk;
    uint64_t cap = std::max<uint64_t>(1ull << 16, (uint64_t)(ix.cand_prob * (double)g.lay.total_len * 1.5) + 4096);
    cap = std::max(std::min<uint64_t>(cap, 1ull << 28), s->seed_cap_hint);
    uint64_t dirty_cap = std::max<uint64_t>(1ull << 16, s->seed_dirty_hint);
    uint64_t n_cand = 0, n_dirty = 0;
    unsigned long long *ctr = s->counters.as<unsigned long long>();
    for (int attempt = 0;; attempt++) {
        s->cand.ensure(cap * sizeof(unsigned long long));
        s->dirty.ensure(dirty_cap * sizeof(unsigned long long));
        CU(cudaMemsetAsync(s->counters.p, 0, 8 * sizeof(unsigned long long), st));
        CU(cudaEventRecord(s->ev[3], st));
        SeedScanArgs a{};
        a.codes = g.codes.as<uint32_t>();
        a.inval = g.inval.as<uint32_t>();
        a.n_chunks = n_chunks;
        a.S = (uint32_t)ix.S;
        a.d = (uint32_t)(ix.S - ix.S1);
        a.kmer_mask = ix.kmer_mask;
        a.bloom_mul = seed_bloom_mul(ix.S1);
        a.bloom = t->seed_bloom[variant].as<uint32_t>();
        a.bitmap2 = t->seed_bitmap2[variant].as<uint32_t>();
        a.bitmap2_mul = ix.bitmap2_mul;
        a.bitmap2_shift = ix.bitmap2_shift;
        a.table = t->seed_table[variant].as<unsigned long long>();
        a.table_shift = 32u - ix.table_bits;
        a.table_mask = (uint32_t)((1ull << ix.table_bits) - 1);
        a.cand = s->cand.as<unsigned long long>();
        a.cand_count = ctr;
        a.cand_cap = cap;
        a.total_len = g.lay.total_len;
        a.max_mm = (uint32_t)std::max(0, h.cfg.max_mm);
        a.dirty = s->dirty.as<unsigned long long>();
        a.dirty_count = ctr + 2;
        a.dirty_cap = dirty_cap;
        launch_seed_scan(s, ix, a);
        CU(cudaMemcpyAsync(s->h_counts.p, s->counters.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        n_cand = s->h_counts.as<unsigned long long>()[0];
        n_dirty = s->h_counts.as<unsigned long long>()[2];
        if (n_dirty > seed_dirty_max()) return false;
        if (n_cand <= cap && n_dirty <= dirty_cap) break;
        if (attempt >= 2 || n_cand * sizeof(unsigned long long) > (32ull << 30))
            throw Error(fmt("scan produced %llu candidate sites; the panel is too unspecific for this input",
                            (unsigned long long)n_cand));
        cap = std::max(cap, n_cand + 1024);
        dirty_cap = std::max(dirty_cap, n_dirty + 1024);
    }
    s->seed_cap_hint = std::max(s->seed_cap_hint, n_cand + n_cand / 4);
    s->seed_dirty_hint = std::max(s->seed_dirty_hint, n_dirty + n_dirty / 4);
    CU(cudaEventRecord(s->ev[4], st));
    // dirty windows: replacement lookups (k_seed_dirty_probe); the few windows it cannot resolve
    // are expanded against every orientation
    uint64_t n_dcand = 0, n_over = 0;
    if (n_dirty) {
        uint64_t dcap = n_dirty * 4 + 1024;
        s->dirty_over.ensure(n_dirty * sizeof(unsigned long long));
        for (int attempt = 0;; attempt++) {
            s->dcand.ensure(dcap * sizeof(unsigned long long));
            CU(cudaMemsetAsync(ctr + 3, 0, sizeof(unsigned long long), st));
            CU(cudaMemsetAsync(ctr + 5, 0, sizeof(unsigned long long), st));
            SeedDirtyProbeArgs d{};
            d.dirty = s->dirty.as<unsigned long long>();
            d.n_dirty = n_dirty;
            d.codes = g.codes.as<uint32_t>();
            d.inval = g.inval.as<uint32_t>();
            d.S = (uint32_t)ix.S;
            d.kmer_mask = ix.kmer_mask;
            d.bitmap2 = t->seed_bitmap2[variant].as<uint32_t>();
            d.bitmap2_mul = ix.bitmap2_mul;
            d.bitmap2_shift = ix.bitmap2_shift;
            d.table = t->seed_table[variant].as<unsigned long long>();
            d.table_shift = 32u - ix.table_bits;
            d.table_mask = (uint32_t)((1ull << ix.table_bits) - 1);
            d.cand = s->dcand.as<unsigned long long>();
            d.cand_count = ctr + 3;
            d.cand_cap = dcap;
            d.over = s->dirty_over.as<unsigned long long>();
            d.over_count = ctr + 5;
            const unsigned grid = (unsigned)std::min<uint64_t>((n_dirty + 127) / 128, (uint64_t)s->sm_count * 16);
            k_seed_dirty_probe<<<grid, 128, 0, st>>>(d);
            check_launch("k_seed_dirty_probe");
            s->launches++;
            CU(cudaMemcpyAsync(s->h_counts.p, s->counters.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            n_dcand = s->h_counts.as<unsigned long long>()[3];
            n_over = s->h_counts.as<unsigned long long>()[5];
            if (n_dcand <= dcap) break;
            if (attempt >= 1) throw Error("dirty-window candidates overflowed twice");
            dcap = n_dcand + 1024;
        }
        if (n_over * U > seed_dirty_max()) return false;
    }
    const uint64_t n_ocand = n_over * U;
    s->n_cand = n_cand;  // the k-mer hits; dirty-window candidates live in s->dcand / s->ocand
    s->timing.candidates = n_cand + n_dcand + n_ocand;
    n_sites = 0;
    if (n_cand + n_dcand + n_ocand == 0) return true;
    s->sites.ensure((n_cand + n_dcand + n_ocand) * sizeof(ipcr_site));
    v.sites = s->sites.as<ipcr_site>();
    v.site_cap = n_cand + n_dcand + n_ocand;
    auto verify = [&](const DevBuf &buf, uint64_t n) {
        if (!n) return;
        v.cand = buf.as<unsigned long long>();
        v.n_cand = n;
        const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)s->sm_count * 16);
        k_verify<<<grid, 256, 0, st>>>(v);
        check_launch("k_verify");
        s->launches++;
    };
    verify(s->cand, n_cand);
    verify(s->dcand, n_dcand);
    if (n_ocand) {
        s->ocand.ensure(n_ocand * sizeof(unsigned long long));
        SeedExpandArgs e{};
        e.dirty = s->dirty_over.as<unsigned long long>();
        e.n_dirty = n_over;
        e.col_off = t->seed_coloff[variant].as<uint16_t>();
        e.orients = t->orients.as<OrientDev>();
        e.n_orient = (uint32_t)U;
        e.S = (uint32_t)ix.S;
        e.cand = s->ocand.as<unsigned long long>();
        const unsigned grid = (unsigned)std::min<uint64_t>((n_ocand + 255) / 256, (uint64_t)s->sm_count * 16);
        k_seed_dirty_expand<<<grid, 256, 0, st>>>(e);
        check_launch("k_seed_dirty_expand");
        s->launches++;
        verify(s->ocand, n_ocand);
    }
    n_sites = read_counter(s, 1);
    return true;
}

static uint64_t find_sites(ipcr_scratch *s, GenomeData &g, bool force_raw) {
    const Panel &h = s->panel->host;
    cudaStream_t st = s->stream;
    s->n_cand = 0;
    s->timing.used_jit = 0;
    if (g.lay.total_len == 0 || h.orients.empty()) return 0;
    const ipcr_cfg &cfg = h.cfg;
    bool any_rc_unseeded = false;
    for (const Orient &o : h.orients)
        if (o.rc_type && !o.seeded) any_rc_unseeded = true;
    // rc orientations must keep window-violating matches only where the reference caps
    // BEFORE filterLeftTW (compiled.go:185-190,250-257)
    const bool raw_possible = cfg.hit_cap > 0 && cfg.max_mm > 0 && cfg.terminal_window > 0;
    const int variant = raw_possible && (force_raw || g.any_reset || any_rc_unseeded) ? 1 : 0;
    s->timing.plan_variant = (uint32_t)variant;

    s->counters.ensure(8 * sizeof(unsigned long long));
    s->h_counts.ensure(8 * sizeof(unsigned long long));

    VerifyArgs v{};
    v.ascii = g.ascii;
    v.segs = g.segs_dev.as<SegmentDev>();
    v.n_segs = (uint32_t)g.segs.size();
    v.orients = s->tables->orients.as<OrientDev>();
    v.cls = s->tables->cls.as<uint8_t>();
    v.prm.max_mm = cfg.max_mm;
    v.prm.terminal_window = cfg.terminal_window;
    v.prm.hit_cap = cfg.hit_cap;
    v.prm.circular = cfg.circular;
    v.prm.raw_rc = (uint8_t)variant;
    v.site_count = s->counters.as<unsigned long long>() + 1;

    CU(cudaEventRecord(s->ev[3], st));
    uint64_t n_sites = 0;
    if (prefer_seed(s->panel) && seed_filter(s, g, variant, v, n_sites)) {
        s->timing.used_jit = kPathSeed;
        CU(cudaEventRecord(s->ev[5], st));
        s->timing.sites = n_sites;
        return n_sites;
    }
    ensure_planes(s, g);
    JitRun run;
    if (jit_filter(s, g, variant, run)) {
        // ---- specialised filter: per-thread lists of 32-position candidate words ----
        s->timing.used_jit = kPathJit;
        s->n_cand = 0;
        s->timing.candidates = run.n_bits;
        if (run.n_entries > 0 && run.n_bits > 0) {
            s->sites.ensure(run.n_bits * sizeof(ipcr_site));
            v.sites = s->sites.as<ipcr_site>();
            v.site_cap = run.n_bits;
            WordListArgs w{};
            w.lists = s->lists.as<uint4>();
            w.counts = s->list_counts.as<uint32_t>();
            w.clamped = s->list_clamped.as<uint32_t>();
            w.offs = s->list_offs.as<uint32_t>();
            w.n_threads = run.n_threads;
            w.K = run.K;
            w.stats = s->counters.as<unsigned long long>() + 4;
            w.lay = g.lay;
            w.variant = (uint32_t)variant;
            const unsigned grid = (unsigned)std::min<uint64_t>((run.n_entries + 255) / 256, (uint64_t)s->sm_count * 16);
            k_verify_words<<<grid, 256, 0, st>>>(v, w, run.n_entries);
            check_launch("k_verify_words");
            s->launches++;
            n_sites = read_counter(s, 1);
        }
        CU(cudaEventRecord(s->ev[5], st));
        s->timing.sites = n_sites;
        return n_sites;
    }

    // ---- table-driven filter: candidate positions ----
    double expect = 0.0;
    for (const OrientPlan &op : h.plan[variant].o) expect += op.cand_prob;
    uint64_t cap = std::max<uint64_t>(1ull << 16, (uint64_t)(expect * (double)g.lay.total_len * 4.0) + 4096);
    cap = std::min<uint64_t>(cap, 1ull << 28);
    CU(cudaEventRecord(s->ev[3], st));
    uint64_t n_cand = 0;
    for (int attempt = 0;; attempt++) {
        s->cand.ensure(cap * sizeof(unsigned long long));
        CU(cudaMemsetAsync(s->counters.p, 0, 8 * sizeof(unsigned long long), st));
        launch_filter_generic(s, g, variant, cap);
        n_cand = read_counter(s, 0);
        if (n_cand <= cap) break;
        if (attempt >= 2 || n_cand * sizeof(unsigned long long) > (32ull << 30))
            throw Error(fmt("scan produced %llu candidate sites; the panel is too unspecific for this input",
                            (unsigned long long)n_cand));
        cap = n_cand + 1024;
    }
    CU(cudaEventRecord(s->ev[4], st));
    s->n_cand = n_cand;
    s->timing.candidates = n_cand;
    if (n_cand > 0) {
        s->sites.ensure(n_cand * sizeof(ipcr_site));
        v.cand = s->cand.as<unsigned long long>();
        v.n_cand = n_cand;
        v.sites = s->sites.as<ipcr_site>();
        v.site_cap = n_cand;
        const unsigned grid = (unsigned)std::min<uint64_t>((n_cand + 255) / 256, (uint64_t)s->sm_count * 16);
        k_verify<<<grid, 256, 0, st>>>(v);
        check_launch("k_verify");
        s->launches++;
        n_sites = read_counter(s, 1);
    }
    CU(cudaEventRecord(s->ev[5], st));
    s->timing.sites = n_sites;
    return n_sites;
}


// =====================================================================================
// The asynchronous chain (pair_small.cuh).  Stage 1 leaves a SITE BLOCK in device memory,
// every count stays on the device, stage 2 is one single-CTA kernel over any number of
// blocks, and the host waits ONCE for [header | products].  Anything the chain cannot take
// (more than 8192 sites, inputs too dirty for the seed index, ...) is reported in the
// header and served by the multi-kernel path.
// =====================================================================================
constexpr uint64_t kChainMaxEntries = 1ull << 18;  // records x pairs x 2 one CTA walks

static void begin_call(ipcr_scratch *s) {
    bind(s);
    if (s->inflight) CU(cudaStreamSynchronize(s->stream));
    s->inflight = false;
    s->args_used = 0;
    s->launches = 0;
    s->timing = ipcr_timing{};
}

// small per-call tables travel through pinned staging so that the copy never blocks the host
static void upload_staged(ipcr_scratch *s, DevBuf &b, const void *src, size_t bytes) {
    b.ensure(std::max(bytes, (size_t)16));
    if (!bytes) return;
    size_t off = (s->args_used + 63) & ~(size_t)63;
    if (off + bytes > s->h_args.cap) {
        CU(cudaStreamSynchronize(s->stream));  // nothing reads the old staging buffer any more
        s->h_args.ensure(std::max<size_t>(2 * (off + bytes), 1 << 16));
        off = 0;
    }
    memcpy(s->h_args.as<uint8_t>() + off, src, bytes);
    s->args_used = off + bytes;
    CU(cudaMemcpyAsync(b.p, s->h_args.as<uint8_t>() + off, bytes, cudaMemcpyHostToDevice, s->stream));
}

static bool raw_possible(const ipcr_cfg &cfg) { return cfg.hit_cap > 0 && cfg.max_mm > 0 && cfg.terminal_window > 0; }

static int chain_variant(ipcr_scratch *s, bool force_raw) {
    const Panel &h = s->panel->host;
    bool any_rc_unseeded = false;
    for (const Orient &o : h.orients)
        if (o.rc_type && !o.seeded) any_rc_unseeded = true;
    return raw_possible(h.cfg) && (force_raw || s->raw_hint || any_rc_unseeded) ? 1 : 0;
}

// can this panel / record set run the chain at all?
static bool chain_ok(ipcr_scratch *s, const ipcr_record_info *recs, size_t n_recs) {
    const Panel &h = s->panel->host;
    if (getenv("IPCR_NO_CHAIN")) return false;
    if (!prefer_seed(s->panel) || h.roles.empty() || n_recs == 0 || n_recs > 65535) return false;
    if ((uint64_t)n_recs * h.roles.size() * 2ull > kChainMaxEntries) return false;
    if (key_layout(h, recs, n_recs).key_bits + 1 + kSortSmallIdxBits > 64) return false;
    return seed_host(s->panel, 0).S > 0;
}

// Stage 1 of the chain on a laid-out genome: init -> pack -> seed scan -> dirty windows -> verify,
// all into `d_block`.  Nothing is waited for.
static void stage1_async(ipcr_scratch *s, GenomeData &g, int variant, uint8_t *d_block, uint64_t site_cap,
                         uint64_t publish_seq = 0) {
    const Panel &h = s->panel->host;
    const SeedIndex &ix = seed_get(s, variant);
    cudaStream_t st = s->stream;
    DevTables *t = s->tables;
    const size_t n_recs = g.recs.size();
    const bool packed = g.codes_ext != nullptr;
    const uint8_t *rr_init = nullptr;
    {  // ONE small copy: the segment table and, behind it, the record table stage 2 reads (+ record flags)
        const size_t seg_bytes = (g.segs.size() * sizeof(SegmentDev) + 63) & ~(size_t)63;
        const size_t rec_bytes = n_recs * sizeof(ipcr_record_info);
        std::vector<uint8_t> tab(seg_bytes + rec_bytes + (packed ? n_recs : 0), 0);
        memcpy(tab.data(), g.segs.data(), g.segs.size() * sizeof(SegmentDev));
        for (size_t r = 0; r < n_recs; r++) {
            ipcr_record_info ri{};
            ri.len = g.recs[r].len;
            memcpy(tab.data() + seg_bytes + r * sizeof(ipcr_record_info), &ri, sizeof ri);
            if (packed) tab[seg_bytes + rec_bytes + r] = g.recs[r].has_reset_byte ? 1 : 0;
        }
        if (tab != g.last_tab || !g.segs_dev.p) {  // unchanged tables (same shards every call) are already there
            upload_staged(s, g.segs_dev, tab.data(), tab.size());
            g.last_tab = tab;
        }
        g.recs_dev_ptr = reinterpret_cast<const ipcr_record_info *>(g.segs_dev.as<uint8_t>() + seg_bytes);
        if (packed) rr_init = g.segs_dev.as<uint8_t>() + seg_bytes + rec_bytes;
    }
    g.seg_reset.ensure(std::max<size_t>(g.segs.size(), 4) * sizeof(uint32_t));
    uint64_t cap = std::max<uint64_t>(1ull << 16, (uint64_t)(ix.cand_prob * (double)g.lay.total_len * 1.5) + 4096);
    cap = std::max(std::min<uint64_t>(cap, 1ull << 28), s->seed_cap_hint);
    const uint64_t dirty_cap = std::max<uint64_t>(1ull << 16, s->seed_dirty_hint);
    s->cand.ensure(cap * sizeof(unsigned long long));
    s->dirty.ensure(dirty_cap * sizeof(unsigned long long));
    Stage1Ctl init{};
    init.cand_cap = cap;
    init.dirty_cap = dirty_cap;
    init.site_cap = site_cap;
    init.variant = (uint32_t)variant;
    init.n_recs = (uint32_t)n_recs;
    Stage1Ctl *ctl = reinterpret_cast<Stage1Ctl *>(d_block);
    uint8_t *rec_reset = d_block + sizeof(Stage1Ctl);
    k_block_init<<<1, 256, 0, st>>>(d_block, init, g.seg_reset.as<uint32_t>(), (uint32_t)g.segs.size(), rr_init);
    check_launch("k_block_init");
    CU(cudaEventRecord(s->ev[1], st));
    if (!packed) launch_pack_codes(s, g, rec_reset);
    CU(cudaEventRecord(s->ev[2], st));
    const uint32_t *d_codes = packed ? g.codes_ext : g.codes.as<uint32_t>();
    const uint32_t *d_inval = packed ? g.inval_ext : g.inval.as<uint32_t>();
    CU(cudaEventRecord(s->ev[3], st));
    const uint64_t n_chunks = (g.lay.total_len + kSeedChunk - 1) / kSeedChunk;
    SeedScanArgs a{};
    a.codes = d_codes;
    a.inval = d_inval;
    a.n_chunks = n_chunks;
    a.S = (uint32_t)ix.S;
    a.d = (uint32_t)(ix.S - ix.S1);
    a.kmer_mask = ix.kmer_mask;
    a.bloom_mul = seed_bloom_mul(ix.S1);
    a.bloom = t->seed_bloom[variant].as<uint32_t>();
    a.bitmap2 = t->seed_bitmap2[variant].as<uint32_t>();
    a.bitmap2_mul = ix.bitmap2_mul;
    a.bitmap2_shift = ix.bitmap2_shift;
    a.table = t->seed_table[variant].as<unsigned long long>();
    a.table_shift = 32u - ix.table_bits;
    a.table_mask = (uint32_t)((1ull << ix.table_bits) - 1);
    a.cand = s->cand.as<unsigned long long>();
    a.cand_count = &ctl->n_cand;
    a.cand_cap = cap;
    a.total_len = g.lay.total_len;
    a.max_mm = (uint32_t)std::max(0, h.cfg.max_mm);
    a.dirty = s->dirty.as<unsigned long long>();
    a.dirty_count = &ctl->n_dirty;
    a.dirty_cap = dirty_cap;
    launch_seed_scan(s, ix, a);
    CU(cudaEventRecord(s->ev[4], st));
    VerifyArgs v{};
    v.ascii = g.ascii;
    v.segs = g.segs_dev.as<SegmentDev>();
    v.n_segs = (uint32_t)g.segs.size();
    v.orients = t->orients.as<OrientDev>();
    v.cls = t->cls.as<uint8_t>();
    v.prm.max_mm = h.cfg.max_mm;
    v.prm.terminal_window = h.cfg.terminal_window;
    v.prm.hit_cap = h.cfg.hit_cap;
    v.prm.circular = h.cfg.circular;
    v.prm.raw_rc = (uint8_t)variant;
    v.sites = reinterpret_cast<ipcr_site *>(d_block + block_ctl_bytes(n_recs));
    v.site_count = &ctl->n_sites;
    v.site_cap = site_cap;
    if (packed || !getenv("IPCR_VERIFY_ASCII")) {
        v.codes = d_codes;
        v.inval = d_inval;
        v.oplanes = t->oplanes.as<uint64_t>();
    }
    if (h.cfg.max_mm > 0) {  // dirty windows exist only with a mismatch budget (halo.go:76-80)
        SeedDirtyResolveArgs r{};
        r.p.dirty = s->dirty.as<unsigned long long>();
        r.p.codes = a.codes;
        r.p.inval = a.inval;
        r.p.S = a.S;
        r.p.kmer_mask = a.kmer_mask;
        r.p.bitmap2 = a.bitmap2;
        r.p.bitmap2_mul = a.bitmap2_mul;
        r.p.bitmap2_shift = a.bitmap2_shift;
        r.p.table = a.table;
        r.p.table_shift = a.table_shift;
        r.p.table_mask = a.table_mask;
        r.n_dirty_ptr = &ctl->n_dirty;
        r.dirty_cap = dirty_cap;
        r.col_off = t->seed_coloff[variant].as<uint16_t>();
        r.n_orient = (uint32_t)h.orients.size();
        r.v = v;
        k_seed_dirty_resolve<<<(unsigned)s->sm_count * 2, 128, 0, st>>>(r);
        check_launch("k_seed_dirty_resolve");
        s->launches++;
    }
    v.cand = s->cand.as<unsigned long long>();
    v.n_cand = cap;
    v.n_cand_ptr = &ctl->n_cand;
    if (publish_seq) {  // peer exchange: k_verify's last block publishes the finished block
        if (!s->verify_done.p) {
            s->verify_done.ensure(64);
            CU(cudaMemsetAsync(s->verify_done.p, 0, 64, st));
        }
        v.publish_seq_ptr = &ctl->seq;
        v.publish_seq = publish_seq;
        v.blocks_done = s->verify_done.as<unsigned int>();
    }
    k_verify<<<(unsigned)s->sm_count * 8, 256, 0, st>>>(v);
    check_launch("k_verify");
    CU(cudaEventRecord(s->ev[5], st));
    s->launches += 2;  // k_block_init, k_verify
    s->timing.used_jit = kPathSeed;
    s->timing.plan_variant = (uint32_t)variant;
    s->inflight = true;
}

// Stage 2 of the chain over `n_blocks` site blocks + the readback of [header | products].
static void pair_async(ipcr_scratch *s, const uint8_t *d_blocks, uint32_t n_blocks, uint64_t block_bytes,
                       uint64_t block_site_cap, const ipcr_record_info *recs, size_t n_recs,
                       const ipcr_record_info *d_recs = nullptr, const uint8_t *const *d_peer_blocks = nullptr,
                       uint64_t wait_seq = 0) {
    const Panel &h = s->panel->host;
    cudaStream_t st = s->stream;
    const size_t n_pairs = h.roles.size();
    const uint64_t n_entries = (uint64_t)n_recs * n_pairs * 2ull;
    const uint32_t task_cap = (uint32_t)std::max<uint64_t>(1u << 15, s->task_cap_hint);
    const uint64_t prod_cap = std::max<uint64_t>(1u << 12, s->prod_cap_hint);
    if (!d_recs) {
        upload_staged(s, s->recs_dev, recs, n_recs * sizeof(ipcr_record_info));
        d_recs = s->recs_dev.as<ipcr_record_info>();
    }
    s->rec_reset_dev.ensure(n_recs + 16);
    s->fs.ensure((size_t)kPairSmallMax * sizeof(FinalSite));
    s->entry_tasks.ensure(std::max<uint64_t>(n_entries * 4, (uint64_t)kPairMaxProductive * 8));
    s->task_off.ensure((std::max<uint64_t>(n_entries, task_cap) + 1) * 4);  // general mode: entry scan; bucket mode: live tasks
    s->task_cnt.ensure((size_t)task_cap * 4);
    s->prod_off.ensure(((size_t)task_cap + 1) * 4);
    s->products.ensure((prod_cap + 1) * sizeof(ipcr_product));
    s->h_products.ensure((prod_cap + 1) * sizeof(ipcr_product));
    PairSmallArgs a{};
    a.blocks = d_blocks;
    a.peer_blocks = d_peer_blocks;
    a.wait_seq = wait_seq;
    if (d_peer_blocks) {
        s->gather.ensure((size_t)n_blocks * block_bytes + 64);
        a.stage = s->gather.as<uint8_t>();
    }
    a.n_blocks = n_blocks;
    a.block_bytes = block_bytes;
    a.ctl_bytes = block_ctl_bytes(n_recs);
    a.block_site_cap = block_site_cap;
    a.orients = s->tables->orients.as<OrientDev>();
    a.rec_reset = s->rec_reset_dev.as<uint8_t>();
    a.j.fs = s->fs.as<FinalSite>();
    a.j.pairs = s->tables->pairs.as<PairDev>();
    a.j.n_pairs = (uint32_t)n_pairs;
    a.j.n_recs = (uint32_t)n_recs;
    a.j.recs = d_recs;
    a.j.prm.max_mm = h.cfg.max_mm;
    a.j.prm.terminal_window = h.cfg.terminal_window;
    a.j.prm.hit_cap = h.cfg.hit_cap;
    a.j.prm.circular = h.cfg.circular;
    a.j.kl = key_layout(h, recs, n_recs);
    a.j.entry_tasks = s->entry_tasks.as<uint32_t>();
    a.j.task_off = s->task_off.as<uint32_t>();
    a.j.task_cnt = s->task_cnt.as<uint32_t>();
    a.j.prod_off = s->prod_off.as<uint32_t>();
    a.j.products = s->products.as<ipcr_product>() + 1;  // the header occupies the first 64 bytes
    a.j.prod_cap = prod_cap;
    a.task_cap = task_cap;
    a.use_off = s->tables->use_off.as<uint32_t>();
    a.use_list = s->tables->use_list.as<uint32_t>();
    a.header = s->products.as<PairHeader>();
    // shared memory: 8192 sort words + either the final keys (general mode) or the two bucket tables
    constexpr size_t kPairSmemMax = 222 * 1024;  // + ~1.5 KB of static shared memory in the kernel: 227 KB per CTA at most
    const size_t words = (size_t)kPairSmallMax * sizeof(unsigned long long);
    const uint64_t n_buckets = (uint64_t)n_recs * h.orients.size();
    size_t smem = 2 * words;
    if (words + 8 * (n_buckets + 1) <= kPairSmemMax && !getenv("IPCR_NO_BUCKETS")) {
        a.bucket_cap = (uint32_t)n_buckets;
        smem = std::max(smem, words + 8 * ((size_t)n_buckets + 1));
        // + the panel tables of the join (uses, pairs, record lengths) when they still fit
        const size_t tables = 4 * ((h.orients.size() + 1) + 2 * n_pairs + 8 * n_pairs + 4 * n_recs + 16);
        if (words + 8 * ((size_t)n_buckets + 1) + tables <= kPairSmemMax) smem = std::max(smem, words + 8 * ((size_t)n_buckets + 1) + tables);
    }
    a.smem_words = (uint32_t)(smem / 4);
    if (getenv("IPCR_CHAIN_DEBUG")) {
        s->pair_dbg.ensure(32 * 8);
        CU(cudaMemsetAsync(s->pair_dbg.p, 0, 32 * 8, st));
        a.dbg = s->pair_dbg.as<unsigned long long>();
    }
    static std::mutex attr_mu;
    {
        std::lock_guard<std::mutex> lock(attr_mu);
        CU(cudaFuncSetAttribute(k_pair_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPairSmemMax));
    }
    k_pair_small<<<1, 1024, smem, st>>>(a);
    check_launch("k_pair_small");
    s->launches++;
    CU(cudaEventRecord(s->ev[6], st));
    // one readback: the header and as many products as the last calls needed (+ slack)
    const uint64_t n_copy = std::min<uint64_t>(prod_cap, std::max<uint64_t>(256, s->prod_cap_hint));
    CU(cudaMemcpyAsync(s->h_products.p, s->products.p, (n_copy + 1) * sizeof(ipcr_product), cudaMemcpyDeviceToHost, st));
    s->last_copied = n_copy;
    CU(cudaEventRecord(s->ev[7], st));
    s->inflight = true;
}

enum ChainResult { kChainDone, kChainRetry, kChainTooDirty };

// Waits for the chain and reads the header.  kChainDone: *n_prod products are at s->out_ptr.
// kChainRetry: a capacity (or the plan variant) was adjusted; run stage 1 again if *redo_stage1,
// else only pair_async.  Larger site lists are paired by the multi-kernel path right here.
static ChainResult chain_finish(ipcr_scratch *s, const uint8_t *d_blocks, uint32_t n_blocks, uint64_t block_bytes,
                                ipcr_record_info *recs, size_t n_recs, uint64_t *n_prod, bool *redo_stage1,
                                const uint8_t *const *d_peer_blocks = nullptr) {
    const Panel &h = s->panel->host;
    cudaStream_t st = s->stream;
    CU(cudaStreamSynchronize(st));
    s->inflight = false;
    const PairHeader hd0 = *s->h_products.as<PairHeader>();
    struct { uint64_t n_sites, n_final, n_tasks, n_prod, n_cand, n_dirty; uint32_t flags, min_variant; } hd{
        hd0.n_sites, hd0.n_final, hd0.n_tasks, hd0.n_prod, hd0.n_cand, hd0.n_dirty, hd0.flags, hd0.min_variant};
    *redo_stage1 = false;
    *n_prod = 0;
    if (getenv("IPCR_CHAIN_DEBUG") && s->pair_dbg.p) {
        unsigned long long d[32];
        if (cudaMemcpy(d, s->pair_dbg.p, sizeof d, cudaMemcpyDeviceToHost) == cudaSuccess) {
            fprintf(stderr, "[chain] k_pair_small phase ends (cycles of thread 0: zeroed, keys, max, scan, sorted, keep, scan, final, tables, joined, filled):");
            for (unsigned long long i = 1; i <= d[0] && i < 32; i++) fprintf(stderr, " %llu", d[i]);
            fprintf(stderr, "\n");
        }
    }
    if (getenv("IPCR_CHAIN_DEBUG"))
        fprintf(stderr, "[chain] sites %llu final %llu tasks %llu prod %llu cand %llu dirty %llu flags %x variant %u\n",
                (unsigned long long)hd.n_sites, (unsigned long long)hd.n_final, (unsigned long long)hd.n_tasks,
                (unsigned long long)hd.n_prod, (unsigned long long)hd.n_cand, (unsigned long long)hd.n_dirty, hd.flags,
                hd.min_variant);
    s->timing.candidates = hd.n_cand;
    s->timing.sites = hd.n_sites;
    if (hd.flags & kPairPeerTimeout)
        throw Error("peer exchange: a rank did not publish its site block within 4 s (did every rank run stage 1 of this step?)");
    if (hd.n_dirty > seed_dirty_max()) return kChainTooDirty;
    if (hd.flags & kPairStage1Overflow) {
        s->seed_cap_hint = std::max(s->seed_cap_hint, hd.n_cand + hd.n_cand / 4 + 1024);
        s->seed_dirty_hint = std::max(s->seed_dirty_hint, hd.n_dirty + hd.n_dirty / 4 + 1024);
        if (hd.n_cand * sizeof(unsigned long long) > (32ull << 30))
            throw Error(fmt("scan produced %llu candidate sites; the panel is too unspecific for this input",
                            (unsigned long long)hd.n_cand));
        *redo_stage1 = true;
        return kChainRetry;
    }
    s->seed_cap_hint = std::max(s->seed_cap_hint, hd.n_cand + hd.n_cand / 4);
    s->seed_dirty_hint = std::max(s->seed_dirty_hint, hd.n_dirty + hd.n_dirty / 4);
    if ((hd.flags & kPairAnyReset) && raw_possible(h.cfg)) {
        s->raw_hint = true;  // later calls start with the raw-rc index
        if (hd.min_variant == 0) {  // hit-cap parity needs raw rc matches from every block
            *redo_stage1 = true;
            return kChainRetry;
        }
    }
    if (hd.flags & kPairSitesLost) {
        s->site_cap_hint = std::max<uint64_t>(s->site_cap_hint, hd.n_sites + hd.n_sites / 4 + 1024);
        *redo_stage1 = true;
        return kChainRetry;
    }
    if (hd.flags & kPairNotSmall) {  // complete site lists, too long for one CTA: multi-kernel stage 2
        if (hd.n_sites >= (1ull << 32)) throw Error("too many primer sites for one pairing pass");
        const uint64_t ctl_bytes = block_ctl_bytes(n_recs);
        const ipcr_site *d_sites = reinterpret_cast<const ipcr_site *>(d_blocks + ctl_bytes);
        if (n_blocks > 1 || d_peer_blocks) {
            s->sites.ensure(hd.n_sites * sizeof(ipcr_site));
            UnpackArgs u{d_blocks, d_peer_blocks, n_blocks, block_bytes, ctl_bytes, s->sites.as<ipcr_site>(), hd.n_sites};
            k_unpack_blocks<<<(unsigned)s->sm_count, 256, 0, st>>>(u);
            check_launch("k_unpack_blocks");
            s->launches++;
            d_sites = s->sites.as<ipcr_site>();
        }
        s->h_reset.ensure(n_recs + 16);
        CU(cudaMemcpyAsync(s->h_reset.p, s->rec_reset_dev.p, n_recs, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        for (size_t r = 0; r < n_recs; r++) recs[r].has_reset_byte = s->h_reset.as<uint8_t>()[r] ? 1 : 0;
        *n_prod = pair_sites(s, d_sites, hd.n_sites, recs, n_recs);
        if (*n_prod == 0) CU(cudaEventRecord(s->ev[6], st));
        CU(cudaEventRecord(s->ev[7], st));
        CU(cudaStreamSynchronize(st));
        return kChainDone;
    }
    if (hd.flags & (kPairTaskOverflow | kPairProdOverflow)) {
        s->task_cap_hint = std::max<uint64_t>(s->task_cap_hint, hd.n_tasks + hd.n_tasks / 4 + 1024);
        s->prod_cap_hint = std::max<uint64_t>(s->prod_cap_hint, hd.n_prod + hd.n_prod / 4 + 1024);
        if (hd.n_tasks >= (1ull << 31) || hd.n_prod >= (1ull << 31)) throw Error("too many products for one pairing pass");
        return kChainRetry;
    }
    const uint64_t copied = s->last_copied;
    if (hd.n_prod > copied) {  // the rest of the products
        CU(cudaMemcpyAsync(s->h_products.as<ipcr_product>() + 1 + copied, s->products.as<ipcr_product>() + 1 + copied,
                           (hd.n_prod - copied) * sizeof(ipcr_product), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    s->prod_cap_hint = std::max<uint64_t>(s->prod_cap_hint, hd.n_prod + hd.n_prod / 4);
    s->task_cap_hint = std::max<uint64_t>(s->task_cap_hint, hd.n_tasks + hd.n_tasks / 4);
    s->out_ptr = s->h_products.as<ipcr_product>() + 1;
    *n_prod = hd.n_prod;
    return kChainDone;
}

static void chain_timing(ipcr_scratch *s, uint64_t n_prod, bool had_h2d) {
    ipcr_timing &t = s->timing;
    t.h2d_ms = had_h2d ? ev_ms(s->ev[0], s->ev[1]) : 0.f;
    t.pack_ms = ev_ms(s->ev[1], s->ev[2]);
    t.filter_ms = ev_ms(s->ev[3], s->ev[4]);
    t.verify_ms = ev_ms(s->ev[4], s->ev[5]);
    t.pair_ms = ev_ms(s->ev[5], s->ev[6]);
    t.d2h_ms = ev_ms(s->ev[6], s->ev[7]);
    t.total_ms = ev_ms(had_h2d ? s->ev[0] : s->ev[1], s->ev[7]);
    t.products = n_prod;
    t.kernel_launches = s->launches;
}

// One laid-out genome through the chain.  false: the input is too dirty for the seed index.
static bool scan_chain(ipcr_scratch *s, GenomeData &g, bool had_h2d, uint64_t *n_prod) {
    const size_t n_recs = g.recs.size();
    for (int attempt = 0; attempt < 6; attempt++) {
        const uint64_t site_cap = std::max<uint64_t>(kPairSmallMax, s->site_cap_hint);
        const uint64_t bb = block_bytes_for(n_recs, site_cap);
        s->block.ensure(bb);
        stage1_async(s, g, chain_variant(s, false), s->block.as<uint8_t>(), site_cap);
        for (int again = 0; again < 4; again++) {
            pair_async(s, s->block.as<uint8_t>(), 1, bb, site_cap, g.recs.data(), n_recs, g.recs_dev_ptr);
            bool redo = false;
            const ChainResult r = chain_finish(s, s->block.as<uint8_t>(), 1, bb, g.recs.data(), n_recs, n_prod, &redo);
            if (r == kChainTooDirty) return false;
            if (r == kChainDone) {
                chain_timing(s, *n_prod, had_h2d);
                return true;
            }
            if (redo) break;
        }
    }
    throw Error("the asynchronous scan chain did not converge");
}

// a laid-out genome (ASCII resident or on its way) -> products: the chain when it applies,
// else pack (synchronous) -> find -> pair
static uint64_t scan_laidout(ipcr_scratch *s, GenomeData &g, bool had_h2d) {
    if (!s->inflight) s->args_used = 0;
    if (g.lay.total_len == 0) {
        s->out_ptr = s->h_products.as<ipcr_product>();
        return 0;
    }
    if (chain_ok(s, g.recs.data(), g.recs.size())) {
        uint64_t n = 0;
        if (scan_chain(s, g, had_h2d, &n)) return n;
    }
    pack_genome(s, g);
    CU(cudaEventRecord(s->ev[2], s->stream));
    return scan_genome_impl(s, g, true);
}
}  // namespace ipcr

#include "multi.hpp"

using namespace ipcr;

#define ABI_TRY try {
#define ABI_CATCH(ret)                                   \
    }                                                    \
    catch (const std::exception &e) {                    \
        set_err(err, errlen, e.what());                  \
        return ret;                                      \
    }                                                    \
    catch (...) {                                        \
        set_err(err, errlen, "unknown error");           \
        return ret;                                      \
    }

extern "C" {

int ipcr_gpu_abi_version(void) { return IPCR_GPU_ABI_VERSION; }

int ipcr_gpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int ipcr_gpu_init(int n_devices_or_0, char *err, size_t errlen) {
    ABI_TRY
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        throw Error("no CUDA device visible: the ipcr GPU engine has no CPU fallback");
    }
    if (n_devices_or_0 < 0 || n_devices_or_0 > n)
        throw Error(fmt("%d devices requested, %d visible", n_devices_or_0, n));
    const int use = n_devices_or_0 == 0 ? n : n_devices_or_0;
    int cur = 0;
    CU(cudaGetDevice(&cur));
    check_device(cur);
    for (int d = 0; d < n && use > 1; d++) check_device(d);
    CU(cudaFree(0));
    g_n_devices.store(use);
    return 0;
    ABI_CATCH(1)
}

int ipcr_gpu_set_device(int device, char *err, size_t errlen) {
    ABI_TRY
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        throw Error("no CUDA device visible: the ipcr GPU engine has no CPU fallback");
    }
    if (device < 0 || device >= n) throw Error(fmt("device %d out of range (0..%d)", device, n - 1));
    CU(cudaSetDevice(device));
    check_device(device);
    CU(cudaFree(0));
    return 0;
    ABI_CATCH(1)
}

int ipcr_gpu_num_devices(void) { return g_n_devices.load(); }

ipcr_panel *ipcr_gpu_panel_create(const ipcr_cfg *cfg, const ipcr_pair *pairs, size_t n_pairs, char *err,
                                  size_t errlen) {
    ABI_TRY
    if (!cfg) throw Error("cfg is NULL");
    if (n_pairs && !pairs) throw Error("pairs is NULL");
    auto p = std::make_unique<ipcr_panel>();
    std::string m = p->host.build(*cfg, pairs, n_pairs);
    if (!m.empty()) throw Error(m);
    return p.release();
    ABI_CATCH(nullptr)
}

void ipcr_gpu_panel_destroy(ipcr_panel *p) { delete p; }
size_t ipcr_gpu_panel_num_pairs(const ipcr_panel *p) { return p ? p->host.roles.size() : 0; }
size_t ipcr_gpu_panel_num_orientations(const ipcr_panel *p) { return p ? p->host.orients.size() : 0; }

int ipcr_gpu_panel_have(const ipcr_panel *p, size_t pair_idx, int which) {
    if (!p || pair_idx >= p->host.have.size()) return 0;
    int bit = which == 'A' ? 0 : which == 'B' ? 1 : which == 'a' ? 2 : which == 'b' ? 3 : -1;
    return bit < 0 ? 0 : (p->host.have[pair_idx] >> bit) & 1;
}

void ipcr_gpu_panel_set_jit(ipcr_panel *p, int mode) {
    if (p) p->host.jit_mode = mode;
}

void ipcr_gpu_panel_set_filter(ipcr_panel *p, int filter) { ipcr_gpu_panel_set_jit(p, filter); }

ipcr_scratch *ipcr_gpu_scratch_create(ipcr_panel *p, char *err, size_t errlen) {
    ABI_TRY
    if (!p) throw Error("panel is NULL");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        throw Error("no CUDA device visible: the ipcr GPU engine has no CPU fallback");
    }
    int device = 0;
    CU(cudaGetDevice(&device));
    return make_scratch(p, device);
    ABI_CATCH(nullptr)
}

void ipcr_gpu_scratch_destroy(ipcr_scratch *s) {
    if (!s) return;
    for (ipcr_scratch *q : s->peers) {
        cudaSetDevice(q->device);
        delete q;
    }
    s->peers.clear();
    cudaSetDevice(s->device);
    for (void *m : s->peer_map)
        if (m) cudaIpcCloseMemHandle(m);
    delete s;
}

void *ipcr_gpu_scratch_stream(ipcr_scratch *s) { return s ? (void *)s->stream : nullptr; }

struct ipcr_fasta {
    std::vector<ipcr::FastaRecord> recs;
    uint8_t *seq = nullptr;
    uint64_t total = 0;
    bool pinned = false;
    // packed form of the whole sequence buffer (IPCR_FASTA_PACK): codes + hard-mismatch bitmap, pinned
    uint32_t *codes = nullptr, *inval = nullptr;
    bool packed_pinned = false;  // page-locking costs ~0.5 s per GB: only worth it for handles scanned many times
    // the bitmap as a run list (a real genome's N runs are few): shipped instead of the 0.125 B/bp bitmap
    std::vector<uint64_t> inval_runs;
    bool runs_ok = false;
    // per-record reset flags: from the run list when there is one (O(runs log records)), else from the bitmap
    void find_rec_reset() {
        rec_reset.assign(recs.size(), 0);
        if (!runs_ok) {
            for (size_t r = 0; r < recs.size(); r++)
                rec_reset[r] = ipcr::packed_any_invalid(inval, recs[r].off, recs[r].off + recs[r].len) ? 1 : 0;
            return;
        }
        for (size_t i = 0; i + 1 < inval_runs.size(); i += 2) {
            const uint64_t a = inval_runs[i], b = std::min<uint64_t>(inval_runs[i + 1], total);
            if (a >= b) continue;  // the padding behind the last record
            size_t lo = 0, hi = recs.size();  // first record whose end lies beyond a
            while (lo < hi) {
                const size_t mid = (lo + hi) / 2;
                if (recs[mid].off + recs[mid].len > a) hi = mid; else lo = mid + 1;
            }
            for (size_t r = lo; r < recs.size() && recs[r].off < b; r++)
                if (recs[r].len) rec_reset[r] = 1;
        }
    }
    void find_runs(int threads, const std::vector<std::pair<uint64_t, uint64_t>> *dirty = nullptr) {
        const uint64_t groups = ipcr::packed_groups(total);
        const uint64_t max_runs = std::max<uint64_t>(4096, groups / 64);
        uint64_t covered = 0;
        if (dirty)
            for (const auto &r : *dirty) covered += r.second - r.first;
        if (dirty && covered <= groups / 8)  // the feeder knows where the set bits are: look only there
            runs_ok = ipcr::packed_invalid_runs_ranges(inval, *dirty, max_runs, inval_runs);
        else
            runs_ok = ipcr::packed_invalid_runs_mt(inval, groups, max_runs, inval_runs, threads);
        if (!runs_ok) inval_runs.clear();
    }
    std::vector<uint8_t> rec_reset;  // per record: holds a non-ACGT byte (the feeder upper-cased a-z already)
    std::vector<ipcr_product> products;
    // IPCR_FASTA_PACK_ONLY: no ASCII was kept.  The paths that need it (ipcr_gpu_fasta_record's seq, the scan of a
    // handle the packed chain cannot take) rebuild it once from the codes, non-ACGT bytes as 'N'.
    std::mutex ascii_mu;
    const uint8_t *ascii(int threads = 0) {
        std::lock_guard<std::mutex> lock(ascii_mu);
        if (seq || !codes) return seq;
        uint8_t *buf = static_cast<uint8_t *>(malloc((size_t)std::max<uint64_t>(total, 1) + 64));
        if (!buf) throw std::runtime_error("out of host memory for the FASTA records");
        if (threads <= 0) threads = (int)std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
        if (total < (1u << 22)) threads = 1;
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++) {
            const uint64_t a = total * (uint64_t)t / (uint64_t)threads, b = total * (uint64_t)(t + 1) / (uint64_t)threads;
            if (t + 1 < threads) th.emplace_back([=] { ipcr::unpack_host(codes, inval, a, b, buf + a); });
            else ipcr::unpack_host(codes, inval, a, b, buf + a);
        }
        for (auto &x : th) x.join();
        seq = buf;
        return seq;
    }
    void alloc_packed(uint64_t total, bool pin) {
        const uint64_t groups = ipcr::packed_groups(total);
        if (pin) {
            void *c = nullptr, *i = nullptr;
            if (cudaMallocHost(&c, groups * 8 + 64) == cudaSuccess && cudaMallocHost(&i, groups * 4 + 64) == cudaSuccess) {
                codes = static_cast<uint32_t *>(c);
                inval = static_cast<uint32_t *>(i);
                packed_pinned = true;
                return;
            }
            cudaGetLastError();
            if (c) cudaFreeHost(c);
        }
        codes = static_cast<uint32_t *>(malloc(groups * 8 + 64));
        inval = static_cast<uint32_t *>(malloc(groups * 4 + 64));
        if (!codes || !inval) throw std::runtime_error("out of host memory for the packed records");
    }
    ~ipcr_fasta() {
        if (seq) {
            if (pinned) cudaFreeHost(seq);
            else free(seq);
        }
        if (packed_pinned) {
            if (codes) cudaFreeHost(codes);
            if (inval) cudaFreeHost(inval);
        } else {
            free(codes);
            free(inval);
        }
    }
};

// the feeder's second output: the packed form of its sequence buffer, in pinned memory
static void fasta_pack(ipcr_fasta *f, int threads, bool pin) {
    f->alloc_packed(f->total, pin);
    pack_host(f->seq, f->total, f->codes, f->inval, threads);
    f->find_runs(threads);
    f->find_rec_reset();
}

// host records -> products in s->h_products (shared by scan_batch and scan_fasta)
static uint64_t scan_records(ipcr_scratch *s, const ipcr_seq *recs, size_t n_recs) {
    std::vector<ipcr_segment> segs(n_recs);
    for (size_t i = 0; i < n_recs; i++) {
        segs[i].seq = recs[i].seq;
        segs[i].len = recs[i].len;
        segs[i].own_len = recs[i].len;
        segs[i].rec_off = 0;
        segs[i].rec_idx = (uint32_t)i;
        segs[i].reserved = 0;
    }
    uint64_t total = 0;
    for (size_t i = 0; i < n_recs; i++) total += recs[i].len;
    if (!s->inflight) s->args_used = 0;
    std::vector<ipcr_record_info> rinfo(n_recs, ipcr_record_info{});
    for (size_t i = 0; i < n_recs; i++) rinfo[i].len = recs[i].len;
    const bool chain = total < (1ull << kCandPosBits) && chain_ok(s, rinfo.data(), n_recs);
    uint64_t n = 0;
    // several devices: one slice each (every device pulls its slice over its own PCIe link)
    if (chain && g_n_devices.load() > 1 && total >= multi_min_bases() && scan_multi_device(s, recs, n_recs, &n)) return n;
    // large inputs: overlap the PCIe copy with the scan, slab by slab
    int n_slabs = (int)std::min<uint64_t>(5, total / (640ull << 20));
    // pageable records are worth streaming much earlier: the staged copy (HostStager) is what makes them fast
    if (n_slabs < 2 && chain && total >= staged_min_bases() && n_recs && host_is_pageable(recs[0].seq)) n_slabs = 2;
    if (const char *e = getenv("IPCR_STREAM_SLABS")) n_slabs = atoi(e);
    if (n_slabs >= 2 && total < (1ull << kCandPosBits) && n_recs <= 65535) {
        if (chain && n_slabs <= (int)kPairMaxBlocks && scan_streamed_chain(s, recs, n_recs, n_slabs, &n)) return n;
        n = scan_batch_streamed(s, recs, n_recs, n_slabs);
    } else {
        layout_genome(s, s->own, segs.data(), n_recs, n_recs, false);
        n = scan_laidout(s, s->own, true);
    }
    return n;
}

int ipcr_gpu_scan_batch(ipcr_panel *p, ipcr_scratch *s, const ipcr_seq *recs, size_t n_recs,
                        const ipcr_product **out, uint64_t *n_out, char *err, size_t errlen) {
    ABI_TRY
    if (!p || !s || s->panel != p) throw Error("scratch does not belong to this panel");
    if (out) *out = nullptr;
    if (n_out) *n_out = 0;
    begin_call(s);
    const uint64_t n = scan_records(s, recs, n_recs);
    if (out) *out = s->out_ptr;
    if (n_out) *n_out = n;
    return 0;
    ABI_CATCH(1)
}

ipcr_fasta *ipcr_gpu_fasta_open(const char *path, int threads, int pin, char *err, size_t errlen) {
    ABI_TRY
    if (!path) throw Error("path is NULL");
    auto f = std::make_unique<ipcr_fasta>();
    FastaBytes fb;
    fasta_read_file(path, fb);
    if (threads <= 0) threads = (int)std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    ipcr_fasta *fp = f.get();
    const bool pack_only = (pin & IPCR_FASTA_PACK_ONLY) != 0;
    const bool want_pack = pack_only || (pin & IPCR_FASTA_PACK) != 0;
    pin &= IPCR_FASTA_PIN;
    FastaPackTarget pk;
    pk.only = pack_only;
    const auto t_begin = std::chrono::steady_clock::now();
    double alloc_s = 0.0;
    fasta_parse(fb.data, fb.size, threads, f->recs, f->total, [&](uint64_t total) -> uint8_t * {
        const auto ta = std::chrono::steady_clock::now();
        struct Stop {
            double &acc; std::chrono::steady_clock::time_point t0;
            ~Stop() { acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
        } stop{alloc_s, ta};
        if (want_pack) {  // the packed form is written by the copy pass itself, while its input is still in cache
            fp->alloc_packed(total, pin != 0);
            pk.codes = fp->codes;
            pk.inval = fp->inval;
        }
        if (pack_only) return nullptr;  // no ASCII buffer: pass 2 packs from per-thread staging
        const size_t bytes = (size_t)std::max<uint64_t>(total, 1) + 64;
        if (pin) {
            void *q = nullptr;
            if (cudaMallocHost(&q, bytes) == cudaSuccess) {
                fp->pinned = true;
                fp->seq = static_cast<uint8_t *>(q);
                return fp->seq;
            }
            cudaGetLastError();
        }
        fp->seq = static_cast<uint8_t *>(malloc(bytes));
        if (!fp->seq) throw Error("out of host memory for the FASTA records");
        return fp->seq;
    }, &pk);
    if (getenv("IPCR_FEEDER_TIMING"))
        fprintf(stderr, "[feeder] %zu bytes, %d threads: allocations %.3f s (packed: %s, page-locked: %s), parse + copy + pack %.3f s\n",
                fb.size, threads, alloc_s, want_pack ? "yes" : "no", pin ? "yes" : "no",
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count() - alloc_s);
    const auto t_parsed = std::chrono::steady_clock::now();
    if (want_pack) {
        f->find_runs(threads, pk.dirty_known ? &pk.dirty : nullptr);
        f->find_rec_reset();
    }
    const auto t_runs = std::chrono::steady_clock::now();
    if (fb.map) {  // unmapping 3 GB of touched pages takes ~20 ms: not on the caller's clock
        void *m = fb.map;
        const size_t len = fb.map_len;
        fb.map = nullptr;
        std::thread([m, len] { munmap(m, len); }).detach();
    }
    if (getenv("IPCR_FEEDER_TIMING"))
        fprintf(stderr, "[feeder] run list + record flags %.3f s (%zu runs)\n",
                std::chrono::duration<double>(t_runs - t_parsed).count(), f->inval_runs.size() / 2);
    return f.release();
    ABI_CATCH(nullptr)
}

ipcr_fasta *ipcr_gpu_fasta_from_memory(const ipcr_seq *recs, const char *const *ids, size_t n_recs, int threads,
                                       int flags, char *err, size_t errlen) {
    ABI_TRY
    if (n_recs && !recs) throw Error("recs is NULL");
    auto f = std::make_unique<ipcr_fasta>();
    if (threads <= 0) threads = (int)std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    uint64_t total = 0;
    for (size_t r = 0; r < n_recs; r++) {
        FastaRecord fr;
        fr.id = ids && ids[r] ? ids[r] : std::to_string(r);
        fr.off = total;
        fr.len = recs[r].len;
        total += recs[r].len;
        f->recs.push_back(std::move(fr));
    }
    f->total = total;
    const size_t bytes = (size_t)std::max<uint64_t>(total, 1) + 64;
    if (flags & IPCR_FASTA_PIN) {
        void *q = nullptr;
        if (cudaMallocHost(&q, bytes) == cudaSuccess) {
            f->pinned = true;
            f->seq = static_cast<uint8_t *>(q);
        } else {
            cudaGetLastError();
        }
    }
    if (!f->seq) f->seq = static_cast<uint8_t *>(malloc(bytes));
    if (!f->seq) throw Error("out of host memory for the records");
    {  // the reader's normalisation (a-z upper-cased, normalize.go:5-14), one thread per slice of records
        std::vector<std::thread> th;
        std::atomic<size_t> next{0};
        auto work = [&] {
            for (size_t r = next++; r < n_recs; r = next++) {
                const uint8_t *src = recs[r].seq;
                uint8_t *dst = f->seq + f->recs[r].off;
                for (uint64_t i = 0; i < recs[r].len; i++) {
                    const uint8_t c = src[i];
                    dst[i] = (uint8_t)(c - (((uint8_t)(c - 'a') < 26u) ? 32u : 0u));
                }
            }
        };
        for (int t = 1; t < threads && (size_t)t < n_recs; t++) th.emplace_back(work);
        work();
        for (auto &x : th) x.join();
    }
    if (flags & IPCR_FASTA_PACK) fasta_pack(f.get(), threads, (flags & IPCR_FASTA_PIN) != 0);
    return f.release();
    ABI_CATCH(nullptr)
}

int ipcr_gpu_fasta_is_packed(const ipcr_fasta *f) { return f && f->codes ? 1 : 0; }

int ipcr_gpu_fasta_runs(const ipcr_fasta *f, const uint64_t **runs, uint64_t *n_runs) {
    if (runs) *runs = nullptr;
    if (n_runs) *n_runs = 0;
    if (!f || !f->codes || !f->runs_ok) return 1;
    if (runs) *runs = f->inval_runs.data();
    if (n_runs) *n_runs = f->inval_runs.size() / 2;
    return 0;
}

int ipcr_gpu_fasta_packed(const ipcr_fasta *f, const uint32_t **codes, const uint32_t **inval, uint64_t *n_groups) {
    if (!f || !f->codes) return 1;
    if (codes) *codes = f->codes;
    if (inval) *inval = f->inval;
    if (n_groups) *n_groups = packed_groups(f->total);
    return 0;
}

void ipcr_gpu_packed_words(uint64_t n_bases, uint64_t *code_words, uint64_t *inval_words) {
    const uint64_t g = packed_groups(n_bases);
    if (code_words) *code_words = 2 * g;
    if (inval_words) *inval_words = g;
}

int64_t ipcr_gpu_packed_runs(const uint32_t *inval, uint64_t n_bases, uint64_t *runs, uint64_t max_runs, int threads) {
    if (!inval || (!runs && max_runs)) return -1;
    if (threads <= 0) threads = (int)std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    std::vector<uint64_t> r;
    if (!packed_invalid_runs_mt(inval, packed_groups(n_bases), max_runs, r, threads)) return -1;
    memcpy(runs, r.data(), r.size() * sizeof(uint64_t));
    return (int64_t)(r.size() / 2);
}

int ipcr_gpu_inval_from_runs(ipcr_scratch *s, uint32_t *d_inval, uint64_t n_bases, const uint64_t *runs, uint64_t n_runs,
                             char *err, size_t errlen) {
    ABI_TRY
    if (!s || !d_inval || (!runs && n_runs)) throw Error("scratch / bitmap / runs is NULL");
    bind(s);
    if (s->inflight) CU(cudaStreamSynchronize(s->stream));  // the staging buffer below may still be in use
    s->inflight = false;
    s->args_used = 0;
    const uint64_t groups = packed_groups(n_bases);
    for (uint64_t i = 0; i < n_runs; i++)
        if (runs[2 * i] > runs[2 * i + 1] || runs[2 * i + 1] > groups * 32ull) throw Error("run outside the packed buffer");
    CU(cudaMemsetAsync(d_inval, 0, (size_t)groups * 4, s->stream));
    if (n_runs) {
        upload_staged(s, s->runs_dev, runs, n_runs * 16);
        const unsigned grid = (unsigned)std::min<uint64_t>((n_runs * 32 + 255) / 256, (uint64_t)s->sm_count * 8);
        k_inval_fill<<<grid, 256, 0, s->stream>>>(s->runs_dev.as<unsigned long long>(), n_runs, d_inval);
        check_launch("k_inval_fill");
    }
    s->inflight = true;
    return 0;
    ABI_CATCH(1)
}

int ipcr_gpu_pack_host(const uint8_t *ascii, uint64_t n_bases, uint32_t *codes, uint32_t *inval, int threads) {
    if ((!ascii && n_bases) || !codes || !inval) return 1;
    if (threads <= 0) threads = (int)std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    pack_host(ascii, n_bases, codes, inval, threads);
    return 0;
}

void ipcr_gpu_fasta_close(ipcr_fasta *f) { delete f; }
size_t ipcr_gpu_fasta_num_records(const ipcr_fasta *f) { return f ? f->recs.size() : 0; }
uint64_t ipcr_gpu_fasta_total_bases(const ipcr_fasta *f) { return f ? f->total : 0; }

int ipcr_gpu_fasta_record(const ipcr_fasta *f, size_t i, const char **id, const uint8_t **seq, uint64_t *len) {
    if (!f || i >= f->recs.size()) return 1;
    if (id) *id = f->recs[i].id.c_str();
    if (seq) {
        try {
            *seq = const_cast<ipcr_fasta *>(f)->ascii() + f->recs[i].off;  // pack-only handles rebuild it on first use
        } catch (const std::exception &) {
            return 1;
        }
    }
    if (len) *len = f->recs[i].len;
    return 0;
}

int ipcr_gpu_scan_fasta(ipcr_panel *p, ipcr_scratch *s, ipcr_fasta *f, const ipcr_product **out, uint64_t *n_out,
                        char *err, size_t errlen) {
    ABI_TRY
    if (!p || !s || !f || s->panel != p) throw Error("bad panel/scratch/fasta");
    if (out) *out = nullptr;
    if (n_out) *n_out = 0;
    begin_call(s);
    f->products.clear();
    // the packed form goes over PCIe when the whole file fits one launch chain
    if (f->codes && f->recs.size() <= 60000 && f->total < (1ull << 43) && f->total > 0 && !getenv("IPCR_NO_PACKED")) {
        std::vector<ipcr_record_info> rinfo(f->recs.size(), ipcr_record_info{});
        std::vector<uint64_t> roff(f->recs.size() + 1, 0);
        for (size_t r = 0; r < f->recs.size(); r++) {
            rinfo[r].len = f->recs[r].len;
            rinfo[r].has_reset_byte = f->rec_reset[r];
            roff[r] = f->recs[r].off;
        }
        roff[f->recs.size()] = f->total;
        if (!s->inflight) s->args_used = 0;
        if (chain_ok(s, rinfo.data(), rinfo.size())) {
            PackedHost ph{f->codes, f->inval, f->total, f->runs_ok && !getenv("IPCR_NO_RUNS") ? f->inval_runs.data() : nullptr,
                          f->inval_runs.size() / 2, f->packed_pinned};
            uint64_t n = 0;
            // ipcr_gpu_init(n > 1): every device takes a slice of the packed arrays over its own PCIe link
            const bool multi = g_n_devices.load() > 1 && f->total >= multi_min_bases();
            if ((multi && scan_packed_multi(s, ph, rinfo, roff, &n)) || scan_packed_chain(s, ph, rinfo, roff, &n)) {
                f->products.assign(s->out_ptr, s->out_ptr + n);
                if (out) *out = f->products.data();
                if (n_out) *n_out = n;
                return 0;
            }
        }
    }
    ipcr_timing acc{};
    // batches: <= 60000 records and < 2^43 bases per launch chain (sort-key / candidate-word budgets)
    size_t i = 0;
    while (i < f->recs.size()) {
        std::vector<ipcr_seq> batch;
        uint64_t bases = 0;
        const size_t base = i;
        while (i < f->recs.size() && batch.size() < 60000 && bases + f->recs[i].len < (1ull << 43)) {
            batch.push_back(ipcr_seq{f->ascii() + f->recs[i].off, f->recs[i].len});
            bases += f->recs[i].len;
            i++;
        }
        if (batch.empty()) throw Error("a single FASTA record is too large for one device pass");
        const uint64_t n = scan_records(s, batch.data(), batch.size());
        const ipcr_product *pr = s->out_ptr;
        for (uint64_t k = 0; k < n; k++) {
            ipcr_product q = pr[k];
            q.rec_idx += (uint32_t)base;
            f->products.push_back(q);
        }
        const ipcr_timing &t = s->timing;
        acc.h2d_ms += t.h2d_ms; acc.pack_ms += t.pack_ms; acc.filter_ms += t.filter_ms; acc.verify_ms += t.verify_ms;
        acc.pair_ms += t.pair_ms; acc.d2h_ms += t.d2h_ms; acc.total_ms += t.total_ms;
        acc.candidates += t.candidates; acc.sites += t.sites; acc.used_jit |= t.used_jit;
        acc.plan_variant |= t.plan_variant;
    }
    acc.products = f->products.size();
    acc.kernel_launches = s->launches;
    s->timing = acc;
    if (out) *out = f->products.data();
    if (n_out) *n_out = f->products.size();
    return 0;
    ABI_CATCH(1)
}

int ipcr_gpu_scan_device(ipcr_panel *p, ipcr_scratch *s, const ipcr_segment *segs, size_t n_segs, size_t n_recs,
                         const ipcr_product **out, uint64_t *n_out, char *err, size_t errlen) {
    ABI_TRY
    if (!p || !s || s->panel != p) throw Error("scratch does not belong to this panel");
    if (out) *out = nullptr;
    if (n_out) *n_out = 0;
    begin_call(s);
    CU(cudaEventRecord(s->ev[0], s->stream));
    layout_genome(s, s->own, segs, n_segs, n_recs, true);
    const uint64_t n = scan_laidout(s, s->own, true);
    if (out) *out = s->out_ptr;
    if (n_out) *n_out = n;
    return 0;
    ABI_CATCH(1)
}

int ipcr_gpu_scan(ipcr_panel *p, ipcr_scratch *s, const uint8_t *seq, uint64_t len, const ipcr_product **out,
                  uint64_t *n_out, char *err, size_t errlen) {
    ipcr_seq r{seq, len};
    return ipcr_gpu_scan_batch(p, s, &r, 1, out, n_out, err, errlen);
}

static ipcr_genome *genome_make(ipcr_scratch *s, const ipcr_segment *segs, size_t n_segs, size_t n_recs, bool dev,
                                char *err, size_t errlen) {
    ABI_TRY
    if (!s) throw Error("scratch is NULL");
    bind(s);
    auto g = std::make_unique<ipcr_genome>();
    build_genome(s, g->g, segs, n_segs, n_recs, dev);
    CU(cudaStreamSynchronize(s->stream));
    return g.release();
    ABI_CATCH(nullptr)
}

ipcr_genome *ipcr_gpu_genome_upload(ipcr_scratch *s, const ipcr_segment *segs, size_t n_segs, size_t n_recs,
                                    char *err, size_t errlen) {
    return genome_make(s, segs, n_segs, n_recs, false, err, errlen);
}

ipcr_genome *ipcr_gpu_genome_adopt(ipcr_scratch *s, const ipcr_segment *segs, size_t n_segs, size_t n_recs,
                                   char *err, size_t errlen) {
    return genome_make(s, segs, n_segs, n_recs, true, err, errlen);
}

void ipcr_gpu_genome_destroy(ipcr_genome *g) {
    if (!g) return;
    cudaSetDevice(g->g.device);
    delete g;
}

uint64_t ipcr_gpu_genome_bases(const ipcr_genome *g) { return g ? g->g.own_bases : 0; }

int ipcr_gpu_genome_records(const ipcr_genome *g, ipcr_record_info *out, size_t n_recs) {
    if (!g || !out) return 1;
    for (size_t i = 0; i < n_recs && i < g->g.recs.size(); i++) out[i] = g->g.recs[i];
    return 0;
}

int ipcr_gpu_scan_genome(ipcr_panel *p, ipcr_scratch *s, const ipcr_genome *g, const ipcr_product **out,
                         uint64_t *n_out, char *err, size_t errlen) {
    ABI_TRY
    if (!p || !s || !g || s->panel != p) throw Error("bad panel/scratch/genome");
    if (g->g.device != s->device) throw Error("genome lives on another device than the scratch");
    if (out) *out = nullptr;
    if (n_out) *n_out = 0;
    begin_call(s);
    // the handle is logically const: packing the other format on demand is cached state (pack_mu)
    const uint64_t n = scan_genome_impl(s, const_cast<GenomeData &>(g->g), false);
    if (out) *out = s->out_ptr;
    if (n_out) *n_out = n;
    return 0;
    ABI_CATCH(1)
}

int ipcr_gpu_find_sites(ipcr_panel *p, ipcr_scratch *s, const ipcr_genome *g, const ipcr_site **sites,
                        uint64_t *n_sites, char *err, size_t errlen) {
    ABI_TRY
    if (!p || !s || !g || s->panel != p) throw Error("bad panel/scratch/genome");
    if (g->g.device != s->device) throw Error("genome lives on another device than the scratch");
    begin_call(s);
    for (int i = 0; i < 8; i++) CU(cudaEventRecord(s->ev[i], s->stream));
    const uint64_t n = find_sites(s, const_cast<GenomeData &>(g->g), true);
    if (n) {
        s->h_sites.ensure(n * sizeof(ipcr_site));
        CU(cudaMemcpyAsync(s->h_sites.p, s->sites.p, n * sizeof(ipcr_site), cudaMemcpyDeviceToHost, s->stream));
    }
    CU(cudaEventRecord(s->ev[7], s->stream));
    CU(cudaStreamSynchronize(s->stream));
    s->timing.filter_ms = ev_ms(s->ev[3], s->ev[4]);
    s->timing.verify_ms = ev_ms(s->ev[4], s->ev[5]);
    s->timing.total_ms = ev_ms(s->ev[3], s->ev[7]);
    s->timing.kernel_launches = s->launches;
    if (sites) *sites = n ? s->h_sites.as<ipcr_site>() : nullptr;
    if (n_sites) *n_sites = n;
    return 0;
    ABI_CATCH(1)
}

int ipcr_gpu_find_sites_device(ipcr_panel *p, ipcr_scratch *s, const ipcr_segment *segs, size_t n_segs,
                               size_t n_recs, int force_raw, const ipcr_site **sites, uint64_t *n_sites,
                               ipcr_record_info *recs_out, char *err, size_t errlen) {
    ABI_TRY
    if (!p || !s || s->panel != p) throw Error("bad panel/scratch");
    begin_call(s);
    for (int i = 0; i < 8; i++) CU(cudaEventRecord(s->ev[i], s->stream));
    build_genome(s, s->own, segs, n_segs, n_recs, true);
    const uint64_t n = find_sites(s, s->own, force_raw != 0);
    if (n) {
        s->h_sites.ensure(n * sizeof(ipcr_site));
        CU(cudaMemcpyAsync(s->h_sites.p, s->sites.p, n * sizeof(ipcr_site), cudaMemcpyDeviceToHost, s->stream));
    }
    CU(cudaEventRecord(s->ev[7], s->stream));
    CU(cudaStreamSynchronize(s->stream));
    s->timing.pack_ms = ev_ms(s->ev[1], s->ev[2]);
    s->timing.filter_ms = ev_ms(s->ev[3], s->ev[4]);
    s->timing.verify_ms = ev_ms(s->ev[4], s->ev[5]);
    s->timing.total_ms = ev_ms(s->ev[0], s->ev[7]);
    s->timing.kernel_launches = s->launches;
    if (recs_out)
        for (size_t i = 0; i < n_recs && i < s->own.recs.size(); i++) recs_out[i] = s->own.recs[i];
    if (sites) *sites = n ? s->h_sites.as<ipcr_site>() : nullptr;
    if (n_sites) *n_sites = n;
    return 0;
    ABI_CATCH(1)
}

int ipcr_gpu_pair_sites(ipcr_panel *p, ipcr_scratch *s, const ipcr_site *sites, uint64_t n_sites,
                        const ipcr_record_info *recs, size_t n_recs, const ipcr_product **out, uint64_t *n_out,
                        char *err, size_t errlen) {
    ABI_TRY
    if (!p || !s || s->panel != p) throw Error("bad panel/scratch");
    if (out) *out = nullptr;
    if (n_out) *n_out = 0;
    bind(s);
    s->launches = 0;
    const size_t n_orient = p->host.orients.size();
    for (uint64_t i = 0; i < n_sites; i++) {
        if (sites[i].rec_idx >= n_recs) throw Error(fmt("site %llu: rec_idx out of range", (unsigned long long)i));
        if (sites[i].orient >= n_orient) throw Error(fmt("site %llu: orientation out of range", (unsigned long long)i));
        // a position outside its record would spill into the neighbouring bits of the sort key
        if (sites[i].pos + (uint64_t)p->host.orients[sites[i].orient].L > recs[sites[i].rec_idx].len)
            throw Error(fmt("site %llu: position %llu + primer length exceeds its record", (unsigned long long)i,
                            (unsigned long long)sites[i].pos));
    }
    CU(cudaEventRecord(s->ev[5], s->stream));
    if (n_sites) {
        s->sites.ensure(n_sites * sizeof(ipcr_site));
        CU(cudaMemcpyAsync(s->sites.p, sites, n_sites * sizeof(ipcr_site), cudaMemcpyHostToDevice, s->stream));
    }
    const uint64_t n = pair_sites(s, s->sites.as<ipcr_site>(), n_sites, recs, n_recs);
    CU(cudaEventRecord(s->ev[7], s->stream));
    CU(cudaStreamSynchronize(s->stream));
    s->timing.pair_ms = ev_ms(s->ev[5], s->ev[7]);
    s->timing.products = n;
    s->timing.kernel_launches = s->launches;
    if (out) *out = s->out_ptr;
    if (n_out) *n_out = n;
    return 0;
    ABI_CATCH(1)
}

uint64_t ipcr_gpu_site_block_bytes(size_t n_recs, uint64_t site_cap) { return block_bytes_for(n_recs, site_cap); }

uint64_t ipcr_gpu_scratch_site_cap(const ipcr_scratch *s) {
    return std::max<uint64_t>(1024, s ? s->site_cap_hint : 0);
}

int ipcr_gpu_find_sites_block(ipcr_panel *p, ipcr_scratch *s, const ipcr_segment *segs, size_t n_segs, size_t n_recs,
                              int force_raw, void *d_block, uint64_t site_cap, char *err, size_t errlen) {
    ABI_TRY
    if (!p || !s || s->panel != p || !d_block) throw Error("bad panel/scratch/block");
    begin_call(s);
    layout_genome(s, s->own, segs, n_segs, n_recs, true);
    s->last_valid = false;
    std::vector<ipcr_record_info> rinfo(n_recs, ipcr_record_info{});
    for (size_t r = 0; r < n_recs; r++) rinfo[r].len = std::max<uint64_t>(1, s->own.recs[r].len);
    if (!chain_ok(s, rinfo.data(), n_recs))
        throw Error("ipcr_gpu_find_sites_block needs a seed-index panel; use ipcr_gpu_find_sites_device");
    if (s->own.lay.total_len == 0) {
        Stage1Ctl init{};
        init.site_cap = site_cap;
        init.variant = (uint32_t)chain_variant(s, force_raw != 0);
        init.n_recs = (uint32_t)n_recs;
        s->seg_reset_dummy.ensure(16);
        for (int i = 1; i <= 5; i++) CU(cudaEventRecord(s->ev[i], s->stream));
        k_block_init<<<1, 256, 0, s->stream>>>(static_cast<uint8_t *>(d_block), init, s->seg_reset_dummy.as<uint32_t>(), 0, nullptr);
        check_launch("k_block_init");
        s->inflight = true;
    } else {
        stage1_async(s, s->own, chain_variant(s, force_raw != 0), static_cast<uint8_t *>(d_block), site_cap, s->pending_publish);
        s->pending_publish = 0;  // k_verify's last block publishes: ipcr_gpu_peer_publish has nothing left to do
    }
    s->timing.kernel_launches = s->launches;
    return 0;
    ABI_CATCH(1)
}

int ipcr_gpu_find_sites_block_packed(ipcr_panel *p, ipcr_scratch *s, const uint32_t *d_codes, const uint32_t *d_inval,
                                     uint64_t n_bases, const ipcr_packed_segment *segs, size_t n_segs, size_t n_recs,
                                     int force_raw, void *d_block, uint64_t site_cap, char *err, size_t errlen) {
    ABI_TRY
    if (!p || !s || s->panel != p || !d_block || !d_codes || !d_inval) throw Error("bad panel/scratch/block/packed buffers");
    begin_call(s);
    GenomeData &g = s->own;
    g.device = s->device;
    g.segs.clear();
    g.recs.assign(n_recs, ipcr_record_info{});
    g.own_bases = 0;
    for (size_t i = 0; i < n_segs; i++) {
        const ipcr_packed_segment &in = segs[i];
        if (in.rec_idx >= n_recs) throw Error(fmt("segment %zu: rec_idx %u out of range", i, in.rec_idx));
        if (in.own_len > in.len || in.off + in.len > n_bases) throw Error(fmt("segment %zu out of range", i));
        if (i > 0 && in.off < segs[i - 1].off + segs[i - 1].len) throw Error("packed segments must be ascending and non-overlapping");
        SegmentDev d{};
        d.global_off = in.off;
        d.len = in.len;
        d.own_len = in.own_len;
        d.rec_off = in.rec_off;
        d.rec_idx = in.rec_idx;
        g.segs.push_back(d);
        g.recs[in.rec_idx].len = std::max<uint64_t>(g.recs[in.rec_idx].len, in.rec_off + in.len);
        g.recs[in.rec_idx].has_reset_byte |= in.has_reset_byte ? 1 : 0;
        g.own_bases += in.own_len;
    }
    g.lay.total_len = n_bases;
    g.ascii = nullptr;
    g.codes_ext = d_codes;
    g.inval_ext = d_inval;
    s->last_valid = false;
    std::vector<ipcr_record_info> rinfo(n_recs, ipcr_record_info{});
    for (size_t r = 0; r < n_recs; r++) rinfo[r].len = std::max<uint64_t>(1, g.recs[r].len);
    if (!chain_ok(s, rinfo.data(), n_recs)) throw Error("ipcr_gpu_find_sites_block_packed needs a seed-index panel");
    if (n_bases >= (1ull << kCandPosBits)) throw Error("genome too large for one device buffer");
    if (n_segs == 0 || n_bases == 0) {
        Stage1Ctl init{};
        init.site_cap = site_cap;
        init.variant = (uint32_t)chain_variant(s, force_raw != 0);
        init.n_recs = (uint32_t)n_recs;
        s->seg_reset_dummy.ensure(16);
        for (int i = 1; i <= 5; i++) CU(cudaEventRecord(s->ev[i], s->stream));
        k_block_init<<<1, 256, 0, s->stream>>>(static_cast<uint8_t *>(d_block), init, s->seg_reset_dummy.as<uint32_t>(), 0, nullptr);
        check_launch("k_block_init");
        s->inflight = true;
    } else {
        stage1_async(s, g, chain_variant(s, force_raw != 0), static_cast<uint8_t *>(d_block), site_cap, s->pending_publish);
        s->pending_publish = 0;
    }
    g.codes_ext = g.inval_ext = nullptr;  // the scratch's own genome goes back to packing for itself
    s->timing.kernel_launches = s->launches;
    return 0;
    ABI_CATCH(1)
}

int ipcr_gpu_pair_blocks(ipcr_panel *p, ipcr_scratch *s, const void *d_blocks, uint32_t n_blocks, uint64_t site_cap,
                         const uint64_t *rec_lens, size_t n_recs, const ipcr_product **out, uint64_t *n_out,
                         char *err, size_t errlen) {
    ABI_TRY
    if (!p || !s || s->panel != p || !d_blocks || (!rec_lens && n_recs)) throw Error("bad panel/scratch/blocks");
    if (n_blocks == 0 || n_blocks > kPairMaxBlocks) throw Error(fmt("1..%u blocks per pairing pass", kPairMaxBlocks));
    if (out) *out = nullptr;
    if (n_out) *n_out = 0;
    bind(s);  // no begin_call: the blocks may still be on their way on this stream
    std::vector<ipcr_record_info> rinfo(n_recs, ipcr_record_info{});
    for (size_t r = 0; r < n_recs; r++) rinfo[r].len = rec_lens[r];
    if (!chain_ok(s, rinfo.data(), n_recs)) throw Error("ipcr_gpu_pair_blocks needs a seed-index panel");
    const uint64_t bb = block_bytes_for(n_recs, site_cap);
    const uint8_t *blocks = static_cast<const uint8_t *>(d_blocks);
    for (int again = 0; again < 4; again++) {
        pair_async(s, blocks, n_blocks, bb, site_cap, rinfo.data(), n_recs);
        bool redo = false;
        uint64_t n = 0;
        const ChainResult r = chain_finish(s, blocks, n_blocks, bb, rinfo.data(), n_recs, &n, &redo);
        if (r == kChainTooDirty)
            throw Error("input too dirty for the seed index: use the host-gathered split phase (ipcr_gpu_find_sites_device)");
        if (r == kChainDone) {
            chain_timing(s, n, false);
            if (out) *out = s->out_ptr;
            if (n_out) *n_out = n;
            return 0;
        }
        if (redo) return 2;
    }
    throw Error("stage 2 did not converge");
    ABI_CATCH(1)
}

// ---- peer exchange (one process per GPU; the collective is fused into the pairing kernel) ----
static void peer_close(ipcr_scratch *s) {
    for (void *m : s->peer_map)
        if (m) cudaIpcCloseMemHandle(m);
    s->peer_map.clear();
    s->peer_n = 0;
}

int ipcr_gpu_peer_create(ipcr_scratch *s, size_t n_recs, uint64_t site_cap, void *handle_out, char *err, size_t errlen) {
    ABI_TRY
    if (!s || !handle_out) throw Error("scratch / handle is NULL");
    static_assert(sizeof(cudaIpcMemHandle_t) == IPCR_GPU_PEER_HANDLE_BYTES, "IPC handle size");
    begin_call(s);
    peer_close(s);
    s->peer_recs = n_recs;
    s->peer_cap = site_cap;
    s->peer_bb = block_bytes_for(n_recs, site_cap);
    s->peer_step = 0;
    s->peer_buf.release();  // a fresh allocation: an exported one cannot be resized under its importers
    s->peer_buf.ensure(2 * s->peer_bb);
    CU(cudaMemsetAsync(s->peer_buf.p, 0, 2 * s->peer_bb, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, s->peer_buf.p));
    memcpy(handle_out, &h, sizeof h);
    return 0;
    ABI_CATCH(1)
}

int ipcr_gpu_peer_open(ipcr_scratch *s, const void *handles, uint32_t n_ranks, uint32_t my_rank, char *err, size_t errlen) {
    ABI_TRY
    if (!s || !handles || !s->peer_buf.p) throw Error("ipcr_gpu_peer_create first");
    if (n_ranks == 0 || n_ranks > kPairMaxBlocks || my_rank >= n_ranks) throw Error(fmt("1..%u ranks", kPairMaxBlocks));
    bind(s);
    peer_close(s);
    s->peer_map.assign(n_ranks, nullptr);
    std::vector<const uint8_t *> tab[2];
    for (uint32_t r = 0; r < n_ranks; r++) {
        const uint8_t *base = s->peer_buf.as<uint8_t>();
        if (r != my_rank) {
            cudaIpcMemHandle_t h;
            memcpy(&h, static_cast<const uint8_t *>(handles) + (size_t)r * sizeof h, sizeof h);
            void *m = nullptr;
            CU(cudaIpcOpenMemHandle(&m, h, cudaIpcMemLazyEnablePeerAccess));
            s->peer_map[r] = m;
            base = static_cast<const uint8_t *>(m);
        }
        tab[0].push_back(base);
        tab[1].push_back(base + s->peer_bb);
    }
    for (int k = 0; k < 2; k++) upload(s->peer_tab[k], tab[k].data(), tab[k].size() * sizeof(void *), s->stream);
    CU(cudaStreamSynchronize(s->stream));
    s->peer_n = n_ranks;
    s->peer_rank = my_rank;
    s->peer_step = 0;
    return 0;
    ABI_CATCH(1)
}

void *ipcr_gpu_peer_begin_step(ipcr_scratch *s) {
    if (!s || !s->peer_n) return nullptr;
    s->peer_step++;
    s->pending_publish = s->peer_step;
    return s->peer_buf.as<uint8_t>() + (s->peer_step & 1ull) * s->peer_bb;
}

int ipcr_gpu_peer_publish(ipcr_scratch *s, char *err, size_t errlen) {
    ABI_TRY
    if (!s || !s->peer_n || s->peer_step == 0) throw Error("no peer step in progress");
    bind(s);
    if (s->pending_publish) {  // stage 1 did not publish from inside its verify kernel (empty shard): do it here
        k_block_publish<<<1, 1, 0, s->stream>>>(s->peer_buf.as<uint8_t>() + (s->peer_step & 1ull) * s->peer_bb, s->peer_step);
        check_launch("k_block_publish");
        s->launches++;
        s->pending_publish = 0;
    }
    s->inflight = true;
    return 0;
    ABI_CATCH(1)
}

int ipcr_gpu_pair_peers(ipcr_panel *p, ipcr_scratch *s, const uint64_t *rec_lens, size_t n_recs, const ipcr_product **out,
                        uint64_t *n_out, char *err, size_t errlen) {
    ABI_TRY
    if (!p || !s || s->panel != p || (!rec_lens && n_recs)) throw Error("bad panel/scratch");
    if (!s->peer_n || s->peer_step == 0) throw Error("no peer step in progress");
    if (n_recs != s->peer_recs) throw Error("record count differs from ipcr_gpu_peer_create");
    if (out) *out = nullptr;
    if (n_out) *n_out = 0;
    bind(s);
    std::vector<ipcr_record_info> rinfo(n_recs, ipcr_record_info{});
    for (size_t r = 0; r < n_recs; r++) rinfo[r].len = rec_lens[r];
    if (!chain_ok(s, rinfo.data(), n_recs)) throw Error("ipcr_gpu_pair_peers needs a seed-index panel");
    const uint8_t *const *tab = s->peer_tab[s->peer_step & 1ull].as<const uint8_t *>();
    const uint8_t *own = s->peer_buf.as<uint8_t>() + (s->peer_step & 1ull) * s->peer_bb;
    for (int again = 0; again < 4; again++) {
        pair_async(s, own, s->peer_n, s->peer_bb, s->peer_cap, rinfo.data(), n_recs, nullptr, tab, s->peer_step);
        bool redo = false;
        uint64_t n = 0;
        const ChainResult r = chain_finish(s, own, s->peer_n, s->peer_bb, rinfo.data(), n_recs, &n, &redo, tab);
        if (r == kChainTooDirty)
            throw Error("input too dirty for the seed index: use the host-gathered split phase (ipcr_gpu_find_sites_device)");
        if (r == kChainDone) {
            chain_timing(s, n, false);
            if (out) *out = s->out_ptr;
            if (n_out) *n_out = n;
            return 0;
        }
        if (redo) return 2;
    }
    throw Error("stage 2 did not converge");
    ABI_CATCH(1)
}

void ipcr_gpu_peer_close(ipcr_scratch *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    peer_close(s);
}

int ipcr_gpu_sort_products(ipcr_scratch *s, ipcr_product *products, uint64_t n, const uint32_t *rec_rank,
                           const int64_t *rec_off, size_t n_recs, const uint32_t *pair_rank, size_t n_pairs,
                           uint8_t *tie_out, char *err, size_t errlen) {
    ABI_TRY
    if (!s) throw Error("scratch is NULL");
    if (n == 0) return 0;
    if (!products || !rec_rank || !pair_rank) throw Error("products / rank tables are NULL");
    if (n >= (1ull << 32)) throw Error("too many products for one sort");
    uint32_t max_rr = 0, max_pr = 0;
    for (size_t r = 0; r < n_recs; r++) max_rr = std::max(max_rr, rec_rank[r]);
    for (size_t q = 0; q < n_pairs; q++) max_pr = std::max(max_pr, pair_rank[q]);
    if (max_rr >= (1u << 20) || max_pr >= (1u << 19)) throw Error("rank tables exceed the sort key (2^20 sequences, 2^19 experiments)");
    for (uint64_t i = 0; i < n; i++) {
        const ipcr_product &p = products[i];
        if (p.rec_idx >= n_recs || p.pair_idx >= n_pairs) throw Error(fmt("product %llu: index out of range", (unsigned long long)i));
        const long long off = rec_off ? rec_off[p.rec_idx] : 0;
        if (p.start + off < 0 || p.end + off < 0 || p.length < 0 || ((uint64_t)(p.start + off) >> 44) || ((uint64_t)(p.end + off) >> 44) ||
            ((uint64_t)p.length >> 44) || p.fwd_mm > 255 || p.rev_mm > 255 || p.type > 1)
            throw Error(fmt("product %llu does not fit the sort key", (unsigned long long)i));
    }
    begin_call(s);
    cudaStream_t st = s->stream;
    s->sp_in.ensure(n * sizeof(ipcr_product));
    s->sp_out.ensure(n * sizeof(ipcr_product));
    for (auto &b : s->sp_k) b.ensure(n * 8);
    s->sp_ka.ensure(n * 8);
    s->sp_kb.ensure(n * 8);
    s->sp_ia.ensure(n * 4);
    s->sp_ib.ensure(n * 4);
    s->sp_tie.ensure(n);
    s->tmp.ensure(std::max(sort_tmp_words(n), scan_tmp_words(n)) * 4);
    CU(cudaMemcpyAsync(s->sp_in.p, products, n * sizeof(ipcr_product), cudaMemcpyHostToDevice, st));
    upload(s->sp_rr, rec_rank, n_recs * 4, st);
    upload(s->sp_pr, pair_rank, n_pairs * 4, st);
    if (rec_off) upload(s->sp_ro, rec_off, n_recs * 8, st);
    ProductKeyArgs ka{};
    ka.products = s->sp_in.as<ipcr_product>();
    ka.n = n;
    ka.rec_rank = s->sp_rr.as<uint32_t>();
    ka.rec_off = rec_off ? s->sp_ro.as<long long>() : nullptr;
    ka.pair_rank = s->sp_pr.as<uint32_t>();
    ka.k1 = s->sp_k[0].as<uint64_t>();
    ka.k2 = s->sp_k[1].as<uint64_t>();
    ka.k3 = s->sp_k[2].as<uint64_t>();
    ka.k4 = s->sp_k[3].as<uint64_t>();
    ka.idx = s->sp_ia.as<uint32_t>();
    const unsigned grid = (unsigned)((n + 255) / 256);
    k_product_keys<<<grid, 256, 0, st>>>(ka);
    check_launch("k_product_keys");
    s->launches++;
    // least significant key first; every pass is a stable sort of (key[idx], idx)
    const uint32_t bits[4] = {44u + bit_length(max_rr), 44u, 64u, 16u};
    uint32_t *idx = s->sp_ia.as<uint32_t>(), *idx_alt = s->sp_ib.as<uint32_t>();
    for (int k = 3; k >= 0; k--) {
        k_gather_u64<<<grid, 256, 0, st>>>(s->sp_k[k].as<uint64_t>(), idx, n, s->sp_ka.as<uint64_t>());
        check_launch("k_gather_u64");
        s->launches++;
        const int where = radix_sort(st, s->sp_ka.as<uint64_t>(), idx, s->sp_kb.as<uint64_t>(), idx_alt, n, bits[k],
                                     s->tmp.as<uint32_t>(), &s->launches);
        check_launch("radix_sort");
        if (where) std::swap(idx, idx_alt);
    }
    k_gather_products<<<grid, 256, 0, st>>>(s->sp_in.as<ipcr_product>(), idx, n, s->sp_out.as<ipcr_product>(),
                                           s->sp_tie.as<uint8_t>(), ka.k1, ka.k2, ka.k3, ka.k4);
    check_launch("k_gather_products");
    s->launches++;
    CU(cudaMemcpyAsync(products, s->sp_out.p, n * sizeof(ipcr_product), cudaMemcpyDeviceToHost, st));
    std::vector<uint8_t> tie_local;
    uint8_t *tie = tie_out;
    if (!tie) {
        tie_local.resize(n);
        tie = tie_local.data();
    }
    CU(cudaMemcpyAsync(tie, s->sp_tie.p, n, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    s->timing.kernel_launches = s->launches;
    return 0;
    ABI_CATCH(1)
}

// uploads the product list and the record table of the last scan; validates the products
static void visitor_setup(ipcr_scratch *s, const ipcr_product *products, uint64_t n, AmpliconArgs &a) {
    if (!s->last_valid || !s->last_ascii)
        throw Error("no resident records: run ipcr_gpu_scan / _scan_batch / _scan_device on this scratch first");
    const size_t nr = s->last_rec_len.size();
    for (uint64_t i = 0; i < n; i++) {
        const ipcr_product &p = products[i];
        if (p.rec_idx >= nr) throw Error(fmt("product %llu: rec_idx out of range", (unsigned long long)i));
        const int64_t rl = (int64_t)s->last_rec_len[p.rec_idx];
        if (p.start < 0 || p.start >= rl || p.length < 0 || p.length > rl || p.end < 0 || p.end > rl)
            throw Error(fmt("product %llu does not lie inside its record", (unsigned long long)i));
    }
    cudaStream_t st = s->stream;
    upload(s->v_rec_base, s->last_rec_base.data(), nr * 8, st);
    upload(s->v_rec_len, s->last_rec_len.data(), nr * 8, st);
    upload(s->v_products, products, n * sizeof(ipcr_product), st);
    a.ascii = s->last_ascii;
    a.rec_base = s->v_rec_base.as<uint64_t>();
    a.rec_len = s->v_rec_len.as<uint64_t>();
    a.products = s->v_products.as<ipcr_product>();
    a.n_products = n;
}

int ipcr_gpu_probe_annotate(ipcr_scratch *s, const ipcr_product *products, uint64_t n, const char *probe,
                            uint32_t probe_len, int32_t max_mm, ipcr_probe_hit *out, char *err, size_t errlen) {
    ABI_TRY
    if (!s) throw Error("scratch is NULL");
    if (n == 0) return 0;
    if (!products || !out) throw Error("products/out is NULL");
    bind(s);
    std::string pr(probe ? probe : "", probe_len), rc;
    if (pr.empty()) {  // oligo.go:21-23: empty probe => no hit
        for (uint64_t i = 0; i < n; i++) out[i] = ipcr_probe_hit{};
        return 0;
    }
    size_t bad = 0;
    if (!revcomp(pr, rc, bad))
        throw Error(fmt("invalid probe base at position %zu; allowed: A C G T R Y S W K M B D H V N", bad));
    if (max_mm < 0) max_mm = 0;
    std::vector<uint8_t> cls(2 * pr.size());
    bool strict = true;
    for (size_t i = 0; i < pr.size(); i++) {
        cls[i] = iupac_class((uint8_t)pr[i]);
        cls[pr.size() + i] = iupac_class((uint8_t)rc[i]);
        if (__builtin_popcount(cls[i]) != 1) strict = false;
    }
    ProbeArgs a{};
    visitor_setup(s, products, n, a.amp);
    cudaStream_t st = s->stream;
    upload(s->v_cls, cls.data(), cls.size(), st);
    s->v_hits.ensure(n * sizeof(ipcr_probe_hit));
    a.cls_plus = s->v_cls.as<uint8_t>();
    a.cls_minus = s->v_cls.as<uint8_t>() + pr.size();
    a.L = (uint32_t)pr.size();
    a.max_mm = max_mm;
    a.exact_strict = (max_mm == 0 && strict) ? 1u : 0u;
    a.out = s->v_hits.as<ipcr_probe_hit>();
    k_probe_annotate<<<(unsigned)n, 128, 0, st>>>(a);
    check_launch("k_probe_annotate");
    s->launches++;
    CU(cudaMemcpyAsync(out, s->v_hits.p, n * sizeof(ipcr_probe_hit), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
    ABI_CATCH(1)
}

int ipcr_gpu_amplicons(ipcr_scratch *s, const ipcr_product *products, uint64_t n, uint8_t *out, uint64_t cap,
                       uint64_t *offsets, uint64_t *need, char *err, size_t errlen) {
    ABI_TRY
    if (!s) throw Error("scratch is NULL");
    std::vector<uint64_t> offs(n + 1, 0);
    for (uint64_t i = 0; i < n; i++) offs[i + 1] = offs[i] + (uint64_t)std::max<int64_t>(0, products[i].length);
    if (need) *need = offs[n];
    if (offsets) memcpy(offsets, offs.data(), (n + 1) * 8);
    if (!out || n == 0 || offs[n] == 0) return 0;
    if (cap < offs[n]) throw Error("amplicon buffer too small");
    bind(s);
    AmpliconArgs a{};
    visitor_setup(s, products, n, a);
    cudaStream_t st = s->stream;
    upload(s->v_offs, offs.data(), (n + 1) * 8, st);
    s->v_amp.ensure(offs[n]);
    k_gather_amplicons<<<(unsigned)n, 128, 0, st>>>(a, s->v_offs.as<uint64_t>(), s->v_amp.as<uint8_t>());
    check_launch("k_gather_amplicons");
    s->launches++;
    CU(cudaMemcpyAsync(out, s->v_amp.p, offs[n], cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
    ABI_CATCH(1)
}

int ipcr_gpu_nested(ipcr_scratch *outer, const ipcr_product *products, uint64_t n, ipcr_panel *inner_panel,
                    ipcr_scratch *inner, const char *const *inner_ids, ipcr_product *best, char *err, size_t errlen) {
    ABI_TRY
    if (!outer || !inner || !inner_panel || inner->panel != inner_panel) throw Error("bad scratch/panel");
    if (n == 0) return 0;
    if (!products || !best) throw Error("products/best is NULL");
    // 1. all amplicons, gathered on the device that still holds the outer records
    std::vector<uint64_t> offs(n + 1);
    uint64_t need = 0;
    char e2[512];
    if (ipcr_gpu_amplicons(outer, products, n, nullptr, 0, offs.data(), &need, e2, sizeof e2)) throw Error(e2);
    outer->h_amp.ensure(std::max<uint64_t>(need, 1));
    if (ipcr_gpu_amplicons(outer, products, n, outer->h_amp.as<uint8_t>(), need, offs.data(), &need, e2, sizeof e2))
        throw Error(e2);
    // 2. one batched inner scan (the reference builds an engine per amplicon, nested.go:19-20)
    for (uint64_t i = 0; i < n; i++) {
        best[i] = ipcr_product{};
        best[i].length = -1;
    }
    bind(inner);
    inner->launches = 0;
    auto better = [&](const ipcr_product &x, const ipcr_product &y) {  // nested.go:35-51
        const int mx = x.fwd_mm + x.rev_mm, my = y.fwd_mm + y.rev_mm;
        if (mx != my) return mx < my;
        if (x.length != y.length) return x.length > y.length;
        if (x.start != y.start) return x.start < y.start;
        if (x.end != y.end) return x.end < y.end;
        if (inner_ids && x.pair_idx != y.pair_idx) {
            const int c = strcmp(inner_ids[x.pair_idx], inner_ids[y.pair_idx]);
            if (c != 0) return c < 0;
        }
        return false;  // stable: the earlier emission stays
    };
    const uint8_t *amp = outer->h_amp.as<uint8_t>();
    for (uint64_t b0 = 0; b0 < n; b0 += 60000) {
        const uint64_t b1 = std::min<uint64_t>(n, b0 + 60000);
        std::vector<ipcr_seq> recs;
        for (uint64_t i = b0; i < b1; i++) recs.push_back(ipcr_seq{amp + offs[i], offs[i + 1] - offs[i]});
        const uint64_t np = scan_records(inner, recs.data(), recs.size());
        const ipcr_product *pr = inner->out_ptr;
        for (uint64_t k = 0; k < np; k++) {
            ipcr_product &cur = best[b0 + pr[k].rec_idx];
            if (cur.length < 0 || better(pr[k], cur)) {
                cur = pr[k];
                cur.rec_idx = (uint32_t)(b0 + pr[k].rec_idx);
            }
        }
    }
    return 0;
    ABI_CATCH(1)
}

int ipcr_gpu_product_mismatch_idx(const ipcr_product *pr, int which, int32_t *out, int cap) {
    if (!pr || !out) return 0;
    int n = 0;
    if (which == 0) {
        for (int j = 0; j < 64 && n < cap; j++)
            if ((pr->fwd_mm_mask >> j) & 1ull) out[n++] = j;
    } else {
        // flip() keeps list order, so ascending partner indices come out descending (engine.go:127-136)
        for (int j = 63; j >= 0 && n < cap; j--)
            if ((pr->rev_mm_mask >> j) & 1ull) out[n++] = j;
    }
    return n;
}

int ipcr_gpu_last_timing(const ipcr_scratch *s, ipcr_timing *out) {
    if (!s || !out) return 1;
    *out = s->timing;
    return 0;
}

int ipcr_gpu_synth_fill(ipcr_scratch *s, void *d_dst, uint64_t n, uint64_t first_pos, uint64_t seed, char *err,
                        size_t errlen) {
    ABI_TRY
    if (!s) throw Error("scratch is NULL");
    bind(s);
    if (n) {
        const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)s->sm_count * 32);
        k_synth_fill<<<grid, 256, 0, s->stream>>>(reinterpret_cast<uint8_t *>(d_dst), n, first_pos, seed);
        check_launch("k_synth_fill");
    }
    CU(cudaStreamSynchronize(s->stream));
    return 0;
    ABI_CATCH(1)
}

int ipcr_gpu_debug_planes(const ipcr_genome *g, uint32_t *words, uint64_t max_words, uint64_t *n_words,
                          uint32_t *stripe_rows, uint64_t *n_tiles) {
    if (!g) return 1;
    const GenomeData &d = g->g;
    const uint64_t nw = d.lay.total_len && d.planes_ready ? d.lay.n_tiles * d.lay.rows * kLanes * 4ull : 0;
    if (n_words) *n_words = nw;
    if (stripe_rows) *stripe_rows = d.lay.W;
    if (n_tiles) *n_tiles = d.lay.n_tiles;
    if (words && nw && nw <= max_words) {
        cudaSetDevice(d.device);
        if (cudaMemcpy(words, d.planes.p, nw * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
    }
    return 0;
}

int ipcr_gpu_debug_codes(const ipcr_genome *g, uint32_t *codes, uint32_t *inval, uint64_t max_groups,
                         uint64_t *n_groups) {
    if (!g) return 1;
    const GenomeData &d = g->g;
    const uint64_t ng = d.codes_ready ? d.n_groups : 0;
    if (n_groups) *n_groups = ng;
    if (codes && inval && ng && ng <= max_groups) {
        cudaSetDevice(d.device);
        if (cudaMemcpy(codes, d.codes.p, ng * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
        if (cudaMemcpy(inval, d.inval.p, ng * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
    }
    return 0;
}

int ipcr_gpu_debug_candidates(const ipcr_scratch *s, uint64_t *out, uint64_t cap, uint64_t *n) {
    if (!s) return 1;
    if (n) *n = s->n_cand;
    const uint64_t k = std::min<uint64_t>(cap, s->n_cand);
    if (out && k) {
        cudaSetDevice(s->device);
        if (cudaMemcpy(out, s->cand.p, k * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
    }
    return 0;
}

int ipcr_gpu_debug_plan_eval(const ipcr_panel *p, uint32_t orient, const uint8_t *seq, uint64_t len, int64_t start,
                             int raw_mode) {
    if (!p || orient >= p->host.orients.size()) return -1;
    const int v = raw_mode ? 1 : 0;
    return plan_eval(p->host.orients[orient], p->host.plan[v].o[orient], seq, len, start) ? 1 : 0;
}

int ipcr_gpu_debug_seed_eval(ipcr_panel *p, int raw_mode, const uint8_t *seq, uint64_t len, uint64_t *out,
                             uint64_t cap, uint64_t *n, int32_t *info) {
    if (!p) return 1;
    try {
        const SeedIndex &ix = seed_host(p, raw_mode ? 1 : 0);
        std::vector<uint64_t> c;
        seed_eval(p->host, ix, seq, len, c);
        if (n) *n = c.size();
        for (uint64_t i = 0; out && i < c.size() && i < cap; i++) out[i] = c[i];
        if (info) {
            info[0] = ix.S;
            info[1] = (int32_t)std::min<uint64_t>(ix.n_entries, 0x7FFFFFFF);
            info[2] = (int32_t)ix.table_bits;
            info[3] = ix.use_bloom ? 1 : 0;
            info[4] = (int32_t)(ix.bloom_fill * 1e6);
            info[5] = ix.bloom_k | (ix.stride << 8);
            info[6] = ix.S1;
            info[7] = (int32_t)std::min<uint64_t>(ix.n_keys1, 0x7FFFFFFF);
        }
    } catch (const std::exception &) {
        return 1;
    }
    return 0;
}

int ipcr_gpu_debug_orientation(const ipcr_panel *p, uint32_t orient, char *buf, size_t buflen, int *type) {
    if (!p || orient >= p->host.orients.size()) return 1;
    const Orient &o = p->host.orients[orient];
    if (buf && buflen) snprintf(buf, buflen, "%s", o.seq.c_str());
    if (type) *type = o.rc_type ? 'a' : 'A';
    return 0;
}

int ipcr_gpu_debug_jit_source(ipcr_panel *p, int raw_mode, char *buf, size_t buflen, size_t *needed) {
    if (!p) return 1;
    const std::string src = jit_generate_source(p->host, raw_mode ? 1 : 0);
    if (needed) *needed = src.size() + 1;
    if (buf && buflen) snprintf(buf, buflen, "%s", src.c_str());
    return 0;
}

int ipcr_gpu_debug_jit_compile(ipcr_panel *p, int raw_mode, char *cubin, size_t cap, size_t *cubin_size, char *err,
                               size_t errlen) {
    ABI_TRY
    if (!p) throw Error("panel is NULL");
    JitProgram prog = jit_generate(p->host, raw_mode ? 1 : 0, jit_config_from_env());
    const std::string log = jit_compile_all(prog);
    if (!log.empty()) throw Error(log);
    if (prog.cubins.empty()) throw Error("panel has no orientations");
    const std::vector<char> &bin = prog.cubins[0];  // first group; all groups compiled
    if (cubin_size) *cubin_size = bin.size();
    if (cubin && cap >= bin.size()) memcpy(cubin, bin.data(), bin.size());
    return 0;
    ABI_CATCH(1)
}

}  // extern "C"
