// multi.hpp -- one call, many pieces: streamed slabs on one device, and slices over several
// devices, both on top of the asynchronous chain (pair_small.cuh).  Included by engine.cu.
//
// The reference splits large inputs into overlapping chunks handled by a pool of workers and
// de-duplicates the products afterwards (/root/reference/core/fasta/path_ctx.go:83-179,
// internal/runutil/runutil.go:20-59, internal/pipeline/pipeline.go:60-161).  Here the
// concatenated records are cut into pieces with a primer-length halo; a site belongs to the piece
// holding its start, every piece leaves a site block in device memory, the blocks meet on ONE
// device (peer copies over NVLink, or the all-gather of a multi-process run) and a single pairing
// pass sees all sites -- hit cap, circular wrap and unbounded lengths work across pieces, and
// nothing is de-duplicated because nothing is found twice.
#pragma once

namespace ipcr {

constexpr uint64_t kPieceHalo = 64;  // >= IPCR_GPU_MAX_PRIMER_LEN - 1

struct PieceCut {
    std::vector<uint64_t> roff;  // record offsets in the concatenated coordinate, n_recs + 1
    std::vector<uint64_t> cut;   // piece boundaries, n_pieces + 1 (16-byte aligned inside)
    uint64_t total = 0;
};

static PieceCut cut_pieces(const ipcr_seq *recs, size_t n_recs, int n_pieces) {
    PieceCut c;
    c.roff.assign(n_recs + 1, 0);
    for (size_t r = 0; r < n_recs; r++) c.roff[r + 1] = c.roff[r] + recs[r].len;
    c.total = c.roff[n_recs];
    c.cut.resize((size_t)n_pieces + 1);
    for (int k = 0; k <= n_pieces; k++)
        c.cut[k] = k == n_pieces ? c.total : ((c.total * (uint64_t)k / (uint64_t)n_pieces) & ~15ull);
    return c;
}

// segments of piece [lo, hi) (+ halo) in piece-local coordinates; seq pointers are left NULL
static void piece_segments(const PieceCut &c, size_t n_recs, uint64_t lo, uint64_t hi, std::vector<SegmentDev> &out) {
    out.clear();
    const uint64_t end = std::min(c.total, hi + kPieceHalo);
    for (size_t r = 0; r < n_recs; r++) {
        const uint64_t s0 = std::max(c.roff[r], lo), s1 = std::min(c.roff[r + 1], end), own1 = std::min(c.roff[r + 1], hi);
        if (s0 >= s1 || own1 <= s0) continue;
        SegmentDev d{};
        d.global_off = s0 - lo;
        d.len = s1 - s0;
        d.own_len = own1 - s0;
        d.rec_off = s0 - c.roff[r];
        d.rec_idx = (uint32_t)r;
        out.push_back(d);
    }
}

// ---------------------------------------------------------------------------------------
// Streamed slabs on ONE device: the H2D copy of slab k+1.. runs under stage 1 of slab k
// (SURVEY.md 8f row 1); every slab leaves a block, stage 2 runs once, the host waits once.
// ---------------------------------------------------------------------------------------
// is this host pointer pageable memory (neither cudaMallocHost'ed nor registered)?  The Go host's []byte always is.
static bool host_is_pageable(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

// smallest pageable input that goes through the pinned staging ring (HostStager) instead of the driver's own staging
static uint64_t staged_min_bases() {
    const char *e = getenv("IPCR_STAGED_H2D_MIN");  // tests: tiny inputs too; a huge value: never
    return e ? (uint64_t)atoll(e) : (64ull << 20);
}

struct StagerJoin {  // on every way out: threads joined, their copies landed (the caller may free the source)
    HostStager *h;
    ~StagerJoin() {
        if (!h) return;
        h->finish();
        for (int t = 0; t < HostStager::kThreads; t++)
            if (h->stream[t]) cudaStreamSynchronize(h->stream[t]);
    }
};

static bool scan_streamed_chain(ipcr_scratch *s, const ipcr_seq *recs, size_t n_recs, int n_slabs, uint64_t *n_prod) {
    cudaStream_t st = s->stream, cs = s->copy_stream;
    const PieceCut c = cut_pieces(recs, n_recs, n_slabs);
    std::vector<ipcr_record_info> rinfo(n_recs, ipcr_record_info{});
    for (size_t r = 0; r < n_recs; r++) rinfo[r].len = recs[r].len;
    s->own.device = s->device;
    s->own.ascii_owned.ensure((size_t)c.total + 64);
    uint8_t *base = s->own.ascii_owned.as<uint8_t>();
    while ((int)s->slab_ev.size() < n_slabs) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        s->slab_ev.push_back(e);
    }
    while ((int)s->slab_g.size() < n_slabs) s->slab_g.emplace_back(new GenomeData());
    while ((int)s->slab_t.size() < 2 * n_slabs) {
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        s->slab_t.push_back(e);
    }
    // 1. every copy is queued now; slab k's event fires once [cut[k], cut[k+1] + halo) is resident
    CU(cudaEventRecord(s->ev[0], st));
    CU(cudaEventRecord(s->copy_ev[0], cs));
    // pageable records (what a Go / Python caller passes): several host threads stage them through a pinned ring --
    // the driver's single-threaded staging runs at ~11 GB/s, a fifth of the link
    bool staged = c.total >= staged_min_bases();
    for (size_t r = 0; r < n_recs && staged; r++)
        if (recs[r].len) {
            staged = host_is_pageable(recs[r].seq);
            break;
        }
    if (staged && !s->stager) s->stager.reset(new HostStager());
    StagerJoin stager_join{staged ? s->stager.get() : nullptr};
    uint64_t copied_to = 0;
    for (int k = 0; k < n_slabs; k++) {
        const uint64_t want = std::min(c.total, c.cut[k + 1] + kPieceHalo);
        for (size_t r = 0; r < n_recs && copied_to < want; r++) {
            const uint64_t a = std::max(copied_to, c.roff[r]), b = std::min(want, c.roff[r + 1]);
            if (a >= b) continue;
            if (staged) s->stager->add(base + a, recs[r].seq + (a - c.roff[r]), b - a, k);
            else CU(cudaMemcpyAsync(base + a, recs[r].seq + (a - c.roff[r]), b - a, cudaMemcpyHostToDevice, cs));
        }
        copied_to = std::max(copied_to, want);
        if (!staged) CU(cudaEventRecord(s->slab_ev[k], cs));
    }
    if (staged) s->stager->start(s->device, n_slabs);
    else CU(cudaEventRecord(s->copy_ev[1], cs));
    // 2. stage 1 per slab into its block, stage 2 once; capacities / plan variant may ask for a second round
    bool done = false;
    for (int attempt = 0; attempt < 6 && !done; attempt++) {
        const uint64_t site_cap = std::max<uint64_t>(1024, s->site_cap_hint);
        const uint64_t bb = block_bytes_for(n_recs, site_cap);
        s->gather.ensure((size_t)n_slabs * bb);
        const int variant = chain_variant(s, false);
        uint32_t n_blocks = 0;
        for (int k = 0; k < n_slabs; k++) {
            const uint64_t lo = c.cut[k], hi = c.cut[k + 1];
            if (lo >= hi) continue;
            GenomeData &g = *s->slab_g[k];
            g.device = s->device;
            piece_segments(c, n_recs, lo, hi, g.segs);
            if (g.segs.empty()) continue;
            g.recs = rinfo;
            g.lay.total_len = std::min(c.total, hi + kPieceHalo) - lo;
            g.ascii = base + lo;
            g.codes_ext = g.inval_ext = nullptr;
            g.own_bases = hi - lo;
            g.planes_ready = g.codes_ready = false;
            if (attempt == 0) {
                if (staged) s->stager->wait_mark(k, st);
                else CU(cudaStreamWaitEvent(st, s->slab_ev[k], 0));
            }
            CU(cudaEventRecord(s->slab_t[2 * k], st));
            stage1_async(s, g, variant, s->gather.as<uint8_t>() + (size_t)n_blocks * bb, site_cap);
            CU(cudaEventRecord(s->slab_t[2 * k + 1], st));
            n_blocks++;
        }
        if (staged && attempt == 0) {  // the copy stream's end-of-copies event, for the timing
            s->stager->wait_mark(n_slabs - 1, cs);
            CU(cudaEventRecord(s->copy_ev[1], cs));
        }
        if (n_blocks == 0) {
            *n_prod = 0;
            s->out_ptr = s->h_products.as<ipcr_product>();
            done = true;
            break;
        }
        for (int again = 0; again < 4 && !done; again++) {
            pair_async(s, s->gather.as<uint8_t>(), n_blocks, bb, site_cap, rinfo.data(), n_recs);
            bool redo = false;
            const ChainResult r = chain_finish(s, s->gather.as<uint8_t>(), n_blocks, bb, rinfo.data(), n_recs, n_prod, &redo);
            if (r == kChainTooDirty) {
                CU(cudaStreamSynchronize(cs));
                return false;
            }
            if (r == kChainDone) done = true;
            else if (redo) break;
        }
    }
    if (!done) throw Error("the streamed scan chain did not converge");
    CU(cudaStreamSynchronize(cs));
    s->own.recs = rinfo;
    s->last_valid = true;
    s->last_ascii = base;
    s->last_rec_base.assign(c.roff.begin(), c.roff.begin() + n_recs);
    s->last_rec_len.resize(n_recs);
    for (size_t r = 0; r < n_recs; r++) s->last_rec_len[r] = recs[r].len;
    ipcr_timing &t = s->timing;
    float stage1 = 0.f;
    for (int k = 0; k < n_slabs; k++)
        if (c.cut[k] < c.cut[k + 1]) stage1 += ev_ms(s->slab_t[2 * k], s->slab_t[2 * k + 1]);
    t.h2d_ms = ev_ms(s->copy_ev[0], s->copy_ev[1]);
    t.pack_ms = 0.f;
    t.filter_ms = stage1;  // pack + scan + verify of every slab (the chain does not time them apart here)
    t.verify_ms = 0.f;
    t.pair_ms = ev_ms(s->ev[5], s->ev[6]);
    t.d2h_ms = ev_ms(s->ev[6], s->ev[7]);
    t.total_ms = ev_ms(s->ev[0], s->ev[7]);
    t.products = *n_prod;
    t.kernel_launches = s->launches;
    return true;
}

// ---------------------------------------------------------------------------------------
// Packed host genome (pack_host.hpp: what the native feeder leaves in pinned memory) -> products.
// 0.375 B/bp cross PCIe instead of 1 B/bp, there is no pack kernel, and the exact verify reads the
// same codes.  Slabs are cut on 4096-base boundaries so that every slab is a whole number of packed
// words; the copy of slab k+1.. runs under stage 1 of slab k.
// ---------------------------------------------------------------------------------------
struct PackedHost {
    const uint32_t *codes = nullptr;  // [2 * packed_groups(total)]
    const uint32_t *inval = nullptr;  // [packed_groups(total)]
    uint64_t total = 0;               // bases in the concatenated buffer
    const uint64_t *runs = nullptr;   // optional: the bitmap as [start, end) runs (then only the codes cross PCIe)
    uint64_t n_runs = 0;
    bool pinned = true;               // false: codes / inval are pageable (the feeder page-locks on request only)
};

static bool scan_packed_chain(ipcr_scratch *s, const PackedHost &ph, const std::vector<ipcr_record_info> &rinfo,
                              const std::vector<uint64_t> &rec_off, uint64_t *n_prod) {
    cudaStream_t st = s->stream, cs = s->copy_stream;
    const size_t n_recs = rinfo.size();
    const uint64_t total = ph.total, groups = packed_groups(total);
    int n_slabs = (int)std::min<uint64_t>(8, std::max<uint64_t>(1, total / (256ull << 20)));
    if (const char *e = getenv("IPCR_STREAM_SLABS")) n_slabs = std::max(1, atoi(e));
    n_slabs = std::min<int>(n_slabs, (int)kPairMaxBlocks);
    PieceCut c;
    c.roff = rec_off;  // n_recs + 1 entries
    c.total = total;
    c.cut.resize((size_t)n_slabs + 1);
    for (int k = 0; k <= n_slabs; k++)
        c.cut[k] = k == n_slabs ? total : ((total * (uint64_t)k / (uint64_t)n_slabs) & ~(kPackChunk - 1));
    GenomeData &own = s->own;
    own.device = s->device;
    own.codes.ensure((size_t)groups * 8 + 64);
    own.inval.ensure((size_t)groups * 4 + 64);
    while ((int)s->slab_ev.size() < n_slabs) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        s->slab_ev.push_back(e);
    }
    while ((int)s->slab_g.size() < n_slabs) s->slab_g.emplace_back(new GenomeData());
    while ((int)s->slab_t.size() < 2 * n_slabs) {
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        s->slab_t.push_back(e);
    }
    // 1. copies: slab k is ready once every group its scan may touch is resident
    //    (its chunks, the halo's chunk and one look-ahead group)
    CU(cudaEventRecord(s->ev[0], st));
    CU(cudaEventRecord(s->copy_ev[0], cs));
    const bool by_runs = ph.runs != nullptr;
    if (by_runs) {  // the bitmap is rebuilt on the device from its run list: 0.25 B/bp cross PCIe instead of 0.375
        CU(cudaMemsetAsync(own.inval.p, 0, (size_t)groups * 4, st));
        if (ph.n_runs) {
            upload(s->runs_dev, ph.runs, ph.n_runs * 16, st);
            const unsigned grid = (unsigned)std::min<uint64_t>((ph.n_runs * 32 + 255) / 256, (uint64_t)s->sm_count * 8);
            k_inval_fill<<<grid, 256, 0, st>>>(s->runs_dev.as<unsigned long long>(), ph.n_runs, own.inval.as<uint32_t>());
            check_launch("k_inval_fill");
            s->launches++;
        }
    }
    // pageable source (a one-shot FASTA run): several host threads stage the chunks through a pinned ring instead
    // of the driver's single-threaded staging
    const bool staged = !ph.pinned && total >= staged_min_bases();
    if (staged && !s->stager) s->stager.reset(new HostStager());
    StagerJoin stager_join{staged ? s->stager.get() : nullptr};
    uint64_t g_done = 0;
    for (int k = 0; k < n_slabs; k++) {
        const uint64_t end = std::min(total, c.cut[k + 1] + kPieceHalo);
        const uint64_t g_want = std::min(groups, (end + kPackChunk - 1) / kPackChunk * (kPackChunk / 32u) + 1);
        if (g_want > g_done) {
            if (staged) {
                s->stager->add(own.codes.as<uint32_t>() + 2 * g_done, ph.codes + 2 * g_done, (g_want - g_done) * 8, k);
                if (!by_runs) s->stager->add(own.inval.as<uint32_t>() + g_done, ph.inval + g_done, (g_want - g_done) * 4, k);
            } else {
                CU(cudaMemcpyAsync(own.codes.as<uint32_t>() + 2 * g_done, ph.codes + 2 * g_done, (g_want - g_done) * 8,
                                   cudaMemcpyHostToDevice, cs));
                if (!by_runs)
                    CU(cudaMemcpyAsync(own.inval.as<uint32_t>() + g_done, ph.inval + g_done, (g_want - g_done) * 4,
                                       cudaMemcpyHostToDevice, cs));
            }
            g_done = g_want;
        }
        if (!staged) CU(cudaEventRecord(s->slab_ev[k], cs));
    }
    if (staged) s->stager->start(s->device, n_slabs);
    else CU(cudaEventRecord(s->copy_ev[1], cs));
    bool done = false;
    for (int attempt = 0; attempt < 6 && !done; attempt++) {
        const uint64_t site_cap = std::max<uint64_t>(1024, s->site_cap_hint);
        const uint64_t bb = block_bytes_for(n_recs, site_cap);
        s->gather.ensure((size_t)n_slabs * bb);
        const int variant = chain_variant(s, false);
        uint32_t n_blocks = 0;
        for (int k = 0; k < n_slabs; k++) {
            const uint64_t lo = c.cut[k], hi = c.cut[k + 1];
            if (lo >= hi) continue;
            GenomeData &g = *s->slab_g[k];
            g.device = s->device;
            piece_segments(c, n_recs, lo, hi, g.segs);
            if (g.segs.empty()) continue;
            g.recs = rinfo;
            g.lay.total_len = std::min(total, hi + kPieceHalo) - lo;
            g.ascii = nullptr;
            g.codes_ext = own.codes.as<uint32_t>() + lo / 16;
            g.inval_ext = own.inval.as<uint32_t>() + lo / 32;
            g.own_bases = hi - lo;
            if (attempt == 0) {
                if (staged) s->stager->wait_mark(k, st);
                else CU(cudaStreamWaitEvent(st, s->slab_ev[k], 0));
            }
            CU(cudaEventRecord(s->slab_t[2 * k], st));
            stage1_async(s, g, variant, s->gather.as<uint8_t>() + (size_t)n_blocks * bb, site_cap);
            CU(cudaEventRecord(s->slab_t[2 * k + 1], st));
            n_blocks++;
        }
        if (staged && attempt == 0) {  // the copy stream's end-of-copies event, for the timing
            s->stager->wait_mark(n_slabs - 1, cs);
            CU(cudaEventRecord(s->copy_ev[1], cs));
        }
        if (n_blocks == 0) {
            *n_prod = 0;
            s->out_ptr = s->h_products.as<ipcr_product>();
            done = true;
            break;
        }
        for (int again = 0; again < 4 && !done; again++) {
            std::vector<ipcr_record_info> ri = rinfo;
            pair_async(s, s->gather.as<uint8_t>(), n_blocks, bb, site_cap, ri.data(), n_recs);
            bool redo = false;
            const ChainResult r = chain_finish(s, s->gather.as<uint8_t>(), n_blocks, bb, ri.data(), n_recs, n_prod, &redo);
            if (r == kChainTooDirty) {
                CU(cudaStreamSynchronize(cs));
                return false;
            }
            if (r == kChainDone) done = true;
            else if (redo) break;
        }
    }
    if (!done) throw Error("the packed scan chain did not converge");
    CU(cudaStreamSynchronize(cs));
    s->last_valid = false;  // no ASCII on the device: the visitors need ipcr_gpu_scan / _scan_batch
    ipcr_timing &t = s->timing;
    float stage1 = 0.f;
    for (int k = 0; k < n_slabs; k++)
        if (c.cut[k] < c.cut[k + 1]) stage1 += ev_ms(s->slab_t[2 * k], s->slab_t[2 * k + 1]);
    t.h2d_ms = ev_ms(s->copy_ev[0], s->copy_ev[1]);
    t.pack_ms = 0.f;
    t.filter_ms = stage1;  // scan + verify of every slab
    t.verify_ms = 0.f;
    t.pair_ms = ev_ms(s->ev[5], s->ev[6]);
    t.d2h_ms = ev_ms(s->ev[6], s->ev[7]);
    t.total_ms = ev_ms(s->ev[0], s->ev[7]);
    t.products = *n_prod;
    t.kernel_launches = s->launches;
    return true;
}

// ---------------------------------------------------------------------------------------
// Several devices inside ONE process (the Go host is one process calling the C ABI):
// the records are cut into one slice per device; every device copies its slice from the
// caller's host memory over its own PCIe link and runs stage 1; the site blocks travel to
// the calling scratch's device with peer copies; stage 2 runs there once.
// ---------------------------------------------------------------------------------------
static std::atomic<int> g_n_devices{1};  // devices one call may spread over (ipcr_gpu_init)

static uint64_t multi_min_bases() {
    const char *e = getenv("IPCR_MULTI_MIN_BP");  // read per call: tests lower it
    return e ? (uint64_t)atoll(e) : (256ull << 20);
}

static void check_device(int device) {
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        throw Error(fmt("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
                        prop.minor));
}

static ipcr_scratch *make_scratch(ipcr_panel *p, int device) {
    CU(cudaSetDevice(device));
    check_device(device);
    auto s = std::make_unique<ipcr_scratch>();
    s->panel = p;
    s->device = device;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    s->sm_count = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    for (auto &e : s->ev) CU(cudaEventCreate(&e));
    for (auto &e : s->copy_ev) CU(cudaEventCreate(&e));
    CU(cudaEventCreateWithFlags(&s->done_ev, cudaEventDisableTiming));
    s->counters.ensure(8 * sizeof(unsigned long long));
    s->h_counts.ensure(8 * sizeof(unsigned long long));
    s->tables = tables_for(p, s->device, s->stream);
    return s.release();
}


// peers of a scratch: one more scratch per further device, created on first use
static void ensure_peers(ipcr_scratch *s, int n_devices) {
    int visible = 0;
    CU(cudaGetDeviceCount(&visible));
    n_devices = std::min(n_devices, visible);
    for (int d = 0; (int)s->peers.size() + 1 < n_devices && d < visible; d++) {
        if (d == s->device) continue;
        bool have = false;
        for (ipcr_scratch *q : s->peers) have |= q->device == d;
        if (have) continue;
        ipcr_scratch *q = make_scratch(s->panel, d);
        s->peers.push_back(q);
        // direct NVLink copies between the two devices (already enabled is fine)
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, d, s->device));
        if (can) {
            CU(cudaSetDevice(d));
            if (cudaDeviceEnablePeerAccess(s->device, 0) != cudaSuccess) cudaGetLastError();
            CU(cudaSetDevice(s->device));
            if (cudaDeviceEnablePeerAccess(d, 0) != cudaSuccess) cudaGetLastError();
        }
    }
    CU(cudaSetDevice(s->device));
}

static bool scan_multi_device(ipcr_scratch *s, const ipcr_seq *recs, size_t n_recs, uint64_t *n_prod) {
    ensure_peers(s, g_n_devices.load());
    const int n_dev = (int)s->peers.size() + 1;
    if (n_dev < 2) return false;
    const PieceCut c = cut_pieces(recs, n_recs, n_dev);
    std::vector<ipcr_record_info> rinfo(n_recs, ipcr_record_info{});
    for (size_t r = 0; r < n_recs; r++) rinfo[r].len = recs[r].len;
    cudaStream_t st = s->stream;
    CU(cudaEventRecord(s->ev[0], st));
    bool done = false;
    for (int attempt = 0; attempt < 6 && !done; attempt++) {
        const uint64_t site_cap = std::max<uint64_t>(1024, s->site_cap_hint);
        const uint64_t bb = block_bytes_for(n_recs, site_cap);
        s->gather.ensure((size_t)n_dev * bb);
        const int variant = chain_variant(s, false);
        std::vector<std::string> errs((size_t)n_dev);
        auto job = [&](int k) {
            try {
                ipcr_scratch *q = k == 0 ? s : s->peers[(size_t)k - 1];
                CU(cudaSetDevice(q->device));
                if (k > 0) {
                    if (q->inflight) CU(cudaStreamSynchronize(q->stream));
                    q->inflight = false;
                    q->args_used = 0;
                    q->launches = 0;
                    q->seed_cap_hint = s->seed_cap_hint;
                    q->seed_dirty_hint = s->seed_dirty_hint;
                }
                std::vector<SegmentDev> sd;
                piece_segments(c, n_recs, c.cut[k], c.cut[k + 1], sd);
                std::vector<ipcr_segment> segs(sd.size());
                for (size_t i = 0; i < sd.size(); i++) {
                    segs[i].seq = recs[sd[i].rec_idx].seq + sd[i].rec_off;
                    segs[i].len = sd[i].len;
                    segs[i].own_len = sd[i].own_len;
                    segs[i].rec_off = sd[i].rec_off;
                    segs[i].rec_idx = sd[i].rec_idx;
                    segs[i].reserved = 0;
                }
                uint8_t *dst = s->gather.as<uint8_t>() + (size_t)k * bb;
                if (attempt == 0 || !q->own_laid) layout_genome(q, q->own, segs.data(), segs.size(), n_recs, false);
                q->own_laid = true;
                // whole-record lengths, not this slice's view: stage 2 and the sort key need them
                for (size_t r = 0; r < n_recs; r++) q->own.recs[r].len = recs[r].len;
                if (q->own.lay.total_len == 0) {  // an empty slice still owes an (empty) block
                    Stage1Ctl init{};
                    init.site_cap = site_cap;
                    init.variant = (uint32_t)variant;
                    init.n_recs = (uint32_t)n_recs;
                    q->block.ensure(bb);
                    uint8_t *blk = k == 0 ? dst : q->block.as<uint8_t>();
                    q->seg_reset_dummy.ensure(16);
                    k_block_init<<<1, 256, 0, q->stream>>>(blk, init, q->seg_reset_dummy.as<uint32_t>(), 0, nullptr);
                    check_launch("k_block_init");
                } else if (k == 0) {
                    stage1_async(q, q->own, variant, dst, site_cap);
                } else {
                    q->block.ensure(bb);
                    stage1_async(q, q->own, variant, q->block.as<uint8_t>(), site_cap);
                }
                if (k > 0) {
                    CU(cudaMemcpyPeerAsync(dst, s->device, q->block.p, q->device, bb, q->stream));
                    CU(cudaEventRecord(q->done_ev, q->stream));
                    q->inflight = true;
                }
            } catch (const std::exception &e) {
                errs[(size_t)k] = e.what();
                if (errs[(size_t)k].empty()) errs[(size_t)k] = "unknown error";
            }
        };
        {  // one host thread per further device: pageable caller memory makes the H2D copies blocking
            std::vector<std::thread> th;
            for (int k = 1; k < n_dev; k++) th.emplace_back(job, k);
            job(0);
            for (auto &t : th) t.join();
        }
        CU(cudaSetDevice(s->device));
        for (const std::string &e : errs)
            if (!e.empty()) throw Error(e);
        for (ipcr_scratch *q : s->peers) CU(cudaStreamWaitEvent(st, q->done_ev, 0));
        for (int again = 0; again < 4 && !done; again++) {
            pair_async(s, s->gather.as<uint8_t>(), (uint32_t)n_dev, bb, site_cap, rinfo.data(), n_recs);
            bool redo = false;
            const ChainResult r = chain_finish(s, s->gather.as<uint8_t>(), (uint32_t)n_dev, bb, rinfo.data(), n_recs, n_prod, &redo);
            if (r == kChainTooDirty) {
                for (ipcr_scratch *q : s->peers) q->own_laid = false;
                s->own_laid = false;
                return false;
            }
            if (r == kChainDone) done = true;
            else if (redo) break;
        }
    }
    for (ipcr_scratch *q : s->peers) {
        q->own_laid = false;
        q->inflight = false;  // chain_finish waited for the primary stream, which waited for every peer
        s->launches += q->launches;
    }
    s->own_laid = false;
    if (!done) throw Error("the multi-device scan chain did not converge");
    s->last_valid = false;  // no device holds whole records: the device-side visitors need a single-device scan
    chain_timing(s, *n_prod, true);
    s->timing.total_ms = ev_ms(s->ev[0], s->ev[7]);
    s->timing.kernel_launches = s->launches;
    return true;
}

// ---------------------------------------------------------------------------------------
// The packed host genome over several devices of ONE process (ipcr_gpu_scan_fasta after
// ipcr_gpu_init(n)): slices are cut on 4096-base boundaries, so a slice of the packed arrays
// is a plain pointer offset; every device copies ITS slice of the codes over its own PCIe
// link (0.25 B/bp in total, the bitmap from the clipped run list), runs stage 1 on it, and
// peer-copies its site block to the calling scratch's device, where stage 2 runs once.
// ---------------------------------------------------------------------------------------
static bool scan_packed_multi(ipcr_scratch *s, const PackedHost &ph, const std::vector<ipcr_record_info> &rinfo,
                              const std::vector<uint64_t> &rec_off, uint64_t *n_prod) {
    ensure_peers(s, std::min<int>(g_n_devices.load(), (int)kPairMaxBlocks));
    const int n_dev = (int)s->peers.size() + 1;
    if (n_dev < 2) return false;
    const size_t n_recs = rinfo.size();
    const uint64_t total = ph.total, groups = packed_groups(total);
    PieceCut c;
    c.roff = rec_off;
    c.total = total;
    c.cut.resize((size_t)n_dev + 1);
    for (int k = 0; k <= n_dev; k++)
        c.cut[k] = k == n_dev ? total : ((total * (uint64_t)k / (uint64_t)n_dev) & ~(kPackChunk - 1));
    cudaStream_t st = s->stream;
    CU(cudaEventRecord(s->ev[0], st));
    const bool by_runs = ph.runs != nullptr;
    bool done = false;
    for (int attempt = 0; attempt < 6 && !done; attempt++) {
        const uint64_t site_cap = std::max<uint64_t>(1024, s->site_cap_hint);
        const uint64_t bb = block_bytes_for(n_recs, site_cap);
        s->gather.ensure((size_t)n_dev * bb);
        const int variant = chain_variant(s, false);
        std::vector<std::string> errs((size_t)n_dev);
        auto job = [&](int k) {
            try {
                ipcr_scratch *q = k == 0 ? s : s->peers[(size_t)k - 1];
                CU(cudaSetDevice(q->device));
                if (k > 0) {
                    if (q->inflight) CU(cudaStreamSynchronize(q->stream));
                    q->inflight = false;
                    q->args_used = 0;
                    q->launches = 0;
                    q->seed_cap_hint = s->seed_cap_hint;
                    q->seed_dirty_hint = s->seed_dirty_hint;
                }
                const uint64_t lo = c.cut[k], hi = c.cut[k + 1];
                uint8_t *dst = s->gather.as<uint8_t>() + (size_t)k * bb;
                if (q->slab_g.empty()) q->slab_g.emplace_back(new GenomeData());
                GenomeData &g = *q->slab_g[0];
                g.device = q->device;
                if (lo < hi) piece_segments(c, n_recs, lo, hi, g.segs);
                else g.segs.clear();
                if (g.segs.empty()) {  // an empty slice still owes an (empty) block
                    Stage1Ctl init{};
                    init.site_cap = site_cap;
                    init.variant = (uint32_t)variant;
                    init.n_recs = (uint32_t)n_recs;
                    q->block.ensure(bb);
                    uint8_t *blk = k == 0 ? dst : q->block.as<uint8_t>();
                    q->seg_reset_dummy.ensure(16);
                    k_block_init<<<1, 256, 0, q->stream>>>(blk, init, q->seg_reset_dummy.as<uint32_t>(), 0, nullptr);
                    check_launch("k_block_init");
                } else {
                    // the groups this slice's scan may touch: its chunks, the halo's chunk, one look-ahead group
                    const uint64_t end = std::min(total, hi + kPieceHalo);
                    const uint64_t g_lo = lo / 32u;
                    const uint64_t g_hi = std::min(groups, (end + kPackChunk - 1) / kPackChunk * (kPackChunk / 32u) + 1);
                    const uint64_t ng = g_hi - g_lo;
                    GenomeData &own = q->own;
                    own.device = q->device;
                    if (attempt == 0 || !q->own_laid) {
                        own.codes.ensure((size_t)ng * 8 + 64);
                        own.inval.ensure((size_t)ng * 4 + 64);
                        CU(cudaMemcpyAsync(own.codes.p, ph.codes + 2 * g_lo, ng * 8, cudaMemcpyHostToDevice, q->stream));
                        if (!by_runs) {
                            CU(cudaMemcpyAsync(own.inval.p, ph.inval + g_lo, ng * 4, cudaMemcpyHostToDevice, q->stream));
                        } else {
                            CU(cudaMemsetAsync(own.inval.p, 0, (size_t)ng * 4, q->stream));
                            // the runs that reach into [lo, cov), clipped and moved to slice coordinates
                            const uint64_t cov = g_hi * 32u;
                            std::vector<uint64_t> mine;
                            size_t a = 0, b = (size_t)ph.n_runs;  // first run whose end lies behind lo
                            while (a < b) {
                                const size_t m = (a + b) / 2;
                                if (ph.runs[2 * m + 1] <= lo) a = m + 1;
                                else b = m;
                            }
                            for (size_t r = a; r < (size_t)ph.n_runs && ph.runs[2 * r] < cov; r++) {
                                mine.push_back(std::max(ph.runs[2 * r], lo) - lo);
                                mine.push_back(std::min(ph.runs[2 * r + 1], cov) - lo);
                            }
                            if (!mine.empty()) {
                                const uint64_t nr = mine.size() / 2;
                                upload_staged(q, q->runs_dev, mine.data(), mine.size() * 8);
                                const unsigned grid =
                                    (unsigned)std::min<uint64_t>((nr * 32 + 255) / 256, (uint64_t)q->sm_count * 8);
                                k_inval_fill<<<grid, 256, 0, q->stream>>>(q->runs_dev.as<unsigned long long>(), nr,
                                                                          own.inval.as<uint32_t>());
                                check_launch("k_inval_fill");
                                q->launches++;
                            }
                        }
                        q->own_laid = true;
                    }
                    g.recs = rinfo;
                    g.lay.total_len = end - lo;
                    g.ascii = nullptr;
                    g.codes_ext = own.codes.as<uint32_t>();
                    g.inval_ext = own.inval.as<uint32_t>();
                    g.own_bases = hi - lo;
                    if (k == 0) {
                        stage1_async(q, g, variant, dst, site_cap);
                    } else {
                        q->block.ensure(bb);
                        stage1_async(q, g, variant, q->block.as<uint8_t>(), site_cap);
                    }
                }
                if (k > 0) {
                    CU(cudaMemcpyPeerAsync(dst, s->device, q->block.p, q->device, bb, q->stream));
                    CU(cudaEventRecord(q->done_ev, q->stream));
                    q->inflight = true;
                }
            } catch (const std::exception &e) {
                errs[(size_t)k] = e.what();
                if (errs[(size_t)k].empty()) errs[(size_t)k] = "unknown error";
            }
        };
        {  // one host thread per further device: the packed arrays are only pinned on request
            std::vector<std::thread> th;
            for (int k = 1; k < n_dev; k++) th.emplace_back(job, k);
            job(0);
            for (auto &t : th) t.join();
        }
        CU(cudaSetDevice(s->device));
        for (const std::string &e : errs)
            if (!e.empty()) throw Error(e);
        for (ipcr_scratch *q : s->peers) CU(cudaStreamWaitEvent(st, q->done_ev, 0));
        for (int again = 0; again < 4 && !done; again++) {
            std::vector<ipcr_record_info> ri = rinfo;
            pair_async(s, s->gather.as<uint8_t>(), (uint32_t)n_dev, bb, site_cap, ri.data(), n_recs);
            bool redo = false;
            const ChainResult r = chain_finish(s, s->gather.as<uint8_t>(), (uint32_t)n_dev, bb, ri.data(), n_recs, n_prod, &redo);
            if (r == kChainTooDirty) {
                for (ipcr_scratch *q : s->peers) q->own_laid = false;
                s->own_laid = false;
                return false;
            }
            if (r == kChainDone) done = true;
            else if (redo) break;
        }
    }
    for (ipcr_scratch *q : s->peers) {
        q->own_laid = false;
        q->inflight = false;  // chain_finish waited for the primary stream, which waited for every peer
        s->launches += q->launches;
    }
    s->own_laid = false;
    if (!done) throw Error("the multi-device packed scan chain did not converge");
    s->last_valid = false;
    chain_timing(s, *n_prod, true);
    s->timing.pack_ms = 0.f;
    s->timing.total_ms = ev_ms(s->ev[0], s->ev[7]);
    s->timing.kernel_launches = s->launches;
    return true;
}

}  // namespace ipcr
