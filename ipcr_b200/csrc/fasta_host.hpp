// fasta_host.hpp -- native feeder for the scan: FASTA file -> normalised records in one host
// buffer, ready for slab-wise H2D.
//
// Record semantics follow the reference reader exactly
// (/root/reference/core/fasta/scan.go:10-69, normalize.go:5-14, stream.go:125-131,
//  path_ctx.go:40-81, open.go:29-50):
//   * a line whose FIRST byte is '>' is a header; ID = TrimSpace(header) up to the first
//     space or tab; a record with an empty ID is dropped together with its sequence lines;
//     sequence lines before the first header are ignored;
//   * every other line is TrimSpace'd (ASCII \t \n \v \f \r and space at both ends only),
//     a-z are upper-cased, every other byte is kept as is;
//   * gzip input is detected by magic number or a .gz suffix.
// (One deliberate difference: the reference hands lines longer than 1 MiB to TrimSpace in
//  1 MiB fragments; here a line is always trimmed as a whole.)
// The work is split over threads at line boundaries: pass 1 counts (64-byte SSE2 blocks: a block without '>'
// or stray white space contributes its size minus its newlines), pass 2 copies and upper-cases and -- when the
// caller asks for the packed form -- packs every finished 64 KB of output while it is still in cache.
#pragma once

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <emmintrin.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "pack_host.hpp"

namespace ipcr {

struct FastaRecord {
    std::string id;
    uint64_t off = 0;  // into the sequence buffer
    uint64_t len = 0;
};

inline bool fasta_is_space(uint8_t c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

inline std::string fasta_header_id(const uint8_t *h, size_t n) {  // stream.go:125-131
    size_t a = 0, b = n;
    while (a < b && fasta_is_space(h[a])) a++;
    while (b > a && fasta_is_space(h[b - 1])) b--;
    size_t e = a;
    while (e < b && h[e] != ' ' && h[e] != '\t') e++;
    return std::string(reinterpret_cast<const char *>(h + a), e - a);
}

struct FastaPiece {      // what one thread found in its byte range
    uint64_t head_kept = 0;  // kept bytes before the first header of the range (belong to the open record)
    struct Hdr { std::string id; uint64_t kept; };
    std::vector<Hdr> hdrs;
};

// calls f(line_begin, line_end_exclusive_of_newline) for every line starting in [lo, hi)
template <class F>
inline void fasta_for_lines(const uint8_t *d, size_t n, size_t lo, size_t hi, F &&f) {
    size_t p = lo;
    while (p < hi) {
        const uint8_t *nl = static_cast<const uint8_t *>(memchr(d + p, '\n', n - p));
        const size_t q = nl ? (size_t)(nl - d) : n;
        f(p, q);
        p = q + 1;
    }
}

inline void fasta_trim(const uint8_t *d, size_t &a, size_t &b) {
    while (a < b && fasta_is_space(d[a])) a++;
    while (b > a && fasta_is_space(d[b - 1])) b--;
}

// pass 1 over the lines that START in [lo, hi): kept bytes per (open record | header) -> pc
inline void fasta_count_range(const uint8_t *d, size_t n, size_t lo, const size_t hi, FastaPiece &pc) {
    uint64_t *cur = &pc.head_kept;
    // slow path: line by line (headers, lines with white space at their ends)
    auto lines = [&](size_t lo, size_t hi) {
        fasta_for_lines(d, n, lo, hi, [&](size_t a, size_t b) {
            if (b > a && d[a] == '>') {
                pc.hdrs.push_back({fasta_header_id(d + a + 1, b - a - 1), 0});
                cur = &pc.hdrs.back().kept;
                return;
            }
            fasta_trim(d, a, b);
            *cur += b - a;
        });
    };
    // fast path: whole lines that hold no '>' and no white space except their '\n' keep every other byte.
    // p always sits at a line start; blocks of 64 bytes are classified with SSE2 compares.
    size_t p = lo;
    const __m128i v_nl = _mm_set1_epi8('\n'), v_gt = _mm_set1_epi8('>'), v_sp = _mm_set1_epi8(' ');
    const __m128i v_lo = _mm_set1_epi8((char)(-0x09 - 128)), v_span = _mm_set1_epi8((char)(0x0D - 0x09 + 1 - 128));
    while (p < hi) {
        size_t q = p;
        uint64_t kept = 0;
        size_t last_nl = p;  // one past the last newline seen in clean blocks
        bool dirty = false;
        while (q + 64 <= hi) {
            uint32_t nl_bits[4];
            uint32_t bad = 0;
            for (int k = 0; k < 4; k++) {
                const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i *>(d + q + 16 * k));
                const uint32_t nl = (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(v, v_nl));
                // bytes 0x09..0x0D: (v - 0x09) < 5 unsigned, done as a signed compare after biasing by 128
                const __m128i ws = _mm_cmplt_epi8(_mm_add_epi8(v, v_lo), v_span);
                const uint32_t other = (uint32_t)_mm_movemask_epi8(
                    _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(v, v_gt), _mm_cmpeq_epi8(v, v_sp)), ws));
                nl_bits[k] = nl;
                bad |= other & ~nl;
            }
            if (bad) {
                dirty = true;
                break;
            }
            uint32_t cnt = 0;
            for (int k = 0; k < 4; k++) {
                cnt += (uint32_t)__builtin_popcount(nl_bits[k]);
                if (nl_bits[k]) last_nl = q + 16 * k + (32 - __builtin_clz(nl_bits[k]));
            }
            kept += 64 - cnt;
            q += 64;
        }
        // the clean blocks cover [p, q): whole lines up to last_nl count as they are, the unfinished line
        // (from last_nl) goes to the slow path together with whatever follows in the dirty block
        if (last_nl > p) {
            *cur += kept - (q - last_nl);
            p = last_nl;
        }
        if (!dirty && q + 64 > hi) {  // tail of the range
            lines(p, hi);
            break;
        }
        // slow path over the lines that start before the end of the dirty block
        size_t stop = std::min(hi, q + 64);
        const uint8_t *nl = static_cast<const uint8_t *>(memchr(d + stop - 1, '\n', n - (stop - 1)));
        stop = nl ? (size_t)(nl - d) + 1 : n;
        stop = std::min(stop, hi);  // lines START inside [p, hi); fasta_for_lines reads each to its own end
        lines(p, stop);
        p = stop;
        if (stop >= hi) break;
    }
}

// the line start at or after p (range boundaries are snapped to line starts)
inline size_t fasta_snap(const uint8_t *d, size_t n, size_t p) {
    if (p == 0 || p >= n || d[p - 1] == '\n') return std::min(p, n);
    const uint8_t *nl = static_cast<const uint8_t *>(memchr(d + p, '\n', n - p));
    return nl ? (size_t)(nl - d) + 1 : n;
}

// Parses `n` bytes of FASTA text.  `alloc(total)` must return a buffer for the sequence bytes.
struct FastaPackTarget {  // optional: the packed form (pack_host.hpp) of the sequence buffer, filled by pass 2
    uint32_t *codes = nullptr, *inval = nullptr;
    // pack-only: no sequence buffer at all (alloc() may return NULL).  Every thread normalises into a 64 KB staging
    // buffer and packs from there; the (at most two) 32-base groups a thread shares with its neighbours, and the
    // ragged last group, are assembled in 32-byte edge slots and packed after the join.  Saves writing -- and
    // page-faulting in -- 1 B/bp of memory nobody reads when the scan takes the packed form.
    bool only = false;
    // filled by the one-trip parse: sorted, disjoint group ranges outside which the bitmap is all zero
    std::vector<std::pair<uint64_t, uint64_t>> dirty;
    bool dirty_known = false;
};

// pass 2 of the pack-only mode: the normalised bytes of one thread's output range [start, end) -> codes / bitmap
struct FastaPackStage {
    static constexpr size_t kCap = 64 * 1024;
    std::vector<uint8_t> buf;
    uint64_t A = 0, B = 0;      // interior of the range: whole groups [A / 32, B / 32)
    uint64_t w = 0;             // next output offset
    uint64_t base = 0;          // offset of buf[0] (a multiple of 32)
    size_t fill = 0;
    uint8_t *head = nullptr, *tail = nullptr;  // edge slots of the groups holding `start` / `end` (when not aligned)
    uint64_t head_base = 0, tail_base = 0;
    uint32_t *codes = nullptr, *inval = nullptr;

    static void norm(uint8_t *dst, const uint8_t *src, size_t k) {  // normalize.go:8-10 (auto-vectorised)
        for (size_t i = 0; i < k; i++) {
            const uint8_t c = src[i];
            dst[i] = (uint8_t)(c - (((uint8_t)(c - 'a') < 26u) ? 32u : 0u));
        }
    }
    void flush() {
        const size_t ng = fill / 32;
        if (ng) pack_groups_staged(buf.data(), base / 32, ng, codes, inval);
        const size_t rest = fill - ng * 32;
        if (rest) memmove(buf.data(), buf.data() + ng * 32, rest);
        base += (uint64_t)ng * 32;
        fill = rest;
    }
    void put(const uint8_t *src, size_t k) {
        while (k) {
            if (w < A) {  // the group shared with the previous thread(s)
                const size_t m = (size_t)std::min<uint64_t>(k, A - w);
                norm(head + (w - head_base), src, m);
                w += m; src += m; k -= m;
            } else if (w < B) {
                const size_t m = (size_t)std::min<uint64_t>(std::min<uint64_t>(k, B - w), kCap - fill);
                norm(buf.data() + fill, src, m);
                fill += m; w += m; src += m; k -= m;
                if (fill >= kCap) flush();
            } else {  // the group shared with the next thread(s), or the ragged last group of the buffer
                norm(tail + (w - tail_base), src, k);
                w += k; k = 0;
            }
        }
    }
};

// Pack-only parse in ONE trip through memory.  The file is cut into 512 KB blocks (at line starts) that the threads
// claim in order.  A thread walks its block ONCE, line by line (header / TrimSpace / upper-case rules), writing the
// kept bytes into a block-sized staging buffer (which stays in its L2) and counting them per record.  It then waits
// for the predecessor block's published (output offset, open-record state) -- the chain advances at the pace every
// thread works at, so nobody waits long --, publishes its own, and packs its staged bytes into the final arrays at
// that offset.  Whether the bytes before the block's first header are kept depends on the predecessor: they are
// staged anyway and skipped at pack time if the open record was a dropped one.  The two-pass form reads the file
// twice from DRAM and classifies every byte twice; this one once.  The packed arrays are allocated for the upper
// bound n bases (untouched pages cost nothing); a block's bytes in the groups it shares with its neighbours are
// kept in per-block fragments and assembled at the end.
template <class Alloc>
inline void fasta_parse_chained(const uint8_t *d, size_t n, int threads, std::vector<FastaRecord> &recs, uint64_t &total,
                                Alloc &&alloc, FastaPackTarget *pack) {
    recs.clear();
    total = 0;
    if (threads < 1) threads = 1;
    alloc((uint64_t)n);  // the callback allocates the pack target; n bases is an upper bound of the total
    size_t block = (size_t)512 << 10;
    if (const char *e = getenv("IPCR_FEEDER_BLOCK")) block = (size_t)std::max<long long>(1, atoll(e));  // tests
    const size_t nb = std::max<size_t>(1, (n + block - 1) / block);
    std::vector<size_t> cut(nb + 1, n);
    cut[0] = 0;
    for (size_t i = 1; i < nb; i++) cut[i] = std::max(cut[i - 1], fasta_snap(d, n, i * block));
    struct Blk {
        FastaPiece pc;
        uint64_t start = 0, end = 0;
        bool open_in = false, open_out = false;
        uint8_t head[32], tail[32];
        uint32_t head_len = 0, tail_len = 0;
        uint64_t tail_at = 0;
        uint64_t g0 = 0, ng = 0;  // the whole groups this block packed
        bool any_inval = false;   // ... and whether one of their bitmap words is non-zero
        std::atomic<int> ready{0};
    };
    std::vector<Blk> blk(nb);
    std::atomic<size_t> next{0};
    const bool timing = getenv("IPCR_FEEDER_TIMING") != nullptr;
    const auto t_p0 = std::chrono::steady_clock::now();
    auto worker = [&] {
        std::vector<uint8_t> stage;
        for (size_t i = next++; i < nb; i = next++) {
            Blk &b = blk[i];
            const size_t lo = cut[i], hi = cut[i + 1];
            if (stage.size() < hi - lo + 64) stage.resize(hi - lo + 64);
            // the one walk: kept bytes -> stage, counts -> b.pc
            uint8_t *out = stage.data();
            size_t w = 0;
            bool keep = true;  // before the first header: decided later (open_in)
            uint64_t *cur = &b.pc.head_kept;
            fasta_for_lines(d, n, lo, hi, [&](size_t a0, size_t b0) {
                if (b0 > a0 && d[a0] == '>') {
                    b.pc.hdrs.push_back({fasta_header_id(d + a0 + 1, b0 - a0 - 1), 0});
                    cur = &b.pc.hdrs.back().kept;
                    keep = !b.pc.hdrs.back().id.empty();
                    return;
                }
                fasta_trim(d, a0, b0);
                const size_t k = b0 - a0;
                *cur += k;  // counted for dropped records too, like the two-pass form (never added up)
                if (!keep) return;
                FastaPackStage::norm(out + w, d + a0, k);
                w += k;
            });
            if (i > 0) {
                while (!blk[i - 1].ready.load(std::memory_order_acquire)) std::this_thread::yield();
                b.start = blk[i - 1].end;
                b.open_in = blk[i - 1].open_out;
            }
            const uint8_t *src = stage.data() + (b.open_in ? 0 : b.pc.head_kept);
            const uint64_t kept = w - (b.open_in ? 0 : b.pc.head_kept);
            b.end = b.start + kept;
            b.open_out = b.pc.hdrs.empty() ? b.open_in : !b.pc.hdrs.back().id.empty();
            b.ready.store(1, std::memory_order_release);
            // pack at the final offset: fragment up to the next group boundary, whole groups, fragment behind them
            b.head_len = (uint32_t)std::min<uint64_t>(kept, (32 - b.start % 32) % 32);
            memcpy(b.head, src, b.head_len);
            const uint64_t ng = (kept - b.head_len) / 32;
            b.g0 = (b.start + b.head_len) / 32;
            b.ng = ng;
            if (ng) b.any_inval = pack_groups_staged(src + b.head_len, b.g0, ng, pack->codes, pack->inval) != 0;
            b.tail_at = b.start + b.head_len + 32 * ng;
            b.tail_len = (uint32_t)(b.end - b.tail_at);
            memcpy(b.tail, src + b.head_len + 32 * ng, b.tail_len);
        }
    };
    {
        std::vector<std::thread> th;
        const int nt = (int)std::min<size_t>((size_t)threads, nb);
        for (int t = 1; t < nt; t++) th.emplace_back(worker);
        worker();
        for (auto &x : th) x.join();
    }
    const auto t_p1 = std::chrono::steady_clock::now();
    // record table
    bool open = false;
    for (size_t i = 0; i < nb; i++) {
        if (open) recs.back().len += blk[i].pc.head_kept;
        uint64_t off = blk[i].start + (open ? blk[i].pc.head_kept : 0);
        for (auto &h : blk[i].pc.hdrs) {
            open = !h.id.empty();
            if (open) {
                FastaRecord r;
                r.id = h.id;
                r.off = off;
                r.len = h.kept;
                recs.push_back(std::move(r));
                off += h.kept;
            }
        }
    }
    total = blk[nb - 1].end;
    // groups shared by neighbouring blocks (and the ragged last group): assemble the fragments, pack
    struct Frag { uint64_t at; const uint8_t *src; uint32_t len; };
    std::vector<Frag> frags;
    for (size_t i = 0; i < nb; i++) {
        if (blk[i].head_len) frags.push_back(Frag{blk[i].start, blk[i].head, blk[i].head_len});
        if (blk[i].tail_len) frags.push_back(Frag{blk[i].tail_at, blk[i].tail, blk[i].tail_len});
    }
    std::vector<uint64_t> edge_g;
    for (const Frag &f : frags) edge_g.push_back(f.at / 32);
    edge_g.erase(std::unique(edge_g.begin(), edge_g.end()), edge_g.end());  // frags are in offset order
    std::vector<uint8_t> edge(edge_g.size() * 32 + 32, 0);
    for (const Frag &f : frags) {
        const size_t k = (size_t)(std::lower_bound(edge_g.begin(), edge_g.end(), f.at / 32) - edge_g.begin());
        memcpy(edge.data() + k * 32 + f.at % 32, f.src, f.len);
    }
    std::vector<std::pair<uint64_t, uint64_t>> dirty;
    for (size_t k = 0; k < edge_g.size(); k++)
        if (pack_groups_staged(edge.data() + k * 32, edge_g[k], 1, pack->codes, pack->inval)) dirty.push_back({edge_g[k], edge_g[k] + 1});
    pack_groups_host(nullptr, 0, (total + 31) / 32, packed_groups(total), pack->codes, pack->inval);  // padding groups
    for (size_t i = 0; i < nb; i++)
        if (blk[i].any_inval) dirty.push_back({blk[i].g0, blk[i].g0 + blk[i].ng});
    dirty.push_back({(total + 31) / 32, packed_groups(total)});
    std::sort(dirty.begin(), dirty.end());
    pack->dirty.swap(dirty);
    pack->dirty_known = true;
    if (timing)
        fprintf(stderr, "[feeder] one trip: %zu blocks, walk + pack %.3f s, tables + edge groups %.3f s\n", nb,
                std::chrono::duration<double>(t_p1 - t_p0).count(),
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t_p1).count());
}

template <class Alloc>
inline void fasta_parse(const uint8_t *d, size_t n, int threads, std::vector<FastaRecord> &recs, uint64_t &total,
                        Alloc &&alloc, FastaPackTarget *pack = nullptr) {
    if (pack && pack->only && !getenv("IPCR_FEEDER_TWO_PASS")) {
        fasta_parse_chained(d, n, threads, recs, total, alloc, pack);
        return;
    }
    recs.clear();
    total = 0;
    if (threads < 1) threads = 1;
    {
        size_t mt_min = (size_t)(8u << 20);  // below this one thread is faster
        if (const char *e = getenv("IPCR_FEEDER_MT_MIN")) mt_min = (size_t)atoll(e);  // tests: tiny multi-threaded files
        if (n < mt_min) threads = 1;
    }
    // range starts snapped to line starts
    std::vector<size_t> cut((size_t)threads + 1, n);
    cut[0] = 0;
    for (int t = 1; t < threads; t++) {
        size_t p = n / threads * t;
        if (p < cut[t - 1]) p = cut[t - 1];
        if (p > 0 && p < n && d[p - 1] != '\n') {
            const uint8_t *nl = static_cast<const uint8_t *>(memchr(d + p, '\n', n - p));
            p = nl ? (size_t)(nl - d) + 1 : n;
        }
        cut[t] = std::min(p, n);
    }
    std::vector<FastaPiece> pieces((size_t)threads);
    auto pass1 = [&](int t) { fasta_count_range(d, n, cut[t], cut[t + 1], pieces[t]); };
    const bool timing = getenv("IPCR_FEEDER_TIMING") != nullptr;
    const auto t_p0 = std::chrono::steady_clock::now();
    {
        std::vector<std::thread> th;
        for (int t = 1; t < threads; t++) th.emplace_back(pass1, t);
        pass1(0);
        for (auto &x : th) x.join();
    }
    const auto t_p1 = std::chrono::steady_clock::now();
    struct PhaseNote {  // stderr at scope exit: pass 1 (count) / everything after it (allocation, copy or staging, pack)
        bool on; std::chrono::steady_clock::time_point a, b;
        ~PhaseNote() {
            if (on)
                fprintf(stderr, "[feeder] pass 1 (count) %.3f s, allocation + pass 2 (normalise, pack) %.3f s\n",
                        std::chrono::duration<double>(b - a).count(),
                        std::chrono::duration<double>(std::chrono::steady_clock::now() - b).count());
        }
    } phase_note{timing, t_p0, t_p1};
    // merge: record table, and for every thread its output cursor and whether the record that
    // is open at its range start is kept (non-empty ID)
    std::vector<uint64_t> start_off((size_t)threads, 0);
    std::vector<char> start_open((size_t)threads, 0);
    bool open = false;  // a record with a non-empty ID is being filled
    for (int t = 0; t < threads; t++) {
        start_off[t] = total;
        start_open[t] = open;
        if (open) {
            recs.back().len += pieces[t].head_kept;
            total += pieces[t].head_kept;
        }
        for (auto &h : pieces[t].hdrs) {
            open = !h.id.empty();
            if (open) {
                FastaRecord r;
                r.id = h.id;
                r.off = total;
                r.len = h.kept;
                recs.push_back(std::move(r));
                total += h.kept;
            }
        }
    }
    uint8_t *out = alloc(total);  // the callback also allocates the pack target (it knows the total now)
    const bool do_pack = pack && pack->codes && pack->inval;
    std::vector<uint64_t> end_off((size_t)threads, 0);
    if (do_pack && pack->only) {
        // edge groups: those holding a range boundary that is not group-aligned (several boundaries may share one)
        std::vector<uint64_t> bound((size_t)threads + 1, total);
        for (int t = 0; t < threads; t++) bound[t] = start_off[t];
        std::vector<uint64_t> edge_g;
        for (int t = 0; t <= threads; t++)
            if (bound[t] % 32 != 0) edge_g.push_back(bound[t] / 32);
        std::sort(edge_g.begin(), edge_g.end());
        edge_g.erase(std::unique(edge_g.begin(), edge_g.end()), edge_g.end());
        std::vector<uint8_t> edge(edge_g.size() * 32 + 32, 0);
        auto slot = [&](uint64_t off) -> uint8_t * {
            const size_t i = (size_t)(std::lower_bound(edge_g.begin(), edge_g.end(), off / 32) - edge_g.begin());
            return edge.data() + i * 32;
        };
        auto pass2p = [&](int t) {
            FastaPackStage st;
            const uint64_t start = bound[t], end = bound[t + 1];
            st.w = start;
            st.A = std::min(end, (start + 31) / 32 * 32);
            st.B = std::max(st.A, end / 32 * 32);
            // a range that lies inside ONE unaligned group has no interior: everything goes to that group's slot
            if (start / 32 == end / 32 && start % 32 != 0) st.A = st.B = end;
            st.base = st.A;
            st.codes = pack->codes;
            st.inval = pack->inval;
            if (start % 32 != 0) { st.head = slot(start); st.head_base = start / 32 * 32; }
            if (end % 32 != 0) { st.tail = slot(end); st.tail_base = end / 32 * 32; }
            if (st.A == end && start % 32 != 0 && start / 32 == end / 32) { st.tail = st.head; st.tail_base = st.head_base; }
            st.buf.resize(FastaPackStage::kCap + 64);
            bool keep = start_open[t] != 0;
            size_t hi = 0;
            fasta_for_lines(d, n, cut[t], cut[t + 1], [&](size_t a, size_t b) {
                if (b > a && d[a] == '>') {
                    keep = !pieces[t].hdrs[hi++].id.empty();
                    return;
                }
                if (!keep) return;
                fasta_trim(d, a, b);
                st.put(d + a, b - a);
            });
            st.flush();
            end_off[t] = st.w;
        };
        {
            std::vector<std::thread> th;
            for (int t = 1; t < threads; t++) th.emplace_back(pass2p, t);
            pass2p(0);
            for (auto &x : th) x.join();
        }
        for (size_t i = 0; i < edge_g.size(); i++) pack_groups_staged(edge.data() + i * 32, edge_g[i], 1, pack->codes, pack->inval);
        // the padding groups behind the last base: 0x00 bytes = code 0, invalid
        pack_groups_host(nullptr, 0, (total + 31) / 32, packed_groups(total), pack->codes, pack->inval);
        return;
    }
    auto pass2 = [&](int t) {
        uint64_t w = start_off[t];
        // 32-base groups lying entirely inside this thread's output range are packed here, 64 KB at a time,
        // right after they were written (still in cache); the groups shared with a neighbour come last
        uint64_t g_next = (start_off[t] + 31) / 32;
        bool keep = start_open[t] != 0;
        size_t hi = 0;
        fasta_for_lines(d, n, cut[t], cut[t + 1], [&](size_t a, size_t b) {
            if (b > a && d[a] == '>') {
                keep = !pieces[t].hdrs[hi++].id.empty();
                return;
            }
            if (!keep) return;
            fasta_trim(d, a, b);
            const uint8_t *src = d + a;
            uint8_t *dst = out + w;
            const size_t k = b - a;
            for (size_t i = 0; i < k; i++) {  // normalize.go:8-10 (auto-vectorised)
                const uint8_t c = src[i];
                dst[i] = (uint8_t)(c - (((uint8_t)(c - 'a') < 26u) ? 32u : 0u));
            }
            w += k;
            if (do_pack && w / 32 >= g_next + 2048) {
                pack_groups_host(out, total, g_next, w / 32, pack->codes, pack->inval);
                g_next = w / 32;
            }
        });
        if (do_pack && w / 32 > g_next) pack_groups_host(out, total, g_next, w / 32, pack->codes, pack->inval);
        end_off[t] = w;
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < threads; t++) th.emplace_back(pass2, t);
        pass2(0);
        for (auto &x : th) x.join();
    }
    if (do_pack) {  // groups straddling two threads' ranges, the ragged tail and the padding groups
        const uint64_t groups = packed_groups(total);
        for (int t = 0; t < threads; t++) {
            const uint64_t g_lo = start_off[t] / 32, g_first = (start_off[t] + 31) / 32;
            if (g_first > g_lo) pack_groups_host(out, total, g_lo, std::min(g_first, groups), pack->codes, pack->inval);
            const uint64_t g_end = end_off[t] / 32;  // first group this thread did not finish
            const uint64_t g_stop = t + 1 < threads ? (start_off[t + 1] + 31) / 32 : groups;
            if (g_stop > g_end) pack_groups_host(out, total, g_end, std::min(g_stop, groups), pack->codes, pack->inval);
        }
    }
}

// whole-file read: mmap for plain files, zlib for gzip (open.go:39-49)
struct FastaBytes {
    const uint8_t *data = nullptr;
    size_t size = 0;
    void *map = nullptr;
    size_t map_len = 0;
    std::vector<uint8_t> owned;
    ~FastaBytes() {
        if (map) munmap(map, map_len);
    }
};

inline void fasta_read_file(const std::string &path, FastaBytes &fb) {
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) throw std::runtime_error("open " + path + ": " + strerror(errno));
    uint8_t sig[2] = {0, 0};
    const ssize_t got = ::pread(fd, sig, 2, 0);
    const bool gz = (got == 2 && sig[0] == 0x1f && sig[1] == 0x8b) ||
                    (path.size() >= 3 && path.compare(path.size() - 3, 3, ".gz") == 0);
    if (gz) {
        gzFile g = gzdopen(fd, "rb");
        if (!g) {
            ::close(fd);
            throw std::runtime_error("gzip open failed: " + path);
        }
        gzbuffer(g, 1 << 20);
        std::vector<uint8_t> buf(8u << 20);
        for (;;) {
            const int r = gzread(g, buf.data(), (unsigned)buf.size());
            if (r < 0) {
                gzclose(g);
                throw std::runtime_error("gzip read failed: " + path);
            }
            if (r == 0) break;
            fb.owned.insert(fb.owned.end(), buf.begin(), buf.begin() + r);
        }
        gzclose(g);
        fb.data = fb.owned.data();
        fb.size = fb.owned.size();
        return;
    }
    struct stat st;
    if (fstat(fd, &st) != 0) {
        ::close(fd);
        throw std::runtime_error("stat failed: " + path);
    }
    if (st.st_size == 0) {
        ::close(fd);
        return;
    }
    void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (m == MAP_FAILED) throw std::runtime_error("mmap failed: " + path);
    madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
    fb.map = m;
    fb.map_len = (size_t)st.st_size;
    fb.data = static_cast<const uint8_t *>(m);
    fb.size = (size_t)st.st_size;
}

}  // namespace ipcr
