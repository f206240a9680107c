/*
 * ipcr_gpu.h -- C ABI of the B200-native replacement for ipcr's core/engine scan.
 *
 * This is the drop-in boundary: a cgo shim binds these entry
 * points and implements the reference's pipeline.StreamingCompiledSimulator
 * (/root/reference/internal/pipeline/sim.go:11-39) on top of them, so that the
 * one-line swap at internal/appcore/core.go:108 (`sim := engine.New(...)`)
 * replaces the CPU engine with the GPU engine and nothing else changes.
 *
 * Conventions
 *   - plain C types only; no exceptions cross the ABI.
 *   - every fallible call returns 0 on success, non-zero on failure and writes a
 *     NUL-terminated message into (err, errlen) when err != NULL.
 *   - there is NO CPU fallback: without a usable sm_100 device every call fails.
 *   - a panel is immutable and may be shared by any number of host threads; a
 *     scratch (stream + device/pinned buffers) belongs to one thread at a time
 *     (reference protocol: internal/pipeline/pipeline.go:52-69).
 *   - coordinates are record/chunk-local, 0-based, half-open [start,end)
 *     exactly as engine.Product (core/engine/product.go:4-35).
 */
#ifndef IPCR_GPU_H
#define IPCR_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IPCR_GPU_ABI_VERSION 2
#define IPCR_GPU_MAX_PRIMER_LEN 64 /* mismatch positions travel as 64-bit masks */

/* replaces engine.Config (core/engine/engine.go:10-19) */
typedef struct ipcr_cfg {
    int32_t max_mm;          /* --mismatches; < 0 is rejected (reference is inconsistent there) */
    int32_t terminal_window; /* 3' protected bases; <= 0 disables */
    int64_t min_len;         /* 0 = unbounded */
    int64_t max_len;         /* 0 = unbounded */
    int64_t hit_cap;         /* per orientation per record; 0 = unlimited */
    int32_t seed_len;        /* --seed-length: only decides which orientations the reference would
                                have seeded (hit-cap / emission-order semantics); < 0 = no seeds */
    uint8_t circular;
    uint8_t need_sites;      /* accepted; FwdSite/RevSite are slices of the caller's seq */
    uint8_t reserved[2];
} ipcr_cfg;

/* replaces primer.Pair (core/primer/pair.go:4-10); strings need not be NUL-terminated */
typedef struct ipcr_pair {
    const char *fwd;
    uint32_t fwd_len;
    const char *rev;
    uint32_t rev_len;
    int64_t min_product; /* 0 = use cfg.min_len (core/engine/engine.go:113-120) */
    int64_t max_product;
} ipcr_pair;

/* one FASTA record or chunk (fasta.Record.Seq, core/fasta/reader.go:9-12) */
typedef struct ipcr_seq {
    const uint8_t *seq;
    uint64_t len;
} ipcr_seq;

/* A piece of a record placed on one device.  Whole records use rec_off = 0 and
 * own_len = len.  Multi-GPU shards pass len = own_len + halo (halo >= longest
 * primer - 1) so that a site starting in [0, own_len) is fully visible. */
typedef struct ipcr_segment {
    const uint8_t *seq; /* host pointer (upload) or device pointer (adopt) */
    uint64_t len;
    uint64_t own_len;
    uint64_t rec_off; /* record-local coordinate of seq[0] */
    uint32_t rec_idx;
    uint32_t reserved;
} ipcr_segment;

/* replaces the printable part of engine.Product (core/engine/product.go:4-35).
 * Mismatch positions are bit masks over primer 5'->3' indices: bit j of
 * fwd_mm_mask = FwdMismatchIdx contains j; rev_mm_mask is ALREADY flipped
 * (core/engine/engine.go:127-136).  Use ipcr_gpu_product_mismatch_idx to get
 * the lists in the reference's order. */
typedef struct ipcr_product {
    int64_t start, end, length;
    uint64_t fwd_mm_mask, rev_mm_mask;
    uint32_t pair_idx; /* index into the panel's pairs: ExperimentID = pairs[pair_idx].ID */
    uint32_t rec_idx;  /* index of the record within the call (0 for ipcr_gpu_scan) */
    uint16_t fwd_mm, rev_mm;
    uint8_t type; /* 0 = "forward", 1 = "revcomp" */
    uint8_t reserved[11];
} ipcr_product;

/* one verified primer site (stage-1 output; what multi-GPU shards exchange) */
typedef struct ipcr_site {
    uint64_t pos;     /* record-local start */
    uint64_t mm_mask; /* mismatch mask in orientation coordinates */
    uint32_t rec_idx;
    uint16_t orient;  /* unique-orientation id inside the panel */
    uint8_t mm;
    uint8_t flags;    /* IPCR_SITE_* */
} ipcr_site;
#define IPCR_SITE_TWVIOL 1u /* mismatch inside the protected window (kept only for hit-cap parity) */
#define IPCR_SITE_HALO 2u   /* the reference's seed window holds a non-ACGT byte (found by its halo pass) */

typedef struct ipcr_record_info {
    uint64_t len;
    uint8_t has_reset_byte; /* any byte outside ACGTacgt (core/engine/ac.go:141-148) */
    uint8_t reserved[7];
} ipcr_record_info;

/* device-side timings of the last scan on a scratch, milliseconds (CUDA events
 * on the scratch's stream) */
typedef struct ipcr_timing {
    float h2d_ms, pack_ms, filter_ms, verify_ms, pair_ms, d2h_ms, total_ms;
    uint64_t candidates, sites, products;
    uint32_t kernel_launches; /* launches issued by the last scan */
    uint32_t used_jit;        /* filter that ran: 0 table-driven, 1 NVRTC-specialised, 2 seed index */
    uint32_t plan_variant;    /* 0 = protected window enforced in the filter, 1 = rc orientations raw */
    uint32_t reserved;
} ipcr_timing;

typedef struct ipcr_panel ipcr_panel;
typedef struct ipcr_scratch ipcr_scratch;
typedef struct ipcr_genome ipcr_genome;

/* ---- process / device ---------------------------------------------------- */
int ipcr_gpu_abi_version(void);
/* number of visible CUDA devices, or 0 (never fails; no device => 0) */
int ipcr_gpu_device_count(void);
/* Checks the devices (sm_100) and sets how many of them ONE ipcr_gpu_scan_batch /
 * ipcr_gpu_scan_fasta call may spread over: 0 = all visible devices, 1 = only the device of
 * the scratch, n = the scratch's device plus n - 1 others.  Large inputs (>= 256 Mbp) are then
 * cut into one slice per device (64-base halo), every device pulls its slice from the caller's
 * host memory over its own PCIe link and finds its primer sites, the site blocks meet on the
 * scratch's device over NVLink peer copies and are paired once -- same products, same order.
 * Replaces the chunk worker pool of the Go host for the GPU engine (internal/pipeline/pipeline.go:
 * 60-125, internal/runutil/runutil.go:20-59).  Does not change the calling thread's device.
 * The device-side visitors (probe / nested) need the records on one device: a call that was
 * spread over several devices leaves none resident. */
int ipcr_gpu_init(int n_devices_or_0, char *err, size_t errlen);
/* binds the calling thread to `device` (scratches are created on the thread's current device) */
int ipcr_gpu_set_device(int device, char *err, size_t errlen);
/* the device budget set by ipcr_gpu_init (1 before the first call) */
int ipcr_gpu_num_devices(void);

/* ---- panel: engine.CompilePanel (core/engine/compiled.go:96-136) ---------- */
ipcr_panel *ipcr_gpu_panel_create(const ipcr_cfg *cfg, const ipcr_pair *pairs, size_t n_pairs,
                                  char *err, size_t errlen);
void ipcr_gpu_panel_destroy(ipcr_panel *p);
size_t ipcr_gpu_panel_num_pairs(const ipcr_panel *p);
size_t ipcr_gpu_panel_num_orientations(const ipcr_panel *p); /* unique strings scanned */
/* would the reference have seeded this orientation? which in 'A','B','a','b'
 * (compiledHas, core/engine/compiled.go:35-37) */
int ipcr_gpu_panel_have(const ipcr_panel *p, size_t pair_idx, int which);
/* which candidate filter serves the panel: 0 = table-driven bit-parallel kernel only,
 * 1 = automatic (default: the seed index when it is selective, else bit-parallel, specialised
 * by NVRTC once the work pays for the compile), 2 = force the NVRTC-specialised bit-parallel
 * kernel, 3 = force the seed index */
void ipcr_gpu_panel_set_jit(ipcr_panel *p, int mode);
/* the same switch under its proper name */
#define IPCR_FILTER_BITPARALLEL_TABLE 0
#define IPCR_FILTER_AUTO 1
#define IPCR_FILTER_BITPARALLEL_JIT 2
#define IPCR_FILTER_SEED_INDEX 3
void ipcr_gpu_panel_set_filter(ipcr_panel *p, int filter);

/* ---- scratch: engine.NewSimulationScratch (core/engine/hit_collect.go:19-34) */
ipcr_scratch *ipcr_gpu_scratch_create(ipcr_panel *p, char *err, size_t errlen);
void ipcr_gpu_scratch_destroy(ipcr_scratch *s);

/* ---- the hot path: engine.ForEachCompiledProduct (compiled.go:162-267) ---- */
/* Products come back in the reference's emission order.  *out points into
 * scratch-owned pinned memory, valid until the next call on that scratch. */
int ipcr_gpu_scan(ipcr_panel *p, ipcr_scratch *s, const uint8_t *seq, uint64_t len,
                  const ipcr_product **out, uint64_t *n_out, char *err, size_t errlen);
/* many records in one launch chain; products are record-major */
int ipcr_gpu_scan_batch(ipcr_panel *p, ipcr_scratch *s, const ipcr_seq *recs, size_t n_recs,
                        const ipcr_product **out, uint64_t *n_out, char *err, size_t errlen);

/* same, but the records already sit in DEVICE memory (segment seq pointers are device
 * pointers, ascending, inside one allocation): pack -> filter -> verify -> pair.
 * The call scans the span from the first to the last segment as one buffer; bytes between
 * segments belong to no record (their sites are discarded) but are still scanned, so pass
 * segments that lie (nearly) back to back. */
int ipcr_gpu_scan_device(ipcr_panel *p, ipcr_scratch *s, const ipcr_segment *segs, size_t n_segs,
                         size_t n_recs, const ipcr_product **out, uint64_t *n_out, char *err,
                         size_t errlen);

/* ---- resident genomes (pack once, scan with many panels) ------------------ */
/* host segments -> pinned staging -> H2D -> pack */
ipcr_genome *ipcr_gpu_genome_upload(ipcr_scratch *s, const ipcr_segment *segs, size_t n_segs,
                                    size_t n_recs, char *err, size_t errlen);
/* ASCII already in device memory (segment seq pointers are device pointers that
 * must stay valid for the genome's lifetime); only the pack kernel runs */
ipcr_genome *ipcr_gpu_genome_adopt(ipcr_scratch *s, const ipcr_segment *segs, size_t n_segs,
                                   size_t n_recs, char *err, size_t errlen);
void ipcr_gpu_genome_destroy(ipcr_genome *g);
uint64_t ipcr_gpu_genome_bases(const ipcr_genome *g);
/* per-record info for this genome's pieces (len = max rec_off+len seen) */
int ipcr_gpu_genome_records(const ipcr_genome *g, ipcr_record_info *out, size_t n_recs);
int ipcr_gpu_scan_genome(ipcr_panel *p, ipcr_scratch *s, const ipcr_genome *g,
                         const ipcr_product **out, uint64_t *n_out, char *err, size_t errlen);

/* ---- split phase (multi-GPU: find on every shard, gather, pair once) ------ */
/* stage 1: *sites is a HOST pointer (scratch-owned pinned memory) */
int ipcr_gpu_find_sites(ipcr_panel *p, ipcr_scratch *s, const ipcr_genome *g,
                        const ipcr_site **sites, uint64_t *n_sites, char *err, size_t errlen);
/* stage 1 on records/shards that already sit in device memory (pack included; scratch
 * buffers are reused, nothing is allocated per call once warm).  recs_out[n_recs]
 * receives this shard's view of each record (len seen, reset-byte flag). */
int ipcr_gpu_find_sites_device(ipcr_panel *p, ipcr_scratch *s, const ipcr_segment *segs, size_t n_segs,
                               size_t n_recs, int force_raw, const ipcr_site **sites, uint64_t *n_sites,
                               ipcr_record_info *recs_out, char *err, size_t errlen);
/* force_raw: hit-cap parity needs rc orientations to keep matches that violate the 3' window
 * whenever the WHOLE record holds a non-ACGTacgt byte (core/engine/compiled.go:185-190).  A shard
 * only sees its own bytes: pass 0 to decide from the shard (then, if ipcr_timing.plan_variant == 0
 * and another shard of the same record reports a reset byte while hit_cap, max_mm and
 * terminal_window are all > 0, call again with 1), or pass 1 to be safe at ~20 % more filter work. */
/* stage 2: host site list (any order, any shards) + whole-record info */
int ipcr_gpu_pair_sites(ipcr_panel *p, ipcr_scratch *s, const ipcr_site *sites, uint64_t n_sites,
                        const ipcr_record_info *recs, size_t n_recs, const ipcr_product **out,
                        uint64_t *n_out, char *err, size_t errlen);

/* ---- split phase, device-resident (multi-process multi-GPU: stage 1 per rank, ONE all-gather
 * of fixed-size site blocks, stage 2 on every rank -- the host never touches a site) ---------- */
/* A site block is [64-byte control | n_recs flag bytes, padded to 16 | site_cap sites]. */
uint64_t ipcr_gpu_site_block_bytes(size_t n_recs, uint64_t site_cap);
/* sites per block this scratch currently asks for (grows when a call reported lost sites) */
uint64_t ipcr_gpu_scratch_site_cap(const ipcr_scratch *s);
/* the scratch's CUDA stream (cudaStream_t): order foreign work, e.g. the collective, against it */
void *ipcr_gpu_scratch_stream(ipcr_scratch *s);
/* Stage 1 of device-resident segments (as ipcr_gpu_find_sites_device) into d_block, a device
 * buffer of ipcr_gpu_site_block_bytes(n_recs, site_cap) bytes.  ASYNCHRONOUS: returns once the
 * work is queued on the scratch's stream.  Seed-index panels only (the default filter). */
int ipcr_gpu_find_sites_block(ipcr_panel *p, ipcr_scratch *s, const ipcr_segment *segs, size_t n_segs,
                              size_t n_recs, int force_raw, void *d_block, uint64_t site_cap, char *err,
                              size_t errlen);
/* The same from a PACKED device buffer (ipcr_gpu_pack_host output copied to the device as is): segments
 * are base offsets into it.  has_reset_byte: the caller saw the bytes (the packed form cannot tell a
 * lowercase acgt from N; inputs normalised by the reader have no lowercase). */
typedef struct ipcr_packed_segment {
    uint64_t off, len, own_len, rec_off;
    uint32_t rec_idx;
    uint32_t has_reset_byte;
} ipcr_packed_segment;
int ipcr_gpu_find_sites_block_packed(ipcr_panel *p, ipcr_scratch *s, const uint32_t *d_codes, const uint32_t *d_inval,
                                     uint64_t n_bases, const ipcr_packed_segment *segs, size_t n_segs, size_t n_recs,
                                     int force_raw, void *d_block, uint64_t site_cap, char *err, size_t errlen);
/* Stage 2 over n_blocks blocks lying back to back in device memory (what the all-gather left),
 * queued behind whatever the scratch's stream holds; waits once, for the products.
 * rec_lens[n_recs] = whole-record lengths.  Returns 0 = done, 1 = error, 2 = run stage 1 again on
 * EVERY rank and call again: a capacity was too small (ipcr_gpu_scratch_site_cap may have grown) or
 * a record holds a reset byte while some block was scanned with the plain plan variant (hit-cap
 * parity, core/engine/compiled.go:185-190); every rank sees the same blocks and so takes the same
 * decision without talking to the others. */
int ipcr_gpu_pair_blocks(ipcr_panel *p, ipcr_scratch *s, const void *d_blocks, uint32_t n_blocks,
                         uint64_t site_cap, const uint64_t *rec_lens, size_t n_recs,
                         const ipcr_product **out, uint64_t *n_out, char *err, size_t errlen);

/* ---- peer exchange: the same split phase with the collective FUSED into the pairing kernel ------------
 * One process per GPU.  Every rank owns a double-buffered site block in its own device memory and exports it
 * through CUDA IPC; the other ranks map it (NVLink peer access).  A step is
 *   d_block = ipcr_gpu_peer_begin_step(s);  ipcr_gpu_find_sites_block[_packed](.., d_block, site_cap);
 *   ipcr_gpu_peer_publish(s);  ipcr_gpu_pair_peers(..)
 * on EVERY rank: stage 1 ends by publishing the step number in the block's control word (st.release.sys), the
 * pairing kernel of every rank waits for all step numbers (ld.acquire.sys) and reads the peers' sites straight
 * out of their memories -- no collective library, no copy, no host between stage 1 and stage 2.  Double buffering
 * makes the exchange safe without a barrier: a rank can only overwrite a block two steps later, which needs its
 * own pairing of the step in between, which needs every peer's NEXT publication, which follows their pairing of
 * this step.  A rank that never publishes makes the others fail after 4 s instead of hanging.
 * handle_out / handles: IPCR_GPU_PEER_HANDLE_BYTES per rank, exchanged by the caller (any transport). */
#define IPCR_GPU_PEER_HANDLE_BYTES 64
int ipcr_gpu_peer_create(ipcr_scratch *s, size_t n_recs, uint64_t site_cap, void *handle_out, char *err, size_t errlen);
int ipcr_gpu_peer_open(ipcr_scratch *s, const void *handles, uint32_t n_ranks, uint32_t my_rank, char *err, size_t errlen);
void *ipcr_gpu_peer_begin_step(ipcr_scratch *s);
int ipcr_gpu_peer_publish(ipcr_scratch *s, char *err, size_t errlen);
/* like ipcr_gpu_pair_blocks (0 done / 1 error / 2 run the step again on every rank; when
 * ipcr_gpu_scratch_site_cap grew, create + exchange + open again first) */
int ipcr_gpu_pair_peers(ipcr_panel *p, ipcr_scratch *s, const uint64_t *rec_lens, size_t n_recs,
                        const ipcr_product **out, uint64_t *n_out, char *err, size_t errlen);
void ipcr_gpu_peer_close(ipcr_scratch *s);

/* ---- native FASTA feeder (replaces core/fasta for the hot path) ------ */
/* Reads and normalises a FASTA file exactly like the reference reader (header ID = first
 * token, TrimSpace per line, a-z upper-cased, gzip by magic/suffix; core/fasta/scan.go,
 * normalize.go, stream.go:125-131, open.go) into ONE host buffer, multi-threaded.
 * pin != 0 asks for page-locked memory (faster H2D, slower to allocate). */
typedef struct ipcr_fasta ipcr_fasta;
#define IPCR_FASTA_PIN 1  /* page-locked buffers: faster H2D, but locking costs ~0.5 s per GB -- for handles that are
                             scanned many times; a one-shot FASTA -> TSV run is faster without */
#define IPCR_FASTA_PACK 2 /* also keep the PACKED form (2-bit codes + hard-mismatch bitmap, 0.375 B/bp, pinned):
                             ipcr_gpu_scan_fasta then ships that over PCIe, skips the pack kernel and verifies on
                             the codes; the device-side visitors need the ASCII path */
#define IPCR_FASTA_PACK_ONLY 4 /* ipcr_gpu_fasta_open: ONLY the packed form -- the feeder never writes the 1 B/bp ASCII
                             buffer (a third of its memory traffic and all of its page faults).  For runs that hand
                             the handle to ipcr_gpu_scan_fasta and read ids / lengths only.  ipcr_gpu_fasta_record's
                             `seq` is then rebuilt from the codes on first use, every non-ACGT byte as 'N' (the
                             same thing to the engine; NOT the file's IUPAC letters) */
/* pin: IPCR_FASTA_* flags */
ipcr_fasta *ipcr_gpu_fasta_open(const char *path, int threads, int pin, char *err, size_t errlen);
/* the same handle from records already in host memory (copied, a-z upper-cased like the reader does) */
ipcr_fasta *ipcr_gpu_fasta_from_memory(const ipcr_seq *recs, const char *const *ids, size_t n_recs, int threads,
                                       int flags, char *err, size_t errlen);
int ipcr_gpu_fasta_is_packed(const ipcr_fasta *f);
/* the packed form of the handle's whole sequence buffer (records back to back): 2 * n_groups code words,
 * n_groups bitmap words; owned by the handle */
int ipcr_gpu_fasta_packed(const ipcr_fasta *f, const uint32_t **codes, const uint32_t **inval, uint64_t *n_groups);
/* the handle's bitmap as the run list ipcr_gpu_scan_fasta ships (see ipcr_gpu_packed_runs); returns 1 when the handle
 * is not packed or holds too many runs for a list (*n_runs = 0 then) */
int ipcr_gpu_fasta_runs(const ipcr_fasta *f, const uint64_t **runs, uint64_t *n_runs);
/* the host packer the feeder applies (bit-for-bit what the device pack kernel writes):
 * codes[*code_words], inval[*inval_words] for n_bases bases, padding included */
void ipcr_gpu_packed_words(uint64_t n_bases, uint64_t *code_words, uint64_t *inval_words);
int ipcr_gpu_pack_host(const uint8_t *ascii, uint64_t n_bases, uint32_t *codes, uint32_t *inval, int threads);
/* The bitmap of a real genome is nearly empty (few, long N runs): as a list of runs [runs[2i], runs[2i+1]) of
 * invalid bases (padding included) it is a few KB instead of 0.125 B/bp.  Returns the number of runs written, or
 * -1 when there are more than max_runs (ship the bitmap then).  ipcr_gpu_scan_fasta does this by itself. */
int64_t ipcr_gpu_packed_runs(const uint32_t *inval, uint64_t n_bases, uint64_t *runs, uint64_t max_runs, int threads);
/* rebuilds the bitmap words of n_bases bases in DEVICE memory from such a run list; queued on the scratch's stream
 * (call it before ipcr_gpu_find_sites_block_packed of the same step) */
int ipcr_gpu_inval_from_runs(ipcr_scratch *s, uint32_t *d_inval, uint64_t n_bases, const uint64_t *runs, uint64_t n_runs,
                             char *err, size_t errlen);
void ipcr_gpu_fasta_close(ipcr_fasta *f);
size_t ipcr_gpu_fasta_num_records(const ipcr_fasta *f);
uint64_t ipcr_gpu_fasta_total_bases(const ipcr_fasta *f);
int ipcr_gpu_fasta_record(const ipcr_fasta *f, size_t i, const char **id, const uint8_t **seq, uint64_t *len);
/* ForEachCompiledProduct over every record of the file (rec_idx = record number in file
 * order); large files stream to the device slab by slab under the scan.  *out is owned by
 * the fasta handle and valid until the next scan on it. */
int ipcr_gpu_scan_fasta(ipcr_panel *p, ipcr_scratch *s, ipcr_fasta *f, const ipcr_product **out,
                        uint64_t *n_out, char *err, size_t errlen);

/* ---- visitors on the device ---------------------------------------- */
/* These work on the records of the LAST ipcr_gpu_scan / _scan_batch / _scan_device call on the
 * scratch (their bytes are still resident in HBM); `products` are products of that call. */
typedef struct ipcr_probe_hit {
    int64_t pos;    /* 0-based in the amplicon */
    uint16_t mm;
    uint8_t found;
    uint8_t strand; /* '+' or '-' */
    uint8_t reserved[4];
} ipcr_probe_hit;
/* probe.AnnotateAmplicon / oligo.BestHit (core/oligo/oligo.go:19-77) for every product.
 * `probe` must already be normalised uppercase IUPAC (oligo.Validate). */
int ipcr_gpu_probe_annotate(ipcr_scratch *s, const ipcr_product *products, uint64_t n, const char *probe,
                            uint32_t probe_len, int32_t max_mm, ipcr_probe_hit *out, char *err, size_t errlen);
/* Product.Seq of every product (internal/pipeline/pipeline.go:81-88): amplicon i occupies
 * out[offsets[i], offsets[i+1]); offsets has n + 1 entries.  Returns the bytes needed in *need
 * (call with out == NULL first). */
int ipcr_gpu_amplicons(ipcr_scratch *s, const ipcr_product *products, uint64_t n, uint8_t *out, uint64_t cap,
                       uint64_t *offsets, uint64_t *need, char *err, size_t errlen);
/* visitors.Nested (internal/visitors/nested.go:17-66): scan every amplicon with the INNER
 * panel (all amplicons in one batched launch chain on inner_scratch) and pick the best inner
 * product per amplicon: fewest total mismatches, then longest, leftmost start, end, pair ID
 * (inner_ids[pair] supplies the IDs for the last tie-break; may be NULL).  best[i].length == -1
 * when amplicon i has no inner product. */
int ipcr_gpu_nested(ipcr_scratch *outer_scratch, const ipcr_product *products, uint64_t n,
                    ipcr_panel *inner_panel, ipcr_scratch *inner_scratch, const char *const *inner_ids,
                    ipcr_product *best, char *err, size_t errlen);

/* ---- --sort on the device (internal/common/sort.go:34-74) ----------------
 * Sorts n products (host array, IN PLACE) into common.LessProduct order:
 *   (SourceFile, base SequenceID) -> rec_rank[rec_idx] ; Start + off ; End + off ; Length ; Type ;
 *   ExperimentID -> pair_rank[pair_idx] ; FwdMM ; RevMM
 * The strings enter as ranks the caller computes once (equal strings, equal rank); rec_off[rec_idx] is the
 * chunk offset the collector adds (NULL = 0; it is NOT added to the products).  The sort is stable.
 * tie_out (optional, n bytes): tie_out[i] = 1 when sorted product i equals product i - 1 on every key above --
 * only possible between pairs that share one experiment ID or duplicated records; LessProduct then goes on
 * to the mismatch lists, SequenceID and Seq, which the caller settles inside those (tiny) runs. */
int ipcr_gpu_sort_products(ipcr_scratch *s, ipcr_product *products, uint64_t n, const uint32_t *rec_rank,
                           const int64_t *rec_off, size_t n_recs, const uint32_t *pair_rank, size_t n_pairs,
                           uint8_t *tie_out, char *err, size_t errlen);

/* ---- helpers ---------------------------------------------------------------- */
/* which: 0 = FwdMismatchIdx (ascending), 1 = RevMismatchIdx (reference order:
 * flipped, hence descending).  Returns the count written (<= cap). */
int ipcr_gpu_product_mismatch_idx(const ipcr_product *pr, int which, int32_t *out, int cap);
int ipcr_gpu_last_timing(const ipcr_scratch *s, ipcr_timing *out);
/* fills `n` bytes of DEVICE memory with a deterministic uniform ACGT sequence
 * (splitmix64 of seed and absolute position `first_pos + i`); bench/test input */
int ipcr_gpu_synth_fill(ipcr_scratch *s, void *d_dst, uint64_t n, uint64_t first_pos, uint64_t seed,
                        char *err, size_t errlen);

/* ---- debug / test taps (stable, but not part of the drop-in surface) -------- */
/* copies the packed planes of a genome to host: words[4 * units], unit order
 * (tile, row, lane); returns W (rows per stripe) via *stripe_rows */
int ipcr_gpu_debug_planes(const ipcr_genome *g, uint32_t *words, uint64_t max_words,
                          uint64_t *n_words, uint32_t *stripe_rows, uint64_t *n_tiles);
/* candidates of the last find_sites/scan on this scratch: (orient << 44 | global start) */
/* the 2-bit codes (2 words per 32 bases) and hard-mismatch bitmap (1 word per 32 bases) of a genome
 * packed for the seed-index filter; *n_groups = 0 if it holds bit-planes instead */
int ipcr_gpu_debug_codes(const ipcr_genome *g, uint32_t *codes, uint32_t *inval, uint64_t max_groups,
                         uint64_t *n_groups);
int ipcr_gpu_debug_candidates(const ipcr_scratch *s, uint64_t *out, uint64_t cap, uint64_t *n);
/* CPU evaluation of the panel's filter plan at one window (no GPU): 1 if the
 * plan would flag `start` as a candidate for orientation `orient` */
int ipcr_gpu_debug_plan_eval(const ipcr_panel *p, uint32_t orient, const uint8_t *seq, uint64_t len,
                             int64_t start, int raw_mode);
/* orientation string / type ('A' forward-type, 'a' rc-type) of a unique orientation */
/* host evaluation of the panel's seed index on one sequence: the (orientation << 44 | start)
 * candidates the seed filter hands to verify; info[8] = {S, entries, table_bits, bloom used,
 * bloom fill ppm, bitmap bits per key | anchor stride << 8, S1 (first-level key length), distinct first-level keys} */
int ipcr_gpu_debug_seed_eval(ipcr_panel *p, int raw_mode, const uint8_t *seq, uint64_t len, uint64_t *out,
                             uint64_t cap, uint64_t *n, int32_t *info);
int ipcr_gpu_debug_orientation(const ipcr_panel *p, uint32_t orient, char *buf, size_t buflen,
                               int *type);
/* NVRTC source of the specialised kernel for this panel (for inspection) */
int ipcr_gpu_debug_jit_source(ipcr_panel *p, int raw_mode, char *buf, size_t buflen, size_t *needed);
/* runs NVRTC on that source (works without a GPU) and copies the sm_100a cubin out */
int ipcr_gpu_debug_jit_compile(ipcr_panel *p, int raw_mode, char *cubin, size_t cap, size_t *cubin_size,
                               char *err, size_t errlen);

#ifdef __cplusplus
}
#endif
#endif
