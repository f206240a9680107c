// ipcr_gpu_cli.cpp -- C++ host for the B200 scan: the hot flow of ipcr / ipcr-probe /
// ipcr-nested / ipcr-multiplex (FASTA on disk -> TSV) on top of the C ABI only
// (include/ipcr_gpu.h).  It mirrors what the Go host does around the engine:
//   flags            internal/clibase/common.go:61-174 (the subset that reaches the hot path)
//   primer files     core/primer/loader.go:11-59, validate.go:12-40
//   self pairs       internal/common/primers.go:11-74
//   chunk policy     internal/appcore/core.go:46-98, internal/runutil/runutil.go:10-59
//   pipeline         internal/pipeline/pipeline.go:33-194 (chunk ids, global coords, LRU dedupe)
//   visitors         internal/visitors/{pass,probe,nested}.go (device-side through the ABI)
//   writers / --sort internal/writers/product.go:85-153, internal/common/sort.go:34-74,
//                    internal/output/rows.go:22-30, probeoutput/text.go, nestedoutput/text.go
// Everything that matches, annotates or pairs runs on the GPU; this file is policy, ordering
// and formatting.  Exit codes follow appcore.Run (0 ok, 2 usage, 3 runtime, --no-match-exit-code).
#include <algorithm>
#include <chrono>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <list>
#include <map>
#include <sstream>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "../../../include/ipcr_gpu.h"

namespace {

struct Pair {
    std::string id, fwd, rev;
    long long min_product = 0, max_product = 0;
};

struct Product {
    std::string experiment_id, sequence_id, type, seq, source_file;
    long long start = 0, end = 0, length = 0;
    int fwd_mm = 0, rev_mm = 0;
    std::vector<int> fwd_idx, rev_idx;
    // visitor overlays
    bool probe_found = false;
    std::string probe_strand, probe_site;
    long long probe_pos = 0;
    int probe_mm = 0;
    bool inner_found = false;
    std::string inner_id, inner_type;
    long long inner_start = 0, inner_end = 0, inner_length = 0;
    int inner_fwd_mm = 0, inner_rev_mm = 0;
};

struct Options {
    std::string mode = "ipcr";
    std::string primers, fwd, rev;
    std::vector<std::string> files;
    int mismatches = 0, terminal_window = 3, seed_length = 12;
    long long min_len = 0, max_len = 2000, hit_cap = 10000, chunk_size = 0, dedupe_cap = 200000;
    bool self = true, circular = false, sort = false, header = true;
    int no_match_exit = 0;
    std::string probe, probe_name = "probe";
    int probe_max_mm = 0;
    bool require_probe = true;
    std::string inner_fwd, inner_rev, inner_primers;
    bool require_inner = false;
};

[[noreturn]] void die(int code, const std::string &m) {
    fprintf(stderr, "error: %s\n", m.c_str());
    exit(code);
}

std::string normalize(const std::string &raw) {  // primer.Normalize / oligo.Normalize
    std::string o;
    for (unsigned char c : raw)
        if (!isspace(c) && c != '\'' && c != '"') o.push_back((char)toupper(c));
    return o;
}

std::string validate(const std::string &raw, const char *what) {  // primer.Validate (validate.go:26-40)
    std::string s = normalize(raw);
    if (s.empty()) die(2, std::string(what) + ": empty primer");
    for (size_t i = 0; i < s.size(); i++)
        if (!strchr("ACGTRYSWKMBDHVN", s[i]))
            die(2, std::string(what) + ": invalid primer base '" + s[i] + "' at position " + std::to_string(i + 1) +
                       "; allowed: A C G T R Y S W K M B D H V N");
    return s;
}

std::vector<Pair> load_tsv(const std::string &path) {  // primer.LoadTSV (loader.go:11-59)
    std::ifstream in(path);
    if (!in) die(2, "open " + path + ": no such file or directory");
    std::vector<Pair> out;
    std::string line;
    int ln = 0;
    while (std::getline(in, line)) {
        ln++;
        std::istringstream ss(line);
        std::vector<std::string> f;
        std::string tok;
        while (ss >> tok) f.push_back(tok);
        if (f.empty() || f[0][0] == '#') continue;
        if (f.size() < 3 || f.size() > 5) die(2, path + ":" + std::to_string(ln) + " bad field count");
        Pair p;
        p.id = f[0];
        p.fwd = validate(f[1], (path + ":" + std::to_string(ln) + " forward primer").c_str());
        p.rev = validate(f[2], (path + ":" + std::to_string(ln) + " reverse primer").c_str());
        if (f.size() >= 4) p.min_product = atoll(f[3].c_str());
        if (f.size() == 5) p.max_product = atoll(f[4].c_str());
        out.push_back(p);
    }
    return out;
}

std::string upper(std::string s) {
    for (auto &c : s) c = (char)toupper((unsigned char)c);
    return s;
}

std::vector<Pair> add_self_pairs(const std::vector<Pair> &in) {  // common.AddSelfPairs
    std::vector<Pair> out = in;
    for (const Pair &p : in) {
        if (!p.fwd.empty()) out.push_back(Pair{p.id + "+A:self", upper(p.fwd), upper(p.fwd), 0, 0});
        if (!p.rev.empty()) out.push_back(Pair{p.id + "+B:self", upper(p.rev), upper(p.rev), 0, 0});
    }
    return out;
}

std::vector<Pair> add_self_pairs_unique(const std::vector<Pair> &in) {  // common.AddSelfPairsUnique
    std::vector<Pair> out = in;
    std::map<std::string, bool> sa, sb;
    for (const Pair &p : in) {
        std::string f = upper(p.fwd), r = upper(p.rev);
        if (!f.empty() && !sa.count(f)) { sa[f] = true; out.push_back(Pair{p.id + "+A:self", f, f, 0, 0}); }
        if (!r.empty() && !sb.count(r)) { sb[r] = true; out.push_back(Pair{p.id + "+B:self", r, r, 0, 0}); }
    }
    return out;
}

bool split_chunk_suffix(const std::string &id, std::string &base, long long &off) {  // ids.go:11-27
    base = id;
    off = 0;
    size_t colon = id.rfind(':');
    if (colon == std::string::npos || colon == id.size() - 1) return false;
    std::string suf = id.substr(colon + 1);
    size_t dash = suf.find('-');
    if (dash == std::string::npos) return false;
    std::string num = suf.substr(0, dash);
    size_t b = (!num.empty() && (num[0] == '+' || num[0] == '-')) ? 1 : 0;
    if (num.size() == b) return false;
    for (size_t i = b; i < num.size(); i++)
        if (!isdigit((unsigned char)num[i])) return false;
    base = id.substr(0, colon);
    off = atoll(num.c_str());
    return true;
}

std::string ints_csv(const std::vector<int> &a) {
    std::string s;
    for (size_t i = 0; i < a.size(); i++) {
        if (i) s += ",";
        s += std::to_string(a[i]);
    }
    return s;
}

bool less_product(const Product &a, const Product &b) {  // common.LessProduct (sort.go:34-74)
    if (a.source_file != b.source_file) return a.source_file < b.source_file;
    std::string ab, bb;
    long long ao, bo;
    if (!split_chunk_suffix(a.sequence_id, ab, ao)) { ab = a.sequence_id; ao = 0; }
    if (!split_chunk_suffix(b.sequence_id, bb, bo)) { bb = b.sequence_id; bo = 0; }
    if (ab != bb) return ab < bb;
    if (a.start + ao != b.start + bo) return a.start + ao < b.start + bo;
    if (a.end + ao != b.end + bo) return a.end + ao < b.end + bo;
    if (a.length != b.length) return a.length < b.length;
    if (a.type != b.type) return a.type < b.type;
    if (a.experiment_id != b.experiment_id) return a.experiment_id < b.experiment_id;
    if (a.fwd_mm != b.fwd_mm) return a.fwd_mm < b.fwd_mm;
    if (a.rev_mm != b.rev_mm) return a.rev_mm < b.rev_mm;
    const std::string af = ints_csv(a.fwd_idx), bf = ints_csv(b.fwd_idx);
    if (af != bf) return af < bf;
    const std::string ar = ints_csv(a.rev_idx), br = ints_csv(b.rev_idx);
    if (ar != br) return ar < br;
    if (a.sequence_id != b.sequence_id) return a.sequence_id < b.sequence_id;
    return a.seq < b.seq;
}

std::string base_row(const Product &p) {  // output.FormatBaseRowTSV (rows.go:22-30)
    std::ostringstream o;
    o << p.source_file << '\t' << p.sequence_id << '\t' << p.experiment_id << '\t' << p.start << '\t' << p.end << '\t'
      << p.length << '\t' << p.type << '\t' << p.fwd_mm << '\t' << p.rev_mm << '\t' << ints_csv(p.fwd_idx) << '\t'
      << ints_csv(p.rev_idx);
    return o.str();
}

const char *kHeader = "source_file\tsequence_id\texperiment_id\tstart\tend\tlength\ttype\tfwd_mm\trev_mm\tfwd_mm_i\trev_mm_i";

struct LRU {  // runutil.LRUSet (lru_set.go:8-38)
    size_t cap;
    std::list<std::string> ll;
    std::unordered_map<std::string, std::list<std::string>::iterator> m;
    explicit LRU(long long c) : cap(c > 0 ? (size_t)c : 200000) {}
    bool add(const std::string &k) {
        auto it = m.find(k);
        if (it != m.end()) {
            ll.splice(ll.begin(), ll, it->second);
            return true;
        }
        ll.push_front(k);
        m[k] = ll.begin();
        if (ll.size() > cap) {
            m.erase(ll.back());
            ll.pop_back();
        }
        return false;
    }
};

// The same LRU over integer keys: plain runs (ipcr / multiplex) keep products as POD records and never
// build a string per product.  Equal strings were interned to equal integers, so key equality is the same.
struct PodKey {
    uint64_t a, b, c;
    bool operator==(const PodKey &o) const { return a == o.a && b == o.b && c == o.c; }
};
struct PodKeyHash {
    size_t operator()(const PodKey &k) const {
        uint64_t h = k.a * 0x9E3779B97F4A7C15ull;
        h ^= (k.b + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
        h ^= (k.c + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
        return (size_t)h;
    }
};
struct PodLRU {  // runutil.LRUSet (lru_set.go:8-38)
    size_t cap;
    std::list<PodKey> ll;
    std::unordered_map<PodKey, std::list<PodKey>::iterator, PodKeyHash> m;
    explicit PodLRU(long long c) : cap(c > 0 ? (size_t)c : 200000) {}
    bool add(const PodKey &k) {
        auto it = m.find(k);
        if (it != m.end()) {
            ll.splice(ll.begin(), ll, it->second);
            return true;
        }
        ll.push_front(k);
        m[k] = ll.begin();
        if (ll.size() > cap) {
            m.erase(ll.back());
            ll.pop_back();
        }
        return false;
    }
};

struct JobMeta {  // one scanned record / chunk
    std::string id, base;  // SequenceID as scanned; ID without the chunk suffix
    long long off = 0;     // chunk offset (0 when the ID carries no suffix)
    bool chunked = false;
    uint32_t file = 0, base_intern = 0;
};

static void append_int(std::string &s, long long v) {
    char buf[24];
    char *e = buf + sizeof buf, *p = e;
    unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
    do { *--p = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) *--p = '-';
    s.append(p, (size_t)(e - p));
}

// FwdMismatchIdx ascending, RevMismatchIdx in the reference's (flipped, hence descending) order
// (core/engine/engine.go:127-136) -- the same lists ipcr_gpu_product_mismatch_idx returns
static void append_idx_csv(std::string &s, const ipcr_product &pc, int which) {
    bool first = true;
    if (which == 0) {
        for (uint64_t m = pc.fwd_mm_mask; m; m &= m - 1) {
            if (!first) s += ',';
            first = false;
            append_int(s, __builtin_ctzll(m));
        }
    } else {
        for (uint64_t m = pc.rev_mm_mask; m;) {
            const int j = 63 - __builtin_clzll(m);
            m &= ~(1ull << j);
            if (!first) s += ',';
            first = false;
            append_int(s, j);
        }
    }
}

struct Gpu {
    ipcr_panel *panel = nullptr;
    ipcr_scratch *scratch = nullptr;
    char err[1024];
    void check(int rc) { if (rc != 0) die(3, err); }
};

ipcr_panel *make_panel(const Options &o, const std::vector<Pair> &pairs, bool inner, char *err) {
    ipcr_cfg cfg{};
    cfg.max_mm = o.mismatches;
    cfg.terminal_window = o.terminal_window < 1 ? 0 : o.terminal_window;  // EffectiveTerminalWindow
    cfg.seed_len = o.seed_length;
    if (!inner) {
        cfg.min_len = o.min_len;
        cfg.max_len = o.max_len;
        cfg.hit_cap = o.hit_cap;
        cfg.circular = o.circular;
    }  // inner engine: MinLen = MaxLen = HitCap = 0, linear (nestedapp/app.go:155-165)
    std::vector<ipcr_pair> cp(pairs.size());
    for (size_t i = 0; i < pairs.size(); i++) {
        cp[i].fwd = pairs[i].fwd.data();
        cp[i].fwd_len = (uint32_t)pairs[i].fwd.size();
        cp[i].rev = pairs[i].rev.data();
        cp[i].rev_len = (uint32_t)pairs[i].rev.size();
        cp[i].min_product = pairs[i].min_product;
        cp[i].max_product = pairs[i].max_product;
    }
    ipcr_panel *p = ipcr_gpu_panel_create(&cfg, cp.data(), cp.size(), err, 1024);
    if (!p) die(2, err);
    return p;
}

Options parse(int argc, char **argv) {
    Options o;
    int i = 1;
    if (i < argc && argv[i][0] != '-') o.mode = argv[i++];
    auto need = [&](const std::string &flag) -> std::string {
        if (i + 1 >= argc) die(2, "flag needs an argument: " + flag);
        return argv[++i];
    };
    auto boolval = [&](const std::string &arg, size_t eq) { return eq == std::string::npos || arg.substr(eq + 1) != "false"; };
    for (; i < argc; i++) {
        std::string a = argv[i];
        if (a.size() < 2 || a[0] != '-') { o.files.push_back(a); continue; }
        size_t eq = a.find('=');
        std::string name = a.substr(a.find_first_not_of('-'), eq == std::string::npos ? std::string::npos : eq - a.find_first_not_of('-'));
        auto val = [&]() { return eq == std::string::npos ? need(a) : a.substr(eq + 1); };
        if (name == "forward" || name == "f") o.fwd = val();
        else if (name == "reverse" || name == "r") o.rev = val();
        else if (name == "primers" || name == "p") o.primers = val();
        else if (name == "sequences" || name == "s") o.files.push_back(val());
        else if (name == "mismatches" || name == "m") o.mismatches = atoi(val().c_str());
        else if (name == "min-length") o.min_len = atoll(val().c_str());
        else if (name == "max-length") o.max_len = atoll(val().c_str());
        else if (name == "hit-cap") o.hit_cap = atoll(val().c_str());
        else if (name == "terminal-window") o.terminal_window = atoi(val().c_str());
        else if (name == "seed-length") o.seed_length = atoi(val().c_str());
        else if (name == "chunk-size") o.chunk_size = atoll(val().c_str());
        else if (name == "dedupe-cap") o.dedupe_cap = atoll(val().c_str());
        else if (name == "threads" || name == "t") (void)val();  // one GPU stream serves the whole run
        else if (name == "self") o.self = boolval(a, eq);
        else if (name == "circular" || name == "c") o.circular = boolval(a, eq);
        else if (name == "sort") o.sort = boolval(a, eq);
        else if (name == "no-header") o.header = !boolval(a, eq);
        else if (name == "quiet" || name == "q") (void)0;
        else if (name == "no-match-exit-code") o.no_match_exit = atoi(val().c_str());
        else if (name == "output" || name == "o") { if (val() != "text") die(2, "this host writes --output text only"); }
        else if (name == "probe" || name == "P") o.probe = val();
        else if (name == "probe-name") o.probe_name = val();
        else if (name == "probe-max-mm" || name == "M") o.probe_max_mm = atoi(val().c_str());
        else if (name == "require-probe") o.require_probe = boolval(a, eq);
        else if (name == "inner-forward") o.inner_fwd = val();
        else if (name == "inner-reverse") o.inner_rev = val();
        else if (name == "inner-primers") o.inner_primers = val();
        else if (name == "require-inner") o.require_inner = boolval(a, eq);
        else die(2, "flag provided but not defined: " + a);
    }
    // clibase.Validate (common.go:124-174)
    const bool file = !o.primers.empty(), inl = !o.fwd.empty() || !o.rev.empty();
    if (file && inl) die(2, "--primers conflicts with --forward/--reverse");
    if (inl && (o.fwd.empty() || o.rev.empty())) die(2, "--forward and --reverse must be supplied together");
    if (!file && !inl) die(2, "provide --primers or --forward/--reverse");
    if (o.files.empty()) die(2, "at least one sequence file is required");
    if (o.chunk_size < 0) die(2, "--chunk-size must be >= 0");
    if (o.hit_cap < 0) die(2, "--hit-cap must be >= 0");
    if (o.terminal_window < -1) die(2, "--terminal-window must be >= -1");
    if (inl) { o.fwd = validate(o.fwd, "--forward"); o.rev = validate(o.rev, "--reverse"); }
    return o;
}

}  // namespace

int main(int argc, char **argv) {
    Options o = parse(argc, argv);
    // ---- pairs (app.go:90-102, probeapp, nestedapp, multiplexapp) ----
    std::vector<Pair> pairs, inner;
    if (!o.primers.empty()) pairs = load_tsv(o.primers);
    else pairs.push_back(Pair{o.mode == "nested" ? "outer" : "manual", o.fwd, o.rev, o.min_len, o.max_len});
    if (o.self) pairs = o.mode == "multiplex" ? add_self_pairs_unique(pairs) : add_self_pairs(pairs);
    if (o.mode == "nested") {
        if (!o.inner_primers.empty()) inner = load_tsv(o.inner_primers);
        else inner.push_back(Pair{"inner", validate(o.inner_fwd, "--inner-forward"), validate(o.inner_rev, "--inner-reverse"), 0, 0});
        if (o.self) inner = add_self_pairs(inner);
    }
    if (o.mode == "probe") {
        if (normalize(o.probe).empty()) die(2, "--probe is required");
        for (char c : normalize(o.probe))
            if (!strchr("ACGTRYSWKMBDHVN", c)) die(2, "--probe: invalid base");
    }
    // ---- chunk policy (core.go:78-98, runutil.go:39-59) ----
    long long max_plen = 0, eff = o.max_len;
    bool unbounded = false;
    for (const Pair &p : pairs) {
        max_plen = std::max<long long>(max_plen, std::max(p.fwd.size(), p.rev.size()));
        if (p.max_product > 0) eff = std::max(eff, p.max_product);
        else if (o.max_len <= 0) unbounded = true;
    }
    if (unbounded) eff = 0;
    if (eff > 0 && eff < max_plen)
        die(2, "effective maximum product length (" + std::to_string(eff) + ") is smaller than the longest primer length (" +
                   std::to_string(max_plen) + ")");
    if (o.min_len > 0 && o.max_len > 0 && o.min_len > o.max_len) die(2, "--min-length exceeds --max-length");
    long long chunk = o.chunk_size, overlap = eff;
    if (chunk <= 0 || o.circular || eff <= 0 || chunk <= eff) chunk = overlap = 0;

    // ---- engine (core.go:108-117) ----
    Gpu g;
    // probe / nested visitors read amplicons from the records resident on ONE device; plain runs may spread a
    // chromosome-sized record over every visible GPU behind the ABI
    const int n_dev = (o.mode == "probe" || o.mode == "nested") ? 1 : 0;
    if (ipcr_gpu_init(n_dev, g.err, sizeof g.err) != 0) die(3, g.err);
    g.panel = make_panel(o, pairs, false, g.err);
    g.scratch = ipcr_gpu_scratch_create(g.panel, g.err, sizeof g.err);
    if (!g.scratch) die(3, g.err);
    Gpu gi;
    std::vector<const char *> inner_ids;
    if (o.mode == "nested") {
        gi.panel = make_panel(o, inner, true, gi.err);
        gi.scratch = ipcr_gpu_scratch_create(gi.panel, gi.err, sizeof gi.err);
        if (!gi.scratch) die(3, gi.err);
        for (const Pair &p : inner) inner_ids.push_back(p.id.c_str());
    }
    const std::string probe = normalize(o.probe);
    const bool need_seq = o.mode == "probe" || o.mode == "nested";

    std::vector<Product> rows;
    LRU seen(o.dedupe_cap);
    // phase clocks for IPCR_CLI_TIMING=1 (stderr): where a dense-hit run spends its time
    using Clock = std::chrono::steady_clock;
    double t_open = 0, t_scan = 0, t_collect = 0, t_sort = 0, t_write = 0;
    auto secs = [](Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    uint64_t n_raw = 0;
    // plain runs: POD rows (rec_idx rewritten to a global job number), integer dedupe keys, device sort
    const bool plain = o.mode != "probe" && o.mode != "nested";
    std::vector<ipcr_product> pod;
    std::vector<JobMeta> metas;
    PodLRU seen_pod(o.dedupe_cap);
    std::unordered_map<std::string, uint32_t> base_ids, exp_ids;
    std::vector<uint32_t> pair_exp(pairs.size());
    for (size_t i = 0; i < pairs.size(); i++) pair_exp[i] = exp_ids.emplace(pairs[i].id, (uint32_t)exp_ids.size()).first->second;
    uint32_t file_no = 0;
    for (const std::string &fa : o.files) {
        const uint32_t this_file = file_no++;
        const auto tf0 = Clock::now();
        // plain unchunked runs take the feeder's packed form through ipcr_gpu_scan_fasta (0.375 B/bp over PCIe, no
        // pack kernel); nothing is page-locked: locking 1 GB costs more than a one-shot copy from pageable memory saves
        const bool packed_path = plain && chunk <= 0;
        // (and only ids and lengths are read back, so the feeder does not even write the ASCII records: PACK_ONLY)
        ipcr_fasta *ff = ipcr_gpu_fasta_open(fa.c_str(), 0, packed_path ? IPCR_FASTA_PACK_ONLY : 0, g.err, sizeof g.err);
        if (!ff) die(3, g.err);
        t_open += secs(tf0, Clock::now());
        if (packed_path) {
            const size_t nrec = ipcr_gpu_fasta_num_records(ff);
            const size_t meta_base = metas.size();
            for (size_t r = 0; r < nrec; r++) {
                const char *id; uint64_t len;
                ipcr_gpu_fasta_record(ff, r, &id, nullptr, &len);
                JobMeta m;
                m.id = id;
                m.chunked = split_chunk_suffix(m.id, m.base, m.off);  // the collector strips ANY :a-b suffix (pipeline.go:140-156)
                if (!m.chunked) { m.base = m.id; m.off = 0; }
                m.file = this_file;
                m.base_intern = base_ids.emplace(m.base, (uint32_t)base_ids.size()).first->second;
                metas.push_back(std::move(m));
            }
            const ipcr_product *out = nullptr;
            uint64_t n = 0;
            const auto ts0 = Clock::now();
            g.check(ipcr_gpu_scan_fasta(g.panel, g.scratch, ff, &out, &n, g.err, sizeof g.err));
            const auto ts1 = Clock::now();
            t_scan += secs(ts0, ts1);
            n_raw += n;
            for (uint64_t i = 0; i < n; i++) {
                ipcr_product pc = out[i];
                const JobMeta &m = metas[meta_base + pc.rec_idx];
                const PodKey key{((uint64_t)m.file << 32) | m.base_intern, (uint64_t)(pc.start + m.off),
                                 ((uint64_t)(pc.end + m.off) << 20) | ((uint64_t)pc.type << 19) | pair_exp[pc.pair_idx]};
                if (seen_pod.add(key)) continue;
                pc.rec_idx = (uint32_t)(meta_base + pc.rec_idx);
                pod.push_back(pc);
            }
            t_collect += secs(ts1, Clock::now());
            ipcr_gpu_fasta_close(ff);
            continue;
        }
        const size_t nrec = ipcr_gpu_fasta_num_records(ff);
        // jobs = records or rolling chunks (path_ctx.go:83-179), scanned in batches of <= 4096
        struct Job { std::string id; const uint8_t *seq; uint64_t len; };
        std::vector<Job> jobs;
        for (size_t r = 0; r < nrec; r++) {
            const char *id; const uint8_t *seq; uint64_t len;
            ipcr_gpu_fasta_record(ff, r, &id, &seq, &len);
            const long long step = chunk - overlap;
            if (chunk <= 0 || overlap >= chunk || step <= 0 || (long long)len <= chunk) {
                jobs.push_back(Job{id, seq, len});
                continue;
            }
            long long start = 0, last_end = 0;
            while ((long long)len - start > chunk) {
                jobs.push_back(Job{std::string(id) + ":" + std::to_string(start) + "-" + std::to_string(start + chunk), seq + start, (uint64_t)chunk});
                last_end = start + chunk;
                start += step;
            }
            if (last_end < (long long)len)
                jobs.push_back(Job{std::string(id) + ":" + std::to_string(start) + "-" + std::to_string(len), seq + start, len - (uint64_t)start});
        }
        for (size_t j0 = 0; j0 < jobs.size(); j0 += 4096) {
            const size_t j1 = std::min(jobs.size(), j0 + 4096);
            std::vector<ipcr_seq> recs;
            for (size_t j = j0; j < j1; j++) recs.push_back(ipcr_seq{jobs[j].seq, jobs[j].len});
            const ipcr_product *out = nullptr;
            uint64_t n = 0;
            const auto ts0 = Clock::now();
            g.check(ipcr_gpu_scan_batch(g.panel, g.scratch, recs.data(), recs.size(), &out, &n, g.err, sizeof g.err));
            const auto ts1 = Clock::now();
            t_scan += secs(ts0, ts1);
            n_raw += n;
            struct CollectClock {
                double &acc; Clock::time_point t0;
                ~CollectClock() { acc += std::chrono::duration<double>(Clock::now() - t0).count(); }
            } collect_clock{t_collect, ts1};
            if (n == 0) continue;
            if (plain) {
                const size_t meta_base = metas.size();
                for (size_t j = j0; j < j1; j++) {
                    JobMeta m;
                    m.id = jobs[j].id;
                    m.chunked = split_chunk_suffix(m.id, m.base, m.off);
                    if (!m.chunked) { m.base = m.id; m.off = 0; }
                    m.file = this_file;
                    m.base_intern = base_ids.emplace(m.base, (uint32_t)base_ids.size()).first->second;
                    metas.push_back(std::move(m));
                }
                for (uint64_t i = 0; i < n; i++) {
                    ipcr_product pc = out[i];
                    const JobMeta &m = metas[meta_base + pc.rec_idx];
                    // collector key (pipeline.go:26-30,146-149): base, file, global start / end, type, experiment
                    const PodKey key{((uint64_t)m.file << 32) | m.base_intern, (uint64_t)(pc.start + m.off),
                                     ((uint64_t)(pc.end + m.off) << 20) | ((uint64_t)pc.type << 19) | pair_exp[pc.pair_idx]};
                    if (seen_pod.add(key)) continue;
                    pc.rec_idx = (uint32_t)(meta_base + pc.rec_idx);
                    pod.push_back(pc);
                }
                continue;
            }
            std::vector<ipcr_product> raw(out, out + n);  // the visitors reuse the scratch's buffers
            // collector: global coordinates + LRU dedupe (pipeline.go:134-161)
            std::vector<ipcr_product> kept;
            std::vector<Product> prods;
            for (const ipcr_product &pc : raw) {
                const Job &job = jobs[j0 + pc.rec_idx];
                Product p;
                p.experiment_id = pairs[pc.pair_idx].id;
                p.sequence_id = job.id;
                p.source_file = fa;
                p.start = pc.start; p.end = pc.end; p.length = pc.length;
                p.type = pc.type == 0 ? "forward" : "revcomp";
                p.fwd_mm = pc.fwd_mm; p.rev_mm = pc.rev_mm;
                int32_t buf[64];
                int k = ipcr_gpu_product_mismatch_idx(&pc, 0, buf, 64);
                p.fwd_idx.assign(buf, buf + k);
                k = ipcr_gpu_product_mismatch_idx(&pc, 1, buf, 64);
                p.rev_idx.assign(buf, buf + k);
                std::string base; long long off;
                const bool ok = split_chunk_suffix(p.sequence_id, base, off);
                if (!ok) { base = p.sequence_id; off = 0; }
                const std::string key = base + '\x1f' + fa + '\x1f' + std::to_string(p.start + off) + '\x1f' +
                                        std::to_string(p.end + off) + '\x1f' + p.type + '\x1f' + p.experiment_id;
                if (seen.add(key)) continue;
                if (ok) { p.sequence_id = base; p.start += off; p.end += off; }
                kept.push_back(pc);
                prods.push_back(std::move(p));
            }
            if (kept.empty()) continue;
            const uint64_t nk = kept.size();
            if (need_seq) {  // Product.Seq (pipeline.go:81-88), gathered on the device
                std::vector<uint64_t> offs(nk + 1);
                uint64_t need = 0;
                g.check(ipcr_gpu_amplicons(g.scratch, kept.data(), nk, nullptr, 0, offs.data(), &need, g.err, sizeof g.err));
                std::vector<uint8_t> amp(need + 1);
                g.check(ipcr_gpu_amplicons(g.scratch, kept.data(), nk, amp.data(), need, offs.data(), &need, g.err, sizeof g.err));
                for (uint64_t i = 0; i < nk; i++) prods[i].seq.assign((const char *)amp.data() + offs[i], offs[i + 1] - offs[i]);
            }
            if (o.mode == "probe") {  // visitors.Probe (probe.go:18-32)
                std::vector<ipcr_probe_hit> hits(nk);
                g.check(ipcr_gpu_probe_annotate(g.scratch, kept.data(), nk, probe.data(), (uint32_t)probe.size(), o.probe_max_mm,
                                                hits.data(), g.err, sizeof g.err));
                for (uint64_t i = 0; i < nk; i++) {
                    if (o.require_probe && !hits[i].found) continue;
                    Product &p = prods[i];
                    if (hits[i].found) {
                        p.probe_found = true;
                        p.probe_strand = std::string(1, (char)hits[i].strand);
                        p.probe_pos = hits[i].pos;
                        p.probe_mm = hits[i].mm;
                        p.probe_site = upper(p.seq).substr((size_t)hits[i].pos, probe.size());
                    }
                    rows.push_back(p);
                }
            } else if (o.mode == "nested") {  // visitors.Nested (nested.go:17-66), one batched inner scan
                std::vector<ipcr_product> best(nk);
                g.check(ipcr_gpu_nested(g.scratch, kept.data(), nk, gi.panel, gi.scratch, inner_ids.data(), best.data(), g.err,
                                        sizeof g.err));
                for (uint64_t i = 0; i < nk; i++) {
                    Product &p = prods[i];
                    if (best[i].length < 0) {
                        if (o.require_inner) continue;
                    } else {
                        p.inner_found = true;
                        p.inner_id = inner[best[i].pair_idx].id;
                        p.inner_start = best[i].start; p.inner_end = best[i].end; p.inner_length = best[i].length;
                        p.inner_type = best[i].type == 0 ? "forward" : "revcomp";
                        p.inner_fwd_mm = best[i].fwd_mm; p.inner_rev_mm = best[i].rev_mm;
                    }
                    rows.push_back(p);
                }
            } else {
                for (auto &p : prods) rows.push_back(std::move(p));
            }
        }
        ipcr_gpu_fasta_close(ff);
    }
    const auto tq0 = Clock::now();
    if (o.sort) std::stable_sort(rows.begin(), rows.end(), less_product);
    if (plain && o.sort && pod.size() > 1) {
        // common.LessProduct (sort.go:34-74) with the strings as ranks; the integer part runs on the device
        std::vector<uint32_t> job_rank(metas.size()), pair_rank(pairs.size());
        {
            std::vector<std::pair<std::string, std::string>> keys;  // (file, base)
            for (const JobMeta &m : metas) keys.emplace_back(o.files[m.file], m.base);
            std::vector<std::pair<std::string, std::string>> uniq(keys);
            std::sort(uniq.begin(), uniq.end());
            uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
            for (size_t j = 0; j < metas.size(); j++)
                job_rank[j] = (uint32_t)(std::lower_bound(uniq.begin(), uniq.end(), keys[j]) - uniq.begin());
            std::vector<std::string> ue;
            for (const Pair &p : pairs) ue.push_back(p.id);
            std::sort(ue.begin(), ue.end());
            ue.erase(std::unique(ue.begin(), ue.end()), ue.end());
            for (size_t i = 0; i < pairs.size(); i++)
                pair_rank[i] = (uint32_t)(std::lower_bound(ue.begin(), ue.end(), pairs[i].id) - ue.begin());
        }
        std::vector<int64_t> job_off(metas.size());
        for (size_t j = 0; j < metas.size(); j++) job_off[j] = metas[j].off;
        auto tail_less = [&](const ipcr_product &a, const ipcr_product &b) {  // what LessProduct looks at after RevMM
            std::string x, y;
            append_idx_csv(x, a, 0); append_idx_csv(y, b, 0);
            if (x != y) return x < y;
            x.clear(); y.clear();
            append_idx_csv(x, a, 1); append_idx_csv(y, b, 1);
            if (x != y) return x < y;
            const JobMeta &ma = metas[a.rec_idx], &mb = metas[b.rec_idx];
            return (ma.chunked ? ma.base : ma.id) < (mb.chunked ? mb.base : mb.id);
        };
        if (metas.size() < (1u << 20) && pairs.size() < (1u << 19)) {
            std::vector<uint8_t> tie(pod.size());
            g.check(ipcr_gpu_sort_products(g.scratch, pod.data(), pod.size(), job_rank.data(), job_off.data(), metas.size(),
                                           pair_rank.data(), pairs.size(), tie.data(), g.err, sizeof g.err));
            for (size_t i = 0; i < pod.size();) {  // runs equal on every integer key (rare): finish on the host
                size_t j = i + 1;
                while (j < pod.size() && tie[j]) j++;
                if (j - i > 1) std::stable_sort(pod.begin() + (long)i, pod.begin() + (long)j, tail_less);
                i = j;
            }
        } else {
            std::stable_sort(pod.begin(), pod.end(), [&](const ipcr_product &a, const ipcr_product &b) {
                const JobMeta &ma = metas[a.rec_idx], &mb = metas[b.rec_idx];
                if (job_rank[a.rec_idx] != job_rank[b.rec_idx]) return job_rank[a.rec_idx] < job_rank[b.rec_idx];
                if (a.start + ma.off != b.start + mb.off) return a.start + ma.off < b.start + mb.off;
                if (a.end + ma.off != b.end + mb.off) return a.end + ma.off < b.end + mb.off;
                if (a.length != b.length) return a.length < b.length;
                if (a.type != b.type) return a.type < b.type;
                if (pair_rank[a.pair_idx] != pair_rank[b.pair_idx]) return pair_rank[a.pair_idx] < pair_rank[b.pair_idx];
                if (a.fwd_mm != b.fwd_mm) return a.fwd_mm < b.fwd_mm;
                if (a.rev_mm != b.rev_mm) return a.rev_mm < b.rev_mm;
                return tail_less(a, b);
            });
        }
    }
    t_sort = secs(tq0, Clock::now());
    const auto tw0 = Clock::now();

    // ---- writers ----
    std::string out;
    if (o.header) {
        if (o.mode == "probe")
            out += std::string(kHeader) + "\tprobe_name\tprobe_seq\tprobe_found\tprobe_strand\tprobe_pos\tprobe_mm\tprobe_site\n";
        else if (o.mode == "nested")
            out += "source_file\tsequence_id\touter_experiment_id\touter_start\touter_end\touter_length\touter_type\t"
                   "inner_experiment_id\tinner_found\tinner_start\tinner_end\tinner_length\tinner_type\tinner_fwd_mm\tinner_rev_mm\n";
        else out += std::string(kHeader) + "\n";
    }
    if (plain) {  // output.FormatBaseRowTSV (rows.go:22-30) straight from the POD records
        out.reserve(out.size() + pod.size() * 96);
        for (const ipcr_product &pc : pod) {
            const JobMeta &m = metas[pc.rec_idx];
            out += o.files[m.file];
            out += '\t';
            out += m.chunked ? m.base : m.id;
            out += '\t';
            out += pairs[pc.pair_idx].id;
            out += '\t';
            append_int(out, pc.start + m.off);
            out += '\t';
            append_int(out, pc.end + m.off);
            out += '\t';
            append_int(out, pc.length);
            out += pc.type == 0 ? "\tforward\t" : "\trevcomp\t";
            append_int(out, pc.fwd_mm);
            out += '\t';
            append_int(out, pc.rev_mm);
            out += '\t';
            append_idx_csv(out, pc, 0);
            out += '\t';
            append_idx_csv(out, pc, 1);
            out += '\n';
        }
    }
    for (const Product &p : rows) {
        if (o.mode == "probe") {
            out += base_row(p) + '\t' + o.probe_name + '\t' + upper(o.probe) + '\t' + (p.probe_found ? "true" : "false") + '\t' +
                   p.probe_strand + '\t' + (p.probe_found ? std::to_string(p.probe_pos) : "") + '\t' +
                   (p.probe_found ? std::to_string(p.probe_mm) : "") + '\t' + p.probe_site + '\n';
        } else if (o.mode == "nested") {
            const bool f = p.inner_found;
            out += p.source_file + '\t' + p.sequence_id + '\t' + p.experiment_id + '\t' + std::to_string(p.start) + '\t' +
                   std::to_string(p.end) + '\t' + std::to_string(p.length) + '\t' + p.type + '\t' + p.inner_id + '\t' +
                   (f ? "true" : "false") + '\t' + (f ? std::to_string(p.inner_start) : "") + '\t' +
                   (f ? std::to_string(p.inner_end) : "") + '\t' + (f ? std::to_string(p.inner_length) : "") + '\t' +
                   (f ? p.inner_type : "") + '\t' + (f ? std::to_string(p.inner_fwd_mm) : "") + '\t' +
                   (f ? std::to_string(p.inner_rev_mm) : "") + '\n';
        } else {
            out += base_row(p) + '\n';
        }
    }
    fwrite(out.data(), 1, out.size(), stdout);
    t_write = secs(tw0, Clock::now());
    if (getenv("IPCR_CLI_TIMING"))
        fprintf(stderr, "[ipcr_gpu_cli] feeder %.3f s | scan (H2D + device + D2H) %.3f s | collect (coordinates, dedupe, rows; visitors) "
                        "%.3f s | sort %.3f s | format + write %.3f s | %llu products from the device, %zu rows, %zu bytes\n",
                t_open, t_scan, t_collect, t_sort, t_write, (unsigned long long)n_raw, plain ? pod.size() : rows.size(), out.size());
    if (gi.scratch) ipcr_gpu_scratch_destroy(gi.scratch);
    if (gi.panel) ipcr_gpu_panel_destroy(gi.panel);
    ipcr_gpu_scratch_destroy(g.scratch);
    ipcr_gpu_panel_destroy(g.panel);
    return (plain ? pod.empty() : rows.empty()) ? o.no_match_exit : 0;
}
