// tests/hostsim/hostsim.cpp — runs the product's host/device scalar cores (csrc/*_core.cuh) on the CPU,
// stage by stage in the same order as the CUDA kernels, so the staged zlib-9 formulation (chains ->
// probe table -> table walk -> trees -> bit packing) can be checked against libz without a GPU.
// TEST INFRASTRUCTURE: built and loaded by tests/ only; the product library never contains this file.
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <vector>

#define CCT_STATS 1
#include "deflate_core.cuh"
#include "inflate_core.cuh"

using namespace cct;

long long cct_stat[8];

static void chains_reference(const uint8_t *w, int n, std::vector<uint16_t> &dist16) {
    // dist16[s] = s - (latest p < s, p >= 1, with hash3(p) == hash3(s)), 0 if none or >= 32768
    std::vector<int32_t> head(32768, 0);
    dist16.assign((size_t)n + 1, 0);
    for (int s = 0; s + 2 < n; s++) {
        uint32_t h = hash3(w + s);
        int p = head[h];
        if (p != 0 && s - p < 32768) dist16[s] = (uint16_t)(s - p);
        head[h] = s;  // position 0 is stored as 0 == NIL, exactly like zlib
    }
}

static void runs_reference(const uint8_t *w, int n, std::vector<uint16_t> &rb16, std::vector<uint16_t> &rl16) {
    rb16.assign((size_t)n + 1, 0);
    rl16.assign((size_t)n + 1, 0);
    for (int p = 1; p < n; p++) rb16[p] = w[p] == w[p - 1] ? (uint16_t)(rb16[p - 1] == 65535 ? 65535 : rb16[p - 1] + 1) : 0;
    for (int p = n - 2; p >= 0; p--) rl16[p] = w[p] == w[p + 1] ? (uint16_t)(rl16[p + 1] == 65535 ? 65535 : rl16[p + 1] + 1) : 0;
}

static uint32_t adler32_(const uint8_t *p, int64_t n) {
    uint32_t a = 1, b = 0;
    while (n > 0) {
        int64_t k = n < 5552 ? n : 5552;
        n -= k;
        while (k--) { a += *p++; b += a; }
        a %= 65521; b %= 65521;
    }
    return (b << 16) | a;
}

// zlib's longest_match candidate loop, candidate by candidate over the chain links (no run shortcuts): the ground
// truth the sorted scan is checked against.  Cold start (best_len 2) with the given chain budget.
static uint32_t lm_reference(const uint8_t *w, int n, const uint16_t *dist16, int s, int budget) {
    const int look = n - s;
    if (look < kMinMatch || dist16[s] == 0 || dist16[s] > kMaxDist) return 0;
    const int cap = look < kMaxMatch ? look : kMaxMatch;
    int best = 2, bdist = 0, cur = s - dist16[s];
    for (;;) {
        int len = 0;
        while (len < cap && w[cur + len] == w[s + len]) len++;
        if (len > best) { best = len; bdist = s - cur; if (len >= cap) break; }
        if (dist16[cur] == 0) break;
        cur -= dist16[cur];
        if (s - cur >= kMaxDist) break;
        if (--budget == 0) break;
    }
    return best >= kMinMatch ? mtab_pack(best, bdist) : 0u;
}

// Chain members of a stream sorted by (hash, position) and the run counts, as the sort kernels produce them.
static void sorted_reference(const uint8_t *w, int n, std::vector<SortEnt> &ent, std::vector<uint32_t> &runs) {
    std::vector<uint16_t> rb16, rl16;
    runs_reference(w, n, rb16, rl16);
    runs.assign((size_t)n + 1, 0);
    for (int p = 0; p < n; p++) runs[p] = (uint32_t)rl16[p] | ((uint32_t)rb16[p] << 16);
    std::vector<std::pair<uint32_t, int>> key;
    for (int p = 1; p + 2 < n; p++) key.push_back({sort_key(hash3(w + p)), p});
    std::stable_sort(key.begin(), key.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
    ent.assign(key.size() + 1, SortEnt{0, 0});
    for (size_t i = 0; i < key.size(); i++) {
        const int p = key[i].second;
        uint32_t q = 0;
        for (int k = 0; k < 4; k++) if (p + k < n) q |= (uint32_t)w[p + k] << (8 * k);
        uint32_t a = (uint32_t)p;
        if (quad_is_run(q)) a |= kEntRun;
        if (i == 0 || key[i - 1].first != key[i].first) a |= kEntFirst;
        ent[i] = SortEnt{a, q};
    }
}

// The probe tables of a stream the way the kernels fill them: k_bprobe (quick walk, general walk, the members of byte
// runs left out and announced by the run's start), then k_run_tails.  Returns the number of entries whose budgets differ;
// tail_stats[0] = runs announced, [1] = of those left to the general walk, [2] = records collected.
static int64_t probe_tables(const BucketView &bv, const std::vector<SortEnt> &ent, int n, uint32_t *mlo, uint32_t *mhi,
                            int64_t *tail_stats = nullptr) {
    int64_t n_diff = 0;
    const int members = n >= 4 ? n - 3 : 0;
    std::vector<int> leaders;
    auto put = [&](int p, uint32_t lo, uint32_t hi) {
        if (lo != hi) { mhi[p] = hi; lo |= kProbeHiDiff; n_diff++; }
        mlo[p] = lo;
    };
    for (int i = 0; i < members; i++) {
        uint32_t lo, hi;
        if (bucket_probe_quick(bv, i, ent[i], true, true, &lo)) {
            if (lo >= kProbeTail) { if (lo == kProbeTailLeader) leaders.push_back(i); continue; }
            hi = lo;
        } else bucket_probe(bv, i, ent[i], true, &lo, &hi);
        put(ent_pos(ent[i]), lo, hi);
    }
    for (int i0 : leaders) {
        TailGroup recs[kTailRecords];
        uint32_t lo, hi;
        const int nrec = run_collect(bv, i0, true, recs, 1, &lo, &hi), s0 = ent_pos(ent[i0]);
        const int len = (int)(bv.runs[s0] & 0xFFFFu) + 1;
        if (tail_stats) { tail_stats[0]++; if (nrec < 0) tail_stats[1]++; else tail_stats[2] += nrec; }
        if (nrec < 0) bucket_probe(bv, i0, ent[i0], true, &lo, &hi);
        put(s0, lo, hi);
        for (int j = std::max(1, len - (kMaxMatch - 1)); j <= len - kMinMatch; j++) {
            if (nrec >= 0) tail_entry(recs, 1, nrec, len - j, &lo, &hi);
            else bucket_probe(bv, i0 + j, ent[i0 + j], true, &lo, &hi);
            put(s0 + j, lo, hi);
        }
    }
    return n_diff;
}

extern "C" {

// Staged zlib-9 of src[0..n).
// stats[0] = symbols, [1] = blocks, [2] = positions whose two chain budgets give different matches, [3] = scan steps
// of the probe, [4] = 0 (stats must hold 5 values if non-null).
int64_t hs_deflate2(const uint8_t *src, int n, uint8_t *dst, int64_t cap, int parse_win, int64_t *stats) {
    long long c0 = cct_stat[2];
    std::vector<uint8_t> wbuf(src, src + n);
    wbuf.resize((size_t)n + 16, 0);
    src = wbuf.data();
    std::vector<SortEnt> ent;
    std::vector<uint32_t> runs;
    sorted_reference(src, n, ent, runs);
    BucketView bv{src, ent.data(), runs.data(), n};
    LzView v{src, n};
    std::vector<uint32_t> mlo((size_t)n + 1, 0), mhi((size_t)n + 1, 0xDEADBEEFu);  // (mhi is read only where flagged)
    const int64_t n_incomplete = probe_tables(bv, ent, n, mlo.data(), mhi.data());
    ProbeTab mtab{mlo.data(), mhi.data()};
    long long c1 = cct_stat[2];
    // speculative windows -> stitch -> compaction (the three parse kernels, sequentially)
    const int win = parse_win > 0 ? parse_win : 1024;
    std::vector<uint32_t> mask((size_t)n / 32 + 2, 0xFFFFFFFFu);  // (garbage: the window walks must write every word)
    std::vector<uint32_t> rec((size_t)n + 1, 0);
    for (int p = 0; p < n; p++) rec[p] = walk_precompute(v, mtab, p);
    std::vector<Walk> exits((size_t)n / win + 2);
    for (int k = 0, ws = 0; ws < n; k++, ws += win) exits[k] = walk_window(v, mtab, ws, std::min(ws + win, n), mask.data(), rec.data());
    walk_stitch(v, mtab, win, mask.data(), rec.data(), exits.data());
    std::vector<uint32_t> syms((size_t)n + 1);
    std::vector<int32_t> blk_end((size_t)n / kSymEnd + 2);
    int nsym = 0;
    for (int p = 0; p < n; p++) if ((mask[p >> 5] >> (p & 31)) & 1u) {
        if (nsym > 0 && nsym % kSymEnd == 0) blk_end[nsym / kSymEnd - 1] = p;
        syms[nsym++] = rec[p] & ~kSymSlow;
    }
    int n_flush = finish_blocks(nsym, nsym ? syms[nsym - 1] : 0u, n, blk_end.data());
    long long c2 = cct_stat[2];
    int nblk = n_flush + 1;
    std::vector<BlockDesc> desc((size_t)nblk);
    BlockWork *wk = new BlockWork;
    for (int b = 0; b < nblk; b++) {
        BlockDesc &d = desc[b];
        d.sym_begin = (uint32_t)b * kSymEnd;
        d.nsym = b < n_flush ? (uint32_t)kSymEnd : (uint32_t)(nsym - n_flush * kSymEnd);
        d.pos_begin = b == 0 ? 0u : (uint32_t)blk_end[b - 1];
        d.stored_len = (b < n_flush ? (uint32_t)blk_end[b] : (uint32_t)n) - d.pos_begin;
        d.last = b == nblk - 1;
        memset(wk, 0, sizeof(BlockWork));
        wk->lt.freq[256] = 1;
        for (uint32_t i = 0; i < d.nsym; i++) {
            uint32_t sy = syms[d.sym_begin + i];
            if (sym_dist(sy) == 0) wk->lt.freq[sym_lc(sy)]++;
            else { wk->lt.freq[257 + len_code(sym_lc(sy))]++; wk->dt.freq[dist_code(sym_dist(sy) - 1)]++; }
        }
        const int64_t top = b < n_flush ? (int64_t)blk_end[b] - sym_advance(syms[d.sym_begin + d.nsym - 1]) + 1 : (int64_t)n;
        build_block(*wk, &d, d.stored_len, d.last, (int64_t)d.pos_begin >= window_base_at(top, n));
        fill_block_tables(*wk, &d, 0, 1);
    }
    delete wk;
    // layout + packing
    std::vector<uint8_t> out((size_t)deflate_bound(n) + 64, 0);
    out[0] = 0x78; out[1] = 0xDA;
    uint64_t bit = 16;
    auto put = [&](uint64_t bits, int nb) {
        for (int i = 0; i < nb; i++, bit++) if ((bits >> i) & 1) out[bit >> 3] |= (uint8_t)(1u << (bit & 7));
    };
    for (int b = 0; b < nblk; b++) {
        BlockDesc &d = desc[b];
        d.bit_start = bit;
        for (uint32_t i = 0; i < d.hdr_bits; i++) put((d.hdr[i >> 5] >> (i & 31)) & 1u, 1);
        if (d.type == 0) {
            bit = (bit + 7) & ~7ull;
            put(d.stored_len & 0xFFFF, 16); put(~d.stored_len & 0xFFFF, 16);
            memcpy(&out[bit >> 3], src + d.pos_begin, d.stored_len);
            bit += 8ull * d.stored_len;
        } else {
            uint64_t before = bit;
            for (uint32_t i = 0; i < d.nsym; i++) {
                int nb; uint64_t bits = sym_bits(syms[d.sym_begin + i], d.lcode, d.llen, d.dcode, d.dlen, &nb);
                if (nb != sym_nbits(syms[d.sym_begin + i], d.llen, d.dlen)) return -3;
                put(bits, nb);
            }
            put(d.lcode[256], d.llen[256]);
            if (bit - before != d.body_bits) return -2;  // opt_len / static_len accounting must be exact
        }
    }
    bit = (bit + 7) & ~7ull;
    uint32_t ad = adler32_(src, n);
    int64_t len = (int64_t)(bit >> 3);
    out[len++] = (uint8_t)(ad >> 24); out[len++] = (uint8_t)(ad >> 16); out[len++] = (uint8_t)(ad >> 8); out[len++] = (uint8_t)ad;
    if (stats) { stats[0] = nsym; stats[1] = nblk; stats[2] = n_incomplete; stats[3] = c1 - c0; stats[4] = c2 - c1; }
    if (len > cap) return -1;
    memcpy(dst, out.data(), (size_t)len);
    return len;
}

// Probe tables from the hash-sorted scan (bucket_probe) against zlib's candidate loop over the chain links.
// Returns the number of differing entries (0 = identical).  steps[0] = scan steps of the general walk, steps[1] =
// targets the quick walk settles (steps must hold 2 values if non-null).
int64_t hs_probe_check(const uint8_t *src, int n, int64_t *steps) {
    std::vector<uint8_t> w(src, src + n);
    w.resize((size_t)n + 16, 0);
    std::vector<uint16_t> dist16;
    chains_reference(w.data(), n, dist16);
    std::vector<SortEnt> ent;
    std::vector<uint32_t> runs;
    sorted_reference(w.data(), n, ent, runs);
    BucketView bv{w.data(), ent.data(), runs.data(), n};
    std::vector<uint32_t> lo((size_t)n + 1, 0), hi((size_t)n + 1, 0);
    const long long c0 = cct_stat[2];
    const int members = n >= 4 ? n - 3 : 0;
    int64_t bad = 0, quick = 0;
    for (int i = 0; i < members; i++) {
        const int p = ent_pos(ent[i]);
        bucket_probe(bv, i, ent[i], true, &lo[p], &hi[p]);
        uint32_t q;
        if (bucket_probe_quick(bv, i, ent[i], true, false, &q)) { quick++; if (q != lo[p] || q != hi[p]) bad++; }  // the short cut agrees wherever it settles
    }
    if (steps) { steps[0] = cct_stat[2] - c0; steps[1] = quick; }
    for (int s = 0; s < n; s++)
        if (lm_reference(w.data(), n, dist16.data(), s, kChainFull) != lo[s] || lm_reference(w.data(), n, dist16.data(), s, kChainGood) != hi[s]) bad++;
    // the tables as the kernels fill them (tails of long runs settled once per run) must be the same
    std::vector<uint32_t> klo((size_t)n + 1, 0), khi((size_t)n + 1, 0);
    int64_t ts[3] = {0, 0, 0};
    probe_tables(bv, ent, n, klo.data(), khi.data(), ts);
    for (int s = 0; s < n; s++) {
        const uint32_t want_lo = lo[s] | (lo[s] != hi[s] ? kProbeHiDiff : 0u);
        if (klo[s] != want_lo || (lo[s] != hi[s] && khi[s] != hi[s])) {
            bad++;
            if (getenv("HS_DEBUG") && bad < 12)
                fprintf(stderr, "pos %d runs rl=%u rb=%u: want lo (%d,%d) hi (%d,%d), kernels lo (%d,%d)%s hi (%d,%d)\n", s, runs[s] & 0xFFFF, runs[s] >> 16,
                        mtab_len(lo[s]), mtab_dist(lo[s]), mtab_len(hi[s]), mtab_dist(hi[s]), mtab_len(klo[s]), mtab_dist(klo[s]),
                        (klo[s] & kProbeHiDiff) ? "*" : "", mtab_len(khi[s]), mtab_dist(khi[s]));
        }
    }
    if (steps) { steps[2] = ts[0]; steps[3] = ts[1]; steps[4] = ts[2]; }
    return bad;
}

// What cct_debug_match_tables returns for one string: ent[n] (8-byte entries, n-3 of them used), lo[n], hi[n]
// (lo with kProbeHiDiff where hi differs; hi meaningful only there).
void hs_match_tables(const uint8_t *src, int n, uint64_t *ent_out, uint32_t *lo_out, uint32_t *hi_out) {
    std::vector<uint8_t> w(src, src + n);
    w.resize((size_t)n + 16, 0);
    std::vector<SortEnt> ent;
    std::vector<uint32_t> runs;
    sorted_reference(w.data(), n, ent, runs);
    BucketView bv{w.data(), ent.data(), runs.data(), n};
    const int members = n >= 4 ? n - 3 : 0;
    for (int p = 0; p < n; p++) { lo_out[p] = 0; hi_out[p] = 0; }
    probe_tables(bv, ent, n, lo_out, hi_out);
    for (int i = 0; i < members; i++) ent_out[i] = (uint64_t)ent[i].a | ((uint64_t)ent[i].q << 32);
}

// Where the general walk's work comes from (diagnostics): per class of target -- 0 run member with members of its run
// below it, 1 run start, 2 ordinary position -- out[c] = targets left to the general walk, out[3 + c] = its steps for
// them, out[6] = warp tasks (32 consecutive entries) with at least one such target, out[7] = the sum over those tasks of
// the longest walk in them, out[8] = targets deferred to the run tails, out[9] = all targets.
void hs_general_walk_stats(const uint8_t *src, int n, int64_t *out) {
    std::vector<uint8_t> w(src, src + n);
    w.resize((size_t)n + 16, 0);
    std::vector<SortEnt> ent;
    std::vector<uint32_t> runs;
    sorted_reference(w.data(), n, ent, runs);
    BucketView bv{w.data(), ent.data(), runs.data(), n};
    const int members = n >= 4 ? n - 3 : 0;
    for (int k = 0; k < 10; k++) out[k] = 0;
    long long task_max = 0; bool task_any = false;
    for (int i = 0; i < members; i++) {
        if ((i & 31) == 0) { if (task_any) { out[6]++; out[7] += task_max; } task_any = false; task_max = 0; }
        uint32_t lo, hi;
        out[9]++;
        if (bucket_probe_quick(bv, i, ent[i], true, true, &lo)) { if (lo >= kProbeTail) out[8]++; continue; }
        const long long c0 = cct_stat[2];
        bucket_probe(bv, i, ent[i], true, &lo, &hi);
        const long long st = cct_stat[2] - c0;
        const int s = ent_pos(ent[i]);
        int cls = 2;
        if (ent[i].a & kEntRun) cls = ((runs[s] >> 16) >= 1 && s >= 2) ? 0 : 1;
        out[cls]++; out[3 + cls] += st;
        task_any = true; if (st > task_max) task_max = st;
    }
    if (task_any) { out[6]++; out[7] += task_max; }
}

void hs_stats(long long *out, int reset) { for (int i = 0; i < 8; i++) { out[i] = cct_stat[i]; if (reset) cct_stat[i] = 0; } }

int64_t hs_deflate(const uint8_t *src, int n, uint8_t *dst, int64_t cap, int64_t *stats) {
    return hs_deflate2(src, n, dst, cap, 1024, stats);
}

int hs_inflate(const uint8_t *in, int in_len, uint8_t *out, int out_cap) {
    InflateTables *tb = new InflateTables;
    std::vector<uint32_t> syms((size_t)out_cap + kInfSlack);
    int out_len = 0;
    uint32_t want = 0;
    int ns = inflate_symbols(in, in_len, syms.data(), out_cap + kInfSlack, out_cap, *tb, &out_len, &want);
    delete tb;
    if (ns < 0) return -1;
    lz_expand(syms.data(), ns, out);
    if (adler32_bytes(out, out_len) != want) return -1;
    return out_len;
}

// The lane-parallel Huffman decode of k_inflate_sym, with the kSubLanes lanes stepped one after the other: same
// tables, same decode_span, same confirmation rule.
static bool hs_block_symbols(const std::vector<uint32_t> &words, uint32_t data_end, FastTables &tb, uint32_t &q,
                             uint32_t *syms, int &ns, int &op, int sym_cap, int out_cap, int *max_rounds) {
    for (;;) {
        const uint32_t w0 = q >> 5, origin = w0 << 5;
        for (int j = 0; j < kStageWords; j++) tb.stage[stage_slot(j)] = w0 + j < words.size() ? words[w0 + j] : 0u;
        uint32_t start[kSubLanes], cur[kSubLanes], limit[kSubLanes];
        Span r[kSubLanes];
        for (int l = 0; l < kSubLanes; l++) {
            limit[l] = std::min(q + (uint32_t)(l + 1) * kSubBits, data_end);
            start[l] = q + (uint32_t)l * kSubBits; cur[l] = 0xffffffffu; r[l] = Span{0u, 0, 0, kSpanRun};
        }
        int confirmed = 0, last = -1, rounds = 0;
        static uint32_t marks[kMarkWords][kSubLanes];
        for (;;) {
            rounds++;
            for (int l = confirmed; l < kSubLanes; l++)
                if (start[l] != cur[l]) {
                    decode_span_marks<kSubLanes>(tb.stage, origin, q + (uint32_t)l * kSubBits, start[l], limit[l], data_end, tb,
                                                 &marks[0][l], rounds > 1, r[l]);
                    {  // the shortcut gives what the plain decode gives
                        Span chk;
                        decode_span<false>(tb.stage, origin, std::max(start[l], q + (uint32_t)l * kSubBits), limit[l], data_end, tb, nullptr, 0, chk);
                        if (chk.end != r[l].end || chk.n != r[l].n || chk.extra != r[l].extra || chk.flag != r[l].flag) abort();
                    }
                    cur[l] = start[l];
                }
            int c = 1;
            while (c < kSubLanes && cur[c] == r[c - 1].end) c++;
            for (int l = 0; l < c && last < 0; l++) if (r[l].flag != kSpanRun) last = l;
            if (last >= 0) break;
            if (c == kSubLanes) { last = kSubLanes - 1; break; }
            confirmed = c;
            for (int l = kSubLanes - 1; l >= c; l--) start[l] = r[l - 1].end;
        }
        if (max_rounds && rounds > *max_rounds) *max_rounds = rounds;
        const int end_flag = r[last].flag;
        if (end_flag == kSpanBad) return false;
        int tot_n = 0, tot_b = 0;
        for (int l = 0; l <= last; l++) { tot_n += r[l].n; tot_b += r[l].n + r[l].extra; }
        if (ns + tot_n > sym_cap || op + tot_b > out_cap) return false;
        int on = ns, ob = op;
        for (int l = 0; l <= last; l++) {
            Span w;
            decode_span<true>(tb.stage, origin, cur[l], limit[l], data_end, tb, syms + on, ob, w);
            if (w.flag == kSpanFar) return false;
            if (w.n != r[l].n || w.end != r[l].end) abort();
            on += r[l].n; ob += r[l].n + r[l].extra;
        }
        ns += tot_n; op += tot_b;
        q = r[last].end;
        if (end_flag == kSpanEob) return q <= data_end;
    }
}

// Whole zlib stream through the lane-parallel decoder.  `misalign` (0..3) places the stream at that offset from a
// 4-byte boundary, as streams inside a .cct arena are.  Returns the output length or -1; *rounds = most
// confirmation rounds any group of sub-sequences needed.
int hs_inflate_lanes(const uint8_t *in_, int in_len, uint8_t *out, int out_cap, int misalign, int *rounds) {
    std::vector<uint8_t> buf((size_t)in_len + 16, 0);
    memcpy(buf.data() + 4 + misalign, in_, (size_t)in_len);
    const uint8_t *in = buf.data() + 4 + misalign;
    const uint8_t *base = buf.data() + 4;  // the aligned origin (buf.data() is at least 8-byte aligned)
    const uint32_t skew = 8u * (uint32_t)misalign;
    std::vector<uint32_t> words(((size_t)in_len + misalign + 3) / 4 + 1, 0u);
    for (size_t k = 0; k < (size_t)in_len; k++) words[(k + misalign) >> 2] |= (uint32_t)in[k] << (8 * ((k + misalign) & 3));
    (void)base;
    if (in_len < 6 || (in[0] & 0x0F) != 8 || (((uint32_t)in[0] << 8) | in[1]) % 31 != 0 || (in[1] & 0x20)) return -1;
    const int dlen = in_len - 4;
    const uint32_t data_end = skew + 8u * (uint32_t)dlen;
    const int sym_cap = out_cap + kInfSlack;
    std::vector<uint32_t> syms((size_t)sym_cap);
    FastTables *tb = new FastTables;
    uint32_t bp = 16;
    int ns = 0, op = 0, mr = 0;
    bool last = false, good = true;
    while (good && !last) {
        BitReader br{in, dlen, 0, 0, 0, false};
        br.seek(bp);
        int nlen = 0, ndist = 0;
        const int type = inflate_block_header(br, *tb, &last, &nlen, &ndist, false);
        if (type < 0 || br.bad) { good = false; break; }
        bp = br.tell();
        if (type == 2) {  // as k_inflate_sym: an entry per bit, then the walk
            uint32_t *stage = tb->stage;
            const uint32_t hq = bp + skew, w0 = hq >> 5, sk = hq & 31u;
            for (int j = 0; j < kClWords; j++) stage[kClWordAt + j] = w0 + j < words.size() ? words[w0 + j] : 0u;
            const uint16_t *prim = reinterpret_cast<const uint16_t *>(stage);
            uint16_t *ent = reinterpret_cast<uint16_t *>(stage + kClEntAt);
            for (int b = 0; b < kClBits; b++) {
                const uint32_t rel = sk + (uint32_t)b;
                const uint64_t two = ((uint64_t)stage[kClWordAt + (rel >> 5) + 1] << 32) | stage[kClWordAt + (rel >> 5)];
                ent[b] = cl_entry(prim, (uint32_t)(two >> (rel & 31u)));
            }
            uint32_t used = 0;
            if (!cl_walk(ent, nlen + ndist, tb->lens, &used) || bp + used > 8u * (uint32_t)dlen) { good = false; break; }
            bp += used;
        }
        if (type == 0) {
            const int at = (int)(bp >> 3);
            if (at + nlen > dlen || ns + nlen > sym_cap || op + nlen > out_cap) { good = false; break; }
            for (int k = 0; k < nlen; k++) syms[ns + k] = in[at + k];
            ns += nlen; op += nlen; bp += 8u * (uint32_t)nlen;
            continue;
        }
        for (int pass = 0; pass < 2 && good; pass++) {
            const bool lit = pass == 0;
            FastCanon &c = lit ? tb->lc : tb->dc;
            uint32_t *sy = lit ? tb->lsym : tb->dsym;
            const uint8_t *lens = tb->lens + (lit ? 0 : nlen);
            const int n = lit ? nlen : ndist;
            for (int l = 0; l < 16; l++) c.count[l] = 0;
            fast_count(lens, n, c, 0, 1);
            if (!fast_ranges(c)) { good = false; break; }
            fast_sort(lens, n, c, sy, reinterpret_cast<uint16_t *>(tb->stage), lit);
            fast_fill(c, sy, tb->tab + (lit ? 0 : kFDstAt), lit ? kFLitBits : kFDstBits, 0, 1);
        }
        if (!good) break;
        uint32_t q = bp + skew;
        good = hs_block_symbols(words, data_end, *tb, q, syms.data(), ns, op, sym_cap, out_cap, &mr);
        bp = q - skew;
    }
    delete tb;
    if (rounds) *rounds = mr;
    if (!good || (int)((bp + 7) >> 3) != dlen) return -1;
    lz_expand(syms.data(), ns, out);
    const uint8_t *t = in + dlen;
    const uint32_t want = ((uint32_t)t[0] << 24) | ((uint32_t)t[1] << 16) | ((uint32_t)t[2] << 8) | t[3];
    if (adler32_bytes(out, op) != want) return -1;
    return op;
}

// payload (ops + EOF) -> raster pixels, via the same three steps as the decode kernels
uint32_t hs_decode_payload(const uint8_t *payload, int plen, int n_px, const int32_t *order, uint16_t *out) {
    int nb = n_px / kBlock;
    std::vector<uint16_t> vals((size_t)n_px);
    std::vector<uint8_t> jflag((size_t)nb, 0);
    std::vector<int32_t> slotA((size_t)nb), slotB((size_t)nb);
    std::vector<OpsChunkSum> cs((size_t)plen / kOpsChunk + 2);
    uint32_t st = ops_to_values(payload, plen, n_px, vals.data(), jflag.data(), cs.data());
    if (st & 64u) return st;
    st |= resolve_units(jflag.data(), nb, slotA.data(), slotB.data());
    if (st & 64u) return st;
    for (int k = 0; k < n_px; k++) out[op_raster(k, slotA.data(), slotB.data(), order)] = vals[k];
    return st;
}

uint32_t hs_hilbert(uint32_t n, uint32_t d) { return hilbert_d2raster(n, d); }
int64_t hs_window_base(int64_t t, int64_t n) { return window_base_at(t, n); }
int hs_len_code(int lc) { return len_code(lc); }
int hs_dist_code(int d) { return dist_code(d); }

}  // extern "C"
