/*
 * oracle/deflate9_port.c — from-scratch restatement of zlib 1.3 `compress2(..., level 9)`.
 * TEST INFRASTRUCTURE ONLY (same rule as cct_oracle.c: never on the product path).
 *
 * The reference's entropy stage is `zlib.compress(data, level = 9)` (core.py:340), i.e. the system
 * libz (1.3, third-party, not vendored).  The CUDA path has to reproduce
 * libz's output bit for bit, so this file restates libz's published algorithm (deflate_slow lazy
 * parse with hash chains, the 16 383-symbol block flush, build_tree/gen_bitlen/gen_codes,
 * scan_tree/send_tree, block-type choice, Adler-32) in sequential C and exposes the intermediate
 * results the GPU kernels are compared against stage by stage (match candidates, symbol stream,
 * block boundaries).  It is itself pinned against libz: tests/test_oracle.py runs it against
 * compress2 on the golden payload and a fuzz corpus (stored / static / dynamic / multi-block /
 * end-of-input block-boundary cases).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define WSIZE 32768
#define MAX_DIST (WSIZE - 262)
#define MIN_MATCH 3
#define MAX_MATCH 258
#define TOO_FAR 4096
#define HASH_MASK 0x7FFF
#define SYM_END 16383 /* (lit_bufsize-1) with memLevel 8 */

#define L_CODES 286
#define D_CODES 30
#define BL_CODES 19
#define HEAP_SIZE (2 * L_CODES + 1)
#define MAX_BITS 15

static const int extra_lbits[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const int extra_dbits[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const int extra_blbits[19] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 3, 7};
static const uint8_t bl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

static uint8_t length_code[256];
static int base_length[29];
static uint8_t dist_code_tab[512];
static int base_dist[30];
static uint16_t static_lcode[288], static_dcode[30];
static uint8_t static_llen[288];
static int tables_ready = 0;

static unsigned bi_reverse(unsigned code, int len) {
    unsigned res = 0;
    do { res |= code & 1; code >>= 1; res <<= 1; } while (--len > 0);
    return res >> 1;
}

static void init_tables(void) {
    if (tables_ready) return;
    int length = 0, code, n, dist = 0;
    for (code = 0; code < 28; code++) {
        base_length[code] = length;
        for (n = 0; n < (1 << extra_lbits[code]); n++) length_code[length++] = (uint8_t)code;
    }
    length_code[length - 1] = (uint8_t)code; /* len 258 -> code 28 */
    base_length[28] = 0;
    for (code = 0; code < 16; code++) {
        base_dist[code] = dist;
        for (n = 0; n < (1 << extra_dbits[code]); n++) dist_code_tab[dist++] = (uint8_t)code;
    }
    dist >>= 7;
    for (; code < D_CODES; code++) {
        base_dist[code] = dist << 7;
        for (n = 0; n < (1 << (extra_dbits[code] - 7)); n++) dist_code_tab[256 + dist++] = (uint8_t)code;
    }
    int bl_count[MAX_BITS + 1] = {0};
    for (n = 0; n <= 143; n++) static_llen[n] = 8, bl_count[8]++;
    for (; n <= 255; n++) static_llen[n] = 9, bl_count[9]++;
    for (; n <= 279; n++) static_llen[n] = 7, bl_count[7]++;
    for (; n <= 287; n++) static_llen[n] = 8, bl_count[8]++;
    unsigned next_code[MAX_BITS + 1], c = 0;
    for (int bits = 1; bits <= MAX_BITS; bits++) { c = (c + bl_count[bits - 1]) << 1; next_code[bits] = c; }
    for (n = 0; n < 288; n++) static_lcode[n] = (uint16_t)bi_reverse(next_code[static_llen[n]]++, static_llen[n]);
    for (n = 0; n < D_CODES; n++) static_dcode[n] = (uint16_t)bi_reverse((unsigned)n, 5);
    tables_ready = 1;
}

static inline int d_code(int dist) { return dist < 256 ? dist_code_tab[dist] : dist_code_tab[256 + (dist >> 7)]; }

/* ---------------------------------------------------------------- bit writer */
typedef struct {
    uint8_t *out; int64_t pos, cap; uint64_t buf; int valid; int overflow;
} bitw;

static void put_byte(bitw *w, int b) {
    if (w->pos < w->cap) w->out[w->pos++] = (uint8_t)b; else w->overflow = 1;
}
static void send_bits(bitw *w, unsigned value, int length) {
    w->buf |= (uint64_t)value << w->valid;
    w->valid += length;
    while (w->valid >= 8) { put_byte(w, (int)(w->buf & 0xFF)); w->buf >>= 8; w->valid -= 8; }
}
static void bi_windup(bitw *w) {
    if (w->valid > 0) put_byte(w, (int)(w->buf & 0xFF));
    w->buf = 0; w->valid = 0;
}

/* ---------------------------------------------------------------- Huffman trees (trees.c) */
typedef struct {
    uint16_t freq[HEAP_SIZE];
    uint16_t dad[HEAP_SIZE];
    uint16_t len[HEAP_SIZE];
    uint16_t code[HEAP_SIZE];
    int max_code;
} tree_t;

typedef struct {
    int heap[HEAP_SIZE]; int heap_len, heap_max;
    uint8_t depth[HEAP_SIZE];
    uint16_t bl_count[MAX_BITS + 1];
    uint64_t opt_len, static_len;
} tstate;

#define smaller(t, n, m, depth) \
    ((t)->freq[n] < (t)->freq[m] || ((t)->freq[n] == (t)->freq[m] && depth[n] <= depth[m]))

static void pqdownheap(tstate *s, tree_t *t, int k) {
    int v = s->heap[k], j = k << 1;
    while (j <= s->heap_len) {
        if (j < s->heap_len && smaller(t, s->heap[j + 1], s->heap[j], s->depth)) j++;
        if (smaller(t, v, s->heap[j], s->depth)) break;
        s->heap[k] = s->heap[j]; k = j; j <<= 1;
    }
    s->heap[k] = v;
}

static void build_tree(tstate *s, tree_t *t, int elems, const uint8_t *stat_len, const int *extra, int base,
                       int max_length) {
    int n, m, max_code = -1, node;
    s->heap_len = 0; s->heap_max = HEAP_SIZE;
    for (n = 0; n < elems; n++) {
        if (t->freq[n] != 0) { s->heap[++s->heap_len] = max_code = n; s->depth[n] = 0; }
        else t->len[n] = 0;
    }
    while (s->heap_len < 2) {
        node = s->heap[++s->heap_len] = (max_code < 2 ? ++max_code : 0);
        t->freq[node] = 1; s->depth[node] = 0;
        s->opt_len--;
        if (stat_len) s->static_len -= stat_len[node];
    }
    t->max_code = max_code;
    for (n = s->heap_len / 2; n >= 1; n--) pqdownheap(s, t, n);
    node = elems;
    do {
        n = s->heap[1]; s->heap[1] = s->heap[s->heap_len--]; pqdownheap(s, t, 1);
        m = s->heap[1];
        s->heap[--s->heap_max] = n; s->heap[--s->heap_max] = m;
        t->freq[node] = (uint16_t)(t->freq[n] + t->freq[m]);
        s->depth[node] = (uint8_t)((s->depth[n] >= s->depth[m] ? s->depth[n] : s->depth[m]) + 1);
        t->dad[n] = t->dad[m] = (uint16_t)node;
        s->heap[1] = node++;
        pqdownheap(s, t, 1);
    } while (s->heap_len >= 2);
    s->heap[--s->heap_max] = s->heap[1];

    /* gen_bitlen */
    int h, bits, xbits, overflow = 0;
    for (bits = 0; bits <= MAX_BITS; bits++) s->bl_count[bits] = 0;
    t->len[s->heap[s->heap_max]] = 0;
    for (h = s->heap_max + 1; h < HEAP_SIZE; h++) {
        n = s->heap[h];
        bits = t->len[t->dad[n]] + 1;
        if (bits > max_length) { bits = max_length; overflow++; }
        t->len[n] = (uint16_t)bits;
        if (n > max_code) continue;
        s->bl_count[bits]++;
        xbits = 0;
        if (n >= base) xbits = extra[n - base];
        s->opt_len += (uint64_t)t->freq[n] * (unsigned)(bits + xbits);
        if (stat_len) s->static_len += (uint64_t)t->freq[n] * (unsigned)(stat_len[n] + xbits);
    }
    if (overflow) {
        do {
            bits = max_length - 1;
            while (s->bl_count[bits] == 0) bits--;
            s->bl_count[bits]--; s->bl_count[bits + 1] += 2; s->bl_count[max_length]--;
            overflow -= 2;
        } while (overflow > 0);
        for (bits = max_length; bits != 0; bits--) {
            n = s->bl_count[bits];
            while (n != 0) {
                m = s->heap[--h];
                if (m > max_code) continue;
                if (t->len[m] != (unsigned)bits) {
                    s->opt_len += ((uint64_t)bits - t->len[m]) * t->freq[m];
                    t->len[m] = (uint16_t)bits;
                }
                n--;
            }
        }
    }
    /* gen_codes */
    unsigned next_code[MAX_BITS + 1], code = 0;
    for (bits = 1; bits <= MAX_BITS; bits++) { code = (code + s->bl_count[bits - 1]) << 1; next_code[bits] = code & 0xFFFF; }
    for (n = 0; n <= max_code; n++) {
        int len = t->len[n];
        if (len == 0) continue;
        t->code[n] = (uint16_t)bi_reverse(next_code[len]++, len);
    }
}

/* scan_tree (bl == NULL -> count into blfreq) / send_tree (emit) share one control flow */
static void tree_rle(const tree_t *t, int max_code, uint16_t *blfreq, const tree_t *bl, bitw *w) {
    int prevlen = -1, curlen, nextlen = t->len[0], count = 0, max_count = 7, min_count = 4;
    if (nextlen == 0) { max_count = 138; min_count = 3; }
    for (int n = 0; n <= max_code; n++) {
        curlen = nextlen;
        nextlen = (n + 1 <= max_code) ? t->len[n + 1] : 0xFFFF;
        if (++count < max_count && curlen == nextlen) continue;
        else if (count < min_count) {
            if (blfreq) blfreq[curlen] += (uint16_t)count;
            else do { send_bits(w, bl->code[curlen], bl->len[curlen]); } while (--count != 0);
        } else if (curlen != 0) {
            if (curlen != prevlen) {
                if (blfreq) blfreq[curlen]++;
                else { send_bits(w, bl->code[curlen], bl->len[curlen]); count--; }
            }
            if (blfreq) blfreq[16]++;
            else { send_bits(w, bl->code[16], bl->len[16]); send_bits(w, (unsigned)(count - 3), 2); }
        } else if (count <= 10) {
            if (blfreq) blfreq[17]++;
            else { send_bits(w, bl->code[17], bl->len[17]); send_bits(w, (unsigned)(count - 3), 3); }
        } else {
            if (blfreq) blfreq[18]++;
            else { send_bits(w, bl->code[18], bl->len[18]); send_bits(w, (unsigned)(count - 11), 7); }
        }
        count = 0; prevlen = curlen;
        if (nextlen == 0) { max_count = 138; min_count = 3; }
        else if (curlen == nextlen) { max_count = 6; min_count = 3; }
        else { max_count = 7; min_count = 4; }
    }
}

typedef struct { uint16_t dist; uint8_t lc; } sym_t;

/* instrumentation shared with the tests */
typedef struct {
    int64_t n_blocks, n_stored, n_static, n_dynamic, n_literals, n_matches, n_searches, n_candidates;
} d9_stats;
static d9_stats g_stats;

static void compress_block(bitw *w, const sym_t *sym, int nsym, const uint16_t *lcode, const uint16_t *llen_or_null,
                           const uint8_t *sllen, const uint16_t *dcode, const uint16_t *dlen_or_null) {
    for (int i = 0; i < nsym; i++) {
        unsigned dist = sym[i].dist, lc = sym[i].lc;
        if (dist == 0) {
            send_bits(w, lcode[lc], llen_or_null ? llen_or_null[lc] : sllen[lc]);
        } else {
            unsigned code = length_code[lc];
            send_bits(w, lcode[code + 257], llen_or_null ? llen_or_null[code + 257] : sllen[code + 257]);
            int extra = extra_lbits[code];
            if (extra) send_bits(w, lc - (unsigned)base_length[code], extra);
            dist--;
            code = (unsigned)d_code((int)dist);
            send_bits(w, dcode[code], dlen_or_null ? dlen_or_null[code] : 5);
            extra = extra_dbits[code];
            if (extra) send_bits(w, dist - (unsigned)base_dist[code], extra);
        }
    }
    send_bits(w, lcode[256], llen_or_null ? llen_or_null[256] : sllen[256]);
}

/* _tr_flush_block */
static void flush_block(bitw *w, const sym_t *sym, int nsym, const uint8_t *buf, int64_t stored_len, int last) {
    static const uint8_t static_dlen[30] = {5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5};
    tree_t lt, dt, bt;
    tstate s;
    memset(&lt, 0, sizeof lt); memset(&dt, 0, sizeof dt); memset(&bt, 0, sizeof bt); memset(&s, 0, sizeof s);
    lt.freq[256] = 1;
    for (int i = 0; i < nsym; i++) {
        if (sym[i].dist == 0) lt.freq[sym[i].lc]++;
        else { lt.freq[length_code[sym[i].lc] + 257]++; dt.freq[d_code(sym[i].dist - 1)]++; }
    }
    build_tree(&s, &lt, L_CODES, static_llen, extra_lbits, 257, MAX_BITS);
    build_tree(&s, &dt, D_CODES, static_dlen, extra_dbits, 0, MAX_BITS);
    tree_rle(&lt, lt.max_code, bt.freq, NULL, NULL);
    tree_rle(&dt, dt.max_code, bt.freq, NULL, NULL);
    build_tree(&s, &bt, BL_CODES, NULL, extra_blbits, 0, 7);
    int max_blindex;
    for (max_blindex = BL_CODES - 1; max_blindex >= 3; max_blindex--)
        if (bt.len[bl_order[max_blindex]] != 0) break;
    s.opt_len += 3 * ((uint64_t)max_blindex + 1) + 5 + 5 + 4;
    uint64_t opt_lenb = (s.opt_len + 3 + 7) >> 3, static_lenb = (s.static_len + 3 + 7) >> 3;
    if (static_lenb <= opt_lenb) opt_lenb = static_lenb;
    g_stats.n_blocks++;
    if ((uint64_t)stored_len + 4 <= opt_lenb && buf) {
        g_stats.n_stored++;
        send_bits(w, (0 << 1) + (unsigned)last, 3);
        bi_windup(w);
        put_byte(w, (int)(stored_len & 0xFF)); put_byte(w, (int)((stored_len >> 8) & 0xFF));
        put_byte(w, (int)(~stored_len & 0xFF)); put_byte(w, (int)((~stored_len >> 8) & 0xFF));
        for (int64_t i = 0; i < stored_len; i++) put_byte(w, buf[i]);
    } else if (static_lenb == opt_lenb) {
        g_stats.n_static++;
        send_bits(w, (1 << 1) + (unsigned)last, 3);
        compress_block(w, sym, nsym, static_lcode, NULL, static_llen, static_dcode, NULL);
    } else {
        g_stats.n_dynamic++;
        send_bits(w, (2 << 1) + (unsigned)last, 3);
        send_bits(w, (unsigned)(lt.max_code + 1 - 257), 5);
        send_bits(w, (unsigned)(dt.max_code + 1 - 1), 5);
        send_bits(w, (unsigned)(max_blindex + 1 - 4), 4);
        for (int r = 0; r <= max_blindex; r++) send_bits(w, bt.len[bl_order[r]], 3);
        tree_rle(&lt, lt.max_code, NULL, &bt, w);
        tree_rle(&dt, dt.max_code, NULL, &bt, w);
        compress_block(w, sym, nsym, lt.code, lt.len, NULL, dt.code, dt.len);
    }
    if (last) bi_windup(w);
}

/* ---------------------------------------------------------------- LZ77 lazy parse (deflate_slow) */

/* Optional trace of the parse: one record per emitted symbol. */
/* Start of zlib's 64 KiB window when deflate_slow's loop stands at position t (deflate.c fill_window): the window
 * slides by 32 KiB at the first loop top where fewer than MIN_LOOKAHEAD (262) bytes are left in it -- or, once the
 * input is exhausted, at every loop top -- and strstart has reached wsize + MAX_DIST (65274) inside it.  A block whose
 * first byte lies before the window (block_start < 0) cannot be emitted as a stored block (_tr_flush_block: buf == NULL). */
static int64_t window_base_at(int64_t t, int64_t n) {
    int64_t k = t >= 65275 ? (t - 65275) / 32768 + 1 : 0; /* slides that have certainly happened */
    for (;; k++) {
        int64_t base = 32768 * k;
        int64_t thr = n >= base + 65536 ? base + 65275 : (base + 65274 > n - 261 ? base + 65274 : n - 261);
        if (t < thr) return base;
    }
}

typedef struct { int32_t pos; uint16_t dist; uint16_t len; } d9_symrec; /* len==0: literal at pos */

static int64_t lz_parse(const uint8_t *src, int64_t n, bitw *w, d9_symrec *trace, int64_t trace_cap, int64_t *n_trace) {
    int32_t *head = (int32_t *)calloc(32768, sizeof(int32_t));
    int32_t *prev = (int32_t *)calloc((size_t)(n > 0 ? n : 1), sizeof(int32_t));
    sym_t *sym = (sym_t *)malloc(sizeof(sym_t) * SYM_END);
    int nsym = 0;
    int64_t total_sym = 0, block_start = 0, nt = 0;
    int64_t s = 0;
    int match_len = 2, avail = 0;
    int64_t match_start = 0;
#define HASH(p) ((((unsigned)src[p] << 10) ^ ((unsigned)src[(p) + 1] << 5) ^ (unsigned)src[(p) + 2]) & HASH_MASK)
#define INSERT(p, hh) do { unsigned k_ = HASH(p); (hh) = head[k_]; prev[p] = (hh); head[k_] = (int32_t)(p); } while (0)
#define TRACE(p, d, l) do { if (trace && nt < trace_cap) { trace[nt].pos = (int32_t)(p); trace[nt].dist = (uint16_t)(d); trace[nt].len = (uint16_t)(l); } nt++; } while (0)
    while (s < n) {
        const int64_t top = s; /* strstart at the loop top: fill_window runs here, nowhere else in the iteration */
        int64_t look = n - s;
        int32_t hh = 0;
        if (look >= MIN_MATCH) INSERT(s, hh);
        int prev_len = match_len;
        int64_t prev_match = match_start;
        match_len = MIN_MATCH - 1;
        if (hh != 0 && prev_len < MAX_MATCH && s - hh <= MAX_DIST) {
            /* longest_match */
            g_stats.n_searches++;
            int best = prev_len;
            int chain = prev_len >= 32 ? 1024 : 4096;
            int cap = look < MAX_MATCH ? (int)look : MAX_MATCH;
            int64_t limit = s > MAX_DIST ? s - MAX_DIST : 0;
            int64_t cur = hh;
            if (best < cap) { /* best >= cap: nothing can improve, result is min(best, look) */
                do {
                    g_stats.n_candidates++;
                    const uint8_t *a = src + cur, *b = src + s;
                    if (a[best] == b[best] && a[best - 1] == b[best - 1] && a[0] == b[0] && a[1] == b[1]) {
                        int len = 2;
                        while (len < cap && a[len] == b[len]) len++;
                        if (len > best) {
                            match_start = cur; best = len;
                            if (len >= cap) break;
                        }
                    }
                    cur = prev[cur];
                } while (cur > limit && --chain != 0);
            }
            match_len = best < look ? best : (int)look;
            if (match_len == MIN_MATCH && s - match_start > TOO_FAR) match_len = MIN_MATCH - 1;
        }
        if (prev_len >= MIN_MATCH && match_len <= prev_len) {
            int64_t max_insert = s + look - MIN_MATCH;
            sym[nsym].dist = (uint16_t)(s - 1 - prev_match); sym[nsym].lc = (uint8_t)(prev_len - MIN_MATCH); nsym++;
            TRACE(s - 1, s - 1 - prev_match, prev_len);
            g_stats.n_matches++;
            int reps = prev_len - 2;
            do {
                if (++s <= max_insert) { int32_t t_; INSERT(s, t_); (void)t_; }
            } while (--reps != 0);
            avail = 0; match_len = MIN_MATCH - 1; s++;
            if (nsym == SYM_END) {
                flush_block(w, sym, nsym, block_start >= window_base_at(top, n) ? src + block_start : NULL, s - block_start, 0);
                total_sym += nsym; nsym = 0; block_start = s;
            }
        } else if (avail) {
            sym[nsym].dist = 0; sym[nsym].lc = src[s - 1]; nsym++;
            TRACE(s - 1, 0, 0);
            g_stats.n_literals++;
            if (nsym == SYM_END) {
                flush_block(w, sym, nsym, block_start >= window_base_at(top, n) ? src + block_start : NULL, s - block_start, 0);
                total_sym += nsym; nsym = 0; block_start = s;
            }
            s++;
        } else {
            avail = 1; s++;
        }
    }
    if (avail) { /* the trailing literal never triggers a flush */
        sym[nsym].dist = 0; sym[nsym].lc = src[n - 1]; nsym++;
        TRACE(n - 1, 0, 0);
        g_stats.n_literals++;
    }
    flush_block(w, sym, nsym, block_start >= window_base_at(n, n) ? src + block_start : NULL, n - block_start, 1);
    total_sym += nsym;
    if (n_trace) *n_trace = nt;
    free(head); free(prev); free(sym);
    return total_sym;
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

/* zlib stream of src[0..n) exactly as compress2(dst,&len,src,n,9) would write it.
 * Returns stream length or -1 if cap is too small. */
int64_t d9_compress(const uint8_t *src, int64_t n, uint8_t *dst, int64_t cap, int64_t *nsym_out) {
    init_tables();
    memset(&g_stats, 0, sizeof g_stats);
    bitw w = {dst, 0, cap, 0, 0, 0};
    put_byte(&w, 0x78); put_byte(&w, 0xDA);
    int64_t ns = lz_parse(src, n, &w, NULL, 0, NULL);
    uint32_t ad = adler32_(src, n);
    put_byte(&w, (int)(ad >> 24)); put_byte(&w, (int)((ad >> 16) & 0xFF)); put_byte(&w, (int)((ad >> 8) & 0xFF)); put_byte(&w, (int)(ad & 0xFF));
    if (nsym_out) *nsym_out = ns;
    return w.overflow ? -1 : w.pos;
}

/* Same, additionally returning the LZ77 symbol trace (pos, dist, len; len == 0 means literal). */
int64_t d9_compress_trace(const uint8_t *src, int64_t n, uint8_t *dst, int64_t cap, d9_symrec *trace, int64_t trace_cap,
                          int64_t *n_trace) {
    init_tables();
    memset(&g_stats, 0, sizeof g_stats);
    bitw w = {dst, 0, cap, 0, 0, 0};
    put_byte(&w, 0x78); put_byte(&w, 0xDA);
    lz_parse(src, n, &w, trace, trace_cap, n_trace);
    uint32_t ad = adler32_(src, n);
    put_byte(&w, (int)(ad >> 24)); put_byte(&w, (int)((ad >> 16) & 0xFF)); put_byte(&w, (int)((ad >> 8) & 0xFF)); put_byte(&w, (int)(ad & 0xFF));
    return w.overflow ? -1 : w.pos;
}

void d9_last_stats(d9_stats *out) { *out = g_stats; }
