// deflate_core.cuh — the scalar building blocks of the bit-exact zlib-1.3 level-9 encoder, written as
// host/device functions so the CUDA kernels (cct_deflate.cu) and the CPU test harness
// (tests/hostsim) run the very same code.
//
// What is restated: the reference's entropy stage is `zlib.compress(data, level = 9)`
// (reference src/codec/core.py:340) = system libz 1.3 `deflate_slow` + `trees.c` (third-party, not in the
// reference tree; algorithm recorded in SURVEY.md Appendix B).  The decomposition used here:
//   1. hash chains are parse-independent (every position is inserted exactly once, in order), so
//      `dist16[s]` = distance to the previous position with the same 3-byte hash is a pure function
//      of the data (kernel k_chains);
//   2. `match_probe` runs longest_match from a cold start (best = 2) over the first K chain entries
//      for EVERY position in parallel and records (len, dist, complete?) in a table;
//   3. zlib's lazy-evaluation state machine is walked over symbol starts: speculatively per window in
//      parallel (`walk_window`), then stitched per stream (`walk_stitch`); where the probe was incomplete
//      it falls back to `match_full`, the exact longest_match loop;
//   4. `build_block` restates _tr_flush_block's tree construction and block-type choice;
//   5. the symbol bits are then packed in parallel from the code tables.
#pragma once
#include "cct_common.cuh"

// Work counters for the CPU harness (tests/hostsim defines CCT_STATS); compiled out of the kernels.
#if defined(CCT_STATS) && !defined(__CUDA_ARCH__)
extern long long cct_stat[8];
#define CCT_COUNT(i) (cct_stat[i]++)
#else
#define CCT_COUNT(i) ((void)0)
#endif

namespace cct {

// ------------------------------------------------------------------------------------------------
// symbol / table encodings
// ------------------------------------------------------------------------------------------------
// mtab entry: bits 0..8 best length (0 = no match), bits 9..23 distance, bit 31 = probe incomplete
constexpr uint32_t kProbeIncomplete = 0x80000000u;
CCT_HD uint32_t mtab_pack(int len, int dist) { return (uint32_t)len | ((uint32_t)dist << 9); }
CCT_HD int mtab_len(uint32_t e) { return (int)(e & 511u); }
CCT_HD int mtab_dist(uint32_t e) { return (int)((e >> 9) & 32767u); }

// LZ77 symbol: bits 0..7 literal byte or (len-3), bits 8..23 distance (0 = literal)
CCT_HD uint32_t sym_pack(int lc, int dist) { return (uint32_t)lc | ((uint32_t)dist << 8); }
CCT_HD int sym_lc(uint32_t s) { return (int)(s & 255u); }
CCT_HD int sym_dist(uint32_t s) { return (int)(s >> 8); }

CCT_HD uint32_t hash3(const uint8_t *p) {  // UPDATE_HASH x3 with hash_shift 5, 15-bit mask
    return (((uint32_t)p[0] << 10) ^ ((uint32_t)p[1] << 5) ^ (uint32_t)p[2]) & 0x7FFFu;
}

// length code / extra bits of lc = len-3 (trees.c _length_code, base_length, extra_lbits)
CCT_HD int len_code(int lc) {
    if (lc < 8) return lc;
    if (lc == 255) return 28;
    int m = ilog2((uint32_t)lc);
    return 4 * (m - 2) + 4 + ((lc >> (m - 2)) & 3);
}
CCT_HD int len_extra_bits(int code) { return (code < 8 || code == 28) ? 0 : (code >> 2) - 1; }
// distance code / extra bits of d = dist-1 (trees.c _dist_code, base_dist, extra_dbits)
CCT_HD int dist_code(int d) {
    if (d < 4) return d;
    int m = ilog2((uint32_t)d);
    return 2 * m + ((d >> (m - 1)) & 1);
}
CCT_HD int dist_extra_bits(int code) { return code < 4 ? 0 : (code >> 1) - 1; }
CCT_HD int bl_extra_bits(int code) { return code < 16 ? 0 : (code == 16 ? 2 : (code == 17 ? 3 : 7)); }

CCT_HD uint32_t bit_reverse(uint32_t code, int len) {  // trees.c bi_reverse: the low `len` bits of code, reversed
#if defined(__CUDA_ARCH__)
    return len ? __brev(code) >> (32 - len) : 0u;
#else
    uint32_t res = 0;
    for (int i = 0; i < len; i++) { res = (res << 1) | (code & 1u); code >>= 1; }
    return res;
#endif
}
// static trees (trees.c tr_static_init): literal/length lengths and bit-reversed codes
CCT_HD int static_llen(int n) { return n <= 143 ? 8 : (n <= 255 ? 9 : (n <= 279 ? 7 : 8)); }
CCT_HD uint32_t static_lcode(int n) {
    uint32_t c = n <= 143 ? 0x30u + (uint32_t)n
                          : (n <= 255 ? 0x190u + (uint32_t)(n - 144) : (n <= 279 ? (uint32_t)(n - 256) : 0xC0u + (uint32_t)(n - 280)));
    return bit_reverse(c, static_llen(n));
}

// ------------------------------------------------------------------------------------------------
// longest_match
// ------------------------------------------------------------------------------------------------
// Everything the chain walk needs to know about a position, in one aligned 8-byte load.  (The walk is bound by
// the instructions it issues and the distinct memory sectors its divergent loads touch, so a record carries enough
// of the payload that the common candidate -- same hash, at most three matching bytes -- never touches the payload.)
//   a.lo16  link: distance to the previous position with the same hash, 0 = none / >= 32768.  For a run position
//           (rl >= 2) it is the link of the BOTTOM of its run group, position max(p - rb, 1): the position's own link
//           is then implied (1 inside a run, because p-1 starts the same triple; 0 at position 1) and the group step
//           needs the bottom's.
//   a.hi16  rl: number of bytes equal to w[p] immediately after p (capped at 65535)
//   b       rl <  2: w[p] | w[p+1] << 8 | w[p+2] << 16 | w[p+3] << 24   (0 past the end)
//           rl >= 2: rb | w[p] << 16 | ba << 24 with rb = number of bytes equal to w[p] immediately before p (capped at
//                    65535) and ba = w[p + rl + 1], the first byte after the run (0 past the end or when rl is capped)
struct alignas(8) ChainRec { uint32_t a, b; };
CCT_HD int rec_link(const ChainRec &r) { return (int)(r.a & 0xFFFFu); }
CCT_HD int rec_rl(const ChainRec &r) { return (int)(r.a >> 16); }
CCT_HD bool rec_is_run(const ChainRec &r) { return r.a >= 0x20000u; }
CCT_HD int rec_rb(const ChainRec &r) { return (int)(r.b & 0xFFFFu); }            // run records only
CCT_HD uint32_t rec_run_byte(const ChainRec &r) { return (r.b >> 16) & 0xFFu; }  // run records only
CCT_HD uint32_t rec_after(const ChainRec &r) { return r.b >> 24; }               // run records only
// the position's own chain link
CCT_HD int rec_own_dist(const ChainRec &r, int p) { return (rec_is_run(r) && rec_rb(r)) ? (p >= 2 ? 1 : 0) : rec_link(r); }
// the first four payload bytes at the position (byte 3 of a run record: the run goes on, or the byte after it)
CCT_HD uint32_t rec_quad(const ChainRec &r) {
    if (!rec_is_run(r)) return r.b;
    const uint32_t c = rec_run_byte(r);
    return c * 0x010101u | ((rec_rl(r) >= 3 ? c : rec_after(r)) << 24);
}
CCT_HD ChainRec pack_rec_plain(int link, int rl, uint32_t quad) { return ChainRec{(uint32_t)link | ((uint32_t)rl << 16), quad}; }
CCT_HD ChainRec pack_rec_run(int bottom_link, int rl, int rb, uint32_t c, uint32_t after) {
    return ChainRec{(uint32_t)bottom_link | ((uint32_t)rl << 16), (uint32_t)rb | (c << 16) | (after << 24)};
}
// The record of position p from the plain arrays (dist16 = own links, rl / rb = capped run counts).
CCT_HD ChainRec make_rec(const uint8_t *w, int n, const uint16_t *dist16, int p, int rl, int rb) {
    if (rl >= 2) {
        const int bottom = p - rb < 1 ? 1 : p - rb;
        const int q = p + rl + 1;
        return pack_rec_run(dist16[bottom], rl, rb, w[p], (rl < 65535 && q < n) ? w[q] : 0u);
    }
    uint32_t quad = 0;
    for (int i = 0; i < 4; i++) if (p + i < n) quad |= (uint32_t)w[p + i] << (8 * i);
    return pack_rec_plain(dist16[p], rl, quad);
}

CCT_HD void store_rec(ChainRec *p, const ChainRec &r) {  // one 8-byte store
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint2 *>(p) = make_uint2(r.a, r.b);
#else
    *p = r;
#endif
}

CCT_HD ChainRec load_rec(const ChainRec *p) {
#if defined(__CUDA_ARCH__)
    const uint2 q = __ldg(reinterpret_cast<const uint2 *>(p));
    return ChainRec{q.x, q.y};
#else
    return *p;
#endif
}

struct LzView {
    const uint8_t *w;       // payload bytes
    const ChainRec *recs;   // per-position record (k_run_recs)
    int n;                  // payload length
};

CCT_HD uint8_t load_byte(const uint8_t *p) {  // payload bytes are read-only for the whole deflate stage
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
CCT_HD int match_extend(const uint8_t *a, const uint8_t *b, int from, int cap) {
    int len = from;
    while (len < cap && load_byte(a + len) == load_byte(b + len)) len++;
    return len;
}

// How a chain walk ended.
constexpr int kWalkDone = 0;    // chain end, distance limit or a match of the maximum useful length
constexpr int kWalkBudget = 1;  // chain budget exhausted
constexpr int kWalkSteps = 2;   // step cap of the probe reached

// longest_match's candidate loop (deflate.c), walking the chain of position s from its head.
//   best / bdist   in: best_len to beat (prev_length, or 2 for a cold start); out: result
//   chain          number of candidates that may still be visited (max_chain_length, quartered by the caller)
//   max_steps      cap on LOOP STEPS (not candidates) for the parallel probe
// One loop step handles either a single candidate or a whole RUN GROUP.  When the candidate `cur` starts with
// a triple (c,c,c), the chain entries below cur inside that run are exactly the consecutive positions cur-1,
// cur-2, ... down to the run start (all have the hash of (c,c,c)).  If the target does not itself start with
// (c,c,c) -- the chain only shares the hash -- none of them matches 3 bytes.  If it starts with R >= 3 copies
// of c, the match length of the i-th member is known in closed form from its count of leading c's, m + i:
// m+i if m+i < R, "R + whatever follows" if m+i == R, exactly R if m+i > R (R < cap), cap if R == cap.
// Since only the first-visited candidate attaining the maximum matters, the group is consumed in
// one step; the chain budget and the distance limit are charged as if every member had been visited.  "Whatever
// follows" starts with the byte after either run, which both records carry: the payload is read only when they agree.
// A single candidate is compared through the four payload bytes in its record: a candidate that merely shares the
// hash, or matches exactly three bytes, is settled without touching the payload.
// The walk is resumable (chain_begin / chain_run): match_probe raises the budget of a walk that ran out of it.
struct ChainState {
    int s, cap, best, bdist, chain, steps_left, cur, R;
    uint32_t quad;   // target bytes 0..3
    uint32_t after;  // R != 0: the byte after the target's run (meaningful when R < cap)
};

// Returns the target's own chain link (0: no chain).
CCT_HD int chain_begin(const LzView &v, int s, int cap, int best, int chain, int max_steps, ChainState &c) {
    const ChainRec rs = load_rec(v.recs + s);
    c.s = s; c.cap = cap; c.best = best; c.bdist = 0; c.chain = chain; c.steps_left = max_steps;
    c.quad = rec_quad(rs);
    c.R = 0; c.after = 0;  // leading run of the target, if it is at least one hash triple long
    if (rec_is_run(rs)) { c.R = rec_rl(rs) + 1; if (c.R > cap) c.R = cap; c.after = rec_after(rs); }
    const int own = rec_own_dist(rs, s);
    c.cur = s - own;
    return own;
}

// Runs the walk until it ends (one loop step = a single candidate or a whole run group) and returns how it ended;
// the state is left such that a walk cut short by the budget or the step cap can go on (c.chain / c.steps_left
// raised by the caller).  The chain budget and the step cap both drop by one per step, so the loop carries the step
// count and the (constant) difference: one counter, one exit test.
CCT_HD int chain_run(const LzView &v, ChainState &c) {
    const int s = c.s, cap = c.cap;
    int cur = c.cur, best = c.best, bdist = c.bdist, steps = c.steps_left;
    int diff = c.chain - steps;       // chain = steps + diff
    int lim = diff < 0 ? -diff : 0;   // steps == lim: the budget (diff < 0) or the step cap ran out
    int r;
    for (;;) {
        CCT_COUNT(0);
        const ChainRec rc = load_rec(v.recs + cur);  // one 8-byte load: link, run lengths, first bytes of the candidate
        const int dl = rec_link(rc);
        int link_from = cur;
        if (rec_is_run(rc)) {
            // cur starts a triple (c,c,c): the chain entries below it inside the run form one group
            const int chain = steps + diff;
            const int m = rec_rl(rc) + 1;  // leading c's of the top candidate
            int t = rec_rb(rc);            // further candidates of this run below cur
            if (cur - t < 1) t = cur - 1;  // position 0 is never in a chain
            const int count_dist = kMaxDist - (s - cur);
            int nvis = t + 1;
            if (chain < nvis) nvis = chain;
            if (count_dist < nvis) nvis = count_dist;
            if (c.R && rec_run_byte(rc) == (c.quad & 0xFFu)) {
                const int R = c.R;
                const int i_key = m >= R ? 0 : R - m;  // first member with at least R leading c's
                if (R < cap && m > R) {
                    if (R > best) { best = R; bdist = s - cur; }
                } else if (i_key < nvis) {
                    const int c_al = cur - i_key;  // R < cap: exactly R leading c's, then the byte after the group's run
                    int L = cap;
                    if (R < cap) L = rec_after(rc) != c.after ? R : match_extend(v.w + c_al, v.w + s, R + 1, cap);
                    if (L > best) { best = L; bdist = s - c_al; if (L >= cap) { r = kWalkDone; break; } }
                } else {
                    const int L = m + nvis - 1;  // only members with fewer than R leading c's were reached
                    if (L > best) { best = L; bdist = s - (cur - (nvis - 1)); }
                }
            }  // else: the target does not start with (c,c,c) (a hash collision): no member matches 3 bytes
            if (nvis < t + 1) { r = nvis == count_dist ? kWalkDone : kWalkBudget; break; }
            diff -= t;  // chain -= t
            lim = diff < 0 ? -diff : 0;
            link_from = cur - t;  // the bottom of the run links onwards (its link is the one in the record)
        } else {
            const uint32_t x = rc.b ^ c.quad;
            if ((x & 0xFFFFFFu) == 0) {  // the hash triple matches
                int len = cap < kMinMatch ? cap : kMinMatch;
                if (x == 0 && cap > kMinMatch) {
                    // four bytes match: the payload decides (zlib's scan_end test is only a shortcut for len <= best)
                    const uint8_t *a = v.w + cur, *b = v.w + s;
                    len = (best > kMinMatch && load_byte(a + best) != load_byte(b + best)) ? 0 : match_extend(a, b, kMinMatch + 1, cap);
                }
                if (len > best) {
                    best = len; bdist = s - cur;
                    if (len >= cap) { r = kWalkDone; break; }
                }
            }
        }
        // follow the link, charge the budget and the step cap: one exit test on the common path (precedence as in
        // zlib's loop condition: chain end / distance limit, then the budget; the step cap is ours and comes last)
        cur = link_from - dl;
        const bool done = (dl == 0) | (s - cur >= kMaxDist);  // cur <= limit
        steps--;
        if (done | (steps == lim)) { r = done ? kWalkDone : (steps + diff == 0 ? kWalkBudget : kWalkSteps); break; }
    }
    c.cur = cur; c.best = best; c.bdist = bdist; c.chain = steps + diff; c.steps_left = steps;
    return r;
}

CCT_HD int chain_walk(const LzView &v, int s, int cap, int &best, int &bdist, int chain, int max_steps) {
    ChainState c;
    chain_begin(v, s, cap, best, chain, max_steps, c);
    const int r = chain_run(v, c);
    best = c.best; bdist = c.bdist;
    return r;
}

// Cold-start probe (best_len = 2) over at most K loop steps.  longest_match's answer for any prev_length is
// "the first-visited candidate attaining the maximum length, if that is longer" (SURVEY B.4); prev_length only
// changes the chain budget (4096, quartered to 1024 once prev_length >= good_length 32), so the probe records
// one entry per budget: *lo for prev_length < 32, *hi for prev_length >= 32.  They differ only when the chain
// holds more than 1024 candidates.  An entry is kProbeIncomplete when the step cap ended the walk first.
CCT_HD void match_probe(const LzView &v, int s, int K, uint32_t *lo, uint32_t *hi) {
    *lo = *hi = 0;
    const int look = v.n - s;
    if (look < kMinMatch) return;
    const int cap = look < kMaxMatch ? look : kMaxMatch;
    ChainState c;
    const int d0 = chain_begin(v, s, cap, 2, kChainGood, K, c);
    if (d0 == 0 || d0 > kMaxDist) return;
    int r = chain_run(v, c);
    const uint32_t e = r == kWalkSteps ? kProbeIncomplete : (c.best >= kMinMatch ? mtab_pack(c.best, c.bdist) : 0u);
    *lo = *hi = e;
    if (r != kWalkBudget) return;
    // The quartered budget ran out: the walk with the full budget visits the same candidates up to here, so it
    // simply goes on.  (A run group cut short by the budget left cur and chain untouched and is evaluated again with
    // the longer count -- its members seen so far are a prefix, so best/bdist come out as from a fresh walk.)
    if (c.steps_left == 0) { *lo = kProbeIncomplete; return; }  // budget and step cap ran out on the same candidate
    c.chain += kChainFull - kChainGood;
    r = chain_run(v, c);
    *lo = r == kWalkSteps ? kProbeIncomplete : (c.best >= kMinMatch ? mtab_pack(c.best, c.bdist) : 0u);
}

// Exact longest_match for position s with the parser's prev_length (chain budget 4096, or 1024 once
// prev_length >= good_length 32).  Requires: valid head, prev_len < 258, prev_len < look.
// Returns the improved length (> prev_len) and its distance, or 0.
CCT_HD int match_full(const LzView &v, int s, int prev_len, int *out_dist) {
    const int look = v.n - s;
    const int cap = look < kMaxMatch ? look : kMaxMatch;
    int best = prev_len, bdist = 0;
    CCT_COUNT(1);
    chain_walk(v, s, cap, best, bdist, prev_len >= kGoodLen ? kChainGood : kChainFull, 0x7FFFFFFF);
    *out_dist = bdist;
    return best > prev_len ? best : 0;
}

// ------------------------------------------------------------------------------------------------
// deflate_slow as a walk over symbol starts
// ------------------------------------------------------------------------------------------------
// deflate_slow (deflate.c) emits the literal for byte s-1 while standing at s; re-indexed by the position a
// symbol STARTS at, its lazy state machine reads:
//     at symbol start p, with either no pending match ("cold") or the match (pl, pd) found at p by the
//     previous step's look-ahead:
//         cur = pending, or the cold search at p (prev_length 2, TOO_FAR rule)
//         no cur                      -> LITERAL(w[p]);  next start p+1, cold
//         search at p+1 beats cur     -> LITERAL(w[p]);  next start p+1, pending = that longer match
//         otherwise                   -> MATCH(cur);     next start p+len, cold
// A search result comes from the probe table when the probe was complete (then it does not depend on
// prev_length beyond "must be longer", so a pending match at p is simply the table entry of p and the walk
// state is a function of p alone) and from match_full otherwise.  Because the state is (almost always) a
// function of p, walks started at different positions coincide from the first symbol start they share: the
// parse is done speculatively per window in parallel (walk_window), then stitched by one sequential pass per
// stream that only re-walks up to the point where the true path joins the speculative one (walk_stitch).
struct Walk {
    int p, pl, pd;  // symbol start, pending match length (0 = cold) and distance
};

struct ProbeTab {
    const uint32_t *lo, *hi;  // entries for chain budget 4096 (prev_length < 32) and 1024 (prev_length >= 32)
};

CCT_HD bool lookup_match(const LzView &v, const ProbeTab &mt, int p, int prev_len, int *len, int *dist) {
    const int look = v.n - p;
    if (look < kMinMatch) return false;
    const uint32_t e = prev_len >= kGoodLen ? mt.hi[p] : mt.lo[p];
    if (e == 0 || prev_len >= kMaxMatch || prev_len >= look) return false;
    int ml = 0, md = 0;
    if (e & kProbeIncomplete) ml = match_full(v, p, prev_len, &md);
    else if (mtab_len(e) > prev_len) { ml = mtab_len(e); md = mtab_dist(e); }
    if (ml == kMinMatch && md > kTooFar) ml = 0;
    *len = ml; *dist = md;
    return ml != 0;
}

// One symbol: returns it and advances the walk.
CCT_HD uint32_t walk_step(const LzView &v, const ProbeTab &mtab, Walk &w) {
    const int p = w.p;
    int l0 = w.pl, d0 = w.pd;
    if (!l0 && !lookup_match(v, mtab, p, 2, &l0, &d0)) { w.p = p + 1; return sym_pack(v.w[p], 0); }
    int l1, d1;
    if (lookup_match(v, mtab, p + 1, l0, &l1, &d1)) { w.p = p + 1; w.pl = l1; w.pd = d1; return sym_pack(v.w[p], 0); }
    w.p = p + l0; w.pl = 0; w.pd = 0;
    return sym_pack(l0 - kMinMatch, d0);
}

// Do two walks standing on the same symbol start continue identically?
CCT_HD bool walk_same_state(const ProbeTab &mt, const Walk &a, const Walk &b) {
    if (a.pl == b.pl && a.pd == b.pd) return true;
    const uint32_t e = mt.lo[a.p];  // one complete entry for both budgets: the state is a function of the position
    return e == mt.hi[a.p] && !(e & kProbeIncomplete);
}

// The step as a function of the position alone.  From a cold start at p the machine needs lookup(p, 2) and, if that
// finds a match of length l0, lookup(p + 1, l0); when both come straight from complete table entries the symbol and
// the next start are known without walking -- and a walk that arrives at p+1 with the pending match this step found
// continues exactly like a cold one there, provided the pending entry is the one a cold lookup reads (always the case
// below good_length; above it the two budget tables must agree).  walk_precompute evaluates this for every position
// in parallel (k_next): sym[p] = the symbol a walk standing at p emits (the advance follows from it), or kSymSlow
// when the step needs match_full or depends on how the walk got there.  The walks themselves (walk_window /
// walk_stitch) then chase one 4-byte entry per symbol and fall back to walk_step only on kSymSlow entries, whose
// symbol they record in place with the marker kept.
constexpr uint32_t kSymSlow = 0x80000000u;
CCT_HD int sym_advance(uint32_t sym) { return sym_dist(sym & ~kSymSlow) ? sym_lc(sym) + kMinMatch : 1; }

CCT_HD uint32_t walk_precompute(const LzView &v, const ProbeTab &mt, int p) {
    const uint32_t lit = sym_pack(v.w[p], 0);
    const int look = v.n - p;
    if (look < kMinMatch) return lit;
    const uint32_t e = mt.lo[p];  // lookup(p, 2)
    if (e & kProbeIncomplete) return kSymSlow;
    if (e == 0) return lit;
    const int l0 = mtab_len(e), d0 = mtab_dist(e);
    if (l0 == kMinMatch && d0 > kTooFar) return lit;
    const uint32_t match = sym_pack(l0 - kMinMatch, d0);
    const int look1 = look - 1;  // lookup(p + 1, l0)
    if (look1 < kMinMatch) return match;
    const uint32_t e1 = l0 >= kGoodLen ? mt.hi[p + 1] : mt.lo[p + 1];
    if (e1 == 0 || l0 >= kMaxMatch || l0 >= look1) return match;
    if (e1 & kProbeIncomplete) return kSymSlow;
    if (mtab_len(e1) <= l0) return match;  // (a longer match is at least 4 long: the TOO_FAR rule cannot apply)
    if (l0 >= kGoodLen && mt.lo[p + 1] != e1) return kSymSlow;  // the pending match is not what a cold lookup reads
    return lit;
}

// One symbol of a walk over the precomputed table.  `record`: store the symbol of a slow step (the true walk and the
// speculative window walks do; the stitch's re-run of a speculative walk does not).
CCT_HD void walk_step_tab(const LzView &v, const ProbeTab &mtab, uint32_t *sym, Walk &w, bool record) {
    const int q = w.p;
    if (w.pl == 0) {
        const uint32_t s = sym[q];
        if (!(s & kSymSlow)) { w.p = q + sym_advance(s); return; }
    }
    const uint32_t s = walk_step(v, mtab, w);
    if (record) sym[q] = s | kSymSlow;
}

// Symbol starts are kept as a bit mask, one 32-bit word per 32 positions.
CCT_HD void mark_set(uint32_t *mask, int p) { mask[p >> 5] |= 1u << (p & 31); }
CCT_HD void mark_clear(uint32_t *mask, int p) { mask[p >> 5] &= ~(1u << (p & 31)); }

// Speculative walk of window [ws, we) (ws a multiple of 32): assumes a cold symbol start at ws.  Writes the mask words
// of the window (bits of the symbol starts it visits, zeros elsewhere) and returns the state it leaves the window with.
CCT_HD Walk walk_window(const LzView &v, const ProbeTab &mtab, int ws, int we, uint32_t *mask, uint32_t *sym) {
    Walk w{ws, 0, 0};
    int word = ws >> 5;
    uint32_t bits = 0;
    while (w.p < we) {
        const int q = w.p;
        walk_step_tab(v, mtab, sym, w, true);
        while (word < (q >> 5)) { mask[word++] = bits; bits = 0; }
        bits |= 1u << (q & 31);
    }
    const int last = (we + 31) >> 5;
    while (word < last) { mask[word++] = bits; bits = 0; }
    return w;
}

// Sequential stitch of one stream: window after window, advance the true walk T and (re-running it) the
// speculative walk S of the window in position order; S's marks that T does not confirm are cleared, T's
// symbols are recorded, and as soon as both stand on the same symbol start in the same state the rest of the
// window's speculation is accepted wholesale (T jumps to the recorded exit state).
CCT_HDN inline void walk_stitch(const LzView &v, const ProbeTab &mtab, int win, uint32_t *mask, uint32_t *sym,
                                const Walk *exits) {
    Walk T{0, 0, 0};
    const int n = v.n;
    for (int k = 0, ws = 0; ws < n; k++, ws += win) {
        const int we = ws + win < n ? ws + win : n;
        if (T.p == ws && T.pl == 0) { T = exits[k]; continue; }  // speculation started from the true state
        Walk S{ws, 0, 0};
        for (;;) {
            const bool t_in = T.p < we, s_in = S.p < we;
            if (!t_in && !s_in) break;
            if (t_in && s_in && T.p == S.p) {
                if (walk_same_state(mtab, T, S)) { T = exits[k]; break; }
                // same start, different pending match over an incomplete probe: both move on
                walk_step_tab(v, mtab, sym, S, false);  // (first: T's step may rewrite the entry S must still see as slow)
                walk_step_tab(v, mtab, sym, T, true);
            } else if (s_in && (!t_in || S.p < T.p)) {
                mark_clear(mask, S.p);  // not on the true path
                walk_step_tab(v, mtab, sym, S, false);
            } else {
                mark_set(mask, T.p);
                walk_step_tab(v, mtab, sym, T, true);
            }
        }
    }
}

// Block boundaries.  zlib flushes a block as soon as 16383 symbols have been tallied inside its
// loop; the literal for the last byte is tallied after the loop and never triggers a flush.  With symbols
// numbered in stream order, block j < nflush ends where symbol 16383 (j+1) starts (recorded by the compaction
// pass), or at n when the input ends exactly there (then zlib appends an empty last block).
CCT_HD int finish_blocks(int nsym, uint32_t last_sym, int n, int32_t *blk_end) {
    const int lastlit = (nsym > 0 && sym_dist(last_sym) == 0) ? 1 : 0;
    const int nf = (nsym - lastlit) / kSymEnd;
    if (nf > 0 && nf * kSymEnd == nsym) blk_end[nf - 1] = n;
    return nf;
}

// ------------------------------------------------------------------------------------------------
// trees.c: build_tree / gen_bitlen / gen_codes / scan_tree / send_tree / block-type choice
// ------------------------------------------------------------------------------------------------
constexpr int kLCodes = 286, kDCodes = 30, kBLCodes = 19, kHeapSize = 2 * kLCodes + 1, kMaxBits = 15;
constexpr int kHdrWords = 160;  // dynamic-block header bit string capacity (worst case 4498 bits)

struct Tree {  // one Huffman tree under construction (ct_data with Freq/Dad/Len/Code split out)
    alignas(8) uint16_t freq[kHeapSize];  // freq + dad: dead once the tree is built; k_block_build assembles
    uint16_t dad[kHeapSize];              // the block descriptor in their place
    uint16_t len[kHeapSize];
    uint16_t code[kLCodes + 2];           // leaves only
    int max_code;
};

struct SmallTree {  // distance (30 symbols) and bit-length (19 symbols) trees
    uint16_t freq[2 * kDCodes + 1];
    uint16_t dad[2 * kDCodes + 1];
    uint16_t len[2 * kDCodes + 1];
    uint16_t code[kDCodes + 2];
    int max_code;
};

struct TreeScratch {
    // The priority queue of build_tree, one word per entry: Freq << 16 | depth << 10 | node.  trees.c's
    // smaller(n, m) = Freq[n] < Freq[m] || (equal && depth[n] <= depth[m]) is exactly (q[n] >> 10) <= (q[m] >> 10);
    // with the node in the same word a sift level is one 8-byte load of both children instead of a chain of
    // heap[] -> Freq[] -> depth[] loads.  (Freq <= 16384 + 2 symbols per block, depth <= 22 < 64, node < 573.)
    alignas(8) uint32_t pq[kHeapSize + 3];
    int16_t heap[kHeapSize];  // nodes in the order they left the queue (zlib keeps them in the top of its heap array)
    uint16_t bl_count[kMaxBits + 1];
    int heap_len, heap_max;
    int64_t opt_len, static_len;
};

#if !defined(__CUDACC__)
struct uint2 { uint32_t x, y; };  // host build of the test harness
#endif
constexpr uint32_t kPqNone = 0xFFFFFFFFu;  // sentinel after the last entry: never "smaller" than a real one

CCT_HDN inline void pq_down(TreeScratch &s, int k, const int heap_len) {  // (heap_len by value: s.pq stores would force a reload)
    const uint32_t v = s.pq[k], vk = v >> 10;
    int j = k << 1;
    while (j <= heap_len) {
        // children j (even: 8-byte aligned) and j + 1 (the sentinel when j is the last entry)
        const uint2 ch = *reinterpret_cast<const uint2 *>(&s.pq[j]);
        uint32_t c = ch.x;
        if ((ch.y >> 10) <= (c >> 10)) { j++; c = ch.y; }
        if (vk <= (c >> 10)) break;
        s.pq[k] = c;
        k = j; j <<= 1;
    }
    s.pq[k] = v;
}

// kind: 0 literal/length tree, 1 distance tree, 2 bit-length tree
template <class T>
CCT_HDN inline void build_tree(TreeScratch &s, T &t, int kind) {
    const int elems = kind == 0 ? kLCodes : (kind == 1 ? kDCodes : kBLCodes);
    const int base = kind == 0 ? 257 : 0;
    const int max_length = kind == 2 ? 7 : kMaxBits;
    const int heap_size = 2 * elems + 1;  // zlib uses one HEAP_SIZE array; only relative order matters
    int n, m, max_code = -1, node;
    // the queue length, the stack cursor and the two cost sums live in registers (the struct sits in shared memory)
    int heap_len = 0, heap_max = heap_size;
    int64_t opt_len = s.opt_len, static_len = s.static_len;
    for (n = 0; n < elems; n++) {
        if (t.freq[n] != 0) { s.pq[++heap_len] = ((uint32_t)t.freq[n] << 16) | (uint32_t)(max_code = n); }
        else t.len[n] = 0;
    }
    while (heap_len < 2) {  // force at least two codes
        node = max_code < 2 ? ++max_code : 0;
        s.pq[++heap_len] = (1u << 16) | (uint32_t)node;
        t.freq[node] = 1;
        opt_len--;
        if (kind == 0) static_len -= static_llen(node);
        else if (kind == 1) static_len -= 5;
    }
    t.max_code = max_code;
    s.pq[heap_len + 1] = kPqNone;
    for (n = heap_len / 2; n >= 1; n--) pq_down(s, n, heap_len);
    node = elems;
    do {
        const uint32_t nq = s.pq[1];  // pqremove
        s.pq[1] = s.pq[heap_len]; s.pq[heap_len] = kPqNone; heap_len--;
        pq_down(s, 1, heap_len);
        const uint32_t mq = s.pq[1];
        n = (int)(nq & 1023u); m = (int)(mq & 1023u);
        s.heap[--heap_max] = (int16_t)n; s.heap[--heap_max] = (int16_t)m;
        const uint32_t f = (nq >> 16) + (mq >> 16);
        const uint32_t dn = (nq >> 10) & 63u, dm = (mq >> 10) & 63u;
        t.freq[node] = (uint16_t)f;
        t.dad[n] = t.dad[m] = (uint16_t)node;
        s.pq[1] = (f << 16) | ((((dn >= dm ? dn : dm) + 1u) & 63u) << 10) | (uint32_t)node;
        node++;
        pq_down(s, 1, heap_len);
    } while (heap_len >= 2);
    s.heap[--heap_max] = (int16_t)(s.pq[1] & 1023u);

    // gen_bitlen
    int h, bits, overflow = 0;
    for (bits = 0; bits <= kMaxBits; bits++) s.bl_count[bits] = 0;
    t.len[s.heap[heap_max]] = 0;
    for (h = heap_max + 1; h < heap_size; h++) {
        n = s.heap[h];
        bits = t.len[t.dad[n]] + 1;
        if (bits > max_length) { bits = max_length; overflow++; }
        t.len[n] = (uint16_t)bits;
        if (n > max_code) continue;  // internal node
        s.bl_count[bits]++;
        int xbits = 0;
        if (n >= base) xbits = kind == 0 ? len_extra_bits(n - base) : (kind == 1 ? dist_extra_bits(n) : bl_extra_bits(n));
        const int f = t.freq[n];
        opt_len += f * (bits + xbits);
        if (kind == 0) static_len += f * (static_llen(n) + xbits);
        else if (kind == 1) static_len += f * (5 + xbits);
    }
    if (overflow) {
        do {
            bits = max_length - 1;
            while (s.bl_count[bits] == 0) bits--;
            s.bl_count[bits]--; s.bl_count[bits + 1] += 2; s.bl_count[max_length]--;
            overflow -= 2;
        } while (overflow > 0);
        for (bits = max_length; bits != 0; bits--) {
            n = s.bl_count[bits];
            while (n != 0) {
                m = s.heap[--h];
                if (m > max_code) continue;
                if (t.len[m] != (unsigned)bits) {
                    opt_len += ((int64_t)bits - t.len[m]) * t.freq[m];
                    t.len[m] = (uint16_t)bits;
                }
                n--;
            }
        }
    }
    s.heap_len = heap_len; s.heap_max = heap_max; s.opt_len = opt_len; s.static_len = static_len;
    // gen_codes
    uint32_t next_code[kMaxBits + 1], code = 0;
    next_code[0] = 0;
    for (bits = 1; bits <= kMaxBits; bits++) { code = (code + s.bl_count[bits - 1]) << 1; next_code[bits] = code & 0xFFFFu; }
    for (n = 0; n <= max_code; n++) {
        int len = t.len[n];
        if (len == 0) continue;
        t.code[n] = (uint16_t)bit_reverse(next_code[len]++, len);
    }
}

// tiny LSB-first bit writer into a uint32 word array (dynamic-block header string)
struct HdrBits {
    uint32_t *w; int nbits;
    CCT_HD void put(uint32_t value, int length) {
        int word = nbits >> 5, off = nbits & 31;
        w[word] |= value << off;
        if (off + length > 32) w[word + 1] |= value >> (32 - off);
        nbits += length;
    }
};

// scan_tree (count into blfreq) when bl == nullptr, send_tree (emit) otherwise — one control flow
template <class T>
CCT_HDN inline void tree_rle(const T &t, int max_code, uint16_t *blfreq, const SmallTree *bl, HdrBits *out) {
    int prevlen = -1, curlen, nextlen = t.len[0], count = 0, max_count = 7, min_count = 4;
    if (nextlen == 0) { max_count = 138; min_count = 3; }
    for (int n = 0; n <= max_code; n++) {
        curlen = nextlen;
        nextlen = (n + 1 <= max_code) ? t.len[n + 1] : 0xFFFF;
        if (++count < max_count && curlen == nextlen) continue;
        else if (count < min_count) {
            if (blfreq) blfreq[curlen] += (uint16_t)count;
            else do { out->put(bl->code[curlen], bl->len[curlen]); } while (--count != 0);
        } else if (curlen != 0) {
            if (curlen != prevlen) {
                if (blfreq) blfreq[curlen]++;
                else { out->put(bl->code[curlen], bl->len[curlen]); count--; }
            }
            if (blfreq) blfreq[16]++;
            else { out->put(bl->code[16], bl->len[16]); out->put((uint32_t)(count - 3), 2); }
        } else if (count <= 10) {
            if (blfreq) blfreq[17]++;
            else { out->put(bl->code[17], bl->len[17]); out->put((uint32_t)(count - 3), 3); }
        } else {
            if (blfreq) blfreq[18]++;
            else { out->put(bl->code[18], bl->len[18]); out->put((uint32_t)(count - 11), 7); }
        }
        count = 0; prevlen = curlen;
        if (nextlen == 0) { max_count = 138; min_count = 3; }
        else if (curlen == nextlen) { max_count = 6; min_count = 3; }
        else { max_count = 7; min_count = 4; }
    }
}

// Result of _tr_flush_block's decisions for one block; consumed by the bit-packing kernel.
struct BlockDesc {
    uint32_t sym_begin, nsym;        // symbol range inside the stream's symbol array
    uint32_t pos_begin, stored_len;  // payload bytes covered (for stored blocks)
    uint32_t type, last;             // 0 stored, 1 static, 2 dynamic
    uint32_t hdr_bits;               // header bit count incl. the 3 block-header bits (dynamic: + trees)
    uint32_t body_bits;              // symbol bits + EOB (0 for stored)
    uint64_t bit_start;              // absolute bit offset in the zlib stream (filled by the layout pass)
    uint16_t lcode[kLCodes + 2]; uint8_t llen[kLCodes + 2];
    uint16_t dcode[kDCodes + 2]; uint8_t dlen[kDCodes + 2];
    uint32_t hdr[kHdrWords];
};

struct BlockWork {  // scratch for build_block (shared or local memory)
    Tree lt; SmallTree dt, bt; TreeScratch sc;
};

// lt.freq / dt.freq must hold the block's symbol frequencies (incl. freq[256] = 1) on entry, every
// other field zero.  Fills desc->type/last/stored_len/hdr/hdr_bits/body_bits and the code tables.  `desc` is
// written only after the three trees are built, so it may occupy the storage of wk.lt.freq / wk.lt.dad.
CCT_HDN inline void build_block(BlockWork &wk, BlockDesc *desc, uint32_t stored_len, uint32_t last) {
    const uint8_t bl_order[kBLCodes] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    wk.sc.opt_len = 0; wk.sc.static_len = 0;
    build_tree(wk.sc, wk.lt, 0);
    build_tree(wk.sc, wk.dt, 1);
    tree_rle(wk.lt, wk.lt.max_code, wk.bt.freq, (const SmallTree *)nullptr, (HdrBits *)nullptr);
    tree_rle(wk.dt, wk.dt.max_code, wk.bt.freq, (const SmallTree *)nullptr, (HdrBits *)nullptr);
    build_tree(wk.sc, wk.bt, 2);
    int max_blindex;
    for (max_blindex = kBLCodes - 1; max_blindex >= 3; max_blindex--)
        if (wk.bt.len[bl_order[max_blindex]] != 0) break;
    wk.sc.opt_len += 3 * ((int64_t)max_blindex + 1) + 5 + 5 + 4;
    int64_t opt_lenb = (wk.sc.opt_len + 3 + 7) >> 3, static_lenb = (wk.sc.static_len + 3 + 7) >> 3;
    if (static_lenb <= opt_lenb) opt_lenb = static_lenb;
    desc->stored_len = stored_len; desc->last = last;
    for (int i = 0; i < kHdrWords; i++) desc->hdr[i] = 0;
    HdrBits hb{desc->hdr, 0};
    if ((int64_t)desc->stored_len + 4 <= opt_lenb) {
        desc->type = 0;
        hb.put((0u << 1) + desc->last, 3);
        desc->body_bits = 0;
    } else if (static_lenb == opt_lenb) {
        desc->type = 1;
        hb.put((1u << 1) + desc->last, 3);
        desc->body_bits = (uint32_t)wk.sc.static_len;
    } else {
        desc->type = 2;
        hb.put((2u << 1) + desc->last, 3);
        hb.put((uint32_t)(wk.lt.max_code + 1 - 257), 5);
        hb.put((uint32_t)(wk.dt.max_code + 1 - 1), 5);
        hb.put((uint32_t)(max_blindex + 1 - 4), 4);
        for (int r = 0; r <= max_blindex; r++) hb.put(wk.bt.len[bl_order[r]], 3);
        tree_rle(wk.lt, wk.lt.max_code, (uint16_t *)nullptr, &wk.bt, &hb);
        tree_rle(wk.dt, wk.dt.max_code, (uint16_t *)nullptr, &wk.bt, &hb);
        desc->body_bits = (uint32_t)(wk.sc.opt_len - (hb.nbits - 3));
    }
    desc->hdr_bits = (uint32_t)hb.nbits;
}

// The code tables the bit packer reads, after build_block: (lane, nl) = (lane id, 32) on the device, (0, 1) on the host.
CCT_HD void fill_block_tables(const BlockWork &wk, BlockDesc *desc, int lane, int nl) {
    if (desc->type == 1) {
        for (int n = lane; n < kLCodes + 2; n += nl) { desc->lcode[n] = (uint16_t)static_lcode(n); desc->llen[n] = (uint8_t)static_llen(n); }
        for (int n = lane; n < kDCodes; n += nl) { desc->dcode[n] = (uint16_t)bit_reverse((uint32_t)n, 5); desc->dlen[n] = 5; }
    } else if (desc->type == 2) {
        for (int n = lane; n < kLCodes; n += nl) { desc->lcode[n] = wk.lt.code[n]; desc->llen[n] = (uint8_t)(n <= wk.lt.max_code ? wk.lt.len[n] : 0); }
        for (int n = lane; n < kDCodes; n += nl) { desc->dcode[n] = wk.dt.code[n]; desc->dlen[n] = (uint8_t)(n <= wk.dt.max_code ? wk.dt.len[n] : 0); }
    }
}

// bits of one LZ77 symbol under a block's code tables: returns the bit string (LSB-first, <= 48 bits)
CCT_HD uint64_t sym_bits(uint32_t sym, const uint16_t *lcode, const uint8_t *llen, const uint16_t *dcode,
                         const uint8_t *dlen, int *nbits) {
    int dist = sym_dist(sym), lc = sym_lc(sym);
    if (dist == 0) { *nbits = llen[lc]; return lcode[lc]; }
    int code = len_code(lc);
    uint64_t bits = lcode[code + 257];
    int nb = llen[code + 257];
    int extra = len_extra_bits(code);
    if (extra) { bits |= (uint64_t)(lc & ((1 << extra) - 1)) << nb; nb += extra; }
    int d = dist - 1;
    code = dist_code(d);
    bits |= (uint64_t)dcode[code] << nb; nb += dlen[code];
    extra = dist_extra_bits(code);
    if (extra) { bits |= (uint64_t)(d & ((1 << extra) - 1)) << nb; nb += extra; }
    *nbits = nb;
    return bits;
}

CCT_HD int sym_nbits(uint32_t sym, const uint8_t *llen, const uint8_t *dlen) {
    int dist = sym_dist(sym), lc = sym_lc(sym);
    if (dist == 0) return llen[lc];
    int code = len_code(lc);
    int nb = llen[code + 257] + len_extra_bits(code);
    code = dist_code(dist - 1);
    return nb + dlen[code] + dist_extra_bits(code);
}

}  // namespace cct
